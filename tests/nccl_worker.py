"""torchrun worker for test_two_gpu_nccl_allreduce: each rank runs its packet shard on its own GPU, the packed
estimator buffer is all-reduced over NCCL, rank 0 compares with one oracle run over all packets."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tardis_b200 import parallel  # noqa: E402
from tardis_b200 import synthetic as syn  # noqa: E402
from tardis_b200.engine import Engine  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    model = syn.make_model(8, 3000, "macroatom", mu_tau=-4.0, seed=61)
    packets = syn.make_packets(20001, model.r_inner[0], base_seed=9)
    lo, hi = parallel.shard_bounds(len(packets), rank, world)
    sh = packets.slice(lo, hi)
    for algo in (0, 1):
        eng = Engine(local)
        eng.set_option("algorithm", algo)
        eng.set_model_from(model, number_of_vpackets=2)
        eng.upload_packets(sh.initial_radii, sh.initial_nus, sh.initial_mus, sh.initial_energies, sh.packet_seeds)
        eng.transport(True)
        exact = parallel.all_reduce_estimators(eng, dist)
        assert exact == (algo == 1)
        res = eng.download()
        if algo == 1:
            # the exact path: J_blue / Edotlu of the 2-GPU run == those of ONE engine over all packets, bit for bit,
            # and identical on both ranks
            one = Engine(local)
            one.set_model_from(model, number_of_vpackets=2)
            full = one.run_packets(packets)
            one.close()
            for k in ("j_blue", "edotlu"):
                assert np.array_equal(res[k], full[k]), f"{k}: sharded + integer all-reduce differs from the single-engine run"
                t = torch.from_numpy(np.ascontiguousarray(res[k])).cuda()
                lo_t, hi_t = t.clone(), t.clone()
                dist.all_reduce(lo_t, op=dist.ReduceOp.MIN)
                dist.all_reduce(hi_t, op=dist.ReduceOp.MAX)
                assert torch.equal(lo_t, hi_t), f"{k} differs between ranks"
            for k in ("j", "nu_bar"):
                np.testing.assert_allclose(res[k], full[k], rtol=1e-12, atol=0)
        if rank == 0:
            from oracle import cpu_oracle

            ref = cpu_oracle.run_oracle(model, packets, number_of_vpackets=2, nthreads=8)
            for k in parallel.ESTIMATOR_KEYS:
                np.testing.assert_allclose(res[k], ref[k], rtol=1e-10, atol=0, err_msg=f"{k} algo {algo}")
            np.testing.assert_allclose(res["output_nus"], ref["output_nus"][lo:hi], rtol=1e-11)
        eng.close()
    dist.barrier()
    if rank == 0:
        print("NCCL_PARITY_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
