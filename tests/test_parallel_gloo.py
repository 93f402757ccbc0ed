"""world_size-2 gloo test (CPU) of the multi-GPU host logic: packet sharding + the single all-reduce of
the estimators + gathering the per-packet outputs.  Each rank's shard is computed by the CPU oracle
(the engine itself needs a GPU); the reduced result must equal one oracle run over all packets."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from oracle import cpu_oracle
    from tardis_b200 import parallel
    from tardis_b200 import synthetic as syn

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = syn.make_model(6, 1500, "macroatom", mu_tau=-4.0, seed=31)
    packets = syn.make_packets(1001, model.r_inner[0], base_seed=5)
    lo, hi = parallel.shard_bounds(len(packets), rank, world)
    local = cpu_oracle.run_oracle(model, packets.slice(lo, hi), number_of_vpackets=2)
    red = parallel.all_reduce_host_results(local, dist)
    nus, energies = parallel.gather_packet_outputs(local["output_nus"], local["output_energies"], len(packets), dist)
    if rank == 0:
        full = cpu_oracle.run_oracle(model, packets, number_of_vpackets=2)
        ok = True
        for k in parallel.ESTIMATOR_KEYS:
            ok &= np.allclose(red[k], full[k], rtol=1e-12, atol=0)
        ok &= np.array_equal(nus, full["output_nus"]) and np.array_equal(energies, full["output_energies"])
        ret.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


class _FakeEngine:
    """Host stand-in with the four calls parallel.all_reduce_estimators makes (the real engine needs a GPU)."""

    def __init__(self, rank, scales):
        import torch

        self.est = torch.arange(10, dtype=torch.float64) + 100.0 * rank
        self.words = torch.arange(8, dtype=torch.int64) * (rank + 1) - 3
        self.scales = scales
        self.finalized = 0

    def sync(self):
        pass

    def estimator_layout(self):
        return {"off_j_blue": 6}

    def finalize_line_estimators(self):
        self.finalized += 1
        self.est[6:] = self.words[:4].double()  # "difference arrays -> doubles"


def _collective_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from tardis_b200 import parallel

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    parallel.estimator_tensor = lambda e: e.est
    parallel.line_accumulator_tensor = lambda e: (e.words, e.scales)
    ok = True
    # same scales on both ranks: the integer words are summed, the doubles before off_j_blue are summed, the line part of
    # the double buffer is NOT summed but rebuilt by finalize
    e = _FakeEngine(rank, (2.0 ** 10, 2.0 ** 20))
    exact = parallel.all_reduce_estimators(e, dist)
    words = torch.arange(8, dtype=torch.int64) * 3 - 6
    ok &= exact and e.finalized == 1 and torch.equal(e.words, words)
    ok &= torch.equal(e.est[:6], 2 * torch.arange(6, dtype=torch.float64) + 100.0) and torch.equal(e.est[6:], words[:4].double())
    # scales differ between the ranks: one f64 all-reduce over the whole buffer, no finalize
    e = _FakeEngine(rank, (2.0 ** (10 + rank), 2.0 ** 20))
    exact = parallel.all_reduce_estimators(e, dist)
    ok &= (not exact) and e.finalized == 0 and torch.equal(e.est, 2 * torch.arange(10, dtype=torch.float64) + 100.0)
    ok &= torch.equal(e.words, torch.arange(8, dtype=torch.int64) * (rank + 1) - 3)
    # caller opts out
    e = _FakeEngine(rank, (2.0 ** 10, 2.0 ** 20))
    ok &= (not parallel.all_reduce_estimators(e, dist, exact_lines=False)) and e.finalized == 0
    if rank == 0:
        ret.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def _formal_integral_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from oracle import formal_integral_oracle as fio
    from tardis_b200 import parallel
    from tardis_b200 import synthetic as syn

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = syn.make_model(6, 1500, "downbranch", mu_tau=-3.0, seed=41)
    rng = np.random.default_rng(42)
    tau = np.asarray(model.tau_sobolev)
    att, jblue = rng.random(tau.shape) * 1e-6, rng.random(tau.shape) * 1e-5
    jred = jblue * np.exp(-tau) + att

    class OracleEngine:  # the engine needs a GPU; the host logic under test only needs its call
        def __init__(self):
            self.asked = None

        def formal_integral(self, *, frequencies, inner_temperature, points, interpolate_shells=0, **k):
            self.asked = np.array(frequencies)
            if len(frequencies) == 0:
                return dict(luminosity_densities=np.zeros(0), intensities_nu_p=None, interpolation_ms=0.0, integral_ms=0.0)
            o = fio.solve(model.r_inner, model.r_outer, float(model.time_explosion), model.line_list_nu, inner_temperature, frequencies, att, jred,
                          jblue, tau, model.electron_density, points, interpolate_shells)
            return dict(luminosity_densities=o["luminosity_densities"], intensities_nu_p=None, interpolation_ms=1.0, integral_ms=2.0 + rank)

    ok = True
    for n in (7, 1, 0):  # odd split, fewer frequencies than ranks, none
        freq = np.linspace(model.line_list_nu[-1] * 1.1, model.line_list_nu[0] * 0.95, max(n, 2))[:n]
        eng = OracleEngine()
        got = parallel.formal_integral_sharded(eng, dist, frequencies=freq, inner_temperature=1.0e4, points=40, interpolate_shells=0)
        lo, hi = parallel.shard_bounds(n, rank, world)
        ok &= got["frequency_range"] == (lo, hi) and np.array_equal(eng.asked, freq[lo:hi])
        if n:
            full = fio.solve(model.r_inner, model.r_outer, float(model.time_explosion), model.line_list_nu, 1.0e4, freq, att, jred, jblue, tau,
                             model.electron_density, 40, 0)["luminosity_densities"]
        else:
            full = np.zeros(0)
        ok &= np.array_equal(got["luminosity_densities"], full)  # the same arithmetic per frequency: bit-identical
    try:
        parallel.formal_integral_sharded(OracleEngine(), dist, frequencies=np.ones(3), want_intensities=True)
        ok = False
    except ValueError:
        pass
    ret.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_formal_integral_shards_over_frequencies(oracle):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_formal_integral_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    got = dict(ret.get(timeout=5) for _ in range(2))
    assert got == {0: True, 1: True}


def test_estimator_collective_takes_the_exact_integer_path_when_scales_agree():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_collective_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert ret.get(timeout=5) is True


def test_shard_bounds_cover_everything():
    from tardis_b200.parallel import shard_bounds

    for n in (0, 1, 7, 1000, 10**8 + 3):
        for world in (1, 2, 3, 8):
            edges = [shard_bounds(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def test_two_rank_gloo_reduction_matches_single_run(oracle):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert ret.get(timeout=5) is True
