#!/usr/bin/env python
"""bench.py -- MC packets/sec of the packet-propagation path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA engine
    python bench.py --impl reference [--gpus N] ...                # CPU arm (oracle port, all host threads)
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N   # one rank per GPU

One "step" = one Monte Carlo iteration of the hot path (`montecarlo_transport_with_vpackets`,
tardis/transport/montecarlo/modes/montecarlo_transport.py:239) over one batch of synthetic packets:
workload = BASELINE.json configs[2]: 1e8 packets per GPU, 20 shells, 5e5-line synthetic list, macroatom
(SURVEY.md §8d generator, tardis_b200/synthetic.py).  Packets shard across ranks (weak scaling: every
rank gets its own `--packets`), the only collective is one all-reduce of the packed estimator buffer.

  value   : packets/s with the packets already resident in HBM (kernel + estimator all-reduce)
  e2e     : packets/s through the reference-facing call `tb200_run` with pinned HOST buffers:
            H2D of the 5 packet arrays, seed expansion, transport kernel, D2H of output_nus/energies
            and of all estimators ([L,S] layout), every step
  roofline: algorithmic bytes (SURVEY.md §8d formula, from the kernel's exact integer work counters)
            / CUDA-event duration of the transport kernel, against the measured HBM copy bandwidth
  cpu_baseline: oracle/tardis_oracle.c ("port" of the reference loop) on all host threads, bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tardis_b200 import synthetic as syn  # noqa: E402

METRIC = "MC packets/sec at 1e8 packets, 20 shells, 5e5 lines; spectrum L2 vs ref"


def alg_bytes(c: dict, n_packets: int) -> int:
    """48 B per line-step, 16 B per virtual-packet line-step, 32 B per event,
    8 B per scanned macro-atom transition, 24 B per macro-atom jump, 56 B per packet."""
    events = c["n_boundary_events"] + c["n_line_events"] + c["n_escat_events"]
    return (48 * c["n_line_steps"] + 16 * c["n_vpacket_line_steps"] + 32 * events
            + 8 * c["n_macro_scanned"] + 24 * c["n_macro_jumps"] + 80 * c.get("n_bf_estimator_updates", 0) + 56 * n_packets)


def make_packets_chunked(n: int, r_inner0: float, seed_base: int, chunk: int = 10_000_000) -> syn.Packets:
    """n packets as a concatenation of BlackBodySimpleSource batches (bounded host memory)."""
    parts = []
    done = 0
    it = 0
    while done < n:
        m = min(chunk, n - done)
        parts.append(syn.make_packets(m, r_inner0, base_seed=seed_base, iteration=it))
        done += m
        it += 1
    if len(parts) == 1:
        p = parts[0]
    else:
        p = syn.Packets(*(np.concatenate([getattr(q, f) for q in parts]) for f in
                          ("initial_radii", "initial_nus", "initial_mus", "initial_energies", "packet_seeds")),
                        parts[0].radiation_field_luminosity)
    p.initial_energies[:] = 1.0 / n
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.device_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p > 250.0] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(power))}


def read_peak() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def read_traffic(workload_key: str, n_packets: int):
    """Per-launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the transport kernel from the committed
    ncu capture of the same model, scaled linearly from the captured packet count to this launch's."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(workload_key)
        return None if t is None else t["bytes"] * n_packets / t["packets"]
    except Exception:
        return None


def emitted_spectrum(output_nus, output_energies, grid, time_of_simulation):
    """SpectrumSolver.montecarlo_emitted_luminosity: np.histogram of emitted packets
    (tardis/spectrum/base.py:151-159)."""
    m = output_energies >= 0
    h, _ = np.histogram(output_nus[m], weights=output_energies[m] / time_of_simulation, bins=grid)
    return h


def cpu_leg(model, args, n_threads: int, seed_base: int, vp: int, target_seconds: float = 12.0):
    """Time the CPU oracle (port of the reference loop) on a bounded sample of the workload."""
    from oracle import cpu_oracle

    cpu_oracle.build()
    # Calibrate: thread count x estimator layout (per-thread tables as in the reference, or one shared table with atomic
    # adds).  More threads are not always faster for this memory-latency-bound loop; the timed run uses the fastest.
    cores = n_threads
    candidates = sorted({(t, private) for t in (cores, cores // 2, cores // 4, 32, 16, 8) if 1 <= t <= cores
                         for private in (True, False) if not (private and t > 32)}, reverse=True)
    best = getattr(cpu_leg, "calibrated", {}).get((cores, vp))  # calibrate once per process, not once per step
    for t, private in ([] if best else candidates):
        n0 = max(1_000, 100 * t)
        calib = make_packets_chunked(n0, model.r_inner[0], seed_base + 1)
        t0 = time.perf_counter()
        cpu_oracle.run_oracle(model, calib, number_of_vpackets=vp, nthreads=t, track_last_interaction=False,
                              private_tables_max_threads=(t if private else 0))
        rate = n0 / max(time.perf_counter() - t0, 1e-3)
        if best is None or rate > best[0]:
            best = (rate, t, private)
    cpu_leg.calibrated = {**getattr(cpu_leg, "calibrated", {}), (cores, vp): best}
    rate0, n_threads, private = best
    n = int(min(max(rate0 * target_seconds, 2_000), 4_000_000))
    sample = make_packets_chunked(n, model.r_inner[0], seed_base)
    t0 = time.perf_counter()
    res = cpu_oracle.run_oracle(model, sample, number_of_vpackets=vp, nthreads=n_threads, track_last_interaction=False,
                                private_tables_max_threads=(n_threads if private else 0))
    dt = time.perf_counter() - t0
    cpu_leg.last_choice = {"threads": n_threads, "layout": "per-thread tables" if private else "shared table + atomic adds",
                           "host_cores": cores}
    return n / dt, n, dt, sample, res, n_threads


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--packets", type=int, default=100_000_000, help="packets per GPU per step")
    ap.add_argument("--lines", type=int, default=500_000)
    ap.add_argument("--shells", type=int, default=20)
    ap.add_argument("--mode", default="macroatom", choices=["scatter", "downbranch", "macroatom"])
    ap.add_argument("--vpackets", type=int, default=0)
    ap.add_argument("--continuum", action="store_true", help="IIP mode (BASELINE config 5): bound-free / free-free continuum")
    ap.add_argument("--mu-tau", type=float, default=-7.5)
    ap.add_argument("--algorithm", default="jump", choices=["jump", "scan"],
                    help="jump: prefix-table search + range updates (default, fastest); scan: stream the line list")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scan-reference", action="store_true", help="skip the short scan-kernel roofline measurement")
    ap.add_argument("--device-source", action="store_true",
                    help="also time the step with the packets generated on the device (tb200_create_packets: no H2D of packets); "
                         "reported as e2e.device_source, the contract's e2e is unchanged")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        args.gpus = world

    workload = (f"{args.packets:.0e} packets/GPU, {args.shells} shells, {args.lines} lines, {args.mode}"
                + (f", {args.vpackets} vpackets" if args.vpackets else "") + (", continuum (IIP mode)" if args.continuum else "")
                + f", tau~10^N({args.mu_tau},2)")
    config = {"workload": workload, "packets_per_gpu": args.packets, "n_shells": args.shells, "n_lines": args.lines,
              "line_interaction_type": args.mode, "number_of_vpackets": args.vpackets, "continuum": bool(args.continuum),
              "algorithm": args.algorithm,
              "parallelism": f"packet-sharded x{args.gpus}",
              "l2_policy": "inputs larger than L2 (tables 80-400 MB + 5.6 GB of packets per step)"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        model = syn.make_model(args.shells, args.lines, args.mode, mu_tau=args.mu_tau)
        if args.continuum:
            syn.add_continuum(model)
        n_threads = os.cpu_count() or 1
        rates = []
        sample_n = 0
        for i in range(args.warmup + args.steps):
            rate, n, dt, _, _, used = cpu_leg(model, args, n_threads, syn.BASE_SEED + i, args.vpackets, target_seconds=8.0)
            sample_n = n
            if i >= args.warmup:
                rates.append((rate, dt))
        value = float(np.mean([r for r, _ in rates]))
        line = {"metric": METRIC, "value": value, "unit": "packets/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": float(np.mean([d for _, d in rates]) * 1e3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "impl": "reference",
                "cpu_baseline": {"value": value, "unit": "packets/s", "cores": used, "kind": "port",
                                 "sample": f"{sample_n} packets of the same workload per step (oracle/tardis_oracle.c; fastest of "
                                           f"the calibrated thread counts / table layouts on {n_threads} host cores: "
                                           f"{cpu_leg.last_choice}; the reference's Numba loop cannot travel to this box)"},
                "e2e": {"value": value, "unit": "packets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch

    from tardis_b200.engine import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))

    model = syn.make_model(args.shells, args.lines, args.mode, mu_tau=args.mu_tau)
    if args.continuum:
        syn.add_continuum(model)
    eng = Engine(local_rank)
    eng.set_option("algorithm", {"scan": 0, "jump": 1}[args.algorithm])
    # CTAs per SM, park / refill thresholds: the engine's own defaults (measured best per kernel, engine.cu launch_range)
    eng.set_model_from(model, number_of_vpackets=args.vpackets)

    # host packets of this rank's shard, in pinned memory
    n = args.packets
    pk = make_packets_chunked(n, model.r_inner[0], syn.BASE_SEED + 1000 * rank)

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64 if a.dtype == np.float64 else torch.int64, pin_memory=True)
        v = t.numpy()
        v[...] = a
        return t, v

    pins = [pinned(getattr(pk, f)) for f in ("initial_radii", "initial_nus", "initial_mus", "initial_energies", "packet_seeds")]
    host_in = [v for _, v in pins]
    h2d_bytes = int(sum(v.nbytes for v in host_in))

    from tardis_b200 import parallel

    est_tensor = parallel.estimator_tensor(eng)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step():
        eng.transport(True)
        eng.sync()
        if dist is not None:
            dist.all_reduce(est_tensor)  # the one collective of an MC iteration
            torch.cuda.synchronize()

    # ---- device-resident measurement ----
    eng.upload_packets(*host_in)
    for _ in range(args.warmup):
        resident_step()
    sampler = ClockSampler(local_rank)
    launches0 = eng.kernel_launches()
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    kernel_ms = []
    for _ in range(args.steps):
        resident_step()
        kernel_ms.append(eng.last_kernel_ms())
    barrier()
    elapsed = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = eng.kernel_launches() - launches0
    counters = eng.counters()
    if dist is not None:
        t = torch.tensor([elapsed], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    value = world * n * args.steps / elapsed

    # ---- end-to-end through the reference-facing call with host buffers ----
    L, S, G = model.n_lines, model.n_shells, len(model.spectrum_frequency_grid)
    d2h_bytes = int(2 * n * 8 + (2 * S + 2 * L * S + G) * 8)
    if args.continuum:
        d2h_bytes += int((5 * len(model.continuum.bf_threshold_list_nu) * S + S) * 8)
    e2e_steps = max(1, min(args.steps, 2))
    out_pins = {k: torch.empty(shape, dtype=torch.float64, pin_memory=True) for k, shape in eng.output_shapes(n).items()}
    host_out = {k: t.numpy() for k, t in out_pins.items()}
    res = eng.run(*host_in, buffers=host_out)  # warm-up (also sizes the staging buffers)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = eng.run(*host_in, buffers=host_out)
        if dist is not None:
            dist.all_reduce(est_tensor)
            torch.cuda.synchronize()
    barrier()
    e2e_elapsed = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_elapsed], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_elapsed = float(t.item())
    e2e_value = world * n * e2e_steps / e2e_elapsed

    # same call, but the caller only wants the estimators and the fused spectrum histograms (no per-packet D2H)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.run(*host_in, per_packet=False, buffers={k: v for k, v in host_out.items() if not k.startswith("output_")})
        if dist is not None:
            dist.all_reduce(est_tensor)
            torch.cuda.synchronize()
    barrier()
    lean_elapsed = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([lean_elapsed], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        lean_elapsed = float(t.item())
    e2e_lean_value = world * n * e2e_steps / lean_elapsed

    # optional: the same step with the device-side packet source.  Inputs per step: a seed.
    device_source = None
    if args.device_source and not args.continuum:
        t_inner = 1.0e4
        eng.create_packets(n, syn.BASE_SEED + 1000 * rank, float(model.r_inner[0]), t_inner)  # warm-up
        eng.transport(True); eng.sync()
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            eng.create_packets(n, syn.BASE_SEED + 1000 * rank + i + 1, float(model.r_inner[0]), t_inner)
            eng.transport(True)
            eng.sync()
            if dist is not None:
                dist.all_reduce(est_tensor)
                torch.cuda.synchronize()
            eng.download(buffers=host_out)
        barrier()
        ds_elapsed = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([ds_elapsed], dtype=torch.float64, device=f"cuda:{local_rank}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ds_elapsed = float(t.item())
        device_source = {"value": world * n * e2e_steps / ds_elapsed, "unit": "packets/s", "h2d_bytes_per_step": 64,
                         "d2h_bytes_per_step": d2h_bytes,
                         "note": "packets generated in HBM by tb200_create_packets (BlackBodySimpleSource on the device, T = 1e4 K); "
                                 "not pipelined: generation, transport and the D2H of the results run back to back"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    peak, peak_src = read_peak()
    k_ms = float(np.mean(kernel_ms))
    events = counters["n_boundary_events"] + counters["n_line_events"] + counters["n_escat_events"]
    ref_equiv = alg_bytes(counters, n)  # what the reference's loop touches for the same packets
    if args.algorithm == "scan":
        ab = ref_equiv
        note = "streaming kernel: SURVEY.md §8(d) bytes (48 B per line-step ...)"
    else:
        # the jump algorithm never streams the line list.  Its own algorithmic traffic per unit:
        #   40 B per stopping-predicate probe (nu_line 8 B + two 16 B double-double prefix entries; the one-entry
        #   verification of the common trace end counts as one probe: 2 x 8 B nu_line + 16 B prefix + 16 B prefix[start]),
        #   128 B per trace for the two fixed-point range updates (2 endpoints x 32 B read-modify-write),
        #   32 B per event for J / nu_bar, macro-atom and virtual-packet terms as in §8(d), 56 B per packet.
        ab = (40 * counters["n_search_probes"] + 128 * events + 32 * events + 8 * counters["n_macro_scanned"]
              + 24 * counters["n_macro_jumps"] + 16 * counters["n_vpackets"] * 8 + 80 * counters.get("n_bf_estimator_updates", 0) + 56 * n)
        note = ("jump kernel: 40 B/probe + 160 B/trace + macro-atom/vpacket terms + 56 B/packet; latency-bound by design. "
                "reference_equivalent_GBps is the SURVEY.md §8(d) byte count of the SAME packets (what the streaming "
                "formulation would have to move) divided by this kernel's time")
    achieved = ab / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": read_traffic(f"{args.algorithm}_{args.mode}_{args.lines}_{args.shells}", n),
                "kernel": ("tb::transport_pool_kernel" if args.algorithm == "jump" and not args.continuum and args.vpackets == 0
                           else f"tb::transport_{args.algorithm}_kernel"), "kernel_ms": k_ms, "algorithmic_bytes_per_launch": ab,
                "peak_source": peak_src, "definition": note,
                "reference_equivalent_GBps": ref_equiv / (k_ms * 1e-3) / 1e9,
                "per_packet": {"line_steps": counters["n_line_steps"] / n, "events": events / n,
                               "search_probes": counters["n_search_probes"] / n}}

    # ---- the streaming ("scan") kernel on a slice of the same packets: its roofline on the SURVEY.md §8(d) bytes ----
    scan_block = None
    # (skipped with virtual packets: both kernels resolve a volley by prefix search, so the streaming byte count does not apply)
    if args.algorithm == "jump" and not args.no_scan_reference and not args.continuum and args.vpackets == 0:
        ns = int(min(n, max(2_000_000, n // 20)))
        eng.set_option("algorithm", 0)
        eng.upload_packets(*(a[:ns] for a in host_in))
        scan_ms = []
        for i in range(3):
            eng.transport(True)
            eng.sync()
            if i > 0:
                scan_ms.append(eng.last_kernel_ms())
        sc = eng.counters()
        sab = alg_bytes(sc, ns)
        s_ms = float(np.mean(scan_ms))
        scan_block = {"kernel": "tb::transport_scan_kernel", "packets": ns, "kernel_ms": s_ms, "packets_per_s": ns / s_ms * 1e3,
                      "roofline": {"bound": "hbm", "achieved": sab / (s_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": sab / (s_ms * 1e-3) / 1e9 / peak,
                                   "traffic": read_traffic(f"scan_{args.mode}_{args.lines}_{args.shells}", ns),
                                   "definition": "SURVEY.md §8(d): 48 B/line-step + 32 B/event + macro-atom terms + 56 B/packet"}}
        eng.set_option("algorithm", 1)

    # ---- CPU baseline + spectrum parity on the same sample ----
    cpu = None
    spectrum_l2 = None
    if not args.no_cpu_baseline:
        n_threads = os.cpu_count() or 1
        rate, ns, dt, sample, ref, used = cpu_leg(model, args, n_threads, syn.BASE_SEED + 777, args.vpackets)
        cpu = {"value": rate, "unit": "packets/s", "cores": used, "kind": "port",
               "sample": f"{ns} packets of the same workload in {dt:.1f} s (oracle/tardis_oracle.c restatement of the "
                         f"reference loop; fastest calibrated configuration on {n_threads} host cores: {cpu_leg.last_choice})"}
        g = eng.run_packets(sample)
        a = emitted_spectrum(g["output_nus"], g["output_energies"], model.spectrum_frequency_grid, sample.time_of_simulation)
        b = emitted_spectrum(ref["output_nus"], ref["output_energies"], model.spectrum_frequency_grid, sample.time_of_simulation)
        spectrum_l2 = float(np.linalg.norm(a - b) / np.linalg.norm(b))

    line = {"metric": METRIC, "value": value, "unit": "packets/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "packets/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps,
                    "device_source": device_source,
                    "fused_spectrum_only": {"value": e2e_lean_value, "d2h_bytes_per_step": d2h_bytes - 2 * n * 8,
                                            "note": "same call without the per-packet output arrays: estimators + in-kernel "
                                                    "emitted/reabsorbed spectrum histograms come back (SURVEY.md §8f rank 2)"}},
            "roofline": roofline, "scan_kernel": scan_block, "cpu_baseline": cpu, "spectrum_l2_vs_oracle": spectrum_l2,
            "counters": counters}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
