"""Packet sharding across GPUs: one process per GPU, contiguous packet ranges per rank, full table
replicas, and ONE all-reduce of the packed estimator buffer per MC iteration.
`torch.distributed` is plumbing only (NCCL on GPUs, gloo in the CPU tests of this host logic)."""
from __future__ import annotations

import numpy as np

ESTIMATOR_KEYS = ("j", "nu_bar", "j_blue", "edotlu", "vhist")


def shard_bounds(n_packets: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous index range [g*N/G, (g+1)*N/G) of rank g."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    return (n_packets * rank) // world_size, (n_packets * (rank + 1)) // world_size


class _DeviceBuffer:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can alias it."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3}


def estimator_tensor(engine):
    """torch.float64 CUDA tensor aliasing the engine's packed estimator buffer (everything that is summed over packets:
    j, nu_bar, vhist, the fused spectra, the luminosity sums, the continuum estimators, j_blue and edotlu shell-major;
    `engine.estimator_layout()` gives the offsets).  `set_model` may reallocate the buffer: call this again after it."""
    import torch

    ptr, count = engine.estimator_buffer()
    return torch.as_tensor(_DeviceBuffer(ptr, count), device=f"cuda:{engine.device}")


class _DeviceWords(_DeviceBuffer):
    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<i8", "data": (ptr, False), "version": 3}


def line_accumulator_tensor(engine):
    """(torch.int64 CUDA tensor aliasing the fixed-point difference arrays behind J_blue / Edotlu, their two scales)"""
    import torch

    ptr, count, s1, s2 = engine.line_accumulators()
    return torch.as_tensor(_DeviceWords(ptr, count), device=f"cuda:{engine.device}"), (s1, s2)


def scales_agree(scales, dist, device=None) -> bool:
    """True when every rank accumulated its line estimators at the same two fixed-point scales."""
    import torch

    mine = torch.tensor(scales, dtype=torch.float64, device=device)
    lo, hi = mine.clone(), mine.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool(torch.equal(lo, hi))


def all_reduce_estimators(engine, dist, exact_lines: bool = True) -> bool:
    """Sum the estimators over all ranks, in place on the device.  Returns whether the line estimators took the exact path.

    exact_lines (jump algorithm): J_blue / Edotlu are summed as the 64-bit integer words of their fixed-point difference
    arrays and finalised afterwards, which makes them bit-identical on every rank AND to a single-GPU run over the same
    packets for any number of GPUs (integer addition is associative; the f64 sums of the other tables are not).  The
    doubles before `off_j_blue` of the packed buffer (j, nu_bar, vhist, spectra, luminosities, continuum estimators) take
    the ordinary f64 all-reduce.  Falls back to one f64 all-reduce over the whole buffer when the engine runs the scan
    algorithm or the ranks' scales differ."""
    import torch

    engine.sync()
    est = estimator_tensor(engine)
    exact = False
    if exact_lines:
        try:
            words, scales = line_accumulator_tensor(engine)
        except (RuntimeError, ValueError):  # scan algorithm: no difference arrays (every rank runs the same algorithm)
            words, scales = None, (0.0, 0.0)
        exact = words is not None and scales_agree(scales, dist, device=est.device)
    if exact:
        dist.all_reduce(words)
        dist.all_reduce(est[: engine.estimator_layout()["off_j_blue"]])
        if est.is_cuda:
            torch.cuda.synchronize()
        engine.finalize_line_estimators()
        engine.sync()
    else:
        dist.all_reduce(est)
        if est.is_cuda:
            torch.cuda.synchronize()
    return exact


def pack_host_estimators(res: dict) -> np.ndarray:
    return np.concatenate([np.ascontiguousarray(res[k], dtype=np.float64).ravel() for k in ESTIMATOR_KEYS])


def unpack_host_estimators(buf: np.ndarray, like: dict) -> dict:
    out, off = {}, 0
    for k in ESTIMATOR_KEYS:
        n = like[k].size
        out[k] = buf[off:off + n].reshape(like[k].shape).copy()
        off += n
    return out


def all_reduce_host_results(res: dict, dist) -> dict:
    """Host-array variant of the collective (used by the gloo tests and by callers that already
    downloaded per-rank estimators): sums the five estimator arrays over ranks."""
    import torch

    t = torch.from_numpy(pack_host_estimators(res))
    dist.all_reduce(t)
    out = dict(res)
    out.update(unpack_host_estimators(t.numpy(), res))
    return out


def _collective_device(dist):
    """Where the tensors of a host-array collective must live: NCCL moves device memory only, gloo host memory."""
    import torch

    try:
        backend = str(dist.get_backend())
    except Exception:
        backend = "gloo"
    if "nccl" in backend and torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def _all_gather_rows(local_rows, sizes, dist):
    """all_gather of per-rank [k, n_r] float64 host rows with different n_r: -> list over the ranks of [k, n_r] numpy arrays."""
    import torch

    dev = _collective_device(dist)
    world, m, k = dist.get_world_size(), max(max(sizes), 1), len(local_rows)
    send = torch.zeros(k, m, dtype=torch.float64, device=dev)
    for j, row in enumerate(local_rows):
        row = np.ascontiguousarray(row, dtype=np.float64)
        send[j, : len(row)] = torch.from_numpy(row).to(dev)
    recv = [torch.zeros(k, m, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(recv, send)
    return [recv[r][:, : sizes[r]].cpu().numpy() for r in range(world)]


def gather_packet_outputs(local_nus: np.ndarray, local_energies: np.ndarray, n_packets: int, dist):
    """Per-packet outputs stay sharded during the run; this assembles them on every rank in index order."""
    world = dist.get_world_size()
    sizes = [shard_bounds(n_packets, r, world)[1] - shard_bounds(n_packets, r, world)[0] for r in range(world)]
    parts = _all_gather_rows([local_nus, local_energies], sizes, dist)
    return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])


def formal_integral_sharded(engine, dist, *, frequencies, **kwargs):
    """The formal integral over all ranks: frequencies are independent (numba_formal_integral's own `prange` runs over them,
    spectrum/formal_integral/formal_integral_numba.py:462), so rank g integrates the contiguous slice `shard_bounds` gives it on ITS
    copy of the tables -- after `all_reduce_estimators` + the source function every rank holds the same tables, bit for bit on the
    exact path -- and the luminosity densities are gathered: one small all_gather, no collective on the data path.  Every rank
    returns the full `luminosity_densities` [n]; `interpolation_ms` / `integral_ms` are this rank's.  kwargs: Engine.formal_integral's
    (inner_temperature, points, interpolate_shells, tables, electron_densities, sigma_thomson)."""
    if kwargs.get("want_intensities"):
        raise ValueError("formal_integral_sharded gathers luminosity densities only; ask one engine for I(nu, p)")
    freq = np.ascontiguousarray(frequencies, dtype=np.float64)
    if freq.ndim != 1:
        raise ValueError("frequencies must be one-dimensional")
    world, rank = dist.get_world_size(), dist.get_rank()
    bounds = [shard_bounds(len(freq), r, world) for r in range(world)]
    lo, hi = bounds[rank]
    res = engine.formal_integral(frequencies=freq[lo:hi], **kwargs)
    parts = _all_gather_rows([res["luminosity_densities"]], [b - a for a, b in bounds], dist)
    return dict(luminosity_densities=np.concatenate([p[0] for p in parts]), interpolation_ms=res["interpolation_ms"],
                integral_ms=res["integral_ms"], frequency_range=(lo, hi))
