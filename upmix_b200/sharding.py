"""Multi-GPU sharding of the extraction path: one process per GPU, no data-path collective.

The reference keeps the whole signal in one process (center_extraction.py:444-445, 499-501); bands are
its only unit of parallelism.  Here the independent units are
  * tracks of a batch: track i -> rank i mod world (tracks_for_rank), and
  * time segments of one long track: rank r computes output samples [a_r, b_r) and reads its input
    with a halo on each side (plan_segments / input_range).  Frames keep their global index, so a
    sharded run is bit-identical to the unsharded one (tests/test_gpu_parity.py, tests/test_sharding.py).
Results are gathered by plain copies into disjoint slices of the host output (gather_segments uses
torch.distributed point-to-point/gather on host tensors only to move the finished slices between
processes; nothing is exchanged while computing).
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np


def tracks_for_rank(n_tracks: int, rank: int, world: int) -> range:
    """Track indices rank `rank` of `world` processes owns (round-robin)."""
    if not 0 <= rank < world:
        raise ValueError("rank outside [0, world)")
    return range(rank, n_tracks, world)


def plan_segments(n_total: int, n_shards: int, align: int) -> List[Tuple[int, int]]:
    """Cut [0, n_total) into n_shards contiguous output segments whose interior boundaries are
    multiples of `align` (use the largest hop: every band's hop divides it).  Trailing shards may be
    empty when the track is short."""
    if n_shards < 1 or align < 1 or n_total < 0:
        raise ValueError("bad arguments")
    units = -(-n_total // align)
    per = -(-units // n_shards) if units else 0
    out = []
    for r in range(n_shards):
        a = min(n_total, r * per * align)
        b = min(n_total, (r + 1) * per * align)
        out.append((a, b))
    return out


def input_range(a: int, b: int, halo: int, n_total: int) -> Tuple[int, int]:
    """Input samples segment [a, b) needs: halo on each side, clipped to the track."""
    return max(0, a - halo), min(n_total, b + halo)


def largest_hop(band_extractors: Sequence) -> int:
    return max(int(b.hop_size) for b in band_extractors)


def extract_segment(L: np.ndarray, R: np.ndarray, sr: float, band_extractors: Sequence, a: int, b: int):
    """Output samples [a, b) of (centre, left, right) for the host signal L, R on the current CUDA
    device: copies only the halo'd slice to the device.  Returns float32 numpy arrays of length b-a."""
    from . import _native
    from . import center_extraction as ce
    torch = _native._torch()
    plan = ce.plan_for(band_extractors)
    n_total = len(L)
    if b <= a:
        z = np.zeros(0, dtype=np.float32)
        return z, z.copy(), z.copy()
    lo, hi = input_range(a, b, plan.halo, n_total)
    dev = f"cuda:{plan.device}"
    dl = torch.from_numpy(np.ascontiguousarray(L[lo:hi], dtype=np.float32)).to(dev)
    dr = torch.from_numpy(np.ascontiguousarray(R[lo:hi], dtype=np.float32)).to(dev)
    out = plan.process_segment(dl, dr, lo, n_total, a, b)
    return tuple(o.cpu().numpy() for o in out)


def run_rank(compute: Callable, L, R, n_total: int, rank: int, world: int, align: int):
    """Compute this rank's segment with `compute(a, b) -> tuple of arrays of length b-a`."""
    a, b = plan_segments(n_total, world, align)[rank]
    return (a, b), compute(a, b)


def gather_segments(local_bounds: Tuple[int, int], local_out: Sequence[np.ndarray], n_total: int, group=None,
                    dst: int = 0):
    """Assemble the full outputs on rank `dst` from every rank's finished segment.  Works on host
    arrays through torch.distributed (gloo or the CPU side of any backend); returns the tuple of full
    arrays on `dst`, None elsewhere.  With no process group it just checks that the segment is whole."""
    import torch
    import torch.distributed as dist
    n_ch = len(local_out)
    if not (dist.is_available() and dist.is_initialized()):
        if local_bounds != (0, n_total):
            raise ValueError("no process group, but the local segment is not the whole track")
        return tuple(np.asarray(x) for x in local_out)
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    payload = (tuple(int(v) for v in local_bounds), [np.ascontiguousarray(x, dtype=np.float32) for x in local_out])
    gathered = [None] * world if rank == dst else None
    dist.gather_object(payload, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    full = [np.zeros(n_total, dtype=np.float32) for _ in range(n_ch)]
    covered = 0
    for (a, b), chans in gathered:
        for ch in range(n_ch):
            full[ch][a:b] = chans[ch]
        covered += b - a
    if covered != n_total:
        raise RuntimeError(f"segments cover {covered} of {n_total} samples")
    return tuple(full)
