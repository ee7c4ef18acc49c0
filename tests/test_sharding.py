"""CPU: the N>1 host logic (upmix_b200/sharding.py) with world_size-2 gloo processes.  The compute
function is the oracle (no GPU here); what is tested is the partitioning, the halo arithmetic and the
gather: a 2-rank run must reproduce the unsharded result bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import upmix_oracle as uo
from upmix_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_segments_and_tracks():
    assert list(sharding.tracks_for_rank(10, 1, 4)) == [1, 5, 9]
    assert sum(len(sharding.tracks_for_rank(512, r, 8)) for r in range(8)) == 512
    for n_total, shards, align in ((172_800_000, 8, 16384), (100_003, 2, 16384), (5, 4, 64), (0, 3, 64), (16384 * 8, 8, 16384)):
        segs = sharding.plan_segments(n_total, shards, align)
        assert len(segs) == shards and segs[0][0] == 0 and segs[-1][1] == n_total
        for (a, b), (c, d) in zip(segs[:-1], segs[1:]):
            assert b == c and a <= b and (b % align == 0 or b == n_total)
    assert sharding.input_range(100, 200, 150, 1000) == (0, 350)
    assert sharding.input_range(800, 1000, 150, 1000) == (650, 1000)
    with pytest.raises(ValueError):
        sharding.tracks_for_rank(4, 4, 4)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sr = 48000
        bands = uo.chain([0, 500, 4000], 0.75, uo.blackman_harris, sr, max_block=4096)
        L, R = uo.synth_stereo(n_total, 77)
        L64, R64 = L.astype(np.float64), R.astype(np.float64)
        halo = max(b.n_fft for b in bands)
        align = max(b.hop for b in bands)

        def compute(a, b):
            # stand-in for sharding.extract_segment: the oracle on the halo'd excerpt.  The excerpt
            # starts on the frame grid (multiple of every hop), so [a, b) sees the same frames.
            lo, hi = sharding.input_range(a, b, halo, n_total)
            lo -= lo % align
            res = uo.upmix_multiband(bands, L64[lo:hi], R64[lo:hi])
            return tuple(x[a - lo:b - lo] for x in res)

        bounds, out = sharding.run_rank(compute, L64, R64, n_total, rank, world, align)
        full = sharding.gather_segments(bounds, out, n_total)
        if rank == 0:
            ref = uo.upmix_multiband(bands, L64, R64)
            # the last hop before a cut sees frames that reach past the excerpt's end in the sharded
            # run unless the halo covers them: it does (halo = n_fft), so equality is exact
            q.put([bool(np.array_equal(a, b)) for a, b in zip(ref, full)])
        else:
            assert full is None
        # batch-of-tracks assignment is disjoint and complete
        mine = list(sharding.tracks_for_rank(7, rank, world))
        allr = [None] * world
        dist.all_gather_object(allr, mine)
        assert sorted(sum(allr, [])) == list(range(7))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [48000 + 123, 3000])
def test_two_rank_gloo_time_sharding_matches_unsharded(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == [True, True, True]
