"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the reference-facing
surface (upmix_b200.center_extraction / bela) and hence through the C ABI; results are compared with
the golden fixtures made from the unmodified reference and with the oracle on seeded inputs.

Bar (BASELINE.json north_star): >= 100 dB SNR per output channel, max abs error <= 1e-5 of full scale.
"""
import contextlib
import io
import os

import numpy as np
import pytest

from oracle import upmix_oracle as uo
from tests.parity import assert_parity

pytestmark = pytest.mark.gpu


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


@pytest.fixture(scope="module")
def ce():
    import torch
    assert torch.cuda.is_available()
    import upmix_b200.center_extraction as mod
    from upmix_b200 import _native
    _native.load_library()          # fail loudly if the extension is missing
    return mod


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


# ---------------------------------------------------------------------------------------------------
# golden fixtures of the reference
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,mode", [("cfg1_default6.npz", "raised_cosine"), ("cfg2_3band.npz", "raised_cosine"),
                                       ("cfg4_8band96k.npz", "raised_cosine"), ("hardzero_4band.npz", "hard_zero")])
def test_multiband_vs_reference_fixture(ce, golden_dir, name, mode):
    g = _load(golden_dir, name)
    ext = quiet(ce.chain_bands, list(g["edges"]), 0.75, ce.make_blackman_harris, float(g["sr"]), mode,
                max_block_size=int(g["max_block"]))
    assert [e.block_size for e in ext] == list(g["sizes"])
    L, R = g["in_L"].astype(np.float64), g["in_R"].astype(np.float64)     # float64, as sf.read hands over
    c, l, r = ce.extract_center_left_right_multi_band_in_memory(L, R, float(g["sr"]), ext)
    assert c.dtype == np.float32 and c.shape == L.shape
    peak = float(max(np.abs(L).max(), np.abs(R).max()))
    rep = assert_parity((g["ref_C"], g["ref_Ls"], g["ref_Rs"]), (c, l, r), peak, what=name)
    print(name, [(n, round(s, 1), e) for n, s, e in rep])


def test_single_band_variants_vs_reference_fixture(ce, golden_dir):
    g = _load(golden_dir, "single_band.npz")
    L, R = g["in_L"], g["in_R"]
    peak = float(max(np.abs(L).max(), np.abs(R).max()))
    e = ce.MultiBandExtractorAccu(512, 0.5, ce.make_sqrt_hann, 300.0, 5000.0, 48000, "raised_cosine", 100.0, 800.0)
    got = e.process_all_blocks(L, R)
    # sqrt-Hann ends in exact zeros: the reference's synthesis window is 0/EPS-safe there, compare finite parts
    assert_parity([g[f"ref_sqrt_hann50_{k}"] for k in ("C", "Ls", "Rs")], got, peak, what="sqrt_hann50")
    e = ce.MultiBandExtractorAccu(256, 0.75, ce.make_hann, 1000.0, 24000.0, 48000, "bogus_mode")
    got = e.process_all_blocks(L, R)
    assert_parity([g[f"ref_hann75_{k}"] for k in ("C", "Ls", "Rs")], got, peak, what="hann75")


def test_chunk_api_vs_reference_fixture(ce, golden_dir):
    """process_stereo_chunk / flush_final (center_extraction.py:353-424) with device-carried state."""
    g = _load(golden_dir, "single_band.npz")
    L, R = g["in_L"], g["in_R"]
    peak = float(max(np.abs(L).max(), np.abs(R).max()))
    e = ce.MultiBandExtractorAccu(1024, 0.75, ce.make_blackman_harris, 200.0, 2000.0, 48000, "raised_cosine", 50.0, 500.0)
    chunks = [np.stack(e.process_stereo_chunk(L[f * 256:f * 256 + 1024], R[f * 256:f * 256 + 1024])) for f in range(12)]
    got = np.stack(chunks)
    ref = g["ref_stream_chunks"]
    assert got.shape == ref.shape
    for ch, name in enumerate(("C", "Ls", "Rs")):
        assert_parity([ref[:, ch].reshape(-1)], [got[:, ch].reshape(-1)], peak, names=(name,), what="chunks")
    acc = np.stack([e.accumC, e.accumL, e.accumR])
    flush = np.stack(e.flush_final())
    assert np.array_equal(acc, flush)
    for ch, name in enumerate(("C", "Ls", "Rs")):
        assert_parity([g["ref_stream_flush"][ch]], [flush[ch]], peak, names=(name,), what="flush")
    assert not np.any(np.stack(e.flush_final()))


def test_bela_mode_vs_compiled_reference_fixture(ce, golden_dir):
    from upmix_b200 import bela
    for hw in (2048, 512):
        g = _load(golden_dir, f"bela_hw{hw}.npz")
        L, R = g["in_L"], g["in_R"]
        ol, orr = quiet(bela.bela_offline, L, R, 48000.0, hw)
        peak = float(max(np.abs(L).max(), np.abs(R).max()))
        assert_parity((g["ref_outL"], g["ref_outR"]), (ol, orr), peak, names=("outL", "outR"), what=f"bela hw{hw}")
        # block-by-block streaming gives the same stream
        up = bela.MultiBandUpmix()
        up.setThresholdMultiplier(bela.THRESHOLD_MULTI)
        quiet(up.setup, hw, 48000.0, 4, [0.0, 500.0, 2000.0, 8000.0, 24000.0])
        blocks = [up.process(L[i:i + hw], R[i:i + hw], hw) for i in range(0, len(L), hw)]
        sl = np.concatenate([b[0] for b in blocks])
        sr_ = np.concatenate([b[1] for b in blocks])
        assert_parity((g["ref_outL"], g["ref_outR"]), (sl, sr_), peak, names=("outL", "outR"), what=f"bela stream hw{hw}")
        # ... to float32 rounding against the offline call (whose band-limited bands take the decimated kernels) and
        # bit for bit against an offline plan that keeps the full-size kernels, the ones block streaming runs
        assert max(np.max(np.abs(sl - ol)), np.max(np.abs(sr_ - orr))) < 1e-6
        from upmix_b200 import _native
        n = (len(L) // hw) * hw
        plan = ce.plan_for(up.bands, _native.OUT_FOLD, _native.PLAN_STREAM_KERNELS)
        fl, fr = ce._run_plan(plan, L[:n], R[:n])
        assert np.array_equal(sl[3 * hw:], fl[:n - 3 * hw]) and np.array_equal(sr_[3 * hw:], fr[:n - 3 * hw])


# ---------------------------------------------------------------------------------------------------
# oracle on seeded inputs: every STFT size, ragged lengths
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_fft", [64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536])
def test_every_size_vs_oracle(ce, n_fft):
    sr = 48000
    f_low = 32.0 * sr / n_fft            # puts bin_low at 32 like the dynamic-resolution rule does
    f_high = min(4 * f_low, sr / 2)
    n = max(4 * n_fft + 1237, 9001)
    L, R = uo.synth_stereo(n, n_fft, stress=True)
    e = ce.MultiBandExtractorAccu(n_fft, 0.75, ce.make_blackman_harris, f_low, f_high, sr, "raised_cosine", f_low / 4, f_high / 4)
    b = uo.make_band(n_fft, 0.75, uo.blackman_harris, f_low, f_high, sr, "raised_cosine", f_low / 4, f_high / 4)
    assert np.array_equal(e.band_gain(), b.gain) and np.array_equal(e.synthesis_window, b.syn)
    ref = uo.process_band_batched(b, L.astype(np.float64), R.astype(np.float64))
    got = e.process_all_blocks(L, R)
    peak = float(max(np.abs(L).max(), np.abs(R).max()))
    rep = assert_parity(ref, got, peak, what=f"N={n_fft}")
    print(n_fft, [(nm, round(s, 1), er) for nm, s, er in rep])


@pytest.mark.parametrize("overlap", [0.5, 0.875])
def test_other_overlaps_vs_oracle(ce, overlap):
    """chain_bands takes any overlap (center_extraction.py:518-580; main.py fixes 0.75).  50 % and 87.5 % on bands up to
    8192 points run in the one-frame kernel; 50 % on larger bands runs on the 75 % kernels (decimated / four-step) with
    every other frame absent -- the odd frames add exact zeros, so the sum is the reference's two-frame overlap-add."""
    sr = 48000
    edges, max_block = ([0, 30, 120, 480, 1920, 7680], 65536) if overlap == 0.5 else ([0, 500, 2000, 8000], 8192)
    ext = quiet(ce.chain_bands, edges, overlap, ce.make_blackman_harris, sr, "raised_cosine", max_block_size=max_block)
    bands = uo.chain(edges, overlap, uo.blackman_harris, sr, max_block=max_block)
    assert [e.hop_size for e in ext] == [b.hop for b in bands]
    n = 5 * sr + 4321
    L, R = uo.synth_stereo(n, 17, stress=True)
    got = ce.extract_center_left_right_multi_band_in_memory(L, R, sr, ext)
    ref = uo.upmix_multiband(bands, L.astype(np.float64), R.astype(np.float64))
    peak = float(max(np.abs(L).max(), np.abs(R).max()))
    assert_parity(ref, got, peak, what=f"overlap {overlap}")
    if overlap == 0.5:
        # a dense band above 8192 points takes the four-step kernels
        e = ce.MultiBandExtractorAccu(16384, 0.5, ce.make_blackman_harris, 100.0, 20000.0, sr, "raised_cosine", 25.0, 0.0)
        b = uo.make_band(16384, 0.5, uo.blackman_harris, 100.0, 20000.0, sr, "raised_cosine", 25.0, 0.0)
        ref1 = uo.process_band_batched(b, L.astype(np.float64), R.astype(np.float64))
        assert_parity(ref1, e.process_all_blocks(L, R), peak, what="dense 16384, overlap 0.5")
        # a time shard of the six-band plan is bit-identical to the same samples of the whole call
        import torch
        plan = ce.plan_for(ext)
        dl, dr = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
        whole = plan.process(dl, dr)
        a, b = 32768 * 2, 32768 * 5 + 999
        lo, hi = max(0, a - plan.halo), min(n, b + plan.halo)
        seg = plan.process_segment(dl[lo:hi].contiguous(), dr[lo:hi].contiguous(), lo, n, a, b)
        for w, sg in zip(whole, seg):
            assert torch.equal(w[a:b], sg)


@pytest.mark.parametrize("n", [0, 1, 63, 64, 65, 255, 1000, 4097])
def test_ragged_and_tiny_lengths(ce, n):
    sr = 48000
    ext = quiet(ce.chain_bands, [0, 2000, 8000], 0.75, ce.make_blackman_harris, sr, "raised_cosine", max_block_size=1024)
    bands = uo.chain([0, 2000, 8000], 0.75, uo.blackman_harris, sr, max_block=1024)
    L, R = uo.synth_stereo(max(n, 1), 3)
    L, R = L[:n], R[:n]
    got = ce.extract_center_left_right_multi_band_in_memory(L, R, sr, ext)
    ref = uo.upmix_multiband(bands, L.astype(np.float64), R.astype(np.float64))
    assert all(g.shape == (n,) for g in got)
    if n:
        assert_parity(ref, got, 0.5, what=f"n={n}")


def test_wideband_and_full_gain(ce):
    """A single band covering everything (gain 1 at every bin): exercises DC and Nyquist bins."""
    sr = 48000
    for n_fft in (256, 4096, 16384):
        e = ce.MultiBandExtractorAccu(n_fft, 0.75, ce.make_blackman_harris, 0.0, sr / 2, sr, "raised_cosine", 0.0, 0.0)
        b = uo.make_band(n_fft, 0.75, uo.blackman_harris, 0.0, sr / 2, sr, "raised_cosine", 0.0, 0.0)
        assert np.all(b.gain == 1.0)
        L, R = uo.synth_stereo(3 * n_fft + 11, 9)
        L += np.float32(0.05)                       # DC offset
        R[::2] += np.float32(0.03)                  # energy at Nyquist
        R[1::2] -= np.float32(0.03)
        ref = uo.process_band_batched(b, L.astype(np.float64), R.astype(np.float64))
        got = e.process_all_blocks(L, R)
        assert_parity(ref, got, 1.0, what=f"wideband N={n_fft}")


# ---------------------------------------------------------------------------------------------------
# size-independent properties at larger sizes
# ---------------------------------------------------------------------------------------------------
def test_partition_invariance_segments_and_tracks(ce):
    """Time shards with halos and multi-track batches are bit-identical to the plain run."""
    import torch
    sr = 48000
    ext = quiet(ce.chain_bands, [0, 200, 2000], 0.75, ce.make_blackman_harris, sr, "raised_cosine")
    plan = ce.plan_for(ext)
    n = 5 * sr + 321
    L, R = uo.synth_stereo(n, 21)
    dl, dr = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    whole = [t.clone() for t in plan.process(dl, dr)]
    halo = plan.halo
    cuts = [0, 16384 * 3, 16384 * 7 + 5000, 200000, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        lo, hi = max(0, a - halo), min(n, b + halo)
        seg = plan.process_segment(dl[lo:hi].contiguous(), dr[lo:hi].contiguous(), lo, n, a, b)
        for ch, (w, s) in enumerate(zip(whole, seg)):
            d = (w[a:b] - s).abs()
            assert torch.equal(w[a:b], s), (a, b, ch, float(d.max()), int(d.argmax()))
    # tracks: a batch of 3 different tracks == 3 single runs
    Ls = torch.stack([dl, dr, dl.flip(0)])
    Rs = torch.stack([dr, dl, dr * 0.5])
    batch = plan.process(Ls, Rs)
    for t in range(3):
        single = plan.process(Ls[t].contiguous(), Rs[t].contiguous())
        for bch, sch in zip(batch, single):
            assert torch.equal(bch[t], sch)
    # run-to-run determinism
    again = plan.process(dl, dr)
    for w, s in zip(whole, again):
        assert torch.equal(w, s)


def test_known_answers_on_device(ce):
    import torch
    sr = 48000
    ext = quiet(ce.chain_bands, [0, 30, 120, 480, 1920, 7680], 0.75, ce.make_blackman_harris, sr, "raised_cosine")
    n = 20 * sr
    L, R = uo.synth_stereo(n, 33)
    dl, dr = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    # identical channels: all of the band-limited signal is centre, sides vanish
    c, l, r = ce.extract_center_left_right_multi_band_in_memory(dl, dl, sr, ext)
    assert float(l.abs().max()) < 1e-5 and float(r.abs().max()) < 1e-5 and float(c.abs().max()) > 0.05
    # hard-panned: no centre, no right
    c, l, r = ce.extract_center_left_right_multi_band_in_memory(dl, torch.zeros_like(dl), sr, ext)
    # (float32: L + i*0 through a complex FFT leaves ~1e-8 of leakage in the split, not exact zeros)
    assert float(c.abs().max()) < 1e-6 and float(r.abs().max()) < 1e-6 and float(l.abs().max()) > 0.05
    # Ls + C does not depend on the other channel (per-frame linearity)
    c1, l1, _ = ce.extract_center_left_right_multi_band_in_memory(dl, dr, sr, ext)
    c2, l2, _ = ce.extract_center_left_right_multi_band_in_memory(dl, (dr * 0.3).flip(0).contiguous(), sr, ext)
    assert float(((l1 + c1) - (l2 + c2)).abs().max()) < 2e-6
    # fold-down == Ls + 0.5 C / Rs + 0.5 C
    fl, fr = ce.extract_stereo_fold_down(dl, dr, sr, ext)
    c1, l1, r1 = ce.extract_center_left_right_multi_band_in_memory(dl, dr, sr, ext)
    assert float((fl - (l1 + 0.5 * c1)).abs().max()) < 2e-6 and float((fr - (r1 + 0.5 * c1)).abs().max()) < 2e-6


@pytest.mark.parametrize("seconds,edges", [(300, [0, 200, 2000]), (3600, [0, 200, 2000]), (1800, [0, 30, 120, 480, 1920, 7680])])
def test_long_track_sampled_against_oracle(ce, seconds, edges):
    """cfg 2 at a larger size and at BASELINE's full size (1-hour track: direct band sum, 8192-hop four-step waves,
    whole-wave run lengths), and main.py's default six bands on half an hour: parity on the first and last seconds
    and around an interior cut."""
    import torch
    sr = 48000
    ext = quiet(ce.chain_bands, edges, 0.75, ce.make_blackman_harris, sr, "raised_cosine")
    bands = uo.chain(edges, 0.75, uo.blackman_harris, sr)
    n = seconds * sr + 77
    L, R = uo.synth_stereo(n, 1, stress=True)
    out = ce.extract_center_left_right_multi_band_in_memory(torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda(), sr, ext)
    out = [o.cpu().numpy() for o in out]
    peak = float(max(np.abs(L).max(), np.abs(R).max()))
    margin = 65536
    for a, b in ((0, 3 * sr), (n // 2 - sr, n // 2 + sr), (n - 3 * sr, n)):
        lo, hi = max(0, a - margin), min(n, b + margin)
        # the oracle on an excerpt that starts at a multiple of the largest hop keeps the frame grid
        lo -= lo % 16384
        ref = uo.upmix_multiband(bands, L[lo:hi].astype(np.float64), R[lo:hi].astype(np.float64))
        # excerpt edges differ (ramp-in / missing future frames): compare the interior only
        ia = a if lo == 0 else max(a, lo + 49152)
        ib = b if hi == n else min(b, hi - 65536)
        assert_parity([x[ia - lo:ib - lo] for x in ref], [x[ia:ib] for x in out], peak, what=f"[{a},{b})")


def test_host_buffer_entry_point(ce, golden_dir):
    g = _load(golden_dir, "cfg2_3band.npz")
    ext = quiet(ce.chain_bands, list(g["edges"]), 0.75, ce.make_blackman_harris, 48000.0, "raised_cosine")
    plan = ce.plan_for(ext)
    c, l, r = plan.process_host(g["in_L"], g["in_R"])
    assert_parity((g["ref_C"], g["ref_Ls"], g["ref_Rs"]), (c, l, r), 0.5, what="process_host")


def test_unsupported_inputs_fail_loudly(ce):
    from upmix_b200 import _native
    e = ce.MultiBandExtractorAccu(1000, 0.75, ce.make_hann, 100.0, 1000.0, 48000)
    with pytest.raises(NotImplementedError):
        e.process_all_blocks(np.zeros(5000, np.float32), np.zeros(5000, np.float32))
    with pytest.raises(ValueError):
        ce.MultiBandExtractorAccu(64, 0.999, ce.make_hann, 100.0, 1000.0, 48000)
    w = np.ones(32, np.float32)
    with pytest.raises(_native.UpmixNativeError):
        _native.Plan([(32, 8, w, w, np.ones(17, np.float32))])


def test_pipelined_host_path_is_bit_identical(ce, monkeypatch):
    """extract(..., pinned CPU tensors): segment-pipelined H2D / kernels / D2H == one device-resident call."""
    import torch
    sr = 48000
    ext = quiet(ce.chain_bands, [0, 200, 2000], 0.75, ce.make_blackman_harris, sr, "raised_cosine")
    plan = ce.plan_for(ext)
    n = 95 * sr + 4321
    L, R = uo.synth_stereo(n, 44)
    hl = torch.from_numpy(L).pin_memory()
    hr = torch.from_numpy(R).pin_memory()
    dev = [t.clone() for t in plan.process(hl.cuda(), hr.cuda())]
    host = ce.extract_center_left_right_multi_band_in_memory(hl, hr, sr, ext)
    for d, h in zip(dev, host):
        assert not h.is_cuda and torch.equal(d.cpu(), h)
    monkeypatch.setenv("UPMIX_HOST_SEG", str(7 * sr))        # fixed 7 s segments
    host2 = plan.process_host_tensors(hl, hr)
    for d, h in zip(dev, host2):
        assert torch.equal(d.cpu(), h)
    # full segments above the direct-sum threshold, a shorter last one below it (it stages its bands):
    # one workspace sized for the full segment must serve both
    monkeypatch.setenv("UPMIX_DIRECT_MIN", str(6 * sr))
    host3 = plan.process_host_tensors(hl, hr)
    for d, h in zip(dev, host3):
        assert torch.equal(d.cpu(), h)
    # results are fresh allocations: a second call must not overwrite the first one's tensors
    keep = [h.clone() for h in host3]
    other = plan.process_host_tensors(hr, hl)
    assert all(torch.equal(k, h) for k, h in zip(keep, host3)) and not torch.equal(other[1], host3[1])
    # the call main.py makes (main.py:43-50, 78-80): float64 strided views of one interleaved array, pageable memory
    monkeypatch.delenv("UPMIX_HOST_SEG")
    monkeypatch.delenv("UPMIX_DIRECT_MIN")
    wave = np.stack([L, R], axis=1).astype(np.float64)
    got = ce.extract_center_left_right_multi_band_in_memory(wave[:, 0], wave[:, 1], sr, ext)
    for d, g in zip(dev, got):
        assert isinstance(g, np.ndarray) and g.dtype == np.float32 and np.array_equal(d.cpu().numpy(), g)
    got32 = ce.extract_center_left_right_multi_band_in_memory(L[::-1][::-1], R, sr, ext)     # pageable float32
    for d, g in zip(dev, got32):
        assert np.array_equal(d.cpu().numpy(), g)


def test_sharding_extract_segment_matches_whole(ce):
    from upmix_b200 import sharding
    sr = 48000
    ext = quiet(ce.chain_bands, [0, 30, 120, 480, 1920, 7680], 0.75, ce.make_blackman_harris, sr, "raised_cosine")
    n = 12 * sr + 99
    L, R = uo.synth_stereo(n, 55)
    whole = ce.extract_center_left_right_multi_band_in_memory(L, R, sr, ext)
    align = sharding.largest_hop(ext)
    for world in (2, 4, 8):
        parts = [sharding.extract_segment(L, R, sr, ext, a, b) for a, b in sharding.plan_segments(n, world, align)]
        for ch in range(3):
            assert np.array_equal(np.concatenate([p[ch] for p in parts]), whole[ch]), (world, ch)


def test_main_driver_writes_reference_named_files(ce, tmp_path):
    """upmix_b200.main.run: same modes and file names as the reference's main.py (main.py:110-157)."""
    from scipy.io import wavfile
    from upmix_b200 import main as drv
    sr = 48000
    L, R = uo.synth_stereo(2 * sr, 66)
    pcm = np.round(np.clip(np.stack([L, R], axis=1), -1, 1) * 32767).astype(np.int16)
    in_dir, out_dir = tmp_path / "in", tmp_path / "out"
    in_dir.mkdir()
    wavfile.write(str(in_dir / "noise.wav"), sr, pcm)
    bands = "b8192(0-500)_b4096(500-4000)_b512(4000-24000)"
    for mode, names in (("stereo_sum", [f"noise_Sum_{bands}_ov0.75.wav"]), ("AB", [f"noise_AB_{bands}_ov0.75.wav"]),
                        ("split", [f"noise_Ls_{bands}.wav", f"noise_C_{bands}.wav", f"noise_Rs_{bands}.wav"])):
        written = quiet(drv.run, "noise.wav", mode, str(in_dir), str(out_dir), [0, 500, 4000], max_block_size=8192)
        assert [os.path.basename(w) for w in written] == names
    # stereo_sum content against the oracle with main.py's peak normalisation (main.py:85-97, 143-146)
    wave = pcm.astype(np.float64) / 32768.0
    ob = uo.chain([0, 500, 4000], 0.75, uo.blackman_harris, sr, max_block=8192)
    c, l, r = uo.upmix_multiband(ob, wave[:, 0], wave[:, 1])
    scale = np.max(np.abs(wave)) / max(np.max(np.abs(l)), np.max(np.abs(c)), np.max(np.abs(r)), 1e-9)
    want = np.stack([(l + 0.5 * c) * scale, (r + 0.5 * c) * scale], axis=1)
    _, got = wavfile.read(str(out_dir / f"noise_Sum_{bands}_ov0.75.wav"))
    assert got.shape == want.shape
    assert np.max(np.abs(got.astype(np.float64) - np.rint(want * 32767.0))) <= 1.0          # sf.write's float -> PCM_16 scale
    with pytest.raises(FileNotFoundError):
        drv.run("missing.wav", "AB", str(in_dir), str(out_dir))
    assert quiet(drv.run, "noise.wav", "bogus", str(in_dir), str(out_dir), [0, 500, 4000], max_block_size=8192) == []


def test_device_peak_and_export_mixes(ce):
    """upmix_peak3 / upmix_export_mix against main.py's numpy arithmetic (main.py:85-97, 110-157)."""
    import torch
    from upmix_b200 import _native
    rng = np.random.default_rng(3)
    n = 100003
    c, l, r, il, ir = (rng.standard_normal(n).astype(np.float32) * 0.2 for _ in range(5))
    c[777] = -1.7
    dc, dl, dr, dil, dir_ = (torch.from_numpy(x).cuda() for x in (c, l, r, il, ir))
    pk = _native.peak3(dc, dl, dr).cpu().numpy()
    assert np.array_equal(pk, np.array([np.abs(c).max(), np.abs(l).max(), np.abs(r).max()], dtype=np.float32))
    scale = np.float32(0.37)
    sc, sl, sr_ = c * scale, l * scale, r * scale
    ab = _native.export_mix("AB", float(scale), dc, dl, dr, dil, dir_)[0].cpu().numpy()
    assert np.array_equal(ab[:, 0], (sl + sc) + sr_) and np.array_equal(ab[:, 1], il + ir)
    sp = [o.cpu().numpy() for o in _native.export_mix("split", float(scale), dc, dl, dr)]
    assert np.array_equal(sp[0][:, 0], sl) and not sp[0][:, 1].any()
    assert np.array_equal(sp[1][:, 0], sc) and np.array_equal(sp[1][:, 1], sc)
    assert np.array_equal(sp[2][:, 1], sr_) and not sp[2][:, 0].any()
    ss = _native.export_mix("stereo_sum", float(scale), dc, dl, dr)[0].cpu().numpy()
    # FMA contraction on the device: within one rounding of the two-step numpy result
    assert np.max(np.abs(ss[:, 0] - (sl + np.float32(0.5) * sc))) <= 1.2e-7 and np.max(np.abs(ss[:, 1] - (sr_ + np.float32(0.5) * sc))) <= 1.2e-7
    with pytest.raises(ValueError):
        _native.export_mix("nope", 1.0, dc, dl, dr)


def test_equal_stft_bands_are_merged_and_still_match(ce, golden_dir):
    """Bands with the same size/hop/windows run as one pipeline (shared forward and inverse transforms)."""
    g = _load(golden_dir, "cfg4_8band96k.npz")
    ext = quiet(ce.chain_bands, list(g["edges"]), 0.75, ce.make_blackman_harris, float(g["sr"]), "raised_cosine",
                max_block_size=8192)
    plan = ce.plan_for(ext)
    assert [e.block_size for e in ext].count(8192) == 4 and plan.n_pipelines == 5
    ext1 = quiet(ce.chain_bands, [0, 30, 120, 480, 1920, 7680], 0.75, ce.make_blackman_harris, 48000, "raised_cosine")
    assert ce.plan_for(ext1).n_pipelines == 5
    # different windows are not merged
    a = ce.MultiBandExtractorAccu(1024, 0.75, ce.make_hann, 100.0, 1000.0, 48000)
    b = ce.MultiBandExtractorAccu(1024, 0.75, ce.make_blackman_harris, 1000.0, 5000.0, 48000)
    assert ce.plan_for([a, b]).n_pipelines == 2
    # merged result == sum of the single-band results (up to float32 rounding of the two summation orders)
    L, R = uo.synth_stereo(60000, 8, stress=True)
    whole = ce.extract_center_left_right_multi_band_in_memory(L, R, 96000.0, ext)
    parts = [e.process_all_blocks(L, R) for e in ext]
    for ch in range(3):
        want = np.sum([p[ch].astype(np.float64) for p in parts], axis=0)
        assert uo.snr_db(want, whole[ch]) > 120


def test_multi_track_batch_with_large_band_and_strides(ce):
    """Batches of tracks through the four-step path (waves shared by tracks), non-contiguous track strides."""
    import torch
    sr = 48000
    ext = quiet(ce.chain_bands, [0, 1000], 0.75, ce.make_blackman_harris, sr, "raised_cosine", max_block_size=16384)
    plan = ce.plan_for(ext)
    n = 70000
    tracks = 5
    big = torch.zeros((tracks, 2, n + 64), dtype=torch.float32, device="cuda")
    for t in range(tracks):
        L, R = uo.synth_stereo(n, 300 + t)
        big[t, 0, :n] = torch.from_numpy(L).cuda()
        big[t, 1, :n] = torch.from_numpy(R).cuda()
    Ls, Rs = big[:, 0, :n], big[:, 1, :n]            # row stride 2*(n+64): strided views
    out = plan.process(Ls, Rs)
    bands = uo.chain([0, 1000], 0.75, uo.blackman_harris, sr, max_block=16384)
    for t in (0, tracks - 1):
        ref = uo.upmix_multiband(bands, Ls[t].cpu().numpy().astype(np.float64), Rs[t].cpu().numpy().astype(np.float64))
        assert_parity(ref, [o[t].cpu().numpy() for o in out], 0.5, what=f"track {t}")
        single = plan.process(Ls[t].contiguous(), Rs[t].contiguous())
        for o, s in zip(out, single):
            assert torch.equal(o[t], s)


def test_c_abi_argument_errors(ce):
    """Error codes and messages instead of exceptions or crashes across the boundary."""
    import ctypes
    import torch
    from upmix_b200 import _native
    lib = _native.load_library()
    e = ce.MultiBandExtractorAccu(1024, 0.75, ce.make_blackman_harris, 200.0, 2000.0, 48000)
    plan = ce.plan_for([e])
    n = 5000
    x = torch.zeros(n, device="cuda")
    out = torch.zeros((3, n), device="cuda")
    ws = torch.zeros(plan.workspace_bytes(n, 1), dtype=torch.uint8, device="cuda")
    args = [plan._h, x.data_ptr(), x.data_ptr(), n, 1, n, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), n]
    assert lib.upmix_process(*args, ws.data_ptr(), 16, None) == -4 and b"workspace too small" in lib.upmix_last_error()
    assert lib.upmix_process(plan._h, None, x.data_ptr(), n, 1, n, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), n,
                             ws.data_ptr(), ws.numel(), None) == -1
    assert lib.upmix_process(*args[:4], 0, *args[5:], ws.data_ptr(), ws.numel(), None) == -1          # n_tracks = 0
    # a shard whose input does not cover its halo
    assert lib.upmix_process_segment(plan._h, x.data_ptr(), x.data_ptr(), 1000, 2000, n, 1000, 3000, 1, n, out[0].data_ptr(),
                                     out[1].data_ptr(), out[2].data_ptr(), n, ws.data_ptr(), ws.numel(), None) == -1
    assert b"halo" in lib.upmix_last_error()
    # hop that does not divide the size, size above the supported range
    w = np.ones(1024, np.float32)
    gn = np.ones(513, np.float32)
    with pytest.raises(_native.UpmixNativeError):
        _native.Plan([(1024, 300, w, w, gn)])
    with pytest.raises(_native.UpmixNativeError):
        _native.Plan([(131072, 32768, np.ones(131072, np.float32), np.ones(131072, np.float32), np.ones(65537, np.float32))])
    # fine call still works afterwards
    assert lib.upmix_process(*args, ws.data_ptr(), ws.numel(), None) == 0
    torch.cuda.synchronize()


def test_fold_down_with_large_band_and_stream_multitrack(ce):
    import torch
    sr = 48000
    ext = quiet(ce.chain_bands, [0, 300, 3000], 0.75, ce.make_blackman_harris, sr, "raised_cosine", max_block_size=16384)
    L, R = uo.synth_stereo(50000, 77)
    c, l, r = ce.extract_center_left_right_multi_band_in_memory(L, R, sr, ext)
    fl, fr = ce.extract_stereo_fold_down(L, R, sr, ext)
    assert np.max(np.abs(fl - (l + 0.5 * c))) < 1e-6 and np.max(np.abs(fr - (r + 0.5 * c))) < 1e-6
    # block streaming, two tracks at once == offline result delayed by stream.delay: bit for bit on a plan whose offline
    # call runs the kernels block streaming runs (PLAN_STREAM_KERNELS), to float32 rounding on the default plan
    from upmix_b200 import _native
    small = quiet(ce.chain_bands, [0, 1000, 6000], 0.75, ce.make_blackman_harris, sr, "raised_cosine", max_block_size=2048)
    plan = ce.plan_for(small, _native.OUT_LSCRS, _native.PLAN_STREAM_KERNELS)
    st = plan.stream_open(2)
    n = 2048 * 12
    A = torch.from_numpy(np.stack([uo.synth_stereo(n, 1)[0], uo.synth_stereo(n, 2)[0]])).cuda()
    B = torch.from_numpy(np.stack([uo.synth_stereo(n, 1)[1], uo.synth_stereo(n, 2)[1]])).cuda()
    off = plan.process(A, B)
    blocks = [st.block(A[:, i:i + 512].contiguous(), B[:, i:i + 512].contiguous()) for i in range(0, n, 512)]
    d = st.delay
    dflt = ce.plan_for(small).process(A, B)
    for ch in range(3):
        got = torch.cat([b[ch] for b in blocks], dim=1)
        assert torch.equal(got[:, d:], off[ch][:, :n - d]) and not got[:, :d].any()
        assert float((dflt[ch] - off[ch]).abs().max()) < 1e-6


def test_stream_graph_replay_and_split_runs_are_bit_identical(ce):
    """Steady-state blocks are replayed as CUDA graphs (position folded modulo the largest STFT size, caller buffers
    patched into two graph nodes) and small bands run as several CTAs per block: same stream, bit for bit, as a plan
    that launches every block the plain way -- with input / output buffers that move from block to block, two block
    sizes, and a block-size change in mid-stream."""
    import torch
    from upmix_b200 import _native
    sr = 48000
    bands = quiet(ce.chain_bands, [0, 700, 3000, 9000], 0.75, ce.make_blackman_harris, sr, "raised_cosine", max_block_size=4096)
    n = 4096 * 14
    L, R = uo.synth_stereo(n, 31, stress=True)
    A, B = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    tables = [b.plan_tables() for b in bands]
    os.environ["UPMIX_GRAPHS"] = "0"
    try:
        plain = _native.Plan(tables, _native.OUT_LSCRS, flags=_native.PLAN_STREAM_KERNELS)
    finally:
        os.environ.pop("UPMIX_GRAPHS")
    graphed = _native.Plan(tables, _native.OUT_LSCRS, flags=_native.PLAN_STREAM_KERNELS)
    off = graphed.process(A, B)
    for sizes in ([1024] * 56, [2048] * 28, [1024] * 20 + [4096] * 6 + [1024] * 12):
        outs = []
        for plan in (plain, graphed):
            st = plan.stream_open(1)
            pos, keep, blocks = 0, [], []
            for k, m in enumerate(sizes):
                # fresh (moving) buffers: slices at odd offsets of a scratch tensor, kept alive so that pointers differ
                scratch = torch.empty(2, m + 8, device="cuda")
                o = 4 * (k % 2)
                scratch[0, o:o + m] = A[pos:pos + m]
                scratch[1, o:o + m] = B[pos:pos + m]
                keep.append(scratch)
                blocks.append(st.block(scratch[0, o:o + m], scratch[1, o:o + m]))
                pos += m
            assert pos == n
            outs.append([torch.cat([b[ch] for b in blocks]) for ch in range(3)])
        d = graphed.stream_open(1).delay
        for ch in range(3):
            assert torch.equal(outs[0][ch], outs[1][ch]), (sizes[:3], ch)
            assert torch.equal(outs[1][ch][d:], off[ch][:n - d])


def test_stream_very_long_blocks(ce):
    """Blocks above 65536 samples take the copy path of upmix_stream_block (2-D copies instead of the staging kernel, no
    graph replay): same stream as the offline call, bit for bit."""
    import torch
    from upmix_b200 import _native
    sr = 48000
    bands = quiet(ce.chain_bands, [0, 1000, 6000], 0.75, ce.make_blackman_harris, sr, "raised_cosine", max_block_size=2048)
    plan = ce.plan_for(bands, _native.OUT_LSCRS, _native.PLAN_STREAM_KERNELS)
    m = 131072
    n = 3 * m
    L, R = uo.synth_stereo(n, 41)
    A, B = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    off = plan.process(A, B)
    st = plan.stream_open(1)
    blocks = [st.block(A[i:i + m], B[i:i + m]) for i in range(0, n, m)]
    d = st.delay
    for ch in range(3):
        got = torch.cat([b[ch] for b in blocks])
        assert torch.equal(got[d:], off[ch][:n - d]) and not got[:d].any()


def test_repeated_runs_are_bit_identical(ce):
    """compute-sanitizer's racecheck is closed on this pool, so races are looked for by their symptom: a kernel with a
    shared-memory race (a missing barrier between passes, a tile read before its cp.async / TMA copy landed, an overlap-add
    carry taken from the wrong tile) gives results that change from run to run or with what else is resident.  Every kernel
    family -- decimated, frame-batched, one-frame, four-step, band sum, with and without accumulation -- runs the same input
    eight times, alone and with a second plan busy on another stream, and must return the same bits."""
    import torch
    sr = 48000
    L, R = uo.synth_stereo(25 * sr + 123, 5, stress=True)
    dl, dr = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    noise = torch.randn(2, 40 * sr, device="cuda")
    other = ce.plan_for(quiet(ce.chain_bands, [0, 900, 5000], 0.75, ce.make_blackman_harris, sr, "raised_cosine", max_block_size=4096))
    side = torch.cuda.Stream()
    cases = [([0, 200, 2000], 65536, {}), ([0, 30, 120, 480, 1920, 7680], 65536, {}), ([0, 400, 4000], 16384, {"UPMIX_DEC": "0"}),
             ([0, 1000, 6000], 2048, {"UPMIX_FB": "0"})]
    for edges, max_block, env in cases:
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            ext = quiet(ce.chain_bands, edges, 0.75, ce.make_blackman_harris, sr, "raised_cosine", max_block_size=max_block)
            from upmix_b200 import _native
            plan = _native.Plan([b.plan_tables() for b in ext], _native.OUT_LSCRS)
            for direct_min in ("1", str(1 << 60)):                 # pipelines adding into the outputs / per-band slots + band sum
                os.environ["UPMIX_DIRECT_MIN"] = direct_min
                ref = None
                for rep in range(8):
                    if rep % 2:
                        with torch.cuda.stream(side):
                            other.process(noise[0], noise[1])
                    got = torch.stack(plan.process(dl, dr))
                    if ref is None:
                        ref = got.clone()
                    assert torch.equal(ref, got), (edges, direct_min, rep)
                torch.cuda.synchronize()
        finally:
            os.environ.pop("UPMIX_DIRECT_MIN", None)
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v


def test_pcm16_edge_kernels(ce):
    """WAV edge on the device: int16 stereo -> planar float32 + peak, float32 stereo -> int16."""
    import torch
    from upmix_b200 import _native
    rng = np.random.default_rng(9)
    pcm = rng.integers(-32768, 32768, size=(70001, 2), dtype=np.int16)
    pcm[5] = (-32768, 32767)
    l, r, peak = _native.pcm16_to_planar(torch.from_numpy(pcm).cuda())
    assert np.array_equal(l.cpu().numpy(), pcm[:, 0].astype(np.float32) / 32768.0)
    assert np.array_equal(r.cpu().numpy(), pcm[:, 1].astype(np.float32) / 32768.0)
    assert float(peak.cpu()[0]) == 1.0
    x = (rng.standard_normal((50001, 2)) * 0.5).astype(np.float32)
    x[0] = (1.5, -1.5)
    got = _native.stereo_to_pcm16(torch.from_numpy(x).cuda()).cpu().numpy()
    want = np.clip(np.rint((x * np.float32(32767.0)).astype(np.float64)), -32768, 32767).astype(np.int16)
    assert np.array_equal(got, want) and tuple(got[0]) == (32767, -32768)          # out of range: saturates
    # read scale 1/32768, write scale 32767 (libsndfile's pair): a round trip moves a sample by at most one step
    back = _native.stereo_to_pcm16(torch.stack([l, r], dim=1).contiguous()).cpu().numpy()
    assert np.max(np.abs(back.astype(np.int32) - pcm.astype(np.int32))) <= 1
    # empty input through the drop-in call
    e = ce.MultiBandExtractorAccu(1024, 0.75, ce.make_blackman_harris, 200.0, 2000.0, 48000)
    z = torch.zeros(0, device="cuda")
    assert all(o.numel() == 0 for o in ce.extract_center_left_right_multi_band_in_memory(z, z, 48000, [e]))


@pytest.mark.parametrize("n_fft", [2048, 16384])
def test_chunk_api_every_path_vs_oracle(ce, n_fft):
    """process_stereo_chunk on the fused and on the four-step path: consecutive frames of a signal give the
    offline result hop by hop, flush_final gives the tail (center_extraction.py:353-424)."""
    sr = 48000
    f_low = 32.0 * sr / n_fft
    e = ce.MultiBandExtractorAccu(n_fft, 0.75, ce.make_blackman_harris, f_low, 4 * f_low, sr, "raised_cosine", f_low / 4, f_low)
    b = uo.make_band(n_fft, 0.75, uo.blackman_harris, f_low, 4 * f_low, sr, "raised_cosine", f_low / 4, f_low)
    H = n_fft // 4
    n = 6 * H + n_fft + 123
    L, R = uo.synth_stereo(n, 5)
    ref = uo.process_band_frames(b, L.astype(np.float64), R.astype(np.float64))
    frames = -(-n // H)                                     # every frame that starts inside the signal (CE:448-460)
    Lp = np.concatenate([L, np.zeros(frames * H + n_fft - n, np.float32)])
    Rp = np.concatenate([R, np.zeros(frames * H + n_fft - n, np.float32)])
    got = [np.stack(e.process_stereo_chunk(Lp[f * H:f * H + n_fft], Rp[f * H:f * H + n_fft])) for f in range(frames)]
    got = np.concatenate(got, axis=1)                       # [3, frames*H]
    tail = np.stack(e.flush_final())                        # [3, n_fft]
    assert tail.shape == (3, n_fft)
    assert_parity(ref, list(got[:, :n]), 0.5, what=f"chunks N={n_fft}")


@pytest.mark.parametrize("n_fft,top_bin", [(16384, 511), (16384, 512), (32768, 40), (65536, 511), (65536, 700)])
def test_large_band_row_pruning_boundary(ce, n_fft, top_bin):
    """Four-step path on both sides of the band-limited (pruned) row kernel's eligibility rule: every
    non-zero gain below bin 512 -> pruned row transforms, otherwise the full ones.  Hard-zero pass band
    [3, top_bin] so the last kept bin is exactly top_bin (rows 0 and 8 and their self-mirrored bins included)."""
    sr = 48000
    df = sr / n_fft
    f_low, f_high = 3 * df, top_bin * df
    e = ce.MultiBandExtractorAccu(n_fft, 0.75, ce.make_blackman_harris, f_low, f_high, sr, "hard_zero", 0.0, 0.0)
    b = uo.make_band(n_fft, 0.75, uo.blackman_harris, f_low, f_high, sr, "hard_zero", 0.0, 0.0)
    g = e.band_gain()
    assert np.array_equal(g, b.gain) and int(np.nonzero(g)[0].max()) == top_bin
    n = 2 * n_fft + 4321
    L, R = uo.synth_stereo(n, 100 + top_bin, stress=True)
    ref = uo.process_band_batched(b, L.astype(np.float64), R.astype(np.float64))
    got = e.process_all_blocks(L, R)
    peak = float(max(np.abs(L).max(), np.abs(R).max()))
    assert_parity(ref, got, peak, what=f"N={n_fft} top_bin={top_bin}")


def test_direct_band_sum_is_identical_to_staged(ce, monkeypatch):
    """The two ways of summing the bands (per-band slots + band_sum_kernel for short inputs, pipelines
    adding straight into the outputs for long ones) do the same float32 additions in the same order:
    Ls/C/Rs, fold-down with and without a four-step band, merged bands, track batches, time shards."""
    import torch
    sr = 48000
    n = 3 * sr + 777
    L, R = uo.synth_stereo(n, 31, stress=True)
    dl, dr = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    Ls = torch.stack([dl, dr, dl.flip(0)])
    Rs = torch.stack([dr, dl, dr * 0.5])
    cases = [([0, 200, 2000], 65536), ([0, 30, 120, 480, 1920, 7680], 65536), ([0, 300, 3000], 16384),
             ([0, 500, 2000, 8000], 8192)]
    for edges, max_block in cases:
        ext = quiet(ce.chain_bands, edges, 0.75, ce.make_blackman_harris, sr, "raised_cosine", max_block_size=max_block)
        for fold in (False, True):
            from upmix_b200 import _native
            plan = ce.plan_for(ext, _native.OUT_FOLD if fold else _native.OUT_LSCRS)
            res = {}
            for mode, thr in (("staged", str(1 << 60)), ("direct", "1")):
                monkeypatch.setenv("UPMIX_DIRECT_MIN", thr)
                whole = [t.clone() for t in plan.process(dl, dr)]
                batch = [t.clone() for t in plan.process(Ls, Rs)]
                a, b = 16384 * 3, 16384 * 6 + 1234
                lo, hi = max(0, a - plan.halo), min(n, b + plan.halo)
                seg = [t.clone() for t in plan.process_segment(dl[lo:hi].contiguous(), dr[lo:hi].contiguous(), lo, n, a, b)]
                for w, s in zip(whole, seg):
                    assert torch.equal(w[a:b], s), (edges, fold, mode)
                res[mode] = whole + batch
            for x, y in zip(res["staged"], res["direct"]):
                assert torch.equal(x, y), (edges, fold)


def test_fir_filter_vs_reference_fixture(ce, golden_dir):
    """apply_fir_filter (filter_design.py:54-59, scipy.signal.lfilter in float64) on the device in float32:
    >= 100 dB SNR against the reference's own outputs; shapes, dtypes and the pass-through taps."""
    import torch
    from upmix_b200 import filter_design as fd
    g = _load(golden_dir, "fir.npz")
    L, R = g["in_L"], g["in_R"]
    for x, taps, ref in ((L.astype(np.float64), g["ref_hp"], g["ref_y_hp"]), (R.astype(np.float64), g["ref_lp"], g["ref_y_lp"]),
                         (L, g["ref_hp_short"], g["ref_y_short"])):
        y = fd.apply_fir_filter(x, taps)
        assert y.shape == ref.shape and y.dtype == ref.dtype
        snr = uo.snr_db(ref, y)
        err = float(np.max(np.abs(ref - y)))
        assert snr >= 100.0 and err <= 1e-5, (snr, err)
    assert np.array_equal(fd.apply_fir_filter(L, g["ref_pass"]), L)
    # tracks along the first axis, ragged lengths around the tile size, CUDA tensors in and out
    for n in (1, 7, 2047, 2048, 2049, 5000):
        x = np.stack([L[:n], R[:n]])
        y = fd.apply_fir_filter(torch.from_numpy(x).cuda(), g["ref_hp_short"])
        assert y.is_cuda and y.dtype == torch.float32 and tuple(y.shape) == (2, n)
        from scipy.signal import lfilter
        ref = lfilter(g["ref_hp_short"].astype(np.float64), 1.0, x.astype(np.float64))
        assert float(np.max(np.abs(ref - y.cpu().numpy()))) <= 1e-6


def test_fused_emit_is_bit_identical_to_ring_copy_out(ce, tmp_path):
    """The fused-emit kernels (finished hops leave from the last inverse passes) and the ring + copy-out kernels
    do the same float32 additions in the same order: bit-identical outputs.  The switch is read once per
    process, so the two variants run in subprocesses."""
    import subprocess
    import sys
    script = tmp_path / "run.py"
    script.write_text(
        "import sys, contextlib, io, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "import upmix_b200.center_extraction as ce\n"
        "from oracle import upmix_oracle as uo\n"
        "sr = 48000\n"
        "outs = []\n"
        "for edges, mb in (([0, 200, 2000], 8192), ([0, 500, 2000, 8000], 4096), ([0, 3000, 9000], 512)):\n"
        "    with contextlib.redirect_stdout(io.StringIO()):\n"
        "        ext = ce.chain_bands(edges, 0.75, ce.make_blackman_harris, sr, 'raised_cosine', max_block_size=mb)\n"
        "    L, R = uo.synth_stereo(2 * sr + 333, 5, stress=True)\n"
        "    outs += list(ce.extract_center_left_right_multi_band_in_memory(L, R, sr, ext))\n"
        "    outs += list(ce.extract_stereo_fold_down(L, R, sr, ext))\n"
        "np.savez(sys.argv[1], *outs)\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    res = {}
    for flag in ("0", "1"):
        out = tmp_path / f"out{flag}.npz"
        env = dict(os.environ, UPMIX_DIRECT_EMIT=flag, UPMIX_DIRECT_MIN="1")
        subprocess.run([sys.executable, str(script), str(out)], check=True, env=env)
        res[flag] = np.load(out)
    assert len(res["0"].files) == len(res["1"].files) > 0
    for k in res["0"].files:
        assert np.array_equal(res["0"][k], res["1"][k]), k


def test_no_writes_outside_outputs_and_workspace(ce):
    """compute-sanitizer is not available on the GPU pool, so the bounds are checked the plain way: outputs and
    workspace sit between canary regions, and after runs that exercise every kernel family (fused sizes with
    and without the fused middle, the four-step path, staged and direct band sums, track batches with strides,
    a time shard) the canaries must be untouched."""
    import ctypes
    import torch
    from upmix_b200 import _native
    lib = _native.load_library()
    sr = 48000
    G = 4096                                       # canary elements on each side
    canary = torch.tensor([0x7fc0dead], dtype=torch.int32).view(torch.float32).item()

    def guarded(n_elems, dtype=torch.float32):
        buf = torch.empty(n_elems + 2 * G, dtype=dtype, device="cuda")
        if dtype == torch.float32:
            buf.view(torch.int32).fill_(0x7fc0dead)
        else:
            buf.fill_(0xA5)
        return buf

    def intact(buf, n_elems, dtype=torch.float32):
        if dtype == torch.float32:
            v = buf.view(torch.int32)
            return bool((v[:G] == 0x7fc0dead).all()) and bool((v[G + n_elems:] == 0x7fc0dead).all())
        return bool((buf[:G] == 0xA5).all()) and bool((buf[G + n_elems:] == 0xA5).all())

    cases = [([0, 30, 120, 480, 1920, 7680], 65536, 1, 3 * 65536 + 4321),      # default six bands: merged four-step, 16384, fused
             ([0, 200, 2000], 8192, 3, 40000),                                  # fused middle sizes, track batch
             ([0, 3000, 9000], 512, 2, 9999),                                   # small sizes
             ([0, 1000], 65536, 1, 70001)]                                      # four-step band with gain beyond bin 512
    for direct_min in ("1", str(1 << 60)):
        os.environ["UPMIX_DIRECT_MIN"] = direct_min
        try:
            for edges, mb, tracks, n in cases:
                ext = quiet(ce.chain_bands, edges, 0.75, ce.make_blackman_harris, sr, "raised_cosine", max_block_size=mb)
                plan = ce.plan_for(ext)
                stride = n + 37                                                 # odd stride: scalar store paths
                L = torch.randn(tracks * stride, device="cuda") * 0.1
                R = torch.randn(tracks * stride, device="cuda") * 0.1
                for a, b in ((0, n), (n // 3 + 5, 2 * n // 3 + 11)):            # whole track, then a shard
                    outs = [guarded(tracks * stride) for _ in range(3)]
                    wsb = plan.workspace_bytes(b - a, tracks)
                    ws = guarded(wsb, torch.uint8)
                    ws_ptr = ws.data_ptr() + G
                    pad = (-ws_ptr) % 256                                      # the ABI wants a 256-byte aligned workspace
                    assert pad <= G - 256
                    rc = lib.upmix_process_segment(plan._h, L.data_ptr(), R.data_ptr(), 0, n, n, a, b, tracks, stride,
                                                   *[o.data_ptr() + 4 * G for o in outs], stride, ws_ptr + pad, wsb - pad if wsb else 0, None)
                    if rc == -4:                                               # alignment padding ate into a tight workspace: retry larger
                        ws = guarded(wsb + 256, torch.uint8)
                        ws_ptr = ws.data_ptr() + G
                        pad = (-ws_ptr) % 256
                        rc = lib.upmix_process_segment(plan._h, L.data_ptr(), R.data_ptr(), 0, n, n, a, b, tracks, stride,
                                                       *[o.data_ptr() + 4 * G for o in outs], stride, ws_ptr + pad, wsb, None)
                        wsb += 256
                    assert rc == 0, lib.upmix_last_error()
                    torch.cuda.synchronize()
                    assert intact(ws, wsb, torch.uint8), ("workspace", edges, a, b, direct_min)
                    for o in outs:
                        assert intact(o, tracks * stride), ("output", edges, a, b, direct_min)
                        body = o[G:G + tracks * stride].view(tracks, stride)
                        # the gaps between tracks (stride - n elements) and, for a shard, everything past its length stay canaries
                        assert bool((body[:, b - a:].view(torch.int32) == 0x7fc0dead).all()), ("gap", edges, a, b, direct_min)
                        assert bool(torch.isfinite(body[:, :b - a]).all())
        finally:
            os.environ.pop("UPMIX_DIRECT_MIN", None)
    assert canary != canary                        # the canary is a NaN: any arithmetic on it would have shown


def test_large_batch_goes_through_in_track_groups(ce):
    """A batch of more tracks than one four-step wave holds (32) is processed in groups of tracks; every track
    must equal its own single-track run, bit for bit, in both ways of summing the bands."""
    import torch
    sr = 48000
    ext = quiet(ce.chain_bands, [0, 300, 3000], 0.75, ce.make_blackman_harris, sr, "raised_cosine", max_block_size=16384)
    plan = ce.plan_for(ext)
    n, tracks = 30000, 40
    g = torch.Generator(device="cuda").manual_seed(3)
    L = 0.1 * torch.randn((tracks, n), device="cuda", generator=g)
    R = 0.1 * torch.randn((tracks, n), device="cuda", generator=g)
    for direct_min in ("1", str(1 << 60)):
        os.environ["UPMIX_DIRECT_MIN"] = direct_min
        try:
            batch = [t.clone() for t in plan.process(L, R)]
            for t in (0, 17, 31, 32, 39):
                single = plan.process(L[t].contiguous(), R[t].contiguous())
                for bch, sch in zip(batch, single):
                    assert torch.equal(bch[t], sch), (t, direct_min)
        finally:
            os.environ.pop("UPMIX_DIRECT_MIN", None)


# ---------------------------------------------------------------------------------------------------
# decimated path (upmix_dec.cu): eligibility boundaries, every (P, Q) family, waves, merged bands, fold-down
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_fft,top_bin", [(2048, 5), (2048, 127), (2048, 128), (4096, 255), (4096, 256), (8192, 127),
                                           (8192, 128), (8192, 511), (8192, 512), (16384, 255), (16384, 256),
                                           (32768, 300), (65536, 127), (65536, 256), (65536, 511), (65536, 512)])
def test_decimated_band_boundaries(ce, n_fft, top_bin):
    """Both sides of the decimated kernels' eligibility rule (every live bin below P = 128 / 256 / 512 <= n_fft/16):
    hard-zero pass band [3, top_bin], so the last live bin is exactly top_bin; compared with the oracle, and with the
    full-size kernels of a plan made with PLAN_NO_DECIMATE."""
    import torch
    from upmix_b200 import _native
    sr = 48000
    df = sr / n_fft
    f_low, f_high = 3 * df, top_bin * df
    e = ce.MultiBandExtractorAccu(n_fft, 0.75, ce.make_blackman_harris, f_low, f_high, sr, "hard_zero", 0.0, 0.0)
    b = uo.make_band(n_fft, 0.75, uo.blackman_harris, f_low, f_high, sr, "hard_zero", 0.0, 0.0)
    assert int(np.nonzero(e.band_gain())[0].max()) == top_bin
    n = 2 * n_fft + 4321
    L, R = uo.synth_stereo(n, 200 + top_bin, stress=True)
    ref = uo.process_band_batched(b, L.astype(np.float64), R.astype(np.float64))
    got = e.process_all_blocks(L, R)
    peak = float(max(np.abs(L).max(), np.abs(R).max()))
    assert_parity(ref, got, peak, what=f"N={n_fft} top_bin={top_bin}")
    dl, dr = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    full = ce.plan_for([e], _native.OUT_LSCRS, _native.PLAN_NO_DECIMATE).process(dl, dr)
    for g, f in zip(got, full):
        assert float(np.max(np.abs(g - f.cpu().numpy()))) < 2e-7


def test_decimated_waves_tracks_and_modes_are_consistent(ce, monkeypatch):
    """The decimated kernels under every scheduling the library uses: a scratch cap that forces waves of hops and
    groups of tracks == one wave (bit for bit); accumulating == storing + staged band sum (bit for bit); merged
    bands and the fold-down against the oracle."""
    import torch
    sr = 48000
    edges = [0, 30, 120, 480, 1920, 7680]                       # two 65536-point bands (merged), 16384, 4096: all decimated
    ext = quiet(ce.chain_bands, edges, 0.75, ce.make_blackman_harris, sr, "raised_cosine")
    bands = uo.chain(edges, 0.75, uo.blackman_harris, sr)
    n = 6 * sr + 1234
    tracks = [uo.synth_stereo(n, 300 + t, stress=(t == 0)) for t in range(3)]
    Ls = torch.from_numpy(np.stack([t[0] for t in tracks])).cuda()
    Rs = torch.from_numpy(np.stack([t[1] for t in tracks])).cuda()
    base = [o.clone() for o in ce.extract_center_left_right_multi_band_in_memory(Ls, Rs, sr, ext)]
    ref = uo.upmix_multiband(bands, tracks[0][0].astype(np.float64), tracks[0][1].astype(np.float64))
    assert_parity(ref, [o[0].cpu().numpy() for o in base], 0.5, what="6 bands, decimated")
    from upmix_b200 import _native
    from upmix_b200.center_extraction import _PLAN_CACHE
    for env in ({"UPMIX_DEC_WS_MB": "1"}, {"UPMIX_DIRECT_MIN": "1"}, {"UPMIX_DIRECT_MIN": "1", "UPMIX_DEC_WS_MB": "1"},
                {"UPMIX_DIRECT_MIN": str(1 << 40)}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        got = ce.extract_center_left_right_multi_band_in_memory(Ls, Rs, sr, ext)
        for a, b in zip(base, got):
            assert torch.equal(a, b), env
        for k in env:
            monkeypatch.delenv(k)
    # fold-down (centre folded per bin by the mask kernel) against Ls + 0.5 C / Rs + 0.5 C
    fl, fr = ce.extract_stereo_fold_down(Ls[0].contiguous(), Rs[0].contiguous(), sr, ext)
    want_l, want_r = ref[1].astype(np.float64) + 0.5 * ref[0], ref[2].astype(np.float64) + 0.5 * ref[0]
    assert_parity((want_l, want_r), (fl.cpu().numpy(), fr.cpu().numpy()), 0.5, names=("outL", "outR"), what="fold-down, decimated")


def test_cfg5_sampled_tracks_against_oracle(ce):
    """BASELINE configs[4] at its stated track length: four of the 512 five-minute tracks (seeds 1000 + i), main.py's
    default six bands, processed together as one wave of a batch; parity on windows at the start, an interior cut and
    the end of each track."""
    import torch
    sr = 48000
    edges = [0, 30, 120, 480, 1920, 7680]
    ext = quiet(ce.chain_bands, edges, 0.75, ce.make_blackman_harris, sr, "raised_cosine")
    bands = uo.chain(edges, 0.75, uo.blackman_harris, sr)
    n = 300 * sr
    picks = [0, 137, 300, 511]
    sig = [uo.synth_stereo(n, 1000 + i) for i in picks]
    Ls = torch.from_numpy(np.stack([s[0] for s in sig])).cuda()
    Rs = torch.from_numpy(np.stack([s[1] for s in sig])).cuda()
    out = [o.cpu().numpy() for o in ce.extract_center_left_right_multi_band_in_memory(Ls, Rs, sr, ext)]
    margin = 65536
    for t, (L, R) in enumerate(sig):
        for a, b in ((0, 2 * sr), (n // 2 - sr // 2, n // 2 + sr // 2), (n - 2 * sr, n)):
            lo, hi = max(0, a - margin), min(n, b + margin)
            lo -= lo % 16384
            ref = uo.upmix_multiband(bands, L[lo:hi].astype(np.float64), R[lo:hi].astype(np.float64))
            ia = a if lo == 0 else max(a, lo + 49152)
            ib = b if hi == n else min(b, hi - 65536)
            assert_parity([x[ia - lo:ib - lo] for x in ref], [x[t, ia:ib] for x in out], 0.5, what=f"track {picks[t]} [{a},{b})")


def test_random_crossovers_on_device(ce):
    """Seeded random crossover sets, both crossover modes, 44.1 / 48 / 96 kHz (the GPU twin of
    tests/test_oracle.py::test_live_against_reference_random_crossovers, which pins the oracle to the live reference)."""
    rng = np.random.default_rng(11)
    for trial in range(5):
        sr = int(rng.choice([44100, 48000, 96000]))
        edges = [0.0] + sorted(float(x) for x in rng.uniform(60, 9000, size=int(rng.integers(1, 5))))
        mode = ["raised_cosine", "hard_zero"][trial % 2]
        max_block = int(rng.choice([8192, 16384, 65536]))
        ext = quiet(ce.chain_bands, edges, 0.75, ce.make_blackman_harris, sr, mode, max_block_size=max_block)
        bands = uo.chain(edges, 0.75, uo.blackman_harris, sr, mode=mode, max_block=max_block)
        assert [e.block_size for e in ext] == [b.n_fft for b in bands]
        L, R = uo.synth_stereo(3 * max_block + 20011, 100 + trial, sr=sr, stress=True)
        ref = uo.upmix_multiband(bands, L.astype(np.float64), R.astype(np.float64))
        got = ce.extract_center_left_right_multi_band_in_memory(L, R, sr, ext)
        rep = assert_parity(ref, got, float(max(np.abs(L).max(), np.abs(R).max())), what=f"trial {trial}: sr={sr} edges={edges} {mode}")
        print(trial, sr, [e.block_size for e in ext], mode, [(nm, round(s, 1)) for nm, s, _ in rep])


def test_near_silence_and_digital_zero(ce):
    """Segments at 1e-9 and 1e-12 of full scale and exact zeros: the centre factor's m/(m+EPS) regime and the
    flush-to-zero square root / reciprocal of the mask (fft_device.cuh: sqrt_approx / rcp_approx)."""
    sr = 48000
    edges = [0, 200, 2000]
    ext = quiet(ce.chain_bands, edges, 0.75, ce.make_blackman_harris, sr, "raised_cosine", max_block_size=16384)
    bands = uo.chain(edges, 0.75, uo.blackman_harris, sr, max_block=16384)
    n = 5 * sr
    L, R = uo.synth_stereo(n, 77)
    s = sr
    L[s:2 * s] *= np.float32(1e-9)
    R[s:2 * s] *= np.float32(1e-9)
    L[2 * s:3 * s] *= np.float32(1e-12)
    R[2 * s:3 * s] *= np.float32(1e-12)
    L[3 * s:4 * s] = 0.0
    R[3 * s:4 * s] = 0.0
    ref = uo.upmix_multiband(bands, L.astype(np.float64), R.astype(np.float64))
    got = ce.extract_center_left_right_multi_band_in_memory(L, R, sr, ext)
    assert_parity(ref, got, 0.5, what="near-silence")
    # inside the quiet segments the error is judged against the quiet signal itself
    for a, scale in ((s + 16384, 1e-9), (2 * s + 16384, 1e-12)):      # frames that lie wholly inside the quiet second
        b = a + s - 2 * 16384
        for r_, g_ in zip(ref, got):
            assert float(np.max(np.abs(r_[a:b] - g_[a:b]))) <= 1e-4 * scale
    # an all-zero input gives exact zeros
    z = np.zeros(2 * sr, np.float32)
    for o in ce.extract_center_left_right_multi_band_in_memory(z, z, sr, ext):
        assert not np.any(o)


def test_plan_follows_in_place_edits_of_the_windows(ce):
    """The reference's analysis_window / synthesis_window are plain arrays (CE:257-258): editing them in place must
    rebuild the device tables, not reuse a stale plan."""
    sr = 48000
    e = ce.MultiBandExtractorAccu(1024, 0.75, ce.make_blackman_harris, 500.0, 4000.0, sr, "raised_cosine", 100.0, 500.0)
    L, R = uo.synth_stereo(20000, 8)
    first = e.process_all_blocks(L, R)
    e.synthesis_window[:] *= np.float32(0.5)
    second = e.process_all_blocks(L, R)
    assert np.allclose(second[0], 0.5 * first[0], atol=1e-7) and not np.array_equal(second[0], first[0])
    ext = [e]
    third = ce.extract_center_left_right_multi_band_in_memory(L, R, sr, ext)
    e.analysis_window[:] *= np.float32(2.0)
    fourth = ce.extract_center_left_right_multi_band_in_memory(L, R, sr, ext)
    assert np.array_equal(third[0], second[0]) and not np.array_equal(fourth[0], third[0])


# ---------------------------------------------------------------------------------------------------
# frame-batched kernel of dense bands (upmix_fb.cuh): 16 frames per tile, overlap-add across lanes
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_fft", [256, 512, 1024])
def test_frame_batched_dense_band(ce, n_fft, monkeypatch):
    """Dense band (pass band up to Nyquist) through the frame-batched kernel: against the oracle for lengths around the
    tile boundaries (a run is 16 k - 3 hops), against the one-frame kernel, time shards and multi-track batches bit for
    bit, accumulating == storing + staged sum, fold-down."""
    import torch
    from upmix_b200 import _native
    monkeypatch.setenv("UPMIX_FB_MAX_N", "1024")                # (1024 points take the one-frame kernel by default)
    sr = 48000
    H = n_fft // 4
    f_low = 32.0 * sr / n_fft
    e = ce.MultiBandExtractorAccu(n_fft, 0.75, ce.make_blackman_harris, f_low, sr / 2, sr, "raised_cosine", f_low / 4, 0.0)
    b = uo.make_band(n_fft, 0.75, uo.blackman_harris, f_low, sr / 2, sr, "raised_cosine", f_low / 4, 0.0)
    for n in (H - 1, 13 * H, 13 * H + 1, 29 * H + 7, 100 * H + 321):
        L, R = uo.synth_stereo(n, n_fft + n, stress=n > 4 * n_fft)
        ref = uo.process_band_batched(b, L.astype(np.float64), R.astype(np.float64))
        got = e.process_all_blocks(L, R)
        assert_parity(ref, got, float(max(np.abs(L).max(), np.abs(R).max(), 1e-3)), what=f"N={n_fft} n={n}")
    n = 300 * H + 77
    L, R = uo.synth_stereo(n, 5, stress=True)
    dl, dr = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    plan = ce.plan_for([e])
    whole = [t.clone() for t in plan.process(dl, dr)]
    one = ce.plan_for([e], _native.OUT_LSCRS, _native.PLAN_NO_BATCH).process(dl, dr)
    for w, o in zip(whole, one):
        assert float((w - o).abs().max()) < 2e-7
    halo = plan.halo
    for a_, b_ in ((0, 37 * H), (37 * H, 37 * H + 5), (37 * H + 5, 200 * H + 3), (200 * H + 3, n)):
        lo, hi = max(0, a_ - halo), min(n, b_ + halo)
        seg = plan.process_segment(dl[lo:hi].contiguous(), dr[lo:hi].contiguous(), lo, n, a_, b_)
        for w, s_ in zip(whole, seg):
            assert torch.equal(w[a_:b_], s_), (a_, b_)
    batch = plan.process(torch.stack([dl, dr, dl.flip(0)]), torch.stack([dr, dl, dr * 0.5]))
    for w, bt in zip(whole, batch):
        assert torch.equal(w, bt[0])
    # two bands (the top one dense): the second pipeline accumulates onto the first one's output
    ext = quiet(ce.chain_bands, [0, 32.0 * sr / n_fft], 0.75, ce.make_blackman_harris, sr, "raised_cosine", max_block_size=4 * n_fft)
    bands = uo.chain([0, 32.0 * sr / n_fft], 0.75, uo.blackman_harris, sr, max_block=4 * n_fft)
    assert ext[-1].block_size == n_fft
    ref = uo.upmix_multiband(bands, L.astype(np.float64), R.astype(np.float64))
    monkeypatch.setenv("UPMIX_DIRECT_MIN", "1")
    direct = [t.clone() for t in ce.extract_center_left_right_multi_band_in_memory(dl, dr, sr, ext)]
    monkeypatch.setenv("UPMIX_DIRECT_MIN", str(1 << 40))
    staged = ce.extract_center_left_right_multi_band_in_memory(dl, dr, sr, ext)
    monkeypatch.delenv("UPMIX_DIRECT_MIN")
    assert_parity(ref, [t.cpu().numpy() for t in direct], 0.5, what=f"2 bands, top N={n_fft}")
    for d_, s_ in zip(direct, staged):
        assert torch.equal(d_, s_)
    fl, fr = ce.extract_stereo_fold_down(dl, dr, sr, ext)
    assert_parity((ref[1].astype(np.float64) + 0.5 * ref[0], ref[2].astype(np.float64) + 0.5 * ref[0]),
                  (fl.cpu().numpy(), fr.cpu().numpy()), 0.5, names=("outL", "outR"), what="fold-down")
