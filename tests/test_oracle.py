"""CPU: the oracle (oracle/upmix_oracle.py) against the golden fixtures made from the unmodified
reference (tests/golden/make_golden.py) and, where the reference tree is mounted, against it live."""
import contextlib
import io
import os

import numpy as np
import pytest

from oracle import upmix_oracle as uo
from oracle.ref_loader import load_reference, reference_available


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _bands_for(g, mode="raised_cosine"):
    return uo.chain(list(g["edges"]), 0.75, uo.blackman_harris, float(g["sr"]), mode, max_block=int(g["max_block"]))


@pytest.mark.parametrize("name,mode", [("cfg1_default6.npz", "raised_cosine"), ("cfg2_3band.npz", "raised_cosine"),
                                       ("cfg4_8band96k.npz", "raised_cosine"), ("hardzero_4band.npz", "hard_zero")])
def test_multiband_bit_exact_vs_reference_fixture(golden_dir, name, mode):
    g = _load(golden_dir, name)
    bands = _bands_for(g, mode)
    assert [b.n_fft for b in bands] == list(g["sizes"])
    L, R = g["in_L"].astype(np.float64), g["in_R"].astype(np.float64)
    c, l, r = uo.upmix_multiband(bands, L, R, batched=True)
    assert np.array_equal(c, g["ref_C"]) and np.array_equal(l, g["ref_Ls"]) and np.array_equal(r, g["ref_Rs"])


def test_frame_loop_equals_batched(golden_dir):
    g = _load(golden_dir, "cfg2_3band.npz")
    bands = _bands_for(g)
    L, R = g["in_L"][:40000].astype(np.float64), g["in_R"][:40000].astype(np.float64)
    a = uo.upmix_multiband(bands, L, R, batched=False)
    b = uo.upmix_multiband(bands, L, R, batched=True)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    small = bands[2]
    x = uo.process_band_batched(small, L, R, chunk_frames=7)
    y = uo.process_band_frames(small, L, R)
    for p, q in zip(x, y):
        assert np.array_equal(p, q)


def test_tables_vs_reference_fixture(golden_dir):
    t = _load(golden_dir, "tables.npz")
    mk = dict(blackman_harris=uo.blackman_harris, sqrt_hann=uo.sqrt_hann, hann=uo.hann, blackman=uo.blackman,
              hamming=uo.hamming, rect=uo.rect)
    for name, fn in mk.items():
        for n in (64, 256, 1024):
            w = fn(n)
            assert np.array_equal(w, t[f"win_{name}_{n}"])
            for ov in (50, 75):
                with np.errstate(all="ignore"):
                    s = uo.wola_synthesis_window(w, ov / 100.0)
                assert np.array_equal(s, t[f"syn_{name}_{n}_{ov}"], equal_nan=True), (name, n, ov)
    bands = uo.chain([0, 30, 120, 480, 1920, 7680], 0.75, uo.blackman_harris, 48000)
    for i, b in enumerate(bands):
        assert np.array_equal(b.gain, t[f"gain_default6_{i}"])
        if f"syn_default6_{i}" in t:
            assert np.array_equal(b.syn, t[f"syn_default6_{i}"])
    for f, sr, n in t["sizes_rule"]:
        assert uo.block_size_for_low_freq(float(f), float(sr)) == int(n)
    for f, n, b in t["bins"]:
        assert uo.freq_to_bin(float(f), 48000, int(n)) == int(b)


def test_single_band_variants(golden_dir):
    g = _load(golden_dir, "single_band.npz")
    L, R = g["in_L"].astype(np.float64), g["in_R"].astype(np.float64)
    b = uo.make_band(512, 0.5, uo.sqrt_hann, 300.0, 5000.0, 48000, "raised_cosine", 100.0, 800.0)
    for k, v in zip(("C", "Ls", "Rs"), uo.process_band_frames(b, L, R)):
        assert np.array_equal(v, g[f"ref_sqrt_hann50_{k}"], equal_nan=True)
    for k, v in zip(("C", "Ls", "Rs"), uo.process_band_batched(b, L, R)):
        assert np.array_equal(v, g[f"ref_sqrt_hann50_{k}"], equal_nan=True)
    b = uo.make_band(256, 0.75, uo.hann, 1000.0, 24000.0, 48000, "bogus_mode")
    for k, v in zip(("C", "Ls", "Rs"), uo.process_band_batched(b, L, R)):
        assert np.array_equal(v, g[f"ref_hann75_{k}"], equal_nan=True)


def test_bela_mode_vs_compiled_reference_fixture(golden_dir):
    """bela/upmix.cpp (compiled unmodified against the shim) == prototype with Bela parameters,
    delayed 3*hw.  float32 C++ against float64 numpy: 120 dB is the bar here."""
    for hw in (2048, 512):
        g = _load(golden_dir, f"bela_hw{hw}.npz")
        bands = uo.bela_chain([0, 500, 2000, 8000, 24000], 48000, hw)
        l, r = uo.bela_offline(bands, g["in_L"], g["in_R"], hw)
        assert uo.snr_db(g["ref_outL"], l) > 120 and uo.snr_db(g["ref_outR"], r) > 120


def test_known_answer_properties():
    """Properties that follow from the reference code."""
    sr = 48000
    bands = uo.chain([0, 200, 2000], 0.75, uo.blackman_harris, sr, max_block=8192)
    L, R = uo.synth_stereo(30000, 5)
    L64, R64 = L.astype(np.float64), R.astype(np.float64)
    # WOLA: overlapped ana*syn sums to 1
    b = bands[1]
    w = (b.ana.astype(np.float64) * b.syn).reshape(4, -1).sum(axis=0)
    assert np.max(np.abs(w - 1)) < 1e-6
    # identical channels: everything goes to the centre
    c, l, r = uo.process_band_batched(b, L64, L64)
    assert np.max(np.abs(l)) < 1e-6 and np.max(np.abs(r)) < 1e-6 and np.max(np.abs(c)) > 1e-2
    # hard-panned: nothing goes to the centre
    c, l, r = uo.process_band_batched(b, L64, np.zeros_like(L64))
    assert np.max(np.abs(c)) == 0 and np.max(np.abs(r)) == 0 and np.max(np.abs(l)) > 1e-2
    # Ls + C is the mask-independent WOLA resynthesis of the band-limited L (linearity)
    c, l, r = uo.process_band_batched(b, L64, R64)
    c2, l2, r2 = uo.process_band_batched(b, L64, 0.3 * R64[::-1].copy())
    assert np.max(np.abs((l + c) - (l2 + c2))) < 1e-6
    # empty and shorter-than-a-hop inputs
    assert all(len(x) == 0 for x in uo.process_band_batched(b, L64[:0], R64[:0]))
    for n in (1, 17, b.hop - 1, b.hop + 1):
        x = uo.process_band_batched(b, L64[:n], R64[:n])
        y = uo.process_band_frames(b, L64[:n], R64[:n])
        assert all(np.array_equal(p, q) for p, q in zip(x, y))


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")
def test_live_against_reference_random_crossovers():
    ce = load_reference()
    rng = np.random.default_rng(11)
    for trial in range(3):
        sr = int(rng.choice([44100, 48000]))
        edges = [0.0] + sorted(float(x) for x in rng.uniform(150, 9000, size=int(rng.integers(1, 4))))
        mode = ["raised_cosine", "hard_zero"][trial % 2]
        with contextlib.redirect_stdout(io.StringIO()):
            ext = ce.chain_bands(edges, 0.75, ce.make_blackman_harris, sr, mode)
        for e in ext:       # keep the test fast: shrink the f_low=0 band
            if e.block_size > 8192:
                e.__init__(8192, 0.75, ce.make_blackman_harris, e.f_low, e.f_high, sr, mode, e.xover_width_low_hz,
                           e.xover_width_high_hz)
        L, R = uo.synth_stereo(20011, 100 + trial, sr=sr, stress=True)
        L64, R64 = L.astype(np.float64), R.astype(np.float64)
        ref = ce.extract_center_left_right_multi_band_in_memory(L64, R64, sr, ext)
        bands = [uo.make_band(e.block_size, 0.75, uo.blackman_harris, e.f_low, e.f_high, sr, mode,
                              e.xover_width_low_hz, e.xover_width_high_hz) for e in ext]
        got = uo.upmix_multiband(bands, L64, R64)
        for a, b in zip(ref, got):
            assert np.array_equal(a, b)
