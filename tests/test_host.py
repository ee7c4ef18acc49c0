"""CPU: host-side plan construction of the product (upmix_b200.center_extraction) against the golden
fixtures of the reference -- tables must be bit-identical -- and the C-ABI library's exports."""
import contextlib
import ctypes
import io
import os
import re

import numpy as np
import pytest

import upmix_b200.center_extraction as ce
from upmix_b200 import _native, bela
from oracle import upmix_oracle as uo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def test_windows_and_synthesis_windows_bit_identical(golden_dir):
    t = np.load(os.path.join(golden_dir, "tables.npz"))
    for name in ("blackman_harris", "sqrt_hann", "hann", "blackman", "hamming", "rect"):
        for n in (64, 256, 1024):
            w = getattr(ce, "make_" + name)(n)
            assert w.dtype == np.float32 and np.array_equal(w, t[f"win_{name}_{n}"])
            for ov in (50, 75):
                with np.errstate(all="ignore"):
                    s = ce.design_wola_synthesis_window(w, ov / 100.0)
                assert s.dtype == np.float32 and np.array_equal(s, t[f"syn_{name}_{n}_{ov}"], equal_nan=True)


def test_chain_bands_tables_bit_identical(golden_dir, capsys):
    t = np.load(os.path.join(golden_dir, "tables.npz"))
    ext = ce.chain_bands([0, 30, 120, 480, 1920, 7680], 0.75, ce.make_blackman_harris, 48000, "raised_cosine")
    printed = capsys.readouterr().out.strip().splitlines()
    assert len(printed) == 6 and printed[1].startswith("[Band 2] f_low=30.0 Hz, f_high=120.0 Hz, block_size=65536, "
                                                       "xover_low=7.5 Hz, xover_high=30.0 Hz")
    assert [e.block_size for e in ext] == [65536, 65536, 16384, 4096, 1024, 256]
    assert [e.hop_size for e in ext] == [16384, 16384, 4096, 1024, 256, 64]
    for i, e in enumerate(ext):
        assert np.array_equal(e.band_gain(), t[f"gain_default6_{i}"])
        if f"syn_default6_{i}" in t:
            assert np.array_equal(e.synthesis_window, t[f"syn_default6_{i}"])
        n, h, ana, syn, gain = e.plan_tables()
        assert (n, h) == (e.block_size, e.hop_size) and ana.dtype == syn.dtype == gain.dtype == np.float32
    for f, sr, n in t["sizes_rule"]:
        assert ce.compute_block_size_for_low_freq(float(f), float(sr)) == int(n)
    for f, n, b in t["bins"]:
        assert ce.freq_to_bin(float(f), 48000, int(n)) == int(b)


def test_host_tables_match_oracle_on_random_bands():
    rng = np.random.default_rng(5)
    for _ in range(20):
        sr = float(rng.choice([44100, 48000, 96000]))
        n = int(2 ** rng.integers(6, 14))
        f_low = float(rng.choice([0.0, rng.uniform(10, 4000)]))
        f_high = float(rng.choice([sr / 2, rng.uniform(f_low + 1, sr / 2)]))
        mode = str(rng.choice(["raised_cosine", "hard_zero", "other"]))
        wl, wh = float(rng.uniform(0, 500)), float(rng.uniform(0, 3000))
        e = ce.MultiBandExtractorAccu(n, 0.75, ce.make_blackman_harris, f_low, f_high, sr, mode, wl, wh)
        b = uo.make_band(n, 0.75, uo.blackman_harris, f_low, f_high, sr, mode, wl, wh)
        assert np.array_equal(e.band_gain(), b.gain)
        assert np.array_equal(e.synthesis_window, b.syn) and np.array_equal(e.analysis_window, b.ana)


def test_utilities_and_errors():
    assert [ce.next_power_of_2(x) for x in (-3, 0, 1, 2, 3, 4, 5, 1023, 1024, 1025)] == [1, 1, 1, 2, 4, 4, 8, 1024, 1024, 2048]
    assert ce.hp_freq_to_crossover_width(120.0) == 30.0
    assert ce.compute_block_size_for_low_freq(0.0, 48000, max_block_size=8192) == 8192
    assert ce.compute_block_size_for_low_freq(100.0, 96000, 8192, 32) == 8192
    assert ce.compute_block_size_for_low_freq(100.0, 96000, 2 ** 16, 16) == 16384
    with pytest.raises(ValueError):
        ce.design_wola_synthesis_window(np.ones(8, np.float32), 0.95)
    with pytest.raises(ValueError):
        ce.MultiBandExtractorAccu(64, 0.999, ce.make_hann, 0.0, 100.0, 48000)
    x = np.random.default_rng(0).standard_normal(64)
    w = ce.make_hann(64)
    assert np.allclose(ce.forward_stft(x, w), np.fft.rfft(x * w))
    assert ce.inverse_stft(np.fft.rfft(x), w).dtype == np.float32
    ext = quiet(ce.chain_bands, [0, 1000], 0.75, ce.make_blackman_harris, 48000, max_block_size=4096, threshold_factor=16,
                xo_fraction=0.5)
    assert [e.block_size for e in ext] == [4096, 1024] and ext[0].xover_width_high_hz == 500.0


def test_bela_tables_match_oracle():
    bands = quiet(bela.bela_chain_bands, [0.0, 500.0, 2000.0, 8000.0, 24000.0], 48000.0, 2048)
    ref = uo.bela_chain([0, 500, 2000, 8000, 24000], 48000, 2048)
    assert [b.block_size for b in bands] == [8192, 4096, 1024, 256] == [r.n_fft for r in ref]
    for b, r in zip(bands, ref):
        assert np.array_equal(b.band_gain(), r.gain)
        assert np.max(np.abs(b.analysis_window - r.ana)) < 1e-6        # float32 cosf vs float64 cos
        assert np.array_equal(b.analysis_window, b.synthesis_window)


def test_library_loads_and_exports_every_declared_symbol():
    """include/upmix_b200.h is the boundary: every function it declares must be exported (no compute
    calls here -- there is no GPU in this container)."""
    header = open(os.path.join(ROOT, "include", "upmix_b200.h")).read()
    declared = set(re.findall(r"\b(upmix_[a-z_0-9]+)\s*\(", header))
    assert declared and declared == set(_native.EXPORTS)
    assert os.path.isfile(_native.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib2 = _native.load_library()
    assert lib2.upmix_version() >= 1000
    assert lib2.upmix_last_error() is not None
    # argument validation happens before any CUDA call
    assert lib2.upmix_plan_create(0, None, 0, 0, ctypes.byref(ctypes.c_void_p())) == -1
    assert b"band" in lib2.upmix_last_error()


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    e = ce.MultiBandExtractorAccu(256, 0.75, ce.make_hann, 100.0, 1000.0, 48000)
    with pytest.raises(_native.UpmixNativeError):
        e.process_all_blocks(np.zeros(1000), np.zeros(1000))
    with pytest.raises(_native.UpmixNativeError):
        ce.extract_center_left_right_multi_band_in_memory(np.zeros(1000), np.zeros(1000), 48000, [e])


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "upmix_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"


def test_fir_design_matches_reference_fixture(golden_dir):
    """design_lr4_hp_fir / design_lr4_lp_fir (filter_design.py:25-52) are host code: bit-identical taps
    when scipy is the version the fixture was made with, equal to float32 rounding otherwise."""
    import scipy
    from upmix_b200 import filter_design as fd
    g = np.load(os.path.join(golden_dir, "fir.npz"))
    got = {"ref_hp": fd.design_lr4_hp_fir(48000, 180.0, 1025), "ref_lp": fd.design_lr4_lp_fir(48000, 180.0, 1025),
           "ref_hp_short": fd.design_lr4_hp_fir(44100, 2000.0, 129), "ref_pass": fd.design_lr4_lp_fir(48000, 0.0)}
    for k, v in got.items():
        assert v.dtype == np.float32 and v.shape == g[k].shape
        if scipy.__version__ == str(g["scipy_version"]):
            assert np.array_equal(v, g[k]), k
        else:
            assert np.allclose(v, g[k], rtol=0, atol=1e-7), k
    assert np.array_equal(fd.design_lr4_hp_fir(48000, -1.0), np.array([1.0], dtype=np.float32))


def test_bela_rules_use_float32_like_the_cpp():
    """bela/upmix.cpp evaluates freqToBin and computeBlockSizeForLowFreq in float (BU:45-54, 498-506): the Python
    mirror must round where the C++ rounds.  Checked against a float32 restatement over random edges, and on cases
    where float64 arithmetic would land on the other side."""
    from upmix_b200 import bela
    rng = np.random.default_rng(5)
    f = np.float32
    for _ in range(2000):
        freq, sr = float(rng.uniform(1, 24000)), float(rng.choice([44100.0, 48000.0, 96000.0]))
        n = int(2 ** rng.integers(8, 14))
        b = f(f(f(freq) * f(n)) / f(sr))
        b = min(max(b, f(0)), f(n // 2))
        assert bela.freq_to_bin_bela(freq, sr, n) == int(np.floor(np.float64(b) + 0.5))
        hw = int(2 ** rng.integers(6, 12))
        thr = f(f(f(sr) * f(32.0)) / f(freq))
        want = min(bela.ce.next_power_of_2(int(np.ceil(np.float64(thr)))), hw * 4)
        assert bela.compute_block_size_bela(freq, sr, hw, 32.0) == want
    assert bela.compute_block_size_bela(0.0, 48000.0, 2048) == 8192
    # 48000*32/f_low exactly a power of two in float32 but not in float64 -> the size flips if float64 is used
    f_low = float(np.nextafter(np.float32(1536000.0 / 1024.0), np.float32(0)))       # just below 1500 Hz
    assert bela.compute_block_size_bela(f_low, 48000.0, 2048) in (1024, 2048)


def test_host_conversion_loops_match_numpy():
    """upmix_simd.cpp (AVX2 / non-temporal stores behind a CPU check): float64 / float32 interleaved -> planar float32,
    strided gathers and the copy-out loop give exactly numpy's astype(float32), for ragged lengths and misaligned
    destinations."""
    import ctypes
    from upmix_b200 import _native
    lib = _native.load_library()
    i64, vp = ctypes.c_int64, ctypes.c_void_p
    for name, at in (("upmix_host_pair_f64", [vp, i64, vp, vp]), ("upmix_host_pair_f32", [vp, i64, vp, vp]),
                     ("upmix_host_gather_f64", [vp, i64, i64, vp]), ("upmix_host_gather_f32", [vp, i64, i64, vp]),
                     ("upmix_host_copy", [vp, vp, i64])):
        getattr(lib, name).argtypes = at
        getattr(lib, name).restype = None
    rng = np.random.default_rng(5)
    for n in (0, 1, 7, 8, 33, 1000, 4099):
        for off_l, off_r in ((0, 0), (3, 3), (1, 6)):
            w64 = rng.standard_normal((n, 2)) * 1e3
            w32 = w64.astype(np.float32)
            for src, fn in ((w64, lib.upmix_host_pair_f64), (w32, lib.upmix_host_pair_f32)):
                bl, br = np.full(n + 16, 7.0, np.float32), np.full(n + 16, 7.0, np.float32)
                dl, dr = bl[off_l:off_l + n], br[off_r:off_r + n]
                fn(src.ctypes.data, n, dl.ctypes.data, dr.ctypes.data)
                assert np.array_equal(dl, src[:, 0].astype(np.float32)) and np.array_equal(dr, src[:, 1].astype(np.float32))
                assert (bl[:off_l] == 7).all() and (bl[off_l + n:] == 7).all() and (br[:off_r] == 7).all() and (br[off_r + n:] == 7).all()
            for stride in (1, 3):
                x64 = rng.standard_normal(n * stride + 1)
                x32 = x64.astype(np.float32)
                for src, fn in ((x64, lib.upmix_host_gather_f64), (x32, lib.upmix_host_gather_f32)):
                    buf = np.full(n + 16, 7.0, np.float32)
                    d = buf[off_l:off_l + n]
                    fn(src.ctypes.data, stride, n, d.ctypes.data)
                    assert np.array_equal(d, src[:n * stride:stride].astype(np.float32))
                    assert (buf[:off_l] == 7).all() and (buf[off_l + n:] == 7).all()
    big = rng.standard_normal(70001).astype(np.float32)
    for off in (0, 1, 5):
        buf = np.full(big.size + 16, 7.0, np.float32)
        lib.upmix_host_copy(buf[off:].ctypes.data, big.ctypes.data, big.size)
        assert np.array_equal(buf[off:off + big.size], big) and (buf[:off] == 7).all() and (buf[off + big.size:] == 7).all()
