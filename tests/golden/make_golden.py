#!/usr/bin/env python3
"""Generate the golden fixtures in this directory FROM THE UNMODIFIED REFERENCE.

Run in the development container only (needs /root/reference mounted and, for the Bela fixture,
`make -C oracle` to have built oracle/_ref/libbela_ref.so from bela/upmix.cpp):

    python tests/golden/make_golden.py

Every array named `ref_*` below is an output of the reference's own code
(python-prototype/center_extraction.py via oracle/ref_loader.py, or bela/upmix.cpp via the shim);
inputs are stored next to them so the tests do not depend on a random-number stream.
numpy's version is recorded because the reference's arithmetic depends on it.
"""
import contextlib
import ctypes
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import upmix_oracle as uo            # noqa: E402  (only for synth_stereo)
from oracle.ref_loader import load_reference     # noqa: E402

ce = load_reference()


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def run_multiband(edges, sr, n, seed, max_block=None, mode="raised_cosine"):
    L, R = uo.synth_stereo(n, seed, sr=sr, stress=True)
    if max_block is None:
        ext = quiet(ce.chain_bands, edges, 0.75, ce.make_blackman_harris, sr, mode)
    else:
        # chain_bands does not expose max_block_size (CE:555); rebuild its loop (CE:546-578) with
        # the reference's own pieces for the "STFT sizes up to 8192" configuration.
        e = list(edges) + ([sr / 2.0] if edges[-1] < sr / 2.0 else [])
        ext, prev = [], 0.0
        for lo, hi in zip(e[:-1], e[1:]):
            nfft = ce.compute_block_size_for_low_freq(lo, sr, max_block_size=max_block)
            w = ce.hp_freq_to_crossover_width(hi)
            ext.append(ce.MultiBandExtractorAccu(nfft, 0.75, ce.make_blackman_harris, lo, hi, sr,
                                                 mode, prev, w))
            prev = w
    c, l, r = ce.extract_center_left_right_multi_band_in_memory(
        L.astype(np.float64), R.astype(np.float64), sr, ext)
    return dict(in_L=L, in_R=R, ref_C=c, ref_Ls=l, ref_Rs=r,
                sizes=np.array([x.block_size for x in ext]), edges=np.array(edges, dtype=np.float64),
                sr=np.array(sr), max_block=np.array(max_block or 65536))


def main():
    meta = dict(numpy_version=np.array(np.__version__))

    # cfg 1 shape: main.py's default crossovers (MP:62), 48 kHz, Ls/C/Rs.  2.1 s so that the
    # 65536-point bands see several frames; odd length exercises the ragged tail.
    np.savez_compressed(os.path.join(HERE, "cfg1_default6.npz"), **meta,
                        **run_multiband([0, 30, 120, 480, 1920, 7680], 48000, 100003, 0))
    # cfg 2 shape: 3 bands
    np.savez_compressed(os.path.join(HERE, "cfg2_3band.npz"), **meta,
                        **run_multiband([0, 200, 2000], 48000, 90001, 1))
    # cfg 4 shape: 8 bands, 96 kHz, sizes clamped to 8192
    np.savez_compressed(os.path.join(HERE, "cfg4_8band96k.npz"), **meta,
                        **run_multiband([0, 100, 200, 400, 800, 1600, 3200, 6400], 96000, 50000, 3,
                                        max_block=8192))
    # hard-zero crossover mode on the demo's band set (CE:675)
    np.savez_compressed(os.path.join(HERE, "hardzero_4band.npz"), **meta,
                        **run_multiband([0.0, 40.0, 200.0, 2000.0], 48000, 70001, 4, mode="hard_zero",
                                        max_block=16384))

    # tables: windows, synthesis windows, gains (all from the reference's own functions)
    tab = dict(meta)
    for name in ("blackman_harris", "sqrt_hann", "hann", "blackman", "hamming", "rect"):
        for n in (64, 256, 1024):
            w = getattr(ce, "make_" + name)(n)
            tab[f"win_{name}_{n}"] = w
            for ov in (0.5, 0.75):
                with np.errstate(all="ignore"):
                    tab[f"syn_{name}_{n}_{int(ov * 100)}"] = ce.design_wola_synthesis_window(w, ov)
    ext = quiet(ce.chain_bands, [0, 30, 120, 480, 1920, 7680], 0.75, ce.make_blackman_harris, 48000,
                "raised_cosine")
    for i, e in enumerate(ext):
        a = np.ones(e.block_size // 2 + 1, dtype=np.complex128)
        e._band_limit(a, a.copy())
        tab[f"gain_default6_{i}"] = a.real.copy()
        if e.block_size <= 16384:
            tab[f"syn_default6_{i}"] = e.synthesis_window
    tab["sizes_rule"] = np.array([[f, sr, ce.compute_block_size_for_low_freq(f, sr)]
                                  for sr in (44100, 48000, 96000)
                                  for f in (0, 20, 30, 55.5, 120, 480, 1000, 1920, 7680, 20000)])
    tab["bins"] = np.array([[f, n, ce.freq_to_bin(f, 48000, n)]
                            for n in (256, 4096, 65536) for f in (0, 7.5, 30, 93.75, 281.25, 480, 24000)])
    np.savez_compressed(os.path.join(HERE, "tables.npz"), **tab)

    # single extractor, other windows / overlap 50 %, plus the streaming API (CE:353-424)
    L, R = uo.synth_stereo(6000, 7, stress=False)
    L64, R64 = L.astype(np.float64), R.astype(np.float64)
    one = dict(meta, in_L=L, in_R=R)
    e = ce.MultiBandExtractorAccu(512, 0.5, ce.make_sqrt_hann, 300.0, 5000.0, 48000, "raised_cosine",
                                  100.0, 800.0)
    for k, v in zip(("C", "Ls", "Rs"), e.process_all_blocks(L64, R64)):
        one[f"ref_sqrt_hann50_{k}"] = v
    e = ce.MultiBandExtractorAccu(256, 0.75, ce.make_hann, 1000.0, 24000.0, 48000, "bogus_mode")
    for k, v in zip(("C", "Ls", "Rs"), e.process_all_blocks(L64, R64)):
        one[f"ref_hann75_{k}"] = v
    e = ce.MultiBandExtractorAccu(1024, 0.75, ce.make_blackman_harris, 200.0, 2000.0, 48000,
                                  "raised_cosine", 50.0, 500.0)
    chunks = []
    for f in range(12):
        chunks.append(np.stack(e.process_stereo_chunk(L64[f * 256:f * 256 + 1024],
                                                     R64[f * 256:f * 256 + 1024])))
    one["ref_stream_chunks"] = np.stack(chunks)          # [12, 3(C,L,R), 256]
    one["ref_stream_flush"] = np.stack(e.flush_final())  # [3, 1024]
    np.savez_compressed(os.path.join(HERE, "single_band.npz"), **one)

    # Bela program (bela/upmix.cpp, compiled unmodified against oracle/bela_shim)
    so = os.path.join(ROOT, "oracle", "_ref", "libbela_ref.so")
    lib = ctypes.CDLL(so)
    lib.bela_ref_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_int,
                                 ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
    for hw, nblk in ((2048, 20), (512, 40)):
        n = hw * nblk
        L, R = uo.synth_stereo(n, 2, stress=True)
        oL = np.zeros(n, np.float32)
        oR = np.zeros(n, np.float32)
        rc = lib.bela_ref_run(L.ctypes.data, R.ctypes.data, n, hw, ctypes.c_float(48000.0),
                              oL.ctypes.data, oR.ctypes.data)
        assert rc == 0
        np.savez_compressed(os.path.join(HERE, f"bela_hw{hw}.npz"), **meta, in_L=L, in_R=R,
                            ref_outL=oL, ref_outR=oR, hw=np.array(hw), sr=np.array(48000))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
