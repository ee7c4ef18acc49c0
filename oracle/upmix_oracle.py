"""CPU oracle for the multi-band STFT centre-extraction path  --  TEST INFRASTRUCTURE ONLY.

A float64 numpy restatement of the reference prototype's algorithm
(/root/reference/python-prototype/center_extraction.py, cited below as CE:line).  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may
import this module; the product (`upmix_b200/`) never does, and it has no CPU fallback.

Parity pinning: the reference ships no tests or golden vectors, so this
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: `tests/golden/make_golden.py` imports the
unmodified reference (oracle/ref_loader.py) in the development container and commits its outputs
under `tests/golden/` (numpy 2.3.5 recorded in the fixtures); `tests/test_oracle.py` checks both
entry points below against those fixtures, and, when /root/reference is mounted, live.

Third-party arithmetic the reference leans on and that is not under /root/reference:
numpy.fft.rfft/irfft (pocketfft, CE:122,135) -- version unpinned by the reference (no
requirements file); fixtures were made with numpy 2.3.5, which this oracle also calls.

Two entry points compute the same thing:
  * `process_band_frames`  -- frame-by-frame loop, same operation order and dtypes as CE:353-472.
    This is the "port" that bench.py times as the CPU baseline (one thread per band, CE:499-501).
  * `process_band_batched` -- all frames of a band in a few big numpy calls, for long signals.
"""
from __future__ import annotations

import math
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

EPS = 1e-12  # CE:36


# --------------------------------------------------------------------------------------------
# windows (CE:42-75) -- every generator returns float32
# --------------------------------------------------------------------------------------------
def blackman_harris(n: int) -> np.ndarray:
    """4-term Blackman-Harris, symmetric (denominator n-1), float64 maths then float32 (CE:42-53)."""
    t = np.arange(n)
    den = n - 1
    w = (0.35875 - 0.48829 * np.cos(2 * np.pi * t / den)
         + 0.14128 * np.cos(4 * np.pi * t / den) - 0.01168 * np.cos(6 * np.pi * t / den))
    return w.astype(np.float32)


def hann(n: int) -> np.ndarray:          # CE:61-63
    return np.hanning(n).astype(np.float32)


def sqrt_hann(n: int) -> np.ndarray:     # CE:56-59
    return np.sqrt(np.hanning(n)).astype(np.float32)


def blackman(n: int) -> np.ndarray:      # CE:65-67
    return np.blackman(n).astype(np.float32)


def hamming(n: int) -> np.ndarray:       # CE:69-71
    return np.hamming(n).astype(np.float32)


def rect(n: int) -> np.ndarray:          # CE:73-75
    return np.ones(n, dtype=np.float32)


def wola_synthesis_window(ana: np.ndarray, overlap: float) -> np.ndarray:
    """syn[n] = ana[n] / (sum_{k<K} ana[(n+k*hop) mod L]^2 + EPS)   (CE:80-105).

    The reference accumulates in a Python loop starting from the float 0.0; under NumPy >= 2 a
    Python float is "weak", so every partial sum is rounded to the window's dtype (float32).  The
    same left-to-right float32 additions are done here, one whole vector per k.
    """
    ana = np.asarray(ana)
    L = len(ana)
    hop = int(L * (1.0 - overlap))
    if hop < 1:
        raise ValueError("overlap too large: hop < 1")
    K = int(round(1.0 / (1.0 - overlap)))
    idx = np.arange(L)
    # The reference squares numpy SCALARS (`analysis_window[idx] ** 2`, CE:102).  For float32 that
    # goes through the C library's powf, which is not always correctly rounded (52 of 65536
    # Blackman-Harris samples differ from x*x by one ulp with glibc 2.39), whereas the array
    # expression `ana ** 2` multiplies.  Square element by element to stay bit-identical.
    sq = np.array([v ** 2 for v in ana], dtype=ana.dtype)
    acc = np.zeros(L, dtype=ana.dtype)
    for k in range(K):
        acc = acc + sq[(idx + k * hop) % L]
    return (ana / (acc + ana.dtype.type(EPS))).astype(ana.dtype)


# --------------------------------------------------------------------------------------------
# size rule and bin mapping (CE:142-212)
# --------------------------------------------------------------------------------------------
def freq_to_bin(freq_hz: float, sr: float, n_fft: int) -> int:
    """Nearest bin with Python's round-half-even (CE:142-154)."""
    return int(round(freq_hz / (sr / float(n_fft))))


def next_pow2(x: int) -> int:            # CE:156-171
    p = 1
    while p < x:
        p <<= 1
    return p


def block_size_for_low_freq(f_low: float, sr: float, max_block: int = 2 ** 16,
                            factor: float = 32) -> int:
    """Dynamic-resolution rule: nextpow2(ceil(sr*factor/f_low)) clamped to max_block (CE:173-197)."""
    if f_low <= 0.0:
        return max_block
    return min(next_pow2(int(np.ceil((sr * factor) / f_low))), max_block)


def band_gain(n_fft: int, sr: float, f_low: float, f_high: float, mode: str,
              width_low_hz: float, width_high_hz: float) -> np.ndarray:
    """Per-bin real gain (float64, n_fft/2+1 bins) that CE:273-351 applies to both spectra.

    hard_zero (and any unknown mode, CE:349-351): 1 on [lo, hi], 0 elsewhere.
    raised_cosine (CE:282-332): half-cosine fade-in below lo only if f_low > 0, half-cosine
    fade-out above hi only if f_high < sr/2, zero beyond the fades.
    """
    nb = n_fft // 2 + 1
    lo = freq_to_bin(f_low, sr, n_fft)
    hi = freq_to_bin(f_high, sr, n_fft)
    if lo > hi:
        lo, hi = hi, lo
    g = np.ones(nb, dtype=np.float64)
    if mode != "raised_cosine":
        g[:lo] = 0.0
        g[hi + 1:] = 0.0
        return g
    lo = max(lo, 0)
    hi = min(hi, nb - 1)
    if lo > hi:
        return np.zeros(nb, dtype=np.float64)
    n_in = freq_to_bin(width_low_hz, sr, n_fft)
    n_out = freq_to_bin(width_high_hz, sr, n_fft)
    if f_low > 0:
        start = max(0, lo - n_in)
        g[:start] = 0.0
        for i in range(lo - start):
            g[start + i] *= 0.5 * (1.0 - np.cos(np.pi * ((i + 0.5) / (lo - start))))
    if f_high < sr * 0.5:
        start = hi + 1
        if start < nb:
            stop = min(start + n_out, nb)
            for i in range(stop - start):
                g[start + i] *= 0.5 * (1.0 + np.cos(np.pi * ((i + 0.5) / (stop - start))))
            g[stop:] = 0.0
    return g


# --------------------------------------------------------------------------------------------
# band description
# --------------------------------------------------------------------------------------------
@dataclass
class BandSpec:
    n_fft: int
    hop: int
    ana: np.ndarray      # float32 [n_fft]
    syn: np.ndarray      # float32 [n_fft]
    gain: np.ndarray     # float64 [n_fft/2+1]
    f_low: float = 0.0
    f_high: float = 0.0


def make_band(n_fft: int, overlap: float, window: Callable[[int], np.ndarray], f_low: float,
              f_high: float, sr: float, mode: str = "hard_zero", width_low_hz: float = 50.0,
              width_high_hz: float = 50.0, synthesis: Optional[np.ndarray] = None) -> BandSpec:
    """One band's tables, as MultiBandExtractorAccu.__init__ builds them (CE:240-271)."""
    hop = int(n_fft * (1 - overlap))
    if hop < 1:
        raise ValueError("overlap too large: hop < 1")
    ana = window(n_fft)
    syn = wola_synthesis_window(ana, overlap) if synthesis is None else synthesis
    gain = band_gain(n_fft, sr, f_low, f_high, mode, width_low_hz, width_high_hz)
    return BandSpec(n_fft, hop, ana, syn, gain, f_low, f_high)


def chain(band_edges: Sequence[float], overlap: float, window: Callable[[int], np.ndarray],
          sr: float, mode: str = "raised_cosine", max_block: int = 2 ** 16, factor: float = 32,
          xo_fraction: float = 0.25) -> List[BandSpec]:
    """Adjacent bands from crossover edges (CE:518-580): Nyquist appended, each band's low fade is
    the previous band's high fade width, high fade width = xo_fraction * f_high (CE:200-212)."""
    edges = list(band_edges)
    if edges[-1] < sr / 2.0:
        edges.append(sr / 2.0)
    bands, prev_w = [], 0.0
    for f_low, f_high in zip(edges[:-1], edges[1:]):
        n_fft = block_size_for_low_freq(f_low, sr, max_block, factor)
        w_high = f_high * xo_fraction
        bands.append(make_band(n_fft, overlap, window, f_low, f_high, sr, mode, prev_w, w_high))
        prev_w = w_high
    return bands


def bela_chain(band_edges: Sequence[float], sr: float, hw_block: int,
               factor: float = 32) -> List[BandSpec]:
    """Bands as bela/upmix.cpp configures them (BU:176-226,439-506), expressed with the prototype's
    tables: size rule clamped to hw_block*4, 75 % overlap, Blackman-Harris for BOTH analysis and
    synthesis (BU:200-201), and an effective hard-zero pass band because applyRaisedCosineFilter
    zeroes outside [binLow, binHigh] before fading (BU:319-324).  Bin mapping is lround + clamp
    (BU:45-54), which differs from Python round only at exact .5 bins."""
    bands = []
    for f_low, f_high in zip(band_edges[:-1], band_edges[1:]):
        n_fft = block_size_for_low_freq(f_low, sr, hw_block * 4, factor)
        ana = blackman_harris(n_fft)
        nb = n_fft // 2 + 1

        def cbin(f):
            b = min(max(f * n_fft / sr, 0.0), float(n_fft // 2))
            return int(math.floor(b + 0.5))
        lo, hi = sorted((cbin(f_low), cbin(f_high)))
        hi = min(hi, nb - 1)
        g = np.zeros(nb, dtype=np.float64)
        g[lo:hi + 1] = 1.0
        bands.append(BandSpec(n_fft, n_fft // 4, ana, ana.copy(), g, f_low, f_high))
    return bands


# --------------------------------------------------------------------------------------------
# per-frame mask (CE:373-384)
# --------------------------------------------------------------------------------------------
def _centre_split(sl: np.ndarray, sr_: np.ndarray):
    cross_mag = np.abs(sl * np.conjugate(sr_))
    ml, mr = np.abs(sl), np.abs(sr_)
    coherence = cross_mag / (ml * mr + EPS)
    balance = (ml - mr) / (ml + mr + EPS)
    cf = coherence * (1.0 - np.abs(balance))
    c = 0.5 * cf * (sl + sr_)
    return c, sl - c, sr_ - c


def n_frames(n_sig: int, hop: int) -> int:
    """Frames that can touch output samples [0, n_sig): f*hop < n_sig (loop CE:448-460, trim CE:468)."""
    return -(-n_sig // hop) if n_sig > 0 else 0


# --------------------------------------------------------------------------------------------
# frame loop -- the faithful port (CE:353-472)
# --------------------------------------------------------------------------------------------
def process_band_frames(band: BandSpec, L: np.ndarray, R: np.ndarray
                        ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """One band over the whole signal, one frame at a time.  Returns (centre, left, right) float32."""
    N, H = band.n_fft, band.hop
    L = np.asarray(L, dtype=np.float64)      # the reference is fed float64 (sf.read, main.py:43)
    R = np.asarray(R, dtype=np.float64)
    n_sig = len(L)
    F = n_frames(n_sig, H)
    pad = max(0, (F - 1) * H + N - n_sig) if F else 0
    Lp = np.concatenate([L, np.zeros(pad)])
    Rp = np.concatenate([R, np.zeros(pad)])
    acc = np.zeros((3, N), dtype=np.float32)
    out = np.zeros((3, F * H + N), dtype=np.float32)
    for f in range(F):
        seg = slice(f * H, f * H + N)
        sl = np.fft.rfft(Lp[seg] * band.ana) * band.gain       # CE:121-122, 370
        sr_ = np.fft.rfft(Rp[seg] * band.ana) * band.gain
        for ch, spec in enumerate(_centre_split(sl, sr_)):
            rec = np.fft.irfft(spec).astype(np.float32)         # CE:135
            rec *= band.syn                                     # CE:136
            acc[ch] += rec                                      # CE:392-394
        out[:, f * H:(f + 1) * H] = acc[:, :H]                  # CE:397-399
        acc[:, :-H] = acc[:, H:]                                # CE:402-407
        acc[:, -H:] = 0
    out[:, F * H:F * H + N] = acc                               # flush, CE:411-424
    return out[0, :n_sig].copy(), out[1, :n_sig].copy(), out[2, :n_sig].copy()


# --------------------------------------------------------------------------------------------
# batched restatement (same arithmetic, frames stacked)
# --------------------------------------------------------------------------------------------
def process_band_batched(band: BandSpec, L: np.ndarray, R: np.ndarray, chunk_frames: int = 0
                         ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Same result as process_band_frames (bit-identical float32) with stacked frames.
    Requires hop | n_fft.  The float32 overlap-add keeps the reference's order: for every output
    hop the oldest frame's contribution is added first (CE:392-407)."""
    N, H = band.n_fft, band.hop
    if N % H:
        raise ValueError("batched oracle needs hop | n_fft")
    K = N // H
    n_sig = len(L)
    F = n_frames(n_sig, H)
    out = np.zeros((3, (F + K) * H), dtype=np.float32)
    if F == 0:
        return out[0, :0].copy(), out[1, :0].copy(), out[2, :0].copy()
    total = (F - 1) * H + N
    Lp = np.zeros(total, dtype=np.float64)
    Rp = np.zeros(total, dtype=np.float64)
    Lp[:n_sig] = L
    Rp[:n_sig] = R
    if chunk_frames <= 0:
        chunk_frames = max(K, (1 << 22) // N)
    ana = band.ana.astype(np.float64)
    # every chunk re-derives the K-1 frames before it so that each output hop sees all of its
    # contributions, added oldest-first.
    for f0 in range(0, F, chunk_frames):
        fa = max(0, f0 - (K - 1))
        fb = min(F, f0 + chunk_frames)
        nfr = fb - fa
        idx = (np.arange(nfr)[:, None] + fa) * H + np.arange(N)[None, :]
        sl = np.fft.rfft(Lp[idx] * ana, axis=1) * band.gain
        sr_ = np.fft.rfft(Rp[idx] * ana, axis=1) * band.gain
        for ch, spec in enumerate(_centre_split(sl, sr_)):
            rec = np.fft.irfft(spec, n=N, axis=1).astype(np.float32)
            rec *= band.syn
            rec = rec.reshape(nfr, K, H)
            # output hop h (absolute) = sum over q=K-1..0 of frame (h-q), quarter q
            h_lo, h_hi = f0, fb           # hops finished by this chunk's own frames
            acc = np.zeros((h_hi - h_lo, H), dtype=np.float32)
            for q in range(K - 1, -1, -1):
                fr = np.arange(h_lo, h_hi) - q          # absolute frame index
                ok = fr >= 0
                acc[ok] += rec[fr[ok] - fa, q]
            out[ch, h_lo * H:h_hi * H] = acc.reshape(-1)
            if fb == F:                                  # tail hops F .. F+K-2 (flush)
                for h in range(F, F + K - 1):
                    t = np.zeros(H, dtype=np.float32)
                    for q in range(K - 1, -1, -1):
                        fr = h - q
                        if 0 <= fr < F:
                            t += rec[fr - fa, q]
                    out[ch, h * H:(h + 1) * H] = t
    return out[0, :n_sig].copy(), out[1, :n_sig].copy(), out[2, :n_sig].copy()


# --------------------------------------------------------------------------------------------
# multi-band (CE:477-513) and fold-down (main.py:143-146 / bela/upmix.cpp:295-303)
# --------------------------------------------------------------------------------------------
def upmix_multiband(bands: Sequence[BandSpec], L: np.ndarray, R: np.ndarray, batched: bool = True,
                    threads: bool = False):
    """Sum of the per-band outputs in band order, float32.  Returns (centre, left, right)."""
    fn = process_band_batched if batched else process_band_frames
    if threads:
        with ThreadPoolExecutor() as ex:
            futs = [ex.submit(fn, b, L, R) for b in bands]
            res = [f.result() for f in futs]
    else:
        res = [fn(b, L, R) for b in bands]
    n = len(L)
    c = np.zeros(n, dtype=np.float32)
    l = np.zeros(n, dtype=np.float32)
    r = np.zeros(n, dtype=np.float32)
    for bc, bl, br in res:
        c += bc
        l += bl
        r += br
    return c, l, r


def bela_offline(bands: Sequence[BandSpec], L: np.ndarray, R: np.ndarray, hw_block: int,
                 batched: bool = True):
    """What bela/upmix.cpp's render() emits for the whole signal: per band L+0.5C / R+0.5C
    (BU:295-303), bands summed (BU:487-490), delayed by 3*hw_block samples because every band waits
    for 4*hw buffered samples (BU:232-237); only whole hardware blocks are produced."""
    n_blocks = len(L) // hw_block
    n = n_blocks * hw_block
    fn = process_band_batched if batched else process_band_frames
    outl = np.zeros(n, dtype=np.float32)
    outr = np.zeros(n, dtype=np.float32)
    d = 3 * hw_block
    for b in bands:
        c, l, r = fn(b, np.asarray(L[:n], dtype=np.float64), np.asarray(R[:n], dtype=np.float64))
        vl = l + np.float32(0.5) * c
        vr = r + np.float32(0.5) * c
        if n > d:
            outl[d:] += vl[:n - d]
            outr[d:] += vr[:n - d]
    return outl, outr


# --------------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------------
def synth_stereo(n: int, seed: int, sr: int = 48000, stress: bool = False
                 ) -> Tuple[np.ndarray, np.ndarray]:
    """L = 0.1*N(0,1), R = 0.5*L + 0.05*N(0,1), float32.  With stress=True a half-second is scaled
    by 1e-5 (EPS-dominated bins) and another half-second has R = 0 (hard-panned)."""
    rng = np.random.default_rng(seed)
    L = 0.1 * rng.standard_normal(n)
    R = 0.5 * L + 0.05 * rng.standard_normal(n)
    L = L.astype(np.float32)
    R = R.astype(np.float32)
    if stress and n >= 4:
        q = n // 4
        w = min(sr // 2, q)
        L[q:q + w] *= np.float32(1e-5)
        R[q:q + w] *= np.float32(1e-5)
        R[2 * q:2 * q + w] = 0.0
    return L, R


def snr_db(ref: np.ndarray, got: np.ndarray) -> float:
    ref = np.asarray(ref, dtype=np.float64)
    err = np.asarray(got, dtype=np.float64) - ref
    pe = float(np.sum(err * err))
    ps = float(np.sum(ref * ref))
    if pe == 0.0:
        return float("inf")
    if ps == 0.0:
        return float("-inf")
    return 10.0 * math.log10(ps / pe)
