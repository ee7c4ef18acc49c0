#!/usr/bin/env python3
"""Multi-band STFT upmix (stereo -> centre / left-side / right-side) on B200.

Drop-in for the reference prototype's module of the same name
(/root/reference/python-prototype/center_extraction.py, cited as CE:line): every public name,
signature, default and return order is kept, so `import upmix_b200.center_extraction as ce` can
replace `import center_extraction as ce` (main.py:23).  What differs is where the work happens:

  * band / window / gain construction stays on the host and is bit-identical to the reference
    (it is the parameter half of the path: CE:42-105, 142-212, 240-271, 282-351, 518-580);
  * the per-band STFT chain and the band sum (CE:353-513) run as hand-written sm_100a CUDA
    kernels behind the C ABI in include/upmix_b200.h (float32 arithmetic; the reference computes
    spectra in float64 -- parity is >= 100 dB SNR, see tests/).  There is no CPU fallback: without
    the built library or without a CUDA device these calls raise.

Extensions that do not change the positional signatures: `chain_bands(..., max_block_size=,
threshold_factor=, xo_fraction=)` expose the constants the reference hard-codes (CE:173, 212); the
processing calls also accept float32 CUDA tensors and then return CUDA tensors.
"""
from __future__ import annotations

import math
import os
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import _native

###############################################################################
# Constants
###############################################################################
EPS = 1e-12  # CE:36

###############################################################################
# Window functions (CE:42-75): float64 maths, float32 result
###############################################################################


def make_blackman_harris(N: int) -> np.ndarray:
    """4-term Blackman-Harris window, symmetric (period N-1)."""
    n = np.arange(N)
    coef = (0.35875, 0.48829, 0.14128, 0.01168)
    w = (coef[0] - coef[1] * np.cos(2 * np.pi * n / (N - 1)) + coef[2] * np.cos(4 * np.pi * n / (N - 1))
         - coef[3] * np.cos(6 * np.pi * n / (N - 1)))
    return w.astype(np.float32)


def make_sqrt_hann(N: int) -> np.ndarray:
    """Square-root Hann window (for 50 % overlap)."""
    return np.sqrt(np.hanning(N)).astype(np.float32)


def make_hann(N: int) -> np.ndarray:
    """Hann window."""
    return np.hanning(N).astype(np.float32)


def make_blackman(N: int) -> np.ndarray:
    """Blackman window."""
    return np.blackman(N).astype(np.float32)


def make_hamming(N: int) -> np.ndarray:
    """Hamming window."""
    return np.hamming(N).astype(np.float32)


def make_rect(N: int) -> np.ndarray:
    """Rectangular window."""
    return np.ones(N, dtype=np.float32)


###############################################################################
# WOLA synthesis-window design (CE:80-105)
###############################################################################
_SYN_CACHE: dict = {}


def design_wola_synthesis_window(analysis_window: np.ndarray, overlap: float) -> np.ndarray:
    """w_S[n] = w_A[n] / (sum_k w_A[(n + k*hop) mod L]^2 + EPS), k over the K overlapping frames.

    Bit-identical to the reference's pure-Python L*K loop but O(L) Python work: the squares are taken
    once per sample -- as numpy *scalars*, because the reference squares scalars (CE:102) and the C
    library's float32 pow is not always the correctly rounded product -- and the K partial sums are
    added left to right in the window's dtype, which is what NumPy >= 2 does to the reference's
    Python-float accumulator.  Results are cached per (window bytes, overlap).
    """
    analysis_window = np.asarray(analysis_window)
    L = len(analysis_window)
    hop = int(L * (1.0 - overlap))
    if hop < 1:
        raise ValueError("Overlap too large; resulting hop size < 1.")
    key = (analysis_window.dtype.str, L, float(overlap), analysis_window.tobytes())
    hit = _SYN_CACHE.get(key)
    if hit is not None:
        return hit.copy()
    K = int(round(1.0 / (1.0 - overlap)))
    dt = analysis_window.dtype
    squares = np.array([v ** 2 for v in analysis_window], dtype=dt)
    idx = np.arange(L)
    total = np.zeros(L, dtype=dt)
    for k in range(K):
        total = total + squares[(idx + k * hop) % L]
    eps = dt.type(EPS) if dt.kind == "f" else EPS
    syn = np.asarray(analysis_window / (total + eps), dtype=dt)
    if len(_SYN_CACHE) > 64:
        _SYN_CACHE.clear()
    _SYN_CACHE[key] = syn
    return syn.copy()


###############################################################################
# STFT helpers (CE:110-137) -- host utilities kept for API compatibility; the extraction path
# itself does its transforms on the GPU
###############################################################################
def forward_stft(block: np.ndarray, analysis_win: np.ndarray) -> np.ndarray:
    """rFFT of the windowed block."""
    return np.fft.rfft(block * analysis_win)


def inverse_stft(spec: np.ndarray, synthesis_win: np.ndarray) -> np.ndarray:
    """Inverse rFFT as float32, weighted by the synthesis window."""
    rec = np.fft.irfft(spec).astype(np.float32)
    rec *= synthesis_win
    return rec


###############################################################################
# Utilities (CE:142-212)
###############################################################################
def freq_to_bin(freq_hz: float, sr: float, fft_size: int) -> int:
    """Nearest rFFT bin of a frequency (Python round: ties to even)."""
    return int(round(freq_hz / (sr / float(fft_size))))


def next_power_of_2(x: int) -> int:
    """Smallest power of two >= x (1 for x < 1)."""
    if x < 1:
        return 1
    return 1 << (int(x) - 1).bit_length()


def compute_block_size_for_low_freq(f_low: float, sr: float, max_block_size: int = 2 ** 16,
                                    threshold_factor: float = 32) -> int:
    """Dynamic-resolution rule: enough samples for `threshold_factor` cycles of f_low, rounded up to
    a power of two and clamped to max_block_size; f_low <= 0 gives max_block_size."""
    if f_low <= 0.0:
        return max_block_size
    needed = int(np.ceil((sr * threshold_factor) / f_low))
    return min(next_power_of_2(needed), max_block_size)


XO_FRACTION = 0.25  # CE:212; bela/upmix.cpp:29


def hp_freq_to_crossover_width(hp_freq: float) -> float:
    """Crossover fade width in Hz: 25 % of the crossover frequency."""
    return hp_freq * XO_FRACTION


###############################################################################
# Per-band extractor (CE:217-472)
###############################################################################
def _as_host_f32(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x), dtype=np.float32)


def _is_cuda_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


class MultiBandExtractorAccu:
    """One band: band-limit (hard zero or raised-cosine fades), centre mask, WOLA resynthesis.

    Same constructor, attributes and methods as the reference class (CE:217-472).  The tables
    (`analysis_window`, `synthesis_window`, `band_gain()`) are built on the host exactly as the
    reference builds them; `process_all_blocks`, `process_stereo_chunk` and `flush_final` run on the GPU.
    """

    def __init__(self, block_size: int, overlap: float, window_func: Callable[[int], np.ndarray], f_low: float,
                 f_high: float, sr: float, xover_mode: str = "hard_zero", xover_width_low_hz: float = 50.0,
                 xover_width_high_hz: float = 50.0):
        self.block_size = block_size
        self.overlap = overlap
        self.hop_size = int(block_size * (1 - overlap))
        if self.hop_size < 1:
            raise ValueError("Overlap too large; hop size < 1 is not allowed.")
        self.analysis_window = window_func(block_size)
        self.synthesis_window = design_wola_synthesis_window(self.analysis_window, overlap)
        self.sr = sr
        self.f_low = f_low
        self.f_high = f_high
        self.xover_mode = xover_mode
        self.xover_width_low_hz = xover_width_low_hz
        self.xover_width_high_hz = xover_width_high_hz
        self._plan = None            # single-band native plan (lazy)
        self._plan_key = None
        self._stream_state = None    # device ring [3, block_size] for the chunk API
        self._frames_done = 0

    # ---- tables -------------------------------------------------------------------------------
    def band_gain(self) -> np.ndarray:
        """Real gain per rFFT bin (float64, block_size/2+1) equal to what `_band_limit` (CE:334-351)
        does to a spectrum of ones: the band-limit is frame-invariant, so it is a table here."""
        n_bins = self.block_size // 2 + 1
        fft_size = self.block_size
        lo = freq_to_bin(self.f_low, self.sr, fft_size)
        hi = freq_to_bin(self.f_high, self.sr, fft_size)
        if lo > hi:
            lo, hi = hi, lo
        gain = np.ones(n_bins, dtype=np.float64)
        if self.xover_mode != "raised_cosine":
            # "hard_zero", and the fallback for unknown modes (CE:345-351)
            gain[:lo] = 0.0
            gain[hi + 1:] = 0.0
            return gain
        lo = max(lo, 0)
        hi = min(hi, n_bins - 1)
        if lo > hi:
            gain[:] = 0.0
            return gain
        fade_lo = freq_to_bin(self.xover_width_low_hz, self.sr, fft_size)
        fade_hi = freq_to_bin(self.xover_width_high_hz, self.sr, fft_size)
        if self.f_low > 0:                       # fade in below the band (CE:304-315)
            first = max(0, lo - fade_lo)
            gain[:first] = 0.0
            width = lo - first
            for i in range(width):
                gain[first + i] *= 0.5 * (1.0 - np.cos(np.pi * ((i + 0.5) / width)))
        if self.f_high < self.sr * 0.5:          # fade out above the band (CE:318-332)
            first = hi + 1
            if first < n_bins:
                last = min(first + fade_hi, n_bins)
                width = last - first
                for i in range(width):
                    gain[first + i] *= 0.5 * (1.0 + np.cos(np.pi * ((i + 0.5) / width)))
                gain[last:] = 0.0
        return gain

    def plan_tables(self):
        """(n_fft, hop, ana, syn, gain) as the native plan wants them."""
        return (int(self.block_size), int(self.hop_size), _as_host_f32(self.analysis_window),
                _as_host_f32(self.synthesis_window), self.band_gain().astype(np.float32))

    def _native_plan(self) -> "_native.Plan":
        key = _fingerprint(self) + (_native._torch().cuda.current_device(),)
        if self._plan is None or self._plan_key != key:
            _check_supported([self])
            self._plan = _native.Plan([self.plan_tables()], _native.OUT_LSCRS)
            self._plan_key = key
            self._stream_state = None
        return self._plan

    # ---- streaming state, exposed under the reference's attribute names (CE:269-271) ----------
    def _ring(self):
        if self._stream_state is None:
            torch = _native._torch()
            plan = self._native_plan()
            self._stream_state = torch.zeros((3, self.block_size), dtype=torch.float32, device=f"cuda:{plan.device}")
            self._frames_done = 0
        return self._stream_state

    @property
    def accumC(self) -> np.ndarray:
        return self._ring()[0].cpu().numpy()

    @property
    def accumL(self) -> np.ndarray:
        return self._ring()[1].cpu().numpy()

    @property
    def accumR(self) -> np.ndarray:
        return self._ring()[2].cpu().numpy()

    def process_stereo_chunk(self, blkL, blkR) -> tuple:
        """One frame: forward STFT, band-limit, centre mask, inverse STFT, overlap-add; returns the
        first hop_size finished samples (centre, left, right).  State is carried on the device."""
        torch = _native._torch()
        plan = self._native_plan()
        ring = self._ring()
        as_tensor = _is_cuda_tensor(blkL)
        dev = f"cuda:{plan.device}"
        bl = blkL if as_tensor else torch.from_numpy(_as_host_f32(blkL)).to(dev)
        br = blkR if as_tensor else torch.from_numpy(_as_host_f32(blkR)).to(dev)
        if bl.shape != (self.block_size,) or br.shape != (self.block_size,):
            raise ValueError(f"blocks must have {self.block_size} samples")
        out = plan.frame_step(ring, self._frames_done, bl, br)
        self._frames_done += 1
        if as_tensor:
            return out
        return tuple(o.cpu().numpy() for o in out)

    def flush_final(self) -> tuple:
        """Remaining overlap-add contents (centre, left, right), block_size samples each; clears the state."""
        ring = self._ring()
        left = ring.cpu().numpy().copy()
        ring.zero_()
        return left[0], left[1], left[2]

    def process_all_blocks(self, L, R) -> tuple:
        """The whole signal through this band; returns (centre, left, right) float32 of len(L)."""
        plan = self._native_plan()
        return _run_plan(plan, L, R)


###############################################################################
# Plans over several bands
###############################################################################
def _check_supported(band_extractors: Sequence[MultiBandExtractorAccu]) -> None:
    for i, bex in enumerate(band_extractors):
        n, h = int(bex.block_size), int(bex.hop_size)
        if n & (n - 1) or n < _native.MIN_N or n > _native.LARGE_MAX_N:
            raise NotImplementedError(
                f"band {i}: block_size={n} is not a power of two in [{_native.MIN_N}, {_native.LARGE_MAX_N}]; "
                "the CUDA path implements the sizes the dynamic-resolution rule produces and has no CPU fallback")
        if n % h or h % 2:
            raise NotImplementedError(f"band {i}: hop_size={h} must be even and divide block_size={n}")
        if n > _native.FUSED_MAX_N and 4 * h != n and 2 * h != n:
            raise NotImplementedError(f"band {i}: block_size={n} > {_native.FUSED_MAX_N} needs 75 % or 50 % overlap")


_PLAN_CACHE: "list" = []      # [(key, plan)], most recent last


_RAMPS: dict = {}


def _content_key(arr) -> tuple:
    """Cheap content checksum of a window table (tens of microseconds for 65536 points): the bit patterns'
    sum and xor plus an index-weighted sum, so that in-place edits, swaps and replacements all change it."""
    a = _as_host_f32(arr)
    bits = a.view(np.uint32)
    ramp = _RAMPS.get(a.shape[0])
    if ramp is None:
        ramp = _RAMPS[a.shape[0]] = np.arange(1, a.shape[0] + 1, dtype=np.float64)
    return (a.shape[0], int(np.add.reduce(bits, dtype=np.uint64)), int(np.bitwise_xor.reduce(bits)) if bits.size else 0,
            float(np.dot(a.astype(np.float64), ramp)))


def _fingerprint(bex) -> tuple:
    """What a plan depends on: the extractor's parameters, the CONTENT of its two window tables (the reference's
    attributes are plain arrays, CE:257-258: editing them in place must not leave a stale device copy) and the
    class that supplies band_gain().  If any of this changes, the plan is rebuilt."""
    return (int(bex.block_size), int(bex.hop_size), float(bex.f_low), float(bex.f_high), float(bex.sr),
            str(bex.xover_mode), float(bex.xover_width_low_hz), float(bex.xover_width_high_hz),
            _content_key(bex.analysis_window), _content_key(bex.synthesis_window), type(bex).band_gain)


def plan_for(band_extractors: Sequence[MultiBandExtractorAccu], out_mode: int = _native.OUT_LSCRS,
             flags: int = 0) -> "_native.Plan":
    """Native plan for a list of extractors, cached by content (parameters + window tables), output mode, flags
    and CUDA device.  flags: _native.PLAN_NO_DECIMATE keeps the full-size transform kernels for band-limited bands."""
    torch = _native._torch()
    key = (tuple(_fingerprint(b) for b in band_extractors), out_mode, int(flags), torch.cuda.current_device())
    for k, plan in _PLAN_CACHE:
        if k == key:
            return plan
    _check_supported(band_extractors)
    plan = _native.Plan([b.plan_tables() for b in band_extractors], out_mode, flags=flags)
    _PLAN_CACHE.append((key, plan))
    if len(_PLAN_CACHE) > 8:
        _PLAN_CACHE.pop(0)
    return plan


def _run_plan(plan: "_native.Plan", L, R) -> tuple:
    torch = _native._torch()
    if _is_cuda_tensor(L):
        return plan.process(L, R)
    if isinstance(L, torch.Tensor):                 # host tensors: the library's host pipeline, pinned tensors back
        if L.dtype not in (torch.float32, torch.float64) or R.dtype != L.dtype:
            L, R = L.to(torch.float32), R.to(torch.float32)
        return plan.process_host_tensors(L, R)
    # numpy (what main.py passes: float64 strided views of the interleaved array sf.read returns, MP:43-50):
    # converted, uploaded, processed and downloaded chunk by chunk inside upmix_process_host_ex
    return plan.process_host(L, R)


###############################################################################
# Multi-band extraction (CE:477-513)
###############################################################################
def extract_center_left_right_multi_band_in_memory(L, R, sr: float,
                                                   band_extractors: List[MultiBandExtractorAccu]) -> tuple:
    """All bands over the whole signal, summed in list order.  Returns (centre, left, right), float32,
    len(L) samples each.  numpy in -> numpy out; float32 CUDA tensors in ([n] or [tracks, n]) ->
    CUDA tensors out, left on the device; CPU torch tensors in (pinned for asynchronous copies) ->
    pinned CPU tensors out."""
    plan = plan_for(band_extractors, _native.OUT_LSCRS)
    return _run_plan(plan, L, R)


def extract_stereo_fold_down(L, R, sr: float, band_extractors: List[MultiBandExtractorAccu]) -> tuple:
    """(Ls + 0.5 C, Rs + 0.5 C) summed over bands: the Bela program's output mix (bela/upmix.cpp:295-303)
    and main.py's "stereo_sum" before peak scaling (main.py:143-146)."""
    plan = plan_for(band_extractors, _native.OUT_FOLD)
    return _run_plan(plan, L, R)


###############################################################################
# Chain bands (CE:518-580)
###############################################################################
def chain_bands(band_edges: List[float], overlap: float, window_func: Callable[[int], np.ndarray], sr: float,
                xover_mode: str = "raised_cosine", *, max_block_size: int = 2 ** 16, threshold_factor: float = 32,
                xo_fraction: Optional[float] = None) -> List[MultiBandExtractorAccu]:
    """Consecutive bands from crossover edges: sr/2 appended if missing, block size from the lower
    edge, fade widths chained (a band's low fade is the previous band's high fade, = xo_fraction*f_high)."""
    if band_edges[-1] < (sr / 2.0):
        band_edges = list(band_edges) + [sr / 2.0]
    extractors = []
    prev_width = 0.0
    for i, (f_low, f_high) in enumerate(zip(band_edges[:-1], band_edges[1:])):
        block_size = compute_block_size_for_low_freq(f_low, sr, max_block_size, threshold_factor)
        xover_low = prev_width
        xover_high = hp_freq_to_crossover_width(f_high) if xo_fraction is None else f_high * xo_fraction
        print(f"[Band {i+1}] f_low={f_low:.1f} Hz, f_high={f_high:.1f} Hz, block_size={block_size}, "
              f"xover_low={xover_low:.1f} Hz, xover_high={xover_high:.1f} Hz")
        extractors.append(MultiBandExtractorAccu(block_size=block_size, overlap=overlap, window_func=window_func,
                                                 f_low=f_low, f_high=f_high, sr=sr, xover_mode=xover_mode,
                                                 xover_width_low_hz=xover_low, xover_width_high_hz=xover_high))
        prev_width = xover_high
    return extractors


###############################################################################
# Visualisation / demo (CE:585-737) -- plotting only, imports are lazy
###############################################################################
def visualize_windows(analysis_window: np.ndarray, synthesis_window: np.ndarray, overlap: float):
    """Plot the two windows, the overlapped sum of analysis windows and of analysis*synthesis."""
    import matplotlib.pyplot as plt
    L = len(analysis_window)
    hop = int(L * (1 - overlap))
    K = int(round(1.0 / (1.0 - overlap)))
    total_len = L + (K - 1) * hop
    sums = []
    for w in (analysis_window, analysis_window * synthesis_window):
        acc = np.zeros(total_len, dtype=np.float32)
        for k in range(K):
            acc[k * hop:k * hop + L] += w
        sums.append(acc)
    fig, ax = plt.subplots(3, 1, figsize=(10, 10))
    ax[0].set_title("Analysis vs. Synthesis Window (Single Frame)")
    ax[0].plot(analysis_window, label="Analysis")
    ax[0].plot(synthesis_window, label="Synthesis (WOLA)")
    ax[0].legend(loc="best")
    ax[1].set_title(f"Sum of {K} Overlapped Analysis Windows at {overlap*100:.0f}% Overlap")
    ax[1].plot(sums[0])
    ax[2].set_title(f"Sum of {K} Overlapped Weighted Windows (Analysis*Synthesis)")
    ax[2].plot(sums[1])
    for a in ax:
        a.set_xlabel("Sample index")
        a.set_ylabel("Amplitude")
    fig.tight_layout()
    plt.show()


def main():
    """Demo: in/eyes.wav through bands [0, 40, 200, 2000], then plots of Ls+C+Rs against L+R."""
    from .wavio import read_wav
    in_path = os.path.join("in", "eyes.wav")
    if not os.path.isfile(in_path):
        raise FileNotFoundError(f"Input file not found: {in_path}")
    wave, sr = read_wav(in_path)
    print(f"Loaded '{in_path}' with sample rate {sr} and shape {wave.shape}")
    if wave.ndim == 1:
        wave = np.column_stack([wave, wave])
    L, R = wave[:, 0], wave[:, 1]
    overlap = 0.75
    band_extractors = chain_bands([0.0, 40.0, 200.0, 2000.0], overlap=overlap, window_func=make_blackman_harris,
                                  sr=sr, xover_mode="raised_cosine")
    if band_extractors:
        visualize_windows(band_extractors[0].analysis_window, band_extractors[0].synthesis_window, overlap)
    final_center, final_left, final_right = extract_center_left_right_multi_band_in_memory(L, R, sr, band_extractors)
    import matplotlib.pyplot as plt
    upmix_sum = final_left + final_center + final_right
    orig_sum = L + R
    upmix_norm = upmix_sum / (np.max(np.abs(upmix_sum)) + 1e-12)
    orig_norm = orig_sum / (np.max(np.abs(orig_sum)) + 1e-12)
    t = np.arange(len(upmix_norm)) / sr
    fig, ax = plt.subplots(2, 1, figsize=(12, 8))
    ax[0].plot(t, upmix_norm, label="Upmix (L + C + R)")
    ax[0].plot(t, orig_norm, label="Original (L + R)", alpha=0.75)
    ax[0].set_title("Time Domain Comparison")
    ax[0].legend(loc="upper right")
    freq = np.linspace(0, sr / 2, len(upmix_norm) // 2 + 1)
    ax[1].semilogy(freq, np.abs(np.fft.rfft(upmix_norm)), label="Upmix Spectrum")
    ax[1].semilogy(freq, np.abs(np.fft.rfft(orig_norm)), label="Original Spectrum", alpha=0.75)
    ax[1].set_title("Frequency Domain Comparison")
    ax[1].legend(loc="upper right")
    fig.tight_layout()
    plt.show()


if __name__ == "__main__":
    main()
