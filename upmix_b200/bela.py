"""Bela-equivalent mode: the real-time C++ program of the reference (bela/upmix.cpp, cited BU:line)
served by the same CUDA kernels.

What the Bela program does differently from the Python prototype (all verified by
running upmix.cpp against a shim):
  * STFT size rule clamped to hwBlock*4 (BU:498-506), at most 8 bands (BU:508);
  * 75 % overlap, Blackman-Harris used for BOTH analysis and synthesis (BU:200-201), computed in
    float32 (BU:59-71);
  * applyRaisedCosineFilter zeroes every bin outside [binLow, binHigh] before it fades them
    (BU:319-324), so the band limit is effectively a hard-zero pass band; bins by lround (BU:45-54);
  * per band L + 0.5 C / R + 0.5 C (BU:295-303), bands summed (BU:487-490);
  * every band waits for 4*hw buffered samples (BU:232-237): constant latency 3*hw.

`MultiBandUpmix` mirrors the C++ class (setThresholdMultiplier / setup / process) with device state
carried between hardware blocks; `bela_offline` gives the same stream for a whole signal in one call.
The Bela SDK glue (BelaContext, audioRead/audioWrite) is board I/O and out of scope.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np

from . import _native
from . import center_extraction as ce

MAX_STFT_SIZE = 8192      # BU:24
THRESHOLD_MULTI = 32.0    # BU:27
XO_FRACTION = 0.25        # BU:29 (moot: the fades act on bins that are already zero)
MAX_BANDS = 8             # BU:508


def make_blackman_harris_f32(size: int) -> np.ndarray:
    """Blackman-Harris evaluated in float32 like makeBlackmanHarris (BU:59-71)."""
    f = np.float32
    ratio = np.arange(size, dtype=np.float32) / f(size - 1)
    pi = f(math.pi)
    w = (f(0.35875) - f(0.48829) * np.cos(f(2.0) * pi * ratio) + f(0.14128) * np.cos(f(4.0) * pi * ratio)
         - f(0.01168) * np.cos(f(6.0) * pi * ratio))
    return w.astype(np.float32)


def freq_to_bin_bela(freq_hz: float, sr: float, fft_size: int) -> int:
    """freqToBin (BU:45-54): freqHz * fftSize / sr in FLOAT32 (left to right, like the C++ expression), clamped to
    [0, fftSize/2], std::lround (half away from zero)."""
    f = np.float32
    b = f(f(f(freq_hz) * f(fft_size)) / f(sr))
    b = min(max(b, f(0.0)), f(fft_size // 2))
    return int(math.floor(float(b) + 0.5))


def compute_block_size_bela(f_low: float, sr: float, hw_block: int, threshold_multiplier: float = THRESHOLD_MULTI) -> int:
    """computeBlockSizeForLowFreq (BU:498-506): threshold = (sr * multiplier) / f_low in FLOAT32, next power of two
    of its ceiling, clamped to hwBlock*4."""
    f = np.float32
    if f(f_low) <= f(0.0):
        return hw_block * 4
    threshold = f(f(f(sr) * f(threshold_multiplier)) / f(f_low))
    candidate = ce.next_power_of_2(int(math.ceil(float(threshold))))
    return min(candidate, hw_block * 4)


class BelaBandExtractor(ce.MultiBandExtractorAccu):
    """One Overlap75UpmixBand (BU:174-416) expressed with the prototype's tables."""

    def __init__(self, hw_block: int, sr: float, stft_size: int, f_low: float, f_high: float):
        if stft_size > MAX_STFT_SIZE:
            raise RuntimeError("stftSize too large")          # BU:186-187
        self.block_size = stft_size
        self.overlap = 0.75
        self.hop_size = stft_size // 4
        self.analysis_window = make_blackman_harris_f32(stft_size)
        self.synthesis_window = self.analysis_window.copy()
        self.sr = sr
        self.f_low = f_low
        self.f_high = f_high
        self.xover_mode = "bela"
        self.xover_width_low_hz = f_low * XO_FRACTION if f_low > 0 else 0.0
        self.xover_width_high_hz = f_high * XO_FRACTION if f_high < sr * 0.5 else 0.0
        self.hw_block = hw_block
        self._plan = None
        self._plan_key = None
        self._stream_state = None
        self._frames_done = 0

    def band_gain(self) -> np.ndarray:
        n_bins = self.block_size // 2 + 1
        lo = freq_to_bin_bela(self.f_low, self.sr, self.block_size)
        hi = freq_to_bin_bela(self.f_high, self.sr, self.block_size)
        if lo > hi:
            lo, hi = hi, lo
        hi = min(hi, n_bins - 1)
        g = np.zeros(n_bins, dtype=np.float64)
        g[lo:hi + 1] = 1.0
        return g


def bela_chain_bands(band_edges: Sequence[float], sr: float, hw_block: int,
                     threshold_multiplier: float = THRESHOLD_MULTI) -> List[BelaBandExtractor]:
    """Bands as MultiBandUpmix::setup builds them (BU:439-471); band_edges has numBands+1 entries."""
    n_bands = min(len(band_edges) - 1, MAX_BANDS)
    out = []
    for i in range(n_bands):
        n = compute_block_size_bela(band_edges[i], sr, hw_block, threshold_multiplier)
        print(f"Band {i}: fLow = {band_edges[i]:8.1f} Hz, fHigh = {band_edges[i+1]:8.1f} Hz --> STFT Size = {n}")
        out.append(BelaBandExtractor(hw_block, sr, n, band_edges[i], band_edges[i + 1]))
    return out


def bela_offline(L, R, sr: float, hw_block: int, band_edges: Optional[Sequence[float]] = None,
                 threshold_multiplier: float = THRESHOLD_MULTI, bands: Optional[List[BelaBandExtractor]] = None):
    """The two output channels render() would have produced for the whole signal: fold-down of all
    bands delayed by 3*hw_block; only whole hardware blocks are emitted (like the audio callback)."""
    if bands is None:
        if band_edges is None:
            band_edges = [0.0, 500.0, 2000.0, 8000.0, sr * 0.5]          # BU:525
        bands = bela_chain_bands(band_edges, sr, hw_block, threshold_multiplier)
    n = (len(L) // hw_block) * hw_block
    d = 3 * hw_block
    is_t = ce._is_cuda_tensor(L)
    fl, fr = ce.extract_stereo_fold_down(L[:n], R[:n], sr, bands)
    if is_t:
        import torch
        ol, orr = torch.zeros_like(fl), torch.zeros_like(fr)
    else:
        ol, orr = np.zeros_like(fl), np.zeros_like(fr)
    if n > d:
        ol[d:] = fl[:n - d]
        orr[d:] = fr[:n - d]
    return ol, orr


class MultiBandUpmix:
    """Streaming twin of the C++ MultiBandUpmix (BU:426-514): feed hardware blocks, get blocks back
    with 3*hwBlock latency.  State (input history + per-band overlap-add rings) stays on the device."""

    def __init__(self):
        self._thr = THRESHOLD_MULTI
        self.bands: List[BelaBandExtractor] = []
        self._stream = None
        self._extra = None
        self.hw_block = 0
        self.sr = 0.0

    def setThresholdMultiplier(self, multiplier: float):
        self._thr = float(multiplier)

    def setup(self, hwBlock: int, sr: float, numBands: int, bandEdges: Sequence[float]):
        torch = _native._torch()
        self.hw_block = int(hwBlock)
        self.sr = float(sr)
        numBands = min(numBands, MAX_BANDS)
        self.bands = bela_chain_bands(list(bandEdges)[:numBands + 1], sr, hwBlock, self._thr)
        for b in self.bands:
            if hwBlock % b.hop_size:
                raise ValueError(f"hwBlock={hwBlock} is not a multiple of hop {b.hop_size}")
        plan = ce.plan_for(self.bands, _native.OUT_FOLD)
        self._stream = plan.stream_open(1)
        extra = 3 * self.hw_block - self._stream.delay       # bands smaller than 4*hw finish earlier
        if extra < 0:
            raise ValueError("a band is larger than hwBlock*4")
        self._extra = torch.zeros((2, extra), dtype=torch.float32, device=f"cuda:{plan.device}") if extra else None

    def process(self, inL, inR, frames: Optional[int] = None):
        """One hardware block in, one block out (left, right)."""
        torch = _native._torch()
        is_t = ce._is_cuda_tensor(inL)
        dev = self._stream.state.device
        l = inL if is_t else torch.from_numpy(np.ascontiguousarray(inL, dtype=np.float32)).to(dev)
        r = inR if is_t else torch.from_numpy(np.ascontiguousarray(inR, dtype=np.float32)).to(dev)
        if frames is not None:
            l, r = l[:frames], r[:frames]
        ol, orr = self._stream.block(l, r)
        if self._extra is not None:
            cat = torch.cat([self._extra, torch.stack([ol, orr])], dim=1)
            n = ol.shape[0]
            ol, orr = cat[0, :n], cat[1, :n]
            self._extra = cat[:, n:].contiguous()
        if is_t:
            return ol, orr
        return ol.cpu().numpy(), orr.cpu().numpy()
