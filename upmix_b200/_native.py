"""ctypes binding of the C ABI in include/upmix_b200.h (csrc/libupmix_b200.so).

PyTorch is used only as plumbing: device buffers (tensor.data_ptr()), the current CUDA stream and
pinned host staging.  There is no CPU fallback: if the library is missing, or no CUDA device is
present, the calls below raise.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UPMIX_B200_LIB") or os.path.join(_HERE, "csrc", "libupmix_b200.so")

OUT_LSCRS = 0
OUT_FOLD = 1
PLAN_NO_DECIMATE = 1      # upmix_plan_create_ex flag: band-limited bands keep the full-size transform kernels
PLAN_NO_BATCH = 2         # dense 256/512/1024-point bands keep the one-frame-per-CTA kernel (the one block streaming runs)
PLAN_STREAM_KERNELS = 3   # both: offline results bit-identical to block streaming

FUSED_MAX_N = 8192
LARGE_MAX_N = 65536
MIN_N = 64


class UpmixNativeError(RuntimeError):
    pass


class _BandDesc(ctypes.Structure):
    _fields_ = [("n_fft", ctypes.c_int32), ("hop", ctypes.c_int32), ("ana", ctypes.c_void_p),
                ("syn", ctypes.c_void_p), ("gain", ctypes.c_void_p)]


_lib = None


def _sig(fn, restype, argtypes):
    fn.restype = restype
    fn.argtypes = argtypes


def load_library():
    """Load libupmix_b200.so (built by `__graft_entry__.build()` / csrc/build.sh)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise UpmixNativeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or upmix_b200/csrc/build.sh).  upmix_b200 has no CPU path.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    _sig(lib.upmix_last_error, ctypes.c_char_p, [])
    _sig(lib.upmix_version, i32, [])
    _sig(lib.upmix_plan_create, i32, [i32, ctypes.POINTER(_BandDesc), i32, i32, ctypes.POINTER(vp)])
    _sig(lib.upmix_plan_create_ex, i32, [i32, ctypes.POINTER(_BandDesc), i32, i32, i32, ctypes.POINTER(vp)])
    _sig(lib.upmix_plan_destroy, i32, [vp])
    _sig(lib.upmix_plan_n_bands, i32, [vp])
    _sig(lib.upmix_plan_n_pipelines, i32, [vp])
    _sig(lib.upmix_workspace_bytes, i64, [vp, i64, i32])
    _sig(lib.upmix_segment_halo, i64, [vp])
    _sig(lib.upmix_process, i32, [vp, vp, vp, i64, i32, i64, vp, vp, vp, i64, vp, i64, vp])
    _sig(lib.upmix_process_segment, i32,
         [vp, vp, vp, i64, i64, i64, i64, i64, i32, i64, vp, vp, vp, i64, vp, i64, vp])
    _sig(lib.upmix_stream_state_bytes, i64, [vp, i32])
    _sig(lib.upmix_stream_workspace_bytes, i64, [vp, i32, i32])
    _sig(lib.upmix_stream_delay, i64, [vp])
    _sig(lib.upmix_stream_reset, i32, [vp, vp, i32, vp])
    _sig(lib.upmix_stream_block, i32, [vp, vp, i64, vp, vp, i32, i32, i64, vp, vp, vp, i64, vp, i64, vp])
    _sig(lib.upmix_process_host, i32, [vp, vp, vp, i64, vp, vp, vp])
    _sig(lib.upmix_process_host_ex, i32, [vp, vp, vp, i32, i64, i64, i64, vp, vp, vp, i32])
    _sig(lib.upmix_plan_release_host, i32, [vp])
    _sig(lib.upmix_frame_step, i32, [vp, vp, i64, vp, vp, i32, i64, vp, vp, vp, i64, vp, i64, vp])
    _sig(lib.upmix_peak_workspace_bytes, i64, [])
    _sig(lib.upmix_peak3, i32, [vp, vp, vp, i64, vp, vp, i64, vp])
    _sig(lib.upmix_export_mix, i32, [i32, ctypes.c_float, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp])
    _sig(lib.upmix_pcm16_to_planar, i32, [vp, i64, vp, vp, vp, vp, i64, vp])
    _sig(lib.upmix_stereo_to_pcm16, i32, [vp, i64, vp, vp])
    _sig(lib.upmix_fir_filter, i32, [vp, i64, i32, i64, vp, i32, vp, i64, vp])
    _sig(lib.upmix_debug_launch_count, i64, [i32])
    _sig(lib.upmix_measure_fp32_tflops, i32, [i32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i32)])
    _lib = lib
    return lib


EXPORTS = ("upmix_last_error", "upmix_version", "upmix_plan_create", "upmix_plan_create_ex", "upmix_plan_destroy",
           "upmix_plan_n_bands", "upmix_plan_n_pipelines", "upmix_workspace_bytes", "upmix_segment_halo", "upmix_process",
           "upmix_process_segment", "upmix_stream_state_bytes", "upmix_stream_workspace_bytes",
           "upmix_stream_delay", "upmix_stream_reset", "upmix_stream_block", "upmix_process_host",
           "upmix_process_host_ex", "upmix_plan_release_host",
           "upmix_frame_step", "upmix_debug_launch_count", "upmix_measure_fp32_tflops",
           "upmix_peak_workspace_bytes", "upmix_peak3", "upmix_export_mix", "upmix_pcm16_to_planar",
           "upmix_stereo_to_pcm16", "upmix_fir_filter")

EXPORT_MODES = {"AB": 0, "split": 1, "stereo_sum": 2}


def fir_filter(x, taps):
    """y[..., i] = sum_k taps[k] x[..., i-k] (causal, zero history) for a float32 CUDA tensor x [n] or
    [tracks, n] and float32 CUDA taps [n_taps] (upmix_fir_filter)."""
    torch = _torch()
    lib = load_library()
    x2 = x.reshape(-1, x.shape[-1]).contiguous()
    y = torch.empty_like(x2)
    taps = taps.contiguous()
    with torch.cuda.device(x.device):
        _check(lib.upmix_fir_filter(x2.data_ptr(), x2.shape[1], x2.shape[0], x2.shape[1], taps.data_ptr(), taps.numel(),
                                    y.data_ptr(), x2.shape[1], torch.cuda.current_stream(x.device).cuda_stream))
    return y.reshape(x.shape)


def peak3(c, l, r):
    """max|C|, max|Ls|, max|Rs| of three float32 CUDA tensors, as a 3-element CUDA tensor."""
    torch = _torch()
    lib = load_library()
    n = c.numel()
    out = torch.empty(3, dtype=torch.float32, device=c.device)
    wsb = int(lib.upmix_peak_workspace_bytes())
    ws = torch.empty(wsb, dtype=torch.uint8, device=c.device)
    with torch.cuda.device(c.device):
        _check(lib.upmix_peak3(c.data_ptr(), l.data_ptr(), r.data_ptr(), n, out.data_ptr(), ws.data_ptr(), wsb,
                               torch.cuda.current_stream(c.device).cuda_stream))
    return out


def pcm16_to_planar(pcm):
    """pcm: int16 CUDA tensor [n, 2] (interleaved stereo).  Returns (L, R, peak): planar float32 CUDA
    tensors (x / 32768) and a 1-element tensor with max|x| over both channels."""
    torch = _torch()
    lib = load_library()
    if pcm.dtype != torch.int16 or pcm.dim() != 2 or pcm.shape[1] != 2 or not pcm.is_contiguous():
        raise TypeError("pcm must be a contiguous int16 CUDA tensor [n, 2]")
    n = pcm.shape[0]
    out = torch.empty((2, n), dtype=torch.float32, device=pcm.device)
    peak = torch.zeros(1, dtype=torch.float32, device=pcm.device)
    wsb = int(lib.upmix_peak_workspace_bytes())
    ws = torch.empty(wsb, dtype=torch.uint8, device=pcm.device)
    with torch.cuda.device(pcm.device):
        _check(lib.upmix_pcm16_to_planar(pcm.data_ptr(), n, out[0].data_ptr(), out[1].data_ptr(), peak.data_ptr(),
                                         ws.data_ptr(), wsb, torch.cuda.current_stream(pcm.device).cuda_stream))
    return out[0], out[1], peak


def stereo_to_pcm16(stereo):
    """stereo: float32 CUDA tensor [n, 2].  Returns the int16 CUDA tensor [n, 2] a 16-bit WAV stores."""
    torch = _torch()
    lib = load_library()
    if stereo.dtype != torch.float32 or stereo.dim() != 2 or stereo.shape[1] != 2 or not stereo.is_contiguous():
        raise TypeError("stereo must be a contiguous float32 CUDA tensor [n, 2]")
    out = torch.empty(stereo.shape, dtype=torch.int16, device=stereo.device)
    with torch.cuda.device(stereo.device):
        _check(lib.upmix_stereo_to_pcm16(stereo.data_ptr(), stereo.shape[0], out.data_ptr(),
                                         torch.cuda.current_stream(stereo.device).cuda_stream))
    return out


def export_mix(mode: str, scale: float, c, l, r, in_l=None, in_r=None):
    """Scaled export mix of main.py as interleaved stereo float32 CUDA tensors [n, 2] (one tensor,
    three for "split")."""
    torch = _torch()
    lib = load_library()
    if mode not in EXPORT_MODES:
        raise ValueError(f"unknown export mode {mode!r}")
    n = c.numel()
    outs = [torch.empty((n, 2), dtype=torch.float32, device=c.device) for _ in range(3 if mode == "split" else 1)]
    ptr = [o.data_ptr() for o in outs] + [None, None]
    with torch.cuda.device(c.device):
        _check(lib.upmix_export_mix(EXPORT_MODES[mode], float(scale), c.data_ptr(), l.data_ptr(), r.data_ptr(),
                                    in_l.data_ptr() if in_l is not None else None,
                                    in_r.data_ptr() if in_r is not None else None, n, ptr[0], ptr[1], ptr[2],
                                    torch.cuda.current_stream(c.device).cuda_stream))
    return outs


def launch_count(reset: bool = False) -> int:
    """Kernels launched by the library since the last reset."""
    return int(load_library().upmix_debug_launch_count(1 if reset else 0))


def measure_fp32_tflops(device: int = 0):
    """(TFLOP/s, SM count) of a pure FMA kernel on `device`."""
    _torch()
    v, n = ctypes.c_double(), ctypes.c_int()
    _check(load_library().upmix_measure_fp32_tflops(int(device), ctypes.byref(v), ctypes.byref(n)))
    return float(v.value), int(n.value)


def _check(rc: int):
    if rc < 0:
        raise UpmixNativeError(f"upmix_b200 error {rc}: {load_library().upmix_last_error().decode()}")
    return rc


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise UpmixNativeError("no CUDA device: upmix_b200 runs on B200 (sm_100a) only and has no CPU path")
    return torch


class Plan:
    """An immutable set of bands on one device (UpmixPlan).  `bands` is a sequence of
    (n_fft, hop, ana[n_fft], syn[n_fft], gain[n_fft/2+1]) with float32-convertible tables."""

    def __init__(self, bands: Sequence[Tuple[int, int, np.ndarray, np.ndarray, np.ndarray]],
                 out_mode: int = OUT_LSCRS, device: Optional[int] = None, flags: int = 0):
        torch = _torch()
        lib = load_library()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.out_mode = out_mode
        self.sizes = [int(b[0]) for b in bands]
        self.hops = [int(b[1]) for b in bands]
        keep = []
        descs = (_BandDesc * len(bands))()
        for i, (n_fft, hop, ana, syn, gain) in enumerate(bands):
            ana = np.ascontiguousarray(ana, dtype=np.float32)
            syn = np.ascontiguousarray(syn, dtype=np.float32)
            gain = np.ascontiguousarray(gain, dtype=np.float32)
            if ana.shape != (n_fft,) or syn.shape != (n_fft,) or gain.shape != (n_fft // 2 + 1,):
                raise ValueError(f"band {i}: table shapes do not match n_fft={n_fft}")
            keep += [ana, syn, gain]
            descs[i] = _BandDesc(n_fft, hop, ana.ctypes.data, syn.ctypes.data, gain.ctypes.data)
        handle = ctypes.c_void_p()
        _check(lib.upmix_plan_create_ex(len(bands), descs, out_mode, self.device, int(flags), ctypes.byref(handle)))
        self.flags = int(flags)
        self._h = handle
        self._lib = lib
        self._ws = None
        self.halo = int(lib.upmix_segment_halo(self._h))
        self.n_pipelines = int(lib.upmix_plan_n_pipelines(self._h))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.upmix_plan_destroy(h)
            except Exception:
                pass

    # -- workspace -----------------------------------------------------------------------------
    def workspace_bytes(self, seg_len: int, n_tracks: int = 1) -> int:
        return _check(self._lib.upmix_workspace_bytes(self._h, int(seg_len), int(n_tracks)))

    def _workspace(self, nbytes: int):
        torch = _torch()
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=f"cuda:{self.device}")
        return self._ws

    def release_workspace(self):
        self._ws = None

    # -- device-resident processing ------------------------------------------------------------
    def process(self, L, R, out=None):
        """L, R: float32 CUDA tensors [n] or [tracks, n].  Returns (C, Ls, Rs) -- or (L', R') in
        fold-down mode -- as tensors of the same shape.  Asynchronous on the current stream."""
        return self.process_segment(L, R, 0, L.shape[-1], 0, L.shape[-1], out=out)

    def process_segment(self, L, R, in_begin: int, n_total: int, seg_begin: int, seg_end: int, out=None):
        torch = _torch()
        if L.dtype != torch.float32 or R.dtype != torch.float32 or not L.is_cuda or not R.is_cuda:
            raise TypeError("L and R must be float32 CUDA tensors")
        if L.shape != R.shape or L.dim() not in (1, 2):
            raise ValueError("L and R must have the same shape [n] or [tracks, n]")
        if L.device.index != self.device:
            raise ValueError(f"tensors are on {L.device}, plan is on cuda:{self.device}")
        squeeze = L.dim() == 1
        if squeeze:
            L, R = L[None], R[None]
        if L.stride(1) != 1 or R.stride(1) != 1 or (L.shape[0] > 1 and L.stride(0) != R.stride(0)):
            L, R = L.contiguous(), R.contiguous()
        tracks, in_len = L.shape
        seg_len = seg_end - seg_begin
        n_out = 3 if self.out_mode == OUT_LSCRS else 2
        if out is None:
            out = torch.empty((n_out, tracks, seg_len), dtype=torch.float32, device=L.device)
        elif (out.shape != (n_out, tracks, seg_len) or out.dtype != torch.float32 or
              not all(out[i].is_contiguous() for i in range(n_out))):
            raise ValueError(f"out must be a float32 tensor of shape {(n_out, tracks, seg_len)} with contiguous channels")
        oc, ol, orr = (out[0], out[1], out[2]) if n_out == 3 else (None, out[0], out[1])
        if seg_len == 0:
            return tuple(o[0] for o in out) if squeeze else tuple(out[i] for i in range(n_out))
        wsb = self.workspace_bytes(seg_len, tracks)
        ws = self._workspace(wsb)
        stream = torch.cuda.current_stream(L.device).cuda_stream
        with torch.cuda.device(L.device):
            _check(self._lib.upmix_process_segment(
                self._h, L.data_ptr(), R.data_ptr(), int(in_begin), int(in_len), int(n_total), int(seg_begin),
                int(seg_end), tracks, L.stride(0) if tracks > 1 else in_len,
                oc.data_ptr() if oc is not None else None, ol.data_ptr(), orr.data_ptr(), seg_len,
                ws.data_ptr(), wsb, stream))
        res = tuple(o[0] for o in out) if squeeze else tuple(out[i] for i in range(n_out))
        return res

    def frame_step(self, ring, frame_index: int, blk_l, blk_r):
        """Single-band plans: one frame of the stateful chunk API (upmix_frame_step).  ring: float32
        CUDA tensor [3, n_fft]; blk_l / blk_r: float32 CUDA tensors [n_fft].  Returns (C, Ls, Rs) [hop]."""
        torch = _torch()
        hop = self.hops[0]
        out = torch.empty((3, hop), dtype=torch.float32, device=ring.device)
        wsb = self.workspace_bytes(hop, 1)
        ws = self._workspace(wsb)
        stream = torch.cuda.current_stream(ring.device).cuda_stream
        blk_l, blk_r = blk_l.contiguous(), blk_r.contiguous()
        with torch.cuda.device(ring.device):
            _check(self._lib.upmix_frame_step(self._h, ring.data_ptr(), int(frame_index), blk_l.data_ptr(),
                                              blk_r.data_ptr(), 1, self.sizes[0], out[0].data_ptr(), out[1].data_ptr(),
                                              out[2].data_ptr(), hop, ws.data_ptr(), wsb, stream))
        return out[0], out[1], out[2]

    # -- block streaming -------------------------------------------------------------------------
    def stream_open(self, n_tracks: int = 1):
        return Stream(self, n_tracks)

    # -- host buffers: upmix_process_host_ex (chunked convert / H2D / kernels / D2H pipeline in the library) -------
    def release_host(self):
        """Free the device buffers and pinned staging the host-buffer calls cache in the plan."""
        _check(self._lib.upmix_plan_release_host(self._h))

    def _host_call(self, lp: int, rp: int, dtype: int, sl: int, sr: int, n: int, n_threads: int = 0, pinned_out: bool = True):
        """Runs the host pipeline into FRESH outputs: pinned tensors (torch's caching host allocator hands back blocks
        of results that have been dropped, never one that is still referenced; a first-time block costs ~0.6 s per
        GB to pin) or plain pageable memory (filled by the library's copy-out workers)."""
        torch = _torch()
        n_out = 3 if self.out_mode == OUT_LSCRS else 2
        if pinned_out:
            out = torch.empty((n_out, n), dtype=torch.float32, pin_memory=True)
        else:
            out = torch.from_numpy(np.empty((n_out, n), dtype=np.float32))
        ptrs = [out[i].data_ptr() for i in range(n_out)]
        if n_out == 2:
            ptrs = [None] + ptrs
        with torch.cuda.device(self.device):
            _check(self._lib.upmix_process_host_ex(self._h, lp, rp, dtype, sl, sr, n, ptrs[0], ptrs[1], ptrs[2], int(n_threads)))
        return out

    def process_host_tensors(self, L, R, n_threads: int = 0):
        """L, R: 1-D float32 / float64 CPU torch tensors (any stride; pinned float32 is used in place).  Returns
        freshly allocated pinned float32 CPU tensors; bit-identical to a device-resident call."""
        torch = _torch()
        if L.dtype not in (torch.float32, torch.float64) or R.dtype != L.dtype or L.dim() != 1 or L.shape != R.shape:
            raise TypeError("L and R must be 1-D float32 or float64 CPU tensors of equal length")
        n = L.shape[0]
        n_out = 3 if self.out_mode == OUT_LSCRS else 2
        if n == 0:
            return tuple(torch.empty(0, dtype=torch.float32) for _ in range(n_out))
        if L.stride(0) < 1 or R.stride(0) < 1:
            L, R = L.contiguous(), R.contiguous()
        out = self._host_call(L.data_ptr(), R.data_ptr(), 0 if L.dtype == torch.float32 else 1, L.stride(0), R.stride(0), n,
                              n_threads)
        return tuple(out[i] for i in range(n_out))

    def process_host(self, L: np.ndarray, R: np.ndarray, n_threads: int = 0, pinned_out: Optional[bool] = None):
        """numpy in (float32 or float64, any positive stride -- e.g. the two columns of the interleaved float64
        array sf.read returns, main.py:43-50), fresh float32 numpy arrays out, like the reference's (CE:503-513).
        pinned_out: None = UPMIX_NUMPY_PINNED_OUT (default 0: plain pageable arrays)."""
        if pinned_out is None:
            pinned_out = os.environ.get("UPMIX_NUMPY_PINNED_OUT", "0") not in ("", "0")
        _torch()
        L, R = np.asarray(L), np.asarray(R)
        if L.ndim != 1 or L.shape != R.shape:
            raise ValueError("L and R must be 1-D arrays of equal length")
        if L.dtype != R.dtype or L.dtype not in (np.float32, np.float64):
            L, R = np.ascontiguousarray(L, dtype=np.float32), np.ascontiguousarray(R, dtype=np.float32)
        item = L.dtype.itemsize
        if L.shape[0] and (L.strides[0] < item or R.strides[0] < item or L.strides[0] % item or R.strides[0] % item):
            L, R = np.ascontiguousarray(L), np.ascontiguousarray(R)
        n = L.shape[0]
        n_out = 3 if self.out_mode == OUT_LSCRS else 2
        if n == 0:
            return tuple(np.zeros(0, dtype=np.float32) for _ in range(n_out))
        out = self._host_call(L.ctypes.data, R.ctypes.data, 0 if L.dtype == np.float32 else 1, L.strides[0] // item,
                              R.strides[0] // item, n, n_threads, pinned_out)
        return tuple(out[i].numpy() for i in range(n_out))


class Stream:
    """Block-by-block processing with carried state (upmix_stream_block)."""

    def __init__(self, plan: Plan, n_tracks: int = 1):
        torch = _torch()
        self.plan = plan
        self.n_tracks = n_tracks
        lib = plan._lib
        nbytes = _check(lib.upmix_stream_state_bytes(plan._h, n_tracks))
        self.state = torch.zeros(int(nbytes), dtype=torch.uint8, device=f"cuda:{plan.device}")
        self.delay = int(lib.upmix_stream_delay(plan._h))
        self.samples_done = 0
        self._ws = None

    def reset(self):
        self.state.zero_()
        self.samples_done = 0

    def block(self, in_l, in_r):
        """in_l, in_r: float32 CUDA tensors [n_new] or [tracks, n_new].  Returns the next n_new output
        samples of every output channel (delayed by self.delay)."""
        torch = _torch()
        plan, lib = self.plan, self.plan._lib
        squeeze = in_l.dim() == 1
        if squeeze:
            in_l, in_r = in_l[None], in_r[None]
        in_l, in_r = in_l.contiguous(), in_r.contiguous()
        tracks, n_new = in_l.shape
        if tracks != self.n_tracks:
            raise ValueError("track count differs from the stream's")
        n_out = 3 if plan.out_mode == OUT_LSCRS else 2
        out = torch.empty((n_out, tracks, n_new), dtype=torch.float32, device=in_l.device)
        oc, ol, orr = (out[0], out[1], out[2]) if n_out == 3 else (None, out[0], out[1])
        wsb = _check(lib.upmix_stream_workspace_bytes(plan._h, n_new, tracks))
        if self._ws is None or self._ws.numel() < wsb:
            self._ws = torch.empty(int(wsb), dtype=torch.uint8, device=in_l.device)
        stream = torch.cuda.current_stream(in_l.device).cuda_stream
        with torch.cuda.device(in_l.device):
            _check(lib.upmix_stream_block(plan._h, self.state.data_ptr(), self.samples_done, in_l.data_ptr(),
                                          in_r.data_ptr(), n_new, tracks, n_new,
                                          oc.data_ptr() if oc is not None else None, ol.data_ptr(), orr.data_ptr(),
                                          n_new, self._ws.data_ptr(), wsb, stream))
        self.samples_done += n_new
        return tuple(o[0] for o in out) if squeeze else tuple(out[i] for i in range(n_out))
