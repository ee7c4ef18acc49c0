"""Drop-in for the prototype's filter_design.py (python-prototype/filter_design.py): FIR approximations
of 4th-order Linkwitz-Riley high-/low-pass filters and their application.

Same names, arguments and return types.  The tap design is host-side (scipy.signal.firwin with a
Hamming window, exactly the reference's call, filter_design.py:36-37, 50-51); applying the taps
(`scipy.signal.lfilter(taps, 1.0, wave)`, filter_design.py:59) runs on the GPU (`upmix_fir_filter`,
float32).  There is no CPU path: without a CUDA device `apply_fir_filter` raises.  Like the prototype's
module, nothing on the centre-extraction path imports this.
"""
import numpy as np

from . import _native


def design_lr4_hp_fir(sr: float, cutoff_hz: float = 180.0, numtaps: int = 1025) -> np.ndarray:
    """Approximate 4th-order Linkwitz-Riley high-pass at `cutoff_hz`; [1.0] (pass-through) if
    cutoff_hz <= 0 (filter_design.py:25-38)."""
    if cutoff_hz <= 0:
        return np.array([1.0], dtype=np.float32)
    from scipy.signal import firwin
    return firwin(numtaps, cutoff_hz / (0.5 * sr), pass_zero=False, window="hamming").astype(np.float32)


def design_lr4_lp_fir(sr: float, cutoff_hz: float = 180.0, numtaps: int = 1025) -> np.ndarray:
    """Approximate 4th-order Linkwitz-Riley low-pass at `cutoff_hz`; [1.0] if cutoff_hz <= 0
    (filter_design.py:40-52)."""
    if cutoff_hz <= 0:
        return np.array([1.0], dtype=np.float32)
    from scipy.signal import firwin
    return firwin(numtaps, cutoff_hz / (0.5 * sr), pass_zero=True, window="hamming").astype(np.float32)


def apply_fir_filter(wave, fir_taps):
    """`lfilter(fir_taps, 1.0, wave)` (filter_design.py:54-59): causal convolution along the last axis,
    same shape as `wave`.  numpy in -> float64 numpy out (what lfilter returns for real input here);
    a CUDA tensor in -> a float32 CUDA tensor out.  The arithmetic is float32 on the device."""
    torch = _native._torch()
    if isinstance(wave, torch.Tensor) and wave.is_cuda:
        taps = torch.as_tensor(np.asarray(fir_taps, dtype=np.float32)).to(wave.device) if not isinstance(fir_taps, torch.Tensor) \
            else fir_taps.to(wave.device, torch.float32)
        return _native.fir_filter(wave.to(torch.float32), taps)
    if not torch.cuda.is_available():
        raise _native.UpmixNativeError("upmix_b200 has no CPU path: apply_fir_filter needs a CUDA device")
    w = np.asarray(wave)
    t = np.asarray(fir_taps)
    out_dtype = np.float64
    if w.size == 0:
        return np.zeros(w.shape, dtype=out_dtype)
    y = _native.fir_filter(torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32)).cuda(),
                           torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32)).cuda())
    return y.cpu().numpy().astype(out_dtype)
