"""WAV input/output for main.py.  The reference uses `soundfile` (main.py:43,119); when it is not
installed this falls back to scipy.io.wavfile / the standard library with the same conventions:
reading returns float64 in [-1, 1) shaped [frames, channels] (soundfile's default dtype), writing
stores 16-bit PCM (soundfile's default subtype for .wav)."""
from __future__ import annotations

import numpy as np


def read_wav(path: str):
    try:
        import soundfile as sf
        return sf.read(path)
    except ImportError:
        pass
    from scipy.io import wavfile
    sr, data = wavfile.read(path)
    if data.dtype == np.int16:
        wave = data.astype(np.float64) / 32768.0
    elif data.dtype == np.int32:
        wave = data.astype(np.float64) / 2147483648.0
    elif data.dtype == np.uint8:
        wave = (data.astype(np.float64) - 128.0) / 128.0
    else:
        wave = data.astype(np.float64)
    return wave, int(sr)


def read_wav_pcm16(path: str):
    """(int16 array [frames, 2], sr) when the file is 16-bit PCM stereo and can be read without
    conversion (scipy / stdlib path), else None: main.py then takes the float route."""
    try:
        from scipy.io import wavfile
        sr, data = wavfile.read(path)
    except Exception:
        return None
    if data.dtype != np.int16 or data.ndim != 2 or data.shape[1] != 2:
        return None
    return np.ascontiguousarray(data), int(sr)


def write_wav_pcm16(path: str, pcm: np.ndarray, sr: int) -> None:
    """Store int16 samples [frames, channels] as a 16-bit PCM WAV."""
    try:
        import soundfile as sf
        sf.write(path, pcm, sr, subtype="PCM_16")
        return
    except ImportError:
        pass
    from scipy.io import wavfile
    wavfile.write(path, int(sr), np.ascontiguousarray(pcm, dtype=np.int16))


def write_wav(path: str, data: np.ndarray, sr: int) -> None:
    try:
        import soundfile as sf
        sf.write(path, data, sr)
        return
    except ImportError:
        pass
    from scipy.io import wavfile
    # libsndfile's float -> PCM_16 conversion (what sf.write does with float data): x * 0x7FFF, round to nearest even
    pcm = np.clip(np.rint(np.asarray(data, dtype=np.float64) * 32767.0), -32768, 32767)
    wavfile.write(path, int(sr), pcm.astype(np.int16))
