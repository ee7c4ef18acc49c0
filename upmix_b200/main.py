#!/usr/bin/env python3
"""Upmix driver, same surface as the reference's python-prototype/main.py (cited MP:line):

1) loads a stereo WAV from 'in/' (MP:39-50),
2) builds the multi-band extractors from the crossover list (MP:62-73) and extracts Ls, C, Rs on the
   GPU (MP:78-80),
3) scales Ls, C, Rs by one factor so that none exceeds the input peak (MP:85-97),
4) writes, by export_mode: "AB" (Left = Ls+C+Rs, Right = L+R), "split" (three stereo files) or
   "stereo_sum" (Left = Ls + C/2, Right = Rs + C/2) with the reference's file names (MP:110-157).

Edit the constants in main() -- or pass keyword arguments to run() -- and run `python -m upmix_b200.main`.
"""
import os

import numpy as np

from . import center_extraction as ce
from .wavio import read_wav, read_wav_pcm16, write_wav, write_wav_pcm16


def run(in_filename="eyes.wav", export_mode="stereo_sum", in_dir="in", out_dir="out",
        band_edges=(0, 30, 120, 480, 1920, 7680), overlap=0.75, window_func=None, xover_mode="raised_cosine",
        max_block_size=2 ** 16, threshold_factor=32, xo_fraction=None):
    """The body of the reference's main() with its user-config constants as parameters.
    Returns the list of files written."""
    window_func = window_func or ce.make_blackman_harris
    os.makedirs(out_dir, exist_ok=True)

    in_path = os.path.join(in_dir, in_filename)
    if not os.path.isfile(in_path):
        raise FileNotFoundError(f"File not found: {in_path}")
    # 16-bit PCM stereo (what soundfile writes by default) goes to the device as int16 and is converted
    # there (x/32768, like soundfile's float view); anything else takes the float route.
    from . import _native
    torch = _native._torch()
    pcm = read_wav_pcm16(in_path)
    if pcm is not None:
        pcm_data, sr = pcm
        print(f"Loaded '{in_path}', sr={sr}, shape={pcm_data.shape}")
        if pcm_data.shape[0]:
            dL, dR, dpeak = _native.pcm16_to_planar(torch.from_numpy(pcm_data).cuda())
            peak_in = float(dpeak.cpu()[0])
        else:
            dL = dR = torch.zeros(0, dtype=torch.float32, device="cuda")
            peak_in = 0.0
    else:
        wave, sr = read_wav(in_path)
        print(f"Loaded '{in_path}', sr={sr}, shape={wave.shape}")
        if wave.ndim == 1:
            wave = np.column_stack([wave, wave])
        peak_in = float(np.max(np.abs(wave))) if wave.size else 0.0
        dL = torch.from_numpy(np.ascontiguousarray(wave[:, 0], dtype=np.float32)).cuda()
        dR = torch.from_numpy(np.ascontiguousarray(wave[:, 1], dtype=np.float32)).cuda()
    if peak_in <= 0.0:
        peak_in = 1e-9

    band_extractors = ce.chain_bands(list(band_edges), overlap=overlap, window_func=window_func, sr=sr,
                                     xover_mode=xover_mode, max_block_size=max_block_size,
                                     threshold_factor=threshold_factor, xo_fraction=xo_fraction)
    # Extraction, peak measurement and the export mix all stay on the device; only the three peaks and
    # the finished 16-bit stereo file(s) cross PCIe (main.py:78-97, 110-157 of the reference do this in numpy).
    final_center, final_left, final_right = ce.extract_center_left_right_multi_band_in_memory(dL, dR, sr, band_extractors)

    if dL.numel():
        peak_c, peak_l, peak_r = (float(v) for v in _native.peak3(final_center, final_left, final_right).cpu())
    else:
        peak_c = peak_l = peak_r = 0.0
    overall_peak = max(peak_l, peak_c, peak_r, 1e-9)
    scale_factor = peak_in / overall_peak
    print(f"Original peak = {peak_in:.4f}, L/C/R peak = {overall_peak:.4f}")
    print(f"Applying scale_factor = {scale_factor:.4f}")

    band_info_str = "_".join(f"b{bex.block_size}({int(bex.f_low)}-{int(bex.f_high)})" for bex in band_extractors)
    base_in_name = os.path.splitext(in_filename)[0]
    written = []

    def mix(mode):       # int16 [n, 2] arrays, converted on the device like soundfile's PCM_16 writer
        if dL.numel() == 0:
            return [np.zeros((0, 2), dtype=np.int16)] * (3 if mode == "split" else 1)
        outs = _native.export_mix(mode, scale_factor, final_center, final_left, final_right, dL, dR)
        return [_native.stereo_to_pcm16(o).cpu().numpy() for o in outs]

    if export_mode == "AB":
        out_path = os.path.join(out_dir, f"{base_in_name}_AB_{band_info_str}_ov{overlap:.2f}.wav")
        write_wav_pcm16(out_path, mix("AB")[0], sr)
        written.append(out_path)
        print(f"[AB] Wrote 2-ch => {out_path}\n  Left  = (Ls + C + Rs)\n  Right = (L + R)\n")
    elif export_mode == "split":
        for tag, stereo, what in zip(("Ls", "C", "Rs"), mix("split"), ("Left=Ls, Right=0", "Left=C, Right=C", "Left=0, Right=Rs")):
            path = os.path.join(out_dir, f"{base_in_name}_{tag}_{band_info_str}.wav")
            write_wav_pcm16(path, stereo, sr)
            written.append(path)
            print(f"[Split] Wrote => {path} ({what})")
    elif export_mode == "stereo_sum":
        out_path = os.path.join(out_dir, f"{base_in_name}_Sum_{band_info_str}_ov{overlap:.2f}.wav")
        write_wav_pcm16(out_path, mix("stereo_sum")[0], sr)
        written.append(out_path)
        print(f"[StereoSum] Wrote 2-ch => {out_path}\n  Left  = (Ls + C/2)\n  Right = (Rs + C/2)\n")
    else:
        print(f"Unknown export_mode '{export_mode}' -- no files written.")
    print("Done.")
    return written


def main():
    # --------------------------------------------------------------------------
    # User-Config: adjust these as needed (MP:29-34, 62-65)
    # --------------------------------------------------------------------------
    in_filename = "eyes.wav"        # WAV name in the 'in/' folder
    export_mode = "stereo_sum"      # "AB", "split", or "stereo_sum"
    in_dir = "in"
    out_dir = "out"
    band_edges = [0, 30, 120, 480, 1920, 7680]
    overlap = 0.75
    window_func = ce.make_blackman_harris
    run(in_filename, export_mode, in_dir, out_dir, band_edges, overlap, window_func)


if __name__ == "__main__":
    main()
