"""Shared parity bar (BASELINE.json north_star): >= 100 dB SNR per output channel against the
reference's output and max abs error <= 1e-5 of full scale.  The bound is applied relative to the
input peak (never above digital full scale 1.0), which is stricter than full scale."""
import numpy as np

from oracle.upmix_oracle import snr_db

MIN_SNR_DB = 100.0
MAX_ABS_ERR = 1e-5


def assert_parity(ref_channels, got_channels, peak=1.0, names=("C", "Ls", "Rs"), what=""):
    report = []
    for name, ref, got in zip(names, ref_channels, got_channels):
        ref = np.asarray(ref)
        got = np.asarray(got)
        assert ref.shape == got.shape, f"{what} {name}: shape {got.shape} vs {ref.shape}"
        assert np.all(np.isfinite(got)), f"{what} {name}: non-finite output"
        snr = snr_db(ref, got)
        err = float(np.max(np.abs(ref.astype(np.float64) - got.astype(np.float64)))) if ref.size else 0.0
        report.append((name, snr, err))
        if float(np.max(np.abs(ref))) > 0:
            assert snr >= MIN_SNR_DB, f"{what} {name}: SNR {snr:.1f} dB < {MIN_SNR_DB} dB (max err {err:.3g})"
        assert err <= MAX_ABS_ERR * min(1.0, peak), f"{what} {name}: max abs err {err:.3g} > {MAX_ABS_ERR * min(1.0, peak):.3g}"
    return report
