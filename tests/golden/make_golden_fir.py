#!/usr/bin/env python3
"""Golden fixture for the FIR row FROM THE UNMODIFIED REFERENCE (python-prototype/filter_design.py).
Run in the development container only (needs /root/reference):  python tests/golden/make_golden_fir.py
`ref_*` arrays are outputs of the reference's own functions; scipy / numpy versions are recorded."""
import importlib.util
import os
import sys

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import upmix_oracle as uo            # noqa: E402  (only for synth_stereo)

spec = importlib.util.spec_from_file_location("ref_filter_design", "/root/reference/python-prototype/filter_design.py")
fd = importlib.util.module_from_spec(spec)
spec.loader.exec_module(fd)

sr = 48000
L, R = uo.synth_stereo(30000, 5, stress=True)
hp = fd.design_lr4_hp_fir(sr, 180.0, 1025)
lp = fd.design_lr4_lp_fir(sr, 180.0, 1025)
hp_short = fd.design_lr4_hp_fir(44100, 2000.0, 129)
np.savez_compressed(os.path.join(HERE, "fir.npz"), in_L=L, in_R=R, sr=np.array(sr),
                    ref_hp=hp, ref_lp=lp, ref_hp_short=hp_short, ref_pass=fd.design_lr4_lp_fir(sr, 0.0),
                    ref_y_hp=fd.apply_fir_filter(L.astype(np.float64), hp), ref_y_lp=fd.apply_fir_filter(R.astype(np.float64), lp),
                    ref_y_short=fd.apply_fir_filter(L, hp_short),
                    numpy_version=np.array(np.__version__), scipy_version=np.array(scipy.__version__))
print("wrote fir.npz", hp.dtype, fd.apply_fir_filter(L, hp_short).dtype)
