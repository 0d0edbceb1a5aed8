#!/usr/bin/env python
"""Per-section cycle shares of analysis_kernel, from a -DPNB_ANA_TIMING build of pnb_dsp.cu
(percepnet_b200/libpercepnet_b200_timing.so, built by `python percepnet_b200/build.py --timing`)."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from percepnet_b200 import api  # noqa: E402

NAMES = ["look-ahead FFT (window+transform)", "ring store + band pool", "pitch_downsample decimate", "autocorr 5 lags (5 lanes)",
         "LPC + FIR coefficients (1 lane)", "5-tap FIR", "coarse xcorr 147 lags", "find_best_pitch coarse (1 lane)",
         "fine xcorr + Syy + xx (12 lanes)", "find_best_pitch fine + interp (1 lane)", "remove_doubling dots + yy table",
         "remove_doubling decisions (1 lane)", "final 3-lag refinement (3 lanes)", "comb filter + FFT of P",
         "P store + band pools Ep, Exp", "features + outputs"]


def main():
    api.LIB_PATH = os.path.join(ROOT, "percepnet_b200", "libpercepnet_b200_timing.so")
    from percepnet_b200.synth import synth_pairs
    N, F = 8192, 8
    c, n = synth_pairs(32, F * 3, seed=2024)
    idx = np.arange(N) % 32
    eng = api.Engine(2 * N, F, None, api.TRAIN_DATA)
    L = eng.L
    out = (C.c_ulonglong * 16)()
    for k in range(3):
        eng.train_records(np.ascontiguousarray(c[idx, k * F * 480:(k + 1) * F * 480]), np.ascontiguousarray(n[idx, k * F * 480:(k + 1) * F * 480]))
        if k == 0:
            L.pnb_debug_analysis_cycles(out, 1)       # drop the cold call
    L.pnb_debug_analysis_cycles(out, 1)
    cyc = np.array(list(out), dtype=np.float64)
    frames = 2 * (2 * N) * F
    res = {NAMES[i]: {"cycles_per_frame": cyc[i] / frames, "share": cyc[i] / cyc.sum()} for i in range(16)}
    print(json.dumps({"warp_cycles_per_frame": cyc.sum() / frames, "sections": res}, indent=1))


if __name__ == "__main__":
    main()
