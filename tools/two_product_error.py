#!/usr/bin/env python
"""What a two-product split of the GRU contractions would cost in accuracy (CPU simulation, float64).

The tensor path forms every fp32-accurate product from three half-precision MMAs: xh*wh + xh*wl + xl*wh.  Dropping one of
them would raise the ceiling of the roofline fraction from 1/3 to 1/2.  This script evaluates the network in double
precision on the oracle's features of the test signals, once in full, once with the activations entering the GRU and
output-layer contractions rounded to one fp16 term (= the xl*wh product dropped; the inputs are tanh / sigmoid bounded) and
once with the weights rounded to one (scaled) fp16 term (= xh*wl dropped), and prints the relative distance of g / r
from the full evaluation -- to be read against the 1e-4 bar and the 6.6e-6 the three-product path measures.

    python tools/two_product_error.py            # needs the oracle (make -C oracle), no GPU"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ffi  # noqa: E402
from percepnet_b200.synth import synth_pcm  # noqa: E402
from percepnet_b200.weights import synth_model  # noqa: E402
from util import edge_signals  # noqa: E402


def f16(x):
    return np.asarray(x, np.float64).astype(np.float16).astype(np.float64)


def f16_scaled(w):   # one fp16 term after a power-of-two scale that brings max |w| just below 2 (as tc_prepare scales)
    m = np.abs(w).max()
    k = 2.0 ** np.floor(np.log2(1.999 / m)) if m > 0 else 1.0
    return f16(w * k) / k


def sig(x):
    return 1.0 / (1.0 + np.exp(-x))


def run(A, feats, rx=lambda v: v, rw=lambda v: v):
    """feats [F, 70] -> g [F, 34], r [F, 34]; rx / rw round the GRU and output-layer operands"""
    W = {k: np.asarray(v, np.float64) for k, v in A.items()}
    Wr = {k: (rw(v) if k.endswith("weights") and (k.startswith("gru") or k.startswith("fc_gb") or k.startswith("fc_rb")) else v)
          for k, v in W.items()}
    fc_hist = np.zeros((4, 128)); c1_hist = np.zeros((2, 512))
    h = {n: np.zeros(512) for n in ("gru1", "gru2", "gru3", "gru_gb")}
    h["gru_rb"] = np.zeros(128)
    G, R = [], []

    def gru(name, x, H):
        Wx, U, b = Wr[name + "_weights"], Wr[name + "_recurrent_weights"], W[name + "_bias"]
        xs, hs = rx(x) @ Wx, rx(h[name]) @ U
        z = sig(xs[:H] + hs[:H] + b[:H] + b[3 * H:4 * H])
        r = sig(xs[H:2 * H] + hs[H:2 * H] + b[H:2 * H] + b[4 * H:5 * H])
        n = np.tanh(xs[2 * H:] + b[2 * H:3 * H] + r * (hs[2 * H:] + b[5 * H:]))
        h[name] = z * h[name] + (1 - z) * n
        return h[name]
    for f in np.asarray(feats, np.float64):
        fc = np.maximum(f @ W["fc_weights"] + W["fc_bias"], 0.0)
        taps1 = np.concatenate([fc_hist.reshape(-1), fc])            # oldest tap first (nnet.cpp:182-200)
        c1 = np.maximum(taps1 @ W["conv1_weights"].reshape(5 * 128, 512) + W["conv1_bias"], 0.0)
        taps2 = np.concatenate([c1_hist.reshape(-1), c1])
        c2 = np.tanh(taps2 @ W["conv2_weights"].reshape(3 * 512, 512) + W["conv2_bias"])
        fc_hist = np.vstack([fc_hist[1:], fc]); c1_hist = np.vstack([c1_hist[1:], c1])
        g1 = gru("gru1", c2, 512); g2 = gru("gru2", g1, 512); g3 = gru("gru3", g2, 512)
        gb = gru("gru_gb", g3, 512)
        rb = gru("gru_rb", np.concatenate([g3, c2]), 128)
        G.append(sig(rx(np.concatenate([c2, g1, g2, g3, gb])) @ Wr["fc_gb_weights"] + W["fc_gb_bias"]))
        R.append(sig(rx(rb) @ Wr["fc_rb_weights"] + W["fc_rb_bias"]))
    return np.array(G), np.array(R)


def main():
    F = 16
    model = synth_model(0)
    xs = [v for v in synth_pcm(6, F, seed=321)] + list(edge_signals(F, 1.0).values())
    ffi.build()
    O = ffi.Oracle()
    worst = {"activations as one fp16 term (xl*wh dropped)": 0.0, "weights as one fp16 term (xh*wl dropped)": 0.0}
    for x in xs:
        hdl = O.create(model)
        _, _, taps = O.process_stream(hdl, np.asarray(x, np.float32), True, taps=True)
        O.destroy(hdl)
        feats = np.stack([t.np("features") for t in taps])
        g0, r0 = run(model.arrays, feats)
        for name, kw in (("activations as one fp16 term (xl*wh dropped)", dict(rx=f16)),
                         ("weights as one fp16 term (xh*wl dropped)", dict(rw=f16_scaled))):
            g, r = run(model.arrays, feats, **kw)
            rel = max((np.abs(g - g0) / np.maximum(np.abs(g0), 1e-6)).max(), (np.abs(r - r0) / np.maximum(np.abs(r0), 1e-6)).max())
            worst[name] = max(worst[name], rel)
    for k, v in worst.items():
        print(f"{k}: worst relative distance of g/r from the full evaluation over {len(xs)} signals x {F} hops = {v:.2e}")


if __name__ == "__main__":
    main()
