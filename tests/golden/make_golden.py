#!/usr/bin/env python
"""Regenerates the committed golden vectors from the REAL reference.

Run in the build container (needs /root/reference and oracle/_ref/libpercepnet_ref.so, i.e. the
unmodified reference sources compiled by oracle/Makefile):

    python tests/golden/make_golden.py

Writes (all small, committed):
  toy_layers.npz   the reference's own known-answer vectors, parsed from
                   /root/reference/tests/nnet_data_test.h (tests/testnnet.cpp:19-66)
  e2e.npz          int16 input (speech+noise mix of /root/reference/sampledata, 48 hops), and what the
                   compiled reference returned for it through the float C API at both amplitude
                   scales (out, g/r) and through the CLI's int16 conversions (src/main.cpp:30-39),
                   with weights = percepnet_b200.weights.synth_model(0) (digest stored)
  stages.npz       per-stage outputs of the reference's non-static functions on the same audio
  train.npz        two (speech, noisy) int16 pairs cut from /root/reference/sampledata and the 138-float records
                   the reference's own train() (src/denoise.cpp:600) wrote for them
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ffi import Reference, build  # noqa: E402
from percepnet_b200.weights import synth_model  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def parse_arrays(path):
    txt = open(path).read()
    arrs = {}
    for m in re.finditer(r"static const float (\w+)\[(\d+)\] = \{([^}]*)\}", txt):
        arrs[m.group(1)] = np.array([float(v) for v in m.group(3).replace("\n", " ").split(",") if v.strip()],
                                    dtype=np.float32)
        assert arrs[m.group(1)].size == int(m.group(2))
    return arrs


def make_train(R):
    import tempfile
    n = 40
    sp = np.fromfile(os.path.join(REF, "sampledata/speech/speech.pcm"), dtype=np.int16)
    no = np.fromfile(os.path.join(REF, "sampledata/noise/noise.pcm"), dtype=np.int16)
    speech, noisy, recs = [], [], []
    with tempfile.TemporaryDirectory() as d:
        for k, (off, shift) in enumerate(((480 * 100, 1), (480 * 260, 3))):
            c = sp[off:off + 480 * n].copy()
            y = np.clip(c.astype(np.int32) + (no[off:off + 480 * n].astype(np.int32) >> shift), -32768, 32767).astype(np.int16)
            fc, fn, fo = (os.path.join(d, f"{nm}{k}") for nm in ("c", "n", "o"))
            c.tofile(fc); y.tofile(fn)
            assert R.train_files(fc, fn, n, fo) == 0
            speech.append(c); noisy.append(y)
            recs.append(np.fromfile(fo, np.float32).reshape(n, 138))
    np.savez_compressed(os.path.join(OUT, "train.npz"), speech=np.stack(speech), noisy=np.stack(noisy),
                        records=np.stack(recs))
    print("train.npz", os.path.getsize(os.path.join(OUT, "train.npz")))


def main():
    build()
    if "--train-only" in sys.argv:
        return make_train(Reference())
    R = Reference()
    np.savez_compressed(os.path.join(OUT, "toy_layers.npz"), **parse_arrays(os.path.join(REF, "tests/nnet_data_test.h")))

    m = synth_model(0)
    R.set_model(m)
    n = 48
    sp = np.fromfile(os.path.join(REF, "sampledata/speech/speech.pcm"), dtype=np.int16)
    no = np.fromfile(os.path.join(REF, "sampledata/noise/noise.pcm"), dtype=np.int16)
    off = 480 * 100
    x16 = (sp[off:off + 480 * n].astype(np.int32) // 2 + no[off:off + 480 * n].astype(np.int32) // 2).astype(np.int16)
    res = {"x16": x16, "digest": np.frombuffer(m.digest().encode(), dtype=np.uint8)}
    for name, scale in (("unit", np.float32(1.0 / 32768.0)), ("int16", np.float32(1.0))):
        x = x16.astype(np.float32) * scale
        h = R.create()
        out, gr = R.process_stream(h, x, True)
        R.destroy(h)
        res[f"out_{name}"] = out
        res[f"gr_{name}"] = gr
    o16, gr16 = R.run_pcm16(x16)
    res["cli_out16"] = o16
    res["cli_gr"] = gr16
    np.savez_compressed(os.path.join(OUT, "e2e.npz"), **res)

    # stage taps from reference functions
    st = {}
    x = x16.astype(np.float32)
    bufs = np.stack([x[480 * k:480 * k + 1728] for k in (3, 9, 17, 30)])
    st["pitch_buf"] = bufs
    lps, ps, cs, ts, gs = [], [], [], [], []
    prev_p, prev_g = 0, 0.0
    for b in bufs:
        lp = R.pitch_downsample(b)
        p, c = R.pitch_search(lp)
        T, g = R.remove_doubling(lp, 768 - p, prev_p, prev_g)
        prev_p, prev_g = T, g
        lps.append(lp); ps.append(p); cs.append(c); ts.append(T); gs.append(g)
    st["lp"] = np.stack(lps)
    st["pitch"] = np.array(ps, np.int32)
    st["corr"] = np.array(cs, np.float32)
    st["T"] = np.array(ts, np.int32)
    st["gain"] = np.array(gs, np.float32)
    rng = np.random.RandomState(5)
    z = rng.randn(1920).astype(np.float32)
    st["fft_in"] = z
    st["fft_out"] = R.fft960(z)
    X = st["fft_out"][:962]
    P = R.fft960(rng.randn(1920).astype(np.float32))[:962]
    st["P"] = P
    st["band_energy"] = R.band_energy(X)
    st["band_corr"] = R.band_corr(X, P)
    gb = rng.rand(34).astype(np.float32)
    st["gains"] = gb
    st["interp"] = R.interp_band_gain(gb)
    st["pitch_filter"] = R.pitch_filter(X, P, gb)
    feat = (rng.rand(70) * 3).astype(np.float32)
    state = np.zeros(512 + 1024 + 4 * 512 + 128, np.float32)
    outs = []
    for _ in range(3):
        g, r = R.compute_rnn(state, feat)
        outs.append(np.concatenate([g, r]))
    st["rnn_feat"] = feat
    st["rnn_out"] = np.stack(outs)
    st["borders"] = R.erb_borders()
    np.savez_compressed(os.path.join(OUT, "stages.npz"), **st)
    make_train(R)
    for f in ("toy_layers.npz", "e2e.npz", "stages.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
