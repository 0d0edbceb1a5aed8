#!/usr/bin/env python
"""Small run of every kernel for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):

    compute-sanitizer --tool racecheck python tools/sanitizer_smoke.py

Exercises the fp32 path, the tensor path (one CTA pair's worth of streams), the post-filter, the int16 entry,
the pipelined host entry and the training-data generator on tiny batches; prints a checksum per leg."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from percepnet_b200 import api  # noqa: E402
from percepnet_b200.synth import synth_pairs, synth_pcm, to_int16  # noqa: E402
from percepnet_b200.weights import synth_model  # noqa: E402


def main():
    legs = sys.argv[1:] or ["fp32", "tensor", "train"]
    m = synth_model(0)
    if "fp32" in legs:
        x = synth_pcm(5, 4, seed=3)
        e = api.Engine(5, 2, m, api.NN_FP32 | api.POSTFILTER | api.KEEP_TAPS)
        o1, gr = e.process(x[:, :960], want_gr=True)
        o2 = e.process(to_int16(x[:, 960:]))[0]
        print("fp32", float(np.abs(o1).sum()), int(np.abs(o2.astype(np.int64)).sum()), float(gr.sum()))
        e.close()
    if "tensor" in legs:
        S = 300
        x = synth_pcm(S, 3, seed=4)
        e = api.Engine(S, 2, m, api.NN_TENSOR)
        o1, gr = e.process(x[:, :960], want_gr=True)
        o2, _ = e.process(x[:, 960:], want_gr=True)
        print("tensor", float(np.abs(o1).sum()), float(np.abs(o2).sum()), float(gr.sum()))
        e.close()
    if "train" in legs:
        c, n = synth_pairs(6, 5, seed=9)
        e = api.Engine(12, 3, None, api.TRAIN_DATA)
        r1 = e.train_records(c[:, :3 * 480], n[:, :3 * 480])
        r2 = e.train_records(c[:, 3 * 480:], n[:, 3 * 480:])
        print("train", float(r1.sum()), float(r2.sum()))
        e.close()


if __name__ == "__main__":
    main()
