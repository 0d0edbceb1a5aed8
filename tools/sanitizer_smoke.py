#!/usr/bin/env python
"""Small run of every kernel for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):

    compute-sanitizer --tool racecheck python tools/sanitizer_smoke.py

Exercises the fp32 path, the tensor path (one CTA pair's worth of streams), the post-filter, the int16 entry,
the pipelined host entry, the training-data generator, the persistent fp32 GRU chain (both block sizes), the chunked
overlap schedule with submitted calls, state export / import and the pitch-only kernel on tiny batches; prints a
checksum per leg."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from percepnet_b200 import api  # noqa: E402
from percepnet_b200.synth import synth_pairs, synth_pcm, to_int16  # noqa: E402
from percepnet_b200.weights import synth_model  # noqa: E402


def main():
    legs = sys.argv[1:] or ["fp32", "tensor", "train", "fp32chain", "overlap", "pitch"]
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
    if "fp32chain" in legs:  # the persistent fp32 GRU chain: 128-stream blocks (partial last block) and 16-stream blocks
        for S, F in ((130, 5), (5, 3)):
            x = synth_pcm(S, 2 * F, seed=6)
            e = api.Engine(S, F, m, api.NN_FP32)
            o1 = e.process(x[:, :F * 480])[0]
            o2 = e.process(x[:, F * 480:])[0]
            blob = e.get_state(S - 1)
            e.set_state(0, blob)
            print("fp32chain", S, float(np.abs(o1).sum()), float(np.abs(o2).sum()), len(blob))
            e.close()
    if "overlap" in legs:    # chunked schedule on the SM partition, submitted calls overlapping each other, anti-diagonal unit order
        os.environ.update({"PNB_OVERLAP": "2", "PNB_CHUNK": "2", "PNB_NET_SMS": "64"})
        S, F = 300, 6
        x = synth_pcm(S, 3 * F, seed=8)
        e = api.Engine(S, F, m, api.NN_TENSOR | api.CONV_WIDE)
        outs = [np.empty((S, F * 480), np.float32) for _ in range(3)]
        for k in range(3):
            e.submit(np.ascontiguousarray(x[:, k * F * 480:(k + 1) * F * 480]), outs[k])
        e.wait()
        print("overlap", e.overlap_info(), [float(np.abs(o).sum()) for o in outs])
        e.close()
        for k in ("PNB_OVERLAP", "PNB_CHUNK", "PNB_NET_SMS"):
            os.environ.pop(k)
    if "pitch" in legs:
        buf = synth_pcm(40, 4, seed=12)[:, :1728].copy()
        r = api.pitch_only(buf, np.zeros(40, np.int32), np.zeros(40, np.float32))
        print("pitch", [float(np.asarray(a, np.float64).sum()) for a in r])
    if "train" in legs:
        c, n = synth_pairs(6, 5, seed=9)
        e = api.Engine(12, 3, None, api.TRAIN_DATA)
        r1 = e.train_records(c[:, :3 * 480], n[:, :3 * 480])
        r2 = e.train_records(c[:, 3 * 480:], n[:, 3 * 480:])
        print("train", float(r1.sum()), float(r2.sum()))
        e.close()


if __name__ == "__main__":
    main()
