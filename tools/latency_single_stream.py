#!/usr/bin/env python
"""Per-call latency of small batches (what the rnnoise_* shim pays per 10 ms frame): S streams, one hop per call, host
float buffers in and out through the blocking public call."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from percepnet_b200 import api  # noqa: E402
from percepnet_b200.synth import synth_pcm  # noqa: E402
from percepnet_b200.weights import synth_model  # noqa: E402


def main():
    m = synth_model(0)
    res = {}
    for S in (1, 16, 256):
        x = synth_pcm(min(S, 16), 400, seed=5)
        if S > 16:
            import numpy as np
            x = np.tile(x, (S // 16, 1))
        for name, flags in (("fp32", api.NN_FP32), ("tensor", api.NN_TENSOR)):
            e = api.Engine(S, 1, m, flags)
            for t in range(100):
                e.process(x[:, t * 480:(t + 1) * 480])
            t0 = time.perf_counter()
            for t in range(100, 400):
                e.process(x[:, t * 480:(t + 1) * 480])
            dt = (time.perf_counter() - t0) / 300
            # where the time goes: CUDA events around every launch of a few more calls
            e.profile(True)
            for t in range(20):
                e.process(x[:, t * 480:(t + 1) * 480])
            prof = {k: round(1e3 * v[0] / 20, 1) for k, v in e.profile_read().items()}
            e.profile(False)
            res[f"S{S}_{name}"] = {"us_per_call": round(dt * 1e6, 1), "x_realtime_per_stream": round(0.01 / dt, 1),
                                   "launches_per_call": e.launches_per_call(1), "kernel_us_per_call": prof}
            e.close()
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
