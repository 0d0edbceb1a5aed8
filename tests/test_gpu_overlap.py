"""The chunked two-stream schedule (network of chunk k on its own SMs while analysis k+1 / synthesis k-1 run on the
others, green contexts) must not change a single bit relative to the serial schedule."""
import os

import numpy as np
import pytest

from test_gpu_parity import _inputs, api  # noqa: F401

pytestmark = pytest.mark.gpu


def _run(api, model, x, F, env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        eng = api.Engine(x.shape[0], F, model, api.NN_TENSOR)
        info = eng.overlap_info()
        out, gr = eng.process_stream_chunks(x, want_gr=True)
        eng.check()
        eng.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return out, gr, info


@pytest.mark.parametrize("chunk", ["4", "8"])
def test_overlapped_schedule_is_bit_identical(api, model0, chunk):
    base = _inputs(1.0, 44, n_synth=6)
    x = np.tile(base, (20, 1))[:270]                       # 270 streams: not a multiple of the 256-row tile pair
    F = 22                                                 # calls of 22 + 22 hops: 6 (3) chunks, the last one short
    serial = _run(api, model0, x, F, {"PNB_OVERLAP": "0"})
    assert serial[2]["net_sms"] == 0
    over = _run(api, model0, x, F, {"PNB_OVERLAP": "2", "PNB_CHUNK": chunk, "PNB_NET_SMS": "64"})
    if over[2]["net_sms"] == 0:
        pytest.skip("green contexts are not available on this driver")
    assert over[2]["chunk_hops"] == int(chunk) and over[2]["net_sms"] >= 64 and over[2]["dsp_sms"] >= 8
    assert np.array_equal(serial[0], over[0]) and np.array_equal(serial[1], over[1])
