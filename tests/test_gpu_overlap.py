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


def _with_env(env, fn):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("chunk", ["2", "4", "8"])
def test_submitted_calls_overlap_each_other_bit_identically(api, model0, chunk):
    """pnb_submit_host_* / pnb_submit_device_*: the analysis of call i+1 runs while the network and synthesis of call i are
    still in flight.  Calls of different lengths (chunked and serial ones mixed) must give the serial engine's bits."""
    import torch
    from percepnet_b200.api import FRAME
    lens = [22, 22, 22, 5, 22, 17, 22, 22]                 # hops per call; 5 is below two chunks of 4/8: serial schedule
    total = sum(lens)
    base = _inputs(1.0, total, n_synth=6)
    x = np.ascontiguousarray(np.tile(base, (20, 1))[:270])
    S, Fm = x.shape[0], max(lens)

    def serial():
        eng = api.Engine(S, Fm, model0, api.NN_TENSOR)
        outs, t = [], 0
        for n in lens:
            outs.append(eng.process(x[:, t * FRAME:(t + n) * FRAME])[0])
            t += n
        eng.close()
        return np.concatenate(outs, axis=1)
    ref = _with_env({"PNB_OVERLAP": "0"}, serial)

    def submitted_host():
        eng = api.Engine(S, Fm, model0, api.NN_TENSOR)
        if eng.overlap_info()["net_sms"] == 0:
            eng.close()
            return None
        ins, outs, t = [], [], 0
        for n in lens:
            ins.append(np.ascontiguousarray(x[:, t * FRAME:(t + n) * FRAME]))
            outs.append(np.empty_like(ins[-1]))
            eng.submit(ins[-1], outs[-1])
            t += n
        eng.wait()
        eng.close()
        return np.concatenate(outs, axis=1)
    env = {"PNB_OVERLAP": "2", "PNB_CHUNK": chunk, "PNB_NET_SMS": "64"}
    got = _with_env(env, submitted_host)
    if got is None:
        pytest.skip("green contexts are not available on this driver")
    assert np.array_equal(ref, got)

    def submitted_device():
        eng = api.Engine(S, Fm, model0, api.NN_TENSOR)
        xd = torch.from_numpy(x).cuda()
        yd = torch.zeros_like(xd)
        st = torch.cuda.current_stream().cuda_stream
        t = 0
        for i, n in enumerate(lens):
            a, b = xd[:, t * FRAME:], yd[:, t * FRAME:]
            if i == 4:                                     # a plain call in the middle of submitted ones
                eng.process_device(a.data_ptr(), xd.stride(0), b.data_ptr(), yd.stride(0), n, stream=st)
            else:
                eng.submit_device(a.data_ptr(), xd.stride(0), b.data_ptr(), yd.stride(0), n, stream=st)
            t += n
        eng.flush(st)
        torch.cuda.synchronize()
        eng.check()
        out = yd.cpu().numpy()
        eng.close()
        return out
    got = _with_env(env, submitted_device)
    assert np.array_equal(ref, got)
