"""Row f1: the training-data generator on the GPU (pnb_train_records_*, flag PNB_TRAIN_DATA)
against the oracle's restatement of the reference's train() loop and against the records the reference itself
wrote (tests/golden/train.npz).

Bars: the 70 input features of a record (Ey_lookahead, Ephaty, T, pitchcorr) and the r labels are BIT-EXACT;
the g labels pass through the post-filter's sinf (denoise.cpp:227), where libm and the device differ by at most
an ulp in rare cases: 1e-6 relative, and at least 95 % of them bit-identical (measured: 99.4 %).
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from util import same_bits

pytestmark = pytest.mark.gpu

G_RTOL = 1e-6


@pytest.fixture(scope="module")
def api():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    from percepnet_b200 import api
    api.load_library()
    return api


def _check(got, want):
    assert got.shape == want.shape
    assert same_bits(got[..., :70], want[..., :70]), "record inputs (Ey_lookahead, Ephaty, T, pitchcorr) differ"
    assert same_bits(got[..., 104:], want[..., 104:]), "r labels differ"
    g, w = got[..., 70:104], want[..., 70:104]
    assert np.all(np.abs(g - w) <= G_RTOL * np.abs(w) + 1e-12), np.abs(g - w).max()
    # one differently-rounded sine moves the post-filter's normaliser G, i.e. all 34 gains of that frame, by an ulp
    assert np.mean(g.view(np.int32) == w.view(np.int32)) > 0.95


def test_records_match_reference_golden(api):
    g = np.load(os.path.join(GOLDEN, "train.npz"))
    speech, noisy, want = g["speech"], g["noisy"], g["records"]
    N, F = speech.shape[0], speech.shape[1] // 480
    eng = api.Engine(2 * N, F, None, api.TRAIN_DATA)
    got = eng.train_records(speech, noisy)
    eng.close()
    _check(got, want)


def test_records_match_oracle_chunked(api, oracle):
    """Synthetic pairs at several SNRs plus degenerate pairs, fed in uneven chunks (state carries across calls)."""
    from percepnet_b200.synth import synth_pairs
    F = 36
    clean, noisy = synth_pairs(5, F, seed=4242)
    clean = np.concatenate([clean, np.zeros_like(clean[:1]), noisy[1:2], clean[2:3]])
    noisy = np.concatenate([noisy, noisy[:1], noisy[1:2], (noisy[2:3] // 64).astype(np.int16)])
    N = clean.shape[0]
    want = np.stack([oracle.train_records(clean[k], noisy[k]) for k in range(N)])
    eng = api.Engine(2 * N, 16, None, api.TRAIN_DATA)
    parts, t0 = [], 0
    for n in (1, 16, 7, 12):
        parts.append(eng.train_records(clean[:, t0 * 480:(t0 + n) * 480], noisy[:, t0 * 480:(t0 + n) * 480]))
        t0 += n
    assert t0 == F
    got = np.concatenate(parts, axis=1)
    _check(got, want)
    # the pipelined entry gives the same records (three calls in flight over two slots)
    eng.reset()
    outs = [np.empty((N, n, 138), np.float32) for n in (12, 12, 12)]
    ins = [(np.ascontiguousarray(clean[:, k * 12 * 480:(k + 1) * 12 * 480]), np.ascontiguousarray(noisy[:, k * 12 * 480:(k + 1) * 12 * 480])) for k in range(3)]
    for (c, y), o in zip(ins, outs):
        rc = eng.L.pnb_submit_train_records(eng.h, c.ctypes.data, c.shape[1], y.ctypes.data, y.shape[1], 12, o.ctypes.data, 12 * 138)
        assert rc == 0, eng.L.pnb_last_error()
    assert eng.L.pnb_wait(eng.h) == 0
    assert same_bits(np.concatenate(outs, axis=1), got)
    # reset gives the same records again
    eng.reset()
    again = eng.train_records(clean[:, :16 * 480], noisy[:, :16 * 480])
    assert same_bits(again, got[:, :16])
    eng.close()


def test_train_mode_argument_checks(api, model0):
    with pytest.raises(api.PnbError):
        api.Engine(3, 4, None, api.TRAIN_DATA)                      # odd stream count
    with pytest.raises(api.PnbError):
        api.Engine(4, 4, None, api.TRAIN_DATA | api.NN_TENSOR)      # no network in this mode
    with pytest.raises(api.PnbError):
        api.Engine(4, 4, None, api.NN_FP32)                         # enhancement needs a model
    eng = api.Engine(4, 4, None, api.TRAIN_DATA)
    with pytest.raises(api.PnbError):
        eng.process(np.zeros((4, 480), np.float32))                 # enhancement entry refused
    z16 = np.zeros((4, 480), np.int16)
    assert eng.L.pnb_submit_host_i16(eng.h, z16.ctypes.data, 480, z16.ctypes.data, 480, 1) != 0   # and its pipelined form
    with pytest.raises(api.PnbError):
        eng.train_records(np.zeros((2, 5 * 480), np.int16), np.zeros((2, 5 * 480), np.int16))   # > max_frames
    eng.close()
    enh = api.Engine(2, 4, model0, api.NN_FP32)
    with pytest.raises(api.PnbError):
        enh.train_records(np.zeros((1, 480), np.int16), np.zeros((1, 480), np.int16))
    enh.close()


def test_cli_writes_reference_files(api, tmp_path):
    """python -m percepnet_b200.gen_features with the reference's argv and with a job list (one batch)."""
    from percepnet_b200 import gen_features
    g = np.load(os.path.join(GOLDEN, "train.npz"))
    jobs = []
    for k in range(2):
        fc, fn, fo = (str(tmp_path / f"{nm}{k}") for nm in ("c", "n", "o"))
        g["speech"][k].tofile(fc); g["noisy"][k].tofile(fn)
        jobs.append((fc, fn, 40 - 7 * k, fo))
    assert gen_features.main([jobs[0][0], jobs[0][1], "40", jobs[0][3]]) == 0
    _check(np.fromfile(jobs[0][3], np.float32).reshape(1, 40, 138), g["records"][:1])
    lst = tmp_path / "jobs.txt"
    lst.write_text("".join(f"{a} {b} {c} {d}\n" for a, b, c, d in jobs))
    assert gen_features.main(["--list", str(lst), "--chunk", "16"]) == 0
    for k in range(2):
        n = jobs[k][2]
        _check(np.fromfile(jobs[k][3], np.float32).reshape(1, n, 138), g["records"][k:k + 1, :n])
    assert gen_features.main(["only", "three", "args"]) == 1
