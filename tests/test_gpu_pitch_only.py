"""BASELINE.json config 5: the pitch analysis alone (pnb_pitch_only_device = pitch_downsample + pitch_search +
remove_doubling, /root/reference/src/pitch.cpp:148-216, 283-386, 423-527) against the oracle's stage functions,
which are pinned bit for bit to the reference's own (tests/test_oracle_vs_reference.py).  Integer results must be
equal, float results bit-identical."""
import numpy as np
import pytest

from test_gpu_parity import api  # noqa: F401
from util import edge_signals, random_mixtures, same_bits

pytestmark = pytest.mark.gpu


def _oracle_pitch(oracle, buf, prev_period, prev_gain):
    lp = oracle.pitch_downsample(buf)
    lag, corr, _, _ = oracle.pitch_search(lp)
    T, gain = oracle.remove_doubling(lp, 768 - lag, int(prev_period), float(prev_gain))
    return lag, corr, T, gain


def _windows(seed, n, scale):
    """n pitch buffers of 1728 samples cut from edge signals and random mixtures at amplitude `scale`"""
    sig = list(edge_signals(12, scale).values()) + list(random_mixtures(24, 12, seed=seed) * np.float32(scale))
    rng = np.random.RandomState(seed)
    out = np.empty((n, 1728), np.float32)
    for k in range(n):
        s = sig[k % len(sig)]
        o = rng.randint(0, s.size - 1728)
        out[k] = s[o:o + 1728]
    return out


@pytest.mark.parametrize("scale", [1.0, 32768.0], ids=["unit", "int16scale"])
def test_pitch_only_matches_oracle(api, oracle, scale):
    import torch
    n = 300
    buf = _windows(5 + int(scale), n, scale)
    rng = np.random.RandomState(3)
    prev_T = rng.randint(60, 768, n).astype(np.int32)
    prev_T[::3] = 0                                         # a fresh stream
    prev_g = rng.rand(n).astype(np.float32)
    prev_g[::3] = 0
    stride = 1728 + 32                                      # rows need not be packed
    d_buf = torch.zeros((n, stride), dtype=torch.float32, device="cuda")
    d_buf[:, :1728] = torch.from_numpy(buf).cuda()
    d_pT, d_pg = torch.from_numpy(prev_T).cuda(), torch.from_numpy(prev_g).cuda()
    d_T = torch.empty(n, dtype=torch.int32, device="cuda")
    d_lag = torch.empty(n, dtype=torch.int32, device="cuda")
    d_corr = torch.empty(n, dtype=torch.float32, device="cuda")
    d_gain = torch.empty(n, dtype=torch.float32, device="cuda")
    api.pitch_only_device(d_buf.data_ptr(), stride, n, d_T.data_ptr(), d_corr.data_ptr(), d_gain.data_ptr(), d_lag.data_ptr(),
                          d_pT.data_ptr(), d_pg.data_ptr())
    torch.cuda.synchronize()
    T, lag, corr, gain = d_T.cpu().numpy(), d_lag.cpu().numpy(), d_corr.cpu().numpy(), d_gain.cpu().numpy()
    want = [_oracle_pitch(oracle, buf[k], prev_T[k], prev_g[k]) for k in range(n)]
    w_lag, w_corr, w_T, w_gain = (np.array([w[i] for w in want]) for i in range(4))
    assert np.array_equal(lag, w_lag) and np.array_equal(T, w_T)
    assert same_bits(corr, w_corr.astype(np.float32)) and same_bits(gain, w_gain.astype(np.float32))
    assert len(set(T.tolist())) > 20                        # the windows do exercise many periods


def test_pitch_only_null_prev_and_tail_block(api, oracle):
    """prev pointers NULL (= 0) and a unit count that is not a multiple of the block's 8 warps"""
    import torch
    n = 13
    buf = _windows(77, n, 1.0)
    d_buf = torch.from_numpy(buf).cuda()
    d_T = torch.empty(n, dtype=torch.int32, device="cuda")
    d_corr = torch.empty(n, dtype=torch.float32, device="cuda")
    d_gain = torch.empty(n, dtype=torch.float32, device="cuda")
    api.pitch_only_device(d_buf.data_ptr(), 1728, n, d_T.data_ptr(), d_corr.data_ptr(), d_gain.data_ptr())
    torch.cuda.synchronize()
    want = [_oracle_pitch(oracle, buf[k], 0, 0.0) for k in range(n)]
    assert np.array_equal(d_T.cpu().numpy(), np.array([w[2] for w in want]))
    assert same_bits(d_gain.cpu().numpy(), np.array([w[3] for w in want], np.float32))
    with pytest.raises(api.PnbError):
        api.pitch_only_device(d_buf.data_ptr(), 100, n, d_T.data_ptr(), d_corr.data_ptr(), d_gain.data_ptr())
