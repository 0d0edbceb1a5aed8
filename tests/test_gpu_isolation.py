"""Serving properties of the batched engine that the single-stream reference has by construction: a stream's
output depends on nothing but its own samples -- not on the batch size, its row index, the other rows' content
(even NaN/Inf), nor on how the hops are grouped into calls."""
import numpy as np
import pytest

from util import same_bits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    from percepnet_b200 import api
    api.load_library()
    return api


@pytest.mark.parametrize("nn", ["fp32", "tensor"])
def test_stream_is_independent_of_batch_and_neighbours(api, model0, nn):
    from percepnet_b200.synth import synth_pcm
    flags = api.NN_FP32 if nn == "fp32" else api.NN_TENSOR
    F = 14
    probe = synth_pcm(3, F, seed=777)                       # the streams we follow
    # (a) alone in a small batch
    e = api.Engine(3, F, model0, flags)
    out_a, gr_a = e.process(probe, want_gr=True)
    e.close()
    # (b) scattered through a batch of 261 (crosses two 128-row tiles of the tensor path) among other signals
    S = 261
    others = synth_pcm(S, F, seed=4000)
    rows = [5, 130, 260]
    x = others.copy()
    x[rows] = probe
    e = api.Engine(S, F, model0, flags)
    out_b, gr_b = e.process(x, want_gr=True)
    # (c) same batch, but the neighbours are poisoned with NaN / Inf / full-scale garbage from hop 3 on
    y = x.copy()
    bad = [r for r in range(S) if r not in rows]
    y[bad[0::3], 3 * 480 + 7] = np.nan
    y[bad[1::3], 4 * 480:] = np.inf
    y[bad[2::3], 3 * 480:] = 1e30
    e.reset()
    out_c, gr_c = e.process(y, want_gr=True)
    e.close()
    assert same_bits(out_b[rows], out_a) and same_bits(gr_b[:, rows], gr_a)
    assert same_bits(out_c[rows], out_a) and same_bits(gr_c[:, rows], gr_a)
    assert np.isfinite(out_c[rows]).all()


@pytest.mark.parametrize("nn", ["fp32", "tensor"])
def test_call_granularity_does_not_change_the_output(api, model0, nn):
    from percepnet_b200.synth import synth_pcm
    flags = api.NN_FP32 if nn == "fp32" else api.NN_TENSOR
    F = 23
    x = synth_pcm(9, F, seed=31)
    e = api.Engine(9, F, model0, flags)
    whole, gr_whole = e.process(x, want_gr=True)
    e.close()
    e = api.Engine(9, 9, model0, flags)                      # 9 = one group of 8 hops + 1: exercises the hop-group pre-pass
    outs, grs, t0 = [], [], 0
    for n in (1, 9, 2, 8, 3):
        o, g = e.process(x[:, t0 * 480:(t0 + n) * 480], want_gr=True)
        outs.append(o); grs.append(g); t0 += n
    e.close()
    assert t0 == F
    assert same_bits(np.concatenate(outs, axis=1), whole)
    assert same_bits(np.concatenate(grs, axis=0), gr_whole)
