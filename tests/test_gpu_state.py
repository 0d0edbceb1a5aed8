"""pnb_get_state / pnb_set_state: a stream moved between engines at a call boundary continues exactly where it left
off (the live part of the reference's DenoiseState + RNNState, /root/reference/src/denoise.cpp:71-85)."""
import numpy as np
import pytest

from test_gpu_parity import GR_RTOL, _inputs, api  # noqa: F401

pytestmark = pytest.mark.gpu


def _migrate(api, model, flags_a, flags_b, scale):
    x = _inputs(scale, 13, n_synth=5)[:5]
    other = _inputs(scale, 6, n_synth=3)[:3] * np.float32(0.5)
    a = api.Engine(5, 7, model, flags_a)
    a.process(x[:, :7 * 480])
    blob = a.get_state(2)
    assert len(blob) == a.L.pnb_state_size()
    want, want_gr = a.process(np.ascontiguousarray(x[:, 7 * 480:]), want_gr=True)     # A simply carries on
    a.close()
    b = api.Engine(3, 6, model, flags_b)                                               # other batch size, ring length, hop count
    b.process(other)
    b.set_state(1, blob)
    y = np.zeros((3, 6 * 480), np.float32)
    y[1] = x[2, 7 * 480:]
    got, got_gr = b.process(y, want_gr=True)
    b.close()
    return want[2], want_gr[:, 2], got[1], got_gr[:, 1]


@pytest.mark.parametrize("mode", ["fp32", "tensor"])
@pytest.mark.parametrize("scale", [1.0, 32768.0], ids=["unit", "int16scale"])
def test_stream_migrates_bit_exactly(api, model0, mode, scale):
    if mode == "tensor" and scale != 1.0:
        scale = 256.0
    f = api.NN_FP32 if mode == "fp32" else api.NN_TENSOR
    w, wg, g, gg = _migrate(api, model0, f, f, scale)
    assert np.array_equal(w, g) and np.array_equal(wg, gg)
    assert np.abs(w).max() > 0


def test_stream_migrates_between_network_modes(api, model0):
    w, wg, g, gg = _migrate(api, model0, api.NN_FP32, api.NN_TENSOR, 1.0)
    rel = np.abs(gg - wg) / np.maximum(np.abs(wg), 1e-6)
    assert rel.max() < GR_RTOL
    assert np.abs(np.trunc(w.astype(np.float64) * 32768) - np.trunc(g.astype(np.float64) * 32768)).max() <= 1


def test_state_argument_checks(api, model0):
    e = api.Engine(2, 2, model0)
    blob = e.get_state(0)
    with pytest.raises(api.PnbError):
        e.get_state(2)
    with pytest.raises(api.PnbError):
        e.set_state(0, blob[:100])
    with pytest.raises(api.PnbError):
        e.set_state(0, b"\0" * len(blob))
    e.set_state(1, blob)
    e.close()
