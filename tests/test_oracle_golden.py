"""The oracle against the committed golden vectors (produced by the real reference, see
tests/golden/make_golden.py) and the reference's own toy-layer known answers
(/root/reference/tests/nnet_data_test.h, tests/testnnet.cpp:19-66).  CPU only, runs anywhere."""
import ctypes as C
import os

import numpy as np

from conftest import GOLDEN
from percepnet_b200.weights import Conv1DLayerC, DenseLayerC, GRULayerC
from util import same_bits


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def test_reference_toy_known_answers(oracle):
    """Same calls and tolerance (1e-5, two-sided) as tests/testnnet.cpp."""
    g = dict(np.load(os.path.join(GOLDEN, "toy_layers.npz")))   # materialise: the layer structs hold raw pointers
    eps = 1e-5
    fc = DenseLayerC(_p(g["fc_bias"]), _p(g["fc_weights"]), 2, 3, 1)
    out = oracle.dense_layer(fc, np.full(2, 0.5, np.float32), 3)
    assert np.max(np.abs(out - g["fc_output"])) < eps
    conv = Conv1DLayerC(_p(g["conv1_bias"]), _p(g["conv1_weights"]), 2, 3, 3, 1)
    mem = np.zeros(6, np.float32)
    x = np.full(2, 0.5, np.float32)
    oracle.conv1d_layer(conv, mem, x, 3)
    out = oracle.conv1d_layer(conv, mem, x, 3)
    assert np.max(np.abs(out - g["conv1_output"][:3])) < eps        # testnnet.cpp:38-41
    out = oracle.conv1d_layer(conv, mem, x, 3)
    assert np.max(np.abs(out - g["conv1_output"][3:6])) < eps       # testnnet.cpp:42-46
    gru = GRULayerC(_p(g["gru1_bias"]), _p(g["gru1_weights"]), _p(g["gru1_recurrent_weights"]), 2, 3, 2, 1)
    h = np.zeros(3, np.float32)
    oracle.gru_layer(gru, h, x)
    assert np.max(np.abs(h - g["gru1_output"][:3])) < eps            # testnnet.cpp:55-60 (two-sided here)
    oracle.gru_layer(gru, h, x)
    assert np.max(np.abs(h - g["gru1_output"][3:6])) < eps


def test_golden_end_to_end(oracle, model0):
    g = np.load(os.path.join(GOLDEN, "e2e.npz"))
    assert bytes(g["digest"]).decode() == model0.digest(), "synthetic weights differ from the fixture's"
    x16 = g["x16"]
    for name, scale in (("unit", np.float32(1 / 32768.0)), ("int16", np.float32(1.0))):
        h = oracle.create(model0)
        out, gr, _ = oracle.process_stream(h, x16.astype(np.float32) * scale, True)
        oracle.destroy(h)
        assert same_bits(out, g[f"out_{name}"]) and same_bits(gr, g[f"gr_{name}"])
    o16, gr = oracle.run_pcm16(model0, x16)
    assert np.array_equal(o16, g["cli_out16"]) and same_bits(gr, g["cli_gr"])


def test_golden_stages(oracle, model0):
    g = np.load(os.path.join(GOLDEN, "stages.npz"))
    assert np.array_equal(oracle.erb_borders(), g["borders"])
    prev = (0, 0.0)
    for i, buf in enumerate(g["pitch_buf"]):
        lp = oracle.pitch_downsample(buf)
        assert same_bits(lp, g["lp"][i])
        p, c, _, _ = oracle.pitch_search(lp)
        assert p == g["pitch"][i] and same_bits([c], [g["corr"][i]])
        T, gain = oracle.remove_doubling(lp, 768 - p, *prev)
        assert T == g["T"][i] and same_bits([gain], [g["gain"][i]])
        prev = (T, gain)
    assert same_bits(oracle.fft960(g["fft_in"]), g["fft_out"])
    X = g["fft_out"][:962]
    assert same_bits(oracle.band_energy(X), g["band_energy"])
    assert same_bits(oracle.band_corr(X, g["P"]), g["band_corr"])
    assert same_bits(oracle.interp_band_gain(g["gains"]), g["interp"])
    assert same_bits(oracle.pitch_filter(X, g["P"], g["gains"]), g["pitch_filter"])
    state = np.zeros(3712, np.float32)
    for k in range(3):
        gg, rr = oracle.compute_rnn(model0, state, g["rnn_feat"])
        assert same_bits(np.concatenate([gg, rr]), g["rnn_out"][k])


def test_post_filter_properties(oracle):
    """denoise.cpp:216-250 has no reference-side test; check its defining relations."""
    rng = np.random.RandomState(3)
    g = rng.rand(34).astype(np.float32)
    Ey = rng.rand(34).astype(np.float32) + 0.1
    out = oracle.post_filter(g, Ey)
    gw = g * np.sin(np.pi / 2 * g)
    q = np.dot(g, Ey) / (np.dot(gw, Ey) + 1e-6)
    G = np.sqrt(1.02 * q / (1 + 0.02 * q * q))
    assert np.allclose(out, G * gw, rtol=2e-6, atol=1e-7)
    assert np.allclose(oracle.post_filter(np.ones(34, np.float32), Ey), 1.0, atol=1e-5)   # g = 1 is a fixed point


def test_golden_training_records(oracle):
    """Row f1: the restated train() loop against the records the reference's train() wrote (tests/golden/train.npz)."""
    g = np.load(os.path.join(GOLDEN, "train.npz"))
    for k in range(g["speech"].shape[0]):
        got = oracle.train_records(g["speech"][k], g["noisy"][k])
        assert same_bits(got, g["records"][k])
    rec = g["records"]
    assert np.any(rec[..., 104:] == np.float32(0.99)) and np.any(rec[..., 104:] < 0.5)    # both label branches occur
