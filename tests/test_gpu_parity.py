"""Parity of the CUDA hot path (through the C-ABI, include/percepnet_b200.h) against the oracle.

Bars (BASELINE.json north_star / SURVEY.md 8c):
  * DSP intermediates -- spectra X and P, band energies, pitch lag / period, the 70 features -- are
    BIT-EXACT (the kernels keep the reference's summation order and use no FMA);
  * network outputs g, r: 1e-4 relative;
  * PCM: +-1 LSB of int16 after the CLI's truncating conversion.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from util import edge_signals, same_bits

pytestmark = pytest.mark.gpu

GR_RTOL = 1e-4        # north_star: "1e-4 relative fp32 on band gains"
PCM_LSB = 1           # north_star: "+-1 LSB int16"


@pytest.fixture(scope="module")
def api():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    from percepnet_b200 import api
    api.load_library()
    return api


def _inputs(scale, n_frames, n_synth=4):
    from percepnet_b200.synth import synth_pcm
    xs = [v for v in synth_pcm(n_synth, n_frames, seed=321, scale=scale)]
    # int16-scale floats go through the C API the way the reference's train() feeds it; keep the edge
    # signals at a quarter of full scale there: beyond |pre-activation| ~ 8.6e7 the reference's
    # tansig_approx (vec.h:63, float -> int conversion) is undefined behaviour and returns garbage/NaN
    xs += list(edge_signals(n_frames, scale if scale == 1.0 else scale * 0.25).values())
    return np.stack(xs).astype(np.float32)


def _oracle_run(oracle, model, x, flags=0):
    S, T = x.shape
    outs, grs, taps = [], [], []
    for s in range(S):
        h = oracle.create(model)
        o, g, tp = oracle.process_stream(h, x[s], True, flags=flags, taps=True)
        oracle.destroy(h)
        outs.append(o); grs.append(g); taps.append(tp)
    return np.stack(outs), np.stack(grs, axis=1), taps      # out [S,T], gr [F,S,68]


def _lsb_diff(a, b, unit_scale):
    """difference in int16 LSBs after the CLI's conversion (src/main.cpp:36: truncation)."""
    k = 32768.0 if unit_scale else 1.0
    ia = np.trunc(a.astype(np.float64) * k)
    ib = np.trunc(b.astype(np.float64) * k)
    return np.abs(ia - ib).max()


@pytest.mark.parametrize("scale", [1.0, 32768.0], ids=["unit", "int16scale"])
def test_dsp_taps_bit_exact_and_outputs(api, oracle, model0, scale):
    F, Fmax = 20, 8
    x = _inputs(scale, F)
    S = x.shape[0]
    ref_out, ref_gr, taps = _oracle_run(oracle, model0, x)
    eng = api.Engine(S, Fmax, model0, api.NN_FP32 | api.KEEP_TAPS)
    outs, grs = [], []
    t0 = 0
    while t0 < F:                      # ragged chunking: 8, 8, 4 hops -> state carried across calls
        n = min(Fmax, F - t0)
        o, g = eng.process(x[:, t0 * 480:(t0 + n) * 480], want_gr=True)
        feats = eng.read_tap("features", n)
        pitch = eng.read_tap("pitch", n)
        pitchf = eng.read_tap("pitchf", n)
        X = eng.read_tap("X", n)
        P = eng.read_tap("P", n)
        Ex = eng.read_tap("Ex", n)
        for s in range(S):
            for k in range(n):
                tp = taps[s][t0 + k]
                tag = f"stream {s} hop {t0 + k}"
                assert pitch[k, s, 0] == tp.pitch_search, tag
                assert pitch[k, s, 1] == tp.pitch_index, tag
                assert pitch[k, s, 2] == tp.silence, tag
                assert same_bits(pitchf[k, s], [tp.pitch_corr, tp.pitch_gain]), tag
                assert same_bits(X[k, s], tp.np("X")[:800]), tag
                assert same_bits(P[k, s], tp.np("P")[:800]), tag
                assert same_bits(Ex[k, s], tp.np("Ex")), tag
                assert same_bits(feats[k, s], tp.np("features")), tag
        outs.append(o); grs.append(g)
        t0 += n
    out = np.concatenate(outs, axis=1)
    gr = np.concatenate(grs, axis=0)
    rel = np.abs(gr - ref_gr) / np.maximum(np.abs(ref_gr), 1e-6)
    assert rel.max() < GR_RTOL, f"g/r relative error {rel.max():.3g}"
    assert _lsb_diff(out, ref_out, scale == 1.0) <= PCM_LSB
    # and, as floats, far inside one LSB
    lsb = 1.0 / 32768.0 if scale == 1.0 else 1.0
    assert np.abs(out - ref_out).max() < 0.25 * lsb
    if scale != 1.0:
        assert sum(1 for s in range(S) for t in taps[s] if not t.silence) > 20   # comb-filter branch exercised
    base = sum(eng.launches_per_call(n) for n in (8, 8, 4))
    assert base <= eng.launches <= base + 3          # + the occasional move of the history lines
    eng.close()


def test_golden_fixture_float_and_cli(api, model0):
    """Outputs of the REAL reference, committed under tests/golden (make_golden.py)."""
    g = np.load(os.path.join(GOLDEN, "e2e.npz"))
    assert bytes(g["digest"]).decode() == model0.digest()
    x16 = g["x16"]
    F = x16.size // 480
    for name, scale in (("unit", np.float32(1 / 32768.0)), ("int16", np.float32(1.0))):
        eng = api.Engine(1, 16, model0)
        out, gr = eng.process_stream_chunks((x16.astype(np.float32) * scale)[None, :], want_gr=True)
        eng.close()
        ref_gr = g[f"gr_{name}"]
        rel = np.abs(gr[:, 0, :] - ref_gr) / np.maximum(np.abs(ref_gr), 1e-6)
        assert rel.max() < GR_RTOL
        assert _lsb_diff(out[0], g[f"out_{name}"], name == "unit") <= PCM_LSB
    # int16 wire format == src/main.cpp:30-39 (first output hop dropped by the caller)
    eng = api.Engine(1, F, model0)
    o16, gr = eng.process(x16[None, :].copy(), want_gr=True)
    eng.close()
    d = np.abs(o16[0, 480:].astype(np.int32) - g["cli_out16"].astype(np.int32))
    assert d.max() <= PCM_LSB
    assert (d != 0).mean() < 0.01


def test_hot_weights_saturated_activations(api, oracle, model_hot):
    """Weights scaled x3 drive the GRU gates into saturation (table tails, clamp at index 200).  Unit-scale
    input keeps the pre-activations inside the domain where the reference's float->int conversion is defined."""
    x = _inputs(1.0, 10, n_synth=3)[:5]
    ref_out, ref_gr, _ = _oracle_run(oracle, model_hot, x)
    assert np.isfinite(ref_gr).all()
    eng = api.Engine(x.shape[0], 10, model_hot)
    out, gr = eng.process(x, want_gr=True)
    eng.close()
    assert ref_gr.max() - ref_gr.min() > 0.5
    assert np.abs(gr - ref_gr).max() < 1e-4
    assert _lsb_diff(out, ref_out, True) <= PCM_LSB


def test_batch_properties_at_config_size(api, model0):
    """BASELINE.json config 2 size (1024 streams): size-independent properties --
    a stream's result does not depend on what else is in the batch, nor on how the hops are chunked,
    and the run is deterministic."""
    base = _inputs(32768.0, 8, n_synth=8)[:8]
    S = 1024
    x = np.tile(base, (S // 8, 1))
    eng = api.Engine(S, 8, model0)
    out, gr = eng.process(x, want_gr=True)
    eng.reset()
    out2, gr2 = eng.process(x, want_gr=True)
    eng.close()
    assert np.array_equal(out, out2) and np.array_equal(gr, gr2)                   # deterministic / reset
    assert np.array_equal(out.reshape(S // 8, 8, -1), np.broadcast_to(out[:8], (S // 8, 8, out.shape[1])))
    small = api.Engine(8, 3, model0)                                                 # different batch + chunking 3,3,2
    o_s, g_s = small.process_stream_chunks(base, want_gr=True)
    small.close()
    assert np.array_equal(o_s, out[:8])
    assert np.array_equal(g_s, gr[:, :8])


def test_latency_and_impulse(api, model0):
    """Output hop t carries input hop t-6 (5 hops look-ahead + overlap-add): silence in, silence out; an
    impulse first shows up six hops later."""
    x = np.zeros((1, 12 * 480), np.float32)
    eng = api.Engine(1, 12, model0)
    out, _ = eng.process(x)
    assert np.all(out == 0)
    eng.reset()
    x[0, 100] = 10000.0
    out, _ = eng.process(x)
    eng.close()
    first = np.nonzero(out[0])[0][0]
    assert 5 * 480 <= first < 6 * 480 + 100


def test_postfilter_flag(api, oracle, model0):
    x = _inputs(32768.0, 9, n_synth=2)[:3]
    ref_out, _, taps = _oracle_run(oracle, model0, x, flags=1)
    ref_g = np.stack([[t.np("g_used") for t in tp] for tp in taps], axis=1)          # [F, S, 34]
    eng = api.Engine(3, 9, model0, api.NN_FP32 | api.POSTFILTER | api.KEEP_TAPS)
    out, _ = eng.process(x)
    gr_used = eng.read_tap("g_used", 9)
    eng.close()
    assert np.abs(np.trunc(out.astype(np.float64)) - np.trunc(ref_out.astype(np.float64))).max() <= PCM_LSB
    rel = np.abs(gr_used - ref_g) / np.maximum(np.abs(ref_g), 1e-6)
    # the post-filtered gains themselves: the fp32 network path reproduces g to 5e-7 (its contraction sums in another
    # order than the reference), the filter's common factor G couples the bands -- 1e-6 was measured, 3e-6 is the bar
    assert rel.max() < 3e-6
    plain = api.Engine(3, 9, model0)
    out_p, _ = plain.process(x)
    plain.close()
    assert np.abs(out - out_p).max() > 1.0           # and it does change the signal


def test_device_pointer_entry_and_stream(api, model0):
    import torch
    x = _inputs(1.0, 6, n_synth=4)[:4]
    eng = api.Engine(4, 6, model0)
    ref, _ = eng.process(x)
    eng.reset()
    d_in = torch.from_numpy(x).cuda()
    d_out = torch.empty_like(d_in)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        eng.process_device(d_in.data_ptr(), d_in.stride(0), d_out.data_ptr(), d_out.stride(0), 6,
                           stream=st.cuda_stream)
    st.synchronize()
    eng.close()
    assert np.array_equal(d_out.cpu().numpy(), ref)


def test_pipelined_submit_matches_blocking(api, model0):
    """pnb_submit_host_* / pnb_wait (copies overlapped with kernels) gives the blocking call's bits, in both
    wire formats, across more calls than pipeline slots."""
    from percepnet_b200.synth import to_int16
    x = _inputs(1.0, 20, n_synth=4)[:6]
    S = x.shape[0]
    for cast in (lambda a: a, to_int16):
        xi = cast(x)
        eng = api.Engine(S, 4, model0)
        want, _ = eng.process_stream_chunks(xi)
        eng.reset()
        outs = [np.empty((S, 4 * 480), xi.dtype) for _ in range(5)]
        ins = [np.ascontiguousarray(xi[:, k * 1920:(k + 1) * 1920]) for k in range(5)]
        for k in range(5):
            eng.submit(ins[k], outs[k])
        eng.wait()
        eng.close()
        assert np.array_equal(np.concatenate(outs, axis=1), want)


def test_argument_errors(api, model0):
    eng = api.Engine(2, 4, model0)
    with pytest.raises(api.PnbError):
        eng.process(np.zeros((2, 5 * 480), np.float32))      # more hops than max_frames_per_call
    with pytest.raises(api.PnbError):
        eng.read_tap("pitch", 1)                              # taps not enabled
    eng.close()
    with pytest.raises(api.PnbError):
        api.Engine(0, 4, model0)


def test_random_mixtures_batch(api, oracle, model0):
    """48 seeded random streams (harmonics, noise, DC, silent gaps, clicks; amplitudes 1e-4 .. 1) in one batch:
    DSP taps bit-exact, g/r and PCM within the bars."""
    from util import random_mixtures
    F = 18
    x = random_mixtures(48, F)
    eng = api.Engine(48, F, model0, api.NN_FP32 | api.KEEP_TAPS)
    out, gr = eng.process(x, want_gr=True)
    feats, pitch, ex = eng.read_tap("features", F), eng.read_tap("pitch", F), eng.read_tap("Ex", F)
    eng.close()
    ref_out, ref_gr, taps = _oracle_run(oracle, model0, x)
    for s in range(48):
        assert same_bits(feats[:, s], np.stack([t.np("features") for t in taps[s]])), s
        assert same_bits(ex[:, s], np.stack([t.np("Ex") for t in taps[s]])), s
        assert np.array_equal(pitch[:, s, 1], np.array([t.pitch_index for t in taps[s]])), s
        assert np.array_equal(pitch[:, s, 2], np.array([t.silence for t in taps[s]])), s
    rel = np.abs(gr - ref_gr) / np.maximum(np.abs(ref_gr), 1e-6)
    assert rel.max() < GR_RTOL
    assert _lsb_diff(out, ref_out, True) <= PCM_LSB
