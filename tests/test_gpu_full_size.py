"""BASELINE.json's full single-GPU size (config 3: 16 384 concurrent streams, tensor-core network; and 8 192
file pairs for the training-data generator) checked through properties that do not need 16 384 oracle runs:
every replica of a base stream must come out bit-identical wherever it sits in the batch (every tile, CTA pair and
row position computes the same thing), and the base streams themselves are anchored to the oracle."""
import numpy as np
import pytest

from util import same_bits

pytestmark = pytest.mark.gpu

S_FULL, N_BASE, F = 16384, 48, 8


@pytest.fixture(scope="module")
def api():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    from percepnet_b200 import api
    api.load_library()
    return api


def test_full_batch_enhancement_replicas_and_oracle(api, oracle, model0):
    from percepnet_b200.synth import synth_pcm
    base = synth_pcm(N_BASE, 2 * F, seed=2468)
    idx = np.arange(S_FULL) % N_BASE
    x = base[idx]
    eng = api.Engine(S_FULL, F, model0, api.NN_TENSOR)
    outs, grs = [], []
    for c in range(2):                                        # two calls: the carried state is exercised at full size
        o, g = eng.process(np.ascontiguousarray(x[:, c * F * 480:(c + 1) * F * 480]), want_gr=True)
        outs.append(o); grs.append(g)
    eng.close()
    out, gr = np.concatenate(outs, axis=1), np.concatenate(grs, axis=0)          # [S, 16*480], [16, S, 68]
    assert same_bits(out, out[:N_BASE][idx]), "replicas of a stream differ across the batch"
    assert same_bits(gr, gr[:, :N_BASE][:, idx]), "replica g/r differ across the batch"
    for s in range(6):                                        # anchor: the oracle on the base streams
        h = oracle.create(model0)
        ref, ref_gr, _ = oracle.process_stream(h, base[s], True)
        oracle.destroy(h)
        rel = np.abs(gr[:, s] - ref_gr) / np.maximum(np.abs(ref_gr), 1e-6)
        assert rel.max() < 1e-4, (s, rel.max())
        lsb = np.abs(np.trunc(out[s].astype(np.float64) * 32768) - np.trunc(ref.astype(np.float64) * 32768)).max()
        assert lsb <= 1, (s, lsb)
    assert np.abs(out).max() > 1e-3                            # the batch did produce audio


def test_full_batch_training_records_replicas_and_oracle(api, oracle):
    from percepnet_b200.synth import synth_pairs
    N = S_FULL // 2
    clean, noisy = synth_pairs(N_BASE, F, seed=1357)
    idx = np.arange(N) % N_BASE
    eng = api.Engine(2 * N, F, None, api.TRAIN_DATA)
    rec = eng.train_records(clean[idx], noisy[idx])
    eng.close()
    assert same_bits(rec, rec[:N_BASE][idx])
    for k in range(4):
        want = oracle.train_records(clean[k], noisy[k])
        assert same_bits(rec[k][:, :70], want[:, :70]) and same_bits(rec[k][:, 104:], want[:, 104:])
        assert np.allclose(rec[k][:, 70:104], want[:, 70:104], rtol=1e-6, atol=1e-12)
