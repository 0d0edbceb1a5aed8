"""python -m percepnet_b200.enhance: the batch counterpart of the reference's percepNet_run (src/main.cpp:30-39
framing: whole frames until the first short read, first output frame dropped, truncating int16 conversion),
against the oracle's restatement of that CLI (pn_oracle_run_pcm16, pinned to the real binary in
tests/test_dropin_cli.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    from percepnet_b200 import api
    api.load_library()
    return api


def test_enhance_list_and_single_file(api, oracle, model0, tmp_path):
    from percepnet_b200 import enhance
    from percepnet_b200.synth import synth_pcm, to_int16
    lengths = [37 * 480, 12 * 480 + 133, 25 * 480, 480, 200]          # whole frames, a partial tail, one frame, < one frame
    pcm = [to_int16(synth_pcm(1, 40, seed=600 + k)[0])[:n] for k, n in enumerate(lengths)]
    wpath = str(tmp_path / "w.pnbw")
    model0.save_blob(wpath)
    jobs = []
    for k, p in enumerate(pcm):
        fi, fo = str(tmp_path / f"in{k}.pcm"), str(tmp_path / f"out{k}.pcm")
        p.tofile(fi)
        jobs.append((fi, fo))
    (tmp_path / "jobs.txt").write_text("".join(f"{a} {b}\n" for a, b in jobs))
    want = []
    for p in pcm:
        nf = p.size // 480
        want.append(oracle.run_pcm16(model0, p[:nf * 480]) if nf else (np.zeros(0, np.int16), np.zeros((0, 68), np.float32)))
    for nn in ("fp32", "tensor"):
        assert enhance.main(["--weights", wpath, "--list", str(tmp_path / "jobs.txt"), "--nn", nn, "--chunk", "16"]) == 0
        for k, (_, fo) in enumerate(jobs):
            got = np.fromfile(fo, np.int16)
            ref = want[k][0]
            assert got.shape == ref.shape == (max(lengths[k] // 480 - 1, 0) * 480,), (nn, k)
            if got.size:
                assert np.abs(got.astype(np.int32) - ref.astype(np.int32)).max() <= 1, (nn, k)
    # the reference's argv, with the g/r dump
    fo, fg = str(tmp_path / "single.pcm"), str(tmp_path / "single.gr")
    assert enhance.main(["--weights", wpath, "--nn", "fp32", "--gr", fg, jobs[0][0], fo]) == 0
    assert np.abs(np.fromfile(fo, np.int16).astype(np.int32) - want[0][0].astype(np.int32)).max() <= 1
    gr = np.fromfile(fg, np.float32).reshape(-1, 68)
    ref_gr = want[0][1]
    assert gr.shape == ref_gr.shape
    assert (np.abs(gr - ref_gr) / np.maximum(np.abs(ref_gr), 1e-6)).max() < 1e-4
    assert enhance.main(["only-one-arg"]) == 1 and enhance.main(["a", "b"]) == 1      # usage / no weights
