"""Batch enhancement CLI: the GPU counterpart of the reference's ``percepNet_run`` binary
(/root/reference/src/main.cpp), for one file with the reference's argv or for a list of files as one batch.

  python -m percepnet_b200.enhance --weights model.pnbw <in.pcm> <out.pcm>
  python -m percepnet_b200.enhance --weights model.pnbw --list jobs.txt [--batch 4096] [--chunk 100] [--nn tensor]

File format and framing are main.cpp's (:30-39): raw int16 mono 48 kHz in; frames of 480 samples are read until
the first short read (a trailing partial frame is dropped); every processed frame except the first is written, so
the output holds (n_frames - 1) * 480 samples, converted with C truncation of x * 32768.  ``--gr <file>`` also
stores the per-frame g[34], r[34] the reference dumps to feature_test.raw (denoise.cpp:533-534).

Weights come from a binary weight file (percepnet_b200.weights.PackedModel.save_blob) -- the
reference binary has them compiled in.  ``--seed N`` uses the synthetic random-init model instead (tests, benches).
All files of a list run as independent streams of one engine, padded to the longest file of their batch.
There is no CPU path: without the CUDA library / a GPU this exits with an error.
"""
from __future__ import annotations

import argparse
import sys

import numpy as np

FRAME = 480


def run_jobs(jobs, model, nn: str = "tensor", batch: int = 4096, chunk: int = 100, device: int = 0, gr_paths=None) -> None:
    """jobs: list of (in_path, out_path); gr_paths: optional list of paths (or None) for the g/r dump of each job."""
    from . import api
    flags = api.NN_TENSOR if nn == "tensor" else api.NN_FP32
    order = sorted(range(len(jobs)), key=lambda k: -_n_frames(jobs[k][0]))       # similar lengths share a batch
    for b0 in range(0, len(order), batch):
        group = order[b0:b0 + batch]
        nfr = [_n_frames(jobs[k][0]) for k in group]
        S, Fmax = len(group), max(nfr)
        if Fmax == 0:
            for k in group:
                open(jobs[k][1], "wb").close()
            continue
        x = np.zeros((S, Fmax * FRAME), np.int16)
        for row, k in enumerate(group):
            x[row, :nfr[row] * FRAME] = np.fromfile(jobs[k][0], dtype="<i2", count=nfr[row] * FRAME)
        eng = api.Engine(S, min(chunk, Fmax), model, flags, device=device)
        outs, grs = [], []
        want_gr = gr_paths is not None and any(gr_paths[k] for k in group)
        for t0 in range(0, Fmax, chunk):
            t1 = min(Fmax, t0 + chunk)
            o, g = eng.process(np.ascontiguousarray(x[:, t0 * FRAME:t1 * FRAME]), want_gr=want_gr)
            outs.append(o)
            grs.append(g)
        eng.close()
        out = np.concatenate(outs, axis=1)
        gr = np.concatenate(grs, axis=0) if want_gr else None                       # [F, S, 68]
        for row, k in enumerate(group):
            n = nfr[row]
            out[row, FRAME:n * FRAME].astype("<i2").tofile(jobs[k][1])            # first frame dropped, main.cpp:37-38
            if want_gr and gr_paths[k]:
                gr[:n, row].astype("<f4").tofile(gr_paths[k])


def _n_frames(path: str) -> int:
    import os
    return os.path.getsize(path) // (2 * FRAME)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("args", nargs="*", help="<in.pcm> <out.pcm>")
    ap.add_argument("--list", help="text file, one '<in.pcm> <out.pcm>' job per line")
    ap.add_argument("--weights", help="binary weight file (PackedModel.save_blob)")
    ap.add_argument("--seed", type=int, help="use the synthetic random-init model with this seed instead of --weights")
    ap.add_argument("--nn", default="tensor", choices=["tensor", "fp32"])
    ap.add_argument("--gr", help="single-file mode: also write the per-frame g, r (feature_test.raw of the reference)")
    ap.add_argument("--batch", type=int, default=4096, help="files per GPU batch")
    ap.add_argument("--chunk", type=int, default=100, help="frames per GPU call")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    if a.list:
        jobs = [tuple(line.split()[:2]) for line in open(a.list) if line.split()]
        grp = None
    elif len(a.args) == 2:
        jobs = [(a.args[0], a.args[1])]
        grp = [a.gr]
    else:
        sys.stderr.write("usage: percepnet_b200.enhance <in.pcm> <out.pcm>   (or --list jobs.txt)\n")
        return 1
    from . import api
    from .weights import synth_model
    if a.weights:
        model = api.BlobModel(a.weights)
    elif a.seed is not None:
        model = synth_model(a.seed)
    else:
        sys.stderr.write("give --weights <file.pnbw> (or --seed N for the synthetic model)\n")
        return 1
    run_jobs(jobs, model, a.nn, a.batch, a.chunk, a.device, grp)
    return 0


if __name__ == "__main__":
    sys.exit(main())
