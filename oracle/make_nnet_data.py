#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Writes oracle/_ref/nnet_data.cpp with the REFERENCE's own exporter
(/root/reference/dump_percepnet.py, run unmodified via runpy) from the synthetic seed-0 parameter set of
percepnet_b200.weights -- the file an unmodified reference build would compile as src/nnet_data.cpp.
Needs /root/reference (build container only)."""
import os
import runpy
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)


def main():
    import torch
    from percepnet_b200.weights import synth_state_dict
    for m in ("h5py", "tensorboardX", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["tensorboardX"].SummaryWriter = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib.pyplot"].switch_backend = lambda *a, **k: None
    out_dir = os.path.join(HERE, "_ref")
    os.makedirs(out_dir, exist_ok=True)
    pt = os.path.join(out_dir, "model_seed0.pt")
    torch.save({k: torch.from_numpy(v) for k, v in synth_state_dict(0).items()}, pt)
    sys.argv = ["dump_percepnet.py", pt, os.path.join(out_dir, "nnet_data.cpp")]
    runpy.run_path(os.path.join(REF, "dump_percepnet.py"), run_name="__main__")


if __name__ == "__main__":
    main()
