#!/usr/bin/env python
"""Per-role cycle shares of gru_chain_kernel from a -DPNB_CHAIN_TIMING build
(percepnet_b200/libpercepnet_b200_chain_timing.so, `python percepnet_b200/build.py --chain-timing`)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from percepnet_b200 import api  # noqa: E402

NAMES = {0: "producer: decode + issue", 1: "producer: dependency wait", 2: "producer: wait free stage",
         4: "mma: decode + issue", 5: "mma: wait accumulator (epilogue)", 6: "mma: wait operands (TMA)",
         8: "epilogue: work", 9: "epilogue: dependency wait + prefetch", 10: "epilogue: wait accumulator (MMA)", 11: "epilogue: publish"}


def main():
    api.LIB_PATH = os.path.join(os.path.dirname(api.LIB_PATH), "libpercepnet_b200_chain_timing.so")
    os.environ["PNB_OVERLAP"] = os.environ.get("PNB_OVERLAP", "0")
    import torch
    from percepnet_b200.weights import synth_model
    from bench import make_step_inputs
    S, F = 16384, 8
    dev = torch.device("cuda", 0)
    eng = api.Engine(S, F, synth_model(0), api.NN_TENSOR)
    bufs = make_step_inputs(S, F, 2, 0, dev)
    out = torch.empty_like(bufs[0])
    st = torch.cuda.current_stream()
    cyc = (C.c_ulonglong * 16)()
    for i in range(4):
        eng.process_device(bufs[i % 2].data_ptr(), bufs[i % 2].stride(0), out.data_ptr(), out.stride(0), F, stream=st.cuda_stream)
        if i == 1:
            eng.L.pnb_debug_chain_cycles(cyc, 1)
    eng.L.pnb_debug_chain_cycles(cyc, 1)
    v = list(cyc)
    res = {}
    for base, n, who in ((0, 3, "producer"), (4, 3, "mma"), (8, 4, "epilogue")):
        tot = sum(v[base:base + n]) or 1
        for k in range(n):
            res[NAMES[base + k]] = round(v[base + k] / tot, 4)
        res[f"{who} total Mcycles (sum over CTAs/launches)"] = round(tot / 1e6, 1)
    print(json.dumps(res, indent=1))
    eng.close()


if __name__ == "__main__":
    main()
