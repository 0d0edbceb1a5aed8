#!/usr/bin/env python
"""bench.py -- 48 kHz frames/sec of the PercepNet enhancement hot path (rnnoise_process_frame) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
                  [--streams S] [--frames F] [--nn fp32|tensor]

One "step" = one pnb_process call: S concurrent streams x F hops of 480 samples per GPU.
Default workload = BASELINE.json config 3/4: 16 384 streams per GPU (weak scaling: 131 072 on 8 GPUs),
synthetic 48 kHz PCM, random-init weights of the reference architecture.  `--streams 1024 --nn fp32`
is config 2.  One JSON line on stdout (rank 0).

  value    whole-job frames/s with the input PCM already resident in HBM (device-pointer entry),
           timed with CUDA events on the launching stream, max over ranks
  e2e      the same metric through the public pipelined host call (Engine.submit/wait ==
           pnb_submit_host_i16 + pnb_wait): int16 PCM in pinned host memory -> H2D -> hot path -> D2H ->
           int16 PCM in pinned host memory, every step, inside the timed region (host clock)
  roofline the dominant kernel (the network contraction) against the measured tensor peak
  cpu_baseline  the reference's CPU path timed on this box's host cores on a bounded sample

`--impl reference` times the reference's own CPU implementation (oracle/_ref when it was compiled in the
build container, else the oracle port) -- the only place besides cpu_baseline where oracle/ is executed.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAME = 480
FLOP_PER_FRAME = 15_896_576       # network MAC*2 per hop
BYTES_PER_FRAME = 3_840           # 480 f32 in + 480 f32 out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------- inputs
def make_step_inputs(n_streams, n_frames, n_buffers, rank, device):
    """n_buffers distinct [S, F*480] float32 device tensors (unit scale like src/main.cpp:34 feeds the API).
    64 base streams (harmonic + noise, percepnet_b200.synth) are tiled with per-stream gains and circular
    time offsets so that every stream differs; resident in HBM before timing."""
    import torch
    from percepnet_b200.synth import synth_pcm
    T = n_frames * FRAME
    base_n = min(64, n_streams)
    base = torch.from_numpy(synth_pcm(base_n, n_frames * n_buffers, seed=1234 + 1000 * rank)).to(device)
    g = torch.Generator(device="cpu").manual_seed(99 + rank)
    gains = (0.25 + 0.75 * torch.rand(n_streams, 1, generator=g)).to(device)
    idx = torch.arange(n_streams, device=device) % base_n
    bufs = []
    for b in range(n_buffers):
        seg = base[:, b * T:(b + 1) * T]
        bufs.append((seg[idx] * gains).contiguous())
    return bufs


# ------------------------------------------------------------------------------------- CPU legs
def cpu_reference_run(n_streams, n_frames, threads, seed=4321, model=None):
    """Times the reference's CPU path on `n_streams` independent streams x n_frames hops, OpenMP over
    streams.  Returns (frames_per_s, seconds, kind)."""
    from oracle import ffi
    from percepnet_b200.synth import synth_pcm
    from percepnet_b200.weights import synth_model
    model = model or synth_model(0)
    x = synth_pcm(min(n_streams, 16), n_frames, seed=seed)
    x = np.ascontiguousarray(np.tile(x, ((n_streams + x.shape[0] - 1) // x.shape[0], 1))[:n_streams])
    if ffi.Reference.available():
        R = ffi.Reference()
        R.set_model(model)
        t0 = time.perf_counter()
        R.process_streams(x, threads)
        dt = time.perf_counter() - t0
        kind = "reference"
    else:
        ffi.build()
        O = ffi.Oracle()
        t0 = time.perf_counter()
        O.process_streams(model, x, threads)
        dt = time.perf_counter() - t0
        kind = "port"
    return n_streams * n_frames / dt, dt, kind


def cpu_baseline_leg(budget_s=14.0):
    cores = os.cpu_count() or 1
    threads = cores
    # calibrate on a short run, then size the sample for ~budget_s of wall time
    fps, dt, kind = cpu_reference_run(threads, 20, threads)
    frames = int(max(40, min(4000, budget_s * fps / threads)))
    fps, dt, kind = cpu_reference_run(threads, frames, threads)
    return {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
            "sample": f"{threads} streams x {frames} hops ({threads * frames} frames, {dt:.1f} s wall), "
                      f"OpenMP over streams, g++ -O3 scalar build (the reference's only buildable configuration)",
            "x_realtime": fps / 100.0}


# ------------------------------------------------------------------------------------- row f1
def run_traindata(args):
    """Training-record generator (pnb_train_records_*, reference: train(), src/denoise.cpp:600-787).
    One step = one call: `--streams` PAIRS of (speech, noisy) int16 streams x `--frames` hops; a record is the
    138 floats train() writes per frame.  Single GPU (pairs are independent; shard them like streams)."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from percepnet_b200 import api
    from percepnet_b200.synth import synth_pairs
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the generator has no CPU fallback"}))
        return 2
    N, F, K, W = (8192 if args.streams == 16384 else args.streams), args.frames, args.steps, max(args.warmup, 3)
    device = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    n_buf = 3
    base_c, base_n = synth_pairs(32, F * n_buf, seed=2024)
    idx = np.arange(N) % 32
    shift = (np.arange(N) // 32 % 5).astype(np.int16)                 # per-pair level, so that every pair differs
    T = F * FRAME
    h_c = [torch.from_numpy(np.ascontiguousarray(base_c[idx, b * T:(b + 1) * T] >> shift[:, None])).pin_memory() for b in range(n_buf)]
    h_n = [torch.from_numpy(np.ascontiguousarray(base_n[idx, b * T:(b + 1) * T] >> shift[:, None])).pin_memory() for b in range(n_buf)]
    d_c = [t.to(device) for t in h_c]
    d_n = [t.to(device) for t in h_n]
    d_rec = [torch.empty((N, F, api.RECORD), dtype=torch.float32, device=device) for _ in range(n_buf)]
    h_rec = [torch.empty((N, F, api.RECORD), dtype=torch.float32).pin_memory() for _ in range(n_buf)]
    eng = api.Engine(2 * N, F, None, api.TRAIN_DATA)
    stream = torch.cuda.current_stream()

    def step(i):
        b = i % n_buf
        eng.train_records_device(d_c[b].data_ptr(), T, d_n[b].data_ptr(), T, F, d_rec[b].data_ptr(), F * api.RECORD,
                                 stream=stream.cuda_stream)
    for i in range(W):
        step(i)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = eng.launches
    ev0.record(stream)
    for i in range(K):
        step(W + i)
    ev1.record(stream)
    torch.cuda.synchronize()
    launches = eng.launches - launches0
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    value = N * F * K / (ms * 1e-3)
    eng.profile(True)
    step(0)
    prof = eng.profile_read()
    eng.profile(False)
    # end to end: int16 files in pinned host memory -> records in pinned host memory through the pipelined public
    # call (pnb_submit_train_records + pnb_wait), copies inside the timed region (host clock)
    L = eng.L

    def submit(i):
        b = i % n_buf
        rc = L.pnb_submit_train_records(eng.h, h_c[b].data_ptr(), T, h_n[b].data_ptr(), T, F, h_rec[b].data_ptr(), F * api.RECORD)
        if rc != 0:
            raise RuntimeError(L.pnb_last_error().decode())
    for i in range(3):
        submit(i)
    L.pnb_wait(eng.h)
    t0 = time.perf_counter()
    for i in range(K):
        submit(i)
    L.pnb_wait(eng.h)
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    for i in range(3):
        L.pnb_train_records_host(eng.h, h_c[i % n_buf].data_ptr(), T, h_n[i % n_buf].data_ptr(), T, F, h_rec[i % n_buf].data_ptr(), F * api.RECORD)
    blocking = N * F * 3 / (time.perf_counter() - t0)
    eng.close()
    peaks = measured_peaks()
    ana_ms = prof.get("analysis_kernel", (0.0, 0))[0]
    bytes_per_record = 2 * FRAME * 2 + api.RECORD * 4                 # two int16 hops in, one record out
    roof = {"bound": "hbm", "kernel": "analysis_kernel", "achieved": N * F * bytes_per_record / (ana_ms * 1e-3) / 1e9 if ana_ms else None,
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "traffic": None,
            "note": "the analysis kernel (2 streams per record) is bound by shared-memory wavefronts and instruction issue, "
                    "not HBM (profiles/README.md); the HBM figure only shows how far from the memory roof it sits",
            "breakdown_ms": {k: round(v[0], 4) for k, v in prof.items()}}
    if roof["achieved"]:
        roof["frac"] = roof["achieved"] / roof["peak"]
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import ffi
        ffi.build()
        O = ffi.Oracle()
        cores = os.cpu_count() or 1
        nfr = 1500
        cc, nn_ = synth_pairs(8, nfr, seed=7)
        jobs = [(cc[k % 8], nn_[k % 8]) for k in range(cores * 6)]
        t0 = time.perf_counter()
        with ThreadPoolExecutor(cores) as ex:                         # the C call releases the GIL
            list(ex.map(lambda j: O.train_records(*j), jobs))
        cdt = time.perf_counter() - t0
        cpu = {"value": len(jobs) * nfr / cdt, "unit": "records/s", "cores": cores, "kind": "port",
               "sample": f"{len(jobs)} file pairs x {nfr} frames over {cores} threads, one pair per task ({cdt:.1f} s wall)"}
    print(json.dumps({
        "metric": "training_records_per_sec", "value": value, "unit": "records/s", "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (+f64 label islands)",
        "data": "synthetic",
        "config": {"workload": f"{N} (speech, noisy) int16 pairs x {F} hops per step, train() record generator (row f1)",
                   "pairs": N, "frames_per_step": F,
                   "l2_policy": f"{n_buf} rotating input/record buffer sets (records {N * F * api.RECORD * 4 / 1e6:.0f} MB each)"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof,
        "e2e": {"value": N * F * K / dt, "unit": "records/s", "h2d_bytes_per_step": 2 * N * T * 2,
                "d2h_bytes_per_step": N * F * api.RECORD * 4, "api": "pnb_submit_train_records + pnb_wait",
                "blocking_call_records_per_s": blocking},
        "cpu_baseline": cpu}))
    return 0


# ------------------------------------------------------------------------------------- main
def run_reference_impl(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    fps0, _, kind = cpu_reference_run(cores, 10, cores)
    frames = int(max(10, min(2000, 4.0 * fps0 / cores)))      # ~4 s per step
    for _ in range(args.warmup):
        cpu_reference_run(cores, max(5, frames // 8), cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_run(cores, frames, cores)
    dt = time.perf_counter() - t0
    fps = cores * frames * args.steps / dt
    line = {
        "impl": "reference", "metric": "48kHz_frames_per_sec", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"reference CPU path, {cores} streams x {frames} hops per step (bounded sample of the "
                               f"{args.streams}-streams-per-GPU workload)", "streams": cores, "frames_per_step": frames},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": f"{cores} streams x {frames} hops x {args.steps} steps, OpenMP over streams"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "x_realtime": fps / 100.0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--streams", type=int, default=16384, help="concurrent streams PER GPU")
    ap.add_argument("--frames", type=int, default=8, help="hops per stream per step")
    ap.add_argument("--nn", default="auto", choices=["auto", "fp32", "tensor"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--path", default="enhance", choices=["enhance", "traindata"],
                    help="enhance = the headline hot path; traindata = SURVEY.md 8 row f1, the training-record generator")
    args = ap.parse_args()
    if args.path == "traindata":
        return run_traindata(args)
    if args.impl == "reference":
        return run_reference_impl(args)

    import torch
    import torch.distributed as dist
    from percepnet_b200 import api
    from percepnet_b200.weights import synth_model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the hot path has no CPU fallback"}))
        return 2
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    S, F, K, W = args.streams, args.frames, args.steps, max(args.warmup, 3)
    model = synth_model(0)
    nn_mode = args.nn
    flags = api.NN_FP32
    if nn_mode in ("auto", "tensor"):
        try:
            eng = api.Engine(S, F, model, api.NN_TENSOR, device=local)
            flags, nn_mode = api.NN_TENSOR, "tensor"
        except api.PnbError:
            if nn_mode == "tensor":
                raise
            eng, nn_mode = None, "fp32"
    if flags == api.NN_FP32:
        nn_mode = "fp32"
        eng = api.Engine(S, F, model, api.NN_FP32, device=local)

    n_buf = 3
    bufs = make_step_inputs(S, F, n_buf, rank, device)
    outs = [torch.empty_like(b) for b in bufs]
    stream = torch.cuda.current_stream()

    def step(i):
        b = i % n_buf
        eng.process_device(bufs[b].data_ptr(), bufs[b].stride(0), outs[b].data_ptr(), outs[b].stride(0), F,
                           stream=stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = eng.launches
    ev0.record(stream)
    for i in range(K):
        step(W + i)
    ev1.record(stream)
    barrier()
    launches = (eng.launches - launches0) * world           # kernels launched inside the timed region (library counter; every rank issues the same schedule)
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    from percepnet_b200.sharding import aggregate_throughput
    _, ms_max, value = aggregate_throughput(S * F * K, ms, device=device)   # sum of frames, max of device time

    # ---- kernel breakdown of one more (untimed) step, CUDA events around every launch -----
    eng.profile(True)
    step(W + K)
    prof = eng.profile_read()
    eng.profile(False)
    nn_cls = "tc_gemm_kernel" if nn_mode == "tensor" else "gemm_f32_kernel"
    nn_ms, nn_n = prof.get(nn_cls, (0.0, 0))
    step_ms_prof = sum(v[0] for v in prof.values())
    peaks = measured_peaks()
    roof = None
    if nn_n:
        flops_per_launch = S * F * FLOP_PER_FRAME / nn_n       # algorithmic flops of the step / contraction launches
        achieved = flops_per_launch / (nn_ms / nn_n * 1e-3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if nn_mode == "tensor" and os.path.exists(tp):   # dram__bytes_read+write of the dominant instance, from the last ncu capture
            tj = json.load(open(tp)).get(nn_cls, {})
            if tj:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        roof = {"bound": "tensor", "kernel": nn_cls, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_note": "DRAM bytes of one GRU-512 launch (ncu, profiles/ncu_traffic.json); the kernel is tensor/L2-bound, DRAM is at 9 % of peak", "peak_source": peaks["source"] + " (sustained bf16 cuBLAS)",
                "launches_per_step": nn_n, "avg_launch_ms": nn_ms / nn_n,
                "share_of_step": nn_ms / step_ms_prof if step_ms_prof else None,
                "pipe": "tcgen05 split-fp16 (3 MMA per product)" if nn_mode == "tensor" else "fp32 FMA (CUDA cores)",
                # the fp32-accurate split issues 3 half-precision MMAs per algorithmic product: what the tensor
                # pipe actually executes, against the same measured peak
                "issued_tflops": achieved * 3 if nn_mode == "tensor" else None,
                "issued_frac": achieved * 3 / peak if nn_mode == "tensor" else None,
                "step_hbm_gbs_algorithmic": S * F * BYTES_PER_FRAME / (ms_max / K * 1e-3) / 1e9,
                "breakdown_ms": {k: round(v[0], 4) for k, v in prof.items()}}
        if nn_mode == "fp32":
            roof["fp32_pipe_frac_of_nominal_80TF"] = achieved / 80.0

    # ---- end to end through the public host-buffer call, pinned memory ---------------------
    e2e = None
    if not args.no_e2e:
        # int16 PCM wire format (what the reference CLI reads and writes, src/main.cpp:30-39), pinned host
        # buffers, through the pipelined public call: every step's PCM crosses PCIe in (H2D) and the enhanced
        # PCM crosses back (D2H) inside the timed region; copies of neighbouring steps overlap the kernels.
        n_host = 3
        h_in = [torch.empty((S, F * FRAME), dtype=torch.int16).pin_memory() for _ in range(n_host)]
        for k in range(n_host):
            h_in[k].copy_((bufs[k % n_buf] * 32768.0).round().clamp(-32768, 32767).to(torch.int16).cpu())
        h_out = [torch.empty((S, F * FRAME), dtype=torch.int16).pin_memory() for _ in range(n_host)]
        L = eng.L

        def e2e_submit(i):
            src, dst = h_in[i % n_host], h_out[i % n_host]
            rc = L.pnb_submit_host_i16(eng.h, src.data_ptr(), src.stride(0), dst.data_ptr(), dst.stride(0), F)
            if rc != 0:
                raise RuntimeError(L.pnb_last_error().decode())

        def e2e_wait():
            if L.pnb_wait(eng.h) != 0:
                raise RuntimeError(L.pnb_last_error().decode())
        Ke = K
        for i in range(3):
            e2e_submit(i)
        e2e_wait()
        barrier()
        t0 = time.perf_counter()
        for i in range(Ke):
            e2e_submit(i)
        e2e_wait()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": world * S * F * Ke / float(t.item()), "unit": "frames/s",
               "h2d_bytes_per_step": S * F * FRAME * 2, "d2h_bytes_per_step": S * F * FRAME * 2,
               "steps": Ke, "ms_per_step": 1e3 * float(t.item()) / Ke,
               "api": "pnb_submit_host_i16 + pnb_wait (percepnet_b200.api.Engine.submit/wait): int16 PCM in pinned host "
                      "memory -> H2D -> hot path -> D2H -> int16 PCM in pinned host memory, timed on the host clock"}
        # the blocking single call, for reference (no overlap between copies and kernels)
        t0 = time.perf_counter()
        for i in range(3):
            src, dst = h_in[i % n_host], h_out[i % n_host]
            L.pnb_process_host_i16(eng.h, src.data_ptr(), src.stride(0), dst.data_ptr(), dst.stride(0), F, None)
        e2e["blocking_call_frames_per_s"] = world * S * F * 3 / (time.perf_counter() - t0)

    eng.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline_leg()
        except Exception as ex:  # the checker must not take the bench down
            cpu = {"error": repr(ex)}

    if rank == 0:
        line = {
            "metric": "48kHz_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if nn_mode == "fp32" else "f32 (network products on tcgen05 as split fp16, fp32 accumulate)",
            "data": "synthetic",
            "config": {"workload": f"{S} concurrent 48 kHz streams per GPU x {F} hops per step "
                                   f"(BASELINE.json config {'2' if S == 1024 else '3/4'}), network={nn_mode}",
                       "streams_per_gpu": S, "frames_per_step": F, "nn": nn_mode, "parallelism": f"streams sharded x{world}, no collective",
                       "l2_policy": f"{n_buf} rotating input buffers of {S * F * FRAME * 4 / 1e6:.0f} MB each (> 126 MB L2 in total)",
                       "weights": "random-init, reference architecture (7,962,564 params)"},
            "x_realtime": value / 100.0, "samples_per_sec": value * FRAME,
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "e2e": e2e, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
