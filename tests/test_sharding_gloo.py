"""world_size-2 gloo test of the N>1 host logic (CPU): the stream partition, the max-over-ranks timing
reduction, and the property that makes sharding legal -- a stream's output does not depend on which other
streams share its batch (checked with the oracle standing in for the device)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_streams, n_frames, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import ffi
    from percepnet_b200.sharding import aggregate_throughput, shard_range
    from percepnet_b200.synth import synth_pcm
    from percepnet_b200.weights import synth_model
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total_streams, world, rank)
    model = synth_model(0)
    x = synth_pcm(hi - lo, n_frames, seed=900, first_stream=lo)     # stream k is the same signal whatever the shard
    y = ffi.Oracle().process_streams(model, x, 1)
    units, ms, rate = aggregate_throughput((hi - lo) * n_frames, 10.0 * (rank + 1))
    np.save(os.path.join(out_dir, f"y{rank}.npy"), y)
    if rank == 0:
        np.save(os.path.join(out_dir, "agg.npy"), np.array([units, ms, rate, lo, hi]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    from oracle import ffi
    from percepnet_b200.sharding import shard_range
    from percepnet_b200.synth import synth_pcm
    from percepnet_b200.weights import synth_model
    ffi.build()
    total, F, world = 5, 4, 2
    assert [shard_range(total, world, r) for r in range(world)] == [(0, 3), (3, 5)]
    assert [shard_range(16384 * 8, 8, r)[0] for r in range(8)] == [16384 * r for r in range(8)]
    mp.spawn(_worker, args=(world, _free_port(), total, F, str(tmp_path)), nprocs=world, join=True)
    agg = np.load(tmp_path / "agg.npy")
    assert agg[0] == total * F            # units summed over ranks
    assert agg[1] == 20.0                 # elapsed = max over ranks
    assert abs(agg[2] - total * F / 0.020) < 1e-6
    whole = ffi.Oracle().process_streams(synth_model(0), synth_pcm(total, F, seed=900), 1)
    got = np.concatenate([np.load(tmp_path / f"y{r}.npy") for r in range(world)])
    assert np.array_equal(got, whole)     # sharded == unsharded, bit for bit
