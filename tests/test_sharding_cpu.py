"""CPU, world size 2, gloo: the host-side logic of the multi-GPU path (bench.py / DESIGN.md section 4) -- contiguous
stream shards, a single broadcast of the model file from rank 0, identical per-rank packing -- without any GPU."""
import json
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, golden_files, load_golden, model_file_for


def shard_bounds(total_streams, world, rank):
    """contiguous blocks, remainder to the first ranks (what bench.py uses with streams_per_gpu fixed = weak scaling)"""
    base, rem = divmod(total_streams, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def test_shards_partition_the_batch():
    for total, world in [(32768, 8), (4096, 2), (4097, 4), (5, 8)]:
        spans = [shard_bounds(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, model_path, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import neuralaudio_b200 as na
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank 0 reads the file; ONE broadcast carries it (bench.py does the same over NCCL)
    if rank == 0:
        data = torch.frombuffer(bytearray(open(model_path, "rb").read()), dtype=torch.uint8)
        size = torch.tensor([data.numel()])
    else:
        size = torch.zeros(1, dtype=torch.int64)
    dist.broadcast(size, src=0)
    if rank != 0:
        data = torch.empty(int(size.item()), dtype=torch.uint8)
    dist.broadcast(data, src=0)
    path = os.path.join(out_dir, "rank%d.nam" % rank)
    with open(path, "wb") as f:
        f.write(data.numpy().tobytes())
    desc = na.describe_model_file(path)            # host-side parse + pack, identical on every rank
    lo, hi = shard_bounds(32768, world, rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, (json.dumps(desc, sort_keys=True), lo, hi))
    if rank == 0:
        with open(os.path.join(out_dir, "result.json"), "w") as f:
            json.dump(gathered, f)
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_load_world_size_2(tmp_path):
    import torch.multiprocessing as mp
    g = load_golden(golden_files("syn_a1_standard")[0])
    mf = model_file_for(g, tmp_path)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, mf, str(tmp_path)), nprocs=2, join=True)
    res = json.load(open(tmp_path / "result.json"))
    assert res[0][0] == res[1][0]                                  # same packed model on both ranks
    assert (res[0][1], res[0][2], res[1][1], res[1][2]) == (0, 16384, 16384, 32768)
    d = json.loads(res[0][0])
    assert d["kind"] == "wavenet" and d["num_weights"] == 13802
