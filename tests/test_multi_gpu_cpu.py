"""The N > 1 path on CPU: world_size-2 gloo processes shard the tracks by contiguous range (no data-path
collective), each rank analyses its own range with the CPU checker standing in for the GPU, and rank 0 gathers
timing the way bench.py does (max over ranks).  The union of the shards must equal the single-process result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (os.path.join(ROOT, "feature-extractor_b200"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import fxb200
import oracle_util as ou


def test_shard_tracks_partitions_exactly():
    for n, w in ((4096, 8), (65536, 8), (10, 3), (7, 8), (1, 1)):
        spans = [fxb200.shard_tracks(n, w, r) for r in range(w)]
        assert spans[0][0] == 0
        for (a, c), (b, _) in zip(spans[:-1], spans[1:]):
            assert a + c == b
        assert spans[-1][0] + spans[-1][1] == n
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        fxb200.shard_tracks(8, 2, 2)


def _worker(rank, world, port, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T, N, H, sr = 6, 1024, 512, 44100.0
    S = 40 * H
    first, count = fxb200.shard_tracks(T, world, rank)
    audio = ou.make_tracks(count, S, sr, first_track=first)         # each rank synthesises only its own tracks
    dist.barrier()
    r = ou.port().analyse(audio, window=N, hop=H, sample_rate=sr, threads=1)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)        # stand-in for the per-rank elapsed time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                        # the only collective: timing, never data
    np.save(os.path.join(tmpdir, f"raw_{rank}.npy"), r["raw"])
    if rank == 0:
        assert t.item() == float(world)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 2
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    shards = [np.load(tmp_path / f"raw_{r}.npy") for r in range(world)]
    whole = ou.port().analyse(ou.make_tracks(6, 40 * 512, 44100.0), window=1024, hop=512, sample_rate=44100.0)["raw"]
    assert np.array_equal(np.concatenate(shards, axis=0), whole, equal_nan=True)
