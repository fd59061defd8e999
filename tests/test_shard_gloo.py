"""N > 1 host logic on CPU: world-size-2 gloo run of the batched (one-image-per-GPU) driver with the
compute call replaced by the CPU oracle (tests may use the oracle; the product path may not)."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "graph-algorithm-image-segmentation-gpgpu_b200"


def test_shard_covers_every_image_once():
    batch = importlib.import_module(PKG + ".batch")
    for n in (0, 1, 7, 256):
        for world in (1, 2, 3, 8):
            got = sorted(i for r in range(world) for i in batch.shard(n, r, world))
            assert got == list(range(n))
            sizes = [len(batch.shard(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        batch.shard(4, 2, 2)


def _worker(rank, world, port, n_items, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch = importlib.import_module(PKG + ".batch")
    from oracle import oracle as O
    seen = []

    def load(i):
        seen.append(i)
        return O.synth(48, 36, 100 + i)

    def seg_one(img):
        return O.pipeline(img, 0.8, 300.0, 20, 8, O.FELZ)["n"]

    counts = batch.segment_sharded(n_items, load, seg_one, dist)
    q.put((rank, seen, counts))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_batch_matches_single_process():
    import torch.multiprocessing as mp
    from oracle import oracle as O
    O.build()
    batch = importlib.import_module(PKG + ".batch")
    n_items = 5
    ref = batch.segment_sharded(n_items, lambda i: O.synth(48, 36, 100 + i),
                                lambda img: O.pipeline(img, 0.8, 300.0, 20, 8, O.FELZ)["n"])
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in ps]
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    res.sort()
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]       # image i -> rank i mod 2
    assert res[0][2] == ref and res[1][2] == ref                 # every rank sees the whole batch's summary
    assert all(c > 0 for c in ref)
