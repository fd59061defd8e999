"""Batched segmentation: host-side scheduling only (no compute, no fallback).

* `shard(n_items, rank, world)`: which images of a global batch a rank owns (image i -> rank i mod N;
  SURVEY.md section 8e "batched: replicas only, no collective").
* `ContextPool`: S gseg contexts (one CUDA stream each) on one GPU; images are issued round-robin
  with `gseg_segment_async`, so the latency-bound late rounds of one image (a single thread-block
  cluster) overlap the bandwidth-bound early rounds of the next.
* `segment_sharded`: the N-GPU driver: every rank segments its shard and the per-image component
  counts are all-gathered so every rank (and the caller) sees the whole batch's summary.  The label
  images stay on the rank that produced them.
"""
import os

import numpy as np


def shard(n_items, rank, world):
    """Global indices owned by `rank`: i with i mod world == rank, in increasing order."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_items, world))


class ContextPool:
    def __init__(self, gseg, max_w, max_h, device=0, contexts=4):
        self.gseg = gseg
        self.segs = [gseg.Segmenter(max_w, max_h, device=device) for _ in range(max(1, contexts))]
        if len(self.segs) >= 4:  # many contexts in flight: size each grid for 2 blocks per SM so kernels overlap
            for s in self.segs:
                s.set_blocks_per_sm(int(os.environ.get("GSEG_POOL_BLOCKS_PER_SM", "2")))

    def close(self):
        for s in self.segs:
            s.close()
        self.segs = []

    def run(self, images, on_result, **params):
        """Segment every image of `images` (arrays / tensors, or `bytes` holding a JPEG file, which nvJPEG
        decodes on the context's stream: SURVEY.md s8f N2); `on_result(i, segmenter)` is called once image i is complete
        (read its labels there; `segmenter.labels(out=..., wait=False)` keeps the copy-out asynchronous: it
        is ordered before the context's next image and completed before run() returns)."""
        S = len(self.segs)
        n = len(images)
        pending = [-1] * S  # image index in flight on each context
        # rolling pipeline: a context gets its next image as soon as its previous one is complete, so S - 1
        # images stay in flight while the host waits for the oldest
        for i in range(n + S):
            j = i % S
            if pending[j] >= 0:
                self.segs[j].wait()
                on_result(pending[j], self.segs[j])
                pending[j] = -1
            if i < n:
                if isinstance(images[i], (bytes, bytearray, memoryview)):  # a JPEG file's bytes: decoded on the GPU
                    self.segs[j].segment_jpeg(bytes(images[i]), wait=False, **params)
                else:
                    self.segs[j].segment(images[i], wait=False, **params)
                pending[j] = i
        for s in self.segs:
            s.sync()
        return n


def segment_sharded(n_items, load_image, segment_one, dist=None):
    """Run `segment_one(image) -> int(num_components)` over this rank's shard of a global batch of
    `n_items` images (`load_image(i)` yields image i) and return the global list of component counts.
    `dist` is torch.distributed (any backend) or None for a single process."""
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None and dist.is_initialized() else (0, 1)
    mine = shard(n_items, rank, world)
    counts = np.full(n_items, -1, np.int64)
    for i in mine:
        counts[i] = int(segment_one(load_image(i)))
    if world > 1:
        import torch
        t = torch.from_numpy(counts.copy())
        # every index is owned by exactly one rank and unowned slots hold -1: MAX merges the shards
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        counts = t.numpy()
    return counts.tolist()
