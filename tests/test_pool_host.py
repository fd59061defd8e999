"""Host logic of batch.ContextPool on a stand-in Segmenter (no GPU, no compute): the rolling pipeline
issues every image exactly once, never has more than one image in flight per context, reports results
in image order per context, routes JPEG bytes to segment_jpeg and sizes the grids for many contexts."""
import importlib

import numpy as np

batch = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200.batch")


class FakeSegmenter:
    log = []

    def __init__(self, max_w, max_h, device=0):
        self.in_flight = None
        self.blocks = None
        self.closed = False

    def set_blocks_per_sm(self, b):
        self.blocks = b

    def segment(self, img, wait=True, **kw):
        assert self.in_flight is None and not wait
        self.in_flight = ("array", int(img[0, 0, 0]), kw)
        FakeSegmenter.log.append(("issue", id(self), self.in_flight[1]))

    def segment_jpeg(self, data, wait=True, **kw):
        assert self.in_flight is None and not wait and isinstance(data, bytes)
        self.in_flight = ("jpeg", data[0], kw)
        FakeSegmenter.log.append(("issue", id(self), self.in_flight[1]))

    def wait(self):
        assert self.in_flight is not None
        self.done, self.in_flight = self.in_flight, None

    def sync(self):
        assert self.in_flight is None

    def close(self):
        self.closed = True


class FakeModule:
    Segmenter = FakeSegmenter


def test_rolling_pipeline_on_stand_in(monkeypatch):
    monkeypatch.delenv("GSEG_POOL_BLOCKS_PER_SM", raising=False)
    for contexts, n in ((1, 5), (3, 10), (4, 4), (8, 3)):
        FakeSegmenter.log = []
        pool = batch.ContextPool(FakeModule, 64, 64, contexts=contexts)
        assert len(pool.segs) == contexts
        assert all(s.blocks == (2 if contexts >= 4 else None) for s in pool.segs)
        items = []
        for i in range(n):
            if i % 2:
                items.append(bytes([i]) + b"jpeg")
            else:
                items.append(np.full((2, 2, 3), i, np.uint8))
        got = []
        assert pool.run(items, lambda i, s: got.append((i, s.done[0], s.done[1], s.done[2]["k"])), k=7.0) == n
        assert sorted(g[0] for g in got) == list(range(n))                       # every image exactly once
        assert all(kind == ("jpeg" if i % 2 else "array") and tag == i and k == 7.0 for i, kind, tag, k in got)
        assert [t for (_, _, t) in FakeSegmenter.log] == list(range(n))          # issued in order
        per_ctx = {}
        for i, *_ in got:                                                        # image i ran on context i mod S ...
            per_ctx.setdefault(i % contexts, []).append(i)
        assert all(v == sorted(v) for v in per_ctx.values())                     # ... and finished there in order
        segs = list(pool.segs)
        pool.close()
        assert all(s.closed for s in segs) and pool.segs == []


def test_pool_blocks_per_sm_override(monkeypatch):
    monkeypatch.setenv("GSEG_POOL_BLOCKS_PER_SM", "3")
    pool = batch.ContextPool(FakeModule, 64, 64, contexts=4)
    assert all(s.blocks == 3 for s in pool.segs)
    pool.close()
