"""Tiled schedule, host logic on CPU: strip geometry, cut-edge weights under the fp32 contract, the record
(un)packing of the exchange, and a world-size-2 gloo run of segment_tiled() with the oracle standing in
for the GPU engine, against the single-process tiled oracle."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "graph-algorithm-image-segmentation-gpgpu_b200"


def test_strip_rows():
    t = importlib.import_module(PKG + ".tiled")
    assert t.strip_rows(10, 3) == [(0, 3), (3, 6), (6, 10)]
    assert t.strip_rows(8, 1) == [(0, 8)]
    with pytest.raises(ValueError):
        t.strip_rows(2, 3)


def test_cut_edge_weights_follow_the_fp32_contract(oracle):
    """A cut edge's weight is computed on the host side of the exchange: it must be the very fp32 value the
    engine / oracle would give the same pixel pair."""
    t = importlib.import_module(PKG + ".tiled")
    img = oracle.synth(97, 41, 5)
    planes = oracle.blur(img, 0.8)
    h, w = 41, 97
    wts, _ = oracle.edges(planes, 8)
    V = h * w
    y = 17
    lab = np.arange(w)
    ea, eb, ww = t.cut_edges(lab, planes[:, y, :], lab + 1000, planes[:, y + 1, :], 8)
    S = wts[1 * V + y * w: 1 * V + y * w + w]
    SE = wts[2 * V + y * w: 2 * V + y * w + w - 1]
    NE = wts[3 * V + (y + 1) * w: 3 * V + (y + 1) * w + w - 1]
    assert np.array_equal(ww.view(np.uint32), np.concatenate([S, SE, NE]).view(np.uint32))
    assert np.array_equal(ea[:w], lab) and np.array_equal(eb[:w], lab + 1000)
    assert np.array_equal(ea[w:2 * w - 1], lab[:-1]) and np.array_equal(eb[w:2 * w - 1], lab[1:] + 1000)       # SE
    assert np.array_equal(ea[2 * w - 1:], lab[:-1] + 1000) and np.array_equal(eb[2 * w - 1:], lab[1:])          # NE
    ea4, eb4, w4 = t.cut_edges(lab, planes[:, y, :], lab + 1000, planes[:, y + 1, :], 4)
    assert len(ea4) == w and np.array_equal(w4.view(np.uint32), S.view(np.uint32))


def test_record_pack_roundtrip(oracle):
    t = importlib.import_module(PKG + ".tiled")
    from tests.tiled_ref import oracle_strip
    lab, graph, top, bot = oracle_strip(oracle, oracle.synth(60, 30, 2), 0.8, 300.0, 20, 8)
    rec = t.strip_record(lab, graph, top, bot)
    back = t._unpack(*t._pack(rec))
    for key in rec:
        a, b = np.asarray(rec[key]), np.asarray(back[key])
        assert a.shape == b.shape or key == "n"
        assert np.array_equal(a.astype(np.float64), b.astype(np.float64)), key


def test_halo_geometry_and_blur_equivalence(oracle):
    """A strip blurred with ceil(4 sigma) halo rows has exactly the untiled image's blurred pixels."""
    t = importlib.import_module(PKG + ".tiled")
    assert t.halo_rows(0.8) == 4 and t.halo_rows(0.5) == 2 and t.halo_rows(2.0) == 8
    assert t.strip_with_halo(100, 4, 0, 0.8) == (0, 25, 0, 4)
    assert t.strip_with_halo(100, 4, 3, 0.8) == (75, 100, 4, 0)
    assert t.strip_with_halo(6, 3, 1, 0.8) == (2, 4, 2, 2)   # halo clipped to the rows that exist
    img = oracle.synth(64, 90, 3)
    whole = oracle.blur(img, 0.8)
    for i in range(3):
        y0, y1, ht, hb = t.strip_with_halo(90, 3, i, 0.8)
        part = oracle.blur(np.ascontiguousarray(img[y0 - ht:y1 + hb]), 0.8)[:, ht:ht + (y1 - y0), :]
        assert np.array_equal(part.view(np.uint32), whole[:, y0:y1, :].view(np.uint32))


def test_tiled_oracle_sanity(oracle):
    from tests.tiled_ref import oracle_tiled
    img = oracle.synth(200, 160, 9)
    for S in (1, 2, 4):
        out, n, joined, stats = oracle_tiled(oracle, img, S, 0.8, 300.0, 20, 4)
        assert out.shape == (160 * 200,) or out.shape == (160, 200)
        cnt = np.bincount(oracle.canon(out.reshape(160, 200))[0].reshape(-1))
        assert cnt.sum() == 160 * 200 and cnt.min() >= 20 and len(cnt) == n


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t = importlib.import_module(PKG + ".tiled")
    from oracle import oracle as O
    from tests.tiled_ref import oracle_strip
    img = O.synth(150, 120, 31)
    y0, y1, ht, hb = t.strip_with_halo(120, world, rank, 0.8)

    def seg_strip(buf):
        return oracle_strip(O, buf, 0.8, 300.0, 20, 8, halo_top=ht, halo_bottom=hb)

    def seg_graph(size, Int, ea, eb, w):
        lab, n, _ = O.boruvka_graph(size, Int, ea, eb, w, O.FELZ, 300.0, 20, 48)
        return lab, n

    lab, n = t.segment_tiled(np.ascontiguousarray(img[y0 - ht:y1 + hb]), seg_strip, seg_graph, 8, dist)
    q.put((rank, lab, n))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_tiled_matches_single_process(oracle):
    import torch.multiprocessing as mp
    from tests.tiled_ref import oracle_tiled
    img = oracle.synth(150, 120, 31)
    ref, nref, _, _ = oracle_tiled(oracle, img, 2, 0.8, 300.0, 20, 8)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=180) for _ in ps], key=lambda x: x[0])
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    full = np.concatenate([res[0][1], res[1][1]])
    a, na = oracle.canon(full.reshape(120, 150))
    b, nb = oracle.canon(ref.reshape(120, 150))
    assert na == nb == nref == res[0][2] == res[1][2]
    assert np.array_equal(a, b)
