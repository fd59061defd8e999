"""Tiled schedule for one image cut into horizontal strips, one strip per GPU (BASELINE.json configs[4];
north_star: "only the gigapixel single-image case is tiled, with cross-tile boundary edges exchanged
over NVLink via NCCL before the final Boruvka rounds").  Host-side logic only: no compute, no fallback.

Two drivers of the same schedule:
  segment_tiled_device  the product path: strip + halo rows on the GPU, strip record written by the engine into
                        device memory, all-gather of device buffers (NCCL over NVLink, no host bounce), join +
                        joined rounds + final relabel on the device (gseg_strip_record / gseg_join_segment);
  segment_tiled         the same steps with the join done in numpy on host arrays: the executable specification the
                        CPU tests run over gloo with the oracle standing in for the engine.

Semantics (restated by the oracle in tests/tiled_ref.py; DESIGN.md "Tiled schedule"):
  phase 1  every strip is segmented on its own (Boruvka-Felzenszwalb to completion) -- no communication.  The
           strip is blurred with ceil(4 sigma) halo rows of the neighbouring strips (device driver; the host
           driver's stand-in takes the strip's rows of the whole image's blur), so its blurred pixels and all edge
           weights, cut edges included, are those of the untiled image;
  exchange every rank contributes its final component graph (sizes, Int(C), live inter-component edges in
           list order) and the blurred colours + labels of its first and last row; all-gather over NCCL;
  phase 2  the strips' graphs are joined: components renumbered strip by strip; edge list = strip 0's
           edges, cut edges between strips 0|1, strip 1's edges, cut edges 1|2, ...; a cut edge's weight is
           the L2 colour distance of its two blurred end pixels; cut edges of one boundary are ordered
           S (x = 0..w-1), then SE, then NE (8-connected only).  The same rounds (predicate, then min-size)
           run on the joined graph; list position is the tie-break.
The graph is the untiled image's graph; the partition still differs from an untiled run (a strip's rounds run
to completion before they see the neighbour strips' components) and is exact with respect to the tiled oracle.
"""
import numpy as np


def strip_rows(h, n_strips):
    """Row ranges [(y0, y1)] of n_strips horizontal strips: equal heights, the last one takes the rest."""
    if n_strips < 1 or h < n_strips:
        raise ValueError("need 1 <= n_strips <= image height")
    hs = h // n_strips
    return [(i * hs, (i + 1) * hs if i < n_strips - 1 else h) for i in range(n_strips)]


def _l2(p, q):
    """||p - q||_2 of (3, n) float32 colours with the engine's operation order (every op rounded to fp32)."""
    d = (p - q).astype(np.float32)
    r2, g2, b2 = (d[0] * d[0]).astype(np.float32), (d[1] * d[1]).astype(np.float32), (d[2] * d[2]).astype(np.float32)
    s = (r2 + g2).astype(np.float32)
    s = (s + b2).astype(np.float32)
    return np.sqrt(s).astype(np.float32)


def cut_edges(bottom_lab, bottom_col, top_lab, top_col, conn):
    """Edges across one strip boundary.  bottom_*: last row of the upper strip, top_*: first row of the lower
    strip; labels already in joined numbering; colours (3, w) float32.  Returns (ea, eb, w)."""
    w = len(bottom_lab)
    ea = [bottom_lab]
    eb = [top_lab]
    ww = [_l2(bottom_col, top_col)]
    if conn == 8 and w > 1:
        ea.append(bottom_lab[:-1]); eb.append(top_lab[1:]); ww.append(_l2(bottom_col[:, :-1], top_col[:, 1:]))   # SE
        ea.append(top_lab[:-1]); eb.append(bottom_lab[1:]); ww.append(_l2(top_col[:, :-1], bottom_col[:, 1:]))   # NE
    return (np.concatenate(ea).astype(np.uint32), np.concatenate(eb).astype(np.uint32), np.concatenate(ww))


def join_strips(strips, conn):
    """strips: list (top to bottom) of dict(n, size, Int, ea, eb, w, top_lab, top_col, bot_lab, bot_col) with
    strip-local dense ids.  Returns dict(size, Int, ea, eb, w, offsets)."""
    offs = np.concatenate([[0], np.cumsum([s["n"] for s in strips])]).astype(np.int64)
    size = np.concatenate([np.asarray(s["size"], np.uint32) for s in strips])
    Int = np.concatenate([np.asarray(s["Int"], np.float32) for s in strips])
    ea, eb, w = [], [], []
    for i, s in enumerate(strips):
        ea.append(np.asarray(s["ea"], np.int64) + offs[i]); eb.append(np.asarray(s["eb"], np.int64) + offs[i])
        w.append(np.asarray(s["w"], np.float32))
        if i + 1 < len(strips):
            t = strips[i + 1]
            ca, cb, cw = cut_edges(np.asarray(s["bot_lab"], np.int64) + offs[i], np.asarray(s["bot_col"], np.float32),
                                   np.asarray(t["top_lab"], np.int64) + offs[i + 1], np.asarray(t["top_col"], np.float32), conn)
            ea.append(ca.astype(np.int64)); eb.append(cb.astype(np.int64)); w.append(cw)
    return dict(size=size, Int=Int, ea=np.concatenate(ea).astype(np.uint32), eb=np.concatenate(eb).astype(np.uint32),
                w=np.concatenate(w).astype(np.float32), offsets=offs)


def strip_record(labels, graph, top_col, bot_col):
    """What a rank contributes to the exchange, from its strip's dense label image and exported graph."""
    return dict(n=len(graph["size"]), size=graph["size"], Int=graph["Int"], ea=graph["ea"], eb=graph["eb"], w=graph["w"],
                top_lab=np.asarray(labels[0]).copy(), bot_lab=np.asarray(labels[-1]).copy(),
                top_col=np.asarray(top_col, np.float32).reshape(3, -1), bot_col=np.asarray(bot_col, np.float32).reshape(3, -1))


def _pack(rec):
    """One flat float64-free int64/float32 pair of buffers per record, for tensor all-gathers."""
    ints = np.concatenate([[rec["n"], len(rec["ea"]), len(rec["top_lab"])], rec["size"], rec["ea"], rec["eb"],
                           rec["top_lab"], rec["bot_lab"]]).astype(np.int64)
    flts = np.concatenate([rec["Int"], rec["w"], rec["top_col"].reshape(-1), rec["bot_col"].reshape(-1)]).astype(np.float32)
    return ints, flts


def _unpack(ints, flts):
    n, ne, w = int(ints[0]), int(ints[1]), int(ints[2])
    o = 3
    size = ints[o:o + n]; o += n
    ea = ints[o:o + ne]; o += ne
    eb = ints[o:o + ne]; o += ne
    top_lab = ints[o:o + w]; o += w
    bot_lab = ints[o:o + w]
    f = 0
    Int = flts[f:f + n]; f += n
    ww = flts[f:f + ne]; f += ne
    top_col = flts[f:f + 3 * w].reshape(3, w); f += 3 * w
    bot_col = flts[f:f + 3 * w].reshape(3, w)
    return dict(n=n, size=size.astype(np.uint32), Int=Int, ea=ea.astype(np.uint32), eb=eb.astype(np.uint32), w=ww,
                top_lab=top_lab, bot_lab=bot_lab, top_col=top_col, bot_col=bot_col)


def exchange(rec, dist, device="cpu"):
    """All-gather the strip records (rank order = strip order).  dist: torch.distributed or None (1 rank)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [rec]
    import torch
    world = dist.get_world_size()
    ints, flts = _pack(rec)
    lens = torch.tensor([len(ints), len(flts)], dtype=torch.int64, device=device)
    all_lens = [torch.zeros_like(lens) for _ in range(world)]
    dist.all_gather(all_lens, lens)
    mi, mf = int(max(l[0] for l in all_lens)), int(max(l[1] for l in all_lens))
    ti = torch.zeros(mi, dtype=torch.int64, device=device); ti[:len(ints)] = torch.from_numpy(ints).to(device)
    tf = torch.zeros(max(mf, 1), dtype=torch.float32, device=device); tf[:len(flts)] = torch.from_numpy(flts).to(device)
    gi = [torch.zeros_like(ti) for _ in range(world)]
    gf = [torch.zeros_like(tf) for _ in range(world)]
    dist.all_gather(gi, ti)      # the boundary-edge exchange: NCCL over NVLink on the GPU box, gloo in CPU tests
    dist.all_gather(gf, tf)
    return [_unpack(gi[r].cpu().numpy()[:int(all_lens[r][0])], gf[r].cpu().numpy()[:int(all_lens[r][1])]) for r in range(world)]


def halo_rows(sigma):
    """Rows of halo a strip needs on every side that has a neighbour: the blur's half width ceil(4 sigma)."""
    import math
    return int(math.ceil(max(float(np.float32(sigma)), 0.01) * 4.0))


def strip_with_halo(h, n_strips, i, sigma):
    """(y0, y1, halo_top, halo_bottom) of strip i: its rows [y0, y1) and the halo rows that exist around them."""
    y0, y1 = strip_rows(h, n_strips)[i]
    r = halo_rows(sigma)
    return y0, y1, min(r, y0), min(r, h - y1)


class DeviceTiler:
    """The tiled schedule on the device for one rank's strip (see module docstring).  Buffers are kept across calls."""

    def __init__(self, seg, dist=None):
        import torch
        self.seg, self.torch = seg, torch
        self.dist = dist if dist is not None and dist.is_initialized() and dist.get_world_size() > 1 else None
        self.rank = self.dist.get_rank() if self.dist else 0
        self.world = self.dist.get_world_size() if self.dist else 1
        self.send = self.recv = None
        self.sizes = torch.zeros(1, dtype=torch.int64, device="cuda")
        self.times = {}
        # one stream for the engine's kernels and the collective, so that they are ordered without host round trips
        # (torch's default stream has handle 0, which gseg_set_stream reads as "the context's own stream")
        self.stream = torch.cuda.Stream()
        seg.set_stream(self.stream.cuda_stream)

    def run(self, buf, halo_top, halo_bottom, out=None, **params):
        """buf: (halo_top + hs + halo_bottom, w, 3) uint8 CUDA tensor; out: (hs, w) CUDA tensor (int32 / int16-as-
        uint16 / uint8) or pinned host array for the strip's final labels.  Returns (n_final, n_joined, e_joined)."""
        with self.torch.cuda.stream(self.stream):
            return self._run(buf, halo_top, halo_bottom, out, params)

    def _run(self, buf, halo_top, halo_bottom, out, params):
        import time
        torch, seg, T = self.torch, self.seg, self.times
        t0 = time.perf_counter()
        seg.segment_strip(buf, halo_top, halo_bottom, **params)
        T["phase1"] = time.perf_counter() - t0
        t1 = time.perf_counter()
        nbytes = seg.strip_record_bytes()               # export + duplicate elimination happen here
        T["export"] = time.perf_counter() - t1
        t1 = time.perf_counter()
        stride = nbytes
        if self.dist:
            self.sizes[0] = nbytes
            self.dist.all_reduce(self.sizes, op=self.dist.ReduceOp.MAX)
            stride = int(self.sizes.item())
        stride = (stride + 255) & ~255
        if self.send is None or self.send.numel() < stride:
            self.send = torch.empty(stride * 2, dtype=torch.uint8, device="cuda")
            self.recv = torch.empty(stride * 2 * self.world, dtype=torch.uint8, device="cuda")
        send = self.send[:stride]
        seg.strip_record(send.data_ptr(), stride)
        if self.dist:
            recv = self.recv[:stride * self.world]
            self.dist.all_gather_into_tensor(recv, send)   # the boundary exchange: device buffers over NVLink
            torch.cuda.current_stream().synchronize()
        else:
            recv = send
        T["exchange"] = time.perf_counter() - t1
        T["exchange_bytes"] = stride * self.world
        t1 = time.perf_counter()
        p = {k: v for k, v in params.items() if k in ("k", "min_size", "connectivity", "variant", "max_rounds", "max_levels", "flags")}
        res = seg.join_segment(recv.data_ptr(), self.world, stride, self.rank, out=out, sigma=params.get("sigma", 0.8), **p)
        T["join_phase2"] = time.perf_counter() - t1
        T["total"] = time.perf_counter() - t0
        return res


def segment_tiled(strip_img, segment_strip, segment_graph, conn, dist=None, device="cpu"):
    """Run the tiled schedule for this rank's strip.
    segment_strip(img) -> (dense labels (hs, w), graph dict(size, Int, ea, eb, w), top_col (3, w), bot_col (3, w))
    segment_graph(size, Int, ea, eb, w) -> (label per joined component, n_final)
    Returns (final label image of the strip in image-global dense ids, n_final)."""
    labels, graph, top_col, bot_col = segment_strip(strip_img)
    recs = exchange(strip_record(labels, graph, top_col, bot_col), dist, device)
    joined = join_strips(recs, conn)
    comp_label, n_final = segment_graph(joined["size"], joined["Int"], joined["ea"], joined["eb"], joined["w"])
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    off = int(joined["offsets"][rank])
    return np.asarray(comp_label)[off + np.asarray(labels, np.int64)], n_final
