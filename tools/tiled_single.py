"""The device path of the tiled schedule with every strip on its own context of ONE GPU (no NCCL: the 'gather' is a copy into
one buffer), for sizes where a multi-GPU box is not needed to hold the strips -- and its check against the tiled CPU oracle.
Usage: python tools/tiled_single.py W H n_strips [conn] [check]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
gseg = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200")
tiled = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200.tiled")
W, H, S = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
conn = int(sys.argv[4]) if len(sys.argv) > 4 else 4
check = int(sys.argv[5]) if len(sys.argv) > 5 else 0
SIGMA, K, MS, SEED = 0.8, 300.0, 20, 5
kw = dict(sigma=SIGMA, k=K, min_size=MS, connectivity=conn, variant=gseg.FELZ)
segs, bufs, geo = [], [], []
for i in range(S):
    y0, y1, ht, hb = tiled.strip_with_halo(H, S, i, SIGMA)
    s = gseg.Segmenter(W, y1 - y0, max_connectivity=conn)
    b = torch.empty((ht + (y1 - y0) + hb, W, 3), dtype=torch.uint8, device="cuda")
    s.synth_rows(W, y0 - ht, b.shape[0], SEED, out=b)
    segs.append(s); bufs.append(b); geo.append((y0, y1, ht, hb))
torch.cuda.synchronize()
t0 = time.perf_counter()
for s, b, (y0, y1, ht, hb) in zip(segs, bufs, geo):
    s.segment_strip(b, ht, hb, wait=False, **kw)
for s in segs:
    s.wait()
t1 = time.perf_counter()
sizes = [s.strip_record_bytes() for s in segs]
stride = (max(sizes) + 255) & ~255
recv = torch.zeros(S * stride, dtype=torch.uint8, device="cuda")
for i, s in enumerate(segs):
    s.strip_record(recv[i * stride:].data_ptr(), stride)
outs = [torch.empty((g[1] - g[0], W), dtype=torch.int32, device="cuda") for g in geo]
res = [s.join_segment(recv.data_ptr(), S, stride, i, out=outs[i], **kw) for i, s in enumerate(segs)]
torch.cuda.synchronize()
t2 = time.perf_counter()
print("tiled %dx%d conn %d, %d strips on one GPU: phase 1 (strips side by side) %.1f ms, records + join + joined rounds + relabel (all strips) %.1f ms; "
      "joined graph %d components, %d edges -> %d final components" % (W, H, conn, S, (t1 - t0) * 1e3, (t2 - t1) * 1e3, res[0][1], res[0][2], res[0][0]), flush=True)
assert len(set(res)) == 1
if check:
    from oracle import oracle as O
    from tests.tiled_ref import oracle_tiled
    full = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
    segs[0].synth_rows(W, 0, H, SEED, out=full)
    img = full.cpu().numpy()
    del full
    got = torch.cat(outs).cpu().numpy()
    for s in segs:
        s.close()
    del bufs, outs, recv
    torch.cuda.empty_cache()
    t0 = time.perf_counter()
    ref, nref, _, _ = oracle_tiled(O, img, S, SIGMA, K, MS, conn)
    a, na = O.canon(got.reshape(H, W)); b, nb = O.canon(ref.reshape(H, W))
    print("tiled CPU oracle: %.1f s; partition identical: %s (%d components)" % (time.perf_counter() - t0, bool(na == nb and np.array_equal(a, b)), nb), flush=True)
