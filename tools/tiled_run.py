"""Tiled schedule across GPUs (BASELINE.json configs[4]): one strip per rank, boundary exchange over NCCL,
final rounds on the joined graph.  Launch:  torchrun --nproc-per-node N tools/tiled_run.py W H [conn] [check]
check=1: rank 0 also runs the tiled CPU oracle on the whole image and compares partitions (moderate sizes)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

W, H = int(sys.argv[1]), int(sys.argv[2])
conn = int(sys.argv[3]) if len(sys.argv) > 3 else 4
check = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gseg = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200")
if rank == 0:
    gseg.build()
if world > 1:
    dist.barrier()
tiled = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200.tiled")
y0, y1 = tiled.strip_rows(H, world)[rank]
hs = y1 - y0
seg = gseg.Segmenter(W, hs, device=local, max_connectivity=conn)
# the strip of the synthetic image: generated whole on the device (rows y0..y1 of seed 5), so that every rank
# sees the same global image as the oracle
full = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda") if H * W <= (1 << 28) else None
if full is not None:
    gen = gseg.Segmenter(W, H, device=local, max_connectivity=4)
    gen.synth(W, H, 5, out=full)
    gen.close()
    strip = full[y0:y1].contiguous()
else:  # too large to hold whole on every rank: an independent strip image per rank (timing only)
    strip = torch.empty((hs, W, 3), dtype=torch.uint8, device="cuda")
    seg.synth(W, hs, 5 + rank, out=strip)
kw = dict(sigma=0.8, k=300.0, min_size=20, connectivity=conn, variant=0)
dlab = torch.empty((hs, W), dtype=torch.int32, device="cuda")
times = {}


def seg_strip(img):
    t0 = time.perf_counter()
    seg.segment(img, **kw)
    seg.labels(out=dlab)
    torch.cuda.synchronize()
    times["phase1"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    g = seg.export_graph()
    top, bot = seg.blurred_rows(0)[:, 0, :], seg.blurred_rows(hs - 1)[:, 0, :]
    lab_rows = torch.stack([dlab[0], dlab[-1]]).cpu().numpy()
    times["export"] = time.perf_counter() - t0
    return lab_rows, g, top, bot           # only the first and last label rows travel


def seg_graph(size, Int, ea, eb, w):
    t0 = time.perf_counter()
    out = seg.segment_graph(size, Int, ea, eb, w, k=300.0, min_size=20, variant=0)
    times["phase2"] = time.perf_counter() - t0
    return out


def run():
    t0 = time.perf_counter()
    lab_rows, graph, top, bot = seg_strip(strip)
    t1 = time.perf_counter()
    recs = tiled.exchange(tiled.strip_record(lab_rows, graph, top, bot), dist if world > 1 else None, device="cuda")
    times["exchange"] = time.perf_counter() - t1
    t1 = time.perf_counter()
    joined = tiled.join_strips(recs, conn)
    times["join"] = time.perf_counter() - t1
    comp, n = seg_graph(joined["size"], joined["Int"], joined["ea"], joined["eb"], joined["w"])
    t1 = time.perf_counter()
    F = torch.from_numpy(comp[int(joined["offsets"][rank]):int(joined["offsets"][rank + 1])].astype(np.int32)).cuda()
    final = F[dlab.long()]                 # the strip's label image in image-global ids, on the device
    torch.cuda.synchronize()
    times["relabel"] = time.perf_counter() - t1
    times["total"] = time.perf_counter() - t0
    return final, n, joined


for rep in range(3):
    if world > 1:
        dist.barrier()
    final, n, joined = run()
tt = torch.tensor([times[k] for k in ("phase1", "export", "exchange", "join", "phase2", "relabel", "total")], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    names = ("phase1", "export", "exchange", "join", "phase2", "relabel", "total")
    print("tiled %dx%d conn %d on %d GPU(s): %d strips of %d rows; joined graph %d components, %d edges -> %d final components" %
          (W, H, conn, world, world, hs, len(joined["size"]), len(joined["ea"]), n))
    print("  max over ranks, ms: " + "  ".join("%s %.1f" % (k, v * 1e3) for k, v in zip(names, tt.tolist())))
    print("  %.0f Mpixel/s whole job" % (W * H / 1e6 / tt[-1].item()), flush=True)
if check and full is not None:
    labs = [torch.empty((b - a, W), dtype=torch.int32, device="cuda") for (a, b) in tiled.strip_rows(H, world)]
    if world > 1:
        dist.all_gather(labs, final.int().contiguous()) if len(set(x.shape for x in labs)) == 1 else None
    else:
        labs = [final.int()]
    if rank == 0 and (world == 1 or len(set(x.shape for x in labs)) == 1):
        from oracle import oracle as O
        from tests.tiled_ref import oracle_tiled
        img = full.cpu().numpy()
        t0 = time.perf_counter()
        ref, nref, _, _ = oracle_tiled(O, img, world, 0.8, 300.0, 20, conn)
        got = torch.cat(labs).cpu().numpy()
        a, na = O.canon(got.reshape(H, W)); b, nb = O.canon(ref.reshape(H, W))
        print("  tiled CPU oracle: %.1f s; partition identical: %s (%d components)" %
              (time.perf_counter() - t0, bool(na == nb and np.array_equal(a, b)), nb), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
