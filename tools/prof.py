"""Per-kernel timing table (CUDA events inside libgseg's host-driven schedule) + whole-run timings.
Usage: python tools/prof.py [w h conn variant] ; writes gpurun_out/prof_<tag>.txt"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gseg = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200")
if not os.environ.get("GSEG_NOBUILD"):
    gseg.build()
import torch

w, h, conn, variant = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (1920, 1080, 4, 0)))
tag = "%dx%d_c%d_v%d" % (w, h, conn, variant)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "prof_%s.txt" % tag), "w")


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s)
    out.write(s + "\n")


seg = gseg.Segmenter(w, h)
dimg = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
seg.synth(w, h, 2, out=dimg)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
kw = dict(sigma=0.8, k=300.0, min_size=20, connectivity=conn, variant=variant)

for flags, name in [(1, "host-driven"), (0, "persistent")]:
    for _ in range(3):
        seg.segment(dimg, flags=flags, **kw)
    ts = []
    for _ in range(10):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        seg.segment(dimg, flags=flags, **kw)
        ts.append((time.perf_counter() - t0) * 1e3)
    P("%-12s wall ms/image (L2 flushed): min %.3f med %.3f  -> %.1f Mpixel/s (med)" %
      (name, min(ts), sorted(ts)[len(ts) // 2], w * h / 1e3 / sorted(ts)[len(ts) // 2]))
P("rounds:", seg.stats())
P("device timeline of the device-driven run (us since round-0 graph kernel start):")
prev = 0.0
for (r, tail, end, us, ur, ue, npg) in seg.timeline():
    P("  round %2d %s end %8.1f  dur %7.1f   S %6.1f R %6.1f E %6.1f   pages %6d" % (r, "tail" if tail else "grid", end, end - prev, us, ur, ue, npg))
    prev = end
P("components:", seg.num_components(), "levels:", seg.num_levels(), "launches so far:", seg.launch_count())

seg.set_profiling(True)
agg = {}
NREP = 5
for rep in range(NREP):
    flush.zero_()
    torch.cuda.synchronize()
    seg.segment(dimg, flags=1, **kw)
    for name, rnd, ms, by in seg.profile():
        key = (name, rnd)
        a = agg.setdefault(key, [0.0, by])
        a[0] += ms / NREP
tot = sum(v[0] for v in agg.values())
P("\nper-kernel (mean of %d runs, L2 flushed before each run); total kernel time %.3f ms" % (NREP, tot))
P("%-12s %5s %10s %8s %12s %10s" % ("kernel", "round", "us", "share", "algo MB", "GB/s"))
for (name, rnd), (ms, by) in sorted(agg.items(), key=lambda kv: (kv[0][1], -kv[1][0])):
    if ms * 1e3 < 1.0 and rnd > 3:
        continue
    P("%-12s %5d %10.1f %7.1f%% %12.2f %10.1f" % (name, rnd, ms * 1e3, 100 * ms / tot, by / 1e6, by / ms / 1e6 if ms > 0 else 0))
byname = {}
for (name, rnd), (ms, by) in agg.items():
    b = byname.setdefault(name, [0.0, 0.0, 0])
    b[0] += ms; b[1] += by; b[2] += 1
P("\nby kernel name (all rounds)")
for name, (ms, by, n) in sorted(byname.items(), key=lambda kv: -kv[1][0]):
    P("%-12s n=%3d %10.1f us %6.1f%% %10.2f MB %9.1f GB/s" % (name, n, ms * 1e3, 100 * ms / tot, by / 1e6, by / ms / 1e6 if ms > 0 else 0))
out.close()
