import importlib, os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gseg = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200")
import torch
seg = gseg.Segmenter(1920, 1080)
dimg = torch.empty((1080, 1920, 3), dtype=torch.uint8, device="cuda")
seg.synth(1920, 1080, 2, out=dimg)
seg.set_tail(0, 0)
for _ in range(3):
    seg.segment(dimg, sigma=0.8, k=300.0, min_size=20, connectivity=4, variant=0)
arr = (C.c_ulonglong * 8)()
seg.L.gseg_debug_seg.argtypes = [C.c_void_p, C.c_void_p]
seg.L.gseg_debug_seg(seg.h, arr)
t = list(arr)[:6]
tot = sum(t)
names = ["ticket", "load+gather+count", "lookback", "stage", "emit", "-"]
print("k_edges round 1: warp-cycles per segment (sum over warps), total %.1f Mcycles" % (tot / 1e6))
for n, v in zip(names, t):
    print("  %-20s %8.2f Mcycles  %5.1f%%" % (n, v / 1e6, 100.0 * v / max(tot, 1)))
