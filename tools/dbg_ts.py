import importlib, os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gseg = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200")
import torch
seg = gseg.Segmenter(1920, 1080)
dimg = torch.empty((1080, 1920, 3), dtype=torch.uint8, device="cuda")
seg.synth(1920, 1080, 2, out=dimg)
for _ in range(3):
    seg.segment(dimg, sigma=0.8, k=300.0, min_size=20, connectivity=4, variant=0)
arr = (C.c_ulonglong * 24)()
seg.L.gseg_debug_ts.argtypes = [C.c_void_p, C.c_void_p]
seg.L.gseg_debug_ts(seg.h, arr)
t = list(arr)
print("phase E block 0 timestamps (us from first):", [round((x - t[0]) / 1e3, 2) for x in t[:8]])
print("emit iterations (store, min a, min b):", [round((x - t[0]) / 1e3, 2) for x in t[8:20]])
print("after phaseE %.2f, after fence %.2f, after sync %.2f" % tuple((t[i] - t[0]) / 1e3 for i in (16, 17, 18)))
print(seg.timeline()[-1])
