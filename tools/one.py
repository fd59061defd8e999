"""Run a few segmentations of one synthetic image (for ncu captures).  args: w h conn variant flags n"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gseg = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200")
import torch
w, h, conn, variant, flags, n = (int(x) for x in (sys.argv[1:7] if len(sys.argv) >= 7 else (1920, 1080, 4, 0, 1, 2)))
seg = gseg.Segmenter(w, h)
dimg = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
seg.synth(w, h, 2, out=dimg)
out = torch.empty((h, w), dtype=torch.int32, device="cuda")
for _ in range(n):
    seg.segment(dimg, sigma=0.8, k=300.0, min_size=20, connectivity=conn, variant=variant, flags=flags)
    seg.labels(out=out)
torch.cuda.synchronize()
print("ok", seg.num_components(), len(seg.stats()))
