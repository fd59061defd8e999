"""Device timeline of one synthetic image (tail S/R/E split per round).  args: w h conn variant"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gseg = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200")
import torch
w, h, conn, variant = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (1920, 1080, 4, 0)))
seg = gseg.Segmenter(w, h)
dimg = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
seg.synth(w, h, 2, out=dimg)
for _ in range(3):
    seg.segment(dimg, sigma=0.8, k=300.0, min_size=20, connectivity=conn, variant=variant)
    seg.sync()
tl = seg.timeline()
prev = 0.0
for (r, tail, end, s, rr, e, pages) in tl:
    print("r%2d %s end %7.1f dur %6.1f  S %5.1f R %5.1f E %5.1f pages %d" % (r, "tail" if tail else "grid", end, end - prev, s, rr, e, pages))
    prev = end
print("components", seg.num_components())
