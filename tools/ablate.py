import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gseg = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200")
import torch
seg = gseg.Segmenter(1920, 1080)
dimg = torch.empty((1080, 1920, 3), dtype=torch.uint8, device="cuda")
seg.synth(1920, 1080, 2, out=dimg)
seg.set_profiling(True)
agg = {}
N = 6
for i in range(N):
    seg.segment(dimg, sigma=0.8, k=300.0, min_size=20, connectivity=4, variant=0, flags=1, max_rounds=1)
    if i == 0:
        continue
    for name, rnd, ms, by in seg.profile():
        agg[name] = agg.get(name, 0.0) + ms * 1e3 / (N - 1)
print("dbg_flags=%s" % os.environ.get("GSEG_DBG_FLAGS", "0"), {k: round(v, 1) for k, v in agg.items()})
