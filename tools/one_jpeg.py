"""A few JPEG-fed segmentations of one synthetic image (for ncu captures of the k_jpeg_* kernels).  args: w h rst sampling(420|444) n"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
gseg = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200")
import cv2
import numpy as np
w, h, rst, samp, n = (int(x) for x in (sys.argv[1:6] if len(sys.argv) >= 6 else (1920, 1080, 8, 420, 3)))
seg = gseg.Segmenter(w, h)
seg.set_jpeg_backend(gseg.JPEG_OWN)
img = seg.synth(w, h, 2)
sf = cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420 if samp == 420 else cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444
ok, e = cv2.imencode(".jpg", np.ascontiguousarray(img[..., ::-1]), [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf, cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
data = e.tobytes()
for _ in range(n):
    seg.segment_jpeg(data, sigma=0.8, k=300.0, min_size=20, connectivity=4, variant=0)
print("ok", len(data), seg.num_components(), np.array_equal(seg.input_rgb(), cv2.imdecode(e, cv2.IMREAD_COLOR)[..., ::-1]))
