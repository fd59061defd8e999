#!/bin/bash
mkdir -p gpurun_out/jpgs
python - <<'P'
import importlib, sys, os
sys.path.insert(0, os.getcwd())
import cv2, numpy as np
from oracle import oracle as O
for i in range(32):
    img = O.synth(1920, 1080, 3000 + i)
    cv2.imwrite("gpurun_out/jpgs/%02d.jpg" % i, np.ascontiguousarray(img[..., ::-1]), [cv2.IMWRITE_JPEG_QUALITY, 90])
P
nvcc -O2 -o gpurun_out/nvjpeg_probe tools/micro/nvjpeg_probe.cu -lnvjpeg && gpurun_out/nvjpeg_probe gpurun_out/jpgs
rm -rf gpurun_out/jpgs gpurun_out/nvjpeg_probe
