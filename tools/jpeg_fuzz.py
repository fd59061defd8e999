"""Randomised check of the in-house JPEG decoder on the GPU against libjpeg (cv2.imdecode): sizes, sampling layouts,
qualities, restart intervals (0 = none: the self-synchronising path), optimised tables, grey-scale; single calls and
pool jobs; a 4K file (several sub-sequences per thread) and files larger than the default staging buffer (grow path).
Usage: python tools/jpeg_fuzz.py [ncases] [seed]"""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "graph-algorithm-image-segmentation-gpgpu_b200"
gseg = importlib.import_module(PKG)
batch = importlib.import_module(PKG + ".batch")
import cv2
import numpy as np

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
SAMP = [cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
        cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411]
kw = dict(sigma=0.8, k=300.0, min_size=20, connectivity=4, variant=0)


def picture(w, h, kind):
    if kind == 0:    # smooth + noise
        y, x = np.mgrid[0:h, 0:w]
        base = np.stack([(x * 3 + y) % 256, (x + y * 2) % 256, (x * y // 7) % 256], -1)
        return np.clip(base + rng.integers(-8, 9, (h, w, 3)), 0, 255).astype(np.uint8)
    if kind == 1:    # pure noise (stuffing-heavy at high quality)
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    return np.full((h, w, 3), int(rng.integers(0, 256)), np.uint8)  # flat


def encode(img, grey=False):
    q = int(rng.choice([20, 50, 75, 90, 95, 100]))
    rst = int(rng.choice([0, 0, 0, 1, 2, 3, 5, 8, 17, 64]))
    p = [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, int(rng.choice(SAMP)), cv2.IMWRITE_JPEG_RST_INTERVAL, rst,
         cv2.IMWRITE_JPEG_OPTIMIZE, int(rng.integers(0, 2))]
    src = img[..., 0] if grey else np.ascontiguousarray(img[..., ::-1])
    ok, e = cv2.imencode(".jpg", np.ascontiguousarray(src), p)
    assert ok
    return e, (q, rst, p[3], p[7], grey)


def libjpeg(e):
    return np.ascontiguousarray(cv2.imdecode(e, cv2.IMREAD_COLOR)[..., ::-1])


t0 = time.time()
bad = 0
seg = gseg.Segmenter(720, 720)
seg.set_jpeg_backend(gseg.JPEG_OWN)
for i in range(ncases):
    w, h = (int(rng.integers(1, 721)), int(rng.integers(1, 721))) if i % 5 else (int(rng.integers(1, 40)), int(rng.integers(1, 40)))
    e, info = encode(picture(w, h, int(rng.integers(0, 3))), grey=(i % 11 == 0))
    seg.segment_jpeg(e.tobytes(), **kw)
    if not np.array_equal(seg.input_rgb(), libjpeg(e)):
        bad += 1
        print("MISMATCH", w, h, info, flush=True)
seg.close()
print("single calls: %d cases, %d mismatches, %.1f s" % (ncases, bad, time.time() - t0), flush=True)

# corrupt files: mutated headers and data -- an error or an image, never a crash; the context must decode a good file afterwards
# (an out-of-bounds access on the device would poison it)
seg = gseg.Segmenter(256, 256)
seg.set_jpeg_backend(gseg.JPEG_OWN)
good, _ = encode(picture(200, 144, 0))
nerr = ncuda = 0
for i in range(ncases):
    w, h = int(rng.integers(8, 200)), int(rng.integers(8, 200))
    e, info = encode(picture(w, h, int(rng.integers(0, 2))))
    f = bytearray(e.tobytes())
    sos = f.find(b"\xff\xda")
    for m in range(int(rng.integers(1, 7))):
        lim = sos + 14 if (i & 1) else len(f)
        p = int(rng.integers(2, max(3, lim)))
        f[p] = int(rng.choice([int(rng.integers(0, 256)), 0xFF, 0x00, f[p] ^ (1 << int(rng.integers(0, 8)))]))
    if i % 9 == 0:
        f = f[:int(rng.integers(4, len(f)))]
    try:
        seg.segment_jpeg(bytes(f), **kw)
    except gseg.GsegError as ex:
        nerr += 1
        if "CUDA" in str(ex) and "no CUDA device" not in str(ex):
            ncuda += 1
            print("CUDA ERROR on a corrupt file:", ex, info, flush=True)
            break
seg.segment_jpeg(good.tobytes(), **kw)
ok = np.array_equal(seg.input_rgb(), libjpeg(good))
bad += ncuda + (0 if ok else 1)
seg.close()
print("corrupt files: %d cases, %d rejected or flagged, %d CUDA errors, context fine afterwards: %s" % (ncases, nerr, ncuda, ok), flush=True)

# pool jobs: mixed files, decode-ahead on the copy streams, device label output; the decoded pixels are checked through the
# partition's component count against a second run of the same file through a plain context
w, h = 480, 320
pool = batch.Pool(gseg, w, h, contexts=4, max_connectivity=4, caps=gseg.CAP_JPEG)
ref = gseg.Segmenter(w, h)
ref.set_jpeg_backend(gseg.JPEG_OWN)
files = [encode(picture(w, h, int(rng.integers(0, 2))))[0] for _ in range(48)]
outs = [np.zeros((h, w), np.int32) for _ in files]
jobs = pool.jobs([batch.Jpeg(np.frombuffer(f.tobytes(), np.uint8).copy()) for f in files], outs, elem_bytes=4, **kw)
for rep in range(3):
    res = pool.run(jobs)
    for i, f in enumerate(files):
        ref.segment_jpeg(f.tobytes(), **kw)
        if res[i].status != 0 or not np.array_equal(outs[i], ref.labels()):
            bad += 1
            print("POOL MISMATCH job", i, res[i].status, flush=True)
pool.close()
ref.close()
print("pool: 3 x %d jobs, total mismatches so far %d" % (len(files), bad), flush=True)

# large files: 4K without restart markers (about 19 000 sub-sequences on 8 192 threads), 1080p noise at quality 100 (a file
# larger than the default staging buffer: grow path), both sampling layouts
big = gseg.Segmenter(3840, 2160, max_connectivity=4)
big.set_jpeg_backend(gseg.JPEG_OWN)
for (w, h, kind, q, rst, sf) in [(3840, 2160, 0, 90, 0, SAMP[1]), (3840, 2160, 0, 90, 4, SAMP[0]), (1920, 1080, 1, 100, 0, SAMP[0]),
                                 (1920, 1080, 1, 100, 7, SAMP[1]), (3840, 2160, 1, 95, 0, SAMP[1])]:
    ok, e = cv2.imencode(".jpg", np.ascontiguousarray(picture(w, h, kind)[..., ::-1]),
                         [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf, cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
    t1 = time.time()
    big.segment_jpeg(e.tobytes(), **kw)
    dt = time.time() - t1
    same = np.array_equal(big.input_rgb(), libjpeg(e))
    bad += 0 if same else 1
    print("large: %dx%d q%d rst %d, %d bytes: pixels == libjpeg %s (%.1f ms with the segmentation)" % (w, h, q, rst, e.size, same, dt * 1e3), flush=True)
big.close()
print("TOTAL MISMATCHES %d" % bad)
sys.exit(1 if bad else 0)
