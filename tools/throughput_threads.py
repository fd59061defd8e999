"""Is the 8-context pipeline bound by the host thread's launch rate?  T host threads, each with its own C++ pool
of S contexts, segment B/T device-resident 1080p images each; whole-GPU Mpixel/s for several (T, S).
Usage: python tools/throughput_threads.py [B]"""
import importlib, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
gseg = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200")
gseg.build()
batch = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200.batch")
W, H = 1920, 1080
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kw = dict(sigma=0.8, k=300.0, min_size=20, connectivity=4, variant=0)
gen = gseg.Segmenter(W, H, max_connectivity=4)
dimgs = torch.empty((B, H, W, 3), dtype=torch.uint8, device="cuda")
for i in range(B):
    gen.synth(W, H, 2000 + i, out=dimgs[i])
gen.close()
for T, S in [(1, 8), (2, 4), (2, 8), (4, 2), (4, 4), (1, 16), (1, 4)]:
    pools = [batch.Pool(gseg, W, H, contexts=S, max_connectivity=4) for _ in range(T)]
    dl = torch.empty((T, S, H, W), dtype=torch.int32, device="cuda")
    per = B // T
    jobs = [pools[t].jobs([dimgs[t * per + i] for i in range(per)], [dl[t, i % S] for i in range(per)], **kw) for t in range(T)]

    def work(t):
        pools[t].run(jobs[t])

    def step():
        ts = [threading.Thread(target=work, args=(t,)) for t in range(T)]
        [x.start() for x in ts]
        [x.join() for x in ts]

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    l = sum(p.launch_count() for p in pools)
    print("threads %d x contexts %d: %.2f ms per %d images -> %.0f Mpixel/s  (%.1f us/image)" % (T, S, dt * 1e3, per * T, per * T * W * H / 1e6 / dt, dt * 1e6 / (per * T)), flush=True)
    for p in pools:
        p.close()
