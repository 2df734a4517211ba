"""Is one batch split into two halves on two streams faster than the whole batch on one stream?  (VERDICT r01 next#4:
"make two-batches-in-flight the library default for F >= 512".)  Sustained clocks: ~1 s of identical load before each
timed region.  Run on the GPU box:  python tools/split_probe.py [F]"""
import sys
import time

import torch

sys.path.insert(0, ".")
import facerecognition_infrenceengine_b200 as frg
from oracle import synth

n, d, k = 1_000_000, 512, 5
F = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
store = frg.GalleryStore(dim=d, capacity=n)
store.fill_synthetic(n, 0, synth.GALLERY_SEED)
m = frg.Matcher(store)
Q = torch.from_numpy(synth.queries(F, n, d)[0]).cuda()
s0 = torch.cuda.current_stream()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h = F // 2
outs = [None, None, None]


def whole():
    outs[0] = m.match_device(Q, k, 0.45, out=outs[0])


def split():
    ev = torch.cuda.Event()
    ev.record(s0)
    s1.wait_event(ev); s2.wait_event(ev)
    outs[1] = m.match_device(Q[:h], k, 0.45, out=outs[1], stream=s1.cuda_stream)
    outs[2] = m.match_device(Q[h:], k, 0.45, out=outs[2], stream=s2.cuda_stream)
    e1, e2 = torch.cuda.Event(), torch.cuda.Event()
    e1.record(s1); e2.record(s2)
    s0.wait_event(e1); s0.wait_event(e2)


def timed(fn, steps=200, preload=1.0):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    i = 0
    while time.perf_counter() - t < preload:
        fn(); i += 1
        if i % 32 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s0)
    for _ in range(steps):
        fn()
    b.record(s0)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


for rep in range(2):
    w = timed(whole)
    s = timed(split)
    print("F=%d rep %d: whole %.4f ms/step (%.0f q/s)   split 2x%d on two streams %.4f ms/step (%.0f q/s)   ratio %.3f"
          % (F, rep, w, F / w * 1e3, h, s, F / s * 1e3, w / s), flush=True)
store.close()
