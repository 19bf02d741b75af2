"""Throughput cost of each stage when S detections are in flight (what the bench's `value` is made of).

For every subset of the detection's stages: capture the subset in one CUDA graph per stream slot,
replay it `--reps` times on S streams, report ms per image (wall = device time of the whole batch /
images).  The cost of the full detection is less than the sum of its stages run alone (kernels of
different streams share the SMs); the per-subset numbers show which stage occupies the GPU.

    python tools/stage_share.py [--streams 16] [--reps 6] [--k 51] [--n 4096]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pychips_b200 import _lib, engine, synth

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=16)
ap.add_argument("--reps", type=int, default=6)
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--k", type=int, default=51)
ap.add_argument("--elem", type=int, default=4)
ap.add_argument("--subsets", default="")
a = ap.parse_args()

dev = engine.require_cuda()
lib = _lib.load()
n, k, S = a.n, a.k, a.streams
c = n // 2
imgs, dets, maps, streams = [], [], [], []
for s in range(S):
    img, r = synth.synthetic_disk(n, s)
    if a.elem == 8:
        img = img.astype(np.float64) * (1.0 + 1e-12)
    imgs.append(torch.from_numpy(img).to(dev))
    dets.append(engine.Detector(engine.Geometry(n, n, c, c, r), k, 500, 0, 20, a.elem, dev))
    maps.append(torch.empty((20, n, n), dtype=torch.uint8, device=dev))
    streams.append(torch.cuda.Stream(device=dev))
    dets[-1].detect(imgs[-1], maps=maps[-1])
torch.cuda.synchronize()
P, SP = engine._ptr, engine._stream_ptr


def med(mask):
    def f(s):
        d = dets[s]
        _lib.check(lib.chips_medfilt2d_stages(P(imgs[s]), P(d.filt), d._pp, *d._g(), P(d.ws), SP(), mask), "med")
    return f


STAGES = {
    "sort": med(1), "huang": med(4), "select": med(8), "median": med(15),
    "hist": lambda s: (dets[s].histogram(), dets[s].select_threshold(5, None)),
    "thr": lambda s: dets[s].threshold_prob(0.8),
    "cleanup": lambda s: dets[s].cleanup(1e-3),
    "expand": lambda s: dets[s].expand_maps(0, maps[s]),
    "all": lambda s: dets[s].detect(imgs[s], maps=maps[s]),
}
SUBSETS = [["sort"], ["huang"], ["select"], ["median"], ["hist", "thr"], ["cleanup"], ["expand"],
           ["median", "hist", "thr"], ["hist", "thr", "cleanup", "expand"], ["all"]]
if a.subsets:
    SUBSETS = [x.split("+") for x in a.subsets.split(",")]

out = {}
for sub in SUBSETS:
    graphs = []
    for s in range(S):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=streams[s], capture_error_mode="thread_local"):
            for name in sub:
                STAGES[name](s)
        graphs.append(g)
    torch.cuda.synchronize()

    def run():
        cur = torch.cuda.current_stream()
        for st in streams:
            st.wait_stream(cur)
        for _ in range(a.reps):
            for s in range(S):
                with torch.cuda.stream(streams[s]):
                    graphs[s].replay()
        for st in streams:
            cur.wait_stream(st)

    run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / (a.reps * S))
    out["+".join(sub)] = round(sorted(ts)[1], 4)
    print("+".join(sub), "ms/image", out["+".join(sub)], flush=True)
    del graphs
print(json.dumps({"streams": S, "n": n, "k": k, "elem": a.elem, "ms_per_image": out}))
