"""Host time to enqueue one detection (is the series path CPU-launch-bound?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pychips_b200 import engine, synth
n = 4096
img, r = synth.synthetic_disk(n, 0)
dev = engine.require_cuda()
t = torch.from_numpy(img).to(dev)
det = engine.Detector(engine.Geometry(n, n, n // 2, n // 2, r), 51, 500, 0, 20, 4, dev)
maps = torch.empty((20, n, n), dtype=torch.uint8, device=dev)
for _ in range(3): det.detect(t, maps=maps)
torch.cuda.synchronize()
N = 20
t0 = time.perf_counter()
for _ in range(N): det.detect(t, maps=maps)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"enqueue {1e3*(t1-t0)/N:.3f} ms per detection (host), total {1e3*(t2-t0)/N:.3f} ms per detection")

det.detect_graphed(t, maps=maps); det.detect_graphed(t, maps=maps)
torch.cuda.synchronize()
ref = det.keep.clone(); rec = det.result_bytes().clone()
t0 = time.perf_counter()
for _ in range(N): det.detect_graphed(t, maps=maps)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"graph replay: enqueue {1e3*(t1-t0)/N:.3f} ms, total {1e3*(t2-t0)/N:.3f} ms per detection; "
      f"same result: {bool(torch.equal(ref, det.keep)) and bool(torch.equal(rec, det.result_bytes()))}")
