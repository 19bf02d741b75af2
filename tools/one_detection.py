"""One warmed-up 4096^2 detection (+ expand, + filament candidates): the target of ncu captures."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pychips_b200 import _lib, engine, synth
n, k = int(os.environ.get("N", 4096)), int(os.environ.get("K", 51))
img, r = synth.synthetic_disk(n, 0)
q = float(os.environ.get("QUANT", 0))
if q > 0:
    img = (np.round(img / q) * q).astype(np.float32)      # instrument-like quantised intensities
dev = engine.require_cuda()
elem = 4
if os.environ.get("F64"):
    img = img.astype(np.float64) * (1.0 + 1e-12)          # not float32-representable: float64 storage path
    elem = 8
t = torch.from_numpy(img).to(dev)
c = n // 2
det = engine.Detector(engine.Geometry(n, n, c, c, r), k, 500, 0, 20, elem, dev)
maps = torch.empty((20, n, n), dtype=torch.uint8, device=dev)
for _ in range(int(os.environ.get("REPS", 1))):
    det.detect(t, maps=maps)
torch.cuda.synchronize()
print("n_peaks", det.result().n_peaks, "median_errors", det.result().median_errors)
if os.environ.get("TIME"):
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); det.medfilt(t); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print("medfilt ms", sorted(ts)[2])
if os.environ.get("TIME"):
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); det.detect(t, maps=maps); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print("detect ms", sorted(ts)[2])
