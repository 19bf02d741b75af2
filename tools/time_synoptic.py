"""Time one synoptic detection (BASELINE config 4: 3600x1440 Carrington map, k=3, h_bins=5000, T=25)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pychips_b200 import engine, synth
img = synth.synthetic_synoptic(1440, 3600, 0)
dev = engine.require_cuda()
t = torch.from_numpy(img).to(dev)
det = engine.Detector(engine.Geometry(1440, 3600, 0, 0, -1), 3, 5000, -5, 20, 4, dev)
for _ in range(3):
    det.detect(t)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record(); det.detect(t); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print("synoptic 3600x1440 k=3 detect ms", sorted(ts)[5], "n_peaks", det.result().n_peaks)
