"""Time one C-ABI stage call in isolation (L2 flushed between repetitions)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pychips_b200 import engine, synth, _lib
n, k = 4096, 51
img, r = synth.synthetic_disk(n, 0)
dev = engine.require_cuda()
t = torch.from_numpy(img).to(dev)
c = n // 2
det = engine.Detector(engine.Geometry(n, n, c, c, r), k, 500, 0, 20, 4, dev)
det.detect(t)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts)) * 1e3
which = sys.argv[1] if len(sys.argv) > 1 else "hist"
if which == "hist":
    print("hist us", timeit(lambda: det.histogram()))
elif which == "thr":
    print("threshold_prob us", timeit(lambda: det.threshold_prob(0.8)))
elif which == "cleanup":
    print("cleanup us", timeit(lambda: det.cleanup(1e-3)))
