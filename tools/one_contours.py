"""One detection + contour passes of thresholds 0 and 19 (for an ncu launch list)."""
import sys
import torch
sys.path.insert(0, ".")
from pychips_b200 import engine as eng, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = eng.require_cuda()
img, r = synth.synthetic_disk(n, 1)
det = eng.Detector(eng.Geometry(n, n, n // 2, n // 2, r), 51, 500, 0, 20, 4, dev)
det.detect(torch.from_numpy(img).to(dev))
torch.cuda.synchronize()
for t in (0, 19):
    cs, hier, a2 = det.contours(t)
    print(t, len(cs))
