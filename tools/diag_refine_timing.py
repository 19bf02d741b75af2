"""Per-tile phase timing of k_refine (needs a library built with -DCHIPS_DEBUG_TIMING)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pychips_b200.build as _b; _b.LIB = os.environ.get("CHIPS_LIB", _b.LIB)
from pychips_b200 import engine, synth

n, k = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 51
img, r = synth.synthetic_disk(n, 0)
dev = engine.require_cuda()
t = torch.from_numpy(img).to(dev)
c = n // 2
det = engine.Detector(engine.Geometry(n, n, c, c, r), k, 500, 0, 20, 4, dev)
p = det.plan
gx, gy = (p.roi_w + 15) // 16, (p.roi_h + 15) // 16
dbg = det.ws[p.off_parent_c:p.off_parent_c + gx * gy * 16].view(torch.int32)
dbg.zero_()
det.medfilt(t)
torch.cuda.synchronize()
d = dbg.cpu().numpy().reshape(gy, gx, 4).astype(np.int64)
act = d[..., 0] > 0
print("active tiles", act.sum(), "of", gx * gy)
for i, name in enumerate(["gather+scan", "sort", "walk"]):
    v = d[..., i][act]
    print(f"{name:12s} mean {v.mean():9.0f} p50 {np.median(v):9.0f} p90 {np.percentile(v,90):9.0f} p99 {np.percentile(v,99):9.0f} max {v.max():9.0f} cycles")
tot = d[..., :3].sum(-1)[act]
print("total mean", tot.mean(), "sum/SMs/3 (ms @1.965GHz):", tot.sum() / 148 / 3 / 1.965e6)
C = d[..., 3][act]
print("C mean", C.mean(), "p99", np.percentile(C, 99), "max", C.max())
# where are the slow tiles
slow = np.argwhere((d[..., :3].sum(-1) > np.percentile(tot, 99)) & act)
print("slow tiles (ty,tx):", slow[:10].tolist())
big = tot > 4 * np.median(tot)
print("share of time in tiles >4x median:", tot[big].sum() / tot.sum(), "count", big.sum())
