"""Diagnostics for the tier-2 median kernel: per-tile candidate counts on a real image."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pychips_b200 import engine, synth

n, k = 4096, 51
img, r = synth.synthetic_disk(n, 0)
dev = engine.require_cuda()
t = torch.from_numpy(img).to(dev)
c = n // 2
det = engine.Detector(engine.Geometry(n, n, c, c, r), k, 500, 0, 20, 4, dev)
det.medfilt(t)
torch.cuda.synchronize()
p = det.plan
bstar = det.view("bstar", torch.uint8, (n, n)).cpu().numpy()
rres = det.view("rres", torch.int16, (n, n)).cpu().numpy()
Bp = det.ws[p.off_bucket:p.off_bucket + p.bp_pitch * p.bp_rows].view(torch.uint8).reshape(p.bp_rows, p.bp_pitch).cpu().numpy()
bounds = det.ws[p.off_bounds:p.off_bounds + 256 * 4].view(torch.int32).cpu().numpy().astype(np.uint32)
yy, xx = np.mgrid[0:n, 0:n]
on = (xx - c) ** 2 + (yy - c) ** 2 <= r * r
half = k // 2
x0, y0, w, h = p.roi_x0, p.roi_y0, p.roi_w, p.roi_h
Cs, needs = [], []
rng = np.random.default_rng(0)
tiles = [(ty, tx) for ty in range(0, h, 16) for tx in range(0, w, 16)]
sel = rng.choice(len(tiles), 3000, replace=False)
for i in sel:
    ty, tx = tiles[i]
    Y0, X0 = y0 + ty, x0 + tx
    a = on[Y0:Y0 + 16, X0:X0 + 16]
    if not a.any():
        continue
    bs = np.unique(bstar[Y0:Y0 + 16, X0:X0 + 16][a])
    reg = Bp[Y0:Y0 + 16 + 2 * half, X0:X0 + 16 + 2 * half]
    Cs.append(int(np.isin(reg, bs).sum()))
    needs.append(len(bs))
Cs, needs = np.array(Cs), np.array(needs)
print("active tiles sampled", len(Cs))
print("need buckets: mean %.1f p50 %d p90 %d max %d" % (needs.mean(), np.median(needs), np.percentile(needs, 90), needs.max()))
print("C: mean %.0f p50 %d p90 %d p99 %d max %d" % (Cs.mean(), np.median(Cs), np.percentile(Cs, 90), np.percentile(Cs, 99), Cs.max()))
n2 = 2 ** np.ceil(np.log2(np.maximum(Cs, 2)))
for v in (256, 512, 1024, 2048, 4096, 8192):
    print("n2=%d: %.3f" % (v, (n2 == v).mean()))
hist = np.bincount(Bp[y0 + half:y0 + half + h, x0 + half:x0 + half + w].ravel(), minlength=256)
print("bucket population min/median/max", hist.min(), np.median(hist), hist.max())
print("rres mean", rres[on].mean(), "max", rres[on].max())
