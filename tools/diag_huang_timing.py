"""Per-block timing of k_huang4 (needs a library built with -DCHIPS_DEBUG_TIMING)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pychips_b200.build as _b; _b.LIB = os.environ.get("CHIPS_LIB", _b.LIB)
from pychips_b200 import engine, synth, _lib
n, k = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 51
img, r = synth.synthetic_disk(n, 0)
dev = engine.require_cuda()
t = torch.from_numpy(img).to(dev)
c = n // 2
det = engine.Detector(engine.Geometry(n, n, c, c, r), k, 500, 0, 20, 4, dev)
p = det.plan
nb = p.blk_nbx * p.blk_nby
dbg = det.ws[p.off_parent_c:p.off_parent_c + nb * 16].view(torch.int32)
lib = _lib.load()
for stages in (1, 2):
    lib.chips_medfilt2d_stages(engine._ptr(t), engine._ptr(det.filt), det._pp, c, c, r, engine._ptr(det.ws), engine._stream_ptr(), stages)
dbg.zero_()
lib.chips_medfilt2d_stages(engine._ptr(t), engine._ptr(det.filt), det._pp, c, c, r, engine._ptr(det.ws), engine._stream_ptr(), 4)
torch.cuda.synchronize()
d = dbg.cpu().numpy().reshape(p.blk_nby, p.blk_nbx, 4).astype(np.int64) & 0xFFFFFFFF
cyc, sm, rows, t0 = d[..., 0], d[..., 1], d[..., 2], d[..., 3]
act = cyc > 0
print("blocks", nb, "active", act.sum(), "L", p.blk_rows)
print("cycles: mean %.0f p50 %.0f p90 %.0f max %.0f" % (cyc[act].mean(), np.median(cyc[act]), np.percentile(cyc[act], 90), cyc[act].max()))
print("cycles per row: mean %.0f max %.0f" % ((cyc[act] / (rows[act] + k - 1)).mean(), (cyc[act] / (rows[act] + k - 1)).max()))
start = (t0 - t0[act].min()) & 0xFFFFFFFF
print("start offsets (cycles): p50 %.0f p90 %.0f max %.0f" % (np.median(start[act]), np.percentile(start[act], 90), start[act].max()))
end = start + cyc
print("end max %.0f" % end[act].max())
per_sm = np.bincount(sm[act].ravel(), minlength=148)
print("warps per SM: min %d max %d" % (per_sm.min(), per_sm.max()))
np.set_printoptions(linewidth=250)
print((cyc // 1000)[:, :])
