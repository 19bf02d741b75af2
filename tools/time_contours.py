"""Time chips_contours on the threshold masks of a 4096^2 detection (GPU) next to cv2.findContours."""
import sys, time
import numpy as np, torch, cv2
sys.path.insert(0, ".")
from pychips_b200 import engine as eng, synth, contours as pc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = eng.require_cuda()
img, r = synth.synthetic_disk(n, 1)
det = eng.Detector(eng.Geometry(n, n, n // 2, n // 2, r), 51, 500, 0, 20, 4, dev)
det.detect(torch.from_numpy(img).to(dev))
torch.cuda.synchronize()
lev = det.level.cpu().numpy()
det.contours(19)                                   # sizes the buffers
tr = det._tracer
for t in (0, 10, 19):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lib = tr.lib
    torch.cuda._sleep(20_000_000)                  # ~10 ms of GPU spin: the launches below queue up behind it,
    e0.record()                                    # so the events bracket device time, not host enqueue time
    for _ in range(5):
        eng._lib.check(lib.chips_contours(eng._ptr(det.level), n, n, n, 0, t, tr.cap, tr.start_rounds,
                                          eng._ptr(tr.recs), tr.max_contours, eng._ptr(tr.points), tr.max_points,
                                          eng._ptr(tr.counts), eng._ptr(tr.ws), tr.ws.numel(), eng._stream_ptr()), "c")
    e1.record(); torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / 5
    t0 = time.perf_counter(); got, hier, a2 = det.contours(t); wall = (time.perf_counter() - t0) * 1e3
    mask = (lev <= t).astype(np.uint8)
    t0 = time.perf_counter(); ref, hi = cv2.findContours(mask, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE); cv_ms = (time.perf_counter() - t0) * 1e3
    cnt = tr.counts.cpu().numpy().view(pc.COUNTS_DTYPE)[0]
    print(f"t={t}: border px {cnt['n_border_px']}, contours {len(got)}, points {cnt['n_points']}: device {dev_ms:.3f} ms, "
          f"call+download+assemble {wall:.2f} ms, cv2.findContours {cv_ms:.2f} ms, cap {tr.cap}", flush=True)

import time
t0 = time.perf_counter()
tot = 0
for t in range(20):
    cs, hier, a2 = det.contours(t)
    tot += len(cs)
print(f"all 20 thresholds through Detector.contours: {(time.perf_counter() - t0) * 1e3:.1f} ms wall, {tot} contours")
det.contours_all()
t0 = time.perf_counter()
tot = sum(len(cs) for cs, _, _ in det.contours_all())
print(f"all 20 thresholds through Detector.contours_all (one enqueue, one sync): {(time.perf_counter() - t0) * 1e3:.1f} ms wall, {tot} contours")
t0 = time.perf_counter()
for t in range(20):
    cv2.findContours((lev <= t).astype(np.uint8), cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
print(f"all 20 thresholds through cv2.findContours (mask build included): {(time.perf_counter() - t0) * 1e3:.1f} ms wall")
