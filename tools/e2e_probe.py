"""Which part of the end-to-end series path costs throughput?  (H2D, D2H of the map, prob stack)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pychips_b200 import series, synth
n, B = 4096, 32
host, radii = [], []
for i in range(B):
    img, r = synth.synthetic_disk(n, i)
    host.append(torch.from_numpy(img).pin_memory()); radii.append(r)
dev = torch.device("cuda", 0)
def rate(**kw):
    skip_h2d = kw.pop("skip_h2d", False)
    rn = series.SeriesRunner(n, n, 51, 500, (0, 20), device=dev, n_streams=8, want_maps=True, **kw)
    if skip_h2d:
        for t in rn.dev_in: t.copy_(host[0])
        orig = torch.Tensor.copy_
        # emulate resident inputs: submit() copies into dev_in; make that copy a no-op by passing device tensors
        imgs = [rn.dev_in[i % 8] for i in range(B)]
    else:
        imgs = host
    for _ in range(2): rn.run(imgs, radii)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    prev = None
    for _ in range(6):
        h = rn.submit(imgs, radii)
        if prev: prev.result()
        prev = h
    prev.result(); torch.cuda.synchronize()
    return 6 * B / (time.perf_counter() - t0)
print("full e2e               ", rate(want_keep=True, want_prob_map=True))
print("no map D2H             ", rate(want_keep=False, want_prob_map=True))
print("no prob stack          ", rate(want_keep=True, want_prob_map=False))
print("device inputs (no H2D) ", rate(want_keep=True, want_prob_map=True, skip_h2d=True))
print("device inputs, no D2H  ", rate(want_keep=False, want_prob_map=True, skip_h2d=True))
