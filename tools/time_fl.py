"""Time the filament-cleanup kernels (csrc/fl.cu) at 4096^2 and report HBM GB/s."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pychips_b200 import _lib, engine, synth
from pychips_b200.cleanup import CleanFl, DEFAULT_PARAMS
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
imgs, r = synth.synthetic_triplet(n, 0)
dev = engine.require_cuda()
lib = _lib.load()
d = [torch.from_numpy(imgs[w]).to(dev) for w in (171, 193, 211)]
bits = torch.empty((n, n), dtype=torch.uint8, device=dev)
stats = torch.zeros(C.sizeof(_lib.FlStats), dtype=torch.uint8, device=dev)
ws = torch.empty(int(lib.chips_fl_workspace_bytes(n, n)), dtype=torch.uint8, device=dev)
cf = CleanFl(synth.make_triplet_aia(64, 0), 64)
prm = cf._c_params()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run():
    _lib.check(lib.chips_fl_candidates(engine._ptr(d[0]), engine._ptr(d[1]), engine._ptr(d[2]), n, n, 4, n // 2, n // 2, r,
                                       C.byref(prm), engine._ptr(bits), engine._ptr(stats), engine._ptr(ws), engine._stream_ptr()), "fl")
for _ in range(3): run()
ts = []
for _ in range(10):
    flush.fill_(1)
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record(); run(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ms = float(np.median(ts))
alg = 3 * n * n * 4 * 2 + n * n
print(f"fl candidates {n}^2 f32: {ms:.3f} ms  alg bytes {alg/1e6:.1f} MB -> {alg/ms/1e6:.0f} GB/s ({alg/ms/1e6/6549.8:.2%} of measured HBM peak)")
