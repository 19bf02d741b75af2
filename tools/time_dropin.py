"""Wall-clock of the drop-in call a user makes: Chips(aia).run_CHIPS() on one 4096^2 float64 disk."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from loguru import logger
logger.remove()
from pychips_b200 import Chips, synth
img, r = synth.synthetic_disk(4096, 0)
torch.cuda.init(); torch.zeros(1, device="cuda")
for rep in range(3):
    d = synth.FakeDisk(img.astype(np.float64), r)
    aia = synth.FakeAIA(d)
    t0 = time.perf_counter()
    c = Chips(aia, base_folder="/tmp/chips_b200_time/")
    c.run_CHIPS()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    keys = list(d.solar_ch_regions.__dict__)
    m = d.solar_ch_regions.__dict__[keys[-1]].map          # lazy D2H of one float64 map
    t2 = time.perf_counter()
    print(f"run {rep}: run_CHIPS {1e3*(t1-t0):.1f} ms, first map access {1e3*(t2-t1):.1f} ms, probs {d.solar_ch_regions.__dict__[keys[0]].prob}")
