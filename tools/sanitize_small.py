import sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as NP, torch
from prisim_b200 import synthetic as S
from prisim_b200.interferometry import InterferometerArray, SimpleTime
from oracle import prisim_oracle as O
cfg = S.config3(nside=8, nchan=256, n_side=3, nsnap=1)
ia = InterferometerArray(cfg["labels"], cfg["baselines"], cfg["channels"], telescope=cfg["telescope"], latitude=cfg["latitude"], skycoords="radec", pointing_coords="hadec", device=0, A_eff=100.0, eff_Q=0.9)
ia.observe(SimpleTime(2451545.0, 10.0), {"Tnet": 200.0}, NP.ones(256), cfg["pointing_hadec"], cfg["skymodel"], cfg["t_acc"])
c1 = S.config1(nsrc=200, nchan=128, nsnap=1)
ib = InterferometerArray(c1["labels"], c1["baselines"], c1["channels"], telescope=c1["telescope"], latitude=c1["latitude"], skycoords="radec", pointing_coords="hadec", device=0, A_eff=100.0, eff_Q=0.9)
ib.precision = "fp32"
ib.observe(SimpleTime(2451545.0, 10.0), {"Tnet": 200.0}, NP.ones(128), c1["pointing_hadec"], c1["skymodel"], c1["t_acc"])
ib.generate_noise(); ib.add_noise(); ib.delay_transform(pad=1.0, freq_wts=O.windowing(128, "bhw", area_normalize=True) * 128, verbose=False)
ib.delay_transform(pad=0.5, verbose=False)
torch.cuda.synchronize()
print("ok", ia.precision_report, float(abs(ib.skyvis_lag).max()))
