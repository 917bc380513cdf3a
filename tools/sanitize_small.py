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
# kernel variants, gradient tables, sub-band transforms, redundant-set expansion, gridded beam, phase rotation
for method in ("recurrence", "recurrence_lift", "recurrence_3term", "recurrence_3term_scalar", "recurrence_scalar", "direct", "fp64"):
    ic = InterferometerArray(c1["labels"], c1["baselines"], c1["channels"], telescope=c1["telescope"], latitude=c1["latitude"], skycoords="radec", pointing_coords="hadec", device=0)
    ic.precision = "fp64" if method == "fp64" else "fp32"
    ic.skyvis_method = "auto" if method == "fp64" else method
    ic.observe(SimpleTime(2451545.0, 10.0), {"Tnet": 200.0}, NP.ones(128), c1["pointing_hadec"], c1["skymodel"], c1["t_acc"],
               gradient_mode="baseline" if method in ("recurrence", "fp64") else None)
ib.multi_window_delay_transform([1.0e6, 2.0e6], freq_center=[c1["channels"][30], c1["channels"][90]], shape="bhw", verbose=False)
ib.rotate_visibilities({"location": NP.asarray([[10.0, -25.0]]), "coords": "hadec"}, do_delay_transform=False, verbose=False)
lab = [tuple(l) for l in (ib.labels.tolist() if hasattr(ib.labels, "tolist") else ib.labels)]
ib.duplicate_measurements({lab[0]: [lab[0], ("x1", "x0"), ("x3", "x2")], lab[1]: [lab[1]] + [("y{0}".format(i), "z") for i in range(len(lab))]})
torch.cuda.synchronize()
print("ok", ia.precision_report, ib.baselines.shape, float(abs(ib.vis_freq).max()))
