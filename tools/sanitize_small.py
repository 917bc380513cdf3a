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
for method in ("recurrence", "recurrence_lift", "recurrence_3term", "recurrence_3term_scalar", "recurrence_quarter", "recurrence_scalar", "direct", "fp64"):
    ic = InterferometerArray(c1["labels"], c1["baselines"], c1["channels"], telescope=c1["telescope"], latitude=c1["latitude"], skycoords="radec", pointing_coords="hadec", device=0)
    ic.precision = "fp64" if method == "fp64" else "fp32"
    ic.skyvis_method = "auto" if method == "fp64" else method
    ic.observe(SimpleTime(2451545.0, 10.0), {"Tnet": 200.0}, NP.ones(128), c1["pointing_hadec"], c1["skymodel"], c1["t_acc"],
               gradient_mode="baseline" if method in ("recurrence", "fp64") else None)
ib.multi_window_delay_transform([1.0e6, 2.0e6], freq_center=[c1["channels"][30], c1["channels"][90]], shape="bhw", verbose=False)
ib.rotate_visibilities({"location": NP.asarray([[10.0, -25.0]]), "coords": "hadec"}, do_delay_transform=False, verbose=False)
lab = [tuple(l) for l in (ib.labels.tolist() if hasattr(ib.labels, "tolist") else ib.labels)]
ib.duplicate_measurements({lab[0]: [lab[0], ("x1", "x0"), ("x3", "x2")], lab[1]: [lab[1]] + [("y{0}".format(i), "z") for i in range(len(lab))]})
# round 2: more output tiles than SMs (whole-tile waves + stream-K tail with head partials), every kernel family; the warp-per-row
# 1024-point delay transform; interleaved shard output (row-strided epilogue); drain with the product ring; sub-band transform
from prisim_b200 import engine
rng = NP.random.default_rng(3)
nbl2, nch2, ns2 = 64 * 150 + 5, 256, 96
bl2 = rng.normal(0.0, 150.0, (nbl2, 3)) * NP.asarray([1.0, 1.0, 0.02])
f2 = 150e6 + (NP.arange(nch2) - nch2 // 2) * 97656.25
altaz2 = NP.stack((NP.degrees(NP.arcsin(rng.uniform(0, 1, ns2))), rng.uniform(0, 360, ns2)), 1)
dc2, _ = engine.sky_cull(altaz2, "altaz")
dense2 = torch.rand((ns2, nch2), device="cuda", dtype=torch.float64)
amp2, amp2d = engine.dense_to_amp_table(dense2), engine.dense_to_amp_table(dense2, dtype=torch.float64)
fw2 = engine._f64(rng.uniform(0.05, 0.5, ns2), 0)
for kw in (dict(method="recurrence", nsrc_bright=40), dict(method="direct"), dict(method="recurrence", src_fwhm_deg=fw2)):
    engine.skyvis(dc2, amp2, ns2, bl2, (0, 0, 1.0), f2, **kw)
engine.skyvis(dc2, amp2d, ns2, bl2, (0, 0, 1.0), f2, method="fp64")
engine.skyvis(dc2, amp2d, ns2, bl2[:2000], (0, 0, 1.0), f2, method="fp64", src_fwhm_deg=fw2)
big = torch.zeros((3 * 700, nch2), dtype=torch.complex128, device="cuda")
engine.skyvis(dc2, amp2, ns2, bl2[:700], (0, 0, 1.0), f2, out=big[1::3])                     # strided rows
x1024 = torch.complex(torch.rand((37, 1024), dtype=torch.float64, device="cuda"), torch.rand((37, 1024), dtype=torch.float64, device="cuda"))
engine.delay_transform(x1024, torch.ones(1024, dtype=torch.float64, device="cuda"), torch.rand(1024, dtype=torch.float64, device="cuda"), 1e5, pad=1.0)
engine.delay_transform(None, torch.rand((37, 1024), dtype=torch.float64, device="cuda"), None, 1e5, pad=0.0, nrows=37, nchan=1024)
idr = InterferometerArray(c1["labels"], c1["baselines"], c1["channels"], telescope=c1["telescope"], latitude=c1["latitude"], skycoords="radec", pointing_coords="hadec", device=0, A_eff=100.0, eff_Q=0.9, noise_seed=3)
for j in range(3):
    idr.observe(SimpleTime(2451545.0 + j, 10.0 + j), {"Tnet": 200.0}, NP.ones(128), c1["pointing_hadec"], c1["skymodel"], c1["t_acc"])
    idr.drain(lambda j, p: None, noise=True, delay_transform={"pad": 1.0}, ring=2)
from prisim_b200.delay_spectrum import DelaySpectrum
ids = InterferometerArray(c1["labels"], c1["baselines"], c1["channels"], telescope=c1["telescope"], latitude=c1["latitude"], skycoords="radec", pointing_coords="hadec", device=0, A_eff=100.0, eff_Q=0.9, noise_seed=3)
ids.observe(SimpleTime(2451545.0, 10.0), {"Tnet": 200.0}, NP.ones(128), c1["pointing_hadec"], c1["skymodel"], c1["t_acc"])
ids.generate_noise(); ids.add_noise()
DelaySpectrum(interferometer_array=ids).subband_delay_transform({"sim": [1.0e6, 2.0e6]}, freq_center={"sim": [c1["channels"][30], c1["channels"][90]]},
                                                                shape={"sim": "bhw"}, verbose=False)
torch.cuda.synchronize()
print("ok", ia.precision_report, ib.baselines.shape, float(abs(ib.vis_freq).max()))
