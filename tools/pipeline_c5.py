"""Config 5 pipeline on one GPU: HERA-350 x 1024 ch x 300k-source catalogue, visibilities + Tsys noise + windowed
delay transform, a few of the 1000 snapshots at full size; prints the per-stage device time per snapshot."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as NP, torch
from prisim_b200 import synthetic as S
from prisim_b200.delay_spectrum import windowing
from prisim_b200.interferometry import InterferometerArray, SimpleTime

nsnap = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cfg = S.config5(nsnap=nsnap)
ia = InterferometerArray(cfg["labels"], cfg["baselines"], cfg["channels"], telescope=cfg["telescope"], latitude=cfg["latitude"],
                         skycoords="radec", pointing_coords="hadec", A_eff=cfg["A_eff"], eff_Q=cfg["eff_Q"], device=0, noise_seed=5)
window = 1024 * windowing(1024, "bhw", area_normalize=True)
def timed(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); return time.perf_counter() - t0
t_obs = []
for j in range(nsnap):
    t_obs.append(timed(lambda: ia.observe(SimpleTime(2451545.0 + j * 10.7 / 86400, j * 10.7 / 240.0), cfg["Tsysinfo"], NP.ones(1024),
                                          cfg["pointing_hadec"], cfg["skymodel"], cfg["t_acc"])))
t_noise = timed(lambda: (ia.generate_noise(), ia.add_noise()))
t_dt = timed(lambda: ia.delay_transform(pad=1.0, freq_wts=window, verbose=False))
nsrc = [int(m.size) for m in ia.obs_catalog_indices]
terms = sum(nsrc) * 61075 * 1024
res = {"config": "C5: HERA-350 (61,075 bl) x 1024 ch x 300k-source catalogue, %d of 1000 snapshots" % nsnap,
       "observe_s_per_snapshot": t_obs, "noise_s_per_snapshot": t_noise / nsnap, "delay_transform_s_per_snapshot": t_dt / nsnap,
       "pipeline_gterms_per_s": terms / (sum(t_obs) + t_noise + t_dt) / 1e9, "precision_report": ia.precision_report,
       "extrapolated_1000_snapshots_1gpu_minutes": (sum(t_obs) + t_noise + t_dt) / nsnap * 1000 / 60}
print(json.dumps(res, indent=1))
