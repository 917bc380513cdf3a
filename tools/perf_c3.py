import sys, time
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as NP, torch
from prisim_b200 import synthetic as S, engine
from prisim_b200 import primary_beams as PB
cfg = S.config3(nsnap=1)
sky = cfg["skymodel"]; sp = sky.spec_parms
d_hadec = engine._f64(NP.stack((20.0 - sky.location[:, 0], sky.location[:, 1]), 1), 0)
spec = {"flux_scale": engine._f64(sp["flux-scale"], 0), "index": engine._f64(sp["power-law-index"], 0), "freq_ref": engine._f64(sp["freq-ref"], 0)}
beam = PB.beam_desc_from_telescope(cfg["telescope"], pointing_center=NP.asarray([90.0, 270.0]), skyunits="altaz", device=0)
dircos, index = engine.sky_cull(d_hadec, "hadec", latitude_deg=cfg["latitude"]); nsrc = int(index.shape[0])
fw = engine._f64(NP.sqrt(sky.src_shape[:, 0] * sky.src_shape[:, 1]), 0).index_select(0, index.long()).contiguous()
bl = engine._f64(cfg["baselines"], 0)
amp32 = engine.amp_table(dircos, index, nsrc, spec, beam, cfg["channels"])
amp64 = engine.amp_table(dircos, index, nsrc, spec, beam, cfg["channels"], dtype=torch.float64)
terms = nsrc * bl.shape[0] * 256
def t(fn):
    fn(); torch.cuda.synchronize(); t0 = time.time(); fn(); torch.cuda.synchronize(); return time.time() - t0
for name, fn in (("fp32 notaper", lambda: engine.skyvis(dircos, amp32, nsrc, bl, (0, 0, 1.0), cfg["channels"])),
                 ("fp32 taper", lambda: engine.skyvis(dircos, amp32, nsrc, bl, (0, 0, 1.0), cfg["channels"], src_fwhm_deg=fw)),
                 ("fp64 notaper", lambda: engine.skyvis(dircos, amp64, nsrc, bl, (0, 0, 1.0), cfg["channels"], method="fp64")),
                 ("fp64 taper", lambda: engine.skyvis(dircos, amp64, nsrc, bl, (0, 0, 1.0), cfg["channels"], src_fwhm_deg=fw, method="fp64"))):
    dt = t(fn); print("%-14s %.3f s  %.2f Tterms/s" % (name, dt, terms / dt / 1e12))
