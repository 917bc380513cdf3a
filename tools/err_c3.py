import sys
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
bl = engine._f64(cfg["baselines"][:2048], 0)
amp32 = engine.amp_table(dircos, index, nsrc, spec, beam, cfg["channels"])
amp64 = engine.amp_table(dircos, index, nsrc, spec, beam, cfg["channels"], dtype=torch.float64)
A2 = torch.sqrt(amp32.double().square().sum() / 256).item()
V64 = engine.skyvis(dircos, amp64, nsrc, bl, (0, 0, 1.0), cfg["channels"], src_fwhm_deg=fw, method="fp64")
rms_b = V64.abs().pow(2).mean(dim=1).sqrt()
for m in ("recurrence", "recurrence_scalar", "direct"):
    V = engine.skyvis(dircos, amp32, nsrc, bl, (0, 0, 1.0), cfg["channels"], src_fwhm_deg=fw, method=m)
    err = (V - V64).abs().max(dim=1).values
    print(m, "A2 %.3e  max err/A2 %.2e  median err/A2 %.2e   max err/rms_b (ratio>0.45) %.2e" % (A2, (err / A2).max().item(), (err / A2).median().item(), (err / rms_b)[rms_b > 0.45 * A2].max().item() if (rms_b > 0.45 * A2).any() else -1))
V = engine.skyvis(dircos, amp32, nsrc, bl, (0, 0, 1.0), cfg["channels"], method="recurrence")
V64n = engine.skyvis(dircos, amp64, nsrc, bl, (0, 0, 1.0), cfg["channels"], method="fp64")
err = (V - V64n).abs().max(dim=1).values
print("no taper: max err/A2 %.2e median %.2e" % ((err / A2).max().item(), (err / A2).median().item()))
print("ratio quantiles", torch.quantile(rms_b / A2, torch.tensor([0, 0.01, 0.1, 0.5, 0.9, 1.0], dtype=torch.float64, device='cuda')).cpu().numpy())
