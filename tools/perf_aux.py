import sys, json
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as NP, torch
from prisim_b200 import engine
nbl, nchan = 61075, 1024
dev = 'cuda'
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
V = torch.complex(torch.rand((nbl, nchan), dtype=torch.float64, device=dev), torch.rand((nbl, nchan), dtype=torch.float64, device=dev))
tsys = torch.rand(nchan, dtype=torch.float64, device=dev) + 100
aeff = torch.full((1,), 100.0, dtype=torch.float64, device=dev); effq = torch.full((1,), 0.96, dtype=torch.float64, device=dev)
res = {}
ms = timeit(lambda: engine.noise(V, tsys, aeff, effq, 97656.25, 10.7, 1, nbl, nchan))
b = nbl * nchan * (16 + 8 + 16 + 16)
res['noise_rms_noise_vis'] = {'ms': ms, 'GBs': b / ms / 1e6, 'bytes': b}
nz = torch.empty_like(V)
ms = timeit(lambda: engine.add_noise(V, nz)); b = nbl * nchan * 48
res['add_noise'] = {'ms': ms, 'GBs': b / ms / 1e6, 'bytes': b}
bp = torch.ones(nchan, dtype=torch.float64, device=dev); w = torch.rand(nchan, dtype=torch.float64, device=dev)
for pad in (1.0, 0.0, 0.5):
    ms = timeit(lambda: engine.delay_transform(V, bp, w, 97656.25, pad=pad)); b = nbl * nchan * 32
    res['delay_transform_pad%g' % pad] = {'ms': ms, 'GBs': b / ms / 1e6, 'bytes': b}
# amp table + cull at C2 size
from prisim_b200 import synthetic as S, primary_beams as PB
cfg = S.config2(); sky = cfg['skymodel']; sp = sky.spec_parms
d_hadec = engine._f64(NP.stack((0.0 - sky.location[:, 0], sky.location[:, 1]), 1), 0)
spec = {"flux_scale": engine._f64(sp["flux-scale"], 0), "index": engine._f64(sp["power-law-index"], 0), "freq_ref": engine._f64(sp["freq-ref"], 0)}
beam = PB.beam_desc_from_telescope(cfg["telescope"], pointing_center=NP.asarray([90.0, 270.0]), skyunits="altaz", device=0)
ms = timeit(lambda: engine.sky_cull(d_hadec, "hadec", latitude_deg=cfg["latitude"]))
res['sky_cull_300k'] = {'ms': ms}
dircos, index = engine.sky_cull(d_hadec, "hadec", latitude_deg=cfg["latitude"]); nsrc = int(index.shape[0])
ms = timeit(lambda: engine.amp_table(dircos, index, nsrc, spec, beam, cfg["channels"]))
res['amp_table_airy'] = {'ms': ms, 'Melem_per_s': nsrc * 1024 / ms / 1e3, 'GBs_written': nsrc * 1024 * 4 / ms / 1e6}
print(json.dumps(res, indent=1))
