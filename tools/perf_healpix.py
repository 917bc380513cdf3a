import sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as NP, torch
from prisim_b200 import engine
nside, nchan, nsrc = 128, 256, 393216
npix = 12 * nside * nside
for dt in (torch.float32, torch.float64):
    m = torch.rand((npix, nchan), dtype=dt, device='cuda')
    rng = NP.random.default_rng(0)
    alt = NP.degrees(NP.arcsin(rng.uniform(0, 1, nsrc))); az = rng.uniform(0, 360, nsrc)
    dircos, _ = engine.sky_cull(NP.stack((alt, az), 1), 'altaz')
    for _ in range(2): engine.healpix_beam(m, nside, dircos, nsrc, nchan)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): engine.healpix_beam(m, nside, dircos, nsrc, nchan)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    b = nsrc * nchan * (4 * m.element_size() + 8 + 8)
    print(str(dt), "gather+colmax %.3f ms  algorithmic %.2f GB -> %.0f GB/s" % (ms, b / 1e9, b / ms / 1e6))
