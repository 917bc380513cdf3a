import sys, time
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as NP, torch
from prisim_b200 import engine, _lib
nsrc = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
nbl = int(sys.argv[2]) if len(sys.argv) > 2 else 61075
nchan = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
rng = NP.random.default_rng(0)
alt = NP.degrees(NP.arcsin(rng.uniform(0, 1, nsrc))); az = rng.uniform(0, 360, nsrc)
dircos, idx = engine.sky_cull(NP.stack((alt, az), 1), 'altaz')
amp = engine.dense_to_amp_table(torch.rand((nsrc, nchan), device='cuda'))
bl = rng.normal(0, 150, (nbl, 3)); bl[:, 2] *= 0.01
if nbl == 61075:      # the HERA-350 baselines of config 2 (sorted by length, as orient_and_sort_baselines leaves them)
    from prisim_b200 import synthetic as S
    bl = NP.asarray(S.config2(nsrc=16)['baselines'])
freqs = 150e6 + (NP.arange(nchan) - nchan // 2) * 97656.25
pc = NP.array([0, 0, 1.0])
out = torch.empty((nbl, nchan), dtype=torch.complex128, device='cuda')
bl_d = engine._f64(bl, 0)
for method in sys.argv[4:] or ['recurrence']:
    for i in range(2):
        engine.skyvis(dircos, amp, nsrc, bl_d, pc, freqs, method=method, out=out)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    nrep = 3
    for i in range(nrep):
        engine.skyvis(dircos, amp, nsrc, bl_d, pc, freqs, method=method, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / nrep
    terms = nsrc * nbl * nchan
    print(method, 'nsrc', nsrc, 'ms', ms, 'Tterms/s', terms / ms / 1e9, 'frac of 6.2', terms / ms / 1e9 / 6.2)
