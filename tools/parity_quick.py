import sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as NP, torch
from prisim_b200 import engine
from oracle import prisim_oracle as O
from prisim_b200 import synthetic as S
rng = NP.random.default_rng(3)
for scale, nsrc, nchan in ((1.0, 700, 128), (25.0, 3000, 256), (60.0, 2100, 64)):
    bl = S.array_baselines(S.hera_layout(3))[0] * scale
    freqs = 150e6 + (NP.arange(nchan) - nchan // 2) * 97656.25
    altaz = NP.stack((NP.degrees(NP.arcsin(rng.uniform(0, 1, nsrc))), rng.uniform(0, 360, nsrc)), 1)
    dense = rng.uniform(0.05, 5.0, (nsrc, nchan)) * rng.uniform(0, 1, (nsrc, 1)) ** 4
    dircos, idx = engine.sky_cull(altaz, 'altaz')
    amp = engine.dense_to_amp_table(torch.as_tensor(dense).cuda())
    amp32 = engine.amp_table_to_dense(amp, nsrc, nchan).double().cpu().numpy()
    Vo = O.skyvis_snapshot(bl, altaz, amp32, freqs, NP.asarray([90.0, 270.0]))
    rms_b = NP.sqrt(NP.mean(NP.abs(Vo) ** 2, axis=1, keepdims=True))
    for m in sys.argv[1:] or ['recurrence', 'recurrence_scalar', 'direct']:
        V = engine.skyvis(dircos, amp, nsrc, bl, (0, 0, 1.0), freqs, method=m).cpu().numpy()
        print(scale, nsrc, nchan, m, 'err/rms_b max %.3e  rms %.3e' % ((NP.abs(V - Vo) / rms_b).max(), NP.sqrt(NP.mean((NP.abs(V - Vo) / rms_b) ** 2))))
