"""Brute-force version of tests/test_gpu_parity.py::test_skyvis_random_shapes_property: every method x slabs-per-CTA over many ragged
shapes against the oracle; prints every failing case.  usage: python tools/fuzz_methods.py [ncases]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as NP, torch
from prisim_b200 import engine as eng, _lib
from oracle import prisim_oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng0 = NP.random.default_rng(12345)
ctx = _lib.get_context(0)
bad = 0
for case in range(n):
    nsrc, nbl, nchan, seed = int(rng0.integers(1, 201)), int(rng0.integers(1, 151)), int(rng0.integers(2, 301)), int(rng0.integers(0, 10 ** 6))
    rng = NP.random.default_rng(seed)
    bl = rng.normal(0, 80.0, (nbl, 3)); bl[:, 2] *= 0.05
    freqs = 150e6 + (NP.arange(nchan) - nchan // 2) * 97656.25
    altaz = NP.stack((NP.degrees(NP.arcsin(rng.uniform(0, 1, nsrc))), rng.uniform(0, 360, nsrc)), 1)
    dense = rng.uniform(0.05, 5.0, (nsrc, nchan))
    dircos, _ = eng.sky_cull(altaz, "altaz")
    Vo = None
    for method in ("auto", "recurrence_lift", "recurrence_3term", "recurrence_3term_scalar", "recurrence_quarter", "recurrence_scalar", "direct", "fp64"):
        amp = eng.dense_to_amp_table(torch.as_tensor(dense).cuda(), dtype=torch.float64 if method == "fp64" else torch.float32)
        amp_used = eng.amp_table_to_dense(amp, nsrc, nchan).double().cpu().numpy()
        Vo = O.skyvis_snapshot(bl, altaz, amp_used, freqs, NP.asarray([90.0, 270.0]))
        rms_b = NP.sqrt(NP.mean(NP.abs(Vo) ** 2, axis=1, keepdims=True))
        for spc in (1, 2, 4):
            ctx.set_option("skyvis_spc", spc)
            V = eng.skyvis(dircos, amp, nsrc, bl, (0.0, 0.0, 1.0), freqs, method=method).cpu().numpy()
            err = float((NP.abs(V - Vo) / rms_b).max())
            tol = 1e-11 if method == "fp64" else 1e-5
            if not err <= tol:
                bad += 1
                print("FAIL nsrc=%d nbl=%d nchan=%d seed=%d method=%s spc=%d err=%.3e" % (nsrc, nbl, nchan, seed, method, spc, err), flush=True)
ctx.set_option("skyvis_spc", 0)
print("cases %d, failures %d" % (n, bad))
