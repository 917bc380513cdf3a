"""Speed (and, against the first library run, bit-level agreement) of the fp64 phase-sum kernel with and without the taper
for the library named by PB200_LIB (tools/variants.sh).  usage: PB200_LIB=build/var/lib_x.so python tools/perf_fp64.py [nsrc nbl nchan]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as NP, torch
from prisim_b200 import engine
nsrc = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
nbl = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
nchan = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
rng = NP.random.default_rng(0)
torch.manual_seed(0)
alt = NP.degrees(NP.arcsin(rng.uniform(0, 1, nsrc))); az = rng.uniform(0, 360, nsrc)
dircos, idx = engine.sky_cull(NP.stack((alt, az), 1), "altaz")
amp = torch.rand((nchan // 128) * ((nsrc + 31) // 32 * 32) * 128, device="cuda", dtype=torch.float64)   # [slab][nsrc_pad][128]
bl = rng.normal(0, 150, (nbl, 3)); bl[:, 2] *= 0.01
freqs = 150e6 + (NP.arange(nchan) - nchan // 2) * 97656.25
fw = engine._f64(rng.uniform(0.3, 1.5, nsrc), 0)
bl_d = engine._f64(bl, 0)
out = torch.empty((nbl, nchan), dtype=torch.complex128, device="cuda")
name = os.path.basename(os.environ.get("PB200_LIB", "default"))
for label, kw in (("fp64", {}), ("fp64+taper", {"src_fwhm_deg": fw})):
    run = lambda: engine.skyvis(dircos, amp, nsrc, bl_d, (0.0, 0.0, 1.0), freqs, method="fp64", out=out, **kw)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); run(); run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    cache = "/dev/shm/pb200_fp64_%s_%d_%d_%d.pt" % (label, nsrc, nbl, nchan)
    if os.path.exists(cache):
        ref = torch.load(cache).cuda()
        rel = ((out - ref).abs().amax(dim=1) / ref.abs().pow(2).mean(dim=1).sqrt()).max().item()
    else:
        torch.save(out.cpu(), cache); rel = 0.0
    print("%-22s %-11s %8.2f ms  %.3f Tterms/s   max|dV|/rms vs first run %.2e" % (name, label, ms, nsrc * nbl * nchan / ms / 1e9, rel), flush=True)
