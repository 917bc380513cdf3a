"""Full-grid error and speed of the fp32 phase-sum kernel on config 2 for the library named by PB200_LIB (tools/variants.sh):
every one of the 61,075 x 1024 cells against the fp64 kernel (the reference values are cached in /dev/shm between
variants).  usage: PB200_LIB=build/var/lib_x.so python tools/err_c2.py [sorted]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as NP, torch
from prisim_b200 import engine, synthetic as S, primary_beams as PB
cfg = S.config2()
sky, sp = cfg["skymodel"], cfg["skymodel"].spec_parms
hadec = engine._f64(NP.stack((0.0 - sky.location[:, 0], sky.location[:, 1]), axis=1), 0)
spec = {"flux_scale": engine._f64(sp["flux-scale"], 0), "index": engine._f64(sp["power-law-index"], 0), "freq_ref": engine._f64(sp["freq-ref"], 0)}
beam = PB.beam_desc_from_telescope(cfg["telescope"], pointing_center=NP.asarray([90.0, 270.0]), skyunits="altaz", device=0)
dircos, index = engine.sky_cull(hadec, "hadec", latitude_deg=cfg["latitude"])
nsrc = int(index.shape[0]); bl = engine._f64(cfg["baselines"], 0); pc = (0.0, 0.0, 1.0)
cache = "/dev/shm/pb200_v64_c2.pt"
if os.path.exists(cache):
    V64 = torch.load(cache).cuda()
else:
    amp64 = engine.amp_table(dircos, index, nsrc, spec, beam, cfg["channels"], dtype=torch.float64)
    V64 = engine.skyvis(dircos, amp64, nsrc, bl, pc, cfg["channels"], method="fp64")
    del amp64
    torch.save(V64.cpu(), cache)
rms_b = V64.abs().pow(2).mean(dim=1, keepdim=True).sqrt()
for mode in sys.argv[1:] or ["plain"]:
    dc, ix, nb = dircos, index, 0
    if mode.startswith("sorted"):
        frac = float(mode.split(":")[1]) if ":" in mode else 0.9
        perm, nb = engine.brightness_order(dircos, index, nsrc, spec, beam, cfg["channels"], power_fraction=frac)
        dc, ix = dircos.index_select(0, perm).contiguous(), index.index_select(0, perm).contiguous()
    amp = engine.amp_table(dc, ix, nsrc, spec, beam, cfg["channels"])
    method = os.environ.get("PB200_METHOD", "auto")
    run = lambda: engine.skyvis(dc, amp, nsrc, bl, pc, cfg["channels"], nsrc_bright=nb, method=method)
    V = run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    err = ((V - V64).abs() / rms_b)
    emax = err.amax(dim=1)
    print("%-28s %-12s bright=%6d  %.1f ms  %.3f Tterms/s   max %.3e  p99.9(bl) %.3e  median(bl) %.3e  rms %.3e" % (
        os.path.basename(os.environ.get("PB200_LIB", "default")), mode, nb, ms, nsrc * bl.shape[0] * 1024 / ms / 1e9,
        emax.max().item(), emax.quantile(0.999).item(), emax.median().item(), err.pow(2).mean().sqrt().item()), flush=True)
    # absolute error in units of the incoherent norm A2 = sqrt(mean_f sum_s a^2): what the 'auto' cancellation threshold rests on
    a2 = torch.sqrt(amp.double().square().sum() / 1024)
    ratio = (rms_b.flatten() / a2)
    abs_a2 = ((V - V64).abs().amax(dim=1) / a2)
    print("    A2 = %.4g; max |dV|/A2 = %.3e (p99.9 %.3e); baselines with rms_b/A2 < 0.45 / 0.3 / 0.2 / 0.15 / 0.1: %d / %d / %d / %d / %d; worst err/rms_b among rms_b/A2 >= 0.2: %.3e" % (
        a2.item(), abs_a2.max().item(), abs_a2.quantile(0.999).item(), (ratio < 0.45).sum().item(), (ratio < 0.3).sum().item(),
        (ratio < 0.2).sum().item(), (ratio < 0.15).sum().item(), (ratio < 0.1).sum().item(), emax[ratio >= 0.2].max().item()), flush=True)
    del amp, V
