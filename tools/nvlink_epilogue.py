"""Peer-memory epilogue of the phase-sum kernel, in one process on two GPUs (so that ncu can count NVLink bytes): the kernels run on
cuda:1, the output tensor lives on cuda:0, k_skyvis_finalize stores the visibilities across NVLink -- what every non-writing rank of
sharding.PeerGatherBuffer does.  usage: [ncu --metrics nvl... -k regex:k_skyvis_finalize] python tools/nvlink_epilogue.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as NP, torch
from prisim_b200 import engine
assert torch.cuda.device_count() >= 2
nbl, nchan, nsrc = 7635, 1024, 2048                       # one eighth of HERA-350, all channels
rng = NP.random.default_rng(0)
out0 = torch.zeros((nbl, nchan), dtype=torch.complex128, device="cuda:0")
torch.zeros(8, device="cuda:0").to("cuda:1")              # first cross-device copy: torch enables peer access 0 <-> 1
torch.zeros(8, device="cuda:1").to("cuda:0")
altaz = NP.stack((NP.degrees(NP.arcsin(rng.uniform(0, 1, nsrc))), rng.uniform(0, 360, nsrc)), 1)
bl = rng.normal(0, 150.0, (nbl, 3)) * NP.asarray([1, 1, 0.02])
freqs = 150e6 + (NP.arange(nchan) - nchan // 2) * 97656.25
with torch.cuda.device(1):
    dircos, _ = engine.sky_cull(altaz, "altaz", device=1)
    amp = engine.dense_to_amp_table(torch.rand((nsrc, nchan), device="cuda:1"))
    for _ in range(3):
        engine.skyvis(dircos, amp, nsrc, bl, (0, 0, 1.0), freqs, out=out0, device=1)      # stores land in cuda:0's HBM
    torch.cuda.synchronize(1)
    ref = engine.skyvis(dircos, amp, nsrc, bl, (0, 0, 1.0), freqs, device=1)
    torch.cuda.synchronize(1)
same = torch.equal(out0.cpu(), ref.cpu())
print("peer-stored result equals the local one: %s; payload per launch %.1f MB (%d x %d complex128)" % (same, nbl * nchan * 16 / 1e6, nbl, nchan))
