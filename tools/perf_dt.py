import sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import torch
from prisim_b200 import engine
nbl, nchan = 61075, 1024
V = torch.complex(torch.rand((nbl, nchan), dtype=torch.float64, device='cuda'), torch.rand((nbl, nchan), dtype=torch.float64, device='cuda'))
bp = torch.ones(nchan, dtype=torch.float64, device='cuda'); w = torch.rand(nchan, dtype=torch.float64, device='cuda')
for _ in range(3): out = engine.delay_transform(V, bp, w, 97656.25, pad=1.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): out = engine.delay_transform(V, bp, w, 97656.25, pad=1.0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("delay transform %.3f ms  %.1f GB/s" % (ms, nbl * nchan * 32 / ms / 1e6))
