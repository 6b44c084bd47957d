#!/usr/bin/env python3
"""One cold 1 M-credential verify (BASELINE configs[1]) between cudaProfilerStart/Stop, for ncu --set full."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import agentfield_b200 as afb
import bench

dev = torch.device("cuda", 0)
ctx = afb.Context(0)
d_pks, d_sigs, d_msgs, d_off, expect = bench.make_workload(ctx, dev, 0)
n = bench.N_ITEMS
d_ok = torch.empty(n, dtype=torch.uint8, device=dev)
for _ in range(2):
    ctx.keycache_clear()
    ctx.verify_dev(d_pks, d_sigs, d_msgs, d_off, n, d_ok)
torch.cuda.synchronize()
assert torch.equal(d_ok, expect)
torch.cuda.cudart().cudaProfilerStart()
ctx.keycache_clear()
ctx.verify_dev(d_pks, d_sigs, d_msgs, d_off, n, d_ok)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
