import sys, os, json
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import agentfield_b200 as afb
import bench
dev = torch.device("cuda", 0)
ctx = afb.Context(0)
d_pks, d_sigs, d_msgs, d_off, expect = bench.make_workload(ctx, dev, 0)
N = bench.N_ITEMS
d_ok = torch.empty(N, dtype=torch.uint8, device=dev)
ctx.verify_dev(d_pks, d_sigs, d_msgs, d_off, N, d_ok); torch.cuda.synchronize()
for n in (16384, 65536, 131072, 262144):
    for _ in range(3): ctx.verify_dev(d_pks[:n], d_sigs[:n], d_msgs, d_off[:n+1], n, d_ok[:n])
    torch.cuda.synchronize()
    ctx.profile_begin(4096)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ctx.verify_dev(d_pks[:n], d_sigs[:n], d_msgs, d_off[:n+1], n, d_ok[:n])
    e1.record(); torch.cuda.synchronize()
    prof = ctx.profile_end()
    print(n, "call %.3f ms" % (e0.elapsed_time(e1) / 20), {k: round(v["avg_ms"], 4) for k, v in prof.items() if v["count"]})
