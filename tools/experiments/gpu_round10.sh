#!/bin/bash
set -x
mkdir -p gpurun_out
for lib in libafcrypto.so libafcrypto_sha2.so; do
AFC_LIB=$PWD/agentfield_b200/$lib python - <<'PY'
import os, numpy as np, torch, agentfield_b200 as afb
from oracle import c_oracle as CO
ctx=afb.Context(0); dev=torch.device("cuda",0)
g=torch.Generator(device=dev); g.manual_seed(3)
n=4_000_000
bodies=torch.randint(0,256,(n,256),dtype=torch.uint8,device=dev,generator=g); keys=torch.randint(0,256,(n,32),dtype=torch.uint8,device=dev,generator=g)
off=torch.arange(n+1,device=dev,dtype=torch.int64)*256; koff=(torch.arange(n+1,device=dev,dtype=torch.int64)*32).to(torch.int32)
tags=torch.empty((n,32),dtype=torch.uint8,device=dev)
def timed(fn,reps=5):
    fn(); torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/reps
ms=timed(lambda: ctx.hmac_sha256_dev(keys.view(-1),koff,bodies.view(-1),off,n,tags))
m=20000
exp=CO.hmac_sha256_batch(keys[:m].cpu().numpy().reshape(-1), np.arange(m+1,dtype=np.uint32)*32, bodies[:m].cpu().numpy().reshape(-1), np.arange(m+1,dtype=np.uint64)*256, 8)
assert (tags[:m].cpu().numpy()==exp).all()
ms2=timed(lambda: ctx.sha256_dev(bodies.view(-1),off,n,tags))
print(os.environ["AFC_LIB"].split("/")[-1], "hmac %.3f ms %.3f G/s | sha256 %.3f ms %.3f G/s | compress %s" % (ms, n/ms/1e6, ms2, n/ms2/1e6, ctx.microbench(3,2000)))
PY
done
