"""GPU box: build times of a cold process as the batch grows (buffer growth, worker contexts), text / packed
host paths.  python scripts/cold_start.py"""
import os, sys, time
sys.path.insert(0,'.')
import numpy as np, torch
from make_prg_b200 import device, synth, hostio
ctx=device.Context(0)
mats=[synth.config_msa(3,i) for i in range(2000)]
packed=[hostio.pack_rows(m) for m in mats]
def run_packed(n):
    flat=np.concatenate([p.reshape(-1) for p,_ in packed[:n]])
    pin=torch.from_numpy(flat).pin_memory().numpy()
    offs=np.cumsum([0]+[p.size for p,_ in packed[:n-1]])
    torch.cuda.synchronize()
    t=time.perf_counter(); b,r=ctx.build_packed(pin,offs,[m.shape[0] for m in mats[:n]],[m.shape[1] for m in mats[:n]],[f for _,f in packed[:n]],5,7); dt=time.perf_counter()-t
    print(f'build_packed n={n}: {dt*1e3:.1f} ms', file=sys.stderr, flush=True); r.free(); b.free()
def run(n):
    b=ctx.upload(mats[:n]); torch.cuda.synchronize()
    t=time.perf_counter(); r=ctx.build(b,5,7); dt=time.perf_counter()-t
    print(f'build n={n}: {dt*1e3:.1f} ms', file=sys.stderr, flush=True); r.free(); b.free()
for n in (200,2000,2000):
    run_packed(n)
for n in (2000,2000):
    run(n)
run_packed(2000)
