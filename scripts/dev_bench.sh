python bench.py --steps 10 --warmup 3 --no-files --no-big > gpurun_out/r2d_bench_dev.json 2> gpurun_out/r2d_bench_dev.err
MPRG_HOST_LOOP=1 python bench.py --steps 10 --warmup 3 --no-files --no-big > gpurun_out/r2d_bench_host.json 2> gpurun_out/r2d_bench_host.err
python - <<'P' > gpurun_out/r2d_trace.txt 2>&1
import os, sys, time
os.environ['MPRG_TRACE']='1'
sys.path.insert(0,'.')
import numpy as np, bench
from make_prg_b200 import device
data=bench.workload(0,1000)
ctx=device.Context(0)
shapes=[(200,1000)]*1000
b=ctx.upload((data.reshape(-1),shapes))
for i in range(3):
    t=time.perf_counter(); r=ctx.build(b,5,7); dt=time.perf_counter()-t; print('build wall ms',dt*1e3, file=sys.stderr); r.free()
P
