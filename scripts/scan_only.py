"""GPU box: repeatedly run the root-level column scan (kernel (a)) alone on the bench workload and
report achieved algorithmic GB/s from CUDA events; used under ncu for the --set full capture."""
import os, sys, time
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import numpy as np, torch
import bench
from make_prg_b200 import device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
base = bench.workload(0, min(n, 1000))
data = np.concatenate([base] * (n // 1000)) if n > 1000 else base
ctx = device.Context(0)
batch = ctx.upload((data.reshape(-1), [(bench.ROWS, bench.COLS)] * n))
tasks = [(i, None, 0, bench.COLS) for i in range(n)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
arr, arena = ctx.make_tasks(batch, tasks)
ctx.scan_log(reset=True)
for r in range(reps):
    flush.zero_(); torch.cuda.synchronize()
    ctx.partition_tasks(batch, tasks, 7)
by, ms = ctx.scan_log()
print("scan launches", len(by), "bytes/launch", by[0], "ms", np.round(ms, 4).tolist())
print("achieved GB/s (median of last)", float(by[0] / (np.median(ms[2:]) * 1e-3) / 1e9))
