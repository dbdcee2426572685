"""GPU box: mprg_build (resident) and mprg_build_ascii (end to end) wall time on the bench workload for the
worker / range settings given in the environment (MPRG_WORKERS, MPRG_RANGES_PER_WORKER, MPRG_DYNAMIC_RANGES)."""
import os, sys, time
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import numpy as np, torch
import bench
from make_prg_b200 import device
n = 1000
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
data = bench.workload(0, n)
host = torch.from_numpy(data.reshape(-1)).pin_memory().numpy()
ctx = device.Context(0)
shapes = [(bench.ROWS, bench.COLS)] * n
batch = ctx.upload((host, shapes))
tr, te = [], []
for it in range(steps + 3):
    t0 = time.perf_counter()
    res = ctx.build(batch, 5, 7)
    tr.append(time.perf_counter() - t0)
    res.free()
for it in range(steps + 3):
    t0 = time.perf_counter()
    b, res = ctx.build_ascii((host, shapes), 5, 7)
    te.append(time.perf_counter() - t0)
    res.free(); b.free()
tr, te = np.array(tr[3:]) * 1e3, np.array(te[3:]) * 1e3
tag = " ".join(f"{k[5:]}={os.environ[k]}" for k in ("MPRG_WORKERS", "MPRG_RANGES_PER_WORKER", "MPRG_DYNAMIC_RANGES") if k in os.environ)
print(f"[{tag or 'default'}] resident ms min {tr.min():.2f} med {np.median(tr):.2f} mean {tr.mean():.2f} max {tr.max():.2f} | "
      f"ascii ms min {te.min():.2f} med {np.median(te):.2f} mean {te.mean():.2f} max {te.max():.2f}", flush=True)
