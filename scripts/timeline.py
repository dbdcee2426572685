"""GPU box: warm kernel timeline of one bench-size build through torch.profiler (CUPTI): per-kernel time inside a
real step (not ncu's cold serialised launches), GPU busy time and the idle gaps between kernels."""
import os, sys, json, collections
from pathlib import Path
os.environ.setdefault("MPRG_WORKERS", "1")
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from make_prg_b200 import device
ctx = device.Context(0)
data = bench.workload(0, 1000)
batch = ctx.upload((data.reshape(-1), [(bench.ROWS, bench.COLS)] * 1000))
for _ in range(3):
    ctx.build(batch, 5, 7).free()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    res = ctx.build(batch, 5, 7)
    torch.cuda.synchronize()
out = REPO / "gpurun_out" / "timeline_trace.json"
prof.export_chrome_trace(str(out))
ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
busy = 0.0; cur_end = t0; gaps = []
for e in ev:
    if e["ts"] > cur_end:
        gaps.append((e["ts"] - cur_end, e["name"][:60]))
    busy += max(0.0, e["ts"] + e["dur"] - max(cur_end, e["ts"]))
    cur_end = max(cur_end, e["ts"] + e["dur"])
tot = collections.defaultdict(float); cnt = collections.Counter()
for e in ev:
    n = e["name"].split("(")[0].replace("void ", "").replace("mprg::", "")
    tot[n] += e["dur"]; cnt[n] += 1
print(f"span {t1 - t0:.1f} us, busy {busy:.1f} us, idle {t1 - t0 - busy:.1f} us, events {len(ev)}")
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:28]:
    print(f"{v:9.1f} us  n={cnt[k]:3d}  {k}")
gaps.sort(reverse=True)
print("largest gaps (us, next event):")
for g in gaps[:25]:
    print(f"  {g[0]:8.1f}  {g[1]}")
print("gaps > 5 us:", sum(1 for g in gaps if g[0] > 5), "sum", sum(g[0] for g in gaps if g[0] > 5))
