"""GPU box: warm kernel timeline of one bench-size build through torch.profiler (CUPTI): per-kernel time inside a
real step (not ncu's cold serialised launches), GPU busy time and the idle gaps between kernels."""
import os, sys, json, collections
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from make_prg_b200 import device
import numpy as np
E2E = len(sys.argv) > 1 and sys.argv[1] in ("e2e", "lanes_e2e")   # from pinned packed host rows (mprg_build_packed), as bench.py's e2e
LANES = len(sys.argv) > 1 and sys.argv[1] in ("lanes", "lanes_e2e")  # several builds in flight (device.BuildPipeline), as bench.py's value / e2e
N_BUILDS = 18  # builds inside the profiled region in the lanes modes
if not E2E:
    os.environ.setdefault("MPRG_WORKERS", "1")
ctx = device.Context(0)
data = bench.workload(0, 1000)
pipe = device.BuildPipeline(0) if LANES else None
if E2E:
    from make_prg_b200 import hostio
    n, R, Cc = 1000, bench.ROWS, bench.COLS
    stride = hostio.packed_stride(Cc)
    packed_np = torch.empty(n * R * stride, dtype=torch.uint8).pin_memory().numpy()
    flags = np.zeros(n, np.int32)
    for i in range(n):
        rows, flags[i] = hostio.pack_rows(data[i])
        packed_np[i * R * stride:(i + 1) * R * stride] = rows.reshape(-1)
    offs = np.arange(n, dtype=np.int64) * (R * stride)
    nr, nc = np.full(n, R, np.int32), np.full(n, Cc, np.int32)
    def run():
        if LANES:
            futs = [pipe.submit_packed(packed_np, offs, nr, nc, flags, 5, 7, consume=lambda b, r: r.statuses()[0].sum())
                    for _ in range(N_BUILDS)]
            [f.result() for f in futs]
            return
        b, r = ctx.build_packed(packed_np, offs, nr, nc, flags, 5, 7)
        r.free(); b.free()
else:
    batch = ctx.upload((data.reshape(-1), [(bench.ROWS, bench.COLS)] * 1000))
    batches = [batch] + ([ctx.upload((data.reshape(-1), [(bench.ROWS, bench.COLS)] * 1000)) for _ in range(pipe.depth - 1)]
                         if LANES else [])
    def run():
        if LANES:
            futs = []
            for _ in range(N_BUILDS):
                futs.append(pipe.submit_resident(batches[pipe.next_lane], 5, 7, consume=lambda b, r: r.statuses()[0].sum()))
            [f.result() for f in futs]
            return
        ctx.build(batch, 5, 7).free()
for _ in range(3):
    run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    run()
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
if E2E:
    h2d = [e for e in ev if "HtoD" in e["name"] and e["dur"] > 50]
    print("big H2D copies:", [(round(e["ts"] - t0), round(e["dur"]), e.get("args", {}).get("stream")) for e in h2d])
    streams = collections.defaultdict(lambda: [1e18, 0])
    for e in ev:
        st = e.get("args", {}).get("stream")
        streams[st][0] = min(streams[st][0], e["ts"] - t0); streams[st][1] = max(streams[st][1], e["ts"] + e["dur"] - t0)
    print("streams (first, last us):", {k: (round(v[0]), round(v[1])) for k, v in streams.items()})
print(f"span {t1 - t0:.1f} us, busy {busy:.1f} us, idle {t1 - t0 - busy:.1f} us, events {len(ev)}")
if LANES:
    ksum = sum(e["dur"] for e in ev if e.get("cat") == "kernel")
    # how many kernels are in flight, time-weighted over the busy time
    pts = sorted([(e["ts"], 1) for e in ev if e.get("cat") == "kernel"] + [(e["ts"] + e["dur"], -1) for e in ev if e.get("cat") == "kernel"])
    depth_time = collections.defaultdict(float); cur = 0; last = pts[0][0]
    for t, d in pts:
        depth_time[cur] += t - last; last = t; cur += d
    tot_t = sum(v for k, v in depth_time.items())
    print(f"{N_BUILDS} builds in flight on {pipe.depth} lanes: {(t1 - t0) / N_BUILDS:.1f} us per build, GPU busy {100 * busy / (t1 - t0):.1f} % of the span, "
          f"sum of kernel durations {ksum:.1f} us = {ksum / N_BUILDS:.1f} us per build")
    print("kernels in flight (share of the span):", {k: round(100 * v / tot_t, 1) for k, v in sorted(depth_time.items())})
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:28]:
    print(f"{v:9.1f} us  n={cnt[k]:3d}  {k}")
gaps.sort(reverse=True)
print("largest gaps (us, next event):")
for g in gaps[:25]:
    print(f"  {g[0]:8.1f}  {g[1]}")
print("gaps > 5 us:", sum(1 for g in gaps if g[0] > 5), "sum", sum(g[0] for g in gaps if g[0] > 5))
