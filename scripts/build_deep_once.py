"""GPU box, under ncu: one warm-up build of a reduced deep-clade locus (1,500 x 4,000, config #4 class), then
cudaProfilerStart and ONE build.  python scripts/build_deep_once.py [rows cols]"""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import torch
from make_prg_b200 import device, synth
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
M = synth.synth_deep_msa(rows, cols, 4_000_000, n_clades=8, n_haps=max(16, rows // 5))
ctx = device.Context(0)
batch = ctx.upload([M])
res = ctx.build(batch, 10, 7); res.free()
torch.cuda.synchronize()
torch.cuda.profiler.start()
res = ctx.build(batch, 10, 7)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", res.status(0), len(res.prg(0)))
