"""GPU box, under ncu: warm-up + ONE profiled build of round 1's config-#4 locus (10,000 x 20,000, flat generator:
whole-grid dedupe / k-mer numbering / one-reference-like check, no KMeans)."""
import sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import torch
from make_prg_b200 import device, synth
M = synth.config_msa("4flat", 0)
ctx = device.Context(0)
batch = ctx.upload([M])
res = ctx.build(batch, 10, 7); res.free()
torch.cuda.synchronize()
torch.cuda.profiler.start()
res = ctx.build(batch, 10, 7)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", res.status(0), len(res.prg(0)))
