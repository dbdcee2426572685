import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from make_prg_b200 import device
ctx = device.Context(0)
rng = np.random.default_rng(1)
for X, K in [(rng.integers(0, 4, (40, 300)).astype(float), 3), (np.eye(3), 2)]:
    a = ctx.kmeans(X, K, mode=1)
    b = ctx.kmeans(X, K, mode=2)
    print(X.shape, K, "single", a[1], "group", b[1], "same labels", bool((a[0] == b[0]).all()), flush=True)
