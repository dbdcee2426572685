"""GPU box: config #3 built twice with different chunk sizes; loci whose PRG differs are checked against the oracle."""
import os, sys, time
from concurrent.futures import ProcessPoolExecutor
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "oracle")); sys.path.insert(0, str(REPO / "scripts"))
import numpy as np
import config3
N = int(sys.argv[1]) if len(sys.argv) > 1 else 25000
def main():
    import multiprocessing as mp
    with ProcessPoolExecutor(os.cpu_count(), mp_context=mp.get_context("spawn")) as pool:
        made = list(pool.map(config3.make_packed, range(N), chunksize=16))
    import torch
    from make_prg_b200 import device, synth
    import make_prg_oracle as mo
    ctx = device.Context(0)
    def build_all(chunk):
        out = {}
        for a in range(0, N, chunk):
            c = made[a:a + chunk]
            flat = np.concatenate([p.reshape(-1) for _i, p, _f, _s in c])
            offs = np.cumsum([0] + [p.size for _i, p, _f, _s in c[:-1]])
            b, r = ctx.build_packed(flat, offs, [s[0] for *_x, s in c], [s[1] for *_x, s in c], [f for _i, _p, f, _s in c], 5, 7)
            for k in range(len(c)):
                out[c[k][0]] = r.prg(k)
            r.free(); b.free()
        return out
    A = build_all(5000)
    B = build_all(int(os.environ.get("CHUNK_B", "1250")))
    C = build_all(5000)
    for rep in range(3):
        D = build_all(5000)
        print("repeat", rep, "differs from A in", sum(1 for i in range(N) if A[i] != D[i]), flush=True)
    diff = [i for i in range(N) if A[i] != B[i]]
    diff2 = [i for i in range(N) if A[i] != C[i]]
    print("differ A(5000) vs B:", len(diff), diff[:20], " A vs A again:", len(diff2), diff2[:20], flush=True)
    for i in (diff + diff2)[:6]:
        M = synth.config_msa(3, i)
        want = mo.build_prg_from_matrix([f"s{r}" for r in range(M.shape[0])], M, 5, 7)[0]
        print(i, M.shape, "A ok" if A[i] == want else "A WRONG", "B ok" if B[i] == want else "B WRONG", "C ok" if C[i] == want else "C WRONG", len(A[i]), len(B[i]), len(want), flush=True)
if __name__ == "__main__":
    main()
