"""Debug helper (GPU box): locate the first tree difference between mprg_build and the oracle."""
import sys
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "oracle")); sys.path.insert(0, str(REPO / "tests"))
import make_prg_oracle as mo, kmeans13
from make_prg_b200 import device
from helpers import REF

path = sys.argv[1] if len(sys.argv) > 1 else str(REF / "amira_MSAs" / "alsB.fasta.gz")
N, L = 5, 7
ids, M = mo.load_msa(path)
ctx = device.Context(0)
batch = ctx.upload([M])
res = ctx.build(batch, N, L)
t = res.nodes(0)
trace = []
prg, ob = mo.build_prg_from_matrix(ids, M, N, L, trace=trace)
print("prg equal", prg == res.prg(0), "nodes", ob.next_node_id, res.n_nodes(0))
onodes = []
def walk(n):
    onodes.append(n)
    for c in n.children: walk(c)
walk(ob.root)
kinds = {"leaf": 0, "interval": 1, "cluster": 2}
for i, n in enumerate(onodes):
    if i >= len(t["kind"]): print("gpu tree shorter"); break
    rows = np.arange(M.shape[0]) if t["row_off"][i] < 0 else t["row_pool"][t["row_off"][i]:t["row_off"][i]+t["n_rows"][i]]
    same = (kinds[n.kind] == t["kind"][i] and n.c0 == t["c0"][i] and n.c1 == t["c1"][i] and len(n.rows) == len(rows) and np.array_equal(np.asarray(n.rows), rows) and n.nesting_level == t["nesting_level"][i])
    if not same:
        print("first diff at preorder", i, "oracle", n.kind, n.c0, n.c1, len(n.rows), n.nesting_level, "gpu", t["kind"][i], t["c0"][i], t["c1"][i], len(rows), t["nesting_level"][i])
        p = n.parent
        print("parent", p.kind, p.c0, p.c1, len(p.rows), p.nesting_level)
        # re-run the parent's clustering on both sides
        S = M[np.asarray(p.rows), p.c0:p.c1]
        pids = [ids[r] for r in p.rows]
        tr = []
        oc, nocl = mo.kmeans_cluster_seqs(S, pids, L, trace=tr)
        pos = {rid: k for k, rid in enumerate(pids)}
        print("oracle clusters", [sorted(pos[x] for x in c) for c in oc])
        gc = ctx.cluster_tasks(batch, [(0, np.asarray(p.rows, np.int32), p.c0, p.c1)], L)[0]
        print("gpu clusters   ", gc)
        Xg = ctx.kmer_counts(batch, (0, np.asarray(p.rows, np.int32), p.c0, p.c1), L)
        for X, K in tr:
            print("K", K, "X equal", Xg.shape == X.shape and np.array_equal(Xg, X))
            want, inertia, cen, fitlab = kmeans13.kmeans_fit_predict(X, K)
            got, gi = ctx.kmeans(X, K)
            print("   oracle", want.tolist(), repr(inertia)); print("   gpu   ", got.tolist(), repr(gi))
        break
else:
    print("trees identical")
