"""GPU box: the one-reference-like check of a deep locus (cluster_sequences.py:59-104) as the ONE-HOT INT8 GEMM that
SURVEY 8(d) and north_star describe, on the tensor pipe through the library (torch._int_mm -> cuBLASLt int8 GEMM),
next to the bit-sliced SIMT kernels of csrc/refcheck_grid.cu that read the 4-bit packed rows.

Hamming(row r, majority of cluster k) = w - onehot(row r) . onehot(majority k):  M = rows, N = clusters (padded to
16), K = columns x sigma with sigma = 8 one-hot int8 values per symbol (5 symbols - A C G T, padded).  The one-hot
operand is 16 x the bytes of the packed rows.  Reported: time to EXPAND the packed rows to one-hot int8 (what a fused
kernel would have to do into shared memory), time of the int8 GEMM alone (operands already expanded in HBM), and the
equality of the distances with a numpy reference on a sample.  Comparison numbers of the SIMT kernels:
profiles/r2_refcheck_grid_variants.txt (majority 44.8 us, Hamming 25.4 us for the same 10,000 x 20,000 locus)."""
import json, sys, time
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import numpy as np, torch
from make_prg_b200 import synth

rows, cols, K = 10_000, 20_000, 8
M = synth.config_msa("4flat", 0, rows, cols)           # uint8 ASCII [rows, cols]
lut = np.full(256, 0, np.uint8)
for i, ch in enumerate(b"-ACGT"):
    lut[ch] = i
codes = torch.from_numpy(lut[M]).cuda()                # 0..4 per symbol
labels = torch.from_numpy(np.arange(rows) % K).cuda()  # any clustering: the arithmetic is what is timed
dev = codes.device

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev])) * 1e3  # us

# majority string per cluster (any exact method; here by one-hot sums through the same GEMM shape)
SIG = 8
def expand(c):  # [r, cols] codes -> [r, cols * SIG] int8 one-hot
    return torch.nn.functional.one_hot(c.long(), SIG).to(torch.int8).reshape(c.shape[0], -1)
t_expand = timed(lambda: expand(codes[:2000]), reps=3) * (rows / 2000)
A = torch.empty(rows, cols * SIG, dtype=torch.int8, device=dev)
for r0 in range(0, rows, 1000):
    A[r0:r0 + 1000] = expand(codes[r0:r0 + 1000])
onehot_cluster = torch.zeros(rows, 16, dtype=torch.int8, device=dev)
onehot_cluster[torch.arange(rows, device=dev), labels] = 1
# counts[cluster][col * SIG + sym] = onehot(cluster)^T @ onehot(symbol): M = 16, N = cols * SIG, K = rows
t_counts = timed(lambda: torch._int_mm(onehot_cluster.t().contiguous().repeat(2, 1)[:32], A))
counts = torch._int_mm(onehot_cluster.t().contiguous().repeat(2, 1)[:32], A)[:K].reshape(K, cols, SIG)
maj = counts.argmax(dim=2)                              # [K, cols] (ties: lowest code; timing only)
B = torch.zeros(cols * SIG, 16, dtype=torch.int8, device=dev)
B.reshape(cols, SIG, 16)[torch.arange(cols, device=dev)[:, None], maj.t(), torch.arange(K, device=dev)[None, :]] = 1
t_gemm = timed(lambda: torch._int_mm(A, B))
matches = torch._int_mm(A, B)                           # [rows, 16] int32
ham = cols - matches[torch.arange(rows, device=dev), labels]
# check on a sample against numpy
c_np, maj_np, lab_np = codes.cpu().numpy(), maj.cpu().numpy(), labels.cpu().numpy()
ok = all(int((c_np[r] != maj_np[lab_np[r]]).sum()) == int(ham[r]) for r in range(0, rows, 997))
packed_bytes = rows * cols / 2
line = {"what": "one-reference-like check of a 10,000 x 20,000 locus as a one-hot int8 GEMM on the tensor pipe (torch._int_mm)",
        "onehot_operand_bytes": rows * cols * SIG, "packed_rows_bytes": packed_bytes,
        "expand_to_onehot_us (torch ops, HBM to HBM, extrapolated from 2,000 rows)": t_expand,
        "int8_gemm_counts_us (M=32, N=160000, K=10000)": t_counts,
        "int8_gemm_hamming_us (M=10000, N=16, K=160000)": t_gemm,
        "gemm_hamming_algorithmic_GBs_on_packed_bytes": packed_bytes / (t_gemm * 1e-6) / 1e9,
        "distances_equal_numpy_on_sample": bool(ok),
        "simt_kernels_us": {"majority_count_kernel": 44.8, "hamming_packed_kernel": 25.4,
                            "source": "profiles/r2_refcheck_grid_variants.txt"}}
print(json.dumps(line))
(REPO / "gpurun_out").mkdir(exist_ok=True)
(REPO / "gpurun_out" / "r2_onehot_gemm_compare.json").write_text(json.dumps(line) + "\n")
