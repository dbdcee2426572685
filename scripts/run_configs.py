"""GPU box: BASELINE.json configs #3 (sub-sample), #4 (reduced and full) and #5 through the public API,
with parity against the oracle on a bounded sample and the size-independent property on more loci.
Writes one JSON object per config to stdout (and gpurun_out/configs_r1.json)."""
import json, os, sys, time
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "oracle")); sys.path.insert(0, str(REPO / "tests"))
import numpy as np
import make_prg_oracle as mo
from helpers import prg_spells_all_rows
from make_prg_b200 import device, synth

which = set(sys.argv[1].split(",")) if len(sys.argv) > 1 else {"3", "5", "4r", "4"}
N3 = int(os.environ.get("N3", "2000"))
ctx = device.Context(0)
out = []


def gpu_build(mats, N, L, reps=3):
    """(results, best resident ms, best end-to-end ms)"""
    best_res, best_e2e, prgs, stats = 1e30, 1e30, None, None
    for _ in range(reps):
        t0 = time.perf_counter()
        batch = ctx.upload(mats)
        t1 = time.perf_counter()
        res = ctx.build(batch, N, L)
        t2 = time.perf_counter()
        prgs = [res.prg(i) for i in range(len(mats))]
        stats = [res.status(i) for i in range(len(mats))]
        t3 = time.perf_counter()
        res.free(); batch.free()
        best_res = min(best_res, 1e3 * (t2 - t1)); best_e2e = min(best_e2e, 1e3 * (t3 - t0))
    return prgs, stats, best_res, best_e2e


def check(mats, prgs, N, L, n_oracle, n_prop):
    t0 = time.perf_counter()
    bad = 0
    for i in range(min(n_oracle, len(mats))):
        want, _ = mo.build_prg_from_matrix([f"s{r}" for r in range(mats[i].shape[0])], mats[i], N, L)
        bad += want != prgs[i]
    dt = time.perf_counter() - t0
    prop_bad = sum(0 if prg_spells_all_rows(prgs[i], mats[i]) else 1 for i in range(min(n_prop, len(mats))))
    return bad, dt, prop_bad


def emit(d):
    out.append(d)
    print(json.dumps(d), flush=True)


if "3" in which:
    t0 = time.perf_counter()
    mats = [synth.config_msa(3, i) for i in range(N3)]
    gen = time.perf_counter() - t0
    prgs, stats, ms_res, ms_e2e = gpu_build(mats, 5, 7)
    n_or = 8
    bad, dt, prop_bad = check(mats, prgs, 5, 7, n_or, 64)
    cols = sum(m.shape[1] for m in mats)
    emit({"config": 3, "what": f"{N3}-locus sub-sample of config #3 (500 rows x U[600,1400] cols, seeds 2,000,000+i), -N 5 -L 7",
          "loci": N3, "ok": sum(s == 0 for s in stats), "resident_ms": ms_res, "e2e_ms": ms_e2e,
          "loci_per_s_resident": N3 / ms_res * 1e3, "loci_per_s_e2e": N3 / ms_e2e * 1e3,
          "columns_per_s_e2e": cols / ms_e2e * 1e3, "oracle_checked": n_or, "oracle_mismatch": bad,
          "oracle_s_per_locus_1core": dt / n_or, "property_checked": 64, "property_fail": prop_bad, "gen_s": gen})
    del mats

if "5" in which:
    mats = [synth.config_msa(5, i) for i in range(200)]
    for L in (3, 5, 7, 9, 11, 13, 15):
        prgs, stats, ms_res, ms_e2e = gpu_build(mats, 5, L)
        n_or = 4
        bad, dt, prop_bad = check(mats, prgs, 5, L, n_or, 32)
        emit({"config": 5, "what": "200 loci x 200 x 1000, 25% variable columns (seeds 3,000,000+i), -N 5",
              "L": L, "loci": 200, "ok": sum(s == 0 for s in stats), "resident_ms": ms_res, "e2e_ms": ms_e2e,
              "loci_per_s_resident": 200 / ms_res * 1e3, "loci_per_s_e2e": 200 / ms_e2e * 1e3,
              "oracle_checked": n_or, "oracle_mismatch": bad, "oracle_s_per_locus_1core": dt / n_or,
              "property_checked": 32, "property_fail": prop_bad})
    del mats

# config #4: ONE deep-clade locus (8 clades at 30 % divergence, 1 % private SNPs): the clustering loop runs KMeans
# on every distinct row (n = rows, F = 4^7 = 16,384 k-mers).  "4r" = reduced (oracle run on the box),
# "4" = BASELINE size (sha256 of the oracle's PRG from the build container, tests/golden/config4_deep_oracle.json),
# "4flat" = round 1's generator (no deep clades: the loop ends before KMeans).
import hashlib
for tag, rows, cols in (("4r", 1500, 4000), ("4", 10000, 20000), ("4flat", 10000, 20000)):
    if tag not in which:
        continue
    t0 = time.perf_counter()
    if tag == "4flat":
        M = synth.config_msa("4flat", 0)
    elif tag == "4r":
        M = synth.synth_deep_msa(rows, cols, 4_000_000, n_clades=8, n_haps=300)
    else:
        M = synth.config_msa(4, 0)
    gen = time.perf_counter() - t0
    try:
        ctx.path_counts(reset=True)
        prgs, stats, ms_res, ms_e2e = gpu_build([M], 10, 7, reps=2)
        d = {"config": tag, "what": f"one deep locus {rows} x {cols}, -N 10 -L 7", "ok": int(stats[0] == 0),
             "resident_ms": ms_res, "e2e_ms": ms_e2e, "prg_len": len(prgs[0]), "gen_s": gen,
             "paths": ctx.path_counts(reset=True)}
        d["prg_sha256"] = hashlib.sha256(prgs[0].encode()).hexdigest()
        gname = {"4r": "deep.json", "4": "config4_deep_oracle.json", "4flat": "config4_oracle.json"}[tag]
        gold = json.loads((REPO / "tests" / "golden" / gname).read_text())
        if tag == "4r":
            gold = gold["deep_1500"]
            d["reference_s"] = gold["reference_seconds"]
        else:
            d["oracle_s"] = gold["oracle_seconds_1core"]
        d["oracle_mismatch"] = int(gold["prg_sha256"] != d["prg_sha256"])
    except Exception as e:  # report, do not hide
        d = {"config": tag, "error": repr(e)[:500], "gen_s": gen}
    emit(d)

(REPO / "gpurun_out").mkdir(exist_ok=True)
with open(REPO / "gpurun_out" / "configs_r2.json", "a") as fh:
    for d in out:
        fh.write(json.dumps(d) + "\n")
