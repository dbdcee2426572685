"""CPU (build container or any host with the staged reference): SURVEY 8(d) "CPU reference timing" for the configs
bench.py does not time -- the UNMODIFIED reference (`make_prg from_msa -t <cores>` under the Biopython stand-in of
oracle/run_reference.py, KMeans n_init=10, OMP_NUM_THREADS=1) on a seeded subsample of config #3 (500 rows x
U[600,1400] columns) and of config #5 (25 % variable columns, L = 3 / 7 / 15), with the core count; the PRGs of the
first loci are compared with the oracle port's.  One JSON line per run -> profiles/r2_reference_cpu_configs.jsonl.

    python scripts/reference_cpu_configs.py [loci_config3] [loci_config5]
"""
import json
import os
import shutil
import sys
import tempfile
import time
from pathlib import Path

os.environ.setdefault("OMP_NUM_THREADS", "1")
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "oracle"))

import make_prg_oracle as mo  # noqa: E402
import run_reference as rr  # noqa: E402
from make_prg_b200 import synth  # noqa: E402


def run(config, n_loci, N, L, cores, full_size):
    tmp = Path(tempfile.mkdtemp(prefix="mprg_refcfg_"))
    try:
        (tmp / "msas").mkdir()
        mats = [synth.config_msa(config, i) for i in range(n_loci)]
        for i, M in enumerate(mats):
            (tmp / "msas" / f"locus{i:05d}.fa").write_text(synth.to_fasta(M))
        saved = os.dup(1)
        os.dup2(2, 1)
        t0 = time.perf_counter()
        try:
            rr.ref_cli(["from_msa", "-i", tmp / "msas", "-o", tmp / "out", "-N", N, "-L", L, "-t", cores, "-F"])
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
        wall = time.perf_counter() - t0
        lines = (tmp / "out.prg.fa").read_text().split("\n")
        prgs = {lines[i][1:]: lines[i + 1] for i in range(0, len(lines) - 1, 2)}
        n_chk = min(4, n_loci)
        equal = all(prgs[f"locus{i:05d}"] ==
                    mo.build_prg_from_matrix([f"s{r}" for r in range(mats[i].shape[0])], mats[i], N, L)[0]
                    for i in range(n_chk))
        cols = sum(M.shape[1] for M in mats)
        return {"config": config, "impl": "unmodified reference, make_prg from_msa -t <cores> (oracle/run_reference.py "
                                          "harness: Biopython stand-in, KMeans n_init=10, OMP_NUM_THREADS=1)",
                "N": N, "L": L, "loci": n_loci, "cores": cores, "wall_s": wall, "loci_per_s": n_loci / wall,
                "columns_per_s": cols / wall, "loci_per_s_per_core": n_loci / wall / cores,
                "full_config_loci": full_size,
                "full_config_extrapolated_s_on_these_cores": full_size * wall / n_loci,
                "prgs_equal_oracle_port": {"checked": n_chk, "equal": bool(equal)}}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    n3 = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    n5 = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    cores = os.cpu_count() or 1
    if not rr.reference_available():
        raise SystemExit("no reference source (neither /root/reference nor oracle/_ref)")
    out = REPO / "profiles" / "r2_reference_cpu_configs.jsonl"
    with open(out, "w") as fh:
        for rec in ([run(3, n3, 5, 7, cores, 25_000)] + [run(5, n5, 5, L, cores, 200) for L in (3, 7, 15)]):
            print(json.dumps(rec))
            fh.write(json.dumps(rec) + "\n")
            fh.flush()


if __name__ == "__main__":
    main()
