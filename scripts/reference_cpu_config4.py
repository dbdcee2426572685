"""CPU: SURVEY 8(d) -- the UNMODIFIED reference on BASELINE config #4 (one 10,000 x 20,000 deep-clade locus, -N 10
-L 7, `make_prg from_msa -t 1`: one locus = one worker process) under a 30-minute cap; "> cap" when it does not
finish.  When it finishes, the sha256 of its PRG is compared with tests/golden/config4_deep_oracle.json (the oracle
port's, which the B200 build reproduces).  -> profiles/r2_reference_cpu_config4.json

    python scripts/reference_cpu_config4.py [cap_seconds]
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
CAP = int(sys.argv[1]) if len(sys.argv) > 1 else 1800

CHILD = r"""
import os, sys
os.environ["OMP_NUM_THREADS"] = "1"
sys.path.insert(0, sys.argv[1] + "/oracle")
import run_reference as rr
rr.ref_cli(["from_msa", "-i", sys.argv[2], "-o", sys.argv[3], "-N", 10, "-L", 7, "-t", 1, "-F"])
"""


def main():
    sys.path.insert(0, str(REPO))
    from make_prg_b200 import synth

    golden = json.loads((REPO / "tests" / "golden" / "config4_deep_oracle.json").read_text())
    tmp = Path(tempfile.mkdtemp(prefix="mprg_ref4_"))
    (tmp / "msas").mkdir()
    M = synth.config_msa(4, 0)
    assert hashlib.sha256(M.tobytes()).hexdigest() == golden["msa_sha256"]
    (tmp / "msas" / "locus4.fa").write_text(synth.to_fasta(M))
    del M
    t0 = time.perf_counter()
    rec = {"config": 4, "what": golden["generator"], "rows": golden["rows"], "cols": golden["cols"], "N": 10, "L": 7,
           "impl": "unmodified reference, make_prg from_msa -t 1 (oracle/run_reference.py harness)", "cores": 1,
           "cap_s": CAP, "oracle_port_seconds_1core": golden["oracle_seconds_1core"]}
    try:
        # own session: the reference's worker pool is a grandchild, the cap must end the whole group
        proc = subprocess.Popen([sys.executable, "-c", CHILD, str(REPO), str(tmp / "msas"), str(tmp / "out")],
                                stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, start_new_session=True)
        try:
            if proc.wait(timeout=CAP) != 0:
                raise subprocess.CalledProcessError(proc.returncode, "make_prg from_msa")
        except subprocess.TimeoutExpired:
            os.killpg(proc.pid, 9)
            proc.wait()
            raise
        rec["wall_s"] = time.perf_counter() - t0
        prg = (tmp / "out.prg.fa").read_text().split("\n")[1]
        rec["prg_len"] = len(prg)
        rec["prg_sha256_equals_oracle_port"] = hashlib.sha256(prg.encode()).hexdigest() == golden["prg_sha256"]
    except subprocess.TimeoutExpired:
        rec["wall_s"] = f"> cap ({CAP} s)"
    except subprocess.CalledProcessError as err:
        rec["wall_s"] = None
        rec["error"] = f"reference exited with {err.returncode} after {time.perf_counter() - t0:.0f} s"
    finally:
        import shutil

        shutil.rmtree(tmp, ignore_errors=True)
    (REPO / "profiles" / "r2_reference_cpu_config4.json").write_text(json.dumps(rec) + "\n")
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
