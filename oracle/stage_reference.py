"""Stages the UNMODIFIED reference package for the GPU box: /root/reference/make_prg -> oracle/_ref/make_prg.

TEST / BASELINE INFRASTRUCTURE ONLY -- nothing under make_prg_b200/ imports it.

`/root/reference` does not exist on the GPU box, but `oracle/_ref/` travels there with the repo snapshot
(git-ignored, not gpurun-ignored).  The staged copy is what `bench.py --impl reference` times through the
reference's own CLI entry (`make_prg.__main__:main`, `from_msa -t <all host cores>`) under the harness of
oracle/run_reference.py (Biopython stand-in oracle/refshim, KMeans forced to n_init=10 = scikit-learn 1.3.0
behaviour, one OpenMP thread per worker).  Only the Python sources are staged: the pre-built MAFFT binaries
(30 MB, `update` only) are left out.  Nothing staged is tracked by git.
"""
import shutil
import sys
from pathlib import Path

SRC = Path("/root/reference/make_prg")
DST = Path(__file__).resolve().parent / "_ref" / "make_prg"


def stage(force=False) -> bool:
    """Returns True when oracle/_ref/make_prg exists afterwards."""
    if not SRC.exists():
        return DST.exists()
    if DST.exists() and not force:
        newest_src = max(p.stat().st_mtime for p in SRC.rglob("*.py"))
        newest_dst = max((p.stat().st_mtime for p in DST.rglob("*.py")), default=0)
        if newest_dst >= newest_src:
            return True
    if DST.exists():
        shutil.rmtree(DST)
    DST.parent.mkdir(parents=True, exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("mafft-*", "__pycache__", "*.pyc"))
    return True


if __name__ == "__main__":
    ok = stage(force="--force" in sys.argv)
    print(f"{DST}: {'staged' if ok else 'reference source not available'}")
