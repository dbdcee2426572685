"""GPU box with >= 2 GPUs: `from_msa --gpus 2` (one process per GPU, LPT shards) writes the same files as
`--gpus 1`, and both equal the reference's truth files."""
import filecmp, shutil, sys, tempfile, zipfile
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO)); sys.path.insert(0, str(REPO / "tests"))
from helpers import REF, truth_multi
from make_prg_b200.__main__ import main


def check():
    tmp = Path(tempfile.mkdtemp())
    src = tmp / "msas"; src.mkdir()
    for d in ("amira_MSAs", "sample_example"):
        for f in (REF / d).iterdir():
            if f.is_file():
                shutil.copy(f, src / f.name)
    for c in ("match.nonmatch.match", "nested_snps_deletion", "contains_n_and_RYKMSW", "fails_2"):
        shutil.copy(REF / f"{c}.fa", src / f"{c}.fa")
    for g in (1, 2):
        main(["from_msa", "-i", str(src), "-o", str(tmp / f"g{g}"), "--gpus", str(g)])
    assert (tmp / "g1.prg.fa").read_bytes() == (tmp / "g2.prg.fa").read_bytes()
    for kind in ("bin", "gfa", ):
        with zipfile.ZipFile(tmp / f"g1.prg.{kind}.zip") as a, zipfile.ZipFile(tmp / f"g2.prg.{kind}.zip") as b:
            assert sorted(a.namelist()) == sorted(b.namelist())
            for m in a.namelist():
                assert a.read(m) == b.read(m), m
    from make_prg_b200.prg_builder import PrgBuilderZipDatabase
    d1, d2 = PrgBuilderZipDatabase(tmp / "g1.update_DS.zip"), PrgBuilderZipDatabase(tmp / "g2.update_DS.zip")
    d1.load(), d2.load()
    names = d2.get_loci_names()
    assert names == d1.get_loci_names()
    truth = {**truth_multi("amira_MSAs"), **truth_multi("sample_example")}
    with zipfile.ZipFile(tmp / "g1.update_DS.zip") as a, zipfile.ZipFile(tmp / "g2.update_DS.zip") as b:
        for n in names:
            assert a.read(n) == b.read(n), n  # the table-shaped records are byte-identical
    for n in names:
        if n in truth and n in ("GC00006032", "GC00010897", "glpG"):
            assert d2.get_PrgBuilder(n).build_prg() == truth[n], n
    d1.close(), d2.close()
    lines = (tmp / "g2.prg.fa").read_text().split("\n")
    got = {lines[i][1:]: lines[i + 1] for i in range(0, len(lines) - 1, 2)}
    for n, prg in truth.items():
        assert got[n] == prg, n
    print("cli --gpus 2 == --gpus 1 == truth:", len(got), "loci,", len(names), "update_DS records")


if __name__ == "__main__":  # the shards are spawned processes: they re-import this module
    check()
