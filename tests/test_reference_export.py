"""CPU: the table-shaped update_DS archive (written on a GPU box by scripts/make_ds_fixture.py, committed as
tests/golden/update_DS_fixture.zip) loads into PrgBuilder objects, and make_prg_b200.utils.reference_export turns
it into pickles of the REFERENCE's own classes that the unmodified reference loads and finds equal
(PrgBuilder.__eq__, prg_builder.py:56-85) to what it builds itself from the same MSA -- run under the Biopython
stand-in of oracle/refshim (the real Biopython is not installed in this image)."""
import pickle
import sys
import zipfile
from pathlib import Path

import pytest

from helpers import GOLDEN, REF

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "oracle"))

FIXTURE = GOLDEN / "update_DS_fixture.zip"
INPUTS = {"GC00006032": REF / "sample_example" / "GC00006032.fa", "GC00010897": REF / "sample_example" / "GC00010897.fa",
          "nested_snps_deletion": REF / "nested_snps_deletion.fa", "contains_RYKMSW": REF / "contains_RYKMSW.fa"}


def fixture_prgs():
    lines = (GOLDEN / "update_DS_fixture.prg.fa").read_text().split("\n")
    return {lines[i][1:]: lines[i + 1] for i in range(0, len(lines) - 1, 2)}


def test_table_archive_loads_into_builders():
    from make_prg_b200.prg_builder import DS_MAGIC, PrgBuilderZipDatabase, parse_ds_record
    from make_prg_b200.utils import io_utils

    db = PrgBuilderZipDatabase(FIXTURE)
    db.load()
    assert db.get_loci_names() == sorted(INPUTS)
    prgs = fixture_prgs()
    for locus in db.get_loci_names():
        blob = db._zip_file.read(locus)
        assert blob[:8] == DS_MAGIC
        rec = parse_ds_record(blob)
        want = io_utils.load_alignment_file(str(INPUTS[locus]))
        assert (rec["matrix"] == want.matrix).all() and rec["titles"] == [r.description for r in want]
        b = db.get_PrgBuilder(locus)
        assert b.build_prg() == prgs[locus] == rec["prg"]
        assert b.next_node_id == rec["n_nodes"] and (b.site_num - 5) // 2 == rec["n_sites"]
        assert len(b.prg_index) > 0 and all(isinstance(k, tuple) for k in b.prg_index)
        assert [r.id for r in b.root.alignment] == want.ids
        assert pickle.loads(pickle.dumps(b, protocol=4)) == b
    db.close()


def test_export_to_reference_pickles(tmp_path):
    import run_reference as rr

    if not rr.reference_available():
        pytest.skip("reference sources not available (neither /root/reference nor oracle/_ref)")
    rr.load_reference()
    from make_prg.prg_builder import PrgBuilderZipDatabase as RefDatabase

    from make_prg_b200.utils.reference_export import export_update_ds

    out = tmp_path / "ref.update_DS.zip"
    assert export_update_ds(FIXTURE, out) == len(INPUTS)
    db = RefDatabase(out)
    db.load()
    assert db.get_loci_names() == sorted(INPUTS)
    prgs = fixture_prgs()
    for locus in db.get_loci_names():
        loaded = db.get_PrgBuilder(locus)  # the reference's own deserialisation
        assert type(loaded).__module__ == "make_prg.prg_builder"
        assert loaded.build_prg() == prgs[locus]
        # what the unmodified reference builds from the same file
        built, prg = rr.ref_build(INPUTS[locus], 5, 7, locus_name=locus)
        assert prg == prgs[locus]
        assert loaded == built and built == loaded
        assert sorted(loaded.prg_index) == sorted(built.prg_index)
    db.close()
