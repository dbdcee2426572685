"""Pins the oracle restatement (oracle/make_prg_oracle.py) to the reference:
its own golden outputs (tests/golden/ref/truth, copied from the reference's test data) and the
outputs of the unmodified reference on seeded synthetic MSAs (tests/golden/synthetic.json)."""
import hashlib

import pytest

import make_prg_oracle as mo
from helpers import (REF, SMALL_CASES, locus_name, rows_to_matrix, sub_build_cases, synthetic_cases, truth_multi,
                     truth_prg)
from make_prg_b200 import synth


@pytest.mark.parametrize("case", sorted(SMALL_CASES))
def test_small_cases_prg_gfa_bin(case):
    prg, _ = mo.build_prg_from_file(REF / f"{case}.fa", 5, SMALL_CASES[case])
    assert prg == truth_prg(case)
    assert mo.prg_gfa(prg) == (REF / "truth" / case / f"{case}.prg.gfa").read_text()
    assert mo.prg_bin(prg) == (REF / "truth" / case / f"{case}.prg.bin").read_bytes()


def test_disallowed_base_skips_locus():
    # tests/integration_tests/test_from_msa.py:183-187 (fails_2.fa => no output for the locus)
    with pytest.raises(mo.SequenceCurationError):
        mo.build_prg_from_file(REF / "fails_2.fa", 5, 7)


def test_empty_msa():
    with pytest.raises(ValueError, match="No records found in handle"):
        mo.build_prg_from_file(REF / "several_empty" / "empty.fa", 5, 7)


@pytest.mark.parametrize("name", ["GC00006032", "GC00010897"])
def test_sample_example(name):
    prg, _ = mo.build_prg_from_file(REF / "sample_example" / f"{name}.fa", 5, 7)
    assert prg == truth_multi("sample_example")[name]


@pytest.mark.parametrize("name", ["glpG", "group_18516", "alsB"])
def test_amira(name):
    prg, _ = mo.build_prg_from_file(REF / "amira_MSAs" / f"{name}.fasta.gz", 5, 7)
    assert prg == truth_multi("amira_MSAs")[name]


def test_gz_input_equals_plain():
    a, _ = mo.build_prg_from_file(REF / "match.fa", 5, 7)
    b, _ = mo.build_prg_from_file(REF / "match.fa.gz", 5, 7)
    assert a == b


@pytest.mark.parametrize("rec", synthetic_cases(),
                         ids=lambda r: f"c{r['config']}i{r['index']}N{r['N']}L{r['L']}")
def test_synthetic_against_reference_run(rec):
    M = synth.config_msa(rec["config"], rec["index"], rec.get("rows"), rec.get("cols"))
    assert hashlib.sha256(M.tobytes()).hexdigest() == rec["msa_sha256"], "generator drifted"
    ids = [f"s{i}" for i in range(M.shape[0])]
    prg, b = mo.build_prg_from_matrix(ids, M, rec["N"], rec["L"])
    assert prg == rec["prg"]
    assert [list(t) for t in b.dump_tree()] == [list(t) for t in rec["tree"]]
    assert hashlib.sha256(mo.prg_gfa(prg).encode()).hexdigest() == rec["gfa_sha256"]
    assert hashlib.sha256(mo.prg_bin(prg)).hexdigest() == rec["bin_sha256"]


def test_build_below_a_parent_node_against_reference_run():
    """NodeFactory.build(alignment, builder, parent_node) (recursion_tree.py:431-432): the re-build of an
    updated leaf is not a root -- no forced MultiIntervalNode, nesting starts at the parent's level."""
    recs = sub_build_cases()
    assert len(recs) >= 200 and {r["tree"][0][0] for r in recs} == {"LeafNode", "MultiIntervalNode", "MultiClusterNode"}
    for r in recs:
        M = rows_to_matrix(r["rows"])
        b = mo.OracleBuilder([f"s{i}" for i in range(M.shape[0])], M, r["N"], r["L"],
                             parent_level=r["parent_level"], first_node_id=r["first_node_id"])
        assert b.build_prg() == r["prg"]
        assert [list(t) for t in b.dump_tree()] == [list(t) for t in r["tree"]]
        assert b.next_node_id == r["next_node_id"]
