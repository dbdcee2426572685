"""Alignment formats other than FASTA (-f clustal / stockholm / phylip*): the parsers of
make_prg_b200/utils/alignment_formats.py against hand-written files in each format's published layout, and the
loader's case / N handling on top of them (make_prg/utils/io_utils.py:22-49 takes any Bio.AlignIO format)."""
import io

import pytest

from make_prg_b200.utils import alignment_formats as af
from make_prg_b200.utils.io_utils import load_alignment_file

ROWS = [("seq_one", "ACGT-ACGTTACGGA-TTACA"), ("seq_two", "ACGTTACGTTAC--ACTTACA"), ("s3", "ACGTnACGTTACGGACTTACA")]

CLUSTAL = """CLUSTAL W (1.83) multiple sequence alignment


seq_one         ACGT-ACGTTAC 11
seq_two         ACGTTACGTTAC 12
s3              ACGTnACGTTAC 12
                ****  ******

seq_one         GGA-TTACA 19
seq_two         --ACTTACA 19
s3              GGACTTACA 21
                    *****
"""

STOCKHOLM = """# STOCKHOLM 1.0
#=GF ID toy
seq_one   ACGT-ACGTTAC
seq_two   ACGTTACGTTAC
#=GR seq_two SS ............
s3        ACGTnACGTTAC

seq_one   GGA-TTACA
seq_two   --ACTTACA
s3        GGACTTACA
//
"""

PHYLIP = """ 3 21
seq_one   ACGT-ACGTT ACGGA
seq_two   ACGTTACGTT AC--A
s3        ACGTnACGTT ACGGA

-TTACA
CTTACA
CTTACA
"""

PHYLIP_SEQ = """3 21
seq_one   ACGT-ACGTTAC
GGA-TTACA
seq_two   ACGTTACGTTAC--ACTTACA
s3        ACGTnACGTT
ACGGACTTAC
A
"""

PHYLIP_RELAXED = """3 21
a_long_identifier_1 ACGT-ACGTTACGGA-TTACA
a_long_identifier_2 ACGTTACGTTAC--ACTTACA
s3 ACGTnACGTTACGGACTTACA
"""


@pytest.mark.parametrize("fmt,text", [("clustal", CLUSTAL), ("stockholm", STOCKHOLM), ("phylip", PHYLIP),
                                      ("phylip-sequential", PHYLIP_SEQ)])
def test_parsers_read_interleaved_and_sequential_layouts(fmt, text):
    assert af.read_records(io.StringIO(text), fmt) == ROWS


def test_relaxed_phylip_keeps_long_identifiers():
    recs = af.read_records(io.StringIO(PHYLIP_RELAXED), "phylip-relaxed")
    assert [r[0] for r in recs] == ["a_long_identifier_1", "a_long_identifier_2", "s3"]
    assert [r[1] for r in recs] == [r[1] for r in ROWS]


def test_gz_and_path_input(tmp_path):
    import gzip

    p = tmp_path / "x.aln.gz"
    with gzip.open(p, "wt") as fh:
        fh.write(CLUSTAL)
    assert af.read_records(p, "clustal") == ROWS


@pytest.mark.parametrize("fmt,text,msg", [
    ("clustal", "seq_one ACGT\n", "not a Clustal file"),
    ("clustal", CLUSTAL.replace("GGA-TTACA 19", "GGA-TTACA 18"), "invalid sequence number"),
    ("clustal", CLUSTAL.replace("s3              GGACTTACA 21", "s3              GGACTTAC"), "same length"),
    ("stockholm", "seq ACGT\n//\n", "STOCKHOLM header"),
    ("phylip", "3 x\n", "two integers"),
    ("phylip", PHYLIP.replace(" 3 21", " 3 22"), "header says"),
    ("nexus", CLUSTAL, "not supported"),
])
def test_malformed_input_raises(fmt, text, msg):
    with pytest.raises(af.AlignmentFormatError, match=msg):
        af.read_records(io.StringIO(text), fmt)


def test_loader_applies_case_and_n_rules_like_fasta(tmp_path):
    fasta = io.StringIO(af.to_fasta_text(ROWS))
    want = [(r.id, r.seq) for r in load_alignment_file(fasta, "fasta")]
    p = tmp_path / "toy.sto"
    p.write_text(STOCKHOLM)
    got = [(r.id, r.seq) for r in load_alignment_file(str(p), "stockholm")]
    assert got == want
    assert got[2][1][4] in "ACGT"  # the N of s3 took the column's majority symbol
