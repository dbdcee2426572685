"""Shared helpers for the test-suite (golden fixture access, seeded synthetic inputs)."""
import gzip
import json
import re
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
GOLDEN = REPO / "tests" / "golden"
REF = GOLDEN / "ref"

# case -> min_match_length used by the reference's integration tests
# (tests/integration_tests/test_from_msa.py:35-220 in the reference)
SMALL_CASES = {
    "match": 7, "match.nonmatch": 7, "match.nonmatch.match": 7, "match.nonmatch.shortmatch": 7,
    "match.staggereddash": 7, "nonmatch": 7, "nonmatch.match": 7, "nonmatch.shortmatch": 7,
    "shortmatch.nonmatch": 7, "shortmatch.nonmatch.match": 7, "contains_n": 7,
    "contains_n_and_RYKMSW": 7, "contains_n_no_variants": 7, "contains_RYKMSW": 7,
    "a_column_full_of_Ns": 7, "nested_snps_seq_backgrounds": 3,
    "nested_snps_seq_backgrounds_more_seqs": 3, "nested_snps_deletion": 1,
}


def truth_prg(case):
    return (REF / "truth" / case / f"{case}.prg.fa").read_text().split("\n")[1]


def truth_multi(setname):
    lines = (REF / "truth" / setname / f"{setname}.prg.fa").read_text().split("\n")
    return {lines[i][1:]: lines[i + 1] for i in range(0, len(lines) - 1, 2)}


def locus_name(path):
    return re.sub(r"\.(fa|fasta)(\.gz)?$", "", Path(path).name)


def synthetic_cases():
    with open(GOLDEN / "synthetic.json") as fh:
        return json.load(fh)


def sub_build_cases():
    """NodeFactory.build(alignment, builder, parent_node) vectors from the unmodified reference
    (oracle/gen_golden_sub.py)."""
    with open(GOLDEN / "sub_builds.json") as fh:
        return json.load(fh)


def unit_cases():
    with open(GOLDEN / "units.json") as fh:
        return json.load(fh)


def kmeans_cases():
    z = np.load(GOLDEN / "kmeans_cases.npz")
    for i in range(int(z["count"])):
        yield (z[f"X{i}"].astype(np.float64), int(z[f"K{i}"]), z[f"labels{i}"],
               float(z[f"inertia{i}"]))


def rows_to_matrix(rows):
    if not rows or len(rows[0]) == 0:
        return np.zeros((len(rows), 0), np.uint8)
    return np.frombuffer("".join(rows).encode(), np.uint8).reshape(len(rows), -1).copy()


def prg_to_regex(prg):
    """A PRG string as a regular expression over ACGT: site ` s a1 s+1 a2 ... s ` -> (?:a1|a2|...).
    Every path through the PRG is a match; used for the size-independent property
    'each input sequence is spelled by the PRG'."""
    import re

    tokens = re.findall(r" \d+ |[ACGT]+", prg)
    out = []
    stack = []  # open odd markers
    for tok in tokens:
        if tok[0] == " ":
            m = int(tok)
            if m % 2 == 1:
                if stack and stack[-1] == m:
                    stack.pop()
                    out.append(")")
                else:
                    stack.append(m)
                    out.append("(?:")
            else:
                out.append("|")
        else:
            out.append(tok)
    assert not stack, "unbalanced PRG"
    return "".join(out)


def prg_spells_all_rows(prg, M):
    import re

    pat = re.compile(prg_to_regex(prg))
    seen = set()
    for row in M:
        s = bytes(row[row != ord("-")]).decode()
        if s in seen:
            continue
        seen.add(s)
        if not pat.fullmatch(s):
            return False
    return True
