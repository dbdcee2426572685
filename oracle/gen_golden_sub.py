"""Golden vectors for NodeFactory.build(alignment, prg_builder, parent_node) -- the re-build of an updated
leaf below its parent (make_prg/recursion_tree.py:353-388, 401-471) -- produced by the UNMODIFIED
reference under the harness of run_reference.py.  Build container only (needs /root/reference):

    python oracle/gen_golden_sub.py        ->  tests/golden/sub_builds.json
"""
import json
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(REPO))

import gen_golden  # noqa: E402
import run_reference as rr  # noqa: E402
from make_prg_b200 import synth  # noqa: E402


class Parent:
    """What NodeFactory.build reads from parent_node: its nesting level."""

    def __init__(self, nesting_level):
        self.nesting_level = nesting_level
        self.node_id = 0


def ref_sub_build(M, max_nesting, L, parent_level, first_node_id):
    rr.load_reference()
    from make_prg.prg_builder import PrgBuilder
    from make_prg.recursion_tree import NodeFactory
    from make_prg.utils.io_utils import load_alignment_file

    with tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False) as fh:
        fh.write(synth.to_fasta(M))
        path = fh.name
    alignment = load_alignment_file(path, "fasta")
    Path(path).unlink()
    builder = PrgBuilder.__new__(PrgBuilder)
    builder._locus_name = "sub"
    builder.max_nesting, builder.min_match_length = max_nesting, L
    builder.aligner = None
    builder.next_node_id = first_node_id
    builder.site_num = 5
    builder.prg_index = {}
    node = NodeFactory.build(alignment, builder, Parent(parent_level))
    parts = []
    node.preorder_traversal_to_build_prg(parts)
    tree = []

    def walk(n):
        tree.append((type(n).__name__, n.node_id, n.nesting_level, len(n.alignment),
                     n.alignment.get_alignment_length(), len(n.children)))
        for c in n.children:
            walk(c)

    walk(node)
    return "".join(parts), tree, builder.next_node_id


def main():
    rng = np.random.default_rng(777)
    out = []
    mats = [gen_golden.random_small_msa(rng) for _ in range(60)]
    mats = [M for M in mats if M.shape[1] > 0 and not (M == ord("N")).any()]
    mats += [synth.synth_msa(40, 160, 900 + i, var_frac=0.12, n_dels=3) for i in range(6)]
    mats += [synth.synth_msa(30, 25, 950 + i, var_frac=0.5, n_dels=1) for i in range(6)]  # one non-match interval
    for k, M in enumerate(mats):
        for parent_level in (0, 2, 4):
            L = int(rng.choice([1, 3, 7]))
            first = int(rng.integers(1, 50))
            try:
                prg, tree, next_id = ref_sub_build(M, 5, L, parent_level, first)
            except Exception as err:  # SequenceCurationError etc.: not a vector
                print("skip", k, parent_level, type(err).__name__)
                continue
            out.append({"rows": [bytes(r).decode() for r in M], "N": 5, "L": L, "parent_level": parent_level,
                        "first_node_id": first, "prg": prg, "tree": tree, "next_node_id": next_id})
    kinds = {}
    for rec in out:
        kinds[rec["tree"][0][0]] = kinds.get(rec["tree"][0][0], 0) + 1
    print(len(out), "vectors; root kinds", kinds)
    with open(REPO / "tests" / "golden" / "sub_builds.json", "w") as fh:
        json.dump(out, fh)


if __name__ == "__main__":
    main()
