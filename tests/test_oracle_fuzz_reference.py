"""CPU: the oracle port against the UNMODIFIED reference on the seeded random alignments of the GPU fuzz test
(tests/test_gpu_fuzz.py compares mprg_build with the oracle on exactly these inputs, so the three agree pairwise):
one-row / one-column alignments, gap-placement twins, all-gap column blocks, N, RYKMSW, lower case, symbols outside
the alphabet.  The reference runs in-process from /root/reference (or the copy staged under oracle/_ref) under the
Biopython stand-in of oracle/run_reference.py; skipped where neither is present."""
import os

import numpy as np
import pytest

import make_prg_oracle as mo
import run_reference as rr
from test_gpu_fuzz import SETTINGS, oracle_outcome, random_msa

pytestmark = pytest.mark.skipif(not rr.reference_available(), reason="no reference source in this container")

PER_SETTING = 48  # the same 48 alignments per setting as the GPU test


def reference_outcome(path, N, L):
    rr.load_reference()
    from make_prg.utils.seq_utils import SequenceCurationError

    try:
        builder, prg = rr.ref_build(path, N, L, locus_name="fuzz")
    except SequenceCurationError:
        return ("curation",)
    return ("ok", prg, builder.next_node_id)


@pytest.mark.parametrize("setting", range(len(SETTINGS)))
def test_oracle_equals_unmodified_reference_on_fuzz_alignments(setting, tmp_path):
    N, L = SETTINGS[setting]
    # the GPU test's streams: MPRG_FUZZ_SEED=<int> draws another set there and here
    rng = np.random.default_rng(77_000 + setting + 1000 * int(os.environ.get("MPRG_FUZZ_SEED", "0")))
    done = n_ok = 0
    while done < PER_SETTING:
        text = random_msa(rng)
        M, out = oracle_outcome(text, N, L)
        if M is None:
            continue
        path = tmp_path / f"fuzz{done}.fa"
        path.write_text(text)
        ref = reference_outcome(path, N, L)
        assert ref == out, (setting, done, M.shape)
        n_ok += out[0] == "ok"
        done += 1
    assert n_ok >= PER_SETTING // 2
