"""-m gpu parity: kernel (a) column scan and the interval partition, through the C ABI, against
the oracle and the golden vectors generated from the reference (tests/golden/units.json)."""
import numpy as np
import pytest

import make_prg_oracle as mo
from helpers import rows_to_matrix, unit_cases
from make_prg_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from make_prg_b200 import device

    return device.Context(0)


def _packed_sym(row_bytes, col):
    """include/mprg.h: column c of a 32-column chunk is nibble c // 4 of 32-bit word c % 4."""
    c = col & 31
    byte = (col >> 5) * 16 + (c & 3) * 4 + (c >> 3)
    return (int(row_bytes[byte]) >> (((c >> 2) & 1) * 4)) & 15


def _iv_list(arr):
    return [[int(a["start"]), int(a["stop"]), int(a["type"])] for a in arr]


def test_pack_roundtrip(ctx):
    rng = np.random.default_rng(0)
    alphabet = np.frombuffer(b"ACGT-RYKMSWNacgtXZ", np.uint8)
    mats = [alphabet[rng.integers(0, len(alphabet), (r, c))] for r, c in [(3, 1), (5, 31), (4, 32), (7, 33), (2, 100)]]
    batch = ctx.upload(mats)
    lut = {ch: i for i, ch in enumerate(b"-AMCSGWTNR?Y?K??") if ch != ord("?")}
    for l, M in enumerate(mats):
        P = batch.packed(l)
        for r in range(M.shape[0]):
            for c in range(M.shape[1]):
                ch = bytes([M[r, c]]).upper()[0]
                want = lut.get(ch, 15)
                assert _packed_sym(P[r], c) == want
        # padding nibbles
        for c in range(M.shape[1], P.shape[1] * 2):
            assert _packed_sym(P[0], c) == 15
    fl = batch.flags()
    assert all(f & 1 for f in fl[1:])  # X/Z are disallowed somewhere in the bigger ones


def test_unit_vectors_consensus_reach_partition(ctx):
    cases = [r for r in unit_cases() if r["all"] is not None and "N" not in "".join(r["rows"])]
    mats = [rows_to_matrix(r["rows"]) for r in cases]
    batch = ctx.upload(mats)
    tasks = [(i, None, 0, m.shape[1]) for i, m in enumerate(mats)]
    scans = ctx.scan_tasks(batch, tasks)
    for rec, (cons, reach) in zip(cases, scans):
        assert cons.decode() == rec["consensus"]
        for a, b, ans in rec["has_empty"]:
            assert bool(reach[a] >= b) == ans
    for L in sorted({r["L"] for r in cases}):
        idx = [i for i, r in enumerate(cases) if r["L"] == L]
        parts = ctx.partition_tasks(batch, [tasks[i] for i in idx], L)
        for i, iv in zip(idx, parts):
            assert _iv_list(iv) == cases[i]["all"], (cases[i]["rows"], L)


def test_windows_and_row_subsets_against_oracle(ctx):
    rng = np.random.default_rng(5)
    mats = [synth.synth_msa(40, 300, 77, var_frac=0.1, n_dels=6),
            synth.synth_msa(64, 1500, 78, var_frac=0.05, n_dels=8),
            synth.synth_msa(9, 70, 79, var_frac=0.3, n_dels=4)]
    # long shared gap runs crossing 1024-column blocks
    mats[1][5:20, 900:1300] = ord("-")
    mats[1][30, 0:1500] = ord("-")
    batch = ctx.upload(mats)
    tasks = []
    for l, M in enumerate(mats):
        R, Ccols = M.shape
        tasks.append((l, None, 0, Ccols))
        for _ in range(25):
            c0 = int(rng.integers(0, Ccols))
            c1 = int(rng.integers(c0 + 1, Ccols + 1))
            k = int(rng.integers(1, R + 1))
            rows = np.sort(rng.choice(R, k, replace=False))
            tasks.append((l, rows, c0, c1))
    scans = ctx.scan_tasks(batch, tasks)
    for (l, rows, c0, c1), (cons, reach) in zip(tasks, scans):
        S = mats[l][:, c0:c1] if rows is None else mats[l][rows, c0:c1]
        assert cons == mo.consensus(S).tobytes()
        assert np.array_equal(reach, mo.gap_reach(S))
    for L in (1, 3, 7, 11):
        parts = ctx.partition_tasks(batch, tasks, L)
        for (l, rows, c0, c1), iv in zip(tasks, parts):
            S = mats[l][:, c0:c1] if rows is None else mats[l][rows, c0:c1]
            want = [[s, e, t] for s, e, t in mo.partition(mo.consensus(S), L, S)[2]]
            assert _iv_list(iv) == want


def test_partition_consensus_reference_vectors(ctx):
    # tests/from_msa/test_interval_partition.py:80-136 of the reference (empty alignment)
    def run(cons, L):
        iv = ctx.partition_consensus(cons, L)
        m = [[int(a["start"]), int(a["stop"])] for a in iv if a["type"] == 0]
        n = [[int(a["start"]), int(a["stop"])] for a in iv if a["type"] == 1]
        return m, n

    assert run("ATATAAA", 3) == ([[0, 6]], [])
    assert run("*******", 3) == ([], [[0, 6]])
    assert run("TTATT**AAAC*", 3) == ([[0, 4], [7, 10]], [[5, 6], [11, 11]])
    assert run("**AT*AAA", 3) == ([[5, 7]], [[0, 4]])
    assert run("TTATT**AA", 3) == ([[0, 4]], [[5, 8]])
    assert run("TT", 5) == ([[0, 1]], [])
    assert run("T*", 5) == ([], [[0, 1]])
    assert run("", 5) == ([], [])


def test_empty_and_degenerate_tasks(ctx):
    mats = [np.frombuffer(b"ACGT", np.uint8).reshape(1, 4).copy(),
            np.full((3, 5), ord("-"), np.uint8)]
    batch = ctx.upload(mats)
    scans = ctx.scan_tasks(batch, [(0, None, 0, 4), (1, None, 0, 5), (0, None, 2, 2)])
    assert scans[0][0] == b"ACGT"
    assert scans[1][0] == b"*****"
    assert list(scans[1][1]) == [4, 4, 4, 4, 4]
    assert scans[2][0] == b""
    parts = ctx.partition_tasks(batch, [(0, None, 0, 4), (1, None, 0, 5)], 3)
    assert _iv_list(parts[0]) == [[0, 3, 0]]
    assert _iv_list(parts[1]) == [[0, 4, 0]]  # all rows spell "" => demoted to a match interval
