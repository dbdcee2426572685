"""-m gpu parity of the whole path (mprg_build through the C ABI): byte-identical PRG strings and
identical recursion trees against the reference's golden outputs and reference runs on synthetic MSAs."""
import hashlib

import numpy as np
import pytest

import make_prg_oracle as mo
from helpers import (REF, SMALL_CASES, prg_spells_all_rows, synthetic_cases, truth_multi,
                     truth_prg)
from make_prg_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from make_prg_b200 import device

    return device.Context(0)


def _build(ctx, mats, N, L):
    batch = ctx.upload(mats)
    res = ctx.build(batch, N, L)
    return batch, res


def _tree_dump(res, locus, M):
    t = res.nodes(locus)
    names = {0: "LeafNode", 1: "MultiIntervalNode", 2: "MultiClusterNode"}
    out = []
    for i in range(len(t["kind"])):
        rows = (np.arange(M.shape[0]) if t["row_off"][i] < 0
                else t["row_pool"][t["row_off"][i]:t["row_off"][i] + t["n_rows"][i]])
        S = M[rows, t["c0"][i]:t["c1"][i]]
        keep = int((~(S == ord("-")).all(axis=0)).sum())
        out.append([names[int(t["kind"][i])], i, int(t["nesting_level"][i]), len(rows), keep,
                    int(t["n_children"][i])])
    return out


def test_small_cases(ctx):
    for L in sorted(set(SMALL_CASES.values())):
        names = [c for c, l in SMALL_CASES.items() if l == L]
        mats = [mo.load_msa(REF / f"{c}.fa")[1] for c in names]
        _, res = _build(ctx, mats, 5, L)
        for i, c in enumerate(names):
            assert res.status(i) == 0
            assert res.prg(i) == truth_prg(c), c


def test_disallowed_base_marks_locus_only(ctx):
    mats = [mo.load_msa(REF / "match.fa")[1]]
    ids, bad = mo.parse_fasta((REF / "fails_2.fa").read_text())
    mats.append(np.frombuffer("".join(bad).upper().encode(), np.uint8).reshape(len(bad), -1))
    _, res = _build(ctx, mats, 5, 7)
    assert res.status(0) == 0 and res.prg(0) == truth_prg("match")
    assert res.status(1) == 1 and res.prg(1) == ""


def test_sample_example_and_amira(ctx):
    for setname, files in (("sample_example", ["GC00006032.fa", "GC00010897.fa"]),
                           ("amira_MSAs", ["alsB.fasta.gz", "glpG.fasta.gz", "group_18516.fasta.gz"])):
        mats = [mo.load_msa(REF / setname / f)[1] for f in files]
        _, res = _build(ctx, mats, 5, 7)
        truth = truth_multi(setname)
        for i, f in enumerate(files):
            name = f.split(".")[0]
            assert res.prg(i) == truth[name], name
            assert prg_spells_all_rows(res.prg(i), mats[i])


def test_synthetic_prg_and_tree(ctx):
    recs = synthetic_cases()
    groups = {}
    for r in recs:
        groups.setdefault((r["N"], r["L"]), []).append(r)
    for (N, L), rs in groups.items():
        mats = [synth.config_msa(r["config"], r["index"], r.get("rows"), r.get("cols")) for r in rs]
        _, res = _build(ctx, mats, N, L)
        for i, r in enumerate(rs):
            assert res.prg(i) == r["prg"], (r["config"], r["index"], N, L)
            assert res.n_nodes(i) == r["n_nodes"] and res.n_sites(i) == r["n_sites"]
            assert _tree_dump(res, i, mats[i]) == [list(t) for t in r["tree"]]


def test_build_below_a_parent_node(ctx):
    """mprg_build_sub against the unmodified reference's NodeFactory.build(alignment, builder, parent_node):
    PRG strings and trees of 216 re-builds below nodes of nesting level 0 / 2 / 4, one batch per (N, L)."""
    from helpers import rows_to_matrix, sub_build_cases

    recs = sub_build_cases()
    groups = {}
    for r in recs:
        groups.setdefault((r["N"], r["L"]), []).append(r)
    for (N, L), rs in groups.items():
        mats = [rows_to_matrix(r["rows"]) for r in rs]
        batch = ctx.upload(mats)
        res = ctx.build_sub(batch, N, L, [r["parent_level"] for r in rs])
        for i, r in enumerate(rs):
            assert res.status(i) == 0
            assert res.prg(i) == r["prg"], (r["rows"], L, r["parent_level"])
            want = [[t[0], t[1] - r["first_node_id"]] + list(t[2:]) for t in r["tree"]]
            assert _tree_dump(res, i, mats[i]) == want
        # the same loci as roots differ wherever the reference would force a MultiIntervalNode
        roots = ctx.build_sub(batch, N, L, [-1] * len(rs))
        plain = ctx.build(batch, N, L)
        for i in range(len(rs)):
            assert roots.prg(i) == plain.prg(i)


def test_batch_order_invariance(ctx):
    mats = [synth.config_msa(2, i) for i in range(4)]
    _, a = _build(ctx, mats, 5, 7)
    _, b = _build(ctx, mats[::-1], 5, 7)
    for i in range(4):
        assert a.prg(i) == b.prg(3 - i)


def test_build_ascii_equals_upload_then_build(ctx):
    """mprg_build_ascii (every worker range copied, packed and built on its own stream) returns exactly
    what mprg_batch_upload + mprg_build return, whatever the number of workers, including a locus with
    a disallowed base and loci of different shapes."""
    mats = [synth.synth_msa(30 + (i % 7) * 9, 120 + (i % 5) * 77, 900 + i, var_frac=0.06, n_dels=3) for i in range(40)]
    bad = mats[11].copy()
    bad[3, 17] = ord("Z")
    mats[11] = bad
    _, want = _build(ctx, mats, 5, 7)
    for workers in (1, 3, 8):
        ctx.set_workers(workers)
        batch, got = ctx.build_ascii(mats, 5, 7)
        for i in range(len(mats)):
            assert got.status(i) == want.status(i), (workers, i)
            assert got.prg(i) == want.prg(i), (workers, i)
        assert got.status(11) == 1
        assert batch.flags()[11] & 1
        got.free()
        batch.free()
    ctx.set_workers(4)


def test_build_ascii_from_pinned_memory_equals_pageable(ctx, monkeypatch):
    """Pinned host buffers copied and packed, pinned buffers read in place by the pack kernel
    (MPRG_ZEROCOPY=1) and pageable buffers: same packed rows, flags and PRGs for odd widths, unaligned
    locus offsets, widths around the 512-column staging span and a locus with a disallowed base."""
    import torch

    widths = [1, 7, 31, 32, 33, 511, 512, 513, 1000, 1023, 1025, 1531]
    mats = [synth.synth_msa(3 + (i * 5) % 17, w, 700 + i, var_frac=0.1, n_dels=2) for i, w in enumerate(widths)]
    mats[4][1, 3] = ord("Z")
    mats[6][2, 500:512] = np.frombuffer(b"RYKMSWRYKMSW", np.uint8)
    flat = np.concatenate([np.zeros(3, np.uint8)] + [m.reshape(-1) for m in mats])  # every offset is odd-ish
    shapes = [m.shape for m in mats]
    pinned = torch.from_numpy(flat.copy()).pin_memory().numpy()[3:]
    pageable = flat[3:].copy()
    results = []
    for buf, zero_copy in ((pinned, True), (pageable, False), (pinned, False)):
        if zero_copy:
            monkeypatch.setenv("MPRG_ZEROCOPY", "1")
        else:
            monkeypatch.delenv("MPRG_ZEROCOPY", raising=False)
        batch, res = ctx.build_ascii((buf, shapes), 5, 7)
        results.append((batch.flags().tolist(), [batch.packed(i).tobytes() for i in range(len(mats))],
                        [res.status(i) for i in range(len(mats))], [res.prg(i) for i in range(len(mats))]))
        res.free()
        batch.free()
    assert results[0] == results[1] == results[2]
    assert results[0][2][4] == 1 and all(st == 0 for i, st in enumerate(results[0][2]) if i != 4)
    for i, M in enumerate(mats):
        if i != 4:
            want, _ = mo.build_prg_from_matrix([f"s{r}" for r in range(M.shape[0])], M, 5, 7)
            assert results[0][3][i] == want, widths[i]


def test_deep_locus_takes_the_whole_grid_paths(ctx):
    """One deep locus of the FLAT generator (every row distinct through private SNPs, but all within 20 % of
    the majority string) sends a single huge clustering problem through the whole-grid de-duplication, k-mer
    numbering and one-reference-like check with ONE cluster; the loop of cluster_sequences.py:256-274 ends
    there, so no KMeans runs (the deep-clade loci of tests/test_gpu_deep.py do reach it).  The PRG must still
    be the oracle's, byte for byte."""
    M = synth.synth_msa(1500, 3000, 4_100_000, n_haps=300, var_frac=0.04, private_snp=0.01)
    want, _ = mo.build_prg_from_matrix([f"s{r}" for r in range(M.shape[0])], M, 10, 7)
    _, res = _build(ctx, [M], 10, 7)
    assert res.status(0) == 0
    assert hashlib.sha256(res.prg(0).encode()).hexdigest() == hashlib.sha256(want.encode()).hexdigest()


def test_repeated_builds_of_a_large_batch_are_identical(ctx):
    """The device-resident loop bump-allocates its arenas with atomics: WHERE an object lands differs from run to
    run, WHAT is built must not (round 2 found per-task scratch regions that could overlap when two allocation
    counters were bumped in different orders: 3 wrong loci in 25,000).  3,000 loci x 4 builds, PRG for PRG, plus
    the oracle on a sample."""
    mats = [synth.config_msa(3, i, rows=120, cols=400 + (i * 37) % 300) for i in range(3000)]
    batch = ctx.upload(mats)
    first = None
    for rep in range(4):
        res = ctx.build(batch, 5, 7)
        prgs = [res.prg(i) for i in range(len(mats))]
        assert all(res.status(i) == 0 for i in range(0, len(mats), 101))
        res.free()
        if first is None:
            first = prgs
        else:
            assert prgs == first, rep
    for i in range(0, len(mats), 500):
        want, _ = mo.build_prg_from_matrix([f"s{r}" for r in range(mats[i].shape[0])], mats[i], 5, 7)
        assert first[i] == want, i
    batch.free()
