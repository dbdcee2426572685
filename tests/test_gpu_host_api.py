"""-m gpu: the Python host layer that mirrors the reference's interface (PrgBuilder, node classes,
IntervalPartitioner, kmeans_cluster_seqs, writers, CLI) on top of the C ABI; the expectations are
the reference's own golden files and unit vectors, so these read like the reference's tests."""
import hashlib
import pickle
import zipfile
from argparse import Namespace
from pathlib import Path

import numpy as np
import pytest

from helpers import REF, SMALL_CASES, rows_to_matrix, truth_multi, truth_prg, unit_cases

pytestmark = pytest.mark.gpu


def test_prg_builder_matches_truth_and_engine_string():
    from make_prg_b200.prg_builder import PrgBuilder
    from make_prg_b200.recursion_tree import LeafNode, MultiClusterNode, MultiIntervalNode

    for case in ("match.nonmatch.match", "nested_snps_seq_backgrounds", "contains_n_and_RYKMSW",
                 "nested_snps_deletion", "match.staggereddash"):
        b = PrgBuilder(case, REF / f"{case}.fa", "fasta", 5, SMALL_CASES[case])
        prg = b.build_prg()
        assert prg == truth_prg(case) == b.engine_prg
        assert b.root.node_id == 0 and b.root.parent is None
        ids = []

        def walk(n):
            ids.append(n.node_id)
            assert isinstance(n, (LeafNode, MultiClusterNode, MultiIntervalNode))
            for c in n.children:
                assert c.parent is n
                walk(c)

        walk(b.root)
        assert ids == list(range(b.next_node_id))  # pre-order numbering
        clone = pickle.loads(pickle.dumps(b, protocol=4))
        assert clone == b and clone.build_prg() == prg


def test_sample_example_builder():
    from make_prg_b200.prg_builder import PrgBuilder

    truth = truth_multi("sample_example")
    for name in ("GC00006032", "GC00010897"):
        b = PrgBuilder(name, REF / "sample_example" / f"{name}.fa", "fasta", 5, 7)
        assert b.build_prg() == truth[name]
        # prg_index keys are character offsets of the leaf alleles (recursion_tree.py:278-300)
        for (s, e), leaf in b.prg_index.items():
            assert set(truth[name][s:e]) <= set("ACGT") and (s, e) in leaf.indexed_PRG_intervals


def test_disallowed_base_raises_curation_error():
    from make_prg_b200.prg_builder import PrgBuilder
    from make_prg_b200.utils.seq_utils import SequenceCurationError

    with pytest.raises(SequenceCurationError):
        PrgBuilder("fails_2", REF / "fails_2.fa", "fasta", 5, 7)


def test_interval_partitioner_reference_vectors():
    # tests/from_msa/test_interval_partition.py of the reference
    from make_prg_b200.from_msa.interval_partition import Interval, IntervalPartitioner, IntervalType
    from make_prg_b200.msa import MSA, SeqRecord

    M, N = IntervalType.Match, IntervalType.NonMatch
    empty = MSA([])
    m, n, a = IntervalPartitioner("TTATT**AAAC*", 3, empty).get_intervals()
    assert m == [Interval(M, 0, 4), Interval(M, 7, 10)] and n == [Interval(N, 5, 6), Interval(N, 11, 11)]
    assert a == sorted(m + n)
    m, n, _ = IntervalPartitioner("**AT*AAA", 3, empty).get_intervals()
    assert m == [Interval(M, 5, 7)] and n == [Interval(N, 0, 4)]
    m, n, _ = IntervalPartitioner("T*", 5, empty).get_intervals()
    assert m == [] and n == [Interval(N, 0, 1)]
    msa = MSA([SeqRecord("TTAAGGTTT-AATTTA", "s1"), SeqRecord("TTAAGGTTTTAATTTA", "s2")])
    m, n, _ = IntervalPartitioner("TTAAGGTTT*AATTTA", 7, msa).get_intervals()
    assert m == [Interval(M, 0, 7)] and n == [Interval(N, 8, 15)]


def test_seq_utils_mirrors_on_unit_vectors():
    from make_prg_b200.msa import MSA, SeqRecord
    from make_prg_b200.utils.seq_utils import get_consensus_from_MSA, has_empty_sequence

    for rec in unit_cases()[:60]:
        if "N" in "".join(rec["rows"]) or len(rec["rows"][0]) == 0:
            continue
        msa = MSA([SeqRecord(s, f"s{i}") for i, s in enumerate(rec["rows"])])
        assert get_consensus_from_MSA(msa) == rec["consensus"]
        a, b, ans = rec["has_empty"][0]
        assert has_empty_sequence(msa, (a, b)) == ans


def test_kmeans_cluster_seqs_mirror_exact_id_order():
    from make_prg_b200.from_msa.cluster_sequences import kmeans_cluster_seqs
    from make_prg_b200.msa import MSA, SeqRecord

    n = 0
    for rec in unit_cases():
        if rec["clustered_ids"] is None or "N" in "".join(rec["rows"]) or len(rec["rows"][0]) == 0:
            continue
        msa = MSA([SeqRecord(s, f"s{i}") for i, s in enumerate(rec["rows"])])
        res = kmeans_cluster_seqs(msa, rec["L"])
        assert res.clustered_ids == rec["clustered_ids"], (rec["rows"], rec["L"])
        n += 1
        if n >= 120:
            break


def _run_cli(tmp_path, input_path, name, **kw):
    from make_prg_b200.subcommands import from_msa
    from make_prg_b200.subcommands.output_type import OutputType

    opts = Namespace(input=str(input_path), suffix="", output_prefix=str(tmp_path / name),
                     alignment_format="fasta", max_nesting=5, min_match_length=7,
                     output_type=OutputType("a"), force=False, threads=1, verbose=False, log=None, gpus=1)
    for k, v in kw.items():
        setattr(opts, k, v)
    from_msa.run(opts)
    return opts


def test_cli_sample_example_outputs_equal_truth(tmp_path):
    _run_cli(tmp_path, REF / "sample_example", "sample_example")
    truth_dir = REF / "truth" / "sample_example"
    assert (tmp_path / "sample_example.prg.fa").read_bytes() == (truth_dir / "sample_example.prg.fa").read_bytes()
    for kind in ("bin", "gfa"):
        with zipfile.ZipFile(tmp_path / f"sample_example.prg.{kind}.zip") as got, \
                zipfile.ZipFile(truth_dir / f"sample_example.prg.{kind}.zip") as want:
            assert sorted(got.namelist()) == sorted(want.namelist())
            for member in want.namelist():
                assert got.read(member) == want.read(member), member
    # update_DS: table-shaped records written by the library, PrgBuilder objects built on load
    from make_prg_b200.prg_builder import DS_MAGIC, PrgBuilder, PrgBuilderZipDatabase

    with zipfile.ZipFile(tmp_path / "sample_example.update_DS.zip") as zf:
        assert sorted(zf.namelist()) == ["GC00006032", "GC00010897"]
        assert zf.read("GC00006032")[:8] == DS_MAGIC
    db = PrgBuilderZipDatabase(tmp_path / "sample_example.update_DS.zip")
    db.load()
    assert db.get_number_of_loci() == 2 and db.get_loci_names() == ["GC00006032", "GC00010897"]
    for locus in db.get_loci_names():
        b = db.get_PrgBuilder(locus)
        assert b.build_prg() == truth_multi("sample_example")[locus]
        # the same object as a build through the Python API of the locus' file
        direct = PrgBuilder(locus, REF / "sample_example" / f"{locus}.fa", "fasta", 5, 7)
        direct.build_prg()
        assert b == direct and b.prg_index.keys() == direct.prg_index.keys()
        assert [r.id for r in b.root.alignment] == [r.id for r in direct.root.alignment]
        assert pickle.loads(pickle.dumps(b, protocol=4)) == b
    assert db == db
    db.close()


def test_cli_pickled_update_ds_and_sharded_merge_of_archives(tmp_path, monkeypatch):
    """MPRG_PICKLE_DS=1 keeps the pickled-object members; parts of a sharded run merge their update_DS archives."""
    from make_prg_b200 import hostio
    from make_prg_b200.prg_builder import PrgBuilderZipDatabase

    monkeypatch.setenv("MPRG_PICKLE_DS", "1")
    _run_cli(tmp_path, REF / "sample_example", "pk")
    with zipfile.ZipFile(tmp_path / "pk.update_DS.zip") as zf:
        b = pickle.loads(zf.read("GC00006032"))
        assert b.build_prg() == truth_multi("sample_example")["GC00006032"]
    monkeypatch.delenv("MPRG_PICKLE_DS")
    _run_cli(tmp_path, REF / "sample_example", "tb")
    pk, tb = PrgBuilderZipDatabase(tmp_path / "pk.update_DS.zip"), PrgBuilderZipDatabase(tmp_path / "tb.update_DS.zip")
    pk.load(), tb.load()
    assert pk == tb
    pk.close(), tb.close()
    # two parts (one locus each) -> one archive with both members
    from make_prg_b200.subcommands import from_msa

    files = sorted((REF / "sample_example").glob("*.fa"))
    opts = _run_cli(tmp_path, REF / "sample_example", "whole", force=True)
    parts = []
    for k, f in enumerate(files):
        from_msa.build_and_write([f], opts, output_prefix=str(tmp_path / f"part{k}"), part=True)
        parts.append(str(tmp_path / f"part{k}"))
    opts.output_prefix = str(tmp_path / "merged")
    assert from_msa.merge_parts(parts, opts) == 2
    for ext in (".prg.fa", ".update_DS.zip", ".prg.bin.zip", ".prg.gfa.zip"):
        got, want = tmp_path / f"merged{ext}", tmp_path / f"whole{ext}"
        if ext == ".prg.fa":
            assert got.read_bytes() == want.read_bytes()
        else:
            with zipfile.ZipFile(got) as a, zipfile.ZipFile(want) as b:
                assert sorted(a.namelist()) == sorted(b.namelist())
                for m in b.namelist():
                    assert a.read(m) == b.read(m), (ext, m)


def test_cli_single_msa_and_skip_semantics(tmp_path):
    from make_prg_b200.subcommands.from_msa import EmptyMSAError

    _run_cli(tmp_path, REF / "match.nonmatch.fa", "one")
    t = REF / "truth" / "match.nonmatch"
    assert (tmp_path / "one.prg.fa").read_text().split("\n")[1] == truth_prg("match.nonmatch")
    assert (tmp_path / "one.prg.bin").read_bytes() == (t / "match.nonmatch.prg.bin").read_bytes()
    assert (tmp_path / "one.prg.gfa").read_bytes() == (t / "match.nonmatch.prg.gfa").read_bytes()
    # a locus with a disallowed base produces no output (tests/integration_tests/test_from_msa.py:183-187)
    _run_cli(tmp_path, REF / "fails_2.fa", "bad")
    assert not (tmp_path / "bad.prg.fa").exists()
    # an empty MSA aborts the run (test_from_msa.py:256-277)
    with pytest.raises(EmptyMSAError):
        _run_cli(tmp_path, REF / "several_empty", "empty")
    # existing outputs are protected unless --force
    with pytest.raises(RuntimeError):
        _run_cli(tmp_path, REF / "match.nonmatch.fa", "one")
    _run_cli(tmp_path, REF / "match.nonmatch.fa", "one", force=True)


def test_cli_other_alignment_formats_equal_fasta_input(tmp_path):
    """-f clustal / stockholm / phylip-relaxed (from_msa.py:48-55 passes the name to Bio.AlignIO): the same
    alignments in another layout give the same .prg.fa / .bin / .gfa as the FASTA files."""
    from make_prg_b200.utils.io_utils import parse_fasta

    names = ["GC00006032", "GC00010897"]
    recs = {}
    for n in names:
        with open(REF / "sample_example" / f"{n}.fa") as fh:
            recs[n] = [(r.id, r.seq) for r in parse_fasta(fh)]

    def clustal(rows):
        out = ["CLUSTAL W (1.83) multiple sequence alignment", "", ""]
        width, pad = len(rows[0][1]), max(len(rid) for rid, _ in rows) + 6
        for a in range(0, width, 60):
            out += [f"{rid:<{pad}}{seq[a:a + 60]}" for rid, seq in rows] + [" " * pad, ""]
        return "\n".join(out) + "\n"

    def stockholm(rows):
        return "# STOCKHOLM 1.0\n" + "".join(f"{rid} {seq}\n" for rid, seq in rows) + "//\n"

    def phylip_relaxed(rows):
        return f"{len(rows)} {len(rows[0][1])}\n" + "".join(f"{rid} {seq}\n" for rid, seq in rows)

    _run_cli(tmp_path, REF / "sample_example", "fasta")
    want = (tmp_path / "fasta.prg.fa").read_bytes()
    for fmt, writer, ext in (("clustal", clustal, "aln"), ("stockholm", stockholm, "sto"),
                             ("phylip-relaxed", phylip_relaxed, "phy")):
        d = tmp_path / f"in_{ext}"
        d.mkdir()
        for n in names:
            (d / f"{n}.{ext}").write_text(writer(recs[n]))
        _run_cli(tmp_path, d, fmt, alignment_format=fmt, suffix=ext)
        got = (tmp_path / f"{fmt}.prg.fa").read_bytes().replace(f".{ext}".encode(), b"")
        assert got == want.replace(b".fa", b""), fmt


def test_cli_chunked_pipeline_equals_one_batch(tmp_path, monkeypatch):
    """The load -> build -> write pipeline cuts a run into chunks of MPRG_CHUNK_MB of input: the final files
    do not depend on the cut (amira_MSAs + sample_example + the small cases, one file per chunk vs one chunk)."""
    import shutil

    src = tmp_path / "msas"
    src.mkdir()
    for f in sorted((REF / "amira_MSAs").iterdir()) + sorted((REF / "sample_example").iterdir()):
        if f.is_file():
            shutil.copy(f, src / f.name)
    for case in ("match.nonmatch.match", "nested_snps_deletion", "contains_n_and_RYKMSW", "fails_2"):
        shutil.copy(REF / f"{case}.fa", src / f"{case}.fa")
    _run_cli(tmp_path, src, "one", skip_update_ds=True)
    monkeypatch.setenv("MPRG_CHUNK_MB", "0.000001")
    from make_prg_b200.subcommands import from_msa

    files = from_msa.get_all_input_files(str(src), "")
    assert len(from_msa.cut_chunks(files)) == len(files) >= 8
    _run_cli(tmp_path, src, "many", skip_update_ds=True)
    assert (tmp_path / "one.prg.fa").read_bytes() == (tmp_path / "many.prg.fa").read_bytes()
    assert (tmp_path / "one.prg.fa").read_text().count(">") == len(files) - 1  # fails_2 is skipped
    assert not (tmp_path / "one.update_DS.zip").exists()
    for kind in ("bin", "gfa"):
        with zipfile.ZipFile(tmp_path / f"one.prg.{kind}.zip") as a, zipfile.ZipFile(tmp_path / f"many.prg.{kind}.zip") as b:
            assert a.namelist() == b.namelist() and a.testzip() is None and b.testzip() is None
            for member in a.namelist():
                assert a.read(member) == b.read(member), member
    truth = truth_multi("amira_MSAs")
    lines = (tmp_path / "many.prg.fa").read_text().split("\n")
    got = {lines[i][1:]: lines[i + 1] for i in range(0, len(lines) - 1, 2)}
    for name, prg in truth.items():
        assert got[name] == prg, name


def test_node_factory_build_below_parent(tmp_path):
    """NodeFactory.build(alignment, prg_builder, parent_node) of the host mirror: node ids count on from
    PrgBuilder.next_node_id and the new node can be swapped in as LeafNode._update_leaf does."""
    from helpers import sub_build_cases
    from make_prg_b200.msa import MSA, SeqRecord
    from make_prg_b200.prg_builder import PrgBuilder
    from make_prg_b200.recursion_tree import NodeFactory

    builder = PrgBuilder("match.nonmatch.match", REF / "match.nonmatch.match.fa", "fasta", 5, 7)
    for r in [c for c in sub_build_cases() if c["L"] == 7][:12]:
        builder.next_node_id = r["first_node_id"]
        parent = builder.root
        parent.nesting_level = r["parent_level"]
        msa = MSA([SeqRecord(s, f"s{i}") for i, s in enumerate(r["rows"])])
        node = NodeFactory.build(msa, builder, parent)
        assert node.parent is parent and node.node_id == r["first_node_id"]
        assert builder.next_node_id == r["next_node_id"]
        builder.site_num = 5
        parts = []
        node.preorder_traversal_to_build_prg(parts)
        assert "".join(parts) == r["prg"]
        dump = []

        def walk(n):
            dump.append([type(n).__name__, n.node_id, n.nesting_level, len(n.alignment),
                         n.alignment.get_alignment_length(), len(n.children)])
            for c in n.children:
                walk(c)

        walk(node)
        assert dump == [list(t) for t in r["tree"]]


def test_packed_host_rows_equal_device_pack_and_text_build():
    """mprg_build_packed: the host packer produces what pack_rows_kernel produces (bytes and alphabet flags),
    and a build from packed host rows equals the build from text (incl. N / RYKMSW / disallowed loci)."""
    import numpy as np
    from make_prg_b200 import device, hostio, synth

    ctx = device.default_context(0)
    rng = np.random.default_rng(21)
    mats = [synth.synth_msa(3 + (i * 7) % 23, w, 900 + i, var_frac=0.08, n_dels=2)
            for i, w in enumerate([1, 31, 32, 33, 100, 255, 1000, 64, 7, 513])]
    mats[2][1, 3] = ord("Z")
    mats[4][0, 5:9] = np.frombuffer(b"RYKM", np.uint8)
    mats[6][2, 500:503] = np.frombuffer(b"SWR", np.uint8)
    batch, res_text = ctx.build_ascii(mats, 5, 7)
    packed, flags = zip(*(hostio.pack_rows(m) for m in mats))
    for i, m in enumerate(mats):
        assert np.array_equal(batch.packed(i), packed[i]), i
    assert batch.flags().tolist() == list(flags)
    flat = np.concatenate([p.reshape(-1) for p in packed])
    offsets = np.cumsum([0] + [p.size for p in packed[:-1]])
    b2, res_packed = ctx.build_packed(flat, offsets, [m.shape[0] for m in mats], [m.shape[1] for m in mats], flags, 5, 7)
    assert b2.flags().tolist() == list(flags)
    for i in range(len(mats)):
        assert res_packed.status(i) == res_text.status(i)
        assert res_packed.prg(i) == res_text.prg(i)
        assert np.array_equal(b2.packed(i), packed[i])
    assert res_text.status(2) == 1 and res_text.status(0) == 0


def test_build_pipeline_lanes_equal_one_at_a_time_builds():
    """device.BuildPipeline: builds in flight side by side on two lanes (own contexts, own host threads) return
    exactly what one build at a time returns -- PRG strings, statuses, trees -- for every submission, whichever
    lane took it, with and without a consume callback."""
    import numpy as np
    from make_prg_b200 import device, hostio, synth

    ctx = device.default_context(0)
    sets = []
    for s in range(5):
        mats = [synth.synth_msa(4 + (i * 5 + s) % 40, 40 + (37 * i + 11 * s) % 700, 7000 + 100 * s + i,
                                var_frac=0.05 + 0.03 * (i % 4), n_dels=i % 3) for i in range(48)]
        if s == 1:
            mats[3][0, 2] = ord("Z")  # a locus that is skipped
        if s == 2:
            mats[5][1, 10:13] = np.frombuffer(b"RYK", np.uint8)  # host assembly (IUPAC product)
        packed, flags = zip(*(hostio.pack_rows(m) for m in mats))
        flat = np.concatenate([p.reshape(-1) for p in packed])
        offsets = np.cumsum([0] + [p.size for p in packed[:-1]])
        args = (flat, offsets, [m.shape[0] for m in mats], [m.shape[1] for m in mats], flags, 5, 7)
        b, res = ctx.build_packed(*args)
        want = [(res.status(i), res.prg(i), res.n_nodes(i)) for i in range(len(mats))]
        want_tree = {k: v.tolist() for k, v in res.nodes(7).items()}
        res.free()
        b.free()
        sets.append((args, want, want_tree))

    def consume(b, res):
        return ([(res.status(i), res.prg(i), res.n_nodes(i)) for i in range(res.n_loci)],
                {k: v.tolist() for k, v in res.nodes(7).items()})

    with device.BuildPipeline(0, depth=2) as pipe:
        order = [0, 1, 2, 3, 4, 4, 3, 2, 1, 0, 2, 2, 0]
        futs = [pipe.submit_packed(*sets[k][0], consume=consume) for k in order]
        for k, f in zip(order, futs):
            got, tree = f.result()
            assert got == sets[k][1], k
            assert tree == sets[k][2], k
        # without a callback the caller owns batch and result
        b, res = pipe.submit_packed(*sets[1][0]).result()
        assert [(res.status(i), res.prg(i), res.n_nodes(i)) for i in range(res.n_loci)] == sets[1][1]
        res.free()
        b.free()
        assert pipe.launch_count() > 0 and pipe.copy_stats()["h2d_bytes"] > 0


def test_c_caller_with_several_contexts_in_flight(tmp_path):
    """tests/c_abi/lanes.c on the GPU: four host threads of a plain C program, one context each, build the same
    loci from packed host rows side by side; every PRG equals the one a single context built."""
    import subprocess

    from test_abi import build_c_caller

    exe = build_c_caller(tmp_path)
    out = subprocess.run([str(exe), "4", "6"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 mismatches" in out.stdout


class _VectorAligner:
    """Stands in for the MSA aligner of `make_prg update`: hands back the alignment of a golden vector."""

    def __init__(self):
        self.next = {}

    def get_updated_alignment(self, current_alignment, new_sequences):
        return self.next[frozenset(new_sequences)]


def test_leaf_update_through_sub_builds(tmp_path):
    """LeafNode._update_leaf / batch_update / batch_update_leaves (recursion_tree.py:345-388) driven by the 216
    NodeFactory.build(alignment, builder, parent_node) vectors of the unmodified reference: the leaf is replaced
    by the re-built node (any subclass), node ids count on from next_node_id, the PRG index is invalidated, and
    the batched form (one device batch for all leaves) gives the same trees as one call per leaf."""
    from helpers import sub_build_cases
    from make_prg_b200.msa import MSA, SeqRecord
    from make_prg_b200.prg_builder import PrgBuilder
    from make_prg_b200.recursion_tree import LeafNode, MultiIntervalNode, batch_update_leaves

    def dump(node):
        out = []

        def walk(n):
            out.append([type(n).__name__, n.node_id, n.nesting_level, len(n.alignment),
                        n.alignment.get_alignment_length(), len(n.children)])
            for c in n.children:
                walk(c)

        walk(node)
        return out

    def make_case(r, k):
        """A builder whose root (nesting level = the vector's parent level) holds one leaf to be updated."""
        builder = PrgBuilder.__new__(PrgBuilder)
        builder._locus_name = f"locus{k}"
        builder.max_nesting, builder.min_match_length = r["N"], r["L"]
        builder.aligner = _VectorAligner()
        builder.next_node_id = r["first_node_id"]
        builder.site_num, builder.prg_index, builder.engine_prg = 5, {}, None
        old = MSA([SeqRecord("ACGT", "old0"), SeqRecord("ACGA", "old1")])
        root = MultiIntervalNode(r["parent_level"], old, None, builder, 0)
        leaf = LeafNode(r["parent_level"], old, root, builder, 1)
        root._children.append(leaf)
        builder.root = root
        builder.update_PRG_index(0, 4, leaf)
        leaf.new_sequences = {f"NEW{k}"}
        builder.aligner.next[frozenset(leaf.new_sequences)] = MSA(
            [SeqRecord(s, f"s{i}") for i, s in enumerate(r["rows"])])
        return builder, root, leaf

    def check(builder, root, leaf, r):
        node = root.children[0]
        assert node is not leaf and node.parent is root and node.node_id == r["first_node_id"]
        assert builder.next_node_id == r["next_node_id"] and builder.prg_index == {}
        builder.site_num = 5
        parts = []
        node.preorder_traversal_to_build_prg(parts)
        assert "".join(parts) == r["prg"]
        assert dump(node) == [list(t) for t in r["tree"]]

    cases = sub_build_cases()
    assert len(cases) == 216
    # one call per leaf (LeafNode.batch_update -> _update_leaf), on a sample
    for k, r in enumerate(cases[::9]):
        builder, root, leaf = make_case(r, k)
        leaf.batch_update()
        check(builder, root, leaf, r)
    # all 216 leaves in one device batch per (N, L)
    made = [make_case(r, k) for k, r in enumerate(cases)]
    assert batch_update_leaves([leaf for _b, _r, leaf in made]) == 216
    for (builder, root, leaf), r in zip(made, cases):
        check(builder, root, leaf, r)
    # a leaf without new sequences is left alone
    builder, root, leaf = make_case(cases[0], 999)
    leaf.new_sequences = set()
    leaf.batch_update()
    assert root.children[0] is leaf and batch_update_leaves([leaf]) == 0
