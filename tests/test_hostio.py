"""CPU: the native loader and writers of libmprg (csrc/hostio.cpp) against the
host-side mirrors of the reference (utils/io_utils.py, utils/gfa.py, utils/prg_encoder.py) and against
the reference's golden output files.  No device is needed: these entry points are plain host code."""
import gzip
import zipfile
from pathlib import Path

import numpy as np
import pytest

from helpers import REF, truth_multi

from make_prg_b200 import hostio
from make_prg_b200.utils import gfa as gfa_py
from make_prg_b200.utils import io_utils
from make_prg_b200.utils.prg_encoder import PrgEncoder


def fixture_fastas():
    files = sorted(p for p in REF.glob("*.fa*") if p.is_file())
    for sub in ("sample_example", "amira_MSAs", "several", "several_compressed"):
        files += sorted(p for p in (REF / sub).rglob("*") if p.is_file())
    return files


def all_truth_prgs():
    out = []
    for f in sorted((REF / "truth").glob("*/*.prg.fa")):
        lines = f.read_text().split("\n")
        out += [lines[i + 1] for i in range(0, len(lines) - 1, 2)]
    return out


def test_loader_equals_python_loader_on_every_fixture():
    files = fixture_fastas()
    assert len(files) >= 30 and any(str(f).endswith(".gz") for f in files)
    msas = hostio.load_fasta_files(files, threads=3, pin=False)
    assert msas.n_loci == len(files)
    with_n = 0
    for i, f in enumerate(files):
        want = io_utils.load_alignment_file(str(f))
        assert msas.status[i] == hostio.LOAD_OK, f
        assert np.array_equal(msas.matrix(i), want.matrix), f
        assert msas.ids(i) == want.ids, f
        assert [r.description for r in msas.alignment(i)] == [r.description for r in want], f
        with_n += int(msas.flags[i] & hostio.FLAG_HAS_N)
    assert not any(msas.flags & hostio.FLAG_DUPLICATE_IDS)
    assert with_n >= 4  # the contains_n* fixtures went through the in-place N replacement
    # the loci sit back to back, as mprg_build_ascii wants them
    sizes = msas.n_rows.astype(np.int64) * msas.n_cols
    assert np.array_equal(msas.offsets, np.concatenate([[0], np.cumsum(sizes)[:-1]]))
    assert msas.ascii_bytes == int(sizes.sum())
    msas.free()


def test_loader_text_edge_cases(tmp_path):
    cases = {
        "crlf.fa": b">a desc one\r\nACgt\r\nAC\r\n>b\r\nAC-TNC\r\n",
        "lone_cr.fa": b">a\rACGT\r>b\rAC-T\r",
        "blank_and_spaces.fa": b"; comment\n\n>a  two  spaces \nAC GT\n\n  \nAC\n>b\tTab\n\nACGTAC\n",
        "no_trailing_newline.fa": b">a\nACGT\n>b\nAC-T",
        "long_lines.fa": b">a\n" + b"ACGT" * 300 + b"\n>b\n" + b"acgt" * 150 + b"\n" + b"AC-T" * 150 + b"\n",
        "empty_title.fa": b">\nAC\n>x\nAG\n",
        "zero_columns.fa": b">a\n>b\n",
        "wrapped_gt.fa": b">a\nAC\n>b\nA\n>\n",
        # Biopython rstrips every line: trailing tabs / form feeds go, inner ones stay
        "trailing_tabs.fa": b">a \t\nAC\t\nGT\t \n>b\x0c\nACGT\x0c\n>c\nAC\x1f\n\t\nGT",
        "inner_tab.fa": b">a\nA\tC\nGT\n>b x\nACG\n-T\n",
        "utf8_title.fa": ">s1 caf\u00e9 \u2003\nACGT\n>\u00e9\u00e8 second\nAC-T\n".encode("utf-8"),
    }
    paths = []
    for name, blob in cases.items():
        (tmp_path / name).write_bytes(blob)
        paths.append(tmp_path / name)
    with gzip.open(tmp_path / "multi.fa.gz", "wb") as fh:
        fh.write(b">a\nACGT\n")
    with gzip.open(tmp_path / "multi.fa.gz", "ab") as fh:  # second gzip member, as `cat a.gz b.gz`
        fh.write(b">b\nACGA\n")
    paths.append(tmp_path / "multi.fa.gz")
    msas = hostio.load_fasta_files(paths, threads=2, pin=False)
    for i, p in enumerate(paths):
        try:
            want = io_utils.load_alignment_file(str(p))
        except ValueError as err:
            assert msas.status[i] in (hostio.LOAD_NO_RECORDS, hostio.LOAD_RAGGED), p
            with pytest.raises(ValueError, match=str(err.args[0])[:20]):
                hostio.raise_for_load_status(msas, i)
            continue
        assert msas.status[i] == hostio.LOAD_OK, p
        assert np.array_equal(msas.matrix(i), want.matrix), p
        assert msas.ids(i) == want.ids, p
        assert msas.titles(i) == [r.description for r in want], p
    msas.free()


def test_loader_statuses(tmp_path):
    (tmp_path / "empty.fa").write_bytes(b"")
    (tmp_path / "ragged.fa").write_bytes(b">a\nACGT\n>b\nACG\n")
    (tmp_path / "latin.fa").write_bytes(">a\nAC\u00e9T\n".encode("utf-8"))
    (tmp_path / "broken.fa.gz").write_bytes(b"not a gzip stream at all")
    (tmp_path / "badtitle.fa").write_bytes(b">s\xff1\nACGT\n")
    paths = [tmp_path / "empty.fa", tmp_path / "ragged.fa", tmp_path / "latin.fa", tmp_path / "missing.fa",
             tmp_path / "broken.fa.gz", tmp_path / "badtitle.fa"]
    msas = hostio.load_fasta_files(paths, threads=2, pin=False)
    assert list(msas.status[:4]) == [hostio.LOAD_NO_RECORDS, hostio.LOAD_RAGGED, hostio.LOAD_NOT_ASCII,
                                     hostio.LOAD_IO_ERROR]
    assert list(msas.n_rows) == [0] * 6 and msas.ascii_bytes == 0
    (tmp_path / "dup.fa").write_bytes(b">a x\nAC\n>b\nAC\n>a y\nAG\n")
    dup = hostio.load_fasta_files([tmp_path / "dup.fa"], threads=1, pin=False)
    assert dup.status[0] == hostio.LOAD_OK and dup.flags[0] & hostio.FLAG_DUPLICATE_IDS
    dup.free()
    # non-ASCII among the sequences: the text loads, the locus is a curation error (skipped, not fatal)
    with pytest.raises(hostio.NonAsciiSequenceError):
        hostio.raise_for_load_status(msas, 2)
    # a title that is not UTF-8: the reference's read fails with the decoder's error
    assert msas.status[5] == hostio.LOAD_NOT_ASCII
    with pytest.raises(UnicodeDecodeError):
        hostio.raise_for_load_status(msas, 5)
    with pytest.raises(ValueError, match="No records found in handle"):
        hostio.raise_for_load_status(msas, 0)
    with pytest.raises(ValueError, match="same length"):
        hostio.raise_for_load_status(msas, 1)
    with pytest.raises(FileNotFoundError):
        hostio.raise_for_load_status(msas, 3)
    msas.free()


def test_loader_vector_and_scalar_paths_agree(tmp_path, monkeypatch):
    rng = np.random.default_rng(5)
    paths = []
    for k in range(6):
        rows, cols, wrap = int(rng.integers(1, 40)), int(rng.integers(0, 700)), int(rng.integers(1, 130))
        M = rng.choice(np.frombuffer(b"ACGTacgtRYKMSW-N", np.uint8), size=(rows, cols))
        with open(tmp_path / f"r{k}.fa", "wb") as fh:
            for r in range(rows):
                s = M[r].tobytes()
                fh.write(b">row%d x\n" % r)
                for c in range(0, cols, wrap):
                    fh.write(s[c:c + wrap] + (b"\r\n" if k % 2 else b"\n"))
        paths.append(tmp_path / f"r{k}.fa")
    fast = hostio.load_fasta_files(paths, threads=2, pin=False)
    monkeypatch.setenv("MPRG_NO_AVX2", "1")
    slow = hostio.load_fasta_files(paths, threads=1, pin=False)
    for i, p in enumerate(paths):
        want = io_utils.load_alignment_file(str(p))
        assert np.array_equal(fast.matrix(i), want.matrix) and np.array_equal(slow.matrix(i), want.matrix), p
        assert b"N" not in fast.matrix(i).tobytes() or fast.n_cols[i] == 0
    fast.free()
    slow.free()


def test_gfa_and_bin_equal_python_mirrors_and_golden_files():
    prgs = all_truth_prgs()
    assert len(prgs) >= 25
    for prg in prgs:
        g = gfa_py.GFA_Output(gfa_py.HEADER)
        g.build_gfa_string(prg_string=prg)
        assert hostio.prg_to_gfa(prg) == g.gfa_string
        assert hostio.encode_prg(prg).tolist() == PrgEncoder().encode(prg)
    t = REF / "truth" / "match.nonmatch"
    prg = (t / "match.nonmatch.prg.fa").read_text().split("\n")[1]
    assert hostio.encode_prg(prg).astype("<u4").tobytes() == (t / "match.nonmatch.prg.bin").read_bytes()
    assert hostio.prg_to_gfa(prg).encode() == (t / "match.nonmatch.prg.gfa").read_bytes()


def test_encoder_reference_unit_vectors_and_errors():
    # tests/test_prg_encoder.py of the reference
    assert hostio.encode_prg("").tolist() == []
    assert hostio.encode_prg("ACGT").tolist() == [1, 2, 3, 4]
    assert hostio.encode_prg("a 5 g 6 t 5 c").tolist() == [1, 5, 3, 6, 4, 6, 2]
    assert hostio.encode_prg("TC 5 ACTC 7 TAGTCA 8 TTGTGA 7  6 AACTAG 5 AG").tolist() == PrgEncoder().encode(
        "TC 5 ACTC 7 TAGTCA 8 TTGTGA 7  6 AACTAG 5 AG")
    with pytest.raises(hostio.EncodeError):
        hostio.encode_prg("AC 5 AX 6 T 5 ")
    with pytest.raises(ValueError):
        hostio.encode_prg("A 5 C 6 T 5 A 5 ")
    with pytest.raises(AssertionError):
        hostio.prg_to_gfa("A 5 C 5 ")  # a site with one allele (gfa.py:64-66)
    with pytest.raises(AssertionError):
        hostio.prg_to_gfa("A 5 C 6 T")  # unterminated site


def test_host_packer_and_packed_loader(tmp_path, monkeypatch):
    """The loader's packed mode: rows in the 4-bit device layout (what pack_rows_kernel produces on the device,
    checked against it in tests/test_gpu_host_api.py), vector and scalar packers agree, the text comes back."""
    rng = np.random.default_rng(11)
    alphabet = np.frombuffer(b"ACGTacgtRYKMSWN-Z@`[{", np.uint8)
    good = np.frombuffer(b"ACGTRYKMSWN-", np.uint8)
    for rows, cols in [(5, 1), (3, 31), (4, 32), (7, 33), (2, 64), (9, 1000), (1, 0), (0, 5), (6, 257)]:
        M = rng.choice(alphabet, size=(rows, cols))
        P, flags = hostio.pack_rows(M)
        assert P.shape == (rows, hostio.packed_stride(cols))
        monkeypatch.setenv("MPRG_NO_AVX2", "1")
        P2, flags2 = hostio.pack_rows(M)
        monkeypatch.delenv("MPRG_NO_AVX2")
        assert np.array_equal(P, P2) and flags == flags2
        upper = np.where((M >= ord("a")) & (M <= ord("z")), M - 32, M)
        want = np.where(np.isin(upper, good), upper, ord("?"))
        assert np.array_equal(hostio.unpack_rows(P, cols), want)
        if M.size:
            assert bool(flags & 1) == bool((~np.isin(upper, good)).any())
            assert bool(flags & 2) == bool((upper == ord("N")).any())
            assert bool(flags & 4) == bool(np.isin(upper, np.frombuffer(b"RYKMSW", np.uint8)).any())
            assert bool(flags & 8) == bool(np.isin(upper, np.frombuffer(b"MSWN", np.uint8)).any())
        # padding columns are MPRG_SYM_PAD (15) and raise no flag
        if cols % 32:
            full = hostio.unpack_rows(P, hostio.packed_stride(cols) * 2)
            assert (full[:, cols:] == ord("?")).all()
    files = fixture_fastas()
    text = hostio.load_fasta_files(files, threads=3, pin=False)
    packed = hostio.load_fasta_files(files, threads=3, pin=False, packed=True)
    both = hostio.load_fasta_files(files, threads=2, pin=False, packed=True, keep_ascii=True)
    assert packed.ascii_bytes == 0 and both.ascii_bytes == text.ascii_bytes
    at = 0
    for i in range(text.n_loci):
        M = text.matrix(i)
        assert packed.packed_offsets[i] == at
        at += M.shape[0] * hostio.packed_stride(M.shape[1])
        want, flags = hostio.pack_rows(M)
        po = int(packed.packed_offsets[i])
        assert np.array_equal(packed.packed[po:po + want.size].reshape(want.shape), want)
        assert packed.alphabet_flags[i] == flags == both.alphabet_flags[i]
        if flags & 1:  # a disallowed character (the locus is skipped anyway): '?' in the unpacked text
            assert ((packed.matrix(i) == M) | (packed.matrix(i) == ord("?"))).all()
        else:
            assert np.array_equal(packed.matrix(i), M)  # unpacked from the 4-bit rows
        assert np.array_equal(both.matrix(i), M) and packed.ids(i) == text.ids(i)
    assert packed.packed_bytes == at
    for m in (text, packed, both):
        m.free()


def test_writer_final_files(tmp_path):
    truth = truth_multi("sample_example")
    names = sorted(truth, reverse=True)  # archive order = order added, .prg.fa = sorted
    strings = hostio.PrgStrings([truth[n] for n in names])
    w = hostio.OutputWriter(tmp_path / "out", threads=2)
    w.add(strings, [0, 1], names)
    n, nbytes = w.close()
    assert n == 2 and nbytes > 0
    tdir = REF / "truth" / "sample_example"
    assert (tmp_path / "out.prg.fa").read_bytes() == (tdir / "sample_example.prg.fa").read_bytes()
    for kind in ("bin", "gfa"):
        with zipfile.ZipFile(tmp_path / f"out.prg.{kind}.zip") as got, \
                zipfile.ZipFile(tdir / f"sample_example.prg.{kind}.zip") as want:
            assert got.testzip() is None
            assert got.namelist() == [f"{n}.{kind}" for n in names]
            assert got.infolist()[0].compress_type == zipfile.ZIP_STORED
            for member in want.namelist():
                assert got.read(member) == want.read(member), member
    # one locus: plain files, no archives (input_output_files.py:107-112,122-126)
    w = hostio.OutputWriter(tmp_path / "one", threads=1)
    w.add(strings, [1], [names[1]])
    assert w.close()[0] == 1
    assert (tmp_path / "one.prg.bin").read_bytes() == hostio.encode_prg(truth[names[1]]).astype("<u4").tobytes()
    assert (tmp_path / "one.prg.gfa").read_text() == hostio.prg_to_gfa(truth[names[1]])
    assert not (tmp_path / "one.prg.bin.zip").exists()
    # nothing added: nothing written; only some outputs asked for
    assert hostio.OutputWriter(tmp_path / "none").close() == (0, 0)
    assert not list(tmp_path.glob("none*"))
    w = hostio.OutputWriter(tmp_path / "g", prg=False, binary=False, gfa=True)
    w.add(strings, [0, 1], names)
    w.add(strings, [0], ["again"])
    w.close()
    assert sorted(p.name for p in tmp_path.glob("g.*")) == ["g.prg.gfa.zip"]
    with zipfile.ZipFile(tmp_path / "g.prg.gfa.zip") as z:
        assert z.namelist() == [f"{names[0]}.gfa", f"{names[1]}.gfa", "again.gfa"]
    with pytest.raises(hostio.EncodeError):
        bad = hostio.PrgStrings(["AC 5 AX 6 T 5 "])
        hostio.OutputWriter(tmp_path / "bad").add(bad, [0], ["bad"])
    strings.free()


def test_writer_parts_merge_and_abort(tmp_path):
    """A sharded run: every shard writes a part, mprg_merge_outputs makes the final files -- byte-identical
    .prg.fa and the same archive members as one writer; unfinished files never carry a final name."""
    prgs = all_truth_prgs()[:7]
    names = [f"locus{c}" for c in "dbagfce"]
    strings = hostio.PrgStrings(prgs)
    w = hostio.OutputWriter(tmp_path / "whole", threads=2)
    w.add(strings, np.arange(7), names)
    w.close()
    shards = [[0, 3, 4], [1], [], [2, 5, 6]]
    parts = []
    for k, idx in enumerate(shards):
        pw = hostio.OutputWriter(tmp_path / f"part{k}", threads=1, part=True)
        if idx:
            pw.add(strings, idx, [names[i] for i in idx])
        pw.close()
        parts.append(tmp_path / f"part{k}")
    assert (tmp_path / "part1.prg.bin.zip").exists()  # a part of one locus is still an archive
    assert hostio.merge_outputs(parts, tmp_path / "merged") == 7
    assert (tmp_path / "merged.prg.fa").read_bytes() == (tmp_path / "whole.prg.fa").read_bytes()
    for kind in ("bin", "gfa"):
        with zipfile.ZipFile(tmp_path / f"merged.prg.{kind}.zip") as got, \
                zipfile.ZipFile(tmp_path / f"whole.prg.{kind}.zip") as want:
            assert got.testzip() is None
            assert got.namelist() == [f"{names[i]}.{kind}" for idx in shards for i in idx]
            for member in want.namelist():
                assert got.read(member) == want.read(member), member
    assert not list(tmp_path.glob("part*"))
    # the whole run holds one locus: plain files
    pw = hostio.OutputWriter(tmp_path / "p0", part=True)
    pw.add(strings, [2], ["only"])
    pw.close()
    assert hostio.merge_outputs([tmp_path / "p0", tmp_path / "p1"], tmp_path / "single") == 1
    assert (tmp_path / "single.prg.bin").read_bytes() == hostio.encode_prg(prgs[2]).astype("<u4").tobytes()
    assert (tmp_path / "single.prg.gfa").read_text() == hostio.prg_to_gfa(prgs[2])
    assert (tmp_path / "single.prg.fa").read_text() == f">only\n{prgs[2]}\n"
    assert not (tmp_path / "single.prg.bin.zip").exists()
    # an aborted writer leaves nothing behind (truncated archives blocked reruns)
    aw = hostio.OutputWriter(tmp_path / "gone", threads=1)
    aw.add(strings, [0, 1, 2], names[:3])
    assert not (tmp_path / "gone.prg.bin.zip").exists()  # still under its temporary name
    aw.abort()
    assert not list(tmp_path.glob("gone*"))
    strings.free()


def test_mprg_error_crosses_process_boundaries():
    import pickle

    from make_prg_b200._lib import MprgError

    err = pickle.loads(pickle.dumps(MprgError(-2, "kernel launch failed")))
    assert isinstance(err, MprgError) and err.code == -2 and "kernel launch failed" in str(err)


def test_writer_many_entries_zip64(tmp_path):
    n = 66000  # > 65535 entries: the archive needs the zip64 end-of-central-directory records
    strings = hostio.PrgStrings(["A 5 C 6 G 5 T"])
    w = hostio.OutputWriter(tmp_path / "big", prg=False, binary=True, gfa=False, threads=2)
    names = [f"locus{i}" for i in range(n)]
    w.add(strings, np.zeros(n, np.int32), names)
    assert w.close()[0] == n
    with zipfile.ZipFile(tmp_path / "big.prg.bin.zip") as z:
        assert len(z.namelist()) == n
        assert z.read("locus65999.bin") == np.array([1, 5, 2, 6, 3, 6, 4], "<u4").tobytes()


def test_native_n_replacement_equals_python_random():
    rng = np.random.default_rng(11)
    alphabet = np.frombuffer(b"ACGTACGTACGT--NNRYKMSW", np.uint8)
    for k in range(40):
        rows, cols = int(rng.integers(1, 30)), int(rng.integers(1, 400))
        M = rng.choice(alphabet, size=(rows, cols))
        if k % 5 == 0:
            M[:, : cols // 3] = ord("N")  # columns full of N: choice("ACGT")
        if k % 7 == 0:
            M[:] = rng.choice(np.frombuffer(b"AC", np.uint8), size=(rows, cols))  # ties everywhere
            M[0, 0] = ord("N")
        a, b = M.copy(), M.copy()
        hostio.replace_n(a)
        hostio.replace_n_in_place(b)
        assert np.array_equal(a, b), k
        assert ord("N") not in a


def test_pipeline_helpers(tmp_path, monkeypatch):
    """Host logic of the from_msa pipeline that needs no device: locus names (input_output_files.py:234-235),
    chunk cutting, thread split between the stages."""
    from pathlib import Path

    from make_prg_b200.subcommands import from_msa as fm

    for name in ["a/b/x.fa", "x.fasta.gz", "y.fa.gz", "z.fasta", "w.txt", "q.fa.fa", "/p/GC0001.fa", "n.fa.gz.fa"]:
        assert fm.locus_name_of(name) == fm.remove_known_input_extensions(Path(name).name), name
    files = []
    for i, size in enumerate([10, 300, 5, 5, 400, 1]):
        p = tmp_path / f"l{i}.fa"
        p.write_bytes(b"A" * size)
        files.append(p)
    chunks = fm.cut_chunks(files, max_bytes=310)
    assert [[f.name for f in c] for c in chunks] == [["l0.fa", "l1.fa"], ["l2.fa", "l3.fa"], ["l4.fa"], ["l5.fa"]]
    assert sum(len(c) for c in fm.cut_chunks(files + [tmp_path / "missing.fa"], max_bytes=10 ** 9)) == 7
    assert fm.cut_chunks(files, max_bytes=10 ** 9, max_loci=4) == [files[:4], files[4:]]
    monkeypatch.setenv("MPRG_CHUNK_MB", "0.0001")  # ~105 bytes
    assert len(fm.cut_chunks(files)) == 5
    assert fm.side_threads(1) >= fm.side_threads(4) >= 1 and fm.side_threads(4, writer=True) >= 1


def test_chunk_generator_with_stub_pipeline(tmp_path, monkeypatch):
    """from_msa.iter_built_chunks without a device: the native loader is real, the build pipeline is a stub.
    Chunk k+1 is submitted before the result of chunk k is handed on, results come in input order, a skipped locus
    (disallowed base) is left out of `ok`, and an error -- in a build or in the consumer -- leaves no build in
    flight and frees what was loaded."""
    from argparse import Namespace
    from concurrent.futures import Future

    from make_prg_b200 import device
    from make_prg_b200._lib import MprgError
    from make_prg_b200.subcommands import from_msa as fm

    files = []
    for i in range(5):
        p = tmp_path / f"l{i}.fa"
        p.write_text(f">a\nACGT{'ACGT'[i % 4]}A\n>b\nACGTTA\n")
        files.append(p)
    chunks = [files[:2], files[2:3], files[3:]]
    events, freed = [], []

    class Res:
        def __init__(self, n, bad):
            self.n, self.bad = n, bad

        def statuses(self):
            st = np.zeros(self.n, np.int32)
            for i in self.bad:
                st[i] = 1
            return st, np.full(self.n, 7, np.int64)

        def free(self):
            freed.append("res")

    class Batch:
        def free(self):
            freed.append("batch")

    class StubPipe:
        fail_at = None

        def __init__(self):
            self.k = 0

        def submit_msa_set(self, msas, N, L, consume=None):
            k, self.k = self.k, self.k + 1
            events.append(("submit", k, msas.n_loci))
            fut = Future()
            if StubPipe.fail_at == k:
                fut.set_exception(MprgError(-2, "injected"))
            else:
                fut.set_result((Batch(), Res(msas.n_loci, [1] if k == 0 else [])))
            return fut

    monkeypatch.setattr(device, "default_pipeline", lambda dev=0, depth=None: StubPipe())
    real_free = hostio.MsaSet.free

    def counted_free(self):  # (also called by __del__: count a set once)
        if not getattr(self, "_counted", False):
            self._counted = True
            freed.append("msas")
        return real_free(self)

    monkeypatch.setattr(hostio.MsaSet, "free", counted_free)
    opts = Namespace(alignment_format="fasta", max_nesting=5, min_match_length=7)
    got = []
    for names, msas, res, ok in fm.iter_built_chunks(files, opts, chunks=chunks):
        events.append(("yield", names[0]))
        got.append((names, ok))
        res.free()
        msas.free()
    assert got == [(["l0", "l1"], [0]), (["l2"], [0]), (["l3", "l4"], [0, 1])]
    # one build of lookahead: chunk 1 is submitted before chunk 0 is yielded, chunk 2 before chunk 1
    assert events == [("submit", 0, 2), ("submit", 1, 1), ("yield", "l0"), ("submit", 2, 2), ("yield", "l2"),
                      ("yield", "l3")]
    assert freed.count("batch") == 3 and freed.count("res") == 3 and freed.count("msas") == 3
    # a build fails: the error surfaces, the chunk that was in flight behind it is waited for and dropped
    del events[:], freed[:]
    StubPipe.fail_at = 1
    with pytest.raises(MprgError):
        for names, msas, res, ok in fm.iter_built_chunks(files, opts, chunks=chunks):
            res.free()
            msas.free()
    assert ("submit", 2, 2) in events and freed.count("msas") == 2 + 1  # chunk 0 by the consumer, chunk 2 by the cleanup
    # the consumer fails: the build in flight is waited for, its batch / result / matrices are released
    del events[:], freed[:]
    StubPipe.fail_at = None
    with pytest.raises(RuntimeError):
        for names, msas, res, ok in fm.iter_built_chunks(files, opts, chunks=chunks):
            res.free()
            msas.free()
            raise RuntimeError("consumer")
    # (chunk 0 by the consumer, chunk 1 in flight, chunk 2 loaded ahead)
    assert freed.count("msas") == 3 and freed.count("batch") == 2 and freed.count("res") == 2


def test_writer_streaming_appends_equal_serial_and_stop_on_a_bad_prg(tmp_path):
    """mprg_writer_add with many loci and several threads appends the archives while the PRGs are still being
    encoded: the files equal those of a one-thread writer (and of a writer fed locus by locus), and a PRG that
    cannot be encoded raises the reference's error and leaves no file behind."""
    prgs = all_truth_prgs()
    many = [prgs[i % len(prgs)] for i in range(300)]
    names = [f"l{i:04d}" for i in range(300)]
    strings = hostio.PrgStrings(many)
    for tag, threads, step in (("serial", 1, 300), ("streamed", 6, 300), ("two_adds", 6, 150), ("one_by_one", 3, 1)):
        w = hostio.OutputWriter(tmp_path / tag, threads=threads)
        for a in range(0, 300, step):
            w.add(strings, np.arange(a, min(300, a + step)), names[a:a + step])
        assert w.close()[0] == 300
    for tag in ("streamed", "two_adds", "one_by_one"):
        for ext in ("prg.fa", "prg.bin.zip", "prg.gfa.zip"):
            a, b = (tmp_path / f"serial.{ext}").read_bytes(), (tmp_path / f"{tag}.{ext}").read_bytes()
            if ext.endswith("zip"):  # (the DOS time stamp of the members may differ by a tick)
                with zipfile.ZipFile(tmp_path / f"serial.{ext}") as za, zipfile.ZipFile(tmp_path / f"{tag}.{ext}") as zb:
                    assert za.namelist() == zb.namelist() and zb.testzip() is None
                    assert all(za.read(m) == zb.read(m) for m in za.namelist())
            else:
                assert a == b
    strings.free()
    bad = list(many)
    bad[211] = "ACGT 5 A 6 X 5 "  # not a PRG: a letter outside ACGT
    strings = hostio.PrgStrings(bad)
    w = hostio.OutputWriter(tmp_path / "bad", threads=6)
    with pytest.raises(Exception) as err:
        w.add(strings, np.arange(300), names)
    assert type(err.value).__name__ in ("EncodeError", "AssertionError", "ValueError", "MprgError")
    one = hostio.OutputWriter(tmp_path / "bad1", threads=1)
    with pytest.raises(type(err.value)):
        one.add(strings, np.arange(300), names)  # the one-thread path raises the same error
    assert not list(tmp_path.glob("bad*"))
    strings.free()
