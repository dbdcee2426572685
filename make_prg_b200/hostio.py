"""Python face of libmprg's host-side I/O (include/mprg.h, csrc/hostio.cu): the native FASTA loader
in front of the hot path and the native .prg.fa / .bin / .gfa writers behind it.

    load_fasta_files  ~ load_alignment_file per file      make_prg/utils/io_utils.py:17-49
    encode_prg        ~ PrgEncoder.encode                 make_prg/utils/prg_encoder.py:74-91
    prg_to_gfa        ~ GFA_Output.write_gfa's text       make_prg/utils/gfa.py:39-109
    OutputWriter      ~ InputOutputFiles.create_final_files  make_prg/utils/input_output_files.py:70-135
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import MprgError, ptr
from .msa import MSA
from .utils import io_utils

LOAD_OK, LOAD_NO_RECORDS, LOAD_RAGGED, LOAD_IO_ERROR, LOAD_NOT_ASCII = 0, 1, 2, 3, 4
FLAG_HAS_N, FLAG_DUPLICATE_IDS = 1, 2
WRITE_PRG, WRITE_BIN, WRITE_GFA, WRITE_PART, WRITE_DS = 1, 2, 4, 8, 16


def default_threads():
    return max(1, min(32, os.cpu_count() or 1))


def _c_strings(strings):
    raw = [s.encode() if isinstance(s, str) else bytes(s) for s in strings]
    arr = (C.c_char_p * max(len(raw), 1))(*raw)
    return arr, raw


def _view(pointer, ctype, n):
    if n == 0 or not pointer:
        return np.zeros(0, np.dtype(ctype))
    return np.ctypeslib.as_array(C.cast(pointer, C.POINTER(ctype)), shape=(n,))


class MsaSet:
    """The loci of a list of FASTA files in one host buffer (pinned when a device is present), laid out
    as mprg_build_ascii takes them.  Views are valid until free()."""

    def __init__(self, handle, paths):
        self.lib = _lib.load()
        self.handle = handle
        self.paths = [str(p) for p in paths]
        n, ascii_p, nbytes = C.c_int32(), C.c_void_p(), C.c_int64()
        offs, nr, nc, st, fl = (C.c_void_p() for _ in range(5))
        rc = self.lib.mprg_fasta_info(handle, C.byref(n), C.byref(ascii_p), C.byref(nbytes), C.byref(offs),
                                      C.byref(nr), C.byref(nc), C.byref(st), C.byref(fl))
        if rc != 0:
            raise MprgError(rc, "mprg_fasta_info failed")
        self.n_loci = n.value
        self.ascii = _view(ascii_p, C.c_uint8, max(nbytes.value, 1))
        self.ascii_bytes = nbytes.value
        self.offsets = _view(offs, C.c_int64, self.n_loci)
        self.n_rows = _view(nr, C.c_int32, self.n_loci)
        self.n_cols = _view(nc, C.c_int32, self.n_loci)
        self.status = _view(st, C.c_int32, self.n_loci)
        self.flags = _view(fl, C.c_int32, self.n_loci)
        # packed mode: the matrices in the 4-bit device layout (mprg_build_packed takes them as they are)
        self.packed = self.packed_offsets = self.alphabet_flags = None
        pk, pbytes, poffs, pflags = C.c_void_p(), C.c_int64(), C.c_void_p(), C.c_void_p()
        if self.lib.mprg_fasta_packed(handle, C.byref(pk), C.byref(pbytes), C.byref(poffs), C.byref(pflags)) == 0:
            self.packed = _view(pk, C.c_uint8, max(pbytes.value, 1))
            self.packed_bytes = pbytes.value
            self.packed_offsets = _view(poffs, C.c_int64, self.n_loci)
            self.alphabet_flags = _view(pflags, C.c_int32, self.n_loci)

    def matrix(self, locus):
        """uint8[rows, cols] of one locus: a (writable) view of the text, or -- when the loader only kept the
        packed rows -- the text unpacked from them."""
        r, c, o = int(self.n_rows[locus]), int(self.n_cols[locus]), int(self.offsets[locus])
        if self.ascii_bytes == 0 and self.packed is not None:
            stride = packed_stride(c)
            po = int(self.packed_offsets[locus])
            return unpack_rows(self.packed[po:po + r * stride].reshape(r, stride), c)
        return self.ascii[o:o + r * c].reshape(r, c)

    def titles(self, locus):
        n = C.c_int64()
        p = self.lib.mprg_fasta_titles(self.handle, locus, C.byref(n))
        text = C.string_at(p, n.value).decode() if p and n.value else ""
        rows = int(self.n_rows[locus]) if self.status[locus] == LOAD_OK else None
        out = [t.rstrip() for t in text.split("\n")]  # Unicode blanks too, as str.rstrip() in Biopython
        return out if rows is None or len(out) == rows else out + [""] * (rows - len(out))

    def ids(self, locus):
        """Record ids as Biopython cuts them: the first whitespace-separated token of the title."""
        out = []
        for title in self.titles(locus):
            tokens = title.split(None, 1)
            out.append(tokens[0] if tokens else "")
        return out

    def alignment(self, locus):
        """The locus as an MSA object (ids, descriptions, rows) for the host-side node classes."""
        # a private copy of the rows: the set's buffer goes back to the pinned pool on free()
        return MSA.from_matrix(self.ids(locus), self.matrix(locus).copy(), descriptions=self.titles(locus))

    def shapes(self):
        """int64[n_loci, 2] (rows, cols)."""
        return np.stack([self.n_rows.astype(np.int64), self.n_cols.astype(np.int64)], axis=1)

    def free(self):
        if self.handle is not None:
            self.ascii = self.offsets = self.n_rows = self.n_cols = self.status = self.flags = None
            self.packed = self.packed_offsets = self.alphabet_flags = None
            self.lib.mprg_fasta_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def replace_n(matrix):
    """io_utils.py:35-47 on an upper-cased uint8 matrix, in place, by the library (mprg_replace_n)."""
    assert matrix.dtype == np.uint8 and matrix.flags["C_CONTIGUOUS"]
    rc = _lib.load().mprg_replace_n(ptr(matrix), matrix.shape[0], matrix.shape[1])
    if rc != 0:
        raise MprgError(rc, "mprg_replace_n failed")


def replace_n_in_place(matrix):
    """The same through the Python mirror (random.Random itself): every N becomes the column's majority
    symbol, one `choice` per column in column order.  Kept as the checker of mprg_replace_n."""
    seqs = [row.tobytes().decode("ascii") for row in matrix]
    consensus = io_utils.majority_consensus_of_rows(seqs)
    cons = np.frombuffer(consensus.encode("ascii"), np.uint8)
    mask = matrix == ord("N")
    matrix[mask] = np.broadcast_to(cons, matrix.shape)[mask]


ALPHABET = np.frombuffer(b"-AMCSGWTNR?Y?K??", np.uint8)  # MPRG_ALPHABET: character of every 4-bit code


def packed_stride(cols):
    """Bytes of a packed row: 16 per chunk of 32 columns."""
    return (int(cols) + 31) // 32 * 16


def pack_rows(matrix):
    """uint8[rows, cols] ASCII -> (uint8[rows, stride] in the 4-bit device layout, alphabet flags), by the
    library's host packer (mprg_pack_rows)."""
    matrix = np.ascontiguousarray(matrix, np.uint8)
    rows, cols = matrix.shape
    out = np.empty((rows, packed_stride(cols)), np.uint8)
    flags = C.c_int32()
    rc = _lib.load().mprg_pack_rows(ptr(matrix) if matrix.size else None, rows, cols, ptr(out) if out.size else ptr(np.zeros(1, np.uint8)),
                                    out.size, C.byref(flags))
    if rc != 0:
        raise MprgError(rc, "mprg_pack_rows failed")
    return out, flags.value


def unpack_rows(packed, cols):
    """Inverse of the packing (numpy; test and archive helper): column c of a chunk is nibble c // 4 of the
    little-endian 32-bit word c % 4."""
    packed = np.ascontiguousarray(packed, np.uint8)
    rows = packed.shape[0]
    if cols == 0 or rows == 0:
        return np.zeros((rows, cols), np.uint8)
    words = packed.view("<u4").reshape(rows, -1, 4)                      # [row, chunk, word]
    shifts = (4 * np.arange(8, dtype=np.uint32)).reshape(1, 1, 8, 1)      # nibble j of every word
    codes = (words[:, :, None, :] >> shifts) & 15                         # [row, chunk, nibble j, word w]: column 4j + w
    return ALPHABET[codes.reshape(rows, -1)[:, :cols]]


def load_fasta_files(paths, threads=None, pin=True, packed=False, keep_ascii=False):
    """Parses every file on host threads (N replaced by the loader).  Loci whose status is not LOAD_OK
    are left to the caller (`raise_for_load_status` re-creates the reference's exception).
    packed: the matrices come out in the 4-bit device layout (for Context.build_msa_set / mprg_build_packed);
    the text is dropped unless keep_ascii."""
    lib = _lib.load()
    arr, _keep = _c_strings([os.fspath(p) for p in paths])
    h = C.c_void_p()
    mode = (1 if pin else 0) | (2 if packed else 0) | (4 if keep_ascii else 0)
    rc = lib.mprg_fasta_load(C.cast(arr, C.c_void_p), len(paths), threads or default_threads(), mode,
                             C.byref(h))
    if rc != 0:
        raise MprgError(rc, "mprg_fasta_load failed")
    return MsaSet(h, paths)


class NonAsciiSequenceError(ValueError):
    """The file parses as text but its rows hold non-ASCII characters: a curation error of that locus."""


def raise_for_load_status(msas, locus):
    """The exception load_alignment_file raises for this file (io_utils.py:17-29 through Biopython)."""
    status = int(msas.status[locus])
    if status == LOAD_OK:
        return
    if status == LOAD_NO_RECORDS:
        raise ValueError("No records found in handle")
    if status == LOAD_RAGGED:
        raise ValueError("Sequences must all be the same length")
    # unreadable, or bytes >= 0x80 among the sequences / a title that is not UTF-8: the Python loader raises
    # what the reference's read raises (OSError, UnicodeDecodeError, ...).  If it loads, the rows hold
    # characters outside the DNA alphabet: the reference skips such a locus (SequenceCurationError,
    # from_msa.py:147-151), which is what LOAD_NOT_ASCII then means to the caller.
    io_utils.load_alignment_file(msas.paths[locus], "fasta")
    if status == LOAD_NOT_ASCII:
        raise NonAsciiSequenceError(f"{msas.paths[locus]}: a sequence has a disallowed (non-ASCII) character")
    raise ValueError(f"{msas.paths[locus]} could not be loaded")


# ---- writers ------------------------------------------------------------------------------------
class EncodeError(Exception):
    pass


def _raise_encoding(rc, what):
    if rc == 1:
        raise EncodeError(f"{what} contains invalid characters")
    if rc == 2:
        raise ValueError("Prg error: odd site marker found >2 times")
    if rc == 3:
        raise OverflowError("int too big to convert")
    raise MprgError(rc, f"{what} failed")


def encode_prg(prg):
    """PRG string -> uint32 array (what PrgEncoder.write stores little-endian)."""
    lib = _lib.load()
    raw = prg.encode() if isinstance(prg, str) else bytes(prg)
    n = C.c_int64()
    out = np.zeros(max(len(raw), 1), np.uint32)  # one value per base or marker: never more than bytes
    rc = lib.mprg_encode_prg(raw, len(raw), ptr(out), out.size, C.byref(n))
    if rc != 0:
        _raise_encoding(rc, "Unit")
    return out[:n.value].copy()


def prg_to_gfa(prg):
    """PRG string -> GFA text, header included."""
    lib = _lib.load()
    raw = prg.encode() if isinstance(prg, str) else bytes(prg)
    n = C.c_int64()
    rc = lib.mprg_prg_to_gfa(raw, len(raw), None, 0, C.byref(n))
    if rc != 0:
        raise AssertionError("Invalid prg sequence")
    out = np.zeros(max(n.value, 1), np.uint8)
    lib.mprg_prg_to_gfa(raw, len(raw), ptr(out), out.size, C.byref(n))
    return out[:n.value].tobytes().decode()


class PrgStrings:
    """A result handle over plain PRG strings (mprg_result_from_prgs), for OutputWriter.add."""

    def __init__(self, prgs):
        self.lib = _lib.load()
        raw = [p.encode() if isinstance(p, str) else bytes(p) for p in prgs]
        arr = (C.c_char_p * max(len(raw), 1))(*raw)
        lens = np.array([len(r) for r in raw], np.int64)
        h = C.c_void_p()
        rc = self.lib.mprg_result_from_prgs(C.cast(arr, C.c_void_p), ptr(lens) if len(raw) else None, len(raw),
                                            C.byref(h))
        if rc != 0:
            raise MprgError(rc, "mprg_result_from_prgs failed")
        self.handle = h

    def free(self):
        if self.handle is not None:
            self.lib.mprg_result_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def merge_outputs(part_prefixes, output_prefix, prg=True, binary=True, gfa=True, update_ds=False):
    """Final .prg.fa / .prg.bin(.zip) / .prg.gfa(.zip) (and .update_DS.zip) from the parts of a sharded run
    (mprg_merge_outputs); returns the number of loci.  The parts are removed."""
    lib = _lib.load()
    what = (WRITE_PRG if prg else 0) | (WRITE_BIN if binary else 0) | (WRITE_GFA if gfa else 0)
    what |= WRITE_DS if update_ds else 0
    arr, _keep = _c_strings([os.fspath(p) for p in part_prefixes])
    n = C.c_int64()
    err = C.create_string_buffer(1024)
    rc = lib.mprg_merge_outputs(C.cast(arr, C.c_void_p), len(part_prefixes), os.fspath(output_prefix).encode(), what,
                                C.byref(n), err, len(err))
    if rc != 0:
        raise OSError(err.value.decode() or "mprg_merge_outputs failed")
    return n.value


class OutputWriter:
    """<prefix>.prg.fa / .prg.bin(.zip) / .prg.gfa(.zip) written by host threads of the library."""

    def __init__(self, output_prefix, prg=True, binary=True, gfa=True, threads=None, part=False):
        """part: this writer produces one part of a sharded run (archives even for one locus); the parts
        become the final files through merge_outputs."""
        self.lib = _lib.load()
        what = (WRITE_PRG if prg else 0) | (WRITE_BIN if binary else 0) | (WRITE_GFA if gfa else 0)
        what |= WRITE_PART if part else 0
        self.threads = threads or default_threads()
        h = C.c_void_p()
        rc = self.lib.mprg_writer_open(os.fspath(output_prefix).encode(), what, C.byref(h))
        if rc != 0:
            raise MprgError(rc, "mprg_writer_open failed")
        self.handle = h

    def add(self, result, loci, names):
        """result: BuildResult or PrgStrings; loci: indices into it; names: locus names (archive order)."""
        loci = np.ascontiguousarray(loci, np.int32)
        arr, _keep = _c_strings(names)
        rc = self.lib.mprg_writer_add(self.handle, result.handle, ptr(loci), C.cast(arr, C.c_void_p), len(loci),
                                      self.threads)
        if rc != 0:
            message = self.lib.mprg_writer_error(self.handle).decode()
            self.abort()
            if rc > 0:
                _raise_encoding(rc, message)
            raise OSError(message)

    def add_ds(self, result, msas, loci, names, max_nesting, min_match_length):
        """Appends the update data of these loci (tables: node table, row subsets, titles, packed root alignment,
        PRG) to <prefix>.update_DS.zip (mprg_writer_add_ds); msas: the MsaSet the result was built from."""
        loci = np.ascontiguousarray(loci, np.int32)
        arr, _keep = _c_strings(names)
        rc = self.lib.mprg_writer_add_ds(self.handle, result.handle, msas.handle, ptr(loci), C.cast(arr, C.c_void_p),
                                         len(loci), max_nesting, min_match_length, self.threads)
        if rc != 0:
            message = self.lib.mprg_writer_error(self.handle).decode()
            self.abort()
            raise OSError(message)

    def close(self):
        n, nbytes = C.c_int64(), C.c_int64()
        rc = self.lib.mprg_writer_close(self.handle, C.byref(n), C.byref(nbytes))
        if rc != 0:
            message = self.lib.mprg_writer_error(self.handle).decode()
            self.abort()
            raise OSError(message)
        self.handle = None
        return n.value, nbytes.value

    def abort(self):
        if self.handle is not None:
            self.lib.mprg_writer_abort(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.abort()
        except Exception:
            pass
