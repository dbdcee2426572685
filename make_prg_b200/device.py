"""Numpy-facing wrapper over the C ABI: one Context per GPU, Batches of loci resident in HBM."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import INTERVAL_DTYPE, TASK_DTYPE, MprgError, ptr


class Batch:
    def __init__(self, ctx, handle, shapes):
        self.ctx, self.handle, self.shapes = ctx, handle, shapes

    @property
    def n_loci(self):
        return len(self.shapes)

    def flags(self):
        out = np.zeros(self.n_loci, np.int32)
        self.ctx._check(self.ctx.lib.mprg_batch_flags(self.ctx.handle, self.handle, ptr(out)))
        return out

    def packed(self, locus):
        stride = C.c_int32(0)
        lib, ctx = self.ctx.lib, self.ctx
        ctx._check(lib.mprg_batch_download_packed(ctx.handle, self.handle, locus, None, 0, C.byref(stride)))
        out = np.zeros((self.shapes[locus][0], stride.value), np.uint8)
        ctx._check(lib.mprg_batch_download_packed(ctx.handle, self.handle, locus, ptr(out), out.size,
                                                  C.byref(stride)))
        return out

    def free(self):
        if self.handle is not None:
            self.ctx.lib.mprg_batch_free(self.ctx.handle, self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """Owns one mprg_ctx (one GPU, one stream).  Fails loudly without a CUDA device."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.mprg_create(device, C.byref(h))
        if rc != 0:
            raise MprgError(rc, f"mprg_create(device={device}) failed; a CUDA device is required")
        self.handle = h

    def close(self):
        if self.handle is not None:
            self.lib.mprg_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise MprgError(rc, self.lib.mprg_last_error(self.handle).decode())

    def device_info(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.mprg_device_info(self.handle, C.byref(a), C.byref(b), C.byref(c)))
        return {"sm_count": a.value, "cc": (b.value, c.value)}

    def launch_count(self):
        return int(self.lib.mprg_launch_count(self.handle))

    def scan_stats(self, reset=False):
        ms, by, n = C.c_double(), C.c_double(), C.c_int64()
        self._check(self.lib.mprg_scan_stats(self.handle, C.byref(ms), C.byref(by), C.byref(n), int(reset)))
        return {"ms": ms.value, "bytes": by.value, "launches": n.value}

    # ---- loader -> HBM ------------------------------------------------------------------------
    def upload(self, matrices):
        """matrices: list of uint8[rows, cols] ASCII arrays (or one flat buffer + shapes tuple)."""
        if isinstance(matrices, tuple):
            flat, shapes = matrices
            shapes = [tuple(s) for s in shapes]
        else:
            shapes = [tuple(m.shape) for m in matrices]
            flat = (np.concatenate([np.ascontiguousarray(m, np.uint8).reshape(-1) for m in matrices])
                    if matrices else np.zeros(0, np.uint8))
        sizes = np.array([r * c for r, c in shapes], np.int64)
        offsets = np.zeros(len(shapes), np.int64)
        if len(shapes):
            offsets[1:] = np.cumsum(sizes)[:-1]
        n_rows = np.array([s[0] for s in shapes], np.int32)
        n_cols = np.array([s[1] for s in shapes], np.int32)
        if flat.size == 0:
            flat = np.zeros(1, np.uint8)
        h = C.c_void_p()
        self._check(self.lib.mprg_batch_upload(self.handle, ptr(flat), ptr(offsets), ptr(n_rows),
                                               ptr(n_cols), len(shapes), C.byref(h)))
        return Batch(self, h, shapes)

    # ---- task helpers -----------------------------------------------------------------------
    @staticmethod
    def make_tasks(batch, tasks):
        """tasks: iterable of (locus, rows-or-None, c0, c1) -> (task array, row arena)."""
        arr = np.zeros(len(tasks), TASK_DTYPE)
        pool = []
        off = 0
        for i, (locus, rows, c0, c1) in enumerate(tasks):
            if rows is None:
                arr[i] = (locus, -1, batch.shapes[locus][0], c0, c1)
            else:
                rows = np.asarray(rows, np.int32)
                arr[i] = (locus, off, len(rows), c0, c1)
                pool.append(rows)
                off += len(rows)
        arena = np.concatenate(pool).astype(np.int32) if pool else np.zeros(0, np.int32)
        return arr, arena

    def scan_tasks(self, batch, tasks):
        """-> list of (consensus bytes, gap_reach int32[]) per task (kernel (a))."""
        arr, arena = self.make_tasks(batch, tasks)
        widths = (arr["c1"] - arr["c0"]).astype(np.int64)
        offs = np.zeros(len(arr), np.int64)
        if len(arr):
            offs[1:] = np.cumsum(widths)[:-1]
        total = int(widths.sum())
        cons = np.zeros(max(total, 1), np.uint8)
        reach = np.zeros(max(total, 1), np.int32)
        self._check(self.lib.mprg_scan_tasks(self.handle, batch.handle, ptr(arr), len(arr),
                                             ptr(arena) if arena.size else None, arena.size,
                                             ptr(offs), ptr(cons), ptr(reach)))
        return [(cons[o:o + w].tobytes(), reach[o:o + w].copy()) for o, w in zip(offs, widths)]

    def partition_tasks(self, batch, tasks, min_match_length):
        """-> per task a structured array of intervals (start, stop, type) sorted by start."""
        arr, arena = self.make_tasks(batch, tasks)
        caps = np.maximum(arr["c1"] - arr["c0"], 1).astype(np.int64)
        offs = np.zeros(len(arr), np.int64)
        if len(arr):
            offs[1:] = np.cumsum(caps)[:-1]
        iv = np.zeros(max(int(caps.sum()), 1), INTERVAL_DTYPE)
        cnt = np.zeros(max(len(arr), 1), np.int32)
        self._check(self.lib.mprg_partition_tasks(self.handle, batch.handle, ptr(arr), len(arr),
                                                  ptr(arena) if arena.size else None, arena.size,
                                                  min_match_length, ptr(offs), ptr(iv), ptr(cnt)))
        return [iv[o:o + n].copy() for o, n in zip(offs, cnt[:len(arr)])]

    def partition_consensus(self, consensus, min_match_length, gap_reach=None):
        cons = np.frombuffer(consensus.encode() if isinstance(consensus, str) else consensus, np.uint8)
        n = len(cons)
        iv = np.zeros(max(n, 1), INTERVAL_DTYPE)
        cnt = C.c_int32(0)
        reach = None if gap_reach is None else np.ascontiguousarray(gap_reach, np.int32)
        self._check(self.lib.mprg_partition_consensus(self.handle, ptr(cons) if n else None, ptr(reach),
                                                      n, min_match_length, ptr(iv), len(iv),
                                                      C.byref(cnt)))
        return iv[:cnt.value].copy()


_default = {}


def default_context(device=0):
    if device not in _default:
        _default[device] = Context(device)
    return _default[device]
