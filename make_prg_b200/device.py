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

    def set_workers(self, n):
        self._check(self.lib.mprg_set_workers(self.handle, int(n)))

    WAIT_MODES = {"spin": 0, "block": 1, "yield": 2}

    def set_wait_mode(self, mode):
        """How the host thread of a build waits at the level loop's synchronisation points (mprg_set_wait_mode):
        "spin" (cudaStreamSynchronize), "block" (sleep on an event), "yield" (poll + sched_yield)."""
        self._check(self.lib.mprg_set_wait_mode(self.handle, self.WAIT_MODES[mode]))

    PATHS = ("kmeans_cta", "kmeans_group", "refcheck_cta", "refcheck_grid", "refcheck_grid_multi", "kmer_grid",
             "dedupe_grid")

    def path_counts(self, reset=False):
        """Launches per kernel variant since the last reset (mprg_path_counts)."""
        out = np.zeros(8, np.int64)
        self._check(self.lib.mprg_path_counts(self.handle, ptr(out), int(reset)))
        return dict(zip(self.PATHS, out.tolist()))

    def kmeans_stats(self, reset=False):
        """Device ms / launches / problems of the level loop's KMeans launches since the last reset."""
        ms, n, p = C.c_double(), C.c_int64(), C.c_int64()
        self._check(self.lib.mprg_kmeans_stats(self.handle, C.byref(ms), C.byref(n), C.byref(p), int(reset)))
        return {"ms": ms.value, "launches": n.value, "problems": p.value}

    def copy_stats(self, reset=False):
        a, b = C.c_int64(), C.c_int64()
        self._check(self.lib.mprg_copy_stats(self.handle, C.byref(a), C.byref(b), int(reset)))
        return {"h2d_bytes": a.value, "d2h_bytes": b.value}

    def timer_start(self):
        self._check(self.lib.mprg_timer(self.handle, 0, None))

    def timer_stop(self):
        ms = C.c_double()
        self._check(self.lib.mprg_timer(self.handle, 1, C.byref(ms)))
        return ms.value

    def scan_log(self, reset=False, capacity=65536):
        by = np.zeros(capacity, np.float64)
        ms = np.zeros(capacity, np.float64)
        n = C.c_int32()
        self._check(self.lib.mprg_scan_log(self.handle, ptr(by), ptr(ms), capacity, C.byref(n), int(reset)))
        return by[:n.value].copy(), ms[:n.value].copy()

    def read_yardstick(self, batch):
        """(bytes, ms) of one bare streaming read of the packed batch (measurement aid, mprg.h)."""
        by, ms = C.c_double(), C.c_double()
        self._check(self.lib.mprg_read_yardstick(self.handle, batch.handle, C.byref(by), C.byref(ms)))
        return by.value, ms.value

    # ---- loader -> HBM ------------------------------------------------------------------------
    @staticmethod
    def _flatten(matrices):
        if isinstance(matrices, tuple):
            flat, shapes = matrices
            shape_arr = np.asarray(shapes, np.int64).reshape(-1, 2)
        else:
            shape_arr = np.array([m.shape for m in matrices], np.int64).reshape(-1, 2)
            flat = (np.concatenate([np.ascontiguousarray(m, np.uint8).reshape(-1) for m in matrices])
                    if matrices else np.zeros(0, np.uint8))
        sizes = shape_arr[:, 0] * shape_arr[:, 1]
        offsets = np.zeros(len(shape_arr), np.int64)
        if len(shape_arr):
            offsets[1:] = np.cumsum(sizes)[:-1]
        n_rows = np.ascontiguousarray(shape_arr[:, 0], np.int32)
        n_cols = np.ascontiguousarray(shape_arr[:, 1], np.int32)
        if flat.size == 0:
            flat = np.zeros(1, np.uint8)
        # shapes: indexable per locus as (rows, cols); an int64[n, 2] array, no per-locus Python objects
        return flat, shape_arr, offsets, n_rows, n_cols

    def upload(self, matrices):
        """matrices: list of uint8[rows, cols] ASCII arrays (or one flat buffer + shapes tuple)."""
        flat, shapes, offsets, n_rows, n_cols = self._flatten(matrices)
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

    # ---- clustering ----------------------------------------------------------------------------
    def _row_offsets(self, arr):
        offs = np.zeros(len(arr), np.int64)
        if len(arr):
            offs[1:] = np.cumsum(arr["n_rows"].astype(np.int64))[:-1]
        return offs, int(arr["n_rows"].astype(np.int64).sum())

    def dedupe_rows(self, batch, tasks):
        """-> per task (group int32[], ungapped_len int32[], n_unique_ungapped, n_unique_gapped)."""
        arr, arena = self.make_tasks(batch, tasks)
        offs, total = self._row_offsets(arr)
        group = np.zeros(max(total, 1), np.int32)
        ulen = np.zeros(max(total, 1), np.int32)
        nu = np.zeros(max(len(arr), 1), np.int32)
        ng = np.zeros(max(len(arr), 1), np.int32)
        self._check(self.lib.mprg_dedupe_rows(self.handle, batch.handle, ptr(arr), len(arr),
                                              ptr(arena) if arena.size else None, arena.size,
                                              ptr(offs), ptr(group), ptr(ulen), ptr(nu), ptr(ng)))
        return [(group[o:o + n].copy(), ulen[o:o + n].copy(), int(nu[i]), int(ng[i]))
                for i, (o, n) in enumerate(zip(offs, arr["n_rows"]))]

    def kmer_counts(self, batch, task, kmer_size):
        """-> float64[n_distinct_long, n_kmers] (kernel (b))."""
        arr, arena = self.make_tasks(batch, [task])
        n, F = C.c_int32(0), C.c_int32(0)
        rows = ptr(arena) if arena.size else None
        self._check(self.lib.mprg_kmer_counts(self.handle, batch.handle, ptr(arr), rows, kmer_size,
                                              C.byref(n), C.byref(F), None, 0))
        X = np.zeros((n.value, F.value), np.float64)
        if X.size:
            self._check(self.lib.mprg_kmer_counts(self.handle, batch.handle, ptr(arr), rows, kmer_size,
                                                  C.byref(n), C.byref(F), ptr(X), X.size))
        return X

    def kmeans(self, X, K, mode=0):
        """KMeans(K, random_state=2, elkan, n_init=10).fit(X).predict(X) -> (labels, inertia).
        mode 0: the engine's choice, 1: one CTA per initialisation, 2: CTA groups (deep loci)."""
        X = np.ascontiguousarray(X, np.float64)
        labels = np.zeros(X.shape[0], np.int32)
        inertia = C.c_double(0)
        self._check(self.lib.mprg_kmeans_mode(self.handle, ptr(X), X.shape[0], X.shape[1], K, ptr(labels),
                                              C.byref(inertia), mode))
        return labels, inertia.value

    def one_ref_like(self, batch, task, cluster_of_row, n_clusters):
        arr, arena = self.make_tasks(batch, [task])
        cl = np.ascontiguousarray(cluster_of_row, np.int32)
        flags = np.zeros(n_clusters, np.int32)
        self._check(self.lib.mprg_one_ref_like(self.handle, batch.handle, ptr(arr),
                                               ptr(arena) if arena.size else None, ptr(cl), n_clusters,
                                               ptr(flags)))
        return flags.astype(bool)

    def cluster_tasks(self, batch, tasks, kmer_size):
        """kmeans_cluster_seqs per task -> list of clustered row positions (ClusteringResult order)."""
        arr, arena = self.make_tasks(batch, tasks)
        offs, total = self._row_offsets(arr)
        cluster = np.zeros(max(total, 1), np.int32)
        ncl = np.zeros(max(len(arr), 1), np.int32)
        self._check(self.lib.mprg_cluster_tasks(self.handle, batch.handle, ptr(arr), len(arr),
                                                ptr(arena) if arena.size else None, arena.size,
                                                kmer_size, ptr(offs), ptr(cluster), ptr(ncl)))
        out = []
        for i, (o, n) in enumerate(zip(offs, arr["n_rows"])):
            c = cluster[o:o + n]
            out.append([np.nonzero(c == k)[0].tolist() for k in range(int(ncl[i]))])
        return out

    # ---- the whole path ------------------------------------------------------------------------
    def build(self, batch, max_nesting, min_match_length):
        h = C.c_void_p()
        self._check(self.lib.mprg_build(self.handle, batch.handle, max_nesting, min_match_length,
                                        C.byref(h)))
        return BuildResult(self, h)

    def build_sub(self, batch, max_nesting, min_match_length, parent_levels):
        """Every locus of the batch built below a node of nesting level parent_levels[l] (-1: as a root):
        NodeFactory.build(alignment, builder, parent_node) for a batch of re-builds (mprg_build_sub)."""
        levels = np.ascontiguousarray(parent_levels, np.int32)
        assert len(levels) == batch.n_loci
        h = C.c_void_p()
        self._check(self.lib.mprg_build_sub(self.handle, batch.handle, max_nesting, min_match_length,
                                            ptr(levels), C.byref(h)))
        return BuildResult(self, h)

    def build_ascii(self, matrices, max_nesting, min_match_length):
        """Host ASCII in, (Batch, BuildResult) out in one call: every worker range is copied, packed and
        built on its own stream, so the copies overlap the kernels (mprg_build_ascii)."""
        flat, shapes, offsets, n_rows, n_cols = self._flatten(matrices)
        hb, hr = C.c_void_p(), C.c_void_p()
        self._check(self.lib.mprg_build_ascii(self.handle, ptr(flat), ptr(offsets), ptr(n_rows), ptr(n_cols),
                                              len(shapes), max_nesting, min_match_length, C.byref(hb),
                                              C.byref(hr)))
        return Batch(self, hb, shapes), BuildResult(self, hr)


    def build_packed(self, packed, offsets, n_rows, n_cols, flags, max_nesting, min_match_length):
        """Host rows already in the 4-bit device layout (hostio.pack_rows / a packed MsaSet) in, (Batch,
        BuildResult) out: half the bytes of build_ascii cross PCIe, no pack kernel (mprg_build_packed)."""
        offsets = np.ascontiguousarray(offsets, np.int64)
        n_rows = np.ascontiguousarray(n_rows, np.int32)
        n_cols = np.ascontiguousarray(n_cols, np.int32)
        flags = np.ascontiguousarray(flags, np.int32)
        hb, hr = C.c_void_p(), C.c_void_p()
        self._check(self.lib.mprg_build_packed(self.handle, ptr(packed), ptr(offsets), ptr(n_rows), ptr(n_cols),
                                               ptr(flags), len(n_rows), max_nesting, min_match_length, C.byref(hb),
                                               C.byref(hr)))
        shapes = np.stack([n_rows.astype(np.int64), n_cols.astype(np.int64)], axis=1)
        return Batch(self, hb, shapes), BuildResult(self, hr)

    def build_msa_set(self, msas, max_nesting, min_match_length):
        """The same straight from the native loader's buffers (hostio.MsaSet): no copy on the host."""
        if msas.packed is not None:
            return self.build_packed(msas.packed, msas.packed_offsets, msas.n_rows, msas.n_cols, msas.alphabet_flags,
                                     max_nesting, min_match_length)
        hb, hr = C.c_void_p(), C.c_void_p()
        self._check(self.lib.mprg_build_ascii(self.handle, ptr(msas.ascii), ptr(msas.offsets), ptr(msas.n_rows),
                                              ptr(msas.n_cols), msas.n_loci, max_nesting, min_match_length,
                                              C.byref(hb), C.byref(hr)))
        return Batch(self, hb, msas.shapes()), BuildResult(self, hr)


class BuildResult:
    """Trees and PRG strings of one mprg_build call."""

    def __init__(self, ctx, handle):
        self.ctx, self.handle = ctx, handle
        self.n_loci = ctx.lib.mprg_result_n_loci(handle)

    def status(self, locus):
        return int(self.ctx.lib.mprg_result_status(self.handle, locus))

    def statuses(self):
        """(status int32[n_loci], PRG length int64[n_loci]) in one call."""
        st = np.zeros(max(self.n_loci, 1), np.int32)
        ln = np.zeros(max(self.n_loci, 1), np.int64)
        self.ctx._check(self.ctx.lib.mprg_result_statuses(self.handle, ptr(st), ptr(ln)))
        return st[:self.n_loci], ln[:self.n_loci]

    def prg(self, locus):
        n = C.c_int64(0)
        p = self.ctx.lib.mprg_result_prg(self.handle, locus, C.byref(n))
        return C.string_at(p, n.value).decode() if p and n.value else ""

    def n_nodes(self, locus):
        return int(self.ctx.lib.mprg_result_n_nodes(self.handle, locus))

    def n_sites(self, locus):
        return int(self.ctx.lib.mprg_result_n_sites(self.handle, locus))

    def nodes(self, locus):
        """Pre-order node table: dict of arrays + the locus row pool."""
        n = self.n_nodes(locus)
        cols = {k: np.zeros(max(n, 1), np.int32) for k in
                ("kind", "parent", "nesting_level", "c0", "c1", "n_rows", "n_children")}
        row_off = np.zeros(max(n, 1), np.int64)
        lib = self.ctx.lib
        self.ctx._check(lib.mprg_result_nodes(self.handle, locus, ptr(cols["kind"]), ptr(cols["parent"]),
                                              ptr(cols["nesting_level"]), ptr(cols["c0"]), ptr(cols["c1"]),
                                              ptr(cols["n_rows"]), ptr(row_off), ptr(cols["n_children"])))
        size = int(lib.mprg_result_row_pool_size(self.handle, locus))
        pool = np.zeros(max(size, 1), np.int32)
        if size:
            self.ctx._check(lib.mprg_result_row_pool(self.handle, locus, ptr(pool)))
        out = {k: v[:n] for k, v in cols.items()}
        out["row_off"] = row_off[:n]
        out["row_pool"] = pool[:size]
        return out

    def free(self):
        if self.handle is not None:
            self.ctx.lib.mprg_result_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def default_lanes():
    """Builds in flight per GPU.  Each lane is a host thread that launches the kernels of its build and waits at
    two synchronisation points per recursion level, so the count follows this process's share of the host cores
    (its CPU affinity, divided among LOCAL_WORLD_SIZE ranks): 6 lanes from 7 cores up, 3 on a 4-core share (eight
    ranks on a 32-core host).  Measured per 1,000-locus step (profiles/r2_lanes_sweep.txt; one build at a time:
    2.7 ms resident, 4.6 ms from host rows): 6 lanes 1.3-1.4 ms / 2.2-2.3 ms on 16-24 cores, 3 lanes on a 4-core
    share 1.8 / 2.7 ms.  MPRG_BUILD_LANES overrides."""
    import os

    env = os.environ.get("MPRG_BUILD_LANES")
    if env:
        return max(1, int(env))
    try:
        cores = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        cores = os.cpu_count() or 1
    share = cores // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
    return max(3, min(6, share - 1))


class BuildPipeline:
    """Several builds in flight on one GPU: `depth` lanes, each a Context (own stream, own scratch, own packed
    arenas) driven by its own host thread.  A build from host rows is upload -> level loop -> PRG strings back;
    a lane runs these in order and the lanes are staggered, so the host-to-device copy of build k+1 crosses PCIe
    while build k is in its level loop (the library lets one big upload at a time onto the link), and the idle
    gaps of one level loop (two host synchronisations per recursion level) are filled by the other's kernels.
    Every submission is still one C-ABI call (`mprg_build_packed` / `mprg_build_ascii`), a lane builds its batch
    as one range.  `consume(batch, result)` runs on the lane's thread as soon as the build is done (ctypes calls
    release the GIL); batch and result are freed when it returns and the future carries its return value; without
    `consume` the future carries (batch, result) and the caller frees them.
    Submissions complete in order per lane, results are independent of the lane that built them."""

    def __init__(self, device=0, depth=None):
        from concurrent.futures import ThreadPoolExecutor

        if depth is None:
            depth = default_lanes()
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.depth = depth
        self.contexts = [Context(device) for _ in range(depth)]
        import os

        # a waiting lane polls and yields its core (mprg_set_wait_mode): with plain spinning, 2 ranks x 6 lanes on a
        # 24-core host ran a resident step in 3.2 ms instead of 1.3, sleeping on an event in 2.5
        wait = os.environ.get("MPRG_LANE_WAIT", "yield") if depth > 1 else "spin"
        for c in self.contexts:
            c.set_workers(1)  # one range per build: the overlap comes from the other lanes
            c.set_wait_mode(wait)
        self._lanes = [ThreadPoolExecutor(1, thread_name_prefix=f"mprg-lane{i}") for i in range(depth)]
        self._next = 0

    @property
    def next_lane(self):
        """The lane the next submission goes to (round robin)."""
        return self._next % self.depth

    def _submit(self, call, consume, owns_batch=True):
        lane = self.next_lane
        self._next += 1
        ctx = self.contexts[lane]

        def job():
            batch, res = call(ctx)
            if consume is None:
                return batch, res  # the caller frees both (from any thread)
            try:
                return consume(batch, res)
            finally:
                res.free()
                if owns_batch:
                    batch.free()

        return self._lanes[lane].submit(job)

    def submit_resident(self, batch, max_nesting, min_match_length, consume=None):
        """Context.build of a batch that is already in HBM (any context of this GPU may build it; the batch stays
        the caller's) -> Future of consume(batch, result), or of (batch, result)."""
        return self._submit(lambda c: (batch, c.build(batch, max_nesting, min_match_length)), consume,
                            owns_batch=False)

    def submit_packed(self, packed, offsets, n_rows, n_cols, flags, max_nesting, min_match_length, consume=None):
        """Context.build_packed on the next lane -> Future of consume(batch, result)."""
        return self._submit(lambda c: c.build_packed(packed, offsets, n_rows, n_cols, flags, max_nesting,
                                                     min_match_length), consume)

    def submit_ascii(self, matrices, max_nesting, min_match_length, consume=None):
        return self._submit(lambda c: c.build_ascii(matrices, max_nesting, min_match_length), consume)

    def submit_msa_set(self, msas, max_nesting, min_match_length, consume=None):
        return self._submit(lambda c: c.build_msa_set(msas, max_nesting, min_match_length), consume)

    def launch_count(self):
        return sum(c.launch_count() for c in self.contexts)

    def copy_stats(self, reset=False):
        out = {"h2d_bytes": 0, "d2h_bytes": 0}
        for c in self.contexts:
            for k, v in c.copy_stats(reset).items():
                out[k] += v
        return out

    def close(self):
        for lane in self._lanes:
            lane.shutdown(wait=True)
        for c in self.contexts:
            c.close()
        self.contexts = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


_default = {}


_pipelines = {}


def default_pipeline(device=0, depth=None):
    """The process-wide BuildPipeline of a GPU with this many lanes (made on first use)."""
    if depth is None:
        depth = default_lanes()
    key = (device, depth)
    if key not in _pipelines:
        _pipelines[key] = BuildPipeline(device, depth)
    return _pipelines[key]


def default_context(device=0):
    if device not in _default:
        _default[device] = Context(device)
    return _default[device]
