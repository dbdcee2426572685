#!/usr/bin/env python
"""bench.py -- `from_msa` hot path on B200: MSA loci/sec with byte-identical-PRG semantics.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): synthetic 1,000-locus set, 200 seqs x 1 kb MSAs, -N 5 -L 7, per
GPU.  A step = one pass of the whole hot path (level-synchronous scan / partition / clustering /
KMeans / PRG emission, `mprg_build`) over that batch.
  value     loci/sec with the packed batches already resident in HBM when the timed region starts: K steps
            issued back to back through device.BuildPipeline (several builds in flight on their own streams,
            one C-ABI call per step), host clock between barrier + synchronize; `one_at_a_time` = one build at
            a time on one stream (device time per step, L2 flushed between steps)
  e2e       loci/sec through the public API from pinned HOST buffers in the loader's 4-bit layout: H2D copy
            + build + PRG strings back on the host, every step, the same pipeline (the upload of step k+1
            overlaps the level loop of step k); `e2e.one_at_a_time` = the latency of one such step;
            from_text: the same from ASCII rows
  roofline  the dominant kernel (column scan, root level launch): algorithmic bytes / CUDA-event time
  files     the same metric file to file (FASTA files in page cache -> .prg.fa/.bin.zip/.gfa.zip on tmpfs)
            through the native loader and writers (scripts/files_e2e.py), host wall clock
  cpu_baseline  the oracle port (oracle/make_prg_oracle.py) on a bounded sample, 1 core
`--impl reference` times the UNMODIFIED reference (`make_prg from_msa -t <all host cores>`, staged into
git-ignored oracle/_ref by oracle/stage_reference.py so that it travels to the GPU box) on a bounded sample
of the same loci per step; the oracle port is the fallback when nothing is staged.
N > 1: launched by torchrun, one rank per GPU, loci sharded (weak scaling: 1,000 loci per GPU, no
data-path collective), time = max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

LOCI_PER_GPU = 1000
ROWS, COLS = 200, 1000
MAX_NESTING, MIN_MATCH = 5, 7


def ncu_counters(name, wanted):
    """Counters of the committed `ncu --set full` summary profiles/r2_ncu_<name>.txt (scripts/ncu_summary.py
    output: metric, value, unit per line); {} when the file is not there."""
    f = REPO / "profiles" / f"r2_ncu_{name}.txt"
    out = {}
    if not f.exists():
        return out
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
    for line in f.read_text().splitlines():
        parts = line.split("\t")
        if len(parts) >= 2 and parts[0] in wanted:
            try:
                out[parts[0]] = float(parts[1].replace(",", "")) * scale.get(parts[2] if len(parts) > 2 else "", 1.0)
            except ValueError:
                pass
    return out


def ncu_scan_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one root-level scan launch of this workload, read from
    the committed capture (per launch, like `achieved`); None if not captured."""
    c = ncu_counters("scan_kernel", ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    return sum(c.values()) if len(c) == 2 else None
CACHE = Path(os.environ.get("MPRG_BENCH_CACHE", "/tmp/mprg_bench_cache"))
CPU_BASELINE_SAMPLE = 100


def workload(rank, n_loci=LOCI_PER_GPU):
    """Config #2 loci [rank*n_loci, (rank+1)*n_loci) as one uint8[n_loci, ROWS, COLS] array (cached)."""
    from make_prg_b200 import synth

    CACHE.mkdir(parents=True, exist_ok=True)
    f = CACHE / f"config2_{rank}_{n_loci}.npy"
    if f.exists():
        return np.load(f)
    out = np.empty((n_loci, ROWS, COLS), np.uint8)
    for i in range(n_loci):
        out[i] = synth.config_msa(2, rank * n_loci + i)
    tmp = f.with_suffix(f".{os.getpid()}.tmp.npy")
    np.save(tmp, out)
    os.replace(tmp, f)
    return out


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md), read through NVML in
    a thread of this process every 20 ms (spawning nvidia-smi from every rank stalls the driver for
    tens of milliseconds on a multi-GPU box and would land inside the timed steps); nvidia-smi is the
    fallback when NVML cannot be loaded."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NVML_REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                    0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop = threading.Event()
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            # NVML indexes physical GPUs: honour CUDA_VISIBLE_DEVICES when it is a list of indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = gpu_index
            if vis and all(t.strip().isdigit() for t in vis.split(",")):
                idx = int(vis.split(",")[gpu_index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)))
        try:
            mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        for bit, name in self.NVML_REASONS.items():
            if mask & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                              "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
        for line in out.strip().splitlines():
            r = [c.strip() for c in line.split(",")]
            self.sm.append(float(r[1]))
            self.mx.append(float(r[2]))
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    self.reasons.add(name)

    def _run(self):
        while not self.stop.is_set():
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self.stop.wait(0.02 if self.nvml is not None else 0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": float(max(self.mx)),
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def hbm_peak():
    f = REPO / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def oracle_loci_per_sec(mats, n_workers, keep=None):
    """Times the oracle port (CPU restatement of the reference algorithm) on `mats`; keep: list that
    receives the PRG strings (single-core run only)."""
    sys.path.insert(0, str(REPO / "oracle"))
    t0 = time.perf_counter()
    if n_workers <= 1:
        import make_prg_oracle as mo

        for M in mats:
            prg, _ = mo.build_prg_from_matrix([f"s{i}" for i in range(M.shape[0])], M, MAX_NESTING, MIN_MATCH)
            if keep is not None:
                keep.append(prg)
    else:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(n_workers) as pool:
            pool.map(_oracle_one, list(mats), chunksize=1)
    dt = time.perf_counter() - t0
    return len(mats) / dt, dt


def _oracle_one(M):
    os.environ["OMP_NUM_THREADS"] = "1"
    sys.path.insert(0, str(REPO / "oracle"))
    import make_prg_oracle as mo

    mo.build_prg_from_matrix([f"s{i}" for i in range(M.shape[0])], M, MAX_NESTING, MIN_MATCH)
    return 0


def reference_sample_size(cores):
    """Loci per step of the reference arm: the unmodified reference does about 0.5-1 locus/s/core on this
    workload, so 2 loci per core keep a step at a few seconds and a 25-step run at a few minutes."""
    return max(2 * cores, 16)


def run_reference(args, rank, world):
    """--impl reference: the UNMODIFIED reference (`make_prg from_msa -t <all host cores>`, staged under
    oracle/_ref by oracle/stage_reference.py, driven through its own CLI entry under the harness of
    oracle/run_reference.py) on the first loci of the same workload, FASTA files in -> PRG files out.
    Falls back to the oracle port (kind "port") only when no staged reference is present."""
    if rank != 0:
        return
    os.environ["OMP_NUM_THREADS"] = "1"
    cores = os.cpu_count() or 1
    sample = reference_sample_size(cores)
    data = workload(0)[:sample]
    sys.path.insert(0, str(REPO / "oracle"))
    import run_reference as rr

    kind = "reference" if rr.reference_available() else "port"
    checked = None
    if kind == "reference":
        import shutil
        import tempfile

        tmp = Path(tempfile.mkdtemp(prefix="mprg_ref_"))
        (tmp / "msas").mkdir()
        for i in range(sample):
            with open(tmp / "msas" / f"locus{i:05d}.fa", "w") as f:
                for r in range(data[i].shape[0]):
                    f.write(f">s{r}\n{data[i][r].tobytes().decode()}\n")

        def one_step(n):
            src = tmp / "msas"
            if n < sample:
                src = tmp / "warm"
                if not src.exists():
                    src.mkdir()
                    for i in range(n):
                        shutil.copy(tmp / "msas" / f"locus{i:05d}.fa", src)
            t0 = time.perf_counter()
            saved = os.dup(1)  # stdout carries exactly one JSON line
            os.dup2(2, 1)
            try:
                rr.ref_cli(["from_msa", "-i", src, "-o", tmp / "out", "-N", MAX_NESTING, "-L", MIN_MATCH,
                            "-t", cores, "-F"])
            finally:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)
            return time.perf_counter() - t0

        try:
            for _ in range(args.warmup):
                one_step(min(cores, sample))
            t_total = sum(one_step(sample) for _ in range(args.steps))
            # the reference's PRGs of the sample against the oracle port's (what the GPU arm is checked against)
            import make_prg_oracle as mo

            lines = (tmp / "out.prg.fa").read_text().split("\n")
            ref_prgs = {lines[i][1:]: lines[i + 1] for i in range(0, len(lines) - 1, 2)}
            n_chk = min(8, sample)
            checked = all(
                ref_prgs.get(f"locus{i:05d}") ==
                mo.build_prg_from_matrix([f"s{r}" for r in range(data[i].shape[0])], data[i], MAX_NESTING, MIN_MATCH)[0]
                for i in range(n_chk))
            if not checked:
                raise SystemExit("reference arm: the reference's PRGs differ from the oracle port's")
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        how = (f"{sample} loci of the workload per step as FASTA files through `make_prg from_msa -N {MAX_NESTING} "
               f"-L {MIN_MATCH} -t {cores} -F` of the unmodified reference (oracle/_ref, Biopython stand-in, "
               f"KMeans n_init=10, OMP_NUM_THREADS=1), files in -> .prg.fa/.bin/.gfa/update_DS out")
    else:
        mats = [data[i] for i in range(sample)]
        for _ in range(args.warmup):
            oracle_loci_per_sec(mats[:cores], cores)
        t_total = 0.0
        for _ in range(args.steps):
            _, dt = oracle_loci_per_sec(mats, cores)
            t_total += dt
        how = f"{sample} loci of the workload per step, oracle port, multiprocessing over loci"
    value = sample * args.steps / t_total
    line = {
        "impl": "reference", "metric": "MSA loci/sec (from_msa, byte-identical PRG)", "value": value,
        "unit": "loci/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config_dict(),
        "cpu_baseline": {"value": value, "unit": "loci/s", "cores": cores, "kind": kind, "sample": how,
                         "prgs_equal_oracle_port": checked},
        "e2e": {"value": value, "unit": "loci/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def config_dict():
    """The same object in both arms (the arms differ in what they time, not in the workload)."""
    cores = os.cpu_count() or 1
    from make_prg_b200.device import default_lanes  # (a rule on the host's core count; loads no library)

    lanes = default_lanes()
    return {"workload": "BASELINE configs[1]: synthetic 1,000-locus set, 200 seqs x 1 kb MSAs, -N 5 -L 7, "
                        "per GPU (make_prg_b200.synth seeds 1000+i)",
            "loci_per_gpu": LOCI_PER_GPU, "rows": ROWS, "cols": COLS, "max_nesting": MAX_NESTING,
            "min_match_length": MIN_MATCH,
            "builds_in_flight": lanes,
            "l2": f"throughput steps (value, e2e): {lanes} builds in flight on {lanes} input arenas of 100 MB each "
                  f"(> the 126 MB L2) plus one 256 MiB flush write per step on a third stream; one-at-a-time steps: "
                  f"L2 flushed between steps by writing a 256 MiB device buffer (packed batch is 100 MB < L2)",
            "cpu_sample": f"CPU arms time a bounded sample of the same loci: --impl reference the first "
                          f"{reference_sample_size(cores)} loci per step on {cores} cores, cpu_baseline the "
                          f"first {CPU_BASELINE_SAMPLE} loci on 1 core"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--loci", type=int, default=LOCI_PER_GPU)
    ap.add_argument("--no-big", action="store_true", help="skip the 8x-size launch of the roofline pass")
    ap.add_argument("--no-files", action="store_true", help="skip the file-to-file (FASTA in, PRG files out) pass")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        # stdout carries exactly one JSON line: this image exports NCCL_DEBUG=VERSION and NCCL prints its
        # version banner on stdout (NCCL_DEBUG_FILE does not move it) when the first communicator is made,
        # so file descriptor 1 points at stderr while the process group comes up
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    from make_prg_b200 import device

    ctx = device.Context(local_rank)
    # builds in flight side by side in the two throughput measurements (the product's own pipeline object)
    pipe = device.BuildPipeline(local_rank)
    LANES = pipe.depth
    n_loci = args.loci
    data = workload(rank, n_loci)
    shapes = [(ROWS, COLS)] * n_loci
    host = torch.from_numpy(data.reshape(-1)).pin_memory()  # pinned host ASCII
    host_np = host.numpy()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    seen_lengths = {}  # PRG lengths of the last step of each path, compared with the checked run below

    def step_resident(batch, keep=None):
        res = ctx.build(batch, MAX_NESTING, MIN_MATCH)
        status, lengths = res.statuses()
        n_ok = int((status == 0).sum())
        total_len = int(lengths.sum()) + sum(len(res.prg(i)) for i in range(0, n_loci, 97))
        seen_lengths["resident"] = lengths
        if keep is not None:
            keep.extend(res.prg(i) for i in range(min(CPU_BASELINE_SAMPLE, n_loci)))
        res.free()
        return n_ok, total_len

    # what the native loader hands to the engine: the matrices in the 4-bit device layout, in pinned host
    # memory (hostio.load_fasta_files(packed=True) / mprg_pack_rows), made once outside the timed region
    # like the FASTA parse itself; the `files` object below times loader + engine + writers together
    from make_prg_b200 import hostio

    stride = hostio.packed_stride(COLS)
    packed_host = torch.empty(n_loci * ROWS * stride, dtype=torch.uint8).pin_memory()
    packed_np = packed_host.numpy()
    pk_flags = np.zeros(n_loci, np.int32)
    for i in range(n_loci):
        rows, pk_flags[i] = hostio.pack_rows(data[i])
        packed_np[i * ROWS * stride:(i + 1) * ROWS * stride] = rows.reshape(-1)
    pk_offsets = np.arange(n_loci, dtype=np.int64) * (ROWS * stride)
    pk_rows = np.full(n_loci, ROWS, np.int32)
    pk_cols = np.full(n_loci, COLS, np.int32)

    def step_e2e(text=False):
        # the public one-call paths: pinned host rows in (packed: mprg_build_packed; text: mprg_build_ascii),
        # PRG strings out
        if text:
            batch, res = ctx.build_ascii((host_np, shapes), MAX_NESTING, MIN_MATCH)
        else:
            batch, res = ctx.build_packed(packed_np, pk_offsets, pk_rows, pk_cols, pk_flags, MAX_NESTING, MIN_MATCH)
        status, lengths = res.statuses()
        n_ok = int((status == 0).sum())
        total_len = int(lengths.sum()) + sum(len(res.prg(i)) for i in range(0, n_loci, 97))
        seen_lengths["e2e"] = lengths
        res.free()
        batch.free()
        return n_ok, total_len

    # ---- kernel-side number: batch resident in HBM ----
    batch = ctx.upload((host_np, shapes))
    for _ in range(args.warmup):
        step_resident(batch)
        flush.zero_()
    barrier()
    ctx.scan_log(reset=True)
    ctx.kmeans_stats(reset=True)
    launches0 = ctx.launch_count()
    with ClockSampler(local_rank) as clocks:
        t_dev = 0.0
        dev_steps = []
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.timer_start()
            n_ok, _ = step_resident(batch)
            dev_steps.append(ctx.timer_stop())
            t_dev += dev_steps[-1]
            flush.zero_()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    barrier()
    launches = ctx.launch_count() - launches0
    log_bytes, log_ms = ctx.scan_log(reset=True)
    km = ctx.kmeans_stats(reset=True)

    # ---- roofline pass: the dominant kernel (root-level column scan of the whole batch) alone on one
    # stream, CUDA events around the kernel (mprg_scan_log), L2 flushed between launches ----
    root_tasks = [(i, None, 0, COLS) for i in range(n_loci)]
    for _ in range(3):
        ctx.partition_tasks(batch, root_tasks, MIN_MATCH)
    ctx.scan_log(reset=True)
    for _ in range(max(args.steps, 5)):
        flush.zero_()
        torch.cuda.synchronize()
        ctx.partition_tasks(batch, root_tasks, MIN_MATCH)
    iso_bytes, iso_ms = ctx.scan_log(reset=True)
    # yardstick: a bare streaming read of the same packed batch (same tile shape, no other work), same
    # timing method -- what the HBM system delivers for a launch of this size
    yard = []
    for _ in range(max(args.steps, 5) + 2):
        flush.zero_()
        torch.cuda.synchronize()
        yard.append(ctx.read_yardstick(batch))
    yard = yard[2:]
    # parity of the timed batch: one more (untimed) build whose PRG strings are kept for the oracle check
    # below; the timed steps must have produced PRGs of exactly these lengths
    timed_lengths = seen_lengths["resident"].copy()
    gpu_prgs = []
    step_resident(batch, keep=gpu_prgs)
    if not np.array_equal(timed_lengths, seen_lengths["resident"]):
        raise SystemExit("bench: PRG lengths differ between two builds of the same batch")
    # ---- `value`: throughput with the batches resident: LANES copies of the batch in HBM (LANES x 100 MB of
    # inputs, larger than L2), lane i builds copy i, K builds issued back to back, LANES of them in flight: the
    # two host synchronisations per recursion level of one build are filled by the kernels of the others ----
    lane_batches = [batch] + [ctx.upload((host_np, shapes)) for _ in range(LANES - 1)]

    def consume_resident(b, res):
        status, lengths = res.statuses()
        total_len = int(lengths.sum()) + sum(len(res.prg(i)) for i in range(0, n_loci, 97))
        return int((status == 0).sum()), lengths.copy(), total_len

    def run_lanes(steps, submit):
        import collections

        futs, done = collections.deque(), []
        for _ in range(steps):
            futs.append(submit())
            flush.zero_()  # one 256 MiB write per step on torch's stream, beside the builds
            while len(futs) > LANES:
                done.append(futs.popleft().result())
        while futs:
            done.append(futs.popleft().result())
        return done

    def submit_resident():
        return pipe.submit_resident(lane_batches[pipe.next_lane], MAX_NESTING, MIN_MATCH, consume=consume_resident)

    run_lanes(max(args.warmup, 2 * LANES), submit_resident)
    barrier()
    lane_launches0 = pipe.launch_count()
    with ClockSampler(local_rank) as lane_clocks:
        t0 = time.perf_counter()
        lane_results = run_lanes(args.steps, submit_resident)
        torch.cuda.synchronize()
        t_lanes = 1e3 * (time.perf_counter() - t0)
    barrier()
    lane_launches = pipe.launch_count() - lane_launches0
    for ok_k, lengths_k, _ in lane_results:
        if ok_k != n_ok or not np.array_equal(lengths_k, timed_lengths):
            raise SystemExit("bench: a build in the lanes differs from the one-at-a-time build")
    for b in lane_batches[1:]:
        b.free()
    batch.free()
    # the same kernel on a launch 8x the size (the batch repeated), to separate launch-size effects from
    # kernel quality: 8,000 root tasks, 840 MB per launch
    big = None
    if rank == 0 and not args.no_big:
        reps8 = 8
        host8 = np.concatenate([host_np] * reps8)
        batch8 = ctx.upload((host8, shapes * reps8))
        tasks8 = [(i, None, 0, COLS) for i in range(n_loci * reps8)]
        for _ in range(2):
            ctx.partition_tasks(batch8, tasks8, MIN_MATCH)
        ctx.scan_log(reset=True)
        for _ in range(5):
            flush.zero_()
            torch.cuda.synchronize()
            ctx.partition_tasks(batch8, tasks8, MIN_MATCH)
        b8, m8 = ctx.scan_log(reset=True)
        y8 = []
        for _ in range(5):
            flush.zero_()
            torch.cuda.synchronize()
            y8.append(ctx.read_yardstick(batch8))
        batch8.free()
        del host8
        big = {"loci": n_loci * reps8, "bytes_per_launch": float(b8.mean()), "ms_per_launch": float(m8.mean()),
               "achieved": float(b8.sum() / (m8.sum() * 1e-3) / 1e9),
               "bare_read_gbs": float(sum(b for b, _ in y8) / (sum(m for _, m in y8) * 1e-3) / 1e9)}

    # ---- end to end: host buffers in, PRG strings out ----
    # (1) one call at a time: the latency of one step (upload, level loop, strings back), device time per step
    for _ in range(2):
        step_e2e()
    barrier()
    ctx.copy_stats(reset=True)
    t_serial = 0.0
    e2e_steps = []
    for _ in range(args.steps):
        ctx.timer_start()
        step_e2e()
        e2e_steps.append(ctx.timer_stop())
        t_serial += e2e_steps[-1]
        flush.zero_()
    barrier()
    serial_copies = ctx.copy_stats(reset=True)
    # (2) the headline: the same K steps, every one the same C-ABI call on the same pinned host rows with its own
    # H2D copy and its own PRG strings back, but LANES of them in flight (device.BuildPipeline, what the
    # from_msa pipeline runs): the upload of step k+1 crosses PCIe while step k is in its level loop.  Timed on
    # the host clock between barrier + synchronize on both sides; a step's result is read on the host (statuses,
    # lengths, sampled PRG strings) before its buffers are released.
    def consume(b, res):
        status, lengths = res.statuses()
        total_len = int(lengths.sum()) + sum(len(res.prg(i)) for i in range(0, n_loci, 97))
        return int((status == 0).sum()), lengths.copy(), total_len

    def submit_e2e():
        return pipe.submit_packed(packed_np, pk_offsets, pk_rows, pk_cols, pk_flags, MAX_NESTING, MIN_MATCH,
                                  consume=consume)

    run_lanes(max(args.warmup, 2 * LANES), submit_e2e)
    barrier()
    pipe.copy_stats(reset=True)
    t0 = time.perf_counter()
    pipelined = run_lanes(args.steps, submit_e2e)
    torch.cuda.synchronize()
    t_e2e = 1e3 * (time.perf_counter() - t0)
    barrier()
    copies = pipe.copy_stats(reset=True)
    for ok_k, lengths_k, _ in pipelined:
        if ok_k != n_ok or not np.array_equal(lengths_k, seen_lengths["e2e"]):
            raise SystemExit("bench: a pipelined end-to-end step differs from the one-at-a-time build")
    pipe.close()
    # the same from pinned host TEXT (ASCII rows copied and packed on the device), for callers without the loader
    for _ in range(2):
        step_e2e(text=True)
    barrier()
    text_steps = []
    for _ in range(max(5, args.steps // 2)):
        ctx.timer_start()
        step_e2e(text=True)
        text_steps.append(ctx.timer_stop())
        flush.zero_()
    barrier()
    text_copies = ctx.copy_stats(reset=True)

    t_dev_s, t_e2e_s, t_serial_s = t_lanes / 1e3, t_e2e / 1e3, t_dev / 1e3
    if dist is not None:
        tt = torch.tensor([t_dev_s, t_e2e_s, t_serial_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev_s, t_e2e_s, t_serial_s = float(tt[0]), float(tt[1]), float(tt[2])
        ok = torch.tensor([n_ok], dtype=torch.int64, device="cuda")
        dist.all_reduce(ok)
        n_ok_total = int(ok[0])
    else:
        n_ok_total = n_ok
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    total_loci = n_loci * world
    value = total_loci * args.steps / t_dev_s
    e2e_value = total_loci * args.steps / t_e2e_s
    # roofline of the dominant launch: the root-level scan (largest algorithmic byte count per step)
    peak, peak_src = hbm_peak()
    roof = None
    if len(iso_bytes):
        achieved = float(iso_bytes.sum() / (iso_ms.sum() * 1e-3) / 1e9)
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_scan_traffic(),
                "traffic_source": "profiles/r2_ncu_scan_kernel.txt (ncu --set full of the same launch: "
                                  "dram__bytes_read.sum + dram__bytes_write.sum)",
                "kernel": "scan_kernel<false>, root-level launch over the whole batch, timed alone",
                "bytes_per_launch": float(iso_bytes.mean()), "ms_per_launch": float(iso_ms.mean()),
                "launches_timed": int(len(iso_bytes)), "peak_source": peak_src,
                "bare_read_same_launch": {
                    "note": "one bare streaming read of the same packed batch (mprg_read_yardstick), same "
                            "timing method: the ceiling for a launch of this size",
                    "gbs": float(sum(b for b, _ in yard) / (sum(m for _, m in yard) * 1e-3) / 1e9),
                    "frac_of_peak": float(sum(b for b, _ in yard) / (sum(m for _, m in yard) * 1e-3) / 1e9) / peak},
                "launch_8x": None if big is None else dict(big, frac=big["achieved"] / peak,
                                                           bare_read_frac=big["bare_read_gbs"] / peak),
                "in_step": {"note": "same kernel inside the timed steps: one launch per recursion level over the "
                                    "whole batch (device-resident level loop); root_level_launch_gbs = the "
                                    "105 MB launch of the root level, as timed inside the steps",
                            "launches_per_step": len(log_bytes) / args.steps,
                            "root_level_launch_gbs": float(
                                log_bytes[log_bytes >= 0.99 * log_bytes.max()].sum() /
                                (log_ms[log_bytes >= 0.99 * log_bytes.max()].sum() * 1e-3) / 1e9) if len(log_bytes) else None,
                            "achieved_gbs_sum_over_launches": float(log_bytes.sum() / (log_ms.sum() * 1e-3) / 1e9)
                            if len(log_bytes) else None}}
    # file to file: FASTA files in -> .prg.fa / .prg.bin.zip / .prg.gfa.zip out through the
    # native loader and writers, one batch (1,000 loci) and a pipelined run (4 x the loci, chunked)
    files = None
    if not args.no_files and n_loci == LOCI_PER_GPU:
        try:
            sys.path.insert(0, str(REPO / "scripts"))
            import files_e2e

            files = {"one_batch": files_e2e.measure(n_loci, reps=5),
                     "one_batch_with_update_ds": files_e2e.measure(n_loci, reps=3, update_ds=True),
                     "pipelined_x4": files_e2e.measure(n_loci, reps=3, copies=4)}
        except Exception as err:  # the headline numbers above do not depend on this pass
            files = {"error": repr(err)}
    # CPU baseline: oracle port, 1 core, bounded sample
    sample = min(CPU_BASELINE_SAMPLE, n_loci)
    oracle_prgs = []
    cpu_v, cpu_dt = oracle_loci_per_sec([data[i] for i in range(sample)], 1, keep=oracle_prgs)
    # the oracle is the checker here: the PRGs of the timed batch must be the reference algorithm's
    bad = [i for i in range(sample) if gpu_prgs[i] != oracle_prgs[i]]
    if bad or not np.array_equal(timed_lengths, seen_lengths["e2e"]):
        raise SystemExit(f"bench: PRG parity failure (loci {bad[:10]}, e2e lengths equal: "
                         f"{np.array_equal(timed_lengths, seen_lengths['e2e'])})")
    line = {
        "metric": "MSA loci/sec (from_msa, byte-identical PRG)", "value": value, "unit": "loci/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_dev_s / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config_dict(),
        "columns_per_sec": value * COLS, "loci_ok": n_ok_total,
        "mode": f"throughput: {args.steps} builds (steps) of resident batches issued back to back, {LANES} in flight "
                f"(device.BuildPipeline: {LANES} lanes = contexts with their own stream and host thread, lane i builds "
                f"its own copy of the batch, {LANES} x 100 MB of inputs > L2), host clock between barrier + "
                f"synchronize, max over ranks",
        "one_at_a_time": {"note": "one build at a time on one stream (latency of a step), device time per step, L2 "
                                  "flushed between steps, max over ranks; roofline.in_step, kmeans and "
                                  "step_ms.resident come from these steps",
                          "ms_per_step": 1e3 * t_serial_s / args.steps,
                          "value": total_loci * args.steps / t_serial_s,
                          "gpu_launches": int(launches)},
        "parity": {"prgs_equal_oracle": sample, "of": sample,
                   "note": "PRG strings of the first loci of the timed batch == oracle port; PRG lengths of "
                           "every locus equal across resident / e2e / checked builds"},
        "e2e": {"value": e2e_value, "unit": "loci/s",
                "h2d_bytes_per_step": copies["h2d_bytes"] // args.steps,
                "d2h_bytes_per_step": copies["d2h_bytes"] // args.steps,
                "ms_per_step": 1e3 * t_e2e_s / args.steps,
                "input": "pinned host rows in the 4-bit layout the native loader emits (mprg_build_packed)",
                "mode": f"{LANES} steps in flight (device.BuildPipeline: one mprg_build_packed call per step, "
                        f"each with its own H2D copy and its own results on the host; the upload of step k+1 "
                        f"overlaps the level loop of step k); {args.steps} steps on the host clock between "
                        f"barrier + synchronize, max over ranks; L2: {LANES} input arenas of 102 MB take turns in a "
                        f"126 MB L2, plus one 256 MiB flush write per step on a third stream",
                "one_at_a_time": {"note": "the same call, one step at a time (latency of a step), device time, "
                                          "L2 flushed between steps, rank 0",
                                  "ms_per_step": t_serial / args.steps,
                                  "value": n_loci / (t_serial / args.steps * 1e-3),
                                  "h2d_bytes_per_step": serial_copies["h2d_bytes"] // args.steps,
                                  "d2h_bytes_per_step": serial_copies["d2h_bytes"] // args.steps},
                "from_text": {"note": "same call from pinned host ASCII rows (mprg_build_ascii: copy + device pack), rank 0",
                              "ms_per_step": float(np.mean(text_steps)),
                              "value": n_loci / (float(np.mean(text_steps)) * 1e-3),
                              "h2d_bytes_per_step": text_copies["h2d_bytes"] // len(text_steps)}},
        "gpu_launches": int(lane_launches),
        "roofline": roof,
        # the time-dominant kernel of a step is not bandwidth-bound: float64 KMeans of thousands of tiny problems
        # (sequential-order sums), reported as problems/s with its pipe counters
        "kmeans": {"kernel": "kmeans_kernel_w32 / kmeans_kernel: one warp (CTA) per initialisation, 10 per problem",
                   "ms_per_step": km["ms"] / args.steps, "launches_per_step": km["launches"] / args.steps,
                   "problems_per_step": km["problems"] / args.steps,
                   "problem_fits_per_s": km["problems"] / (km["ms"] * 1e-3) if km["ms"] > 0 else None,
                   "timing": "CUDA events around every KMeans launch inside the timed steps (rank 0)",
                   "ncu": dict(ncu_counters("kmeans_kernel_w32", (
                       "gpu__time_duration.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
                       "smsp__issue_active.avg.pct_of_peak_sustained_active",
                       "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread")),
                       source="profiles/r2_ncu_kmeans_kernel_w32.txt (largest launch: 2,489 problems x 10)")},
        "cpu_baseline": {"value": cpu_v, "unit": "loci/s", "cores": 1, "kind": "port",
                         "sample": f"first {sample} loci of the workload, oracle/make_prg_oracle.py, "
                                   f"{cpu_dt:.1f} s"},
        "files": files,
        "clocks": lane_clocks.summary(),
        "clocks_one_at_a_time": clocks.summary(),
        "wall_ms_per_step": 1e3 * wall / args.steps,
        "step_ms": {"resident": {"min": float(np.min(dev_steps)), "median": float(np.median(dev_steps)),
                                 "max": float(np.max(dev_steps))},
                    "e2e_one_at_a_time": {"min": float(np.min(e2e_steps)), "median": float(np.median(e2e_steps)),
                                          "max": float(np.max(e2e_steps))}, "note": "rank 0, device time per step"},
        "host": {"cores": os.cpu_count(), "loadavg": list(os.getloadavg())},
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
