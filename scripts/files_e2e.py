"""File-to-file throughput of `from_msa` (FASTA in -> .prg.fa/.bin/.gfa out) on config #2.

    python scripts/files_e2e.py [n_loci] [reps] [copies] [--python-io]

Writes the workload as FASTA files (60-column lines) under /dev/shm, then times
make_prg_b200.subcommands.from_msa.build_and_write (native loader -> mprg_build_ascii -> native writers)
and its stages.  --python-io also times the Python loader / writers of the same package for comparison."""
import json
import os
import shutil
import sys
import time
from argparse import Namespace
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

import numpy as np  # noqa: E402


def write_fastas(directory, n_loci, config=2):
    from make_prg_b200 import synth

    directory.mkdir(parents=True, exist_ok=True)
    paths = []
    for i in range(n_loci):
        p = directory / f"locus{i:06d}.fa"
        paths.append(p)
        if p.exists():
            continue
        M = synth.config_msa(config, i)
        with open(p, "wb") as fh:
            for r, row in enumerate(M):
                s = row.tobytes()
                fh.write(b">seq%d sample %d\n" % (r, r))
                fh.write(b"".join(s[k:k + 60] + b"\n" for k in range(0, len(s), 60)))
    return paths


def replicate(paths, copies, directory):
    """copies x the same loci under distinct names (symlinks): a bigger run without generating more data."""
    directory.mkdir(parents=True, exist_ok=True)
    out = []
    for k in range(copies):
        for p in paths:
            q = directory / f"{p.stem}_{k}.fa"
            if not q.exists():
                os.symlink(p, q)
            out.append(q)
    return out


def measure(n_loci=1000, reps=5, copies=1, python_io=False, update_ds=False):
    root = Path(os.environ.get("MPRG_FILES_DIR", "/dev/shm/mprg_files"))
    paths = write_fastas(root / f"config2_{n_loci}", n_loci)
    if copies > 1:
        paths = replicate(paths, copies, root / f"config2_{n_loci}_x{copies}")
        n_loci *= copies
    in_bytes = sum(os.path.getsize(p) for p in paths)

    from make_prg_b200 import device, hostio
    from make_prg_b200.subcommands import from_msa
    from make_prg_b200.subcommands.output_type import OutputType

    out = root / "out"
    shutil.rmtree(out, ignore_errors=True)
    out.mkdir(parents=True)
    opts = Namespace(input=str(root), suffix="", output_prefix=str(out / "run"), alignment_format="fasta",
                     max_nesting=5, min_match_length=7, output_type=OutputType("a"), force=True, threads=1,
                     gpus=1, skip_update_ds=not update_ds)
    from loguru import logger

    logger.remove()
    ctx = device.default_context(0)
    times = []
    for r in range(reps + 2):
        t0 = time.perf_counter()
        n_ok = from_msa.build_and_write(paths, opts)
        times.append(time.perf_counter() - t0)
    times = times[2:]
    out_bytes = sum(os.path.getsize(p) for p in out.glob("run*"))
    line = {"what": "from_msa file to file, config #2", "n_loci": n_loci, "n_ok": n_ok, "reps": reps,
            "input_bytes": in_bytes, "output_bytes": out_bytes,
            "wall_s": {"min": min(times), "median": float(np.median(times)), "max": max(times)},
            "loci_per_s": n_loci / float(np.median(times)),
            "columns_per_s": n_loci * 1000 / float(np.median(times)),
            "chunks": len(from_msa.cut_chunks(paths)), "host_cores": os.cpu_count(),
            "timing": "host wall clock around build_and_write (files in page cache, outputs to tmpfs), "
                      "warm pinned-buffer pool",
            "outputs": "prg.fa + prg.bin.zip + prg.gfa.zip" +
                       (" + update_DS.zip (table-shaped records, the default -O a run)" if update_ds
                        else " (--skip-update-ds)")}
    if copies > 1:
        return line
    # stages, one after the other
    t0 = time.perf_counter()
    msas = hostio.load_fasta_files(paths, packed=True)
    t_load = time.perf_counter() - t0
    # the context build_and_write's one-chunk run used (a lane of the process-wide pipeline): warm buffers
    ctx = device.default_pipeline(0, 1).contexts[0]
    t0 = time.perf_counter()
    batch, res = ctx.build_msa_set(msas, 5, 7)
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    w = hostio.OutputWriter(out / "stage")
    w.add(res, np.arange(n_loci, dtype=np.int32), [p.stem for p in paths])
    w.close()
    t_write = time.perf_counter() - t0
    line["stages_s"] = {"load": t_load, "build_packed": t_build, "write": t_write}
    if python_io:
        from make_prg_b200.utils.io_utils import load_alignment_file

        k = min(n_loci, 50)
        t0 = time.perf_counter()
        for p in paths[:k]:
            load_alignment_file(str(p)).matrix
        line["python_loader_s_per_locus"] = (time.perf_counter() - t0) / k
        t0 = time.perf_counter()
        for i in range(k):
            from_msa._gfa_text(res.prg(i))
            from_msa._bin_bytes(res.prg(i))
        line["python_writers_s_per_locus"] = (time.perf_counter() - t0) / k
        t0 = time.perf_counter()
        from_msa._update_ds_pickles([f"l{i}" for i in range(k)], msas, res, list(range(k)), opts)
        line["update_ds_pickle_s_per_locus"] = (time.perf_counter() - t0) / k
    res.free()
    batch.free()
    msas.free()
    return line


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    print(json.dumps(measure(int(args[0]) if args else 1000, int(args[1]) if len(args) > 1 else 5,
                             int(args[2]) if len(args) > 2 else 1, "--python-io" in sys.argv)))
