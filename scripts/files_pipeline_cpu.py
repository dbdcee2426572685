"""CPU only: the host side of the file pipeline (native loader -> [build] -> native writers) with the device stage
replaced by a stub that sleeps BUILD_MS and hands back oracle PRG strings, to see how the loader and writer threads
share the host cores when a run has several chunks (MPRG_LOAD_THREADS / MPRG_WRITE_THREADS override side_threads).

    python scripts/files_pipeline_cpu.py [copies=4] [build_ms=3]
"""
import os
import shutil
import sys
import time
from argparse import Namespace
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "oracle"))
sys.path.insert(0, str(REPO / "scripts"))

import files_e2e  # noqa: E402
from make_prg_b200 import device, hostio, synth  # noqa: E402
from make_prg_b200.subcommands import from_msa  # noqa: E402
from make_prg_b200.subcommands.output_type import OutputType  # noqa: E402


def main():
    copies = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    build_ms = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
    import make_prg_oracle as mo
    from loguru import logger

    logger.remove()
    root = Path(os.environ.get("MPRG_FILES_DIR", "/dev/shm/mprg_files"))
    paths = files_e2e.write_fastas(root / "config2_1000", 1000)
    if copies > 1:
        paths = files_e2e.replicate(paths, copies, root / f"config2_1000_x{copies}")
    prgs = []
    for i in range(40):
        M = synth.config_msa(2, i)
        prgs.append(mo.build_prg_from_matrix([f"s{r}" for r in range(M.shape[0])], M, 5, 7)[0])

    class Res(hostio.PrgStrings):
        def __init__(self, n):
            super().__init__([prgs[i % len(prgs)] for i in range(n)])
            self.n = n

        def statuses(self):
            return np.zeros(self.n, np.int32), np.zeros(self.n, np.int64)

    class Batch:
        def free(self):
            pass

    class StubPipe:
        pool = ThreadPoolExecutor(3)

        def submit_msa_set(self, msas, N, L, consume=None):
            n = msas.n_loci

            def job():
                time.sleep(max(2e-3, build_ms * 1e-3 * n / 1000))  # (a build is at least a level loop's latency)
                return Batch(), Res(n)

            return self.pool.submit(job)

    device.default_pipeline = lambda dev=0, depth=None: StubPipe()
    out = root / "out_cpu"
    shutil.rmtree(out, ignore_errors=True)
    out.mkdir(parents=True)
    opts = Namespace(input=str(root), suffix="", output_prefix=str(out / "run"), alignment_format="fasta", max_nesting=5,
                     min_match_length=7, output_type=OutputType("a"), force=True, threads=1, gpus=1, skip_update_ds=True)
    times = []
    for r in range(6):
        t0 = time.perf_counter()
        n_ok = from_msa.build_and_write(paths, opts)
        times.append(time.perf_counter() - t0)
    chunks = len(from_msa.cut_chunks(paths))
    print(f"loci {n_ok} chunks {chunks} load_threads {from_msa.side_threads(chunks)} write_threads "
          f"{from_msa.side_threads(chunks, writer=True)} build_ms/1000 {build_ms}: wall ms "
          f"{[round(1e3 * t, 1) for t in times[1:]]} median {1e3 * float(np.median(times[1:])):.1f}")


if __name__ == "__main__":
    main()
