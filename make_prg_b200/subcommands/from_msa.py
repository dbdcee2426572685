"""`make_prg from_msa` with the reference's flags, errors and output files
(make_prg/subcommands/from_msa.py, make_prg/utils/input_output_files.py), but with the per-locus
process pool replaced by level-synchronous GPU batching: every locus of the run (or of this GPU's
shard) goes through the same kernel launches."""
import io
import os
import re
import shutil
import tempfile
import zipfile
from pathlib import Path

import numpy as np
from loguru import logger

from .. import engine, hostio
from .._lib import LOCUS_CURATION_ERROR, LOCUS_OK
from ..from_msa import MIN_MATCH_LEN, NESTING_LVL
from ..prg_builder import PrgBuilder
from ..utils.gfa import GFA_Output, HEADER
from ..utils.io_utils import load_alignment_file
from ..utils.prg_encoder import PrgEncoder


class EmptyMSAError(Exception):
    pass


def register_parser(subparsers):
    p = subparsers.add_parser("from_msa", usage="make_prg from_msa",
                              help="Make PRG from multiple sequence alignment")
    p.add_argument("-i", "--input", action="store", type=str, required=True,
                   help="Multiple sequence alignment file or a directory containing such files")
    p.add_argument("-s", "--suffix", action="store", type=str, default="",
                   help="If the input parameter (-i, --input) is a directory, then filter for files "
                        "with this suffix. If this parameter is not given, all files in the input "
                        "directory is considered.")
    p.add_argument("-o", "--output-prefix", dest="output_prefix", action="store", type=str,
                   required=True, help="Prefix for the output files")
    p.add_argument("-f", "--alignment-format", dest="alignment_format", action="store", default="fasta",
                   help="Alignment format of MSA. Default: %(default)s")
    p.add_argument("-N", "--max-nesting", dest="max_nesting", action="store", type=int,
                   default=NESTING_LVL,
                   help="Maximum number of levels to use for nesting. Default: %(default)d")
    p.add_argument("-L", "--min-match-length", dest="min_match_length", action="store", type=int,
                   default=MIN_MATCH_LEN,
                   help="Minimum number of consecutive characters which must be identical for a match. "
                        "Default: %(default)d")
    p.add_argument("--gpus", dest="gpus", action="store", type=int, default=1,
                   help="Number of GPUs of this node to shard the loci over. Default: %(default)d")
    p.add_argument("--skip-update-ds", dest="skip_update_ds", action="store_true",
                   help="Do not write <prefix>.update_DS.zip (the pickled PrgBuilders `make_prg update` "
                        "reads); everything else of -O p is still written")
    p.set_defaults(func=run)
    return p


def get_all_input_files(input_path, suffix):
    input_path = Path(input_path)
    if not input_path.exists():
        raise FileNotFoundError(f"{input_path} does not exist")
    if input_path.is_file():
        return [input_path]
    return [p.resolve() for p in input_path.iterdir() if p.is_file() and p.name.endswith(suffix)]


def remove_known_input_extensions(name):
    return re.sub(r"\.(fa|fasta)(\.gz)?$", "", name)


def output_files_already_exist(output_type, output_prefix):
    names = []
    if output_type.prg:
        names += [".prg.fa", ".update_DS.zip"]
    if output_type.gfa:
        names += [".prg.gfa", ".prg.gfa.zip"]
    if output_type.binary:
        names += [".prg.bin", ".prg.bin.zip"]
    return any(Path(output_prefix + n).exists() for n in names)


def _gfa_text(prg):
    g = GFA_Output(HEADER)
    g.build_gfa_string(prg_string=prg)
    return g.gfa_string


def _bin_bytes(prg):
    buf = io.BytesIO()
    enc = PrgEncoder()
    enc.write(enc.encode(prg), buf)
    return buf.getvalue()


def cut_chunks(input_files, max_bytes=None, max_loci=16384):
    """Consecutive runs of input files of about max_bytes each: one chunk = one loader call + one device
    batch, so that loading chunk k+1, building chunk k and writing chunk k-1 overlap."""
    if max_bytes is None:
        max_bytes = int(float(os.environ.get("MPRG_CHUNK_MB", "256")) * (1 << 20))
    chunks, cur, size = [], [], 0
    for path in input_files:
        try:
            nbytes = os.stat(path).st_size
        except OSError:
            nbytes = 0
        if cur and (size + nbytes > max_bytes or len(cur) >= max_loci):
            chunks.append(cur)
            cur, size = [], 0
        cur.append(path)
        size += nbytes
    if cur:
        chunks.append(cur)
    return chunks


def side_threads(n_chunks, writer=False):
    """Host threads of the loader and of the writers.  With one chunk nothing overlaps and each stage may
    use every core; with several, loading chunk k+1 and writing chunk k-1 run beside the build of chunk k.
    The stages do not take the same time and the build threads mostly wait, so a fixed split of the cores
    leaves some idle while the slowest stage (the writers, at a quarter of the cores) sets the pace: both stages
    get nearly all of this process's cores and the scheduler shares them (4,000 loci in 4 chunks on 8 cores
    with the device stage stubbed, scripts/files_pipeline_cpu.py: loader 4 + writers 2 threads 301-323 ms,
    7 + 6 threads 224 ms).  MPRG_LOAD_THREADS / MPRG_WRITE_THREADS override."""
    override = os.environ.get("MPRG_WRITE_THREADS" if writer else "MPRG_LOAD_THREADS")
    if override:
        return max(1, int(override))
    cores = os.cpu_count() or 1
    share = max(1, cores // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)))
    if n_chunks <= 1:
        return max(1, min(32, share))
    return max(1, min(16, share * 3 // 4)) if writer else max(1, min(24, share - 1))


def _load_chunk(paths, alignment_format, threads=None):
    if alignment_format.lower() != "fasta":
        # other formats (utils/alignment_formats.py) are rewritten as FASTA text for the native loader, so that
        # upper-casing, N replacement and packing are the same code as for FASTA input
        import tempfile

        from ..utils.alignment_formats import read_records, to_fasta_text

        tmp = tempfile.TemporaryDirectory(prefix="mprg_fmt_")
        converted = []
        for i, path in enumerate(paths):
            out = os.path.join(tmp.name, f"{i}.fa")
            with open(out, "w") as fh:
                fh.write(to_fasta_text(read_records(path, alignment_format)))
            converted.append(out)
        msas = hostio.load_fasta_files(converted, threads=threads, packed=not os.environ.get("MPRG_TEXT_UPLOAD"))
        msas.converted_from = tmp  # the FASTA copies live as long as the set (its error paths re-read a file)
        return msas
    # the loader emits the matrices in the 4-bit device layout: half the bytes cross PCIe, no pack kernel
    # (MPRG_TEXT_UPLOAD=1 keeps the text path: upload of ASCII rows + pack_rows_kernel)
    return hostio.load_fasta_files(paths, threads=threads, packed=not os.environ.get("MPRG_TEXT_UPLOAD"))


_EXTENSIONS = (".fasta.gz", ".fa.gz", ".fasta", ".fa")


def locus_name_of(path):
    """remove_known_input_extensions(Path(path).name) without the Path object and the regular expression."""
    name = os.path.basename(os.fspath(path))
    for ext in _EXTENSIONS:
        if name.endswith(ext):
            return name[:-len(ext)]
    return name


def iter_built_chunks(input_files, options, device_ordinal=0, chunks=None):
    """Loads (native loader, one chunk ahead on a host thread) and builds the input files chunk by chunk.  The
    builds go through device.BuildPipeline: chunk k+1 is submitted before the result of chunk k is handed on, so
    its host-to-device copy overlaps the level loop of chunk k (one lane for a one-chunk run).  Yields (names,
    msas, result, ok) in input order, ok = indices of the chunk's loci that were built; the consumer frees result
    and msas.  Errors follow from_msa.py:142-151: an empty MSA aborts the run, a locus with a disallowed base
    is skipped with a warning."""
    from collections import deque
    from concurrent.futures import ThreadPoolExecutor

    from .. import device

    if chunks is None:
        chunks = cut_chunks(input_files)
    pipe = device.default_pipeline(device_ordinal, min(3, device.default_lanes(), max(1, len(chunks))))
    in_flight = deque()  # (names, msas, curation, future of (batch, result)), oldest first

    def built(entry):
        names, msas, curation, fut = entry
        try:
            return checked(names, msas, curation, fut)
        except BaseException:
            msas.free()  # this chunk is not handed on: nobody else would release its matrices
            raise

    def checked(names, msas, curation, fut):
        batch, res = fut.result()
        batch.free()
        statuses, _lengths = res.statuses()
        ok = np.nonzero(statuses == LOCUS_OK)[0].tolist()
        for i in np.nonzero(statuses != LOCUS_OK)[0].tolist():
            if statuses[i] == LOCUS_CURATION_ERROR or i in curation:
                logger.warning(f"Skipping building PRG for {names[i]}. Error: a slice of a sequence has a "
                               "disallowed base. Redo sequence curation.")
            else:
                try:
                    engine.LocusBuild(int(statuses[i]), "", 0, 0, None).raise_for_status(names[i])
                except BaseException:
                    res.free()
                    raise
        return names, msas, res, ok

    future = None  # the chunk being loaded ahead
    loaded = None  # a chunk that has been loaded and not yet submitted
    try:
        with ThreadPoolExecutor(1) as pool:
            threads = side_threads(len(chunks))
            future = pool.submit(_load_chunk, chunks[0], options.alignment_format, threads) if chunks else None
            for k, paths in enumerate(chunks):
                msas = loaded = future.result()
                future = (pool.submit(_load_chunk, chunks[k + 1], options.alignment_format, threads)
                          if k + 1 < len(chunks) else None)
                names = [locus_name_of(path) for path in paths]
                logger.info(f"Generating PRGs for {names[0]} ... {names[-1]} ({len(names)} loci)...")
                for i in np.nonzero(msas.flags & hostio.FLAG_DUPLICATE_IDS)[0]:
                    logger.warning(f"{names[int(i)]}: duplicated record ids; clusters are cut by row here, the "
                                   "reference pulls every record with a clustered id (recursion_tree.py:558-572), so "
                                   "its PRG may differ for this locus")
                curation = set()  # loci whose rows hold non-ASCII characters: skipped like any disallowed base
                for i in np.nonzero(msas.status != hostio.LOAD_OK)[0]:
                    try:
                        hostio.raise_for_load_status(msas, int(i))
                    except hostio.NonAsciiSequenceError:
                        curation.add(int(i))
                    except ValueError as err:
                        if "No records found in handle" in str(err.args[0]):
                            raise EmptyMSAError(f"No records found in MSA of locus {names[int(i)]}")
                        raise
                in_flight.append((names, msas, curation,
                                  pipe.submit_msa_set(msas, options.max_nesting, options.min_match_length)))
                loaded = None
                while len(in_flight) > 1:
                    yield built(in_flight.popleft())
            while in_flight:
                yield built(in_flight.popleft())
    finally:
        # an error (here or in the consumer) with builds still in flight: wait for them and drop their results
        while in_flight:
            _names, msas, _curation, fut = in_flight.popleft()
            try:
                batch, res = fut.result()
                batch.free()
                res.free()
            except Exception:
                pass
            msas.free()
        if loaded is not None:  # (an empty MSA in it ended the run)
            loaded.free()
        if future is not None:  # a chunk that was loaded ahead and never built
            try:
                future.result().free()
            except Exception:
                pass


def _update_ds_pickles(names, msas, res, ok, options):
    """(name, pickled PrgBuilder): the update_DS members as Python pickles (prg_builder.py:145-147) -- the
    MPRG_PICKLE_DS=1 alternative to the table-shaped records of the native writer."""
    import pickle

    out = []
    for i in ok:
        build = engine.LocusBuild(LOCUS_OK, res.prg(i), res.n_nodes(i), res.n_sites(i), res.nodes(i))
        builder = PrgBuilder.from_engine(names[i], msas.alignment(i), build, options.max_nesting,
                                         options.min_match_length)
        # build_prg() also fills prg_index and the leaves' indexed_PRG_intervals: a statement, not an assert
        prg = builder.build_prg()
        if prg != build.prg:
            raise RuntimeError(f"PRG emission mismatch for {names[i]}")
        out.append((names[i], pickle.dumps(builder, protocol=4)))
    return out


def build_and_write(input_files, options, device_ordinal=0, output_prefix=None, part=False):
    """One GPU: the whole run, files in -> final files out.  Returns the number of PRGs written.
    The update_DS archive is written by the library as table-shaped records (mprg_writer_add_ds; objects are
    built on load, PrgBuilderZipDatabase) unless options.skip_update_ds; MPRG_PICKLE_DS=1 writes pickled
    PrgBuilder objects instead.
    part: the files are one part of a sharded run (hostio.merge_outputs makes the final files).
    Every file appears under its final name only when it is complete; an aborted run leaves none."""
    from concurrent.futures import ThreadPoolExecutor

    prefix = output_prefix or options.output_prefix
    ot = options.output_type
    want_ds = ot.prg and not getattr(options, "skip_update_ds", False)
    chunks = cut_chunks(input_files)
    writer = hostio.OutputWriter(prefix, prg=ot.prg, binary=ot.binary, gfa=ot.gfa,
                                 threads=side_threads(len(chunks), writer=True), part=part)
    pickle_ds = want_ds and bool(os.environ.get("MPRG_PICKLE_DS"))
    ds_zip = None
    ds_tmp = f"{prefix}.update_DS.zip.tmp{os.getpid()}"
    n_ok = 0
    pending = None  # (future, msas, res): the chunk being encoded / written on the writer thread

    def write_chunk(res, msas, ok, ok_names):
        writer.add(res, ok, ok_names)
        if want_ds and not pickle_ds and len(ok):
            writer.add_ds(res, msas, ok, ok_names, options.max_nesting, options.min_match_length)

    def finish(p):
        fut, msas, res = p
        try:
            fut.result()
        finally:
            res.free()
            msas.free()

    try:
        with ThreadPoolExecutor(1) as pool:
            for names, msas, res, ok in iter_built_chunks(input_files, options, device_ordinal, chunks):
                if pending is not None:
                    finish(pending)
                    pending = None
                if pickle_ds and ok:
                    if ds_zip is None:
                        ds_zip = zipfile.ZipFile(ds_tmp, "w")
                    for name, blob in _update_ds_pickles(names, msas, res, ok, options):
                        ds_zip.writestr(name, blob)
                n_ok += len(ok)
                ok_names = names if len(ok) == len(names) else [names[i] for i in ok]
                pending = (pool.submit(write_chunk, res, msas, ok, ok_names), msas, res)
            if pending is not None:
                finish(pending)
                pending = None
        writer.close()
        if ds_zip is not None:
            ds_zip.close()
            ds_zip = None
            os.replace(ds_tmp, prefix + ".update_DS.zip")
    except BaseException:
        if pending is not None:
            try:
                finish(pending)
            except Exception:
                pass
        writer.abort()
        if ds_zip is not None:
            ds_zip.close()
        if os.path.exists(ds_tmp):
            os.unlink(ds_tmp)
        raise
    return n_ok


def build_loci(input_files, options, device_ordinal=0, want_nodes=None):
    """Loads, builds and returns [(locus_name, alignment, LocusBuild)] for the successful loci, in
    input order (the in-memory form of a run)."""
    if want_nodes is None:
        want_nodes = options.output_type.prg and not getattr(options, "skip_update_ds", False)
    good = []
    for names, msas, res, ok in iter_built_chunks(input_files, options, device_ordinal):
        for i in ok:
            build = engine.LocusBuild(LOCUS_OK, res.prg(i), res.n_nodes(i), res.n_sites(i),
                                      res.nodes(i) if want_nodes else None)
            good.append((names[i], msas.alignment(i) if want_nodes else None, build))
        res.free()
        msas.free()
    return good


def run(options):
    logger.info("Getting input files...")
    input_files = get_all_input_files(options.input, options.suffix)
    if len(input_files) == 0:
        raise FileNotFoundError(f"No input files found in {options.input}")
    if not options.force and output_files_already_exist(options.output_type, options.output_prefix):
        raise RuntimeError("One or more output files already exists, aborting run...")
    Path(options.output_prefix).parent.mkdir(parents=True, exist_ok=True)
    gpus = max(1, int(getattr(options, "gpus", 1) or 1))
    logger.info(f"Using {gpus} GPU(s) to generate PRGs...")
    if gpus == 1:
        n_ok = build_and_write(input_files, options)
    else:
        n_ok = _run_sharded(input_files, options, gpus)
    logger.success("All PRGs generated!")
    if n_ok == 0:
        logger.error("No PRGs were built, please check errors")
    logger.success("All done!")


def _shard_worker(rank, shard_files, options, part_prefix, queue, n_shards=1):
    # every shard process drives its own GPU from this host: share the cores (mprg_create reads this)
    os.environ.setdefault("LOCAL_WORLD_SIZE", str(n_shards))
    try:
        n_ok = build_and_write(shard_files, options, device_ordinal=rank, output_prefix=part_prefix, part=True)
        queue.put((rank, n_ok, None))
    except BaseException as err:  # propagated to the parent like a Pool worker's exception
        import traceback

        queue.put((rank, None, (type(err).__name__, str(err), traceback.format_exc())))


def shard_files_lpt(input_files, gpus):
    """LPT partition of the loci by file size (a proxy of rows x cols), one shard per GPU."""
    costs = [os.path.getsize(p) for p in input_files]
    return engine.lpt_partition(costs, gpus)


def merge_parts(part_prefixes, options):
    """Final files from the shards' parts: native merge of .prg.fa / archives, update_DS members appended."""
    ot = options.output_type
    prefix = options.output_prefix
    want_ds = ot.prg and not getattr(options, "skip_update_ds", False)
    return hostio.merge_outputs(part_prefixes, prefix, prg=ot.prg, binary=ot.binary, gfa=ot.gfa, update_ds=want_ds)


def _run_sharded(input_files, options, gpus):
    """Loci are independent: LPT partition by file size, one process per GPU, no collective.  Every shard
    loads, builds and WRITES its own part (native writers); the parent only merges the parts -- nothing but
    a locus count comes back through the queue.  Returns the number of PRGs written."""
    import multiprocessing as mp
    import queue as queue_mod

    parts = shard_files_lpt(input_files, gpus)
    tmp_dir = tempfile.mkdtemp(prefix=".mprg_parts_", dir=str(Path(options.output_prefix).parent.resolve()))
    part_prefixes = [os.path.join(tmp_dir, f"part{rank}") for rank in range(len(parts))]
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = []
    try:
        for rank, idx in enumerate(parts):
            p = ctx.Process(target=_shard_worker,
                            args=(rank, [input_files[i] for i in idx], options, part_prefixes[rank], queue, gpus))
            p.start()
            procs.append(p)
        counts = {}
        while len(counts) < len(procs):
            try:
                rank, n_ok, err = queue.get(timeout=1.0)
            except queue_mod.Empty:
                # a shard that died without reporting (segfault, OOM killer, CUDA abort) must not hang the run
                dead = [r for r, p in enumerate(procs) if p.exitcode not in (None, 0) and r not in counts]
                if dead:
                    raise RuntimeError(f"shard process of GPU {dead[0]} exited with code {procs[dead[0]].exitcode} "
                                       "without a result")
                if all(p.exitcode is not None for p in procs) and queue.empty():
                    missing = [r for r in range(len(procs)) if r not in counts]
                    if missing:
                        raise RuntimeError(f"shard process of GPU {missing[0]} ended without a result")
                continue
            if err is not None:
                name, message, trace = err
                logger.error(f"shard of GPU {rank} failed:\n{trace}")
                if name == "EmptyMSAError":
                    raise EmptyMSAError(message)
                raise RuntimeError(f"{name}: {message}")
            counts[rank] = n_ok
        for p in procs:
            p.join()
        return merge_parts(part_prefixes, options) if sum(counts.values()) else 0
    finally:
        for p in procs:
            if p.is_alive():
                p.terminate()
        for p in procs:
            p.join(timeout=10)
        shutil.rmtree(tmp_dir, ignore_errors=True)
