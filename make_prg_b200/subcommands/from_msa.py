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
    use every core; with several, loading chunk k+1 and writing chunk k-1 run beside the build of chunk
    k, whose worker threads drive the device and must not be starved of cores."""
    cores = os.cpu_count() or 1
    if n_chunks <= 1:
        return max(1, min(32, cores))
    return max(1, min(4, cores // 8)) if writer else max(1, min(12, cores // 3))


def _load_chunk(paths, alignment_format, threads=None):
    if alignment_format != "fasta":
        raise ValueError(f"only the fasta alignment format is supported, got {alignment_format}")
    return hostio.load_fasta_files(paths, threads=threads)


_EXTENSIONS = (".fasta.gz", ".fa.gz", ".fasta", ".fa")


def locus_name_of(path):
    """remove_known_input_extensions(Path(path).name) without the Path object and the regular expression."""
    name = os.path.basename(os.fspath(path))
    for ext in _EXTENSIONS:
        if name.endswith(ext):
            return name[:-len(ext)]
    return name


def iter_built_chunks(input_files, options, device_ordinal=0, chunks=None):
    """Loads (native loader, one chunk ahead on a host thread) and builds (mprg_build_ascii) the input
    files chunk by chunk.  Yields (names, msas, result, ok) with ok = indices of the chunk's loci that
    were built; the consumer frees result and msas.  Errors follow from_msa.py:142-151: an empty MSA
    aborts the run, a locus with a disallowed base is skipped with a warning."""
    from concurrent.futures import ThreadPoolExecutor

    from .. import device

    ctx = device.default_context(device_ordinal)
    if chunks is None:
        chunks = cut_chunks(input_files)
    with ThreadPoolExecutor(1) as pool:
        threads = side_threads(len(chunks))
        future = pool.submit(_load_chunk, chunks[0], options.alignment_format, threads) if chunks else None
        for k, paths in enumerate(chunks):
            msas = future.result()
            future = (pool.submit(_load_chunk, chunks[k + 1], options.alignment_format, threads)
                      if k + 1 < len(chunks) else None)
            names = [locus_name_of(path) for path in paths]
            logger.info(f"Generating PRGs for {names[0]} ... {names[-1]} ({len(names)} loci)...")
            for i in np.nonzero(msas.status != hostio.LOAD_OK)[0]:
                try:
                    hostio.raise_for_load_status(msas, int(i))
                except ValueError as err:
                    if "No records found in handle" in str(err.args[0]):
                        raise EmptyMSAError(f"No records found in MSA of locus {names[int(i)]}")
                    raise
            batch, res = ctx.build_msa_set(msas, options.max_nesting, options.min_match_length)
            batch.free()
            statuses, _lengths = res.statuses()
            ok = np.nonzero(statuses == LOCUS_OK)[0].tolist()
            for i in np.nonzero(statuses != LOCUS_OK)[0].tolist():
                if statuses[i] == LOCUS_CURATION_ERROR:
                    logger.warning(f"Skipping building PRG for {names[i]}. Error: a slice of a sequence has a "
                                   "disallowed base. Redo sequence curation.")
                else:
                    engine.LocusBuild(int(statuses[i]), "", 0, 0, None).raise_for_status(names[i])
            yield names, msas, res, ok


def _update_ds_pickles(names, msas, res, ok, options):
    """(name, pickled PrgBuilder) for the update_DS archive (prg_builder.py:145-147)."""
    import pickle

    out = []
    for i in ok:
        build = engine.LocusBuild(LOCUS_OK, res.prg(i), res.n_nodes(i), res.n_sites(i), res.nodes(i))
        builder = PrgBuilder.from_engine(names[i], msas.alignment(i), build, options.max_nesting,
                                         options.min_match_length)
        assert builder.build_prg() == build.prg, f"PRG emission mismatch for {names[i]}"
        out.append((names[i], pickle.dumps(builder, protocol=4)))
    return out


def build_and_write(input_files, options, device_ordinal=0, output_prefix=None):
    """One GPU: the whole run, files in -> final files out.  Returns the number of PRGs written.
    The update_DS archive (Python PrgBuilder pickles) is written unless options.skip_update_ds."""
    from concurrent.futures import ThreadPoolExecutor

    prefix = output_prefix or options.output_prefix
    ot = options.output_type
    want_ds = ot.prg and not getattr(options, "skip_update_ds", False)
    chunks = cut_chunks(input_files)
    writer = hostio.OutputWriter(prefix, prg=ot.prg, binary=ot.binary, gfa=ot.gfa,
                                 threads=side_threads(len(chunks), writer=True))
    ds_zip = None
    n_ok = 0
    pending = None  # (future, msas, res): the chunk being encoded / written on the writer thread

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
                if want_ds and ok:
                    if ds_zip is None:
                        ds_zip = zipfile.ZipFile(prefix + ".update_DS.zip", "w")
                    for name, blob in _update_ds_pickles(names, msas, res, ok, options):
                        ds_zip.writestr(name, blob)
                n_ok += len(ok)
                ok_names = names if len(ok) == len(names) else [names[i] for i in ok]
                pending = (pool.submit(writer.add, res, ok, ok_names), msas, res)
            if pending is not None:
                finish(pending)
                pending = None
        writer.close()
    except BaseException:
        if pending is not None:
            try:
                finish(pending)
            except Exception:
                pass
        writer.abort()
        raise
    finally:
        if ds_zip is not None:
            ds_zip.close()
    return n_ok


def build_loci(input_files, options, device_ordinal=0, want_nodes=None):
    """Loads, builds and returns [(locus_name, alignment, LocusBuild)] for the successful loci, in
    input order (the in-memory form of a run, used by the multi-GPU shards)."""
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


def write_outputs(good, options):
    """Final files exactly as InputOutputFiles.create_final_files lays them out, from in-memory builds
    (the native writers take the PRG strings; the update_DS pickles are Python objects)."""
    prefix = options.output_prefix
    ot = options.output_type
    strings = hostio.PrgStrings([b.prg for _n, _a, b in good])
    writer = hostio.OutputWriter(prefix, prg=ot.prg, binary=ot.binary, gfa=ot.gfa)
    try:
        writer.add(strings, np.arange(len(good), dtype=np.int32), [name for name, _a, _b in good])
        writer.close()
    finally:
        writer.abort()
        strings.free()
    if ot.prg and not getattr(options, "skip_update_ds", False):
        import pickle

        with zipfile.ZipFile(prefix + ".update_DS.zip", "w") as zf:
            for name, alignment, b in good:
                builder = PrgBuilder.from_engine(name, alignment, b, options.max_nesting,
                                                 options.min_match_length)
                assert builder.build_prg() == b.prg, f"PRG emission mismatch for {name}"
                zf.writestr(name, pickle.dumps(builder, protocol=4))


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
        logger.success("All PRGs generated!")
        if n_ok == 0:
            logger.error("No PRGs were built, please check errors")
    else:
        good = _run_sharded(input_files, options, gpus)
        logger.success("All PRGs generated!")
        if len(good) == 0:
            logger.error("No PRGs were built, please check errors")
        else:
            write_outputs(good, options)
    logger.success("All done!")


def _shard_worker(rank, shard_files, options, queue, n_shards=1):
    # every shard process drives its own GPU from this host: share the cores (mprg_create reads this)
    os.environ.setdefault("LOCAL_WORLD_SIZE", str(n_shards))
    try:
        queue.put((rank, build_loci(shard_files, options, device_ordinal=rank), None))
    except Exception as err:  # propagated to the parent like a Pool worker's exception
        queue.put((rank, None, err))


def _run_sharded(input_files, options, gpus):
    """Loci are independent: LPT partition by file size (a proxy of rows x cols), one process per GPU,
    no collective; results are gathered on the host and re-ordered to the input order."""
    import multiprocessing as mp

    costs = [os.path.getsize(p) for p in input_files]
    parts = engine.lpt_partition(costs, gpus)
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = []
    for rank, idx in enumerate(parts):
        p = ctx.Process(target=_shard_worker, args=(rank, [input_files[i] for i in idx], options, queue, gpus))
        p.start()
        procs.append(p)
    results = {}
    for _ in procs:
        rank, good, err = queue.get()
        if err is not None:
            for p in procs:
                p.terminate()
            raise err
        results[rank] = good
    for p in procs:
        p.join()
    order = {remove_known_input_extensions(Path(f).name): i for i, f in enumerate(input_files)}
    merged = [g for rank in sorted(results) for g in results[rank]]
    return sorted(merged, key=lambda t: order[t[0]])
