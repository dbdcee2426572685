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

from loguru import logger

from .. import engine
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


def build_loci(input_files, options, device_ordinal=0):
    """Loads, builds and returns [(locus_name, alignment, LocusBuild)] for the successful loci, in
    input order.  Errors follow from_msa.py:142-151: an empty MSA aborts the run, a locus with a
    disallowed base is skipped with a warning."""
    from .. import device

    loci = []
    for path in input_files:
        name = remove_known_input_extensions(Path(path).name)
        logger.info(f"Generating PRG for {name}...")
        try:
            alignment = load_alignment_file(str(path), options.alignment_format)
        except ValueError as err:
            if "No records found in handle" in str(err.args[0]):
                raise EmptyMSAError(f"No records found in MSA of locus {name}")
            raise
        loci.append((name, alignment))
    ctx = device.default_context(device_ordinal)
    want_nodes = options.output_type.prg  # the update_DS pickles need the trees
    builds = engine.build_matrices([a.matrix for _, a in loci], options.max_nesting,
                                   options.min_match_length, ctx=ctx, want_nodes=want_nodes)
    good = []
    for (name, alignment), b in zip(loci, builds):
        if b.status == LOCUS_OK:
            good.append((name, alignment, b))
        elif b.status == LOCUS_CURATION_ERROR:
            logger.warning(f"Skipping building PRG for {name}. Error: a slice of a sequence has a "
                           "disallowed base. Redo sequence curation.")
        else:
            b.raise_for_status(name)
    return good


def write_outputs(good, options):
    """Final files exactly as InputOutputFiles.create_final_files lays them out."""
    prefix = options.output_prefix
    ot = options.output_type
    single = len(good) == 1
    if ot.prg:
        with open(prefix + ".prg.fa", "w") as fh:
            for name, _a, b in sorted(good, key=lambda t: t[0] + ".prg.fa"):
                fh.write(f">{name}\n{b.prg}\n")
        with zipfile.ZipFile(prefix + ".update_DS.zip", "w") as zf:
            for name, alignment, b in good:
                builder = PrgBuilder.from_engine(name, alignment, b, options.max_nesting,
                                                 options.min_match_length)
                assert builder.build_prg() == b.prg, f"PRG emission mismatch for {name}"
                import pickle

                zf.writestr(name, pickle.dumps(builder, protocol=4))
    if ot.binary:
        if single:
            Path(prefix + ".prg.bin").write_bytes(_bin_bytes(good[0][2].prg))
        else:
            with zipfile.ZipFile(prefix + ".prg.bin.zip", "w") as zf:
                for name, _a, b in good:
                    zf.writestr(f"{name}.bin", _bin_bytes(b.prg))
    if ot.gfa:
        if single:
            Path(prefix + ".prg.gfa").write_text(_gfa_text(good[0][2].prg))
        else:
            with zipfile.ZipFile(prefix + ".prg.gfa.zip", "w") as zf:
                for name, _a, b in good:
                    zf.writestr(f"{name}.gfa", _gfa_text(b.prg))


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
        good = build_loci(input_files, options)
    else:
        good = _run_sharded(input_files, options, gpus)
    logger.success("All PRGs generated!")
    if len(good) == 0:
        logger.error("No PRGs were built, please check errors")
    else:
        write_outputs(good, options)
    logger.success("All done!")


def _shard_worker(rank, shard_files, options, queue):
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
        p = ctx.Process(target=_shard_worker, args=(rank, [input_files[i] for i in idx], options, queue))
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
