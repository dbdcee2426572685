"""Loader: FASTA (optionally .gz) -> MSA, upper-cased, N replaced by the per-column majority symbol.
Mirrors make_prg/utils/io_utils.py:17-49 and make_prg/utils/seq_utils.py:242-290 (the private RNG is
seeded with sha256 of the concatenated rows and `choice` is drawn once per column, in column order,
so the stream position is identical to the reference's)."""
import gzip
import hashlib
import random
from collections import Counter
from io import StringIO
from pathlib import Path

from ..msa import MSA, SeqRecord


def parse_fasta(handle):
    titles, chunks = [], []
    # Biopython 1.79 SimpleFastaParser: title = line[1:].rstrip(); every sequence line is rstripped,
    # then blanks and carriage returns are removed from the joined sequence
    for raw in handle:
        if raw.startswith(">"):
            titles.append(raw[1:].rstrip())
            chunks.append([])
        elif titles:
            chunks[-1].append(raw.rstrip().replace(" ", "").replace("\r", ""))
    if not titles:
        raise ValueError("No records found in handle")
    records = []
    for title, parts in zip(titles, chunks):
        tokens = title.split(None, 1)
        rid = tokens[0] if tokens else ""
        records.append(SeqRecord("".join(parts), rid, rid, title))
    return MSA(records)


def majority_consensus_of_rows(seqs):
    """seq_utils.py:246-290 on upper-cased row strings."""
    rng = random.Random()
    rng.seed(hashlib.sha256("".join(seqs).encode()).digest())
    consensus = []
    for i in range(len(seqs[0]) if seqs else 0):
        counts = Counter(s[i] for s in seqs if s[i] != "-" and s[i] != "N")
        if not counts:
            consensus.append(rng.choice("ACGT"))
            continue
        top = counts.most_common(1)[0][1]
        consensus.append(rng.choice([res for res, c in counts.items() if c == top]))
    return "".join(consensus)


def get_majority_consensus_from_MSA(alignment):
    return majority_consensus_of_rows([r.seq.upper() for r in alignment])


def load_alignment_file(msa_file, alignment_format="fasta"):
    if alignment_format.lower() != "fasta":
        # clustal / stockholm / phylip*: utils/alignment_formats.py (no Biopython here)
        from .alignment_formats import read_records

        alignment = MSA([SeqRecord(seq, rid, rid, rid) for rid, seq in read_records(msa_file, alignment_format)])
    elif isinstance(msa_file, StringIO):
        alignment = parse_fasta(msa_file)
    else:
        path = str(msa_file)
        opener = gzip.open if path.endswith(".gz") else open
        with opener(path, "rt") as handle:
            alignment = parse_fasta(handle)
    for record in alignment:
        record.seq = record.seq.upper()
    if any("N" in record.seq for record in alignment):
        consensus = get_majority_consensus_from_MSA(alignment)
        for record in alignment:
            if "N" in record.seq:
                record.seq = "".join(consensus[i] if ch == "N" else ch for i, ch in enumerate(record.seq))
    alignment._matrix = None
    return alignment


def locus_name_of(path):
    """input_output_files.py:234-235: file name minus .fa/.fasta(.gz)."""
    name = Path(path).name
    for suffix in (".gz",):
        if name.endswith(suffix):
            name = name[: -len(suffix)]
    for suffix in (".fasta", ".fa"):
        if name.endswith(suffix):
            name = name[: -len(suffix)]
            break
    return name
