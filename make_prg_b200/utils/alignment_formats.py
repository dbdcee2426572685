"""Alignment formats other than FASTA for `-f/--alignment-format` (make_prg/subcommands/from_msa.py:48-55 hands the
name to Bio.AlignIO.read, make_prg/utils/io_utils.py:22-31).  Biopython is not a dependency here; the parsers below
follow the layout rules of its ClustalIO / StockholmIO / PhylipIO readers for well-formed single-alignment files:
record order = order of first appearance, sequences of interleaved blocks are concatenated per record, characters
are kept as they are (case and N handling happen afterwards, exactly as for FASTA).  One alignment per file."""
import gzip

CLUSTAL_HEADERS = ("CLUSTAL", "PROBCONS", "MUSCLE", "MSAPROBS", "Kalign", "Biopython")
SUPPORTED = ("fasta", "clustal", "stockholm", "phylip", "phylip-sequential", "phylip-relaxed")


class AlignmentFormatError(ValueError):
    pass


def _read_text(path_or_handle):
    if hasattr(path_or_handle, "read"):
        return path_or_handle.read()
    path = str(path_or_handle)
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt") as fh:
        return fh.read()


def _check_rectangular(records, what):
    if not records:
        raise AlignmentFormatError(f"No records found in {what} alignment")
    width = len(records[0][1])
    for rid, seq in records:
        if len(seq) != width:
            raise AlignmentFormatError(f"Sequences must all be the same length ({rid}: {len(seq)} != {width})")
    return records


def parse_clustal(text):
    lines = text.splitlines()
    i = 0
    while i < len(lines) and not lines[i].strip():
        i += 1
    if i == len(lines) or not lines[i].startswith(CLUSTAL_HEADERS):
        raise AlignmentFormatError("not a Clustal file: the first line names no known program")
    order, seqs = [], {}
    block_pos = 0
    for line in lines[i + 1:]:
        if not line.strip():
            block_pos = 0
            continue
        if line[0] in " \t":  # conservation line of the block
            continue
        fields = line.rstrip().split()
        if len(fields) < 2 or len(fields) > 3:
            raise AlignmentFormatError(f"Could not parse line: {line!r}")
        rid, chunk = fields[0], fields[1]
        if rid not in seqs:
            if block_pos != len(order):
                raise AlignmentFormatError(f"record {rid} first appears after the first block")
            order.append(rid)
            seqs[rid] = []
        elif block_pos >= len(order) or order[block_pos] != rid:
            raise AlignmentFormatError(f"records out of order in a later block: {rid}")
        seqs[rid].append(chunk)
        if len(fields) == 3:
            try:
                letters = int(fields[2])
            except ValueError:
                raise AlignmentFormatError(f"Could not parse line, bad sequence number: {line!r}") from None
            have = sum(len(c.replace("-", "")) for c in seqs[rid])
            if have != letters:
                raise AlignmentFormatError(f"Could not parse line, invalid sequence number: {line!r}")
        block_pos += 1
    return _check_rectangular([(rid, "".join(seqs[rid])) for rid in order], "Clustal")


def parse_stockholm(text):
    lines = text.splitlines()
    i = 0
    while i < len(lines) and not lines[i].strip():
        i += 1
    if i == len(lines) or lines[i].strip() != "# STOCKHOLM 1.0":
        raise AlignmentFormatError("Did not find STOCKHOLM header")
    order, seqs = [], {}
    for line in lines[i + 1:]:
        line = line.strip()
        if line == "//":
            break
        if not line or line.startswith("#"):
            continue
        parts = line.split(None, 1)
        if len(parts) != 2:
            raise AlignmentFormatError(f"Could not split line into identifier and sequence: {line!r}")
        rid, chunk = parts[0], parts[1].replace(" ", "")
        if rid not in seqs:
            order.append(rid)
            seqs[rid] = []
        seqs[rid].append(chunk)
    return _check_rectangular([(rid, "".join(seqs[rid])) for rid in order], "Stockholm")


def _phylip_counts(lines):
    i = 0
    while i < len(lines) and not lines[i].strip():
        i += 1
    if i == len(lines):
        raise AlignmentFormatError("empty PHYLIP file")
    fields = lines[i].split()
    if len(fields) != 2 or not all(f.isdigit() for f in fields):
        raise AlignmentFormatError("First line should have two integers")
    return i + 1, int(fields[0]), int(fields[1])


def _phylip_split_id(line, relaxed):
    if relaxed:
        parts = line.strip().split(None, 1)
        if len(parts) != 2:
            raise AlignmentFormatError(f"Could not split line into identifier and sequence: {line!r}")
        return parts[0], parts[1].replace(" ", "")
    return line[:10].strip(), line[10:].strip().replace(" ", "")


def parse_phylip(text, relaxed=False, sequential=False):
    lines = text.splitlines()
    at, n, width = _phylip_counts(lines)
    ids, seqs = [], []
    if sequential:
        body = [ln for ln in lines[at:] if ln.strip()]
        k = 0
        for _ in range(n):
            if k >= len(body):
                raise AlignmentFormatError("Premature end of file")
            rid, chunk = _phylip_split_id(body[k], relaxed)
            k += 1
            parts = [chunk]
            while sum(map(len, parts)) < width:
                if k >= len(body):
                    raise AlignmentFormatError("Premature end of file")
                parts.append(body[k].strip().replace(" ", ""))
                k += 1
            ids.append(rid)
            seqs.append("".join(parts))
    else:
        k = at
        while k < len(lines) and not lines[k].strip():
            k += 1
        for _ in range(n):
            if k >= len(lines) or not lines[k].strip():
                raise AlignmentFormatError("Premature end of file")
            rid, chunk = _phylip_split_id(lines[k], relaxed)
            ids.append(rid)
            seqs.append([chunk])
            k += 1
        while True:
            while k < len(lines) and not lines[k].strip():
                k += 1
            if k >= len(lines):
                break
            for r in range(n):
                if k >= len(lines) or not lines[k].strip():
                    raise AlignmentFormatError("Premature end of file in an interleaved block")
                seqs[r].append(lines[k].strip().replace(" ", ""))
                k += 1
        seqs = ["".join(parts) for parts in seqs]
    for rid, seq in zip(ids, seqs):
        if len(seq) != width:
            raise AlignmentFormatError(f"Sequence {rid} has length {len(seq)}, the header says {width}")
    return _check_rectangular(list(zip(ids, seqs)), "PHYLIP")


def read_records(path_or_handle, alignment_format):
    """[(id, sequence)] of the one alignment in the file."""
    fmt = alignment_format.lower()
    if fmt not in SUPPORTED or fmt == "fasta":
        raise AlignmentFormatError(f"alignment format {alignment_format!r} is not supported (one of {SUPPORTED})")
    text = _read_text(path_or_handle)
    if fmt == "clustal":
        return parse_clustal(text)
    if fmt == "stockholm":
        return parse_stockholm(text)
    return parse_phylip(text, relaxed=fmt == "phylip-relaxed", sequential=fmt == "phylip-sequential")


def to_fasta_text(records):
    return "".join(f">{rid}\n{seq}\n" for rid, seq in records)
