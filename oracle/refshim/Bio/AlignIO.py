from Bio.Align import MultipleSeqAlignment
from Bio.Seq import Seq
from Bio.SeqRecord import SeqRecord


def read(handle, fmt):
    if fmt != "fasta":
        raise ValueError(f"stand-in only reads fasta, got {fmt}")
    titles, chunks = [], []
    for raw in handle:
        line = raw.rstrip("\r\n")
        if line.startswith(">"):
            titles.append(line[1:])
            chunks.append([])
        elif titles:
            chunks[-1].append(line.replace(" ", ""))
    if not titles:
        raise ValueError("No records found in handle")
    records = []
    for title, parts in zip(titles, chunks):
        tokens = title.split(None, 1)
        first = tokens[0] if tokens else ""
        records.append(SeqRecord(Seq("".join(parts)), id=first, name=first, description=title))
    return MultipleSeqAlignment(records)
