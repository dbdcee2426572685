from Bio.SeqRecord import SeqRecord  # noqa: F401


class MultipleSeqAlignment:
    def __init__(self, records=()):
        self._records = list(records)
        if self._records:
            width = len(self._records[0])
            if any(len(r) != width for r in self._records):
                raise ValueError("Sequences must all be the same length")

    def __len__(self):
        return len(self._records)

    def __iter__(self):
        return iter(self._records)

    def get_alignment_length(self):
        return len(self._records[0]) if self._records else 0

    def __getitem__(self, index):
        if isinstance(index, int):
            return self._records[index]
        if isinstance(index, slice):
            return MultipleSeqAlignment(self._records[index])
        rows, cols = index
        if isinstance(rows, slice) and isinstance(cols, slice):
            return MultipleSeqAlignment([rec[cols] for rec in self._records[rows]])
        raise TypeError(f"unsupported index {index!r}")

    def __format__(self, fmt):
        if fmt != "fasta":
            raise ValueError(fmt)
        return "".join(rec.format("fasta") for rec in self._records)

    def format(self, fmt):
        return self.__format__(fmt)
