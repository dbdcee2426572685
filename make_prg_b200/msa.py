"""Minimal in-memory MSA container with the surface of Biopython's MultipleSeqAlignment that the
from_msa path uses (the reference aliases it as make_prg.MSA, make_prg/__init__.py:7-9).  Rows are
kept as one uint8 matrix so that an alignment can be handed to the device without conversion."""
import numpy as np


class SeqRecord:
    __slots__ = ("id", "name", "description", "seq")

    def __init__(self, seq, id="<unknown id>", name=None, description=None):
        self.seq = str(seq)
        self.id = id
        self.name = id if name is None else name
        self.description = id if description is None else description

    def __len__(self):
        return len(self.seq)

    def __iter__(self):
        return iter(self.seq)

    def __getitem__(self, index):
        if isinstance(index, slice):
            return SeqRecord(self.seq[index], self.id, self.name, self.description)
        return self.seq[index]

    def format(self, fmt):
        if fmt != "fasta":
            raise ValueError(f"unsupported format {fmt}")
        desc = self.description
        if desc and desc.split(None, 1)[:1] == [self.id]:
            title = desc
        elif desc and desc != "<unknown description>":
            title = f"{self.id} {desc}"
        else:
            title = self.id
        body = "".join(self.seq[i:i + 60] + "\n" for i in range(0, len(self.seq), 60))
        return f">{title}\n{body}"


class MSA:
    """Rows x columns alignment.  `MSA(records)` like Bio.Align.MultipleSeqAlignment."""

    def __init__(self, records=()):
        self._records = list(records)
        if self._records:
            width = len(self._records[0])
            if any(len(r) != width for r in self._records):
                raise ValueError("Sequences must all be the same length")
        self._matrix = None

    @classmethod
    def from_matrix(cls, ids, matrix, descriptions=None):
        recs = [SeqRecord(matrix[i].tobytes().decode(), ids[i],
                          description=None if descriptions is None else descriptions[i])
                for i in range(len(ids))]
        out = cls(recs)
        out._matrix = np.ascontiguousarray(matrix, np.uint8)
        return out

    @property
    def ids(self):
        return [r.id for r in self._records]

    @property
    def matrix(self):
        """uint8[rows, cols] ASCII view of the alignment."""
        if self._matrix is None:
            n, w = len(self._records), self.get_alignment_length()
            if n == 0 or w == 0:
                self._matrix = np.zeros((n, w), np.uint8)
            else:
                self._matrix = np.frombuffer("".join(r.seq for r in self._records).encode(),
                                             np.uint8).reshape(n, w).copy()
        return self._matrix

    def __len__(self):
        return len(self._records)

    def __iter__(self):
        return iter(self._records)

    def get_alignment_length(self):
        return len(self._records[0]) if self._records else 0

    def __getitem__(self, index):
        if isinstance(index, (int, np.integer)):
            return self._records[int(index)]
        if isinstance(index, slice):
            return MSA(self._records[index])
        rows, cols = index
        if isinstance(rows, slice) and isinstance(cols, slice):
            return MSA([rec[cols] for rec in self._records[rows]])
        raise TypeError(f"unsupported index {index!r}")

    def select_rows(self, row_indices):
        return MSA([self._records[int(i)] for i in row_indices])

    def __format__(self, fmt):
        if fmt != "fasta":
            raise ValueError(f"unsupported format {fmt}")
        return "".join(rec.format("fasta") for rec in self._records)

    def format(self, fmt):
        return self.__format__(fmt)
