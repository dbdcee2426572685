"""IntervalPartitioner with the reference's interface (make_prg/from_msa/interval_partition.py); the
scan, the run state machine, the single-sequence demotion and the bijection check run in libmprg."""
from enum import Enum, auto

from .. import device
from .._lib import MprgError


class PartitioningError(Exception):
    pass


class IntervalType(Enum):
    Match = auto()
    NonMatch = auto()
    Root = auto()

    @classmethod
    def from_char(cls, letter):
        return IntervalType.NonMatch if letter == "*" else IntervalType.Match


def is_type(letter, interval_type):
    return IntervalType.from_char(letter) is interval_type


class Interval:
    """Closed interval [start, stop]."""

    def __init__(self, it_type, start, stop=None):
        self.type = it_type
        self.start = start
        if stop is not None:
            assert stop >= start
        self.stop = stop if stop is not None else start

    def modify_by(self, left_delta, right_delta):
        self.start += left_delta
        self.stop += right_delta

    def contains(self, position):
        return self.start <= position <= self.stop

    def __len__(self):
        return self.stop - self.start + 1

    def __lt__(self, other):
        return self.start < other.start

    def __eq__(self, other):
        return self.start == other.start and self.stop == other.stop and self.type is other.type

    def __hash__(self):
        return hash((self.start, self.stop, self.type))

    def __repr__(self):
        return f"[{self.start}, {self.stop}]"


def _to_intervals(arr):
    return [Interval(IntervalType.Match if int(a["type"]) == 0 else IntervalType.NonMatch,
                     int(a["start"]), int(a["stop"])) for a in arr]


class IntervalPartitioner:
    """IntervalPartitioner(consensus_string, min_match_length, alignment).get_intervals()"""

    def __init__(self, consensus_string, min_match_length, alignment):
        ctx = device.default_context()
        self.mml = min_match_length
        try:
            if len(alignment) == 0:
                # the reference's unit tests drive the state machine with hand-written strings
                arr = ctx.partition_consensus(consensus_string, min_match_length)
            else:
                batch = ctx.upload([alignment.matrix])
                task = (0, None, 0, alignment.get_alignment_length())
                device_consensus = ctx.scan_tasks(batch, [task])[0][0].decode()
                if device_consensus != consensus_string:
                    raise ValueError("consensus_string is not the consensus of the alignment")
                arr = ctx.partition_tasks(batch, [task], min_match_length)[0]
        except MprgError as err:
            if err.code == -4:
                raise PartitioningError(str(err)) from None
            raise
        intervals = _to_intervals(arr)
        self._match_intervals = [i for i in intervals if i.type is IntervalType.Match]
        self._non_match_intervals = [i for i in intervals if i.type is IntervalType.NonMatch]

    def get_intervals(self):
        return (sorted(self._match_intervals), sorted(self._non_match_intervals),
                sorted(self._match_intervals + self._non_match_intervals))
