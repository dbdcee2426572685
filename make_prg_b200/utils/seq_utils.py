"""Sequence utilities with the reference's names (make_prg/utils/seq_utils.py).  The O(rows x cols)
ones run on the GPU through libmprg; the string-level ones (IUPAC expansion of a handful of alleles)
stay on the host as in the north-star split."""
import itertools

import numpy as np

from .. import device
from ..msa import MSA, SeqRecord

NONMATCH = "*"
GAP = "-"


class SequenceCurationError(Exception):
    pass


def is_non_match(letter):
    return letter == NONMATCH


def is_gap(letter):
    return letter == GAP


def ungap(seq):
    return seq.replace(GAP, "")


def remove_duplicates(seqs):
    seen = set()
    for x in seqs:
        if x not in seen:
            seen.add(x)
            yield x


def get_alignment_seqs(alignment):
    for record in alignment:
        yield str(record.seq)


def _single_task(alignment):
    ctx = device.default_context()
    batch = ctx.upload([alignment.matrix])
    return ctx, batch, (0, None, 0, alignment.get_alignment_length())


def get_consensus_from_MSA(alignment):
    """seq_utils.py:219-239 -> kernel (a)."""
    if len(alignment) == 0 or alignment.get_alignment_length() == 0:
        return ""
    ctx, batch, task = _single_task(alignment)
    return ctx.scan_tasks(batch, [task])[0][0].decode()


def gap_reach(alignment):
    """int32[cols]: has_empty_sequence(alignment, (s, e)) == (gap_reach[s] >= e)."""
    ctx, batch, task = _single_task(alignment)
    return ctx.scan_tasks(batch, [task])[0][1]


def has_empty_sequence(alignment, interval):
    """seq_utils.py:37-42 -> kernel (a) in its gap-reach form."""
    if len(alignment) == 0:
        return False
    return bool(gap_reach(alignment)[interval[0]] >= interval[1])


def get_number_of_unique_ungapped_sequences(sub_alignment):
    ctx, batch, task = _single_task(sub_alignment)
    return ctx.dedupe_rows(batch, [task])[0][2]


def get_number_of_unique_gapped_sequences(sub_alignment):
    ctx, batch, task = _single_task(sub_alignment)
    return ctx.dedupe_rows(batch, [task])[0][3]


def remove_columns_full_of_gaps_from_MSA(alignment):
    """seq_utils.py:193-216: only shapes the stored node.alignment (never the PRG)."""
    M = alignment.matrix
    if M.size == 0:
        return MSA([SeqRecord("", r.id, r.name, r.description) for r in alignment])
    keep = ~(M == ord(GAP)).all(axis=0)
    K = M[:, keep]
    return MSA([SeqRecord(K[i].tobytes().decode(), r.id, r.name, r.description)
                for i, r in enumerate(alignment)])


class SequenceExpander:
    """seq_utils.py:77-158."""

    iupac = {"R": "GA", "Y": "TC", "K": "GT", "M": "AC", "S": "GC", "W": "AT",
             "A": "A", "C": "C", "G": "G", "T": "T"}
    expandable_bases = set(iupac)
    allowed_bases = expandable_bases | {"N"}
    standard_bases = {"A", "C", "G", "T"}
    ambiguous_bases = expandable_bases - standard_bases

    @classmethod
    def check_if_there_is_sequence_with_disallowed_bases(cls, sequences):
        for sequence in sequences:
            if not set(sequence) <= cls.allowed_bases:
                raise SequenceCurationError(
                    "A slice of a sequence has a disallowed base.\n"
                    f"Allowed bases: {cls.allowed_bases}.\nSequence: {sequence}\nRedo sequence curation.\n")

    @classmethod
    def get_expanded_sequences(cls, sequences):
        cls.check_if_there_is_sequence_with_disallowed_bases(sequences)
        out, seen = [], set()
        for seq in remove_duplicates(sequences):
            if "N" in seq:
                continue
            for combo in itertools.product(*(cls.iupac[b] for b in seq)):
                expanded = "".join(combo)
                if expanded not in seen:
                    seen.add(expanded)
                    out.append(expanded)
        if not out:
            raise SequenceCurationError(
                f"All sequences in this slice contained N. Redo sequence curation.\nSequences: {sequences}")
        return out

    @classmethod
    def get_expanded_sequences_from_MSA(cls, alignment):
        return cls.get_expanded_sequences([ungap(s) for s in get_alignment_seqs(alignment)])
