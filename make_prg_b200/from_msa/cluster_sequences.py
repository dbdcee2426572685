"""kmeans_cluster_seqs with the reference's interface (make_prg/from_msa/cluster_sequences.py:211-296).
De-duplication, k-mer counting, the one-reference-like test and KMeans run in libmprg; the host only
turns cluster indices back into the reference's id lists (ordering rules of extract_clusters /
merge_clusters, :114-133, :194-208)."""
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .. import device
from ..utils.seq_utils import SequenceExpander, ungap

DISTANCE_THRESHOLD = 0.2
LENGTH_THRESHOLD = 5
MAX_CLUSTERS = 10


@dataclass
class ClusteringResult:
    clustered_ids: List[List[str]]
    sequences: Optional[List[str]] = None

    @property
    def no_clustering(self):
        return len(self.clustered_ids) == 1

    @property
    def have_precomputed_sequences(self):
        return self.sequences is not None


def merge_sequences(*seqlists, first_seq):
    rest = [s for sl in seqlists for s in sl if s != first_seq]
    assert any(s == first_seq for sl in seqlists for s in sl), "first sequence not found"
    return SequenceExpander.get_expanded_sequences([first_seq] + rest)


def count_kmer_matrix(alignment, kmer_size):
    """Dense float64 count matrix of the distinct ungapped sequences of length >= kmer_size
    (count_distinct_kmers + count_kmer_occurrences, :26-56) -> kernel (b)."""
    ctx = device.default_context()
    batch = ctx.upload([alignment.matrix])
    return ctx.kmer_counts(batch, (0, None, 0, alignment.get_alignment_length()), kmer_size)


def kmeans_cluster_seqs(alignment, kmer_size):
    ctx = device.default_context()
    batch = ctx.upload([alignment.matrix])
    task = (0, None, 0, alignment.get_alignment_length())
    clusters = ctx.cluster_tasks(batch, [task], kmer_size)[0]
    group, ulen, _, _ = ctx.dedupe_rows(batch, [task])[0]
    ids = alignment.ids

    def ordered(members, single):
        key = (lambda r: (ulen[r] < kmer_size, group[r], r)) if single else (lambda r: (group[r], r))
        rows = sorted(members, key=key)
        return rows

    single = len(clusters) == 1
    out = []
    for ci, members in enumerate(clusters):
        rows = ordered(members, single)
        if 0 in rows:  # merge_clusters: the first record leads its cluster
            rows.remove(0)
            rows.insert(0, 0)
        out.append([ids[r] for r in rows])
    if single:
        seqs = [ungap(rec.seq) for rec in alignment]
        long_seqs = list(dict.fromkeys(s for s in seqs if len(s) >= kmer_size))
        small_seqs = list(dict.fromkeys(s for s in seqs if len(s) < kmer_size))
        return ClusteringResult(out, merge_sequences(long_seqs, small_seqs, first_seq=seqs[0]))
    return ClusteringResult(out)
