"""Python face of the level-synchronous engine (mprg_build): batches of loci in, per-locus PRG strings
and recursion-tree tables out.  One Context per GPU; loci of one call share every kernel launch."""
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

from . import device
from ._lib import LOCUS_CURATION_ERROR, LOCUS_OK
from .utils.seq_utils import SequenceCurationError


@dataclass
class LocusBuild:
    status: int
    prg: str
    n_nodes: int
    n_sites: int
    nodes: Optional[Dict[str, np.ndarray]]

    def raise_for_status(self, locus_name=""):
        if self.status == LOCUS_CURATION_ERROR:
            raise SequenceCurationError(
                f"A slice of a sequence of {locus_name} has a disallowed base. Redo sequence curation.")
        if self.status != LOCUS_OK:
            raise ValueError(f"locus {locus_name} cannot be built (status {self.status}): the alignment "
                             "is empty or still holds N (load it with load_alignment_file)")


def build_matrices(matrices: List[np.ndarray], max_nesting: int, min_match_length: int,
                   ctx: Optional[device.Context] = None, want_nodes: bool = True,
                   max_batch_bytes: int = 8 << 30, parent_levels=None) -> List[LocusBuild]:
    """Runs the whole hot path on a list of uint8[rows, cols] ASCII matrices.  parent_levels[i] >= 0
    builds matrix i below an existing node of that nesting level (NodeFactory.build with a parent_node,
    recursion_tree.py:431-432), -1 / None as a locus root."""
    ctx = ctx or device.default_context()
    out: List[LocusBuild] = []
    start = 0
    while start < len(matrices):
        size, stop = 0, start
        while stop < len(matrices) and (stop == start or size + matrices[stop].size <= max_batch_bytes):
            size += matrices[stop].size
            stop += 1
        chunk = matrices[start:stop]
        if parent_levels is None:
            batch, res = ctx.build_ascii(chunk, max_nesting, min_match_length)
        else:
            batch = ctx.upload(chunk)
            res = ctx.build_sub(batch, max_nesting, min_match_length, list(parent_levels[start:stop]))
        for i in range(len(chunk)):
            status = res.status(i)
            out.append(LocusBuild(status, res.prg(i) if status == LOCUS_OK else "", res.n_nodes(i),
                                  res.n_sites(i),
                                  res.nodes(i) if (want_nodes and status == LOCUS_OK) else None))
        res.free()
        batch.free()
        start = stop
    return out


def lpt_partition(costs, n_parts):
    """Longest-processing-time-first greedy partition: returns n_parts index lists."""
    order = np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable")
    loads = [0.0] * n_parts
    parts = [[] for _ in range(n_parts)]
    for i in order:
        p = int(np.argmin(loads))
        parts[p].append(int(i))
        loads[p] += float(costs[i])
    return [sorted(p) for p in parts]
