"""Recursion-tree node classes with the reference's fields and PRG framing
(make_prg/recursion_tree.py:27-391).  Trees are not grown by Python recursion: NodeFactory.build sends
the alignment through the level-synchronous engine and materialises the nodes from its pre-order node
table (node ids are the pre-order positions, as in the reference where ids are handed out in the
constructor before the children are built)."""
from abc import ABC, abstractmethod
from typing import List, Optional

import numpy as np

from . import engine
from ._lib import NODE_CLUSTER, NODE_INTERVAL, NODE_LEAF
from .msa import MSA, SeqRecord
from .utils.seq_utils import SequenceExpander, remove_columns_full_of_gaps_from_MSA


def equal_msas(msa_1, msa_2):
    return format(msa_1, "fasta") == format(msa_2, "fasta")


class RootAlignment:
    """The root MSA of a locus as one uint8 matrix + record labels: what every node of the locus' tree
    slices.  One object per locus, shared by its nodes (so a pickled PrgBuilder stores the rows once)."""

    def __init__(self, alignment: MSA):
        self.matrix = np.ascontiguousarray(alignment.matrix, np.uint8)
        self.ids = [r.id for r in alignment]
        self.names = [r.name for r in alignment]
        self.descriptions = [r.description for r in alignment]


class SubAlignment:
    """(row subset, column window) of a RootAlignment: the sub-alignment of a tree node, kept as indices
    and materialised as an MSA (columns full of gaps removed, recursion_tree.py:45) on first use.  Every
    sub-alignment the reference builds is such a slice."""

    __slots__ = ("root", "rows", "c0", "c1")

    def __init__(self, root, rows, c0, c1):
        self.root, self.rows, self.c0, self.c1 = root, rows, int(c0), int(c1)

    def matrix(self):
        M = self.root.matrix
        return M[:, self.c0:self.c1] if self.rows is None else M[self.rows, self.c0:self.c1]

    def row_numbers(self):
        return range(self.root.matrix.shape[0]) if self.rows is None else [int(r) for r in self.rows]

    def ungapped_sequences(self):
        """Ungapped rows in row order (the input of SequenceExpander, seq_utils.py:155-158)."""
        return [row.tobytes().replace(b"-", b"").decode() for row in self.matrix()]

    def materialise(self):
        S = self.matrix()
        root = self.root
        if S.size:
            S = S[:, ~(S == ord("-")).all(axis=0)]
        texts = [row.tobytes().decode() for row in S] if S.size else [""] * S.shape[0]
        return MSA([SeqRecord(text, root.ids[r], root.names[r], root.descriptions[r])
                    for text, r in zip(texts, self.row_numbers())])


class RecursiveTreeNode(ABC):
    def __init__(self, nesting_level, alignment, parent, prg_builder, node_id):
        self.nesting_level = nesting_level
        if isinstance(alignment, SubAlignment):
            self._sub, self._alignment = alignment, None
        else:
            self._sub, self._alignment = None, remove_columns_full_of_gaps_from_MSA(alignment)
        self.parent = parent
        self.prg_builder = prg_builder
        self._node_id = node_id
        self._children: List["RecursiveTreeNode"] = []

    @property
    def alignment(self):
        if self._alignment is None:
            self._alignment = self._sub.materialise()
        return self._alignment

    @alignment.setter
    def alignment(self, value):
        self._sub, self._alignment = None, value

    def __getstate__(self):
        state = self.__dict__.copy()
        if state.get("_sub") is not None:
            state["_alignment"] = None  # re-materialised from the shared root matrix after loading
        return state

    @property
    def node_id(self):
        return self._node_id

    @property
    def children(self):
        return self._children

    def is_leaf(self):
        return len(self.children) == 0

    def is_root(self):
        return self.parent is None

    def replace_child(self, old_child, new_child):
        assert old_child in self.children, f"Failure to replace a child, {old_child} does not exist"
        self.children[self.children.index(old_child)] = new_child

    def __eq__(self, other):
        if (self.nesting_level, self.prg_builder.locus_name, self.node_id) != (
                other.nesting_level, other.prg_builder.locus_name, other.node_id):
            return False
        if (self.parent is None) != (other.parent is None):
            return False
        if self.parent is not None and self.parent.node_id != other.parent.node_id:
            return False
        if not equal_msas(self.alignment, other.alignment):
            return False
        if len(self.children) != len(other.children):
            return False
        return all(a == b for a, b in zip(self.children, other.children))

    def __hash__(self):
        return hash((self.node_id, self.prg_builder.locus_name))

    @abstractmethod
    def preorder_traversal_to_build_prg(self, prg_as_list, delim_char=" "):
        raise NotImplementedError

    def __repr__(self):
        return (f"{self.__class__.__name__}:\nId = {self.node_id}\nNesting level = {self.nesting_level}\n"
                f"Parent = {'None' if self.parent is None else f'Id = {self.parent.node_id}'}\n"
                f"Children = [{', '.join(f'Id = {c.node_id}' for c in self.children)}]\n"
                f"Alignment:\n{format(self.alignment, 'fasta')}")


class MultiIntervalNode(RecursiveTreeNode):
    """Vertical partition of an MSA: the PRG is the concatenation of the children's PRGs."""

    def preorder_traversal_to_build_prg(self, prg_as_list, delim_char=" "):
        for child in self.children:
            child.preorder_traversal_to_build_prg(prg_as_list, delim_char)


class MultiClusterNode(RecursiveTreeNode):
    """Horizontal partition (sequence clusters): opens a site, one allele per child."""

    def preorder_traversal_to_build_prg(self, prg_as_list, delim_char=" "):
        site_num = self.prg_builder.get_next_site_num()
        prg_as_list.extend(f"{delim_char}{site_num}{delim_char}")
        last = len(self.children) - 1
        for child_index, child in enumerate(self.children):
            child.preorder_traversal_to_build_prg(prg_as_list, delim_char)
            marker = site_num + 1 if child_index < last else site_num
            prg_as_list.extend(f"{delim_char}{marker}{delim_char}")


class UpdateError(Exception):
    pass


class LeafNode(RecursiveTreeNode):
    """MSAs that are never partitioned; the only nodes that get indexed."""

    def __init__(self, nesting_level, alignment, parent, prg_builder, node_id):
        super().__init__(nesting_level, alignment, parent, prg_builder, node_id)
        self.new_sequences = set()
        self.indexed_PRG_intervals = set()

    def preorder_traversal_to_build_prg(self, prg_as_list, delim_char=" ", do_indexing=True):
        if self._sub is not None:  # straight from the root matrix: no MSA objects for the leaf
            expanded = SequenceExpander.get_expanded_sequences(self._sub.ungapped_sequences())
        else:
            expanded = SequenceExpander.get_expanded_sequences_from_MSA(self.alignment)
        if len(expanded) == 1:
            start = len(prg_as_list)
            prg_as_list.extend(expanded[0])
            if do_indexing:
                self.prg_builder.update_PRG_index(start, len(prg_as_list), node=self)
            return
        site_num = self.prg_builder.get_next_site_num()
        prg_as_list.extend(f"{delim_char}{site_num}{delim_char}")
        last = len(expanded) - 1
        for seq_index, seq in enumerate(expanded):
            start = len(prg_as_list)
            prg_as_list.extend(seq)
            end = len(prg_as_list)
            marker = site_num + 1 if seq_index < last else site_num
            prg_as_list.extend(f"{delim_char}{marker}{delim_char}")
            if do_indexing:
                self.prg_builder.update_PRG_index(start, end, node=self)

    # ---- update methods (make_prg/recursion_tree.py:303-388) --------------------------------------------
    def add_data_to_batch_update(self, update_data):
        """Process the given update data and add a new sequence to self.new_sequences (:305-343)."""
        interval = update_data.ml_path_node_key
        if interval not in self.indexed_PRG_intervals:
            raise UpdateError(f"PRG interval {interval} not found in indexed PRG intervals for node: "
                              f"{self.indexed_PRG_intervals}")
        parts = []
        for prg_interval in sorted(self.indexed_PRG_intervals):
            if prg_interval == interval:
                parts.append(update_data.new_node_sequence)
            else:
                try:
                    parts.append(update_data.ml_path.get_node_given_interval_in_PRG_space(prg_interval).sequence)
                except Exception as err:  # MLPathError of the (out-of-scope) denovo-path parser
                    if type(err).__name__ != "MLPathError":
                        raise
        self.new_sequences.add("".join(parts))

    def add_indexed_PRG_interval(self, interval):
        self.indexed_PRG_intervals.add(interval)

    def batch_update(self):
        if len(self.new_sequences) == 0:
            return
        self._update_leaf()

    def updated_alignment(self):
        """The leaf's alignment with its new sequences added by the builder's aligner (:362-368)."""
        assert self.prg_builder.aligner is not None, "Cannot make updates without a Multiple Sequence Aligner."
        return self.prg_builder.aligner.get_updated_alignment(current_alignment=self.alignment,
                                                              new_sequences=self.new_sequences)

    def swap_in(self, updated_child):
        """Puts the re-built node in this leaf's place and invalidates the builder's PRG index (:378-388)."""
        if self.is_root():
            self.prg_builder.replace_root(updated_child)
        else:
            self.parent.replace_child(self, updated_child)
        self.prg_builder.clear_PRG_index()

    def _update_leaf(self):
        """Update this leaf, replacing it by an updated node, which can be of a different subclass (:353-388)."""
        updated_child = NodeFactory.build(self.updated_alignment(), self.prg_builder, self.parent)
        self.swap_in(updated_child)

    def clear_PRG_interval_index(self):
        self.indexed_PRG_intervals.clear()


def batch_update_leaves(leaves, ctx=None):
    """LeafNode.batch_update for MANY leaves (of any number of loci) with ONE device batch per (max_nesting,
    min_match_length): every updated alignment is re-built below its leaf's parent in the same level-synchronous
    launches (mprg_build_sub), then swapped in -- the batched form of `make_prg update`'s re-partitioning
    (recursion_tree.py:345-388; leaves are visited in the given order, so node ids are handed out as a loop of
    _update_leaf calls would hand them out).  Returns the number of leaves updated."""
    todo = [leaf for leaf in leaves if len(leaf.new_sequences) > 0]
    groups = {}
    for leaf in todo:
        b = leaf.prg_builder
        groups.setdefault((b.max_nesting, b.min_match_length), []).append(leaf)
    for (max_nesting, mml), group in groups.items():
        alignments = [leaf.updated_alignment() for leaf in group]
        levels = [-1 if leaf.parent is None else leaf.parent.nesting_level for leaf in group]
        results = engine.build_matrices([a.matrix for a in alignments], max_nesting, mml, ctx=ctx,
                                        parent_levels=levels)
        for leaf, alignment, result in zip(group, alignments, results):
            builder = leaf.prg_builder
            result.raise_for_status(builder.locus_name)
            first = builder.next_node_id
            builder.next_node_id += result.n_nodes
            leaf.swap_in(nodes_from_table(alignment, result.nodes, builder, leaf.parent, first))
    return len(todo)


_CLASSES = {NODE_LEAF: LeafNode, NODE_INTERVAL: MultiIntervalNode, NODE_CLUSTER: MultiClusterNode}


def nodes_from_table(alignment: MSA, table, prg_builder, parent_node=None, first_node_id=0) -> RecursiveTreeNode:
    """Materialise the tree of one locus from the engine's pre-order node table.  With parent_node the
    tree hangs below that node and its ids count on from first_node_id."""
    n = len(table["kind"])
    root = RootAlignment(alignment)
    built: List[Optional[RecursiveTreeNode]] = [None] * n
    for i in range(n):
        if table["row_off"][i] < 0:
            rows = None
        else:
            o = int(table["row_off"][i])
            rows = np.array(table["row_pool"][o:o + int(table["n_rows"][i])], np.int32)
        sub = SubAlignment(root, rows, table["c0"][i], table["c1"][i])
        parent = built[int(table["parent"][i])] if table["parent"][i] >= 0 else parent_node
        node = _CLASSES[int(table["kind"][i])](int(table["nesting_level"][i]), sub, parent, prg_builder,
                                               first_node_id + i)
        built[i] = node
        if parent is not None and table["parent"][i] >= 0:
            parent._children.append(node)
    return built[0]


class NodeFactory:
    """NodeFactory.build(alignment, prg_builder, parent_node=None) (recursion_tree.py:401-471)."""

    @staticmethod
    def build(alignment, prg_builder, parent_node=None):
        if parent_node is not None:
            # LeafNode._update_leaf (recursion_tree.py:373-376): the updated alignment is re-built below
            # the leaf's parent; the caller swaps the new node in (replace_child) and re-emits the PRG
            result = engine.build_matrices([alignment.matrix], prg_builder.max_nesting,
                                           prg_builder.min_match_length,
                                           parent_levels=[parent_node.nesting_level])[0]
            result.raise_for_status(prg_builder.locus_name)
            first = prg_builder.next_node_id
            prg_builder.next_node_id += result.n_nodes
            return nodes_from_table(alignment, result.nodes, prg_builder, parent_node, first)
        result = engine.build_matrices([alignment.matrix], prg_builder.max_nesting,
                                       prg_builder.min_match_length)[0]
        result.raise_for_status(prg_builder.locus_name)
        prg_builder.next_node_id = result.n_nodes
        prg_builder.engine_prg = result.prg
        return nodes_from_table(alignment, result.nodes, prg_builder)

    @staticmethod
    def _get_vertical_partition(alignment, min_match_length):
        from .from_msa.interval_partition import IntervalPartitioner
        from .utils.seq_utils import get_consensus_from_MSA

        consensus = get_consensus_from_MSA(alignment)
        match, _non_match, all_intervals = IntervalPartitioner(consensus, min_match_length,
                                                               alignment).get_intervals()
        return all_intervals, match
