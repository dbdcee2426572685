"""PrgBuilder with the reference's constructor, fields and writers (make_prg/prg_builder.py:19-166)."""
import pickle
from pathlib import Path

from .recursion_tree import LeafNode, NodeFactory, RecursiveTreeNode, nodes_from_table
from .utils.io_utils import load_alignment_file
from .utils.prg_encoder import PrgEncoder


class LeafNotFoundException(Exception):
    pass


class PrgBuilder:
    def __init__(self, locus_name, msa_file, alignment_format, max_nesting, min_match_length,
                 aligner=None):
        self._locus_name = locus_name
        self.max_nesting = max_nesting
        self.min_match_length = min_match_length
        self.aligner = aligner
        self.next_node_id = 0
        self.site_num = 5
        self.prg_index = {}
        self.engine_prg = None
        alignment = load_alignment_file(str(msa_file), alignment_format)
        self.root: RecursiveTreeNode = NodeFactory.build(alignment, self, None)

    @classmethod
    def from_engine(cls, locus_name, alignment, locus_build, max_nesting, min_match_length):
        """Wrap the result of a batched engine run (no second pass over the device)."""
        self = cls.__new__(cls)
        self._locus_name = locus_name
        self.max_nesting = max_nesting
        self.min_match_length = min_match_length
        self.aligner = None
        self.next_node_id = locus_build.n_nodes
        self.site_num = 5
        self.prg_index = {}
        self.engine_prg = locus_build.prg
        self.root = nodes_from_table(alignment, locus_build.nodes, self)
        return self

    @property
    def locus_name(self):
        return self._locus_name

    def __getstate__(self):
        state = self.__dict__.copy()
        state["aligner"] = None
        return state

    def __eq__(self, other):
        if (self.locus_name, self.max_nesting, self.min_match_length, self.next_node_id, self.site_num) != (
                other.locus_name, other.max_nesting, other.min_match_length, other.next_node_id,
                other.site_num):
            return False
        return self.prg_index == other.prg_index and self.root == other.root

    def __hash__(self):
        return hash(self.locus_name)

    def replace_root(self, new_root):
        self.root = new_root

    def build_prg(self):
        self.site_num = 5
        prg_as_list = []
        self.root.preorder_traversal_to_build_prg(prg_as_list)
        return "".join(prg_as_list)

    def get_next_site_num(self):
        site_num = self.site_num
        self.site_num += 2
        return site_num

    def get_next_node_id(self):
        self.next_node_id += 1
        return self.next_node_id - 1

    def update_PRG_index(self, start_index, end_index, node: LeafNode):
        interval = (start_index, end_index)
        self.prg_index[interval] = node
        node.add_indexed_PRG_interval(interval)

    def clear_PRG_index(self):
        for node in self.prg_index.values():
            node.clear_PRG_interval_index()
        self.prg_index.clear()

    def get_node_given_interval(self, interval):
        if interval not in self.prg_index:
            raise LeafNotFoundException(
                f"Queried PRG interval {interval} does not exist in PRG index for locus {self.locus_name}.")
        return self.prg_index[interval]

    def serialize(self, filepath):
        with open(filepath, "wb") as fh:
            pickle.dump(self, fh, protocol=4)

    @staticmethod
    def deserialize_from_bytes(array_of_bytes):
        return pickle.loads(array_of_bytes)

    @staticmethod
    def write_prg_as_text(output_prefix, prg_string):
        sample = Path(output_prefix).name
        with Path(output_prefix + ".prg.fa").open("w") as fh:
            print(f">{sample}\n{prg_string}", file=fh)

    @staticmethod
    def write_prg_as_binary(output_prefix, prg_string):
        encoder = PrgEncoder()
        with Path(output_prefix + ".bin").open("wb") as fh:
            encoder.write(encoder.encode(prg_string), fh)
