"""PrgBuilder with the reference's constructor, fields and writers (make_prg/prg_builder.py:19-166), and
PrgBuilderZipDatabase (:169-221).  The update_DS archive holds one member per locus: either a pickle (what
`serialize` writes, as in the reference) or the table-shaped record the native writer produces
(mprg_writer_add_ds, "MPRGDS01": node table + row subsets + titles + packed root alignment + PRG), from which
the objects are built on load."""
import pickle
import struct
from pathlib import Path
from zipfile import ZipFile

import numpy as np

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
    def deserialize_from_bytes(array_of_bytes, locus_name=None):
        if bytes(array_of_bytes[:8]) == DS_MAGIC:
            return builder_from_ds_record(array_of_bytes, locus_name)
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


DS_MAGIC = b"MPRGDS01"


def parse_ds_record(blob):
    """The tables of one update_DS member written by mprg_writer_add_ds (layout: csrc/hostio.cpp)."""
    from . import hostio

    if bytes(blob[:8]) != DS_MAGIC:
        raise ValueError("not a table-shaped update_DS record")
    max_nesting, mml, rows, cols, n_nodes, n_sites, stride, _ = struct.unpack_from("<8i", blob, 8)
    pool_len, titles_len, prg_len = struct.unpack_from("<3q", blob, 40)
    at = 64
    buf = np.frombuffer(blob, np.uint8)
    table = {}
    for key in ("kind", "parent", "nesting_level", "c0", "c1", "n_rows", "n_children"):
        table[key] = buf[at:at + 4 * n_nodes].view("<i4")
        at += 4 * n_nodes
    table["row_off"] = buf[at:at + 8 * n_nodes].view("<i8")
    at += 8 * n_nodes
    table["row_pool"] = buf[at:at + 4 * pool_len].view("<i4")
    at += 4 * pool_len
    titles = bytes(buf[at:at + titles_len]).decode()
    at += titles_len
    packed = buf[at:at + rows * stride].reshape(rows, stride) if rows else np.zeros((0, stride), np.uint8)
    at += rows * stride
    prg = bytes(buf[at:at + prg_len]).decode()
    titles = [t.rstrip() for t in titles.split("\n")] if rows else []
    titles += [""] * (rows - len(titles))
    return {"max_nesting": max_nesting, "min_match_length": mml, "n_nodes": n_nodes, "n_sites": n_sites,
            "table": table, "titles": titles, "matrix": hostio.unpack_rows(packed, cols), "prg": prg}


def builder_from_ds_record(blob, locus_name):
    """A PrgBuilder (tree objects, PRG index) from a table-shaped record."""
    from . import engine
    from .msa import MSA

    rec = parse_ds_record(blob)
    ids = []
    for title in rec["titles"]:
        tokens = title.split(None, 1)
        ids.append(tokens[0] if tokens else "")
    alignment = MSA.from_matrix(ids, rec["matrix"], descriptions=rec["titles"])
    build = engine.LocusBuild(0, rec["prg"], rec["n_nodes"], rec["n_sites"], rec["table"])
    builder = PrgBuilder.from_engine(locus_name, alignment, build, rec["max_nesting"], rec["min_match_length"])
    prg = builder.build_prg()  # fills prg_index and the leaves' indexed_PRG_intervals, as the reference's pickles hold them
    if prg != rec["prg"]:
        raise RuntimeError(f"PRG emission mismatch for {locus_name}")
    return builder


class PrgBuilderZipDatabase:
    """A collection of PrgBuilders saved to / loaded from a zip file (make_prg/prg_builder.py:169-221)."""

    def __init__(self, zip_filepath):
        zip_filepath = Path(zip_filepath)
        assert zip_filepath.suffix == ".zip", "PrgBuilderZipDatabase initialised without a .zip filepath"
        self._zip_filepath = zip_filepath
        self._zip_file = None

    def save(self, locus_to_prg_builder_pickle_path):
        from zipfile import ZIP_STORED

        with ZipFile(self._zip_filepath, "w", ZIP_STORED) as zf:
            for locus, path in locus_to_prg_builder_pickle_path.items():
                zf.write(path, arcname=locus)

    def load(self):
        self._zip_file = ZipFile(self._zip_filepath)

    def close(self):
        if self._zip_file is not None:
            self._zip_file.close()

    def get_number_of_loci(self):
        return len(self.get_loci_names())

    def get_loci_names(self):
        return sorted(self._zip_file.namelist())

    def get_PrgBuilder(self, locus):
        return PrgBuilder.deserialize_from_bytes(self._zip_file.read(locus), locus)

    def __eq__(self, other):
        if self.get_loci_names() != other.get_loci_names():
            return False
        return all(self.get_PrgBuilder(locus) == other.get_PrgBuilder(locus) for locus in self.get_loci_names())
