"""update_DS archive of this package -> the REFERENCE's update_DS archive (pickles of make_prg.prg_builder.PrgBuilder
objects holding Bio.Align.MultipleSeqAlignment sub-alignments, make_prg/prg_builder.py:145-147), for users who run
the reference's `make_prg update` on PRGs built here.

Needs the reference's own environment: `make_prg` and `Bio` must be importable (they are not dependencies of this
package and are not installed in the build image; tests/test_reference_export.py runs this under the Biopython
stand-in of oracle/refshim with the unmodified reference sources).  Every object is an instance of the reference's
own classes, built without running its constructors (which would re-run the whole recursion on the CPU):

    python -m make_prg_b200.utils.reference_export <prefix>.update_DS.zip <out>.update_DS.zip
"""
import pickle
import sys
from zipfile import ZIP_STORED, ZipFile


def _reference_classes():
    try:
        import make_prg.prg_builder as pb
        import make_prg.recursion_tree as rt
        from Bio.Align import MultipleSeqAlignment
        from Bio.Seq import Seq
        from Bio.SeqRecord import SeqRecord
    except ImportError as err:
        raise ImportError("reference_export needs the reference's environment (make_prg, Biopython): "
                          f"{err}") from err
    return pb, rt, MultipleSeqAlignment, Seq, SeqRecord


def to_reference_builder(builder):
    """make_prg_b200.prg_builder.PrgBuilder -> make_prg.prg_builder.PrgBuilder (same tree, ids, alignments, index)."""
    pb, rt, MultipleSeqAlignment, Seq, SeqRecord = _reference_classes()
    classes = {"LeafNode": rt.LeafNode, "MultiIntervalNode": rt.MultiIntervalNode, "MultiClusterNode": rt.MultiClusterNode}
    ref = object.__new__(pb.PrgBuilder)
    ref._locus_name = builder.locus_name
    ref.max_nesting = builder.max_nesting
    ref.min_match_length = builder.min_match_length
    ref.aligner = None
    ref.next_node_id = builder.next_node_id
    ref.site_num = 5
    ref.prg_index = {}

    def bio_msa(alignment):
        return MultipleSeqAlignment([SeqRecord(Seq(str(r.seq)), id=r.id, name=r.name, description=r.description)
                                     for r in alignment])

    def convert(node, parent):
        out = object.__new__(classes[type(node).__name__])
        out.nesting_level = node.nesting_level
        out.alignment = bio_msa(node.alignment)  # columns full of gaps already removed (recursion_tree.py:45)
        out.parent = parent
        out.prg_builder = ref
        out._node_id = node.node_id
        out._children = []
        if isinstance(out, rt.LeafNode):
            out.new_sequences = set()
            out.indexed_PRG_intervals = set()
        out._children = [convert(child, out) for child in node.children]
        return out

    ref.root = convert(builder.root, None)
    prg = ref.build_prg()  # the reference's own traversal fills its prg_index / indexed_PRG_intervals
    if prg != builder.build_prg():
        raise RuntimeError(f"the reference's traversal of the exported tree of {builder.locus_name} gives another PRG")
    return ref


def export_update_ds(src_zip, dst_zip):
    """Every locus of src_zip (table-shaped records or pickles of this package) as a reference pickle in dst_zip."""
    from ..prg_builder import PrgBuilderZipDatabase

    db = PrgBuilderZipDatabase(src_zip)
    db.load()
    n = 0
    with ZipFile(dst_zip, "w", ZIP_STORED) as out:
        for locus in db.get_loci_names():
            out.writestr(locus, pickle.dumps(to_reference_builder(db.get_PrgBuilder(locus)), protocol=4))
            n += 1
    db.close()
    return n


if __name__ == "__main__":
    print(export_update_ds(sys.argv[1], sys.argv[2]), "loci exported")
