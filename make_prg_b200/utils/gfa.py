"""PRG string -> GFA1 (make_prg/utils/gfa.py:4-109).  Same segment / link numbering as the reference's
recursive split (ids in order of left-to-right site expansion), produced from one tokenisation of the
PRG instead of one regex split per site."""
import re

HEADER = "H\tVN:Z:1.0\tbn:Z:--linear --singlearr\n"
_TOKEN = re.compile(r" (\d+) |([^ ]+)")


def _parse(prg_string):
    """-> nested structure: a 'sequence' is a list of items, an item is a literal str or a site
    (list of allele sequences)."""
    root = []
    stack = [(None, root, None)]  # (site marker, current allele sequence, alleles list)
    for m in _TOKEN.finditer(prg_string):
        marker, literal = m.group(1), m.group(2)
        if literal is not None:
            stack[-1][1].append(literal)
            continue
        marker = int(marker)
        top_marker, _cur, alleles = stack[-1]
        if marker % 2 == 1:
            if top_marker == marker:  # closes the site
                stack.pop()
            else:  # opens a site
                alleles = [[]]
                stack[-1][1].append(alleles)
                stack.append((marker, alleles[0], alleles))
        else:  # allele separator of the innermost open site
            assert top_marker is not None and marker == top_marker + 1, "Invalid prg sequence"
            alleles.append([])
            stack[-1] = (top_marker, alleles[-1], alleles)
    assert len(stack) == 1, "Invalid prg sequence"
    return root


class GFA_Output:
    def __init__(self, gfa_string="", gfa_id=0, gfa_site=5):
        self.gfa_string = gfa_string
        self.gfa_id = gfa_id
        self.gfa_site = gfa_site
        self.delim_char = " "
        self._lines = []

    def _segment(self, text):
        self._lines.append("S\t%d\t%s\tRC:i:0\n" % (self.gfa_id, text if text != "" else "*"))

    def _link(self, a, b):
        self._lines.append("L\t%d\t+\t%d\t+\t0M\n" % (a, b))

    def _emit(self, sequence):
        """One (sub)string of the PRG: literal? (site literal?)* -- returns the ids that end it."""
        end_ids = []
        pending = ""
        for item in sequence:
            if isinstance(item, str):
                pending += item
                continue
            self._segment(pending)
            pre_var_id = self.gfa_id
            self.gfa_id += 1
            for e in end_ids:
                self._link(e, pre_var_id)
            end_ids = []
            assert len(item) > 1, "Invalid prg sequence"
            for allele in item:
                self._link(pre_var_id, self.gfa_id)
                end_ids.extend(self._emit(allele))
            pending = ""
        self._segment(pending)
        for e in end_ids:
            self._link(e, self.gfa_id)
        last = self.gfa_id
        self.gfa_id += 1
        return [last]

    def build_gfa_string(self, prg_string, pre_var_id=None):
        self._lines = []
        ids = self._emit(_parse(prg_string))
        self.gfa_string += "".join(self._lines)
        return ids

    @staticmethod
    def write_gfa(prefix, prg_string):
        gfa_obj = GFA_Output(HEADER)
        gfa_obj.build_gfa_string(prg_string=prg_string)
        with open(f"{prefix}.gfa", "w") as fh:
            fh.write(gfa_obj.gfa_string)
