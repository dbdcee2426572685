"""-O letters p/b/g/a (make_prg/subcommands/output_type.py)."""


class UnknownOutputTypeError(Exception):
    pass


class OutputType:
    BINARY, ALL, PRG, GFA = "b", "a", "p", "g"

    def __init__(self, value):
        self.type = set(value.lower())
        if not (self._all() or self.prg or self.binary or self.gfa):
            raise UnknownOutputTypeError(f"{value} is an unknown output type")

    def _all(self):
        return self.ALL in self.type

    @property
    def prg(self):
        return self._all() or self.PRG in self.type

    @property
    def binary(self):
        return self._all() or self.BINARY in self.type

    @property
    def gfa(self):
        return self._all() or self.GFA in self.type
