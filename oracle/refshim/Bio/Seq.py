class Seq:
    """String-like sequence (surface used by make_prg: str, len, iter, index/slice, upper, ==)."""

    def __init__(self, data):
        self._data = str(data)

    def __str__(self):
        return self._data

    def __repr__(self):
        return f"Seq({self._data!r})"

    def __len__(self):
        return len(self._data)

    def __iter__(self):
        return iter(self._data)

    def __getitem__(self, index):
        piece = self._data[index]
        return Seq(piece) if isinstance(index, slice) else piece

    def __eq__(self, other):
        return str(self) == str(other)

    def __hash__(self):
        return hash(self._data)

    def upper(self):
        return Seq(self._data.upper())

    def replace(self, old, new):
        return Seq(self._data.replace(old, new))
