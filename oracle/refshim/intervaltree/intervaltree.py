class _Iv:
    def __init__(self, begin, end, data):
        self.begin, self.end, self.data = begin, end, data


class IntervalTree:
    """Linear-scan stub; only the update sub-command (out of scope) uses it."""

    def __init__(self):
        self._ivs = []

    def addi(self, begin, end, data=None):
        self._ivs.append((begin, end, data))

    def __getitem__(self, point):
        return {_Iv(b, e, d) for (b, e, d) in self._ivs if b <= point < e}
