class _Align:
    def globalms(self, *args, **kwargs):
        raise NotImplementedError("stand-in: the update path is out of scope")


align = _Align()
