class _CM:
    @staticmethod
    def get_cmap(name):
        return lambda v: (float(v), 0.0, 1.0 - float(v), 1.0)


cm = _CM()
