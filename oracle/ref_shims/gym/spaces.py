class Discrete:
    def __init__(self, n):
        self.n = n


class Box:
    def __init__(self, low=None, high=None, shape=None, dtype=None):
        self.low, self.high, self.shape, self.dtype = low, high, shape, dtype


class Dict(dict):
    pass
