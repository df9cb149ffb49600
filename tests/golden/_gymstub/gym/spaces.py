import numpy as np


class Space(object):
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)


class Box(Space):
    def __init__(self, low=None, high=None, shape=None, dtype=None):
        if dtype is None:
            dtype = np.float32
        if shape is None:
            low = np.asarray(low)
            high = np.asarray(high)
            shape = low.shape
        else:
            low = np.full(shape, low) if np.isscalar(low) else np.asarray(low)
            high = np.full(shape, high) if np.isscalar(high) else np.asarray(high)
        with np.errstate(all="ignore"):
            self.low = low.astype(dtype)
            self.high = high.astype(dtype)
        Space.__init__(self, shape, dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool((x >= self.low).all()) and bool((x <= self.high).all())


class Discrete(Space):
    def __init__(self, n):
        self.n = n
        Space.__init__(self, (), np.int64)

    def contains(self, x):
        return 0 <= int(x) < self.n


class Dict(Space):
    def __init__(self, spaces):
        self.spaces = spaces
        Space.__init__(self, None, None)
