"""Observation / action space descriptors.

If `gym` (or `gymnasium`) is importable its space classes are used so that the environments
plug into existing code; otherwise these minimal stand-ins with the same attributes
(low, high, shape, dtype, n, spaces, contains, sample) are used.  The image has neither.
"""
import numpy as np

try:  # pragma: no cover - not available in the build image
    from gym import spaces as _sp
    Box, Discrete, Dict = _sp.Box, _sp.Discrete, _sp.Dict
    BACKEND = "gym"
except Exception:
    try:  # pragma: no cover
        from gymnasium import spaces as _sp
        Box, Discrete, Dict = _sp.Box, _sp.Discrete, _sp.Dict
        BACKEND = "gymnasium"
    except Exception:
        BACKEND = "builtin"

        class Space(object):
            def __init__(self, shape=None, dtype=None):
                self.shape = None if shape is None else tuple(shape)
                self.dtype = None if dtype is None else np.dtype(dtype)

        class Box(Space):
            """Closed box; `contains` is shape-equal and low <= x <= high (inclusive), like gym's."""

            def __init__(self, low, high, shape=None, dtype=np.float32):
                if shape is None:
                    low, high = np.asarray(low), np.asarray(high)
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

            def sample(self):
                lo = np.where(np.isfinite(self.low), self.low, -1.0)
                hi = np.where(np.isfinite(self.high), self.high, 1.0)
                return np.random.uniform(lo, hi).astype(self.dtype)

            def __repr__(self):
                return "Box%s" % (self.shape,)

        class Discrete(Space):
            def __init__(self, n):
                self.n = int(n)
                Space.__init__(self, (), np.int64)

            def contains(self, x):
                try:
                    return 0 <= int(x) < self.n and int(x) == x
                except Exception:
                    return False

            def sample(self):
                return int(np.random.randint(self.n))

            def __repr__(self):
                return "Discrete(%d)" % self.n

        class Dict(Space):
            def __init__(self, spaces):
                self.spaces = dict(spaces)
                Space.__init__(self, None, None)

            def contains(self, x):
                return isinstance(x, dict) and all(k in x and s.contains(x[k]) for k, s in self.spaces.items())

            def sample(self):
                return {k: s.sample() for k, s in self.spaces.items()}

            def __repr__(self):
                return "Dict(%s)" % ", ".join("%s:%r" % kv for kv in self.spaces.items())
