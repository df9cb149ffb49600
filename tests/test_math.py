"""The shared bit-reproducible math (gca_math.h) and Philox, through the oracle library."""
import ctypes as C
import math

import numpy as np

from oracle import oracle as orc


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    L = orc.lib()
    cases = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
             ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
             ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
              (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in cases:
        c = np.array(ctr, np.uint32); k = np.array(key, np.uint32); out = np.zeros(4, np.uint32)
        L.gca_oracle_philox4x32_10(c.ctypes.data, k.ctypes.data, out.ctypes.data)
        assert tuple(int(x) for x in out) == want


def test_sincos_log_accuracy_vs_mpmath():
    import mpmath as mp
    mp.mp.prec = 200
    L = orc.lib()
    rng = np.random.RandomState(0)
    xs = np.concatenate([rng.uniform(-10, 10, 3000), rng.uniform(-1e3, 1e3, 1500), rng.uniform(-1e6, 1e6, 1500)])
    worst = 0.0
    same = 0
    for x in xs:
        s, c = C.c_double(), C.c_double()
        L.gca_oracle_sincos(float(x), orc.TRIG_SHARED, C.byref(s), C.byref(c))
        for got, true in ((s.value, mp.sin(x)), (c.value, mp.cos(x))):
            worst = max(worst, float(abs(mp.mpf(got) - true) / mp.mpf(np.spacing(abs(float(true))))))
        same += (s.value == math.sin(x)) + (c.value == math.cos(x))   # libm, what the reference's math.cos/sin call
    assert worst < 1.0, worst                        # < 1 ulp on |x| < 1e6
    assert same > 0.95 * 2 * len(xs)                 # equals libm except where one of the two is not correctly rounded
    us = np.concatenate([rng.uniform(0, 1, 3000), rng.uniform(0, 1, 2000) * 2.0 ** -rng.randint(0, 52, 2000)])
    worst = 0.0
    for u in us[us > 0]:
        got = L.gca_oracle_log(float(u), orc.TRIG_SHARED)
        true = mp.log(u)
        if true != 0:
            worst = max(worst, float(abs(mp.mpf(got) - true) / mp.mpf(np.spacing(abs(float(true))))))
    assert worst < 1.0, worst


def test_philox_normals_are_standard_normal():
    L = orc.lib()
    g = np.zeros(2)
    vals = []
    for i in range(20000):
        L.gca_oracle_philox_normal2(12345, i, 0, 0x80000000, orc.TRIG_SHARED, g.ctypes.data)
        vals.extend(g.tolist())
    v = np.array(vals)
    assert abs(v.mean()) < 0.02 and abs(v.std() - 1) < 0.02
    assert abs(((v[0::2] * v[1::2]).mean())) < 0.02                       # the pair is uncorrelated
    assert abs((np.abs(v) < 1).mean() - 0.6827) < 0.01
