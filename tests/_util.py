"""Shared helpers for the tests: synthetic calibration + an oracle-built quantised cache."""
import functools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from kvquant_b200 import synth  # noqa: E402
from oracle import kvq_oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


@functools.lru_cache(maxsize=None)
def spec(H=32, seed=0):
    return synth.SynthSpec(H, 128, seed=seed)


@functools.lru_cache(maxsize=None)
def quantizer(bits, H=32, seed=0):
    sp = spec(H, seed)
    cal = synth.calibrate(sp, bits, calib_tokens=512, seed=7)
    up, lo, kc = cal["k"]
    klut = O.build_k_lut(up, lo, kc[0])
    vcent = np.sort(cal["v"][2][0].ravel().astype(np.float32))
    return klut, vcent


@functools.lru_cache(maxsize=None)
def oracle_cache(bits, L, H=32, sparse=True, Lmax=None, seed=0):
    """Token-by-token oracle cache (reference decode semantics) of L tokens."""
    sp = spec(H, seed)
    klut, vcent = quantizer(bits, H, seed)
    Lmax = Lmax or ((L + 63) // 64 * 64 + 64)
    c = O.OracleCache(bits, H, Lmax, klut, vcent, include_sparse=sparse)
    k = sp.k_tokens(L, seed=11)
    v = sp.v_tokens(L, seed=12)
    for t in range(L):
        c.append(k[t], v[t])
    return c, k, v


def rel_err(a, b):
    """max |a-b| / max |b|  (norm-wise relative error) and rel-L2."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.abs(a - b)
    return float(d.max() / max(np.abs(b).max(), 1e-30)), float(np.linalg.norm(d) / max(np.linalg.norm(b), 1e-30))
