"""CPU: the C port (oracle/kvq_oracle_port.c, the cpu_baseline of bench.py) against the numpy oracle."""
import os
import sys

import numpy as np
import pytest

from _util import O, ROOT, oracle_cache, rel_err, spec

sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _p(a):
    return a.ctypes.data


@pytest.mark.parametrize("bits,sparse", [(4, True), (3, True), (2, False)])
def test_c_port_matches_numpy_oracle(bits, sparse):
    import build_oracle_c
    lib = build_oracle_c.load()
    L = 70
    c, k, v = oracle_cache(bits, L, sparse=sparse)
    q = np.ascontiguousarray(O.rope_rotate_q(spec().q_vec(2), L, 10000.0))
    klut = np.ascontiguousarray(c.klut["lut"])
    scores = np.zeros((32, L), np.float32)
    ko, ki = (c.k_out, c.k_idx) if sparse else (None, None)
    lib.kvq_port_k_scores(bits, _p(q), _p(c.kwords), _p(klut), _p(ko) if sparse else None,
                          _p(ki) if sparse else None, c.k_out.shape[1], 32, c.Lmax, L, 10000.0, 0, _p(scores))
    assert rel_err(scores, c.k_scores(q))[0] < 1e-4
    out = np.zeros((32, 128), np.float32)
    scratch = np.zeros((32, L), np.float32)
    lib.kvq_port_attend(bits, _p(q), _p(c.kwords), _p(klut), _p(ko) if sparse else None, _p(ki) if sparse else None,
                        _p(c.vwords), _p(c.vlut), _p(c.v_out) if sparse else None, _p(c.v_idx) if sparse else None,
                        c.k_out.shape[1], 32, c.Lmax, L, 10000.0, 0, _p(out), _p(scratch))
    _, want = O.attend_ideal(c.k_scores(q), c.v_output)
    assert rel_err(out, want)[0] < 1e-4


def test_c_port_k_outliers_only_matches_numpy_oracle():
    """BASELINE configs[4]'s cache form (capped K outliers, dense-only V): the port takes NULL V outlier rows."""
    import build_oracle_c
    from _util import quantizer
    lib = build_oracle_c.load()
    bits, H, L = 3, 32, 50
    klut_d, vcent = quantizer(bits)
    sp = spec()
    c = O.OracleCache(bits, H, 128, klut_d, vcent, include_sparse=True, sparse_v=False)
    k, v = sp.k_tokens(L, 51), sp.v_tokens(L, 52)
    for t in range(L):
        c.append(k[t], v[t])
    q = np.ascontiguousarray(O.rope_rotate_q(sp.q_vec(3), L, 10000.0))
    klut = np.ascontiguousarray(c.klut["lut"])
    out = np.zeros((H, 128), np.float32)
    scratch = np.zeros((H, L), np.float32)
    lib.kvq_port_attend(bits, _p(q), _p(c.kwords), _p(klut), _p(c.k_out), _p(c.k_idx), _p(c.vwords), _p(c.vlut), None, None,
                        c.k_out.shape[1], H, c.Lmax, L, 10000.0, 0, _p(out), _p(scratch))
    _, want = O.attend_ideal(c.k_scores(q), c.v_output)
    assert rel_err(out, want)[0] < 1e-4
