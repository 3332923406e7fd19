"""Host-side reader of the reference's calibration artefact (quantizers.pickle): fixtures produced by the reference's
own SimQuant.quantize (tests/golden/gen_quantizers_golden.py) are parsed and turned into the per-channel K LUT the
oracle builds (oracle = restatement of QuantK.load_lookup_table, modeling_llama.py:447-501)."""
import os
import pickle

import numpy as np
import pytest

from _util import GOLDEN, O
from kvquant_b200 import quantizers as Q


@pytest.mark.parametrize("bits,norm", [(4, False), (3, False), (4, True), (3, True)])
def test_parse_reference_quantizers(bits, norm):
    path = os.path.join(GOLDEN, "quantizers_b%d%s.pkl" % (bits, "_norm" if norm else ""))
    parsed = Q.load_quantizers(path, norm=norm)
    raw = pickle.load(open(path, "rb"))
    assert sorted(parsed) == [0, 1]                      # the '.lut' key is skipped (deployment/llama.py:187-188)
    for n, d in parsed.items():
        for side, name in (("k", "k_proj"), ("v", "v_proj")):
            e, r = d[side], raw["model.layers.%d.self_attn.%s" % (n, name)]
            assert e["bits"] == bits and e["centroids"].shape == (2 ** bits,)
            assert np.all(np.diff(e["centroids"]) >= 0)                       # sorted (modeling_llama.py:450)
            assert np.array_equal(np.sort(np.asarray(r[2][0]).ravel()), e["centroids"])
            assert e["upper"].shape == (512,) and np.array_equal(e["upper"], np.asarray(r[0]).ravel())
            assert np.all(e["upper"] > e["lower"])
            if norm:
                assert e["normscale"] == pytest.approx(float(r[3])) and e["normoffset"] == pytest.approx(float(r[4]))
            else:
                assert e["normscale"] is None
        # the K LUT built from the parsed entry == the oracle's restatement of load_lookup_table
        k = d["k"]
        lut = O.build_k_lut(k["upper"], k["lower"], k["centroids"],
                            *((k["normscale"], k["normoffset"]) if norm else ()))
        assert lut["lut"].shape[-1] == 2 ** bits
        up16 = k["upper"].astype(np.float16).astype(np.float32)
        lo16 = k["lower"].astype(np.float16).astype(np.float32)
        assert np.array_equal(lut["thr_upper"].ravel(), up16) and np.array_equal(lut["thr_lower"].ravel(), lo16)


def test_parse_rejects_malformed_artefacts():
    good = pickle.load(open(os.path.join(GOLDEN, "quantizers_b4.pkl"), "rb"))
    with pytest.raises(KeyError):
        Q.parse_quantizers({"model.layers.0.mlp.up_proj": good["model.layers.0.self_attn.k_proj"]})
    only_k = {k: v for k, v in good.items() if "v_proj" not in k}
    with pytest.raises(KeyError):
        Q.parse_quantizers(only_k)
    with pytest.raises(ValueError):   # uniform-quant artefact (no centroids): the deployment kernels are NUQ-only
        Q.parse_quantizers({"model.layers.0.self_attn.k_proj": good["model.layers.0.self_attn.k_proj"][:2],
                            "model.layers.0.self_attn.v_proj": good["model.layers.0.self_attn.v_proj"][:2]})
    with pytest.raises(ValueError):   # Q-Norm requested, artefact calibrated without --norm
        Q.parse_quantizers(good, norm=True)


@pytest.mark.gpu
@pytest.mark.parametrize("bits", [4, 3])
def test_caches_built_from_artefact_match_oracle(bits):
    """quantizers.pickle -> native LayerCache (hidden 512 = 4 heads): fused appends are bit-exact with the oracle cache
    built from the same artefact, and the fused attend agrees within 1e-3."""
    import torch
    from _util import rel_err
    from kvquant_b200 import synth
    parsed = Q.load_quantizers(os.path.join(GOLDEN, "quantizers_b%d.pkl" % bits))
    L, Lmax, H = 70, 128, 4
    caches = Q.layer_caches_from_quantizers(parsed, Lmax, device="cuda:0")
    assert sorted(caches) == [0, 1]
    sp = synth.SynthSpec(H, 128, seed=3)
    k, v = sp.k_tokens(L, seed=21), sp.v_tokens(L, seed=22)
    for n, lc in caches.items():
        e = parsed[n]
        oc = O.OracleCache(bits, H, Lmax, O.build_k_lut(e["k"]["upper"], e["k"]["lower"], e["k"]["centroids"]),
                           e["v"]["centroids"], include_sparse=True)
        for t in range(L):
            oc.append(k[t], v[t])
            lc.append(torch.from_numpy(k[t]).cuda(), torch.from_numpy(v[t]).cuda())
        torch.cuda.synchronize()
        W = 128 * bits // 32
        assert np.array_equal(lc.kcache.cpu().numpy().reshape(H * W, Lmax)[:, :L], oc.kwords.reshape(H * W, Lmax)[:, :L])
        assert np.array_equal(lc.vcache.cpu().numpy().reshape(H * W, Lmax)[:, :L], oc.vwords.reshape(H * W, Lmax)[:, :L])
        assert np.array_equal(lc.k_outlier_idx.cpu().numpy()[:L], oc.k_idx[:L])
        q = O.rope_rotate_q(sp.q_vec(5), L, 10000.0)
        s = oc.k_scores(q, 10000.0, 0)
        _, want = O.attend_ideal(s, lambda p: oc.v_output(p))
        got = lc.attend(torch.from_numpy(q).cuda()).cpu().numpy()
        assert rel_err(got, want)[0] < 1e-3
