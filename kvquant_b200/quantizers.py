"""Reader for the reference's calibration artefact (`quantizers.pickle`) and the glue that turns it into caches.

The reference writes one entry per projection, `quantizers['model.layers.<n>.self_attn.<k_proj|v_proj>']`
(quant/llama_simquant.py:275), each the return value of `SimQuant.quantize`
(quant/kvquant/simquant_module_quantizer.py:550-552):

    (outlier_threshold_upper [1, hidden], outlier_threshold_lower [1, hidden], [centroids (2^bits, 1)]
     [, normscale, normoffset])                                            # the last two only with Q-Norm

and consumes it in `deployment/llama.py:179-198` -> `QuantK.load_lookup_table` / `QuantV.load_lookup_table`
(modeling_llama.py:437-501, 1045-1066): keys containing '.lut' are skipped, the layer number is the third field from
the end of the key, K uses thresholds + centroids (per-channel LUT), V uses the centroids only (its thresholds are
per token, found at run time).

`parse_quantizers` is pure host code (numpy); `layer_caches_from_quantizers` builds the native `LayerCache`s and needs
the CUDA library.  The file is a pickle: as with the reference, only load artefacts you trust.
"""
from __future__ import annotations

import pickle
import re

import numpy as np

_KEY = re.compile(r"(?:^|\.)layers\.(\d+)\.self_attn\.(k_proj|v_proj)$")


def _np(x, dtype=np.float32):
    """torch tensor / numpy array / nested list -> numpy (without importing torch unless the object is a tensor)."""
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=dtype)


def _scalar(x):
    return float(_np(x, np.float64).reshape(-1)[0])


def parse_entry(entry, norm=False):
    """One pickle entry -> dict(upper f32[hidden], lower f32[hidden], centroids f32[2^bits] sorted ascending,
    normscale, normoffset).  Mirrors modeling_llama.py:447-450 / 1054-1055 (flatten, take centroid list [0], squeeze,
    sort) and 466-467 (Q-Norm parameters)."""
    if not isinstance(entry, (tuple, list)) or len(entry) < 3:
        raise ValueError("quantizer entry must be (upper, lower, [centroids][, normscale, normoffset]); "
                         "calibrate with --nuq (the deployment kernels are NUQ-only)")
    upper = _np(entry[0]).reshape(-1)
    lower = _np(entry[1]).reshape(-1)
    if upper.shape != lower.shape:
        raise ValueError("upper / lower threshold shapes differ: %s vs %s" % (upper.shape, lower.shape))
    cent = np.sort(_np(entry[2][0]).reshape(-1))
    n = cent.size
    if n not in (4, 8, 16):
        raise ValueError("expected 4, 8 or 16 centroids (2/3/4-bit NUQ), got %d" % n)
    out = dict(upper=upper, lower=lower, centroids=cent, bits=int(np.log2(n)), normscale=None, normoffset=None)
    if norm:
        if len(entry) < 5:
            raise ValueError("Q-Norm requested but the entry has no (normscale, normoffset): calibrate with --norm")
        out["normscale"], out["normoffset"] = _scalar(entry[3]), _scalar(entry[4])
    return out


def parse_quantizers(quantizers, norm=False):
    """The whole artefact -> {layer index: {'k': entry, 'v': entry}} (entries as in `parse_entry`).
    Keys containing '.lut' are skipped as in deployment/llama.py:187-188; other non-matching keys raise."""
    layers = {}
    for key, entry in quantizers.items():
        if ".lut" in key:
            continue
        m = _KEY.search(key)
        if m is None:
            raise KeyError("unexpected quantizer key %r (want ...layers.<n>.self_attn.k_proj / v_proj)" % key)
        layers.setdefault(int(m.group(1)), {})["k" if m.group(2) == "k_proj" else "v"] = parse_entry(entry, norm)
    for n, d in layers.items():
        if "k" not in d or "v" not in d:
            raise KeyError("layer %d lacks its %s entry" % (n, "k_proj" if "k" not in d else "v_proj"))
        if d["k"]["bits"] != d["v"]["bits"]:
            raise ValueError("layer %d: K is %d-bit, V is %d-bit" % (n, d["k"]["bits"], d["v"]["bits"]))
    return dict(sorted(layers.items()))


def load_quantizers(path, norm=False):
    """Read `quantizers.pickle` (deployment/llama.py:179-181) and parse it."""
    with open(path, "rb") as f:
        return parse_quantizers(pickle.load(f), norm=norm)


def layer_caches_from_quantizers(parsed, max_len, device="cuda", include_sparse=True, sparsity_threshold=0.99,
                                 n_sink=0, layers=None):
    """Native caches for the given layers (default: all), LUTs built exactly as QuantK.load_lookup_table does
    (fp16-rounded thresholds, modeling_llama.py:447-501).  hidden = number of thresholds; heads = hidden / 128."""
    from .cache import HEAD_DIM, LayerCache, build_k_lookup_table
    out = {}
    for n in (parsed.keys() if layers is None else layers):
        k, v = parsed[n]["k"], parsed[n]["v"]
        hidden = k["upper"].size
        if hidden % HEAD_DIM:
            raise ValueError("hidden size %d is not a multiple of %d" % (hidden, HEAD_DIM))
        H = hidden // HEAD_DIM
        t = build_k_lookup_table(k["upper"], k["lower"], k["centroids"], H, k["normscale"], k["normoffset"], device=device)
        v_norm = (v["normscale"], v["normoffset"]) if v["normscale"] is not None else None
        out[n] = LayerCache.from_luts(k["bits"], H, max_len, t, v["centroids"], device=device,
                                      include_sparse=include_sparse, sparsity_threshold=sparsity_threshold,
                                      n_sink=n_sink, v_norm=v_norm)
    return out
