"""Synthetic K/V activations and calibration artefacts (no weights / datasets are available offline).

Follows SURVEY.md section 8d: K is per-channel Gaussian with a few wide "outlier channels" per head plus an i.i.d.
heavy tail (the per-channel structure KVQuant targets, reference README.md:8); V is per-token Gaussian
plus the same tail.  Thresholds are the per-channel 0.5 / 99.5 percentiles of a calibration draw rounded
through fp16 (reference quant/kvquant/simquant_module_quantizer.py:465-466 and
deployment/.../modeling_llama.py:447-448); centroids are 1-D quantile-initialised Lloyd iterations on the
normalised non-outlier calibration values, 2^b entries, sorted.

The output mimics one entry of the reference's `quantizers.pickle`
(quant/llama_simquant.py:275-283): (upper_thr[hidden], lower_thr[hidden], [centroids(2^b,1)], ...).
"""
from __future__ import annotations

import numpy as np


class SynthSpec:
    def __init__(self, num_heads=32, head_dim=128, seed=0):
        self.H = num_heads
        self.D = head_dim
        self.hidden = num_heads * head_dim
        rng = np.random.default_rng(seed)
        self.mu = rng.normal(0.0, 0.5, self.hidden).astype(np.float32)
        self.sigma = np.exp(rng.normal(0.0, 0.5, self.hidden)).astype(np.float32)
        # 4 outlier channels per head, sigma x8
        for h in range(num_heads):
            ch = rng.choice(head_dim, 4, replace=False) + h * head_dim
            self.sigma[ch] *= 8.0
        self.seed = seed

    def k_tokens(self, T, seed):
        """K activations [T, hidden] fp32 (token-major)."""
        rng = np.random.default_rng([self.seed, 1, seed])
        x = rng.standard_normal((T, self.hidden), dtype=np.float32) * self.sigma + self.mu
        tail = rng.random((T, self.hidden)) < 0.005
        t3 = rng.standard_t(3, size=int(tail.sum())).astype(np.float32)
        x[tail] += t3 * 4.0 * self.sigma[np.nonzero(tail)[1]]
        return x.astype(np.float32)

    def v_tokens(self, T, seed):
        """V activations [T, hidden] fp32."""
        rng = np.random.default_rng([self.seed, 2, seed])
        st = np.exp(rng.normal(0.0, 0.3, (T, 1))).astype(np.float32)
        x = rng.standard_normal((T, self.hidden), dtype=np.float32) * st
        tail = rng.random((T, self.hidden)) < 0.005
        t3 = rng.standard_t(3, size=int(tail.sum())).astype(np.float32)
        x[tail] += t3 * 4.0 * np.broadcast_to(st, x.shape)[tail]
        return x.astype(np.float32)

    def q_vec(self, seed):
        """Query [H, D] fp32 (fp16-representable, as it comes out of an fp16 q_proj)."""
        rng = np.random.default_rng([self.seed, 3, seed])
        return rng.standard_normal((self.H, self.D)).astype(np.float16).astype(np.float32)


def lloyd_centroids(x, nlevels, iters=12):
    """1-D k-means (Lloyd) initialised at quantiles; returns sorted float32 centroids."""
    x = np.sort(np.asarray(x, dtype=np.float64).ravel())
    qs = (np.arange(nlevels) + 0.5) / nlevels
    c = np.quantile(x, qs)
    for _ in range(iters):
        edges = (c[1:] + c[:-1]) / 2
        b = np.searchsorted(edges, x)
        sums = np.bincount(b, weights=x, minlength=nlevels)
        cnt = np.bincount(b, minlength=nlevels)
        nz = cnt > 0
        c[nz] = sums[nz] / cnt[nz]
        c = np.sort(c)
    return c.astype(np.float32)


def calibrate(spec: SynthSpec, bits: int, calib_tokens=2048, seed=1234):
    """Returns a dict shaped like one layer's pair of `quantizers.pickle` entries:
    k: (upper[hidden], lower[hidden], [cent(2^b,1)]), v: (.., .., [cent(2^b,1)])."""
    n = 2 ** bits
    k = spec.k_tokens(calib_tokens, seed)
    up = np.quantile(k, 0.995, axis=0).astype(np.float32)
    lo = np.quantile(k, 0.005, axis=0).astype(np.float32)
    up16 = up.astype(np.float16).astype(np.float32)
    lo16 = lo.astype(np.float16).astype(np.float32)
    zp = (up16 + lo16) / 2
    rg = (up16 - lo16) / 2
    kn = (k - zp) / rg
    kcent = lloyd_centroids(kn[np.abs(kn) <= 1.0][:400000], n)
    v = spec.v_tokens(calib_tokens, seed + 1)
    vhi = np.quantile(v, 0.995, axis=1, keepdims=True)
    vlo = np.quantile(v, 0.005, axis=1, keepdims=True)
    vn = (v - (vhi + vlo) / 2) / ((vhi - vlo) / 2)
    vcent = lloyd_centroids(vn[np.abs(vn) <= 1.0][:400000], n)
    return {
        "k": (up, lo, [kcent.reshape(n, 1)]),
        "v": (np.quantile(v, 0.995, axis=0).astype(np.float32),
              np.quantile(v, 0.005, axis=0).astype(np.float32), [vcent.reshape(n, 1)]),
    }


# ---------------------------------------------------------------------------------------------------------------
# GPU-side synthetic cache fill (bench / probes): same distributions, generated with torch on the device
# ---------------------------------------------------------------------------------------------------------------
def fill_layer_cache_gpu(lc, spec: SynthSpec, L: int, seed: int = 0, chunk: int = 8192):
    """Fill a kvquant_b200.cache.LayerCache with L synthetic tokens through the real prefill packers
    (quant_cuda.vecquantNappendvec{K,V}sparseParallel + device top-k), `chunk` tokens at a time; the first chunk is
    packed for real, later chunks are packed into a scratch cache and copied to their slots."""
    import torch
    from . import cache as kc
    dev = lc.device
    H, hidden, bits = lc.H, lc.hidden, lc.bits
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + seed)
    mu = torch.as_tensor(spec.mu, device=dev)
    sigma = torch.as_tensor(spec.sigma, device=dev)
    qk = kc.QuantK(bits, hidden, H, max_position_embeddings=chunk, include_sparse=True)
    qv = kc.QuantV(bits, hidden, H, max_position_embeddings=chunk, include_sparse=True)
    qk.lookup_table = lc.klut.view(H, 128, -1)
    qk.lookup_table2 = None
    qk.outlier_threshold_lower, qk.outlier_threshold_upper = lc.thr_lower, lc.thr_upper
    qv.lut = lc.v_cent
    done = 0
    while done < L:
        T = min(chunk, L - done)
        k = torch.randn((T, hidden), generator=g, device=dev) * sigma + mu
        tail = torch.rand((T, hidden), generator=g, device=dev) < 0.005
        k = k + tail * torch.randn((T, hidden), generator=g, device=dev) * 6.0 * sigma
        st = torch.exp(torch.randn((T, 1), generator=g, device=dev) * 0.3)
        v = torch.randn((T, hidden), generator=g, device=dev) * st
        tail = torch.rand((T, hidden), generator=g, device=dev) < 0.005
        v = v + tail * torch.randn((T, hidden), generator=g, device=dev) * 6.0 * st
        qk.reset(); qv.reset()
        qk.parallel_pack(k.t().contiguous().view(H, 128, T))
        uv, ui, lv, li = qv.topk_thresholds(v)
        if not lc.sparse_v:
            # dense-only V (modeling_llama.py:1101-1108): the range is the min / max of the vector, nothing lies beyond it
            # -> the same packer with those thresholds produces exactly the dense-only codes and LUT rows
            uv = v.max(dim=-1, keepdim=True).values.expand(-1, uv.shape[1]).contiguous()
            lv = v.min(dim=-1, keepdim=True).values.expand(-1, lv.shape[1]).contiguous()
        qv.parallel_pack(v.t().contiguous().view(H, 128, T), uv, ui, lv, li)
        sl = slice(done, done + T)
        lc.kcache[:, :, sl] = qk.kcache[:, :, :T]
        lc.vcache[:, :, sl] = qv.vcache[:, :, :T]
        lc.vlut[sl] = qv.lookup_table[:T]
        lc.vaff[sl, 0] = (uv[:, -1] - lv[:, -1]) / 2
        lc.vaff[sl, 1] = (uv[:, -1] + lv[:, -1]) / 2
        if lc.sparse_k:   # (K codes do not depend on the outlier split: outliers are clamped to the LUT ends either way)
            lc.k_outliers[sl] = qk.outliers[:T]; lc.k_outlier_idx[sl] = qk.outlier_indices[:T]
        if lc.sparse_v:
            lc.v_outliers[sl] = qv.outliers[:T]; lc.v_outlier_idx[sl] = qv.outlier_indices[:T]
        done += T
    lc.len = L
    return lc
