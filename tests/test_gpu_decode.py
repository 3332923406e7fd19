"""GPU: the decode harness -- fused helper kernels against their torch formulation, and a CUDA-graph-captured decode
step against the eager step."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _stage(L=96, bits=4, n_sink=0, layers=2):
    from kvquant_b200 import decode as kd, synth, cache as kc
    cfg = kd.DecodeConfig(n_layers=layers, hidden=4096, n_heads=32, intermediate=1024, vocab=512, bits=bits,
                          n_sink=n_sink, max_len=L + 64)
    sp = synth.SynthSpec(32, 128, seed=0)
    cal = synth.calibrate(sp, bits, calib_tokens=256, seed=7)
    t = kc.build_k_lookup_table(cal["k"][0], cal["k"][1], cal["k"][2][0], 32, device=DEV)
    quant = dict(klut=dict(lut=t["lut"], lut2=None, thr_lower=t["thr_lower"], thr_upper=t["thr_upper"]), v_cent=cal["v"][2][0])
    st = kd.DecoderStage(cfg, 0, layers, DEV, quant, seed=1, with_head=True)
    for i, ly in enumerate(st.layers):
        synth.fill_layer_cache_gpu(ly.cache, sp, L, seed=i, chunk=64)
    return cfg, st


def test_helper_kernels_match_torch_formulation():
    cfg, st = _stage()
    L = st.layers[0].cache.len
    x = (torch.randn(cfg.hidden, device=DEV) * 0.5).half()
    y_ref = st.forward_torch(x.clone())
    st.set_len(L)
    y = st.forward(x.clone())
    st.set_len(L)
    d = (y.float() - y_ref.float()).abs().max().item()
    assert d <= 2e-2 * max(1.0, y_ref.float().abs().max().item()), d   # fp16 GEMV accumulation-order noise only


def test_graphed_step_replays_identically():
    from kvquant_b200 import decode as kd
    cfg, st = _stage(n_sink=3, bits=3)
    L = st.layers[0].cache.len
    gs = kd.GraphedStage(st, L, first=True, last_to_logits=True)
    gs.tok.fill_(7)
    gs.replay()
    torch.cuda.synchronize()
    a = gs.logits.clone()
    gs.replay()                       # same slot is overwritten: replays are idempotent
    torch.cuda.synchronize()
    # (not bit-identical: the outlier scatter uses shared-memory atomics whose order varies, as in the reference)
    assert (a.float() - gs.logits.float()).abs().max().item() <= 1e-2 * max(1.0, a.float().abs().max().item())
    st.set_len(L)
    y = st.forward(st.embed_token(gs.tok))
    ref = st.head(y)
    assert (ref.float() - a.float()).abs().max().item() <= 1e-2 * max(1.0, ref.float().abs().max().item())


@pytest.mark.parametrize("N,K,kind", [(4096, 4096, 0), (333, 1024, 1), (1000, 11008, 2), (12288, 4096, 3), (7, 256, 3)])
def test_fused_gemv_matches_fp32_torch(N, K, kind):
    """kvq_dec_gemv against a plain PyTorch fp32 formulation of the same op (fp16 weights, fp32 accumulate):
    tolerance 2e-3 of the output scale (fp16 output rounding + accumulation order)."""
    from kvquant_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(N + K + kind)
    W = (torch.randn((N, K), generator=g, device=DEV) * 0.03).half()
    res = torch.randn(N, generator=g, device=DEV).half()
    nw = (1 + 0.1 * torch.randn(K, generator=g, device=DEV)).half()
    eps = 1e-5
    if kind == 1:
        x = torch.randn(K, generator=g, device=DEV)
        fx = x.half().float()
    elif kind == 2:
        x = torch.randn(2 * K, generator=g, device=DEV).half()
        gte = x[:K].float()
        fx = ((gte / (1 + torch.exp(-gte))).half() * x[K:]).float()
    else:
        x = torch.randn(K, generator=g, device=DEV).half()
        fx = x.float()
        if kind == 3:
            fx = ((x.float() * torch.rsqrt((x.float() ** 2).mean() + eps)).half() * nw).float()
    want = res.float() + W.float() @ fx
    st = torch.cuda.current_stream().cuda_stream
    for y_f32 in (0, 1):
        y = torch.empty(N, device=DEV, dtype=torch.float32 if y_f32 else torch.float16)
        _lib.check(lib.kvq_dec_gemv(W.data_ptr(), N, K, x.data_ptr(), kind, nw.data_ptr(), eps, res.data_ptr(),
                                    y.data_ptr(), y_f32, st))
        torch.cuda.synchronize()
        err = (y.float() - want).abs().max().item()
        assert err <= 2e-3 * max(1.0, want.abs().max().item()), (err, y_f32)
    # in place on the residual, no residual, and the loud failures
    r2 = res.clone()
    _lib.check(lib.kvq_dec_gemv(W.data_ptr(), N, K, x.data_ptr(), kind, nw.data_ptr(), eps, r2.data_ptr(), r2.data_ptr(), 0, st))
    y0 = torch.empty(N, device=DEV, dtype=torch.float32)
    _lib.check(lib.kvq_dec_gemv(W.data_ptr(), N, K, x.data_ptr(), kind, nw.data_ptr(), eps, None, y0.data_ptr(), 1, st))
    torch.cuda.synchronize()
    assert (r2.float() - want).abs().max().item() <= 2e-3 * max(1.0, want.abs().max().item())
    assert (y0 - (want - res.float())).abs().max().item() <= 2e-3 * max(1.0, want.abs().max().item())
    assert lib.kvq_dec_gemv(W.data_ptr(), N, K - 8, x.data_ptr(), kind, nw.data_ptr(), eps, None, y0.data_ptr(), 1, st) != 0
    assert lib.kvq_dec_gemv(W.data_ptr(), N, K, x.data_ptr(), 3, None, eps, None, y0.data_ptr(), 1, st) != 0


def test_dynamic_length_graph_decodes_a_growing_cache():
    """ONE captured graph with the device-resident length replayed 4 times == 4 eager steps with host lengths
    (appends at slots L..L+3, attends over L+1..L+4 slots, positions advance)."""
    from kvquant_b200 import decode as kd
    cfg, st = _stage(n_sink=3, bits=3)
    L = st.layers[0].cache.len
    gs = kd.GraphedStage(st, L, first=True, last_to_logits=True, dynamic=True)
    toks = [7, 11, 3, 250]
    got = []
    for t in toks:
        gs.tok.fill_(t)
        gs.replay()
        torch.cuda.synchronize()
        got.append(gs.logits.clone())
    assert st.layers[0].cache.len == L + len(toks)
    assert int(st.dyn["len"].item()) == L + len(toks)
    # eager reference: same tokens from the same starting state (the appends rewrite slots L..L+3 identically)
    st.dyn = None
    st.set_len(L)
    for i, t in enumerate(toks):
        y = st.forward(st.embed_token(torch.tensor([t], device=DEV)))
        ref = st.head(y)
        d = (ref.float() - got[i].float()).abs().max().item()
        assert d <= 1e-2 * max(1.0, ref.float().abs().max().item()), (i, d)
    assert st.layers[0].cache.len == L + len(toks)
