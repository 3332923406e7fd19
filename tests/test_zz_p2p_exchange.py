"""GPU, OPT-IN (KVQ_TEST_P2P=1): the experimental peer-memory exchange + merge kernel (kvq_attend_exchange_merge) in a
single-GPU loopback -- two 'ranks' with their own buffers, counters and streams on one device trade partial attention
results through each other's buffers and must both end up with kvq_attend_merge's answer, over several exchanges
(both buffer parities).  Opt-in because the path has not been validated on hardware yet; the
file name keeps it last in the collection order."""
import ctypes
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("KVQ_TEST_P2P") != "1", reason="experimental path: set KVQ_TEST_P2P=1")]
DEV = "cuda:0"


def test_two_rank_loopback_matches_attend_merge():
    from kvquant_b200 import _lib
    lib = _lib.load()
    world, H = 2, 32
    n = H * 128 + H
    nbytes = lib.kvq_p2p_buffer_bytes(world, H)
    bufs, handles = [], []
    for _ in range(world):
        p, h = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
        _lib.check(lib.kvq_p2p_alloc(ctypes.byref(p), nbytes, h))
        bufs.append(p)
        handles.append(h)
    peers = torch.tensor([b.value for b in bufs], dtype=torch.int64, device=DEV)
    seqs = [torch.zeros(1, dtype=torch.int64, device=DEV) for _ in range(world)]
    errs = [torch.zeros(1, dtype=torch.int32, device=DEV) for _ in range(world)]
    streams = [torch.cuda.Stream(device=DEV) for _ in range(world)]
    g = torch.Generator(device=DEV).manual_seed(0)
    try:
        for it in range(5):
            parts = torch.randn((world, n), generator=g, device=DEV)
            parts[:, H * 128:] = parts[:, H * 128:] * 3 + 10          # lse values
            want = torch.empty(H * 128, device=DEV)
            _lib.check(lib.kvq_attend_merge(parts.data_ptr(), world, H, want.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            outs = [torch.empty(H * 128, device=DEV) for _ in range(world)]
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    _lib.check(lib.kvq_attend_exchange_merge(parts[r].data_ptr(), peers.data_ptr(), world, r, H,
                                                             seqs[r].data_ptr(), outs[r].data_ptr(), errs[r].data_ptr(),
                                                             streams[r].cuda_stream))
            torch.cuda.synchronize()
            for r in range(world):
                assert int(errs[r].item()) == 0 and int(seqs[r].item()) == it + 1
                assert torch.allclose(outs[r], want, rtol=1e-5, atol=1e-6), (it, r)
    finally:
        for b in bufs:
            lib.kvq_p2p_free(b)
