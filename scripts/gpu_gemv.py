"""Micro-benchmark of kvq_dec_gemv against torch.mv (cuBLAS) on the LLaMA-7B decode shapes; weights cycle through
more matrices than L2 holds so every call streams from HBM.  Prints one JSON line per shape."""
import json
import sys
import os
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kvquant_b200 import _lib  # noqa: E402


def main():
    lib = _lib.load()
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    for (N, K, kind) in [(12288, 4096, 3), (4096, 4096, 1), (22016, 4096, 3), (4096, 11008, 2), (32000, 4096, 3)]:
        nmat = max(2, int(600e6 // (N * K * 2)) + 1)
        Ws = [(torch.randn(N, K, device=dev) * 0.02).half() for _ in range(nmat)]
        xin = torch.randn(2 * K if kind == 2 else K, device=dev)
        x = xin if kind == 1 else xin.half()
        nw = torch.ones(K, device=dev).half()
        res = torch.randn(N, device=dev).half()
        y = torch.empty(N, device=dev, dtype=torch.float16)
        xr = (x[:K] if kind != 2 else (torch.nn.functional.silu(x[:K].float()).half() * x[K:])).half()

        def ours(i):
            _lib.check(lib.kvq_dec_gemv(Ws[i % nmat].data_ptr(), N, K, x.data_ptr(), kind, nw.data_ptr(), 1e-5,
                                        res.data_ptr(), y.data_ptr(), 0, st))

        def cublas(i):
            torch.addmv(res, Ws[i % nmat], xr, out=y)

        out = {"N": N, "K": K, "kind": kind, "bytes": N * K * 2}
        for name, fn in (("ours", ours), ("cublas", cublas)):
            for i in range(5):
                fn(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 40
            e0.record()
            for i in range(iters):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            out[name + "_ms"] = ms
            out[name + "_gbs"] = N * K * 2 / ms / 1e6
        # numerics: kind 0/1 path against fp32 torch
        if kind in (1,):
            ours(0)
            ref = res.float() + Ws[0].float() @ x.half().float()
            out["max_abs_err"] = float((y.float() - ref).abs().max())
        print(json.dumps(out), flush=True)
        del Ws


if __name__ == "__main__":
    main()
