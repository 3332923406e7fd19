#!/usr/bin/env python
"""torchrun --nproc-per-node N tests/_p2p_check.py : kvq_attend_exchange_merge (peer-memory exchange fused with the
merge) against NCCL all_gather + kvq_attend_merge on the same random partial results, many rounds (both buffer parities,
advancing sequence numbers).  Prints the max difference on rank 0."""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kvquant_b200 import _lib
from kvquant_b200.p2p import PeerExchange
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
H = int(os.environ.get("P2P_H", "32"))
lib = _lib.load()
x = PeerExchange(rank, world, H, dev)
n = H * 128 + H
g = torch.Generator(device=dev).manual_seed(100 + rank)
worst = 0.0
st = torch.cuda.current_stream().cuda_stream
for it in range(int(os.environ.get("P2P_ROUNDS", "200"))):
    part = torch.randn(n, generator=g, device=dev)
    part[H * 128:] = torch.randn(H, generator=g, device=dev) * 3 + rank          # lse
    gath = torch.empty(world * n, device=dev)
    dist.all_gather_into_tensor(gath, part)
    want = torch.empty(H * 128, device=dev)
    _lib.check(lib.kvq_attend_merge(gath.data_ptr(), world, H, want.data_ptr(), st))
    got = torch.empty(H * 128, device=dev)
    x.exchange_merge(part, got)
    torch.cuda.synchronize()
    worst = max(worst, (got - want).abs().max().item() / want.abs().max().item())
t = torch.tensor([worst], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("P2P_CHECK world=%d rounds ok, max rel diff %.3e" % (world, t.item()), flush=True)
dist.barrier()
os._exit(0)
