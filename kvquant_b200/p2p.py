"""Peer-memory exchange of the sequence-sharded partial attention results (kvq_attend_exchange_merge).

Validated against NCCL all_gather + kvq_attend_merge at 2, 4 and 8 GPUs (tests/test_zz_p2p_exchange.py: bit-identical
over 200 rounds); `DecoderStage` uses it when `stage.xchg` is set (bench.py `--sp-exchange p2p`, the default).  One `PeerExchange` per process: allocates this rank's buffer, trades CUDA IPC handles with the
peers over torch.distributed and maps their buffers."""
from __future__ import annotations

import ctypes

import torch

from . import _lib


class PeerExchange:
    def __init__(self, rank, world, num_heads, device, group=None):
        import torch.distributed as dist
        self.lib = _lib.load()
        self.rank, self.world, self.H = int(rank), int(world), int(num_heads)
        self.device = torch.device(device)
        nbytes = self.lib.kvq_p2p_buffer_bytes(self.world, self.H)
        own = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.kvq_p2p_alloc(ctypes.byref(own), nbytes, handle), "kvq_p2p_alloc")
            mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=self.device)
            allh = torch.empty(self.world * 64, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(allh, mine, group=group)
            allh = allh.cpu().numpy().reshape(self.world, 64)
            self._own, self._opened, ptrs = own, [], []
            for r in range(self.world):
                if r == self.rank:
                    ptrs.append(own.value)
                    continue
                p = ctypes.c_void_p()
                h = (ctypes.c_ubyte * 64)(*allh[r].tolist())
                _lib.check(self.lib.kvq_p2p_open(h, ctypes.byref(p)), "kvq_p2p_open")
                self._opened.append(p)
                ptrs.append(p.value)
            self.peers_dev = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
            self.seq = torch.zeros(1, dtype=torch.int64, device=self.device)      # exchanges done so far (lockstep)
            self.err = torch.zeros(1, dtype=torch.int32, device=self.device)
            dist.barrier(group=group)   # every rank has mapped every buffer before the first push

    def exchange_merge(self, part, out):
        """part: f32 [H*128 + H] = this rank's (out, lse); out: f32 [H*128] receives the merged attention output."""
        _lib.check(self.lib.kvq_attend_exchange_merge(part.data_ptr(), self.peers_dev.data_ptr(), self.world, self.rank,
                                                      self.H, self.seq.data_ptr(), out.data_ptr(), self.err.data_ptr(),
                                                      torch.cuda.current_stream(self.device).cuda_stream),
                   "kvq_attend_exchange_merge")
        return out

    def failed(self):
        return bool(self.err.item())

    def close(self):
        for p in self._opened:
            self.lib.kvq_p2p_close(p)
        self._opened = []
        if self._own is not None:
            self.lib.kvq_p2p_free(self._own)
            self._own = None
