"""Linear layer whose weight stays resident on the parameter-server shards: the "pull" happens inside the GEMM.

Reference shape: after every update the workers pull the full weight (`ncclBcast` root->all, kvstore_nccl.h:380-396;
`KVStoreDist::PullImpl` + `Broadcast`, kvstore_dist.h:217-269) and only then the first consumer (FullyConnected ->
cublasSgemmEx) can start.  Here the weight matrix [out, in] is row-sharded over the ranks of the NVSwitch domain
inside the symmetric peer-HBM heap (shard r = rows [r*R, (r+1)*R), exactly the server shard that owns those rows),
and the consumer GEMM (csrc/cuda/gemm_tcgen05.cu) issues its TMA loads against the owning peer's tensor map:
weight tiles travel over NVLink straight into the smem ring that feeds tcgen05.mma - no broadcast kernel, no local
copy of the gathered weight, transfer overlapped with the math tile by tile.

With an NVLS multicast mapping (default, `staged`): the same kernel first lets its idle epilogue warps PUSH this rank's shard
into a stage buffer of ALL ranks with multimem.st (the switch replicates: every rank sends 1/W of the matrix once instead of
every tile pulling its weights again - peer memory is not cached in the local L2), 64-row chunk by chunk with an arrival
counter per chunk (multimem.red.release); the TMA producer starts a tile as soon as the chunks under it have arrived.

forward  y[M,N]  = x[M,K] @ W^T      B = W shards, K-major, sharded over output columns
dgrad    dx[M,K] = dy[M,N] @ W       B = W shards read MN-major, sharded over the contraction dimension
wgrad    dW[N,K] = dy^T @ x          local tcgen05 GEMM into this worker's gradient buffer in the heap; the fused PS
                                     kernel then reduces the shard over workers, applies the optimizer on the fp32
                                     master and writes the new bf16 shard in place (`step`).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import native


class _ShardedLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mod, dummy):
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        ctx.mod, ctx.x_shape = mod, x.shape
        ctx.save_for_backward(x2)
        if mod.staged:        # all-gather fused into the GEMM over NVLS multicast (every rank pushes its shard, chunk counters)
            y = native.gemm_bf16_staged(mod.heap._ctx, x2, shard_off=mod.w_off, stage_off=mod._next_stage(), flags_off=mod.flags_off,
                                        seq=mod._next_seq(), rows_per_shard=mod.rows_per_shard, cols=mod.in_features, b_mn=False,
                                        n=mod.out_padded, k=mod.in_features, bias=mod.bias_f32, channel=-1)
        else:                 # every tile pulls its weights from the owning peer by TMA
            y = native.gemm_bf16(x2, None, shard_ptrs=mod.weight_ptrs, rows_per_shard=mod.rows_per_shard, ldb=mod.in_features,
                                 n=mod.out_padded, k=mod.in_features, bias=mod.bias_f32, block_n=128)
        return y.view(*x.shape[:-1], mod.out_padded)

    @staticmethod
    def backward(ctx, dy):
        mod = ctx.mod
        (x2,) = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        # dgrad: contraction over N, whose rows live on different peers
        if mod.staged:
            dx = native.gemm_bf16_staged(mod.heap._ctx, dy2, shard_off=mod.w_off, stage_off=mod._next_stage(), flags_off=mod.flags_off,
                                         seq=mod._next_seq(), rows_per_shard=mod.rows_per_shard, cols=mod.in_features, b_mn=True,
                                         n=mod.in_features, k=mod.out_padded, channel=-1)
        else:
            dx = native.gemm_bf16(dy2, None, b_mn=True, shard_ptrs=mod.weight_ptrs, rows_per_shard=mod.rows_per_shard,
                                  ldb=mod.in_features, n=mod.in_features, k=mod.out_padded, block_n=128)
        # wgrad into the symmetric gradient buffer (accumulates across micro-batches until `step`)
        native.gemm_bf16(dy2, x2, a_mn=True, b_mn=True, out=mod.grad_full.view(mod.out_padded, mod.in_features),
                         accumulate=mod._grad_live)
        mod._grad_live = True
        if mod.bias_f32 is not None:
            mod.bias_grad += dy2.sum(0, dtype=torch.float32)
        return dx.view(ctx.x_shape), None, None


class PSResidentLinear(nn.Module):
    def __init__(self, in_features: int, out_features: int, heap, bias: bool = True, init_std: float = 0.02, seed: int = 0):
        super().__init__()
        if in_features % 8:
            raise ValueError("in_features must be a multiple of 8")
        self.heap, self.in_features, self.out_features = heap, in_features, out_features
        W = heap.world
        rows = (out_features + W - 1) // W
        self.rows_per_shard = (rows + 127) // 128 * 128            # a 128-wide tile never straddles two shards
        self.out_padded = self.rows_per_shard * W
        n_shard = self.rows_per_shard * in_features
        self.w_off, self.weight_shard = heap.alloc_tensor(n_shard, torch.bfloat16)
        self.g_off, self.grad_full = heap.alloc_tensor(self.out_padded * in_features, torch.bfloat16)
        self.grad_full.zero_()
        g = torch.Generator(device="cpu").manual_seed(seed)
        full = torch.zeros(self.out_padded, in_features)
        full[:out_features] = torch.randn(out_features, in_features, generator=g) * init_std
        lo = heap.rank * self.rows_per_shard
        mine = full[lo:lo + self.rows_per_shard].to(heap.device)
        self.master = mine.float().reshape(-1).clone()
        self.momentum = torch.zeros_like(self.master)
        self.weight_shard.copy_(mine.reshape(-1).to(torch.bfloat16))
        self.weight_ptrs = [heap.peer_heap_ptr(r) + self.w_off for r in range(W)]
        self.bias_f32 = torch.zeros(self.out_padded, device=heap.device) if bias else None
        self.bias_grad = torch.zeros(self.out_padded, device=heap.device) if bias else None
        self._grad_live = False
        # staged pull (NVLS): a stage buffer for the gathered matrix - scratch shared by every PSResidentLinear of the heap,
        # sized for the largest - and per-layer arrival counters (zeroed once, monotonic)
        import os
        self.staged = bool(heap.has_multicast) and W > 1 and os.environ.get("DLCFN_PULL_GEMM", "staged") == "staged"
        if self.staged:
            # TWO stage buffers used alternately by every staged GEMM of the heap, in launch order: a rank can only reach staged
            # launch i+2 after it received every rank's chunks of launch i+1, which each rank pushes after finishing launch i -
            # so the buffer of launch i is free everywhere without an entry barrier (needs out_features > (W-1) * rows_per_shard,
            # i.e. every GEMM consumes rows of every shard: true by construction of rows_per_shard)
            need = (self.out_padded * in_features * 2 + 1023) // 1024 * 1024
            st = heap.__dict__.get("_gemm_stage")
            if st is None or st[1] < need:
                st = (heap.alloc(2 * need, 1024), need)
                heap.__dict__["_gemm_stage"] = st
                heap.__dict__.setdefault("_gemm_stage_launches", 0)
            nflags = self.out_padded // 64
            self.flags_off, flags = heap.alloc_tensor(nflags, torch.int32)
            flags.zero_()
            self._seq = 0
        self._anchor = nn.Parameter(torch.zeros(1, device=heap.device))      # gives autograd a leaf to reach us through
        heap.barrier()
        torch.cuda.current_stream().synchronize()

    def _next_seq(self) -> int:
        self._seq += 1
        return self._seq

    def _next_stage(self) -> int:
        h = self.heap.__dict__
        off, size = h["_gemm_stage"]
        n = h["_gemm_stage_launches"]
        h["_gemm_stage_launches"] = n + 1
        return off + (n & 1) * size

    def forward(self, x):
        y = _ShardedLinearFn.apply(x, self, self._anchor)
        return y if self.out_padded == self.out_features else y[..., : self.out_features]

    @torch.no_grad()
    def step(self, lr: float, momentum: float = 0.9, wd: float = 0.0) -> None:
        """push (gradient rows of my shard from every worker) + reduce + SGD-momentum on the fp32 master + write the new
        bf16 shard in place - one fused kernel; the next forward pulls it through TMA."""
        h = self.heap
        W = h.world
        lo = h.rank * self.rows_per_shard * self.in_features
        hi = lo + self.rows_per_shard * self.in_features
        # I am the server of rows [lo, hi) of the full gradient that EVERY rank computed; the updated bf16 shard is
        # written only into my own shard storage (pull_mask = me): w_off - lo*2 makes element lo land on w_off.
        h.ps_sync_step(self.g_off, self.w_off - lo * 2, self.out_padded * self.in_features, lo, hi, (1 << W) - 1,
                       self.master, self.momentum, None, torch.bfloat16, "sgd_mom", lr=lr, wd=wd, momentum=momentum,
                       rescale=1.0 / W, algo=0, pull_mask=1 << h.rank)
        torch.cuda.current_stream().synchronize()
        self.grad_full.zero_()
        self._grad_live = False
        if self.bias_f32 is not None:
            import torch.distributed as dist
            if W > 1:
                dist.all_reduce(self.bias_grad)
            self.bias_f32.add_(self.bias_grad, alpha=-lr / W)
            self.bias_grad.zero_()
        h.barrier()      # every shard is updated before anybody's next forward pulls it
