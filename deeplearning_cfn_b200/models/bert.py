"""BERT (encoder + MLM/NSP pre-training heads).  BERT-large = 24 layers, hidden 1024, 16 heads, FFN 4096,
vocab 30522, seq 512: the all-reduce-with-bucket-fusion configuration named in BASELINE.json (the reference
tree has no transformer).  Every projection (QKV, attention output, FFN, MLM transform, tied
decoder) runs on the hand-written tcgen05 GEMM (`ops.linear`), attention itself through
`F.scaled_dot_product_attention`."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..ops.layernorm import dropout_add_layer_norm
from ..ops.linear import Linear, linear


# DLCFN_BERT_QKV_PLANES=1: q / k / v as three dense planes (see _QKVFn).  Written after the round's GPU budget was spent:
# host-tested only, so the packed projection (measured: 412 sequences/s) stays the default.
_QKV_PLANES = __import__("os").environ.get("DLCFN_BERT_QKV_PLANES", "0") == "1"


class _QKVFn(torch.autograd.Function):
    """The three attention projections of one input as three tcgen05 GEMMs that write q, k and v as separate DENSE [tokens, H]
    planes of one buffer (so that the [B, h, S, d] views handed to the attention kernel are dense in B-S-h-d order and neither
    the attention forward nor its backward has to re-lay anything out), with ONE parameter [3H, H] / [3H] behind them:
        fwd    plane_j = x @ W_j^T + b_j                          (3 GEMMs, bias in the epilogue)
        dgrad  dx = dq @ W_q + dk @ W_k + dv @ W_v                (3 GEMMs accumulating into one buffer, no add kernels)
        wgrad  dW_j = dplane_j^T @ x, db_j = colsum(dplane_j)     (3 split-K GEMMs into the row blocks of one [3H, H] gradient)
    The packed form (one GEMM -> [tokens, 3H] -> view / permute) cost 4 layout copies + 3 zero fills per layer and step in
    autograd's select / permute backward (profiles/r2/profile_bert_before.txt: 99 copies, 78 fills per step)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        from ..ops import native
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        H = weight.shape[1]
        M = x2.shape[0]
        out = torch.empty((3, M, H), dtype=torch.bfloat16, device=x.device)
        bf = None if bias is None else (bias if bias.dtype == torch.float32 else bias.float())
        for j in range(3):
            native.gemm_bf16(x2, weight[j * H:(j + 1) * H], bias=None if bf is None else bf[j * H:(j + 1) * H], out=out[j])
        ctx.save_for_backward(x2, weight)
        ctx.has_bias = bias is not None
        ctx.x_shape = x.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        from ..ops import native
        x2, weight = ctx.saved_tensors
        H = weight.shape[1]
        M = x2.shape[0]
        planes = []
        for j in range(3):
            d = dout[j]
            planes.append(d if (d.stride(1) == 1 and d.stride(0) == H) else d.contiguous())
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((M, H), dtype=torch.bfloat16, device=x2.device)
            for j in range(3):
                native.gemm_bf16(planes[j], weight[j * H:(j + 1) * H], b_mn=True, out=dx, accumulate=j > 0)
            dx = dx.view(ctx.x_shape)
        if ctx.needs_input_grad[1]:
            dw = torch.empty_like(weight)
            for j in range(3):
                native.gemm_bf16(planes[j], x2, a_mn=True, b_mn=True, out=dw[j * H:(j + 1) * H])
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.cat([native.colsum(planes[j]) for j in range(3)])
        return dx, dw, db


class QKVProjection(Linear):
    """`Linear(hidden, 3 * hidden)` (same parameter names and shapes: `weight` [3H, H], `bias` [3H]) whose output is the three
    dense planes (3, tokens, hidden)."""

    def __init__(self, hidden: int):
        super().__init__(hidden, 3 * hidden)

    def forward(self, x):
        w, b = self.weight, self.bias
        if _QKV_PLANES and x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16:
            return _QKVFn.apply(x, w, b)
        y = super().forward(x)                                           # host / fp32 path
        return y.reshape(-1, 3, w.shape[1]).permute(1, 0, 2)


class BertLayer(nn.Module):
    def __init__(self, hidden: int, heads: int, ffn: int, dropout: float):
        super().__init__()
        self.heads = heads
        self.qkv = QKVProjection(hidden)
        self.proj = Linear(hidden, hidden)
        self.ln1 = nn.LayerNorm(hidden, eps=1e-12)
        self.fc1 = Linear(hidden, ffn)
        self.fc2 = Linear(ffn, hidden)
        self.ln2 = nn.LayerNorm(hidden, eps=1e-12)
        self.drop = dropout

    def forward(self, x, mask=None):
        B, S, H = x.shape
        qkv = self.qkv(x).reshape(3, B, S, self.heads, H // self.heads).transpose(2, 3)     # three [B, h, S, d] views, dense as B-S-h-d
        a = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2], attn_mask=mask,
                                           dropout_p=self.drop if self.training else 0.0)
        a = a.transpose(1, 2).reshape(B, S, H)
        # residual + dropout + LayerNorm: one fused kernel each way when DLCFN_FUSED_LN=1 (ops/layernorm.py), torch otherwise
        x = dropout_add_layer_norm(self.proj(a), x, self.ln1.weight, self.ln1.bias, self.drop, self.ln1.eps, self.training)
        h = self.fc2(F.gelu(self.fc1(x)))
        return dropout_add_layer_norm(h, x, self.ln2.weight, self.ln2.bias, self.drop, self.ln2.eps, self.training)


class Bert(nn.Module):
    def __init__(self, vocab: int = 30522, hidden: int = 1024, layers: int = 24, heads: int = 16, ffn: int = 4096,
                 max_seq: int = 512, type_vocab: int = 2, dropout: float = 0.1):
        super().__init__()
        self.vocab = vocab
        self.vocab_padded = (vocab + 7) // 8 * 8
        self.tok = nn.Embedding(self.vocab_padded, hidden)
        self.pos = nn.Embedding(max_seq, hidden)
        self.typ = nn.Embedding(type_vocab, hidden)
        self.ln = nn.LayerNorm(hidden, eps=1e-12)
        self.layers = nn.ModuleList(BertLayer(hidden, heads, ffn, dropout) for _ in range(layers))
        self.pooler = Linear(hidden, hidden)
        self.mlm_dense = Linear(hidden, hidden)
        self.mlm_ln = nn.LayerNorm(hidden, eps=1e-12)
        self.mlm_bias = nn.Parameter(torch.zeros(self.vocab_padded))
        self.nsp = Linear(hidden, 2)
        self.drop = dropout
        self.apply(self._init)

    @staticmethod
    def _init(m):
        if isinstance(m, (Linear, nn.Embedding)):
            nn.init.normal_(m.weight, std=0.02)
        if isinstance(m, Linear) and m.bias is not None:
            nn.init.zeros_(m.bias)

    def forward(self, ids, type_ids=None, mlm_positions=None):
        B, S = ids.shape
        pos = torch.arange(S, device=ids.device)
        x = self.tok(ids) + self.pos(pos)[None] + (self.typ(type_ids) if type_ids is not None else self.typ.weight[0])
        x = F.dropout(self.ln(x), self.drop, self.training)
        for l in self.layers:
            x = l(x)
        pooled = torch.tanh(self.pooler(x[:, 0]))
        h = x if mlm_positions is None else torch.gather(x, 1, mlm_positions[..., None].expand(-1, -1, x.shape[-1]))
        h = self.mlm_ln(F.gelu(self.mlm_dense(h)))
        logits = linear(h, self.tok.weight, self.mlm_bias)                        # decoder tied to the embedding
        if getattr(self, "padded_logits", False):
            # keep the TMA padding columns: bert_pretrain_loss masks them inside the fused loss kernel instead of
            # paying a slice copy forward and a zero-pad copy backward on the [tokens, vocab] tensor
            return logits, self.nsp(pooled), self.vocab
        return logits[..., : self.vocab], self.nsp(pooled)


def bert_pretrain_loss(out, target):
    from ..ops.nn import softmax_cross_entropy
    logits, nsp = out[0], out[1]
    vocab = out[2] if len(out) > 2 else None
    mlm_labels, nsp_labels = target
    l1 = softmax_cross_entropy(logits, mlm_labels, ignore_index=-100, num_classes=vocab)
    return l1 + softmax_cross_entropy(nsp, nsp_labels)


def bert_large(**kw) -> Bert:
    return Bert(hidden=1024, layers=24, heads=16, ffn=4096, **kw)


def bert_base(**kw) -> Bert:
    return Bert(hidden=768, layers=12, heads=12, ffn=3072, **kw)
