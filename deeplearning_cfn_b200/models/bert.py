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


class BertLayer(nn.Module):
    def __init__(self, hidden: int, heads: int, ffn: int, dropout: float):
        super().__init__()
        self.heads = heads
        self.qkv = Linear(hidden, 3 * hidden)
        self.proj = Linear(hidden, hidden)
        self.ln1 = nn.LayerNorm(hidden, eps=1e-12)
        self.fc1 = Linear(hidden, ffn)
        self.fc2 = Linear(ffn, hidden)
        self.ln2 = nn.LayerNorm(hidden, eps=1e-12)
        self.drop = dropout

    def forward(self, x, mask=None):
        B, S, H = x.shape
        qkv = self.qkv(x).view(B, S, 3, self.heads, H // self.heads).permute(2, 0, 3, 1, 4)
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
