"""Autograd wrappers over the fused sm_100a Llama-block kernels (src/kernels/model_kernels.cu).

`rope_split(qkv, cos, sin, H, KV, D)`  -> q[B,S,H,D], k[B,S,KV,D] (rotated), v[B,S,KV,D]
`swiglu(gu)`                           -> silu(gu[..., :F]) * gu[..., F:]
Each has a plain-PyTorch reference (`*_reference`) used by the numerics tests and as the
CPU fallback of the model.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .. import native


def rope_split_reference(qkv, cos, sin, n_heads: int, n_kv: int, hd: int):
    B, S, _ = qkv.shape
    q, k, v = qkv.split([n_heads * hd, n_kv * hd, n_kv * hd], dim=-1)

    def rot(x, h):
        x = x.reshape(B, S, h, hd).float()
        x1, x2 = x[..., : hd // 2], x[..., hd // 2:]
        c, s = cos[:S, None, :], sin[:S, None, :]
        return torch.cat((x1 * c - x2 * s, x2 * c + x1 * s), dim=-1).to(qkv.dtype)

    return rot(q, n_heads), rot(k, n_kv), v.reshape(B, S, n_kv, hd)


def swiglu_reference(gu):
    g, u = gu.chunk(2, dim=-1)
    return F.silu(g.float()).mul(u.float()).to(gu.dtype)


class _RopeSplit(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, cos, sin, n_heads, n_kv, hd):
        q, k, v = native().rope_split(qkv.contiguous(), cos, sin, n_heads, n_kv, hd)
        ctx.save_for_backward(cos, sin)
        ctx.dims = (n_heads, n_kv, hd)
        return q, k, v

    @staticmethod
    def backward(ctx, dq, dk, dv):
        cos, sin = ctx.saved_tensors
        h, kv, hd = ctx.dims
        dqkv = native().rope_merge_bwd(dq.contiguous(), dk.contiguous(), dv.contiguous(), cos, sin, h, kv, hd)
        return dqkv, None, None, None, None, None


class _SwiGLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gu):
        gu = gu.contiguous()
        ctx.save_for_backward(gu)
        return native().swiglu_fwd(gu)

    @staticmethod
    def backward(ctx, dout):
        (gu,) = ctx.saved_tensors
        return native().swiglu_bwd(gu, dout.contiguous())


def rope_split(qkv, cos, sin, n_heads: int, n_kv: int, hd: int):
    if qkv.is_cuda and qkv.dtype == torch.bfloat16 and hd % 16 == 0:
        return _RopeSplit.apply(qkv, cos, sin, n_heads, n_kv, hd)
    return rope_split_reference(qkv, cos, sin, n_heads, n_kv, hd)


def swiglu(gu):
    if gu.is_cuda and gu.dtype == torch.bfloat16 and (gu.shape[-1] // 2) % 8 == 0:
        return _SwiGLU.apply(gu)
    return swiglu_reference(gu)
