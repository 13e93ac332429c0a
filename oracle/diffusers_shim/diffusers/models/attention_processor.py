"""Attention + AttnProcessor(2_0) restated."""
import inspect

import torch
import torch.nn.functional as F
from torch import nn


class AttnProcessor2_0:
    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, **_):
        residual = hidden_states
        input_ndim = hidden_states.ndim
        if attn.group_norm is not None:
            b, c, h, w = hidden_states.shape
            hidden_states = attn.group_norm(hidden_states.view(b, c, h * w)).transpose(1, 2)
        q = attn.to_q(hidden_states)
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        k = attn.to_k(ctx)
        v = attn.to_v(ctx)
        B = q.shape[0]
        hd = q.shape[-1] // attn.heads
        q = q.view(B, -1, attn.heads, hd).transpose(1, 2)
        k = k.view(B, -1, attn.heads, hd).transpose(1, 2)
        v = v.view(B, -1, attn.heads, hd).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(B, -1, attn.heads * hd).to(q.dtype)
        o = attn.to_out[0](o)
        o = attn.to_out[1](o)
        if input_ndim == 4:
            o = o.transpose(-1, -2).reshape(b, c, h, w)
        if attn.residual_connection:
            o = o + residual
        return o / attn.rescale_output_factor


class AttnProcessor:
    """baddbmm + softmax equivalent."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, **_):
        q = attn.to_q(hidden_states)
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        k = attn.to_k(ctx)
        v = attn.to_v(ctx)
        B = q.shape[0]
        hd = q.shape[-1] // attn.heads
        q = q.view(B, -1, attn.heads, hd).transpose(1, 2)
        k = k.view(B, -1, attn.heads, hd).transpose(1, 2)
        v = v.view(B, -1, attn.heads, hd).transpose(1, 2)
        s = torch.softmax((q @ k.transpose(-1, -2)) * attn.scale, dim=-1)
        o = (s @ v).transpose(1, 2).reshape(B, -1, attn.heads * hd)
        o = attn.to_out[0](o)
        return attn.to_out[1](o) / attn.rescale_output_factor


AttentionProcessor = AttnProcessor2_0


class AttnAddedKVProcessor:  # name only (unet_2d_condition.py:11-17 imports it; SD-1.5 never selects it)
    pass


ADDED_KV_ATTENTION_PROCESSORS = (AttnAddedKVProcessor,)
CROSS_ATTENTION_PROCESSORS = (AttnProcessor, AttnProcessor2_0)


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False,
                 upcast_attention=False, upcast_softmax=False, norm_num_groups=None, eps=1e-5,
                 rescale_output_factor=1.0, residual_connection=False, out_bias=True, processor=None, **_):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.rescale_output_factor = rescale_output_factor
        self.residual_connection = residual_connection
        kv = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.group_norm = nn.GroupNorm(norm_num_groups, query_dim, eps=eps) if norm_num_groups else None
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(kv, inner, bias=bias)
        self.to_v = nn.Linear(kv, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=out_bias), nn.Dropout(dropout)])
        self.processor = processor or AttnProcessor2_0()

    def set_processor(self, processor):
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        ok = set(inspect.signature(self.processor.__call__).parameters)
        kw = {k: v for k, v in kw.items() if k in ok}
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kw)
