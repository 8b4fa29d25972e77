"""Generic pre-/post-norm transformer encoder on the sm_100a kernels: mirror of the reference's ``model/transformer.py``
(``TransformerEncoder`` :147-171, ``TransformerLayer`` :57-99, ``MultiHeadAttention`` :107-131 with four cloned
``Linear(hidden, hidden)``, ``FeedForward`` :134-143 with exact-erf GELU, LayerNorm eps 1e-12, additive attention mask,
a final LayerNorm in pre-norm mode).  Same constructor contract (``config.hidden_size / num_attention_heads /
intermediate_size / num_hidden_layers / hidden_dropout / attention_dropout / checkpointing``) and ``state_dict`` keys.
MiCo imports this module only for ``GELU`` (mico.py:11); it is built from the same kernels as the towers through the
autograd wrappers in ``functional`` (tcgen05 GEMMs with fused bias / GELU / residual epilogues, fused attention,
LayerNorm).  Dropout must be 0 or the module in eval mode (the kernels have no dropout).
"""
import copy
import math

import torch
import torch.nn as nn

from . import functional as MF
from .ops import BF16


class _Linear(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(o, i))
        self.bias = nn.Parameter(torch.empty(o))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))        # nn.Linear default init
        bound = 1 / math.sqrt(i)
        nn.init.uniform_(self.bias, -bound, bound)


class LayerNorm(nn.Module):
    def __init__(self, d, eps=1e-12):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(d))
        self.bias = nn.Parameter(torch.zeros(d))

    def forward(self, x):
        return MF.layer_norm(x, self.weight, self.bias, self.eps)


class GELU(nn.Module):
    def forward(self, x):
        return MF.gelu(x)


def _mask3(mask, b, S):
    """reference masks are additive and broadcastable to (b, heads, S, S): accept (b,1,1,S), (b,1,S,S), (b,S), (b,S,S)"""
    if mask is None:
        return None
    m = mask.float()
    if m.dim() == 4:
        m = m[:, 0]
        if m.shape[1] == 1:
            m = m[:, 0]
    return m.expand(b, *m.shape[1:]).contiguous()


class MultiHeadAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.linears = nn.ModuleList([_Linear(config.hidden_size, config.hidden_size) for _ in range(4)])
        self.head_num = config.num_attention_heads
        self.hidden_size = config.hidden_size
        self.p_drop = config.attention_dropout

    def forward(self, q, k, v, mask=None, residual=None):
        b, S = q.shape[0], q.shape[1]
        H, d = self.head_num, self.hidden_size // self.head_num
        qh, kh, vh = [MF.linear_tc(x, l.weight, l.bias, out_dtype=BF16).view(b, -1, H, d)
                      for l, x in zip(self.linears, (q, k, v))]
        o = MF.attention(qh, kh, vh, 1.0 / math.sqrt(d), _mask3(mask, b, S))
        l3 = self.linears[-1]
        return MF.linear_tc(o.view(b, S, self.hidden_size), l3.weight, l3.bias, residual=residual)


class FeedForward(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.linear1 = _Linear(config.hidden_size, config.intermediate_size)
        self.linear2 = _Linear(config.intermediate_size, config.hidden_size)
        self.activation = GELU()

    def forward(self, x, residual=None):
        a = MF.linear_gelu(x, self.linear1.weight, self.linear1.bias)
        return MF.linear_tc(a, self.linear2.weight, self.linear2.bias, residual=residual)


class TransformerLayer(nn.Module):
    def __init__(self, config, mode):
        super().__init__()
        self.attention = MultiHeadAttention(config)
        self.ff_layer = FeedForward(config)
        self.p_drop = config.hidden_dropout
        self.layernorm1 = LayerNorm(config.hidden_size, eps=1e-12)
        self.layernorm2 = LayerNorm(config.hidden_size, eps=1e-12)
        self.mode = mode

    def forward(self, hidden_states, attention_mask):
        if self.training and (self.p_drop > 0 or self.attention.p_drop > 0):
            raise NotImplementedError("mico_b200 transformer kernels have no dropout: use p = 0 or eval()")
        if self.mode == 'prenorm':       # transformer.py:75-86; the residual adds ride in the GEMM epilogues
            h = self.layernorm1(hidden_states)
            hidden_states = self.attention(h, h, h, attention_mask, residual=hidden_states)
            return self.ff_layer(self.layernorm2(hidden_states), residual=hidden_states)
        if self.mode == 'postnorm':      # transformer.py:88-99
            hidden_states = self.layernorm1(self.attention(hidden_states, hidden_states, hidden_states, attention_mask,
                                                           residual=hidden_states))
            return self.layernorm2(self.ff_layer(hidden_states, residual=hidden_states))
        raise NotImplementedError


class TransformerEncoder(nn.Module):
    def __init__(self, config, mode='prenorm'):
        super().__init__()
        layer = TransformerLayer(config, mode)
        self.mode = mode
        self.layer = nn.ModuleList([copy.deepcopy(layer) for _ in range(config.num_hidden_layers)])
        if self.mode == 'prenorm':
            self.last_layernorm = LayerNorm(config.hidden_size, eps=1e-12)
        self.checkpointing = config.checkpointing

    def forward(self, input_, attention_mask=None, cross_hidden_states=None, use_cache=False, cache=None,
                cache_first=False, cache_type=None):
        hidden_states = input_
        for layer_module in self.layer:
            if self.checkpointing:
                hidden_states = torch.utils.checkpoint.checkpoint(layer_module, hidden_states, attention_mask,
                                                                  use_reentrant=False)
            else:
                hidden_states = layer_module(hidden_states, attention_mask)
        if self.mode == 'prenorm':
            hidden_states = self.last_layernorm(hidden_states)
        return hidden_states, cache
