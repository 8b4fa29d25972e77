"""OpenAI-CLIP vision tower on the sm_100a kernels: mirror of the reference's ``model/clip/clip.py`` ``VisionTransformer``
(:228-295) with ``ResidualAttentionBlock`` (:173-210: ``nn.MultiheadAttention`` + QuickGELU MLP, pre-norm) for the
``clip_vit_base_16`` / ``clip_vit_large_14_336px`` encoder types (mico.py:354-371).  Same constructor, ``state_dict`` keys
(``conv1.weight``, ``class_embedding``, ``positional_embedding``, ``ln_pre``, ``transformer.resblocks.N.{ln_1, attn.in_proj_*,
attn.out_proj, ln_2, mlp.c_fc, mlp.c_proj}``, ``ln_post``, ``proj``) and ``forward(x, return_all_features)`` contract.

It reuses the EVA tower's launch sequences (eva_vit.py) with four switches: no patch-embed bias, ``ln_pre`` in front
of the blocks, a full ``in_proj_bias`` (q, k and v), and QuickGELU ``x * sigmoid(1.702 x)`` in the fc1 / fc2-dgrad epilogues.
"""
import torch
import torch.nn as nn

from . import functional as MF
from .eva_vit import EVAVisionTransformer, LayerNorm, _ParamLinear
from .ops import ACT_MUL_AUX, ACT_QUICK_GELU_SAVE_GRAD, MicoError


class _MHA(nn.Module):
    """Parameter holder with nn.MultiheadAttention's names."""

    def __init__(self, d):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d, d))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
        self.out_proj = _ParamLinear(d, d)
        nn.init.xavier_uniform_(self.in_proj_weight)


class _Mlp(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.c_fc = _ParamLinear(d, 4 * d)
        self.c_proj = _ParamLinear(4 * d, d)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model, n_head):
        super().__init__()
        self.attn = _MHA(d_model)
        self.ln_1 = LayerNorm(d_model, 1e-5)
        self.mlp = _Mlp(d_model)
        self.ln_2 = LayerNorm(d_model, 1e-5)
        self.drop_prob = 0.0


class _Transformer(nn.Module):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.ModuleList([ResidualAttentionBlock(width, heads) for _ in range(layers)])


class _Conv(nn.Module):
    def __init__(self, width, patch):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(width, 3, patch, patch).normal_(0.0, 0.02))


class VisionTransformer(EVAVisionTransformer):
    def __init__(self, input_resolution, patch_size, width, layers, heads, output_dim, checkpointing=False,
                 adaptor_layers=0, vision_mask=False):
        if adaptor_layers:
            raise NotImplementedError("adaptor blocks (clip.py:189-196) are not used by MiCo")
        nn.Module.__init__(self)
        if width % heads or (width // heads) % 8:
            raise NotImplementedError("head_dim must be a multiple of 8")
        self.input_resolution, self.output_dim, self.patch_size_ = input_resolution, output_dim, patch_size
        self.embed_dim = self.num_features = width
        self.num_heads = heads
        self.eps = 1e-5                       # nn.LayerNorm default (clip.py:159-166)
        scale = width ** -0.5
        self.conv1 = _Conv(width, patch_size)
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width, 1e-5)
        self.transformer = _Transformer(width, layers, heads)
        self.ln_post = LayerNorm(width, 1e-5)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))
        self.grad_checkpointing = checkpointing

        # geometry the shared launch sequences read
        class _PE:
            pass
        pe = _PE()
        pe.img_size = (input_resolution, input_resolution)
        pe.patch_size = (patch_size, patch_size)
        pe.num_patches = (input_resolution // patch_size) ** 2
        object.__setattr__(self, "patch_embed", pe)
        object.__setattr__(self, "blocks", list(self.transformer.resblocks))
        from .eva_vit import _Bf16Cache
        self._bf16 = _Bf16Cache()
        k = 3 * patch_size * patch_size
        self._kpad = (k + 63) // 64 * 64
        self._injected_dp = None
        self._act, self._act_bwd = ACT_QUICK_GELU_SAVE_GRAD, ACT_MUL_AUX
        self._full_qkv_bias = True
        self._ln_pre = True
        self.drop_path_rng = "philox"
        self._dp_rates, self._dp_calls = None, 0
        self.grad_bucket_hook = None

    def _flat_params(self):
        top = [self.class_embedding, self.positional_embedding, self.conv1.weight, None, self.ln_post.weight,
               self.ln_post.bias, self.ln_pre.weight, self.ln_pre.bias]
        for b in self.transformer.resblocks:
            top += [b.ln_1.weight, b.ln_1.bias, b.attn.in_proj_weight, b.attn.in_proj_bias, None, b.attn.out_proj.weight,
                    b.attn.out_proj.bias, b.ln_2.weight, b.ln_2.bias, b.mlp.c_fc.weight, b.mlp.c_fc.bias,
                    b.mlp.c_proj.weight, b.mlp.c_proj.bias]
        return top

    def forward(self, x, return_all_features=False):
        y = EVAVisionTransformer.forward_features(self, x, return_all_features=True)    # ln_post on every token
        if return_all_features:
            return y
        y = y[:, 0, :]
        if self.proj is not None:
            y = MF.linear_f32(y, self.proj.t())        # x @ proj (clip.py:292-293)
        return y
