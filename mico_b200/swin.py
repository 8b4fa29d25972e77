"""Swin Transformer (video frame encoder of BASELINE config 4) on the sm_100a kernels: mirror of the reference's
``model/swin.py`` (``SwinTransformer`` :485-611, ``BasicLayer`` :355-434, ``SwinTransformerBlock`` :175-312,
``WindowAttention`` :77-156, ``PatchMerging`` :315-352, ``PatchEmbed`` :437-475).  Same constructor keywords,
``state_dict`` keys (including the ``relative_position_index`` / ``attn_mask`` buffers) and ``forward_features`` contract
``(B,3,224,224) -> (B, 49, 8*embed_dim)``.

Compute: 4x4 patch embedding = im2col kernel + tcgen05 GEMM; every Linear (qkv, proj, fc1+GELU, fc2+residual, patch-merge
reduction) = tcgen05 GEMM; LayerNorm = K2 kernel; window attention = the fused attention kernel with a 4-D additive bias
``[nW, heads, 49, 49]`` = relative position bias (+ shifted-window mask) addressed per (window, head) inside the kernel
(mico_attention_*, mask_bmod / mask_hs); the bias is a parameter, its gradient comes from mico_attention_dmask.
The cyclic shift, window partition / reverse and the 2x2 patch-merge gather are index permutations done with torch views
and copies on bf16/fp32 tensors (folding them into the kernels' loads is future work); DropPath uses
per-sample multipliers.  Dropout (drop_rate / attn_drop_rate) must be 0, as in swin_base_patch4_window7_224_22k.yaml.
"""
import torch
import torch.nn as nn

from . import functional as MF
from . import ops
from .ops import BF16, F32, MicoError


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


def window_partition(x, window_size):
    B, H, W, C = x.shape
    x = x.view(B, H // window_size, window_size, W // window_size, window_size, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, window_size, window_size, C)


def window_reverse(windows, window_size, H, W):
    B = int(windows.shape[0] / (H * W / window_size / window_size))
    x = windows.view(B, H // window_size, W // window_size, window_size, window_size, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


class _Linear(nn.Module):
    def __init__(self, i, o, bias=True):
        super().__init__()
        self.weight = nn.Parameter(nn.init.trunc_normal_(torch.empty(o, i), std=.02, a=-2.0, b=2.0))
        self.bias = nn.Parameter(torch.zeros(o)) if bias else None


class LayerNorm(nn.Module):
    def __init__(self, d, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(d))
        self.bias = nn.Parameter(torch.zeros(d))

    def forward(self, x, out_dtype=F32):
        return MF.layer_norm(x, self.weight, self.bias, self.eps, out_dtype)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = _Linear(in_features, hidden_features)
        self.fc2 = _Linear(hidden_features, in_features)

    def forward(self, x, residual=None):
        a = MF.linear_gelu(x, self.fc1.weight, self.fc1.bias)
        return MF.linear_tc(a, self.fc2.weight, self.fc2.bias, residual=residual)


class WindowAttention(nn.Module):
    def __init__(self, dim, window_size, num_heads, qkv_bias=True, qk_scale=None):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, window_size, num_heads
        head_dim = dim // num_heads
        if head_dim % 8:
            raise NotImplementedError("head_dim must be a multiple of 8")
        self.scale = qk_scale or head_dim ** -0.5
        self.relative_position_bias_table = nn.Parameter(nn.init.trunc_normal_(
            torch.zeros((2 * window_size[0] - 1) * (2 * window_size[1] - 1), num_heads), std=.02, a=-2.0, b=2.0))
        coords = torch.stack(torch.meshgrid([torch.arange(window_size[0]), torch.arange(window_size[1])], indexing="ij"))
        cf = torch.flatten(coords, 1)
        rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
        rel[:, :, 0] += window_size[0] - 1
        rel[:, :, 1] += window_size[1] - 1
        rel[:, :, 0] *= 2 * window_size[1] - 1
        self.register_buffer("relative_position_index", rel.sum(-1))
        self.qkv = _Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = _Linear(dim, dim)

    def forward(self, x, mask=None):
        """x: bf16 (num_windows*B, N, C); mask: (nW, N, N) 0 / -100 or None  ->  fp32 (num_windows*B, N, C)"""
        B_, N, C = x.shape
        H = self.num_heads
        qkv = MF.linear_tc(x, self.qkv.weight, self.qkv.bias, out_dtype=BF16).view(B_, N, 3, H, C // H)
        bias = self.relative_position_bias_table[self.relative_position_index.view(-1)].view(N, N, -1)
        bias = bias.permute(2, 0, 1).unsqueeze(0)                        # 1, nH, N, N  (swin.py:135-139)
        if mask is not None:
            bias = bias + mask.unsqueeze(1)                              # nW, nH, N, N (swin.py:141-144)
        o = MF.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], self.scale, bias.float().contiguous())
        return MF.linear_tc(o.view(B_, N, C), self.proj.weight, self.proj.bias)


class SwinTransformerBlock(nn.Module):
    def __init__(self, dim, input_resolution, num_heads, window_size=7, shift_size=0, mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop_path=0.):
        super().__init__()
        self.dim, self.input_resolution, self.num_heads = dim, input_resolution, num_heads
        self.window_size, self.shift_size = window_size, shift_size
        if min(self.input_resolution) <= self.window_size:
            self.shift_size = 0
            self.window_size = min(self.input_resolution)
        assert 0 <= self.shift_size < self.window_size, "shift_size must in 0-window_size"
        self.norm1 = LayerNorm(dim)
        self.attn = WindowAttention(dim, to_2tuple(self.window_size), num_heads, qkv_bias, qk_scale)
        self.drop_prob = float(drop_path)
        self.norm2 = LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        attn_mask = None
        if self.shift_size > 0:      # swin.py:227-246
            H, W = self.input_resolution
            img_mask = torch.zeros((1, H, W, 1))
            sl = (slice(0, -self.window_size), slice(-self.window_size, -self.shift_size), slice(-self.shift_size, None))
            cnt = 0
            for h in sl:
                for w in sl:
                    img_mask[:, h, w, :] = cnt
                    cnt += 1
            mw = window_partition(img_mask, self.window_size).view(-1, self.window_size * self.window_size)
            attn_mask = mw.unsqueeze(1) - mw.unsqueeze(2)
            attn_mask = attn_mask.masked_fill(attn_mask != 0, float(-100.0)).masked_fill(attn_mask == 0, float(0.0))
        self.register_buffer("attn_mask", attn_mask)

    def _drop_path(self, x):
        """timm DropPath: per-sample Bernoulli(1-p) / (1-p) in training (swin.py:215)"""
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        m = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * m.div_(keep)

    def forward(self, x):
        H, W = self.input_resolution
        B, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        shortcut = x
        x = self.norm1(x, BF16).view(B, H, W, C)
        if self.shift_size > 0:
            x = torch.roll(x, shifts=(-self.shift_size, -self.shift_size), dims=(1, 2))
        ws = self.window_size
        xw = window_partition(x, ws).view(-1, ws * ws, C)
        aw = self.attn(xw, mask=self.attn_mask).view(-1, ws, ws, C)
        x = window_reverse(aw, ws, H, W)
        if self.shift_size > 0:
            x = torch.roll(x, shifts=(self.shift_size, self.shift_size), dims=(1, 2))
        x = shortcut + self._drop_path(x.view(B, H * W, C))
        if self.drop_prob == 0.0 or not self.training:
            return self.mlp(self.norm2(x, BF16), residual=x)             # residual add in the fc2 epilogue
        return x + self._drop_path(self.mlp(self.norm2(x, BF16)))


class PatchMerging(nn.Module):
    def __init__(self, input_resolution, dim):
        super().__init__()
        self.input_resolution, self.dim = input_resolution, dim
        self.reduction = _Linear(4 * dim, 2 * dim, bias=False)
        self.norm = LayerNorm(4 * dim)

    def forward(self, x):
        H, W = self.input_resolution
        B, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        assert H % 2 == 0 and W % 2 == 0, f"x size ({H}*{W}) are not even."
        x = x.view(B, H, W, C)
        x = torch.cat([x[:, 0::2, 0::2, :], x[:, 1::2, 0::2, :], x[:, 0::2, 1::2, :], x[:, 1::2, 1::2, :]], -1)
        x = self.norm(x.view(B, -1, 4 * C), BF16)
        return MF.linear_tc(x, self.reduction.weight, None)


class BasicLayer(nn.Module):
    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio, qkv_bias, qk_scale, drop_path,
                 downsample, use_checkpoint=False):
        super().__init__()
        self.use_checkpoint = use_checkpoint
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(dim, input_resolution, num_heads, window_size, 0 if (i % 2 == 0) else window_size // 2,
                                 mlp_ratio, qkv_bias, qk_scale, drop_path[i] if isinstance(drop_path, list) else drop_path)
            for i in range(depth)])
        self.downsample = downsample(input_resolution, dim) if downsample is not None else None

    def forward(self, x):
        for blk in self.blocks:
            x = torch.utils.checkpoint.checkpoint(blk, x, use_reentrant=False) if self.use_checkpoint else blk(x)
        return self.downsample(x) if self.downsample is not None else x


class _PatchEmbedFn(torch.autograd.Function):
    """Conv2d(k = s = patch) as im2col + GEMM (swin.py:469); pixels get no gradient."""

    @staticmethod
    def forward(ctx, x, weight, bias, patch, kpad):
        cols = ops.patchify(x.contiguous().float(), patch, kpad)
        wb = ops.cast_bf16_2d(weight.detach().reshape(weight.shape[0], -1), kpad)
        y = ops.gemm(cols, wb, bias=bias.detach(), out_dtype=F32)
        ctx.save_for_backward(cols)
        ctx.k = weight[0].numel()
        ctx.wshape = weight.shape
        return y.view(x.shape[0], -1, weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        (cols,) = ctx.saved_tensors
        dyb = ops.scale_cast_bf16(dy.reshape(-1, dy.shape[-1]).contiguous().float())
        dw = ops.gemm(dyb, cols[:, :ctx.k], a_mn=True, b_mn=True, out_dtype=F32).view(ctx.wshape)
        return None, dw, ops.colsum(dyb), None, None


class PatchEmbed(nn.Module):
    class _Proj(nn.Module):
        def __init__(self, c, d, p):
            super().__init__()
            self.weight = nn.Parameter(torch.empty(d, c, p, p).normal_(0.0, 0.02))
            self.bias = nn.Parameter(torch.zeros(d))

    def __init__(self, img_size=224, patch_size=4, in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        self.img_size, self.patch_size = to_2tuple(img_size), to_2tuple(patch_size)
        self.patches_resolution = [self.img_size[0] // self.patch_size[0], self.img_size[1] // self.patch_size[1]]
        self.num_patches = self.patches_resolution[0] * self.patches_resolution[1]
        self.in_chans, self.embed_dim = in_chans, embed_dim
        self.proj = PatchEmbed._Proj(in_chans, embed_dim, self.patch_size[0])
        self.norm = norm_layer(embed_dim) if norm_layer is not None else None
        k = in_chans * self.patch_size[0] * self.patch_size[1]
        self._kpad = (k + 63) // 64 * 64

    def forward(self, x):
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        x = _PatchEmbedFn.apply(x, self.proj.weight, self.proj.bias, self.patch_size[0], self._kpad)
        return self.norm(x) if self.norm is not None else x


class SwinTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=4, in_chans=3, num_classes=1000, embed_dim=96, depths=[2, 2, 6, 2],
                 num_heads=[3, 6, 12, 24], window_size=7, mlp_ratio=4., qkv_bias=True, qk_scale=None, drop_rate=0.,
                 attn_drop_rate=0., drop_path_rate=0.1, norm_layer=None, ape=False, patch_norm=True, use_checkpoint=False,
                 fused_window_process=False, **kwargs):
        super().__init__()
        if drop_rate or attn_drop_rate:
            raise NotImplementedError("Swin dropout is 0 in the MiCo configuration; the kernels have no dropout")
        self.num_classes, self.num_layers, self.embed_dim = num_classes, len(depths), embed_dim
        self.ape, self.patch_norm = ape, patch_norm
        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))
        self.mlp_ratio = mlp_ratio
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim, LayerNorm if patch_norm else None)
        pr = self.patch_embed.patches_resolution
        self.patches_resolution = pr
        if self.ape:
            self.absolute_pos_embed = nn.Parameter(nn.init.trunc_normal_(
                torch.zeros(1, self.patch_embed.num_patches, embed_dim), std=.02, a=-2.0, b=2.0))
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, sum(depths), device="cpu")]
        self.layers = nn.ModuleList()
        for i in range(self.num_layers):
            self.layers.append(BasicLayer(int(embed_dim * 2 ** i), (pr[0] // (2 ** i), pr[1] // (2 ** i)), depths[i],
                                          num_heads[i], window_size, mlp_ratio, qkv_bias, qk_scale,
                                          dpr[sum(depths[:i]):sum(depths[:i + 1])],
                                          PatchMerging if (i < self.num_layers - 1) else None, use_checkpoint))
        self.norm = LayerNorm(self.num_features)

    def no_weight_decay(self):
        return {'absolute_pos_embed'}

    def no_weight_decay_keywords(self):
        return {'relative_position_bias_table'}

    def forward_features(self, x):
        if not x.is_cuda:
            raise MicoError("mico_b200 runs on CUDA (sm_100a) only")
        x = self.patch_embed(x)
        if self.ape:
            x = x + self.absolute_pos_embed
        for layer in self.layers:
            x = layer(x)
        return self.norm(x)

    def encode_audio(self, x):
        return self.forward_features(x.repeat(1, 3, 1, 1).clone())

    def forward(self, x):
        return self.forward_features(x)
