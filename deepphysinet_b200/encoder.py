"""MetaNet encoder - stays in PyTorch (BASELINE.json north_star: "The encoder stays in PyTorch").

Only here so that `PhysicsNet` keeps the reference's constructor, call surface and `state_dict`
layout: attribute names and parameter creation order follow
DeepPhysiNet/model/{meta_net.py:13-20, transformer_net.py:95-129, embed.py:16-64, attn.py:161-196},
so `torch.manual_seed(s); PhysicsNet(...)` reproduces the reference's initial weights and reference
checkpoints load unchanged.  It is NOT on the hot path and has no CUDA of its own.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .pe import SineCosPE


class _BatchedGradLinear(torch.autograd.Function):
    """y = x W^T + b for x [B, L, in] with the weight gradient computed per sample and summed: dW = sum_b G_b^T X_b.
    Same arithmetic as F.linear's backward (fp32 products, different summation order).  cuBLAS runs the single [out x B L] x [B L x in]
    product of the default backward on 16 CTAs (30 us for 256 x 3320 x 256 on a B200, 25 of them per step = 0.75 ms of an
    end-to-end step); the batched form fills the GPU."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return F.linear(x, w, b)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        gx = g.matmul(w) if ctx.needs_input_grad[0] else None
        gw = torch.bmm(g.transpose(1, 2), x).sum(0) if ctx.needs_input_grad[1] else None
        gb = g.sum((0, 1)) if ctx.needs_input_grad[2] else None
        return gx, gw, gb


def _linear(x, w, b):
    if x.is_cuda and x.dim() == 3 and x.shape[0] > 1 and b is not None and torch.is_grad_enabled() and w.requires_grad:
        return _BatchedGradLinear.apply(x, w, b)
    return F.linear(x, w, b)


class _SinusoidTable(nn.Module):
    """embed.py:16-33 - persistent buffer `pe` [1, max_len, d_model]."""

    def __init__(self, d_model, max_len=5000):
        super().__init__()
        pos = torch.arange(0, max_len).float().unsqueeze(1)
        div = (torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model)).exp()
        table = torch.zeros(max_len, d_model).float()
        table[:, 0::2] = torch.sin(pos * div)
        table[:, 1::2] = torch.cos(pos * div)
        self.register_buffer("pe", table.unsqueeze(0))

    def forward(self, length):
        return self.pe[:, :length]


class _TokenConv(nn.Module):
    """embed.py:36-49 - circular Conv1d(k=3) over the field-token axis."""

    def __init__(self, c_in, d_model):
        super().__init__()
        self.tokenConv = nn.Conv1d(c_in, d_model, kernel_size=3, padding=1, padding_mode="circular")
        nn.init.kaiming_normal_(self.tokenConv.weight, mode="fan_in", nonlinearity="leaky_relu")

    def forward(self, x):                      # x [B, tokens, c_in]
        # The circular k = 3 convolution as ONE GEMM over the three neighbouring tokens: out[l] = sum_k W[:, :, k] x[(l + k - 1) % L].
        # Same arithmetic as F.conv1d (fp32 products, different summation order); cuDNN's implicit-GEMM kernel for this shape
        # (c_in = 2405, 159 tokens) needs 0.62 ms forward + 0.20 ms weight gradient per step, 4 % of an end-to-end step.
        conv = self.tokenConv
        if conv.kernel_size != (3,) or conv.padding_mode != "circular":
            return conv(x.transpose(1, 2)).transpose(1, 2)
        x3 = torch.cat([x.roll(1, dims=1), x, x.roll(-1, dims=1)], dim=-1)               # [B, L, 3 c_in]
        wm = conv.weight.permute(2, 1, 0).reshape(3 * conv.in_channels, conv.out_channels)
        return torch.addmm(conv.bias, x3.reshape(-1, x3.shape[-1]), wm).view(x.shape[0], x.shape[1], -1)


class _FieldEmbedding(nn.Module):
    """embed.py:52-64 (attribute names incl. the reference's `time_embending` spelling)."""

    def __init__(self, c_in, d_model):
        super().__init__()
        self.value_embedding = _TokenConv(c_in, d_model)
        self.position_embedding = _SinusoidTable(d_model)
        self.time_embending = SineCosPE(1, N_freqs=d_model // 2, include_input=False)

    def forward(self, x, forecast_h, learnable_token):
        x = self.value_embedding(x)
        x = torch.cat([learnable_token.expand(x.shape[0], -1, -1), x], dim=1)
        return x + self.position_embedding(x.shape[1]) + self.time_embending(forecast_h)


class _SelfAttention(nn.Module):
    """attn.py:161-196 with FullAttention(mask_flag=False) (attn.py:43-68): softmax(QK^T/sqrt(E))V."""

    def __init__(self, d_model, n_heads):
        super().__init__()
        self.inner_attention = nn.Identity()   # parameter-free in the reference as well
        self.query_projection = nn.Linear(d_model, d_model)
        self.key_projection = nn.Linear(d_model, d_model)
        self.value_projection = nn.Linear(d_model, d_model)
        self.out_projection = nn.Linear(d_model, d_model)
        self.n_heads = n_heads

    def forward(self, x):
        B, L, D = x.shape
        H = self.n_heads
        q = _linear(x, self.query_projection.weight, self.query_projection.bias).view(B, L, H, -1).transpose(1, 2)
        k = _linear(x, self.key_projection.weight, self.key_projection.bias).view(B, L, H, -1).transpose(1, 2)
        v = _linear(x, self.value_projection.weight, self.value_projection.bias).view(B, L, H, -1).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v)
        return _linear(o.transpose(1, 2).reshape(B, L, D), self.out_projection.weight, self.out_projection.bias)


class _EncoderLayer(nn.Module):
    """transformer_net.py:17-44."""

    def __init__(self, d_model, n_heads, d_ff, activation):
        super().__init__()
        self.attention = _SelfAttention(d_model, n_heads)
        self.conv1 = nn.Conv1d(d_model, d_ff, kernel_size=1)
        self.conv2 = nn.Conv1d(d_ff, d_model, kernel_size=1)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.activation = F.relu if activation == "relu" else F.gelu

    def forward(self, x):
        x = self.norm1(x + self.attention(x))
        y = self.activation(_linear(x, self.conv1.weight.squeeze(-1), self.conv1.bias))
        y = _linear(y, self.conv2.weight.squeeze(-1), self.conv2.bias)
        return self.norm2(x + y)


class _EncoderStack(nn.Module):
    """transformer_net.py:47-72 without the (unused) distilling conv layers."""

    def __init__(self, layers, norm_layer):
        super().__init__()
        self.attn_layers = nn.ModuleList(layers)
        self.norm = norm_layer

    def forward(self, x):
        for layer in self.attn_layers:
            x = layer(x)
        return self.norm(x)


class TransformerNet(nn.Module):
    """transformer_net.py:95-129."""

    def __init__(self, enc_in, c_out, d_model=512, n_heads=8, e_layers=6, d_ff=512, activation="gelu",
                 learnable_token_num=128, output_attention=False, **kwargs):
        super().__init__()
        self.enc_embedding = _FieldEmbedding(enc_in, d_model)
        self.learnable_token = nn.Parameter(torch.rand([1, learnable_token_num, d_model]), requires_grad=True)
        self.encoder = _EncoderStack([_EncoderLayer(d_model, n_heads, d_ff, activation) for _ in range(e_layers)],
                                     nn.LayerNorm(d_model))
        self.projection = nn.Linear(d_model, c_out, bias=True)

    def forward(self, x_enc, forecast_h):
        h = self.enc_embedding(x_enc, forecast_h, self.learnable_token)
        return _linear(self.encoder(h), self.projection.weight, self.projection.bias)


class MetaNet(nn.Module):
    """meta_net.py:13-20."""

    def __init__(self, meta_cfg):
        super().__init__()
        cfg = dict(meta_cfg)
        cfg.pop("name", None)
        self.meta_cfg = meta_cfg
        self.model = TransformerNet(**cfg)

    def forward(self, x, forecast_h):
        return self.model(x, forecast_h)
