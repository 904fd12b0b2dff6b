"""UNet2DCondition skeletons (SDXL-Turbo and SD-Turbo / SD2.1-base) in plain PyTorch.

`diffusers` — which supplies the UNet the reference quantizes (reference kernels/mixdq.py:4,35-41)
— is not available offline, so the architecture is restated here from the module tree the
reference ships in mixed_precision_scripts/sensitivity_log/sdxl_turbo/weight/sqnr/
bs32_split_sqnr_weight/generated_images/run.log:8-1048 (every layer with its shape) and the public
UNet configs. Module names are exactly the diffusers names, i.e. the 794 keys of the reference's
bit-width YAMLs (kernels/cfgs/weight/uniform_8.yaml), so `quantize_unet` and PTQ checkpoints apply
unchanged. It is the host of the hot path (every nn.Linear / nn.Conv2d leaf), the FP16 baseline
and the carrier of the CPU oracle — not part of the quantized hot path itself: attention, norms and
activations stay stock PyTorch fp16 ops, as in the reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class UNetConfig:
    name: str = "sdxl-turbo"
    in_channels: int = 4
    out_channels: int = 4
    sample_size: int = 64                      # latent H = W (512x512 image)
    block_out_channels: Tuple[int, ...] = (320, 640, 1280)
    down_block_types: Tuple[str, ...] = ("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D")
    up_block_types: Tuple[str, ...] = ("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D")
    layers_per_block: int = 2
    transformer_layers_per_block: Tuple[int, ...] = (1, 2, 10)
    attention_head_dim: int = 64               # channels per head
    cross_attention_dim: int = 2048
    norm_num_groups: int = 32
    addition_embed: bool = True                # SDXL text_time conditioning
    addition_time_embed_dim: int = 256
    projection_class_embeddings_input_dim: int = 2816


def sdxl_turbo_config() -> UNetConfig:
    return UNetConfig()


def sd_turbo_config() -> UNetConfig:
    """stabilityai/sd-turbo (SD2.1-base UNet) — not in the reference; public unet/config.json."""
    return UNetConfig(
        name="sd-turbo", block_out_channels=(320, 640, 1280, 1280),
        down_block_types=("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",),
        up_block_types=("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3,
        transformer_layers_per_block=(1, 1, 1, 1), cross_attention_dim=1024,
        addition_embed=False)


def tiny_config() -> UNetConfig:
    """A few-layer UNet with the SDXL topology (split shortcuts, cross-attention, up/down
    samplers) small enough for CPU tests and the smoke run."""
    return UNetConfig(
        name="tiny", sample_size=16, block_out_channels=(64, 128),
        down_block_types=("DownBlock2D", "CrossAttnDownBlock2D"),
        up_block_types=("CrossAttnUpBlock2D", "UpBlock2D"), layers_per_block=1,
        transformer_layers_per_block=(1, 1), attention_head_dim=32, cross_attention_dim=96,
        norm_num_groups=16, addition_embed=True, addition_time_embed_dim=16,
        projection_class_embeddings_input_dim=64 + 6 * 16)


def timestep_embedding(timesteps: torch.Tensor, dim: int) -> torch.Tensor:
    """Sinusoidal embedding, diffusers `Timesteps(dim, flip_sin_to_cos=True, freq_shift=0)`."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32,
                                                 device=timesteps.device) / half
    emb = timesteps.float()[:, None] * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim: int, dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(self.act(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, in_ch: int, out_ch: int, temb_ch: int, groups: int):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_ch, eps=1e-5)
        self.conv1 = nn.Conv2d(in_ch, out_ch, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_ch, out_ch)
        self.norm2 = nn.GroupNorm(groups, out_ch, eps=1e-5)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(out_ch, out_ch, 3, padding=1)
        self.nonlinearity = nn.SiLU()
        self.conv_shortcut = nn.Conv2d(in_ch, out_ch, 1) if in_ch != out_ch else None

    def forward(self, x, temb):
        h = self.conv1(self.nonlinearity(self.norm1(x)))
        t = self.time_emb_proj(self.nonlinearity(temb))
        h = h + t[:, :, None, None]
        h = self.conv2(self.dropout(self.nonlinearity(self.norm2(h))))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    def __init__(self, dim: int, ctx_dim: Optional[int], head_dim: int):
        super().__init__()
        self.heads = dim // head_dim
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(ctx_dim or dim, dim, bias=False)
        self.to_v = nn.Linear(ctx_dim or dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        b, t, c = x.shape
        q = self.to_q(x).view(b, t, self.heads, -1).transpose(1, 2)
        k = self.to_k(ctx).view(b, ctx.shape[1], self.heads, -1).transpose(1, 2)
        v = self.to_v(ctx).view(b, ctx.shape[1], self.heads, -1).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v)
        o = o.transpose(1, 2).reshape(b, t, c)
        return self.to_out[1](self.to_out[0](o))


class GEGLU(nn.Module):
    def __init__(self, dim: int, inner: int):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, ctx_dim: int, head_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, head_dim)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, ctx_dim, head_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), ctx)
        return x + self.ff(self.norm3(x))


class Transformer2DModel(nn.Module):
    def __init__(self, dim: int, ctx_dim: int, head_dim: int, layers: int, groups: int):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Linear(dim, dim)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(dim, ctx_dim, head_dim) for _ in range(layers)])
        self.proj_out = nn.Linear(dim, dim)

    def forward(self, x, ctx):
        b, c, h, w = x.shape
        res = x
        y = self.norm(x).permute(0, 2, 3, 1).reshape(b, h * w, c)
        y = self.proj_in(y)
        for blk in self.transformer_blocks:
            y = blk(y, ctx)
        y = self.proj_out(y)
        return y.reshape(b, h, w, c).permute(0, 3, 1, 2) + res


class Downsample2D(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, in_ch, out_ch, temb_ch, layers, tlayers, cross, add_down):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(in_ch if i == 0 else out_ch, out_ch, temb_ch, cfg.norm_num_groups)
             for i in range(layers)])
        if cross:
            self.attentions = nn.ModuleList(
                [Transformer2DModel(out_ch, cfg.cross_attention_dim, cfg.attention_head_dim,
                                    tlayers, cfg.norm_num_groups) for _ in range(layers)])
        else:
            self.attentions = None
        self.downsamplers = nn.ModuleList([Downsample2D(out_ch)]) if add_down else None

    def forward(self, x, temb, ctx):
        outs = []
        for i, r in enumerate(self.resnets):
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, ch, temb_ch, tlayers):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb_ch, cfg.norm_num_groups)
                                      for _ in range(2)])
        self.attentions = nn.ModuleList(
            [Transformer2DModel(ch, cfg.cross_attention_dim, cfg.attention_head_dim, tlayers,
                                cfg.norm_num_groups)])

    def forward(self, x, temb, ctx):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ctx)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, cfg: UNetConfig, in_ch, out_ch, prev_ch, temb_ch, layers, tlayers, cross,
                 add_up):
        super().__init__()
        res = []
        for i in range(layers):
            skip = in_ch if i == layers - 1 else out_ch
            rin = prev_ch if i == 0 else out_ch
            res.append(ResnetBlock2D(rin + skip, out_ch, temb_ch, cfg.norm_num_groups))
        self.resnets = nn.ModuleList(res)
        if cross:
            self.attentions = nn.ModuleList(
                [Transformer2DModel(out_ch, cfg.cross_attention_dim, cfg.attention_head_dim,
                                    tlayers, cfg.norm_num_groups) for _ in range(layers)])
        else:
            self.attentions = None
        self.upsamplers = nn.ModuleList([Upsample2D(out_ch)]) if add_up else None

    def forward(self, x, skips: List[torch.Tensor], temb, ctx):
        for i, r in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class UNet2DConditionModel(nn.Module):
    def __init__(self, cfg: UNetConfig):
        super().__init__()
        self.cfg = cfg
        self.config = cfg  # diffusers-style alias (`unet.config.in_channels`, `.sample_size`)
        ch = cfg.block_out_channels
        temb_ch = ch[0] * 4
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(ch[0], temb_ch)
        if cfg.addition_embed:
            self.add_embedding = TimestepEmbedding(cfg.projection_class_embeddings_input_dim, temb_ch)
        nb = len(ch)
        self.down_blocks = nn.ModuleList()
        out = ch[0]
        for i, t in enumerate(cfg.down_block_types):
            inp, out = out, ch[i]
            self.down_blocks.append(DownBlock(cfg, inp, out, temb_ch, cfg.layers_per_block,
                                              cfg.transformer_layers_per_block[i],
                                              t.startswith("CrossAttn"), i != nb - 1))
        self.mid_block = MidBlock(cfg, ch[-1], temb_ch, cfg.transformer_layers_per_block[-1])
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(ch))
        rev_t = list(reversed(cfg.transformer_layers_per_block))
        out = rev[0]
        for i, t in enumerate(cfg.up_block_types):
            prev, out = out, rev[i]
            inp = rev[min(i + 1, nb - 1)]
            self.up_blocks.append(UpBlock(cfg, inp, out, prev, temb_ch, cfg.layers_per_block + 1,
                                          rev_t[i], t.startswith("CrossAttn"), i != nb - 1))
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, ch[0], eps=1e-5)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)

    def forward(self, sample, timestep, encoder_hidden_states, added_cond_kwargs=None,
                text_embeds=None, time_ids=None, return_dict=False):
        cfg = self.cfg
        b = sample.shape[0]
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.float32, device=sample.device)
        t = timestep.reshape(-1).expand(b)
        emb = self.time_embedding(timestep_embedding(t, cfg.block_out_channels[0]).to(sample.dtype))
        if cfg.addition_embed:
            if added_cond_kwargs is not None:
                text_embeds = added_cond_kwargs["text_embeds"]
                time_ids = added_cond_kwargs["time_ids"]
            tid = timestep_embedding(time_ids.reshape(-1), cfg.addition_time_embed_dim)
            add = torch.cat([text_embeds, tid.reshape(b, -1).to(text_embeds.dtype)], dim=-1)
            emb = emb + self.add_embedding(add.to(sample.dtype))
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, emb, encoder_hidden_states)
            skips.extend(outs)
        x = self.mid_block(x, emb, encoder_hidden_states)
        for blk in self.up_blocks:
            x = blk(x, skips, emb, encoder_hidden_states)
        x = self.conv_out(self.conv_act(self.conv_norm_out(x)))
        return (x,)

    # ---- helpers ----------------------------------------------------------------------
    def quantizable_layers(self):
        """(name, module) of every nn.Linear / nn.Conv2d leaf — the reference's 794 YAML keys."""
        return [(n, m) for n, m in self.named_modules() if isinstance(m, (nn.Linear, nn.Conv2d))]

    def example_inputs(self, batch: int, device, dtype=torch.float16, seed: int = 0):
        """Synthetic inputs of the shapes the reference benchmarks its UNet with
        (reference kernels/mixdq.py:391-414), seeded."""
        g = torch.Generator().manual_seed(seed)
        cfg = self.cfg
        d = dict(
            sample=torch.randn(batch, cfg.in_channels, cfg.sample_size, cfg.sample_size, generator=g),
            timestep=torch.tensor(999.0),
            encoder_hidden_states=torch.randn(batch, 77, cfg.cross_attention_dim, generator=g))
        if cfg.addition_embed:
            te_dim = cfg.projection_class_embeddings_input_dim - 6 * cfg.addition_time_embed_dim
            d["text_embeds"] = torch.randn(batch, te_dim, generator=g)
            d["time_ids"] = torch.tensor([[512., 512., 0., 0., 512., 512.]]).repeat(batch, 1)
        out = {}
        for k, v in d.items():
            v = v.to(device)
            out[k] = v.to(dtype) if k != "timestep" else v
        out["sample"] = out["sample"].contiguous(memory_format=torch.channels_last)
        return out


def build_unet(name: str = "sdxl-turbo", seed: int = 0) -> UNet2DConditionModel:
    cfg = {"sdxl-turbo": sdxl_turbo_config, "sd-turbo": sd_turbo_config, "tiny": tiny_config}[name]()
    torch.manual_seed(seed)
    return UNet2DConditionModel(cfg)
