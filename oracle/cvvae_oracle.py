"""CPU oracle for the CV-VAE encode()/decode() hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``cvvae_b200``) may
import this module; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker / the timed CPU baseline, never as the thing shipped.

It is a functional restatement (state-dict in, tensor out, plain ``torch``
library calls, no ``nn.Module`` graph) of the reference's PyTorch algorithm.
Every function cites the reference ``file:line`` it follows (paths relative to
the upstream repository root).

Pinning: the reference ships no golden vectors or tests.  The
oracle is therefore pinned against outputs of the *reference itself*, executed
in the build container from ``/root/reference`` through the third-party shim in
``tests/ref_shim`` by ``tests/golden/make_golden.py``; the resulting fixtures
are committed under ``tests/golden/`` and ``tests/test_oracle_golden.py``
checks this file against them (CPU fp32: bit-exact for the networks, since the
same ATen kernels run in the same order).

Two network families are restated:

* ``sd21``  - ``models/vae_models.py``  Encoder / Decoder  (4-channel latent)
* ``sd3``   - ``models/vae_blocks3d_sd3.py`` + ``models/vae_models3d_sd3.py``
              Encoder3D / Decoder3D (16-channel latent)

plus the chunk / tile / blend wrapper of ``models/modeling_vae.py`` which is
textually identical for both (``CVVAEModel`` :20-341, ``CVVAESD3Model`` :344-667).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]
Trace = Optional[Callable[[str, Tensor], None]]


# --------------------------------------------------------------------------- #
# configuration (ctor kwargs of models/modeling_vae.py:23-51 and :347-381)
# --------------------------------------------------------------------------- #
@dataclass
class VAEConfig:
    variant: str = "sd21"  # "sd21" | "sd3"
    in_channels: int = 3
    z_channels: int = 4  # sd3: 16 ("out_channels" of CVVAESD3Model)
    ch: int = 128  # base width; block widths are ch*ch_mult
    ch_mult: Tuple[int, ...] = (1, 2, 4, 4)
    num_res_blocks: int = 2
    norm_groups: int = 32
    double_z: bool = True
    causal_encoder: bool = True
    causal_decoder: bool = False
    half_3d: bool = True
    encoder_attn_type: str = "vanilla-xformers"  # sd21 only
    decoder_attn_type: str = "spatial-temporal-xformer"  # sd21 only
    mid_block_add_attention: bool = True  # sd3 only
    en_de_n_frames_a_time: Optional[int] = 16
    time_n_compress: Optional[int] = 4
    spatial_n_compress: Optional[int] = 8
    tile_spatial_size: Optional[int] = 576
    num_video_frames: Optional[int] = None
    tile_overlap_ratio: Optional[float] = 0.2222
    reshape_z_dim_to_4: bool = False
    reshape_x_dim_to_4: bool = False

    @property
    def eps(self) -> float:
        # vae_models.py:192-195 (1e-5) vs vae_models3d_sd3.py:122,137,151 (1e-6)
        return 1e-5 if self.variant == "sd21" else 1e-6

    @property
    def widths(self) -> List[int]:
        return [self.ch * m for m in self.ch_mult]

    @property
    def moments_channels(self) -> int:
        return 2 * self.z_channels if self.double_z else self.z_channels


# --------------------------------------------------------------------------- #
# synthetic, order-independent parameters
# --------------------------------------------------------------------------- #
def _key_seed(key: str, seed: int) -> int:
    h = 1469598103934665603
    for ch in (f"{seed}:{key}").encode():
        h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h & 0x7FFFFFFFFFFFFFFF


def synth_param(key: str, shape: Sequence[int], seed: int) -> Tensor:
    """Deterministic fp32 value for one state-dict entry.

    Independent of module construction order (each key owns its generator), so
    the reference model, the oracle and the CUDA engine can all be filled with
    exactly the same numbers.  Scales follow SURVEY.md section 8d: norm weight U(0.5,1.5),
    norm bias U(-0.2,0.2), conv/linear bias U(-0.1,0.1); conv/linear weights are
    uniform with the fan-in bound PyTorch's default init uses.
    """
    g = torch.Generator().manual_seed(_key_seed(key, seed))
    shape = tuple(int(s) for s in shape)
    u = torch.rand(shape, generator=g, dtype=torch.float32)
    leaf = key.rsplit(".", 1)[-1]
    is_norm = any(t in key for t in ("norm", "group_norm"))
    if leaf == "weight" and is_norm:
        return u + 0.5
    if leaf == "bias" and is_norm:
        return u * 0.4 - 0.2
    if leaf == "bias":
        return u * 0.2 - 0.1
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    bound = 1.0 / math.sqrt(max(fan_in, 1))
    return (u * 2.0 - 1.0) * bound


def _conv3(shapes, name, ci, co, k=3):
    shapes[name + ".weight"] = (co, ci, k, k, k)
    shapes[name + ".bias"] = (co,)


def _conv2(shapes, name, ci, co, k=3):
    shapes[name + ".weight"] = (co, ci, k, k)
    shapes[name + ".bias"] = (co,)


def _norm(shapes, name, c):
    shapes[name + ".weight"] = (c,)
    shapes[name + ".bias"] = (c,)


def _linear(shapes, name, ci, co):
    shapes[name + ".weight"] = (co, ci)
    shapes[name + ".bias"] = (co,)


def _sd21_resblock_shapes(shapes, p, ci, co, half_3d):
    # vae_models.py:343-388
    _norm(shapes, p + ".norm1", ci)
    _conv3(shapes, p + ".conv1", ci, co)
    _norm(shapes, p + ".norm2", co)
    if half_3d:
        _conv2(shapes, p + ".conv2", co, co)
    else:
        _conv3(shapes, p + ".conv2", co, co)
    if ci != co:
        _conv3(shapes, p + ".nin_shortcut", ci, co, k=1)


def _sd21_attn_shapes(shapes, p, c, attn_type):
    # vae_models.py:427-444 / :473-497 / :540-571
    if attn_type == "none":
        return
    _norm(shapes, p + ".norm", c)
    for n in ("q", "k", "v", "proj_out"):
        _conv2(shapes, f"{p}.{n}", c, c, k=1)
    if attn_type == "spatial-temporal-xformer":
        for n in ("q_t", "k_t", "v_t", "proj_out_t"):
            _linear(shapes, f"{p}.{n}", c, c)
        _norm(shapes, p + ".norm_t", c)


def param_shapes(cfg: VAEConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    """State-dict key -> shape, in the key schema of the reference modules.

    sd21: vae_models.py:679-788 (Encoder), :826-944 (Decoder)
    sd3 : vae_models3d_sd3.py:81-158, :238-319 with block classes of
          vae_blocks3d_sd3.py (diffusers naming).
    """
    shapes: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    w = cfg.widths
    L = len(w)
    if cfg.variant == "sd21":
        # ---- encoder
        _conv3(shapes, "encoder.conv_in", cfg.in_channels, cfg.ch)
        cin = cfg.ch
        for lvl in range(L):
            for b in range(cfg.num_res_blocks):
                _sd21_resblock_shapes(shapes, f"encoder.down.{lvl}.block.{b}", cin, w[lvl], cfg.half_3d)
                cin = w[lvl]
            if lvl != L - 1:
                _conv3(shapes, f"encoder.down.{lvl}.downsample.conv", cin, cin)
        _sd21_resblock_shapes(shapes, "encoder.mid.block_1", cin, cin, cfg.half_3d)
        _sd21_attn_shapes(shapes, "encoder.mid.attn_1", cin, cfg.encoder_attn_type)
        _sd21_resblock_shapes(shapes, "encoder.mid.block_2", cin, cin, cfg.half_3d)
        _norm(shapes, "encoder.norm_out", cin)
        _conv3(shapes, "encoder.conv_out", cin, cfg.moments_channels)
        # ---- decoder
        cin = w[-1]
        _conv3(shapes, "decoder.conv_in", cfg.z_channels, cin)
        _sd21_resblock_shapes(shapes, "decoder.mid.block_1", cin, cin, cfg.half_3d)
        _sd21_attn_shapes(shapes, "decoder.mid.attn_1", cin, cfg.decoder_attn_type)
        _sd21_resblock_shapes(shapes, "decoder.mid.block_2", cin, cin, cfg.half_3d)
        for lvl in reversed(range(L)):
            for b in range(cfg.num_res_blocks + 1):
                _sd21_resblock_shapes(shapes, f"decoder.up.{lvl}.block.{b}", cin, w[lvl], cfg.half_3d)
                cin = w[lvl]
            if lvl != 0:
                up_time = lvl % 2 == 1
                _conv3(shapes, f"decoder.up.{lvl}.upsample.conv", cin, cin * (2 if up_time else 1))
        _norm(shapes, "decoder.norm_out", cin)
        _conv3(shapes, "decoder.conv_out", cin, cfg.in_channels)
        return shapes

    assert cfg.variant == "sd3"

    def res(p, ci, co):
        # vae_blocks3d_sd3.py:448-516
        _norm(shapes, p + ".norm1", ci)
        _conv3(shapes, p + ".conv1", ci, co)
        _norm(shapes, p + ".norm2", co)
        if cfg.half_3d:
            _conv2(shapes, p + ".conv2", co, co)
        else:
            _conv3(shapes, p + ".conv2", co, co)
        if ci != co:
            _conv2(shapes, p + ".conv_shortcut", ci, co, k=1)

    def mid(p, c):
        # vae_blocks3d_sd3.py:773-840
        if cfg.mid_block_add_attention:
            a = p + ".attentions.0"
            _norm(shapes, a + ".group_norm", c)
            for n in ("to_q", "to_k", "to_v"):
                _linear(shapes, f"{a}.{n}", c, c)
            _linear(shapes, a + ".to_out.0", c, c)
        res(p + ".resnets.0", c, c)
        res(p + ".resnets.1", c, c)

    _conv3(shapes, "encoder.conv_in", cfg.in_channels, w[0])
    cin = w[0]
    for i in range(L):
        for b in range(cfg.num_res_blocks):
            res(f"encoder.down_blocks.{i}.resnets.{b}", cin, w[i])
            cin = w[i]
        if i != L - 1:
            _conv3(shapes, f"encoder.down_blocks.{i}.downsamplers.0.conv", cin, cin)
    mid("encoder.mid_block", cin)
    _norm(shapes, "encoder.conv_norm_out", cin)
    _conv3(shapes, "encoder.conv_out", cin, cfg.moments_channels)

    rw = list(reversed(w))
    _conv3(shapes, "decoder.conv_in", cfg.z_channels, rw[0])
    cin = rw[0]
    for i in range(L):
        for b in range(cfg.num_res_blocks + 1):
            res(f"decoder.up_blocks.{i}.resnets.{b}", cin, rw[i])
            cin = rw[i]
        if i != L - 1:
            up_time = i % 2 == 0
            _conv3(shapes, f"decoder.up_blocks.{i}.upsamplers.0.conv", cin, cin * (2 if up_time else 1))
    mid("decoder.mid_block", rw[0])
    _norm(shapes, "decoder.conv_norm_out", cin)
    _conv3(shapes, "decoder.conv_out", cin, cfg.in_channels)
    return shapes


def make_state_dict(cfg: VAEConfig, seed: int = 1234, dtype=torch.float32) -> StateDict:
    return OrderedDict((k, synth_param(k, s, seed).to(dtype)) for k, s in param_shapes(cfg).items())


# --------------------------------------------------------------------------- #
# L1 ops
# --------------------------------------------------------------------------- #
def _t(trace: Trace, name: str, x: Tensor) -> Tensor:
    if trace is not None:
        trace(name, x)
    return x


def swish(x: Tensor) -> Tensor:
    """vae_models.py:187-189  (two ops: sigmoid then multiply)."""
    return x * torch.sigmoid(x)


def group_norm(x: Tensor, sd: StateDict, p: str, groups: int, eps: float) -> Tensor:
    """vae_models.py:192-195 Normalize / nn.GroupNorm in the sd3 blocks.

    5-D input: statistics over (C/groups, T, H, W) jointly per sample.
    """
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps)


def sd21_causal_conv3d(x: Tensor, sd: StateDict, p: str, pad: int, stride=1) -> Tensor:
    """vae_models.py:298-328 CausalConv3d.forward.

    H, W zero-padded by ``pad`` in fp32, T front-padded by ``2*pad`` replicating
    the first frame, cast back, then nn.Conv3d with padding 0.
    (The bfloat16 branch at :315 is dead code: dtype is fp32 there.)
    """
    ori = x.dtype
    x = x.to(torch.float32)
    x = F.pad(x, (pad, pad, pad, pad, 0, 0), mode="constant", value=0)
    x = F.pad(x, (0, 0, 0, 0, 2 * pad, 0), mode="replicate")
    x = x.to(ori)
    return F.conv3d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride)


def sd21_conv3d_zero(x: Tensor, sd: StateDict, p: str, pad: int) -> Tensor:
    """Plain nn.Conv3d(padding=pad) of the non-causal decoder (vae_models.py:361,953)."""
    return F.conv3d(x, sd[p + ".weight"], sd[p + ".bias"], padding=pad)


def conv2d_extra_dim(x: Tensor, sd: StateDict, p: str, pad: int) -> Tensor:
    """vae_models.py:331-340 / vae_blocks3d_sd3.py:107-116: fold T into batch, Conv2d, unfold."""
    b, c, t, h, w = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    y = F.conv2d(y, sd[p + ".weight"], sd[p + ".bias"], padding=pad)
    return y.reshape(b, t, y.shape[1], y.shape[2], y.shape[3]).permute(0, 2, 1, 3, 4)


def sd21_conv(x, sd, p, causal: bool, pad: int):
    """conv_cls selection of vae_models.py:361-362,714-715,952-955 (use_3d_conv=True)."""
    return sd21_causal_conv3d(x, sd, p, pad) if causal else sd21_conv3d_zero(x, sd, p, pad)


def sd21_resblock(x, sd, p, cfg: VAEConfig, causal: bool, trace: Trace = None):
    """vae_models.py:390-410 ResnetBlock3D.forward (temb is None, dropout 0)."""
    h = group_norm(x, sd, p + ".norm1", cfg.norm_groups, cfg.eps)
    h = swish(h)
    h = _t(trace, p + ".conv1", sd21_conv(h, sd, p + ".conv1", causal, 1))
    h = group_norm(h, sd, p + ".norm2", cfg.norm_groups, cfg.eps)
    h = swish(h)
    if cfg.half_3d:
        h = conv2d_extra_dim(h, sd, p + ".conv2", 1)
    else:
        h = sd21_conv(h, sd, p + ".conv2", causal, 1)
    if (p + ".nin_shortcut.weight") in sd:
        x = sd21_conv(x, sd, p + ".nin_shortcut", causal, 0)
    return _t(trace, p, x + h)


def sd21_downsample(x, sd, p, down_time: bool):
    """vae_models.py:251-263 Downsample3D.forward (with_conv=True).

    Zero pad right/bottom by 1, replicate-pad time front by 2, Conv3d stride 2 or (1,2,2).
    (bf16 is routed through fp16 for the replicate pad, :254-257 - value preserving
    only for fp16-representable magnitudes; mirrored here.)
    """
    x = F.pad(x, (0, 1, 0, 1, 0, 0), mode="constant", value=0)
    if x.dtype == torch.bfloat16:
        x = F.pad(x.to(torch.float16), (0, 0, 0, 0, 2, 0), mode="replicate").to(torch.bfloat16)
    else:
        x = F.pad(x, (0, 0, 0, 0, 2, 0), mode="replicate")
    stride = 2 if down_time else (1, 2, 2)
    return F.conv3d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], stride=stride)


def _interleave_time(x: Tensor, up_time: int) -> Tensor:
    """'b (n c) t h w -> b c (t n) h w' then drop frame 0 (vae_models.py:230-232)."""
    if up_time == 1:
        return x
    b, nc, t, h, w = x.shape
    c = nc // up_time
    x = x.reshape(b, up_time, c, t, h, w).permute(0, 2, 3, 1, 4, 5).reshape(b, c, t * up_time, h, w)
    return x[:, :, 1:]


def sd21_upsample(x, sd, p, up_time: bool, causal: bool = False):
    """vae_models.py:214-235 Upsample3D.forward.

    NB the Decoder never forwards ``causal`` (vae_models.py:936) so it is always False there.
    """
    ori = x.dtype
    if x.dtype == torch.bfloat16:
        x = x.to(torch.float16)
    x = F.interpolate(x, scale_factor=(1.0, 2.0, 2.0), mode="nearest")
    x = F.pad(x, (1, 1, 1, 1, 0, 0), mode="constant", value=0)
    if not causal:
        x = F.pad(x, (0, 0, 0, 0, 1, 1), mode="replicate")
    else:
        x = F.pad(x, (0, 0, 0, 0, 2, 0), mode="replicate")
    x = x.to(ori)
    x = F.conv3d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"])
    return _interleave_time(x, 2 if up_time else 1)


def _sdpa_tokens(q: Tensor, k: Tensor, v: Tensor) -> Tensor:
    """softmax(q k^T / sqrt(C)) v on [B, N, C]; equals both
    F.scaled_dot_product_attention (vae_models.py:456) and
    xformers.ops.memory_efficient_attention(q,k,v) (:518,:581,:607)."""
    return F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]


def sd21_spatial_attention(h_: Tensor, sd, p, cfg: VAEConfig) -> Tensor:
    """AttnBlock.attention / MemoryEfficientAttnBlock.attention on [(b t), c, h, w]
    (vae_models.py:446-461, :500-528): GN (per frame) -> q,k,v 1x1 -> attention."""
    h_ = F.group_norm(h_, cfg.norm_groups, sd[p + ".norm.weight"], sd[p + ".norm.bias"], cfg.eps)
    q = F.conv2d(h_, sd[p + ".q.weight"], sd[p + ".q.bias"])
    k = F.conv2d(h_, sd[p + ".k.weight"], sd[p + ".k.bias"])
    v = F.conv2d(h_, sd[p + ".v.weight"], sd[p + ".v.bias"])
    B, C, H, W = q.shape
    q, k, v = (t.reshape(B, C, H * W).permute(0, 2, 1).contiguous() for t in (q, k, v))
    o = _sdpa_tokens(q, k, v)
    return o.permute(0, 2, 1).reshape(B, C, H, W)


def sd21_attn(x: Tensor, sd, p, cfg: VAEConfig, attn_type: str, trace: Trace = None) -> Tensor:
    """make_attn dispatch (vae_models.py:641-676) and the block forwards
    (:463-470, :530-537, :619-629)."""
    if attn_type == "none":
        return x
    if attn_type not in ("vanilla", "vanilla-xformers", "spatial-temporal-xformer"):
        raise NotImplementedError(f"oracle: attn_type {attn_type!r} is outside the hot path")
    b, c, t, h, w = x.shape
    h_ = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    h_ = sd21_spatial_attention(h_, sd, p, cfg)
    h_ = F.conv2d(h_, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    if attn_type == "spatial-temporal-xformer":
        # :626-628  "(b t) c h w -> (b h w) t c", temporal attention, back
        h_ = h_.reshape(b, t, c, h, w).permute(0, 3, 4, 1, 2).reshape(b * h * w, t, c)
        _t(trace, p + ".spatial", h_)
        h_ = F.layer_norm(h_, (c,), sd[p + ".norm_t.weight"], sd[p + ".norm_t.bias"], 1e-5)
        q = F.linear(h_, sd[p + ".q_t.weight"], sd[p + ".q_t.bias"]).contiguous()
        k = F.linear(h_, sd[p + ".k_t.weight"], sd[p + ".k_t.bias"]).contiguous()
        v = F.linear(h_, sd[p + ".v_t.weight"], sd[p + ".v_t.bias"]).contiguous()
        o = _sdpa_tokens(q, k, v)
        o = F.linear(o, sd[p + ".proj_out_t.weight"], sd[p + ".proj_out_t.bias"])
        h_ = o.reshape(b, h, w, t, c).permute(0, 4, 3, 1, 2)
    else:
        h_ = h_.reshape(b, t, c, h, w).permute(0, 2, 1, 3, 4)
    return _t(trace, p, x + h_)


def sd21_encoder(x: Tensor, sd: StateDict, cfg: VAEConfig, trace: Trace = None) -> Tensor:
    """vae_models.py:790-823 Encoder.forward."""
    causal = cfg.causal_encoder
    L = len(cfg.ch_mult)
    h = _t(trace, "encoder.conv_in", sd21_conv(x, sd, "encoder.conv_in", causal, 1))
    for lvl in range(L):
        for b in range(cfg.num_res_blocks):
            h = sd21_resblock(h, sd, f"encoder.down.{lvl}.block.{b}", cfg, causal, trace)
        if lvl != L - 1:
            h = _t(trace, f"encoder.down.{lvl}.downsample",
                   sd21_downsample(h, sd, f"encoder.down.{lvl}.downsample", down_time=(lvl % 2 == 0)))
    h = sd21_resblock(h, sd, "encoder.mid.block_1", cfg, causal, trace)
    h = sd21_attn(h, sd, "encoder.mid.attn_1", cfg, cfg.encoder_attn_type, trace)
    h = sd21_resblock(h, sd, "encoder.mid.block_2", cfg, causal, trace)
    h = group_norm(h, sd, "encoder.norm_out", cfg.norm_groups, cfg.eps)
    h = _t(trace, "encoder.norm_out", swish(h))
    return _t(trace, "encoder.conv_out", sd21_conv(h, sd, "encoder.conv_out", causal, 1))


def sd21_decoder(z: Tensor, sd: StateDict, cfg: VAEConfig, trace: Trace = None) -> Tensor:
    """vae_models.py:960-1002 Decoder.forward."""
    causal = cfg.causal_decoder
    L = len(cfg.ch_mult)
    h = _t(trace, "decoder.conv_in", sd21_conv(z, sd, "decoder.conv_in", causal, 1))
    h = sd21_resblock(h, sd, "decoder.mid.block_1", cfg, causal, trace)
    h = sd21_attn(h, sd, "decoder.mid.attn_1", cfg, cfg.decoder_attn_type, trace)
    h = sd21_resblock(h, sd, "decoder.mid.block_2", cfg, causal, trace)
    for lvl in reversed(range(L)):
        for b in range(cfg.num_res_blocks + 1):
            h = sd21_resblock(h, sd, f"decoder.up.{lvl}.block.{b}", cfg, causal, trace)
        if lvl != 0:
            h = _t(trace, f"decoder.up.{lvl}.upsample",
                   sd21_upsample(h, sd, f"decoder.up.{lvl}.upsample", up_time=(lvl % 2 == 1)))
    h = group_norm(h, sd, "decoder.norm_out", cfg.norm_groups, cfg.eps)
    h = _t(trace, "decoder.norm_out", swish(h))
    return _t(trace, "decoder.conv_out", sd21_conv(h, sd, "decoder.conv_out", causal, 1))


# ----------------------------- sd3 family ---------------------------------- #
def sd3_conv3d(x, sd, p, causal: bool, pad: int, stride=1):
    """vae_blocks3d_sd3.py:81-104 (CausalConv3d: replicate pad (p,p,p,p,2p,0), bf16 via fp32)
    and :16-46 (Conv3d with padding_mode='replicate')."""
    if pad > 0:
        dt = x.dtype
        if dt == torch.bfloat16:
            x = x.to(torch.float32)
        tp = (2 * pad, 0) if causal else (pad, pad)
        x = F.pad(x, (pad, pad, pad, pad) + tp, mode="replicate")
        if dt == torch.bfloat16:
            x = x.to(dt)
    return F.conv3d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride)


def sd3_resblock(x, sd, p, cfg: VAEConfig, causal: bool, trace: Trace = None):
    """vae_blocks3d_sd3.py:518-569 ResnetBlock3D.forward (temb None, no up/down, scale 1)."""
    h = group_norm(x, sd, p + ".norm1", cfg.norm_groups, cfg.eps)
    h = F.silu(h)
    h = _t(trace, p + ".conv1", sd3_conv3d(h, sd, p + ".conv1", causal, 1))
    h = group_norm(h, sd, p + ".norm2", cfg.norm_groups, cfg.eps)
    h = F.silu(h)
    if cfg.half_3d:
        h = conv2d_extra_dim(h, sd, p + ".conv2", 1)
    else:
        h = sd3_conv3d(h, sd, p + ".conv2", causal, 1)
    if (p + ".conv_shortcut.weight") in sd:
        x = conv2d_extra_dim(x, sd, p + ".conv_shortcut", 0)
    return _t(trace, p, (x + h) / 1.0)


def sd3_attention(x, sd, p, cfg: VAEConfig, trace: Trace = None):
    """AttentionWithExtraDim (vae_blocks3d_sd3.py:119-147) over diffusers ``Attention``
    with AttnProcessor2_0 semantics for 4-D input (third-party, unpinned version; restated
    per SURVEY.md section 8c): residual; view(B,C,HW); GroupNorm(groups, eps) on [B,C,N];
    Linear q,k,v (bias); SDPA 1 head, scale C^-0.5; to_out[0]; back to [B,C,H,W];
    + residual; / rescale_output_factor (=1)."""
    b, c, t, h, w = x.shape
    hs = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    res = hs
    y = hs.reshape(b * t, c, h * w)
    y = F.group_norm(y, cfg.norm_groups, sd[p + ".group_norm.weight"], sd[p + ".group_norm.bias"], cfg.eps)
    y = y.transpose(1, 2)
    q = F.linear(y, sd[p + ".to_q.weight"], sd[p + ".to_q.bias"])
    k = F.linear(y, sd[p + ".to_k.weight"], sd[p + ".to_k.bias"])
    v = F.linear(y, sd[p + ".to_v.weight"], sd[p + ".to_v.bias"])
    o = _sdpa_tokens(q, k, v)
    o = F.linear(o, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])
    o = o.transpose(-1, -2).reshape(b * t, c, h, w)
    o = (o + res) / 1.0
    return _t(trace, p, o.reshape(b, t, c, h, w).permute(0, 2, 1, 3, 4))


def sd3_mid(x, sd, p, cfg, causal, trace=None):
    """vae_blocks3d_sd3.py:842-856 UNetMidBlock3D.forward."""
    x = sd3_resblock(x, sd, p + ".resnets.0", cfg, causal, trace)
    if cfg.mid_block_add_attention:
        x = sd3_attention(x, sd, p + ".attentions.0", cfg, trace)
    return sd3_resblock(x, sd, p + ".resnets.1", cfg, causal, trace)


def sd3_upsample(x, sd, p, up_time: bool, causal: bool):
    """vae_blocks3d_sd3.py:314-364 Upsample3D.forward (norm None, interpolate True)."""
    dt = x.dtype
    if dt == torch.bfloat16:
        x = x.to(torch.float32)
    x = F.interpolate(x, scale_factor=(1.0, 2.0, 2.0), mode="nearest")
    if dt == torch.bfloat16:
        x = x.to(dt)
    x = sd3_conv3d(x, sd, p + ".conv", causal, 1)
    return _interleave_time(x, 2 if up_time else 1)


def sd3_encoder(x, sd, cfg: VAEConfig, trace: Trace = None):
    """vae_models3d_sd3.py:162-208 Encoder3D.forward (eval path)."""
    causal = cfg.causal_encoder
    L = len(cfg.ch_mult)
    h = _t(trace, "encoder.conv_in", sd3_conv3d(x, sd, "encoder.conv_in", causal, 1))
    for i in range(L):
        for b in range(cfg.num_res_blocks):
            h = sd3_resblock(h, sd, f"encoder.down_blocks.{i}.resnets.{b}", cfg, causal, trace)
        if i != L - 1:
            down_time = i % 2 == 0
            stride = 2 if down_time else (1, 2, 2)
            # Downsample3D.forward :224-239 -> conv_cls(stride, padding=1) :200-210
            h = _t(trace, f"encoder.down_blocks.{i}.downsamplers.0",
                   sd3_conv3d(h, sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", causal, 1, stride))
    h = sd3_mid(h, sd, "encoder.mid_block", cfg, causal, trace)
    h = group_norm(h, sd, "encoder.conv_norm_out", cfg.norm_groups, cfg.eps)
    h = _t(trace, "encoder.conv_norm_out", F.silu(h))
    return _t(trace, "encoder.conv_out", sd3_conv3d(h, sd, "encoder.conv_out", causal, 1))


def sd3_decoder(z, sd, cfg: VAEConfig, trace: Trace = None):
    """vae_models3d_sd3.py:323-388 Decoder3D.forward (eval path, latent_embeds None)."""
    causal = cfg.causal_decoder
    L = len(cfg.ch_mult)
    h = _t(trace, "decoder.conv_in", sd3_conv3d(z, sd, "decoder.conv_in", causal, 1))
    h = sd3_mid(h, sd, "decoder.mid_block", cfg, causal, trace)
    for i in range(L):
        for b in range(cfg.num_res_blocks + 1):
            h = sd3_resblock(h, sd, f"decoder.up_blocks.{i}.resnets.{b}", cfg, causal, trace)
        if i != L - 1:
            h = _t(trace, f"decoder.up_blocks.{i}.upsamplers.0",
                   sd3_upsample(h, sd, f"decoder.up_blocks.{i}.upsamplers.0", up_time=(i % 2 == 0), causal=causal))
    h = group_norm(h, sd, "decoder.conv_norm_out", cfg.norm_groups, cfg.eps)
    h = _t(trace, "decoder.conv_norm_out", F.silu(h))
    return _t(trace, "decoder.conv_out", sd3_conv3d(h, sd, "decoder.conv_out", causal, 1))


def encoder_forward(x, sd, cfg: VAEConfig, trace: Trace = None):
    return sd21_encoder(x, sd, cfg, trace) if cfg.variant == "sd21" else sd3_encoder(x, sd, cfg, trace)


def decoder_forward(z, sd, cfg: VAEConfig, trace: Trace = None):
    return sd21_decoder(z, sd, cfg, trace) if cfg.variant == "sd21" else sd3_decoder(z, sd, cfg, trace)


# --------------------------------------------------------------------------- #
# L3 wrapper: chunk / tile / blend  (models/modeling_vae.py)
# --------------------------------------------------------------------------- #
def blend_h(a: Tensor, b: Tensor, ov: int) -> Tensor:
    """modeling_vae.py:321-330 - IN PLACE on b; weights fp32 arange(ov)/ov."""
    wb = (torch.arange(ov).view(1, 1, 1, 1, -1) / ov).to(b.device)
    b[:, :, :, :, :ov] = (1 - wb) * a[:, :, :, :, -ov:] + wb * b[:, :, :, :, :ov]
    return b


def blend_v(a: Tensor, b: Tensor, ov: int) -> Tensor:
    """modeling_vae.py:332-341 - IN PLACE on b."""
    wb = (torch.arange(ov).view(1, 1, 1, -1, 1) / ov).to(b.device)
    b[:, :, :, :ov, :] = (1 - wb) * a[:, :, :, -ov:, :] + wb * b[:, :, :, :ov, :]
    return b


@dataclass
class TileGeometry:
    """Derived attributes of modeling_vae.py:84-112 and the rounding of :148-150,:234-236."""
    encode_chunk: Optional[int]
    decode_chunk: Optional[int]
    pixel_tile: Optional[int]
    latent_tile: Optional[int]
    ratio: Optional[float]

    @staticmethod
    def of(cfg: VAEConfig) -> "TileGeometry":
        if cfg.en_de_n_frames_a_time is not None:
            assert cfg.time_n_compress is not None
            assert cfg.en_de_n_frames_a_time % cfg.time_n_compress == 0
            ec, dc = cfg.en_de_n_frames_a_time, cfg.en_de_n_frames_a_time // cfg.time_n_compress
        else:
            ec = dc = None
        if cfg.tile_spatial_size is not None:
            assert cfg.spatial_n_compress is not None and cfg.tile_overlap_ratio is not None
            pt, lt, r = cfg.tile_spatial_size, cfg.tile_spatial_size // cfg.spatial_n_compress, cfg.tile_overlap_ratio
        else:
            pt = lt = r = None
        return TileGeometry(ec, dc, pt, lt, r)


def _spatial_tiled(x: Tensor, fn, tile: Optional[int], other_tile: Optional[int], ratio, encode: bool) -> Tensor:
    """modeling_vae.py:144-191 (encode) / :230-277 (decode)."""
    if tile is None:
        return fn(x)
    # encode: :148-150 (tile=pixel, other=latent); decode: :234-236 (tile=latent, other=pixel)
    in_stride = round(tile * (1 - ratio))
    out_overlap = round(other_tile * ratio)
    out_stride = other_tile - out_overlap
    rows = []
    for i in range(0, x.shape[3], in_stride):
        cols = []
        for j in range(0, x.shape[4], in_stride):
            cols.append(fn(x[:, :, :, i:i + tile, j:j + tile]))
            if j + tile >= x.shape[4]:
                break
        rows.append(cols)
        if i + tile >= x.shape[3]:
            break
    res_rows = []
    for i, cols in enumerate(rows):
        res_cols = []
        for j, t in enumerate(cols):
            if i > 0:
                t = blend_v(rows[i - 1][j], t, out_overlap)
            if j > 0:
                t = blend_h(cols[j - 1], t, out_overlap)
            res_cols.append(t)
        res_rows.append(res_cols)
    out = []
    for i, cols in enumerate(res_rows):
        for j, t in enumerate(cols):
            if i < len(res_rows) - 1:
                t = t[:, :, :, :out_stride, :]
            if j < len(cols) - 1:
                t = t[:, :, :, :, :out_stride]
            cols[j] = t
        out.append(torch.cat(cols, dim=4))
    return torch.cat(out, dim=3)


def tiled_encode(x: Tensor, sd, cfg: VAEConfig) -> Tensor:
    """modeling_vae.py:193-210."""
    g = TileGeometry.of(cfg)
    enc = lambda v: encoder_forward(v, sd, cfg)
    sp = lambda v: _spatial_tiled(v, enc, g.pixel_tile, g.latent_tile, g.ratio, True)
    if g.encode_chunk is None:
        return sp(x)
    assert x.dim() == 5
    stride = g.encode_chunk
    n_rounds = math.ceil((x.shape[2] - 1) / stride) or 1
    zs = []
    for n in range(n_rounds):
        z = sp(x[:, :, n * stride:(n + 1) * stride + 1])
        zs.append(z if n == 0 else z[:, :, 1:])
    return torch.cat(zs, dim=2)


def tiled_decode(z: Tensor, sd, cfg: VAEConfig) -> Tensor:
    """modeling_vae.py:279-296."""
    g = TileGeometry.of(cfg)
    dec = lambda v: decoder_forward(v, sd, cfg)
    sp = lambda v: _spatial_tiled(v, dec, g.latent_tile, g.pixel_tile, g.ratio, False)
    if g.decode_chunk is None:
        return sp(z)
    assert z.dim() == 5
    stride = g.decode_chunk
    n_rounds = math.ceil((z.shape[2] - 1) / stride) or 1
    xs = []
    for n in range(n_rounds):
        x = sp(z[:, :, n * stride:(n + 1) * stride + 1])
        xs.append(x if n == 0 else x[:, :, 1:])
    return torch.cat(xs, dim=2)


@dataclass
class Posterior:
    """diffusers DiagonalGaussianDistribution (third-party; twin:
    lvdm/modules/distributions/distributions.py:24-73)."""
    parameters: Tensor
    mean: Tensor = field(init=False)
    logvar: Tensor = field(init=False)
    std: Tensor = field(init=False)
    var: Tensor = field(init=False)

    def __post_init__(self):
        self.mean, self.logvar = torch.chunk(self.parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def mode(self) -> Tensor:
        return self.mean

    def sample(self, generator=None) -> Tensor:
        eps = torch.randn(self.mean.shape, generator=generator, dtype=self.mean.dtype, device=self.mean.device)
        return self.mean + self.std * eps


def encode(x: Tensor, sd, cfg: VAEConfig) -> Posterior:
    """modeling_vae.py:212-228."""
    if x.dim() == 4:
        if cfg.num_video_frames is not None:
            bt, c, h, w = x.shape
            t = cfg.num_video_frames
            x = x.reshape(bt // t, t, c, h, w).permute(0, 2, 1, 3, 4)
        else:
            x = x[:, :, None]
    return Posterior(tiled_encode(x, sd, cfg))


def decode(z: Tensor, sd, cfg: VAEConfig, num_frames: Optional[int] = None) -> Tensor:
    """modeling_vae.py:298-319."""
    if z.dim() == 4:
        t = num_frames
        if t is None and cfg.num_video_frames is not None:
            t = 1 + (cfg.num_video_frames - 1) // cfg.time_n_compress
        if t is not None:
            bt, c, h, w = z.shape
            z = z.reshape(bt // t, t, c, h, w).permute(0, 2, 1, 3, 4)
        else:
            z = z[:, :, None]
    x = tiled_decode(z, sd, cfg)
    if cfg.reshape_x_dim_to_4:
        b, c, t, h, w = x.shape
        x = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    return x


def forward(x: Tensor, sd, cfg: VAEConfig, sample_posterior=False, generator=None, num_frames=None) -> Tensor:
    """modeling_vae.py:114-142."""
    post = encode(x, sd, cfg)
    z = post.sample(generator) if sample_posterior else post.mode()
    return decode(z, sd, cfg, num_frames=num_frames)


# --------------------------------------------------------------------------- #
# helpers shared by tests / bench
# --------------------------------------------------------------------------- #
def synthetic_video(shape: Sequence[int], seed: int = 0) -> Tensor:
    """uniform [-1,1] fp32 on CPU from a seeded generator."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(tuple(shape), generator=g, dtype=torch.float32) * 2.0 - 1.0
