"""State-dict schema of the two CV-VAE checkpoints (key -> shape), so that the published safetensors load
unchanged:

  sd21 (vae3d, vae3d_v1-1): LDM naming  - encoder.down.N.block.N.{norm1,conv1,norm2,conv2,nin_shortcut},
        down.N.downsample.conv, mid.{block_1,attn_1,block_2}, norm_out, conv_out; decoder.up.N...
        (reference models/vae_models.py:679-788, 826-944)
  sd3  (vae3d_sd3): diffusers naming   - encoder.down_blocks.N.resnets.N.*, .downsamplers.0.conv,
        mid_block.{resnets.N,attentions.0.{group_norm,to_q,to_k,to_v,to_out.0}}, conv_norm_out, conv_out;
        decoder.up_blocks.N.resnets.N, .upsamplers.0.conv
        (reference models/vae_models3d_sd3.py:81-158, 238-319; models/vae_blocks3d_sd3.py)
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Tuple

import torch
import torch.nn as nn

from .engine import NetConfig

Shapes = "OrderedDict[str, Tuple[int, ...]]"


class _Schema:
    def __init__(self):
        self.s: Dict[str, Tuple[int, ...]] = OrderedDict()

    def conv(self, name, ci, co, k, nd):
        self.s[name + ".weight"] = (co, ci) + (k,) * nd
        self.s[name + ".bias"] = (co,)

    def vec(self, name, c):
        self.s[name + ".weight"] = (c,)
        self.s[name + ".bias"] = (c,)


def param_shapes(cfg: NetConfig):
    S = _Schema()
    w = list(cfg.widths)
    n = len(w)
    sd3 = cfg.variant == "sd3"

    def resblock(p, ci, co):
        S.vec(p + ".norm1", ci)
        S.conv(p + ".conv1", ci, co, 3, 3)
        S.vec(p + ".norm2", co)
        S.conv(p + ".conv2", co, co, 3, 2 if cfg.half_3d else 3)
        if ci != co:
            if sd3:
                S.conv(p + ".conv_shortcut", ci, co, 1, 2)
            else:
                S.conv(p + ".nin_shortcut", ci, co, 1, 3)

    def attn21(p, c, kind):
        if kind == "none":
            return
        S.vec(p + ".norm", c)
        for q in ("q", "k", "v", "proj_out"):
            S.conv(f"{p}.{q}", c, c, 1, 2)
        if kind == "spatial-temporal-xformer":
            for q in ("q_t", "k_t", "v_t", "proj_out_t"):
                S.conv(f"{p}.{q}", c, c, 1, 0)
            S.vec(p + ".norm_t", c)

    def mid3(p, c):
        if cfg.mid_block_add_attention:
            a = p + ".attentions.0"
            S.vec(a + ".group_norm", c)
            for q in ("to_q", "to_k", "to_v", "to_out.0"):
                S.conv(f"{a}.{q}", c, c, 1, 0)
        resblock(p + ".resnets.0", c, c)
        resblock(p + ".resnets.1", c, c)

    # ---------------- encoder
    S.conv("encoder.conv_in", cfg.in_channels, w[0], 3, 3)
    c = w[0]
    for i in range(n):
        for b in range(cfg.num_res_blocks):
            resblock(f"encoder.down_blocks.{i}.resnets.{b}" if sd3 else f"encoder.down.{i}.block.{b}", c, w[i])
            c = w[i]
        if i != n - 1:
            S.conv(f"encoder.down_blocks.{i}.downsamplers.0.conv" if sd3 else f"encoder.down.{i}.downsample.conv", c, c, 3, 3)
    if sd3:
        mid3("encoder.mid_block", c)
        S.vec("encoder.conv_norm_out", c)
    else:
        resblock("encoder.mid.block_1", c, c)
        attn21("encoder.mid.attn_1", c, cfg.encoder_attn_type)
        resblock("encoder.mid.block_2", c, c)
        S.vec("encoder.norm_out", c)
    S.conv("encoder.conv_out", c, cfg.moments_channels, 3, 3)
    # ---------------- decoder
    c = w[-1]
    S.conv("decoder.conv_in", cfg.z_channels, c, 3, 3)
    if not sd3:
        resblock("decoder.mid.block_1", c, c)
        attn21("decoder.mid.attn_1", c, cfg.decoder_attn_type)
        resblock("decoder.mid.block_2", c, c)
    for i in range(n):
        lvl = n - 1 - i
        for b in range(cfg.num_res_blocks + 1):
            resblock(f"decoder.up_blocks.{i}.resnets.{b}" if sd3 else f"decoder.up.{lvl}.block.{b}", c, w[lvl])
            c = w[lvl]
        if i != n - 1:
            up_time = (i % 2 == 0) if sd3 else (lvl % 2 == 1)
            S.conv(f"decoder.up_blocks.{i}.upsamplers.0.conv" if sd3 else f"decoder.up.{lvl}.upsample.conv",
                   c, c * (2 if up_time else 1), 3, 3)
    if sd3:
        mid3("decoder.mid_block", w[-1])
        S.vec("decoder.conv_norm_out", c)
    else:
        S.vec("decoder.norm_out", c)
    S.conv("decoder.conv_out", c, cfg.out_ch, 3, 3)
    return S.s


class ParamTree(nn.Module):
    """A bare module tree whose only job is to own parameters under the reference's key names
    (so state_dict()/load_state_dict()/.to()/.half() behave like the reference modules)."""


def build_param_tree(root: nn.Module, shapes) -> None:
    for key, shape in shapes.items():
        parts = key.split(".")
        mod = root
        for name in parts[:-1]:
            child = mod._modules.get(name)
            if child is None:
                child = ParamTree()
                mod.add_module(name, child)
            mod = child
        leaf = parts[-1]
        t = torch.empty(shape)
        if len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            bound = 1.0 / math.sqrt(fan_in)
            nn.init.uniform_(t, -bound, bound)
        elif leaf == "weight":
            t.fill_(1.0)  # norm scale
        else:
            t.zero_()
        mod.register_parameter(leaf, nn.Parameter(t, requires_grad=False))
