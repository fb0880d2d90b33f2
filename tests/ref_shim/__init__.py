"""Third-party shim so that /root/reference/models/*.py import VERBATIM.

Used only by ``tests/golden/make_golden.py`` (run in the build container, where
``/root/reference`` exists) to produce the committed golden fixtures.  It
supplies the handful of ``diffusers`` / ``xformers`` symbols the reference's hot
path touches (SURVEY.md section 8c lists them and their semantics); neither package is
installed and there is no network.  Nothing here is reference code.
"""
from __future__ import annotations

import functools
import inspect
import json
import os
import sys
import types
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = os.environ.get("CVVAE_REFERENCE_ROOT", "/root/reference")


class _Config(dict):
    __getattr__ = dict.__getitem__


def register_to_config(init):
    @functools.wraps(init)
    def wrapper(self, *args, **kwargs):
        sig = inspect.signature(init)
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
        object.__setattr__(self, "_shim_config", _Config(cfg))
        init(self, *args, **kwargs)

    return wrapper


class ConfigMixin:
    config_name = "config.json"

    @property
    def config(self):
        return self._shim_config


class ModelMixin(nn.Module):
    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    @classmethod
    def from_pretrained(cls, path, subfolder=None, torch_dtype=None, **_):
        from safetensors.torch import load_file

        d = os.path.join(path, subfolder) if subfolder else path
        with open(os.path.join(d, cls.config_name)) as f:
            cfg = {k: v for k, v in json.load(f).items() if not k.startswith("_")}
        m = cls(**cfg)
        m.load_state_dict(load_file(os.path.join(d, "diffusion_pytorch_model.safetensors")))
        if torch_dtype is not None:
            m = m.to(torch_dtype)
        return m.eval()


def apply_forward_hook(fn):
    return fn


@dataclass
class DecoderOutput:
    sample: torch.Tensor


@dataclass
class AutoencoderKLOutput:
    latent_dist: object


class DiagonalGaussianDistribution:
    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator=None):
        eps = torch.randn(self.mean.shape, generator=generator, dtype=self.mean.dtype, device=self.mean.device)
        return self.mean + self.std * eps

    def mode(self):
        return self.mean


def get_activation(name):
    assert name in ("silu", "swish"), name
    return nn.SiLU()


class RMSNorm(nn.Module):  # imported by the reference, unused at default config
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError


class SpatialNorm(nn.Module):  # idem
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError


class Attention(nn.Module):
    """diffusers Attention + AttnProcessor2_0 for 4-D input, single group-normed self-attention."""

    def __init__(self, query_dim, heads=8, dim_head=64, eps=1e-5, norm_num_groups=None,
                 residual_connection=False, bias=False, rescale_output_factor=1.0,
                 spatial_norm_dim=None, upcast_softmax=False, _from_deprecated_attn_block=False, **_):
        super().__init__()
        assert spatial_norm_dim is None
        inner = heads * dim_head
        self.heads = heads
        self.residual_connection = residual_connection
        self.rescale_output_factor = rescale_output_factor
        self.group_norm = nn.GroupNorm(norm_num_groups, query_dim, eps=eps, affine=True) if norm_num_groups else None
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(query_dim, inner, bias=bias)
        self.to_v = nn.Linear(query_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Dropout(0.0)])

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, **_):
        assert encoder_hidden_states is None and attention_mask is None
        residual = hidden_states
        b, c, h, w = hidden_states.shape
        x = hidden_states.view(b, c, h * w).transpose(1, 2)
        if self.group_norm is not None:
            x = self.group_norm(x.transpose(1, 2)).transpose(1, 2)
        q, k, v = self.to_q(x), self.to_k(x), self.to_v(x)
        hd = q.shape[-1] // self.heads
        q, k, v = (t.view(b, -1, self.heads, hd).transpose(1, 2) for t in (q, k, v))
        x = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)
        x = x.transpose(1, 2).reshape(b, -1, self.heads * hd).to(q.dtype)
        x = self.to_out[1](self.to_out[0](x))
        x = x.transpose(-1, -2).reshape(b, c, h, w)
        if self.residual_connection:
            x = x + residual
        return x / self.rescale_output_factor


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Place the stub packages in sys.modules and put the reference root on sys.path."""
    if "diffusers" in sys.modules and getattr(sys.modules["diffusers"], "_cvvae_shim", False):
        return
    log = types.SimpleNamespace(get_logger=lambda name: __import__("logging").getLogger(name))
    _mod("diffusers", _cvvae_shim=True)
    _mod("diffusers.configuration_utils", ConfigMixin=ConfigMixin, register_to_config=register_to_config)
    _mod("diffusers.models")
    _mod("diffusers.models.modeling_utils", ModelMixin=ModelMixin)
    _mod("diffusers.utils", deprecate=lambda *a, **k: None, is_torch_version=lambda op, v: True,
         logging=log, BaseOutput=object)
    _mod("diffusers.utils.accelerate_utils", apply_forward_hook=apply_forward_hook)
    _mod("diffusers.utils.torch_utils", randn_tensor=lambda shape, generator=None, device=None, dtype=None:
         torch.randn(shape, generator=generator, device=device, dtype=dtype))
    _mod("diffusers.models.autoencoders")
    _mod("diffusers.models.autoencoders.vae", DiagonalGaussianDistribution=DiagonalGaussianDistribution,
         DecoderOutput=DecoderOutput)
    _mod("diffusers.models.modeling_outputs", AutoencoderKLOutput=AutoencoderKLOutput)
    _mod("diffusers.models.activations", get_activation=get_activation)
    _mod("diffusers.models.downsampling", RMSNorm=RMSNorm)
    _mod("diffusers.models.attention_processor", Attention=Attention, SpatialNorm=SpatialNorm)

    def mea(q, k, v, attn_bias=None, op=None):
        return F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]

    xf = _mod("xformers", __version__="0.0.16")
    xf.ops = _mod("xformers.ops", memory_efficient_attention=mea)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))
