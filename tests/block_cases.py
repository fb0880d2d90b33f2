"""Block-level parity cases shared by the CPU (test double, fp32) and GPU (kernels, fp16) suites: one engine method per
reference block type against the oracle's restatement of that block (oracle/cvvae_oracle.py cites the reference lines),
on a width-64 model so that every channel transition of the real networks occurs (64->128 with a 1x1 shortcut, 256->256,
up/down-sampling in space and time, both attention flavours, the sd3 replicate-padded variants)."""
import torch

from cvvae_b200.engine import Act, Engine, NetConfig, prepack_params
from oracle import cvvae_oracle as O

CH = 64


def setup(variant, ops, dtype, device):
    wrap = dict(tile_spatial_size=None, en_de_n_frames_a_time=None)
    cfg = O.VAEConfig(variant=variant, ch=CH, **wrap) if variant == "sd21" else O.VAEConfig(variant="sd3", ch=CH, z_channels=16, **wrap)
    sd = O.make_state_dict(cfg, 1234)
    net = NetConfig(variant=variant, z_channels=cfg.z_channels, widths=tuple(cfg.widths))
    sdd = {k: v.to(device) for k, v in sd.items()}
    eng = Engine(net, prepack_params(sdd, ops, dtype), ops, dtype)
    eng._begin_pass(1, torch.device(device))
    return eng, cfg, sd


def _x(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g)


# name -> (variant, input [B,C,T,H,W], engine call, oracle call)
def cases():
    W = CH
    c = {}
    c["sd21_resblock_shortcut_causal"] = ("sd21", (1, W, 5, 20, 24),
        lambda e, a: e.resblock(a, "encoder.down.1.block.0", True),
        lambda x, sd, cfg: O.sd21_resblock(x, sd, "encoder.down.1.block.0", cfg, True))
    c["sd21_resblock_plain_noncausal"] = ("sd21", (2, 4 * W, 3, 12, 12),
        lambda e, a: e.resblock(a, "decoder.up.3.block.1", False),
        lambda x, sd, cfg: O.sd21_resblock(x, sd, "decoder.up.3.block.1", cfg, False))
    c["sd21_resblock_shortcut_noncausal"] = ("sd21", (1, 4 * W, 3, 16, 16),
        lambda e, a: e.resblock(a, "decoder.up.1.block.0", False),
        lambda x, sd, cfg: O.sd21_resblock(x, sd, "decoder.up.1.block.0", cfg, False))
    c["sd21_downsample_time"] = ("sd21", (1, W, 5, 20, 26),
        lambda e, a: e.downsample(a, 0, True),
        lambda x, sd, cfg: O.sd21_downsample(x, sd, "encoder.down.0.downsample", True))
    c["sd21_downsample_space"] = ("sd21", (1, 2 * W, 3, 18, 14),
        lambda e, a: e.downsample(a, 1, True),
        lambda x, sd, cfg: O.sd21_downsample(x, sd, "encoder.down.1.downsample", False))
    c["sd21_upsample_time"] = ("sd21", (1, 4 * W, 3, 10, 12),
        lambda e, a: e.upsample(a, "decoder.up.3.upsample.conv", 2, False),
        lambda x, sd, cfg: O.sd21_upsample(x, sd, "decoder.up.3.upsample", True))
    c["sd21_upsample_space"] = ("sd21", (1, 4 * W, 3, 10, 12),
        lambda e, a: e.upsample(a, "decoder.up.2.upsample.conv", 1, False),
        lambda x, sd, cfg: O.sd21_upsample(x, sd, "decoder.up.2.upsample", False))
    c["sd21_attn_encoder"] = ("sd21", (1, 4 * W, 3, 9, 8),
        lambda e, a: e.attn_sd21(a, "encoder.mid.attn_1", "vanilla-xformers"),
        lambda x, sd, cfg: O.sd21_attn(x, sd, "encoder.mid.attn_1", cfg, "vanilla-xformers"))
    c["sd21_attn_decoder_spatial_temporal"] = ("sd21", (1, 4 * W, 3, 8, 8),
        lambda e, a: e.attn_sd21(a, "decoder.mid.attn_1", "spatial-temporal-xformer"),
        lambda x, sd, cfg: O.sd21_attn(x, sd, "decoder.mid.attn_1", cfg, "spatial-temporal-xformer"))
    c["sd3_resblock_shortcut_causal"] = ("sd3", (1, W, 5, 20, 24),
        lambda e, a: e.resblock(a, "encoder.down_blocks.1.resnets.0", True),
        lambda x, sd, cfg: O.sd3_resblock(x, sd, "encoder.down_blocks.1.resnets.0", cfg, True))
    c["sd3_resblock_noncausal"] = ("sd3", (1, 4 * W, 3, 12, 12),
        lambda e, a: e.resblock(a, "decoder.up_blocks.0.resnets.1", False),
        lambda x, sd, cfg: O.sd3_resblock(x, sd, "decoder.up_blocks.0.resnets.1", cfg, False))
    c["sd3_downsample_time"] = ("sd3", (1, W, 5, 20, 26),
        lambda e, a: e.downsample(a, 0, True),
        lambda x, sd, cfg: O.sd3_conv3d(x, sd, "encoder.down_blocks.0.downsamplers.0.conv", True, 1, stride=2))
    c["sd3_upsample_time"] = ("sd3", (1, 4 * W, 3, 10, 12),
        lambda e, a: e.upsample(a, "decoder.up_blocks.0.upsamplers.0.conv", 2, False),
        lambda x, sd, cfg: O.sd3_upsample(x, sd, "decoder.up_blocks.0.upsamplers.0", True, False))
    c["sd3_attention"] = ("sd3", (1, 4 * W, 3, 8, 9),
        lambda e, a: e.attn_sd3(a, "encoder.mid_block.attentions.0"),
        lambda x, sd, cfg: O.sd3_attention(x, sd, "encoder.mid_block.attentions.0", cfg))
    return c


def run_case(name, ops, dtype, device):
    variant, shape, eng_fn, ora_fn = cases()[name]
    eng, cfg, sd = setup(variant, ops, dtype, device)
    x = _x(shape, 7)
    with torch.no_grad():
        want = ora_fn(x, sd, cfg)                                   # fp32, NCDHW
        a = Act(x.to(dtype).to(device).permute(0, 2, 3, 4, 1).contiguous())
        got = eng_fn(eng, a).t.permute(0, 4, 1, 2, 3).float().cpu()
    return got, want
