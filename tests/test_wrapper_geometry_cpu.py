"""Property tests (CPU, no compute kernels): the wrapper's chunk / tile / blend / crop / batching structure equals the
reference's loops (oracle `_spatial_tiled` + chunk loop, modeling_vae.py:144-210,230-296) for arbitrary sizes, with a
position-encoding stand-in for the networks so that any mis-ordered, mis-cropped or mis-batched tile shows up."""
import pytest
import torch
from hypothesis import given, settings, strategies as st

from cvvae_b200 import CVVAEModel
from fake_ops import FakeOps
from oracle import cvvae_oracle as O


def _standin(scale_t, scale_s, channels):
    """Deterministic 'network': output depends on the tile CONTENT only (so identical windows give identical outputs,
    like the real networks), resampled to the output geometry of an encoder (down) or decoder (up)."""
    def fn(v):
        B, C, T, H, W = v.shape
        if scale_s < 1:                      # encoder: T -> 1+(T-1)/4, H,W -> /8
            t = v[:, :1, ::4, ::8, ::8]
        else:                                # decoder: T -> 4(T-1)+1, H,W -> x8
            t = v[:, :1].repeat_interleave(4, 2)[:, :, 3:].repeat_interleave(8, 3).repeat_interleave(8, 4)
        return torch.cat([t * (k + 1) + 0.01 * k for k in range(channels)], dim=1).contiguous()
    return fn


@settings(max_examples=25, deadline=None)
@given(frames=st.integers(0, 9), h8=st.integers(2, 30), w8=st.integers(2, 30), batch=st.integers(1, 2),
       tile=st.sampled_from([64, 96]), chunk=st.sampled_from([4, 8]))
def test_tiled_encode_decode_structure(frames, h8, w8, batch, tile, chunk):
    T, H, W = 1 + 4 * frames, 8 * h8, 8 * w8
    wrap = dict(tile_spatial_size=tile, en_de_n_frames_a_time=chunk)
    m = CVVAEModel(ch=32, **wrap)
    m._ops_factory = FakeOps                 # blend runs through the operator test double; networks are stand-ins
    cfg = O.VAEConfig(variant="sd21", ch=32, **wrap)
    g = O.TileGeometry.of(cfg)
    x = torch.rand((batch, 3, T, H, W), generator=torch.Generator().manual_seed(T * 1000 + H + W))

    def ref_chunks(v, stride, fn):
        outs = []
        for n in range(max(1, -(-(v.shape[2] - 1) // stride))):
            o = fn(v[:, :, n * stride:(n + 1) * stride + 1])
            outs.append(o if n == 0 else o[:, :, 1:])
        return torch.cat(outs, dim=2)

    enc = _standin(0.25, 0.125, 8)
    object.__setattr__(m, "_run_net", lambda which, v: (enc if which == "encode" else dec)(v))
    dec = _standin(4, 8, 3)
    got = m.tiled_encode(x)
    want = ref_chunks(x, g.encode_chunk, lambda v: O._spatial_tiled(v, enc, g.pixel_tile, g.latent_tile, g.ratio, True))
    assert got.shape == want.shape          # (the reference's crop arithmetic does not always give W/8 columns)
    assert torch.equal(got, want)
    z = got[:, :4].contiguous()
    got = m.tiled_decode(z)
    want = ref_chunks(z, g.decode_chunk, lambda v: O._spatial_tiled(v, dec, g.latent_tile, g.pixel_tile, g.ratio, False))
    assert got.shape == want.shape
    assert torch.equal(got, want)


def test_usable_cores_is_sane():
    import os
    import bench
    n = bench.usable_cores()
    assert 1 <= n <= (os.cpu_count() or 1)
