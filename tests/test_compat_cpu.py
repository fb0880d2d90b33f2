"""Zero-edit drop-in (north_star: "cvvae_inference_video.py and the SD pipelines call it unchanged").

The reference's inference script is executed VERBATIM from /root/reference (build container only) with `compat/` in
front of it on sys.path, so that its `from models.modeling_vae import CVVAEModel` resolves to this package.  The three
packages the script needs that are absent from the image (decord, fire) or touch the file system (write_video) are
stubbed; the engine runs on the CPU test double of the operator set (tests/fake_ops.py) and `.cuda()` is a no-op -
what is checked here is the host-side contract: from_pretrained(path, subfolder=, torch_dtype=), requires_grad_, the
fp16 input convention, encode(...).latent_dist.sample(), decode(...).sample, shapes and value range.
"""
import importlib
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("CVVAE_REFERENCE_ROOT", "/root/reference")


@pytest.fixture
def compat_path(monkeypatch):
    monkeypatch.syspath_prepend(os.path.join(ROOT, "compat"))
    for name in [n for n in sys.modules if n == "models" or n.startswith("models.") or n.startswith("diffuser_engine")]:
        monkeypatch.delitem(sys.modules, name)
    yield


def test_compat_paths_resolve_to_the_engine(compat_path):
    import cvvae_b200
    m1 = importlib.import_module("models.modeling_vae")
    m2 = importlib.import_module("diffuser_engine.models.modeling_vae")
    assert m1.CVVAEModel is cvvae_b200.CVVAEModel and m1.CVVAESD3Model is cvvae_b200.CVVAESD3Model
    assert m2.CVVAEModel is cvvae_b200.CVVAEModel


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "cvvae_inference_video.py")), reason="reference checkout not present")
def test_reference_inference_script_runs_unchanged(compat_path, monkeypatch, tmp_path):
    from fake_ops import FakeOps
    from oracle import cvvae_oracle as O
    import cvvae_b200
    # a small "published checkpoint": config.json + safetensors under <path>/vae3d, as the script expects
    wrap = dict(tile_spatial_size=72, en_de_n_frames_a_time=4)
    m = cvvae_b200.CVVAEModel(ch=32, **wrap)
    m.load_state_dict(O.make_state_dict(O.VAEConfig(variant="sd21", ch=32, **wrap), 1234))
    m.save_pretrained(str(tmp_path / "ckpt" / "vae3d"))
    monkeypatch.setattr(cvvae_b200.modeling_vae._CVVAEBase, "_ops_factory", FakeOps)
    # stubs of what the image lacks: decord (video reader), fire (CLI), and the file write
    n_frames, H0, W0 = 7, 60, 90                              # 7 frames -> frame_end = 5
    g = torch.Generator().manual_seed(3)
    frames = torch.randint(0, 256, (n_frames, H0, W0, 3), generator=g, dtype=torch.uint8)

    class _Batch:
        def asnumpy(self):
            return frames.numpy()

    class VideoReader:
        def __init__(self, path, ctx=None):
            pass

        def __len__(self):
            return n_frames

        def get_avg_fps(self):
            return 8.0

        def get_batch(self, idx):
            return _Batch()

    decord = types.ModuleType("decord")
    decord.VideoReader, decord.cpu = VideoReader, (lambda i=0: None)
    fire = types.ModuleType("fire")
    fire.Fire = lambda fn: None
    monkeypatch.setitem(sys.modules, "decord", decord)
    monkeypatch.setitem(sys.modules, "fire", fire)
    written = {}
    import torchvision.io
    monkeypatch.setattr(torchvision.io, "write_video", lambda path, arr, fps=None, options=None: written.update(path=path, arr=arr, fps=fps),
                        raising=False)   # (torchvision >= 0.26 no longer ships write_video; the script still imports it)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)
    torch.manual_seed(0)
    src = open(os.path.join(REF, "cvvae_inference_video.py")).read()
    ns = {"__name__": "reference_script"}
    exec(compile(src, os.path.join(REF, "cvvae_inference_video.py"), "exec"), ns)   # verbatim, unmodified
    assert ns["CVVAEModel"] is cvvae_b200.CVVAEModel
    ns["main"](str(tmp_path / "ckpt"), "in.mp4", str(tmp_path / "out" / "rec.mp4"), height=80, width=104)
    arr = written["arr"]
    assert written["fps"] == 8.0 and arr.dtype == torch.uint8 and tuple(arr.shape) == (5, 80, 104, 3)
    assert 0 < arr.float().mean().item() < 255
