"""Block-level wiring of the engine (ResnetBlock3D with / without the fused 1x1 shortcut, Down/Upsample3D, the attention
blocks; sd21 and sd3 variants) against the oracle's restatement of each reference block, through the fp32 test double."""
import pytest
import torch

import block_cases as BC
from fake_ops import FakeOps


@pytest.mark.parametrize("name", sorted(BC.cases()))
def test_block_matches_reference_block(name):
    got, want = BC.run_case(name, FakeOps(), torch.float32, "cpu")
    assert got.shape == want.shape
    torch.testing.assert_close(got, want, rtol=2e-4, atol=2e-4)
