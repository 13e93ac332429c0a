"""Median filter + uint8 frames against the reference-generated golden and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import vx_oracle as O

pytestmark = pytest.mark.gpu


def test_median3d_matches_reference_golden(golden_dir):
    from vexpress_b200 import ops
    m = torch.load(os.path.join(golden_dir, "prologue_small.pt"), weights_only=False)["median"]
    v = torch.rand(*m["shape"], generator=torch.Generator().manual_seed(m["seed_input"]))
    frames, filt = ops.median3d_u8(v.cuda(), want_filtered=True)
    assert torch.equal(filt.cpu(), m["filtered"])
    assert np.array_equal(frames.cpu().numpy(), m["uint8"].numpy())


@pytest.mark.parametrize("T,H,W", [(2, 2, 2), (16, 64, 48), (5, 129, 67)])
def test_median3d_matches_oracle(T, H, W):
    from vexpress_b200 import ops
    v = torch.rand(3, T, H, W, generator=torch.Generator().manual_seed(T * H + W))
    v[:, :, ::3] = v[:, :, ::3].round()                      # plenty of ties
    frames, filt = ops.median3d_u8(v.cuda(), want_filtered=True)
    ref = O.median_filter_3d(v, 3)
    assert torch.equal(filt.cpu(), ref)
    assert np.array_equal(frames.cpu().numpy(), O.video_to_uint8(ref))
