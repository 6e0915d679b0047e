"""GPU input / label pipeline against the reference fixture and the oracle: on-device Gaussian ground
truth bit-for-bit (utils/mpii_data.py:165-181), uint8 images through forward_uint8 == forward on the normalised
fp32 tensor (mpii_data.py:184-185)."""
import os
import warnings

import numpy as np
import pytest
import torch

from oracle import evaluate_oracle as E
from oracle import unipose_oracle as O

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_gaussian_labels_bit_exact_vs_reference_fixture():
    from unipose_b200 import data
    g = np.load(os.path.join(GOLDEN, "labels_mpii.npz"))
    kpts, center = E.synth_keypoints(4, 16, 368, 368, seed=40)
    heat = data.gaussian_heatmaps(torch.from_numpy(kpts).cuda(), 368, 368, stride=8, sigma=3.0).cpu().numpy()
    cm = data.center_map(torch.from_numpy(center).cuda(), 368, 368, stride=8, sigma=3.0).cpu().numpy()
    assert heat.shape == (4, 17, 46, 46) and cm.shape == (4, 1, 46, 46)
    assert np.array_equal(heat, g["heat"]), np.abs(heat - g["heat"]).max()
    assert np.array_equal(cm, g["centermap"])


@pytest.mark.parametrize("n,k,size,stride,sigma", [(8, 14, 256, 8, 3.0), (3, 13, 368, 8, 3.0), (2, 17, 512, 4, 2.0)])
def test_gaussian_labels_vs_oracle(n, k, size, stride, sigma):
    from unipose_b200 import data
    kpts, center = E.synth_keypoints(n, k, size, size, seed=n + k)
    heat = data.gaussian_heatmaps(torch.from_numpy(kpts).cuda(), size, size, stride=stride, sigma=sigma).cpu().numpy()
    cm = data.center_map(torch.from_numpy(center).cuda(), size, size, stride=stride, sigma=3.0).cpu().numpy()
    bad = 0
    for b in range(n):
        rh, rc = E.reference_labels(kpts[b], center[b], size, size, stride, sigma)
        # CUDA's and glibc's double exp are both < 1 ulp but not correctly rounded: after the fp32 store a difference
        # needs a double-rounding coincidence - allow 1 fp32 ulp on a vanishing fraction, everything else bit-exact
        d = np.abs(heat[b].astype(np.float64) - rh)
        assert d.max() <= 6e-8, d.max()
        bad += int((heat[b] != rh).sum())
        assert np.array_equal(cm[b], rc)
    assert bad <= 2, bad


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_forward_uint8_equals_forward_on_normalised_input(precision):
    from unipose_b200.model.unipose import unipose
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = unipose(dataset="MPII", num_classes=16, precision=precision)
    m.load_state_dict(O.synth_state_dict(16, seed=0))
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, (2, 128, 128, 3), generator=g, dtype=torch.uint8)
    x = (u8.permute(0, 3, 1, 2).float() - 128.0) / 256.0            # Mytransforms.to_tensor + normalize
    a = m(x.cuda())
    b = m.forward_uint8(u8.cuda())
    assert torch.equal(a, b)
