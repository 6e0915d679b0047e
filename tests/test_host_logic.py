"""CPU tests of the host-side mirror of the reference interface (no kernels launched)."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from unipose_b200.model.modules.backbone.resnet import stem_s2d_weight, stem_superpixel_weight
from unipose_b200.model.unipose import unipose
from unipose_b200.model import uniposeLSTM

from conftest import GOLDEN


@pytest.fixture(scope="module")
def keys():
    return json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))


def test_image_model_state_dict_identical_to_reference(keys):
    with pytest.warns(UserWarning):   # offline: no ImageNet checkpoint in the hub cache
        m = unipose(dataset="MPII", num_classes=16)
    sd = m.state_dict()
    assert list(sd.keys()) == list(keys["image_mpii_keys"].keys())
    for k, v in sd.items():
        assert list(v.shape) == keys["image_mpii_keys"][k], k
    assert sum(p.numel() for p in m.parameters()) == 47547313


def test_video_model_state_dict_identical_to_reference(keys):
    with pytest.warns(UserWarning):
        m = uniposeLSTM.unipose(num_classes=13)
    sd = m.state_dict()
    assert list(sd.keys()) == list(keys["video_keys"].keys())
    for k, v in sd.items():
        assert list(v.shape) == keys["video_keys"][k], k


def test_reference_error_conventions():
    from unipose_b200.model.modules.backbone import build_backbone
    from unipose_b200.model.modules.wasp import build_wasp
    import torch.nn as nn
    with pytest.raises(NotImplementedError):
        build_backbone('xception', 16, nn.BatchNorm2d)
    with pytest.raises(NotImplementedError):
        build_wasp('resnet', 32, nn.BatchNorm2d)


def test_no_cpu_fallback():
    with pytest.warns(UserWarning):
        m = unipose(dataset="MPII", num_classes=16).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 3, 64, 64))


def test_stem_space_to_depth_weights_are_equivalent():
    torch.manual_seed(0)
    w = torch.randn(8, 3, 7, 7, dtype=torch.float64)
    x = torch.randn(2, 3, 32, 48, dtype=torch.float64)
    ref = F.conv2d(x, w, stride=2, padding=3)
    # 2x2 space-to-depth with channel order (ph, pw, c), padded to 16 channels
    n, c, h, wd = x.shape
    x2 = x.view(n, c, h // 2, 2, wd // 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(n, 12, h // 2, wd // 2)
    x2 = torch.cat([x2, x2.new_zeros(n, 4, h // 2, wd // 2)], 1)
    w2 = stem_s2d_weight(w)
    got = F.conv2d(F.pad(x2, (2, 1, 2, 1)), w2)
    assert torch.allclose(got, ref, atol=1e-12)


def test_stem_superpixel_weights_are_equivalent():
    """The 7x7/s2 stem as a plain 4x2 conv over 64-element super pixels (4 s2d pixels x 16 channels) that emits four
    adjacent output pixels per position - the layout resnet._emit_image feeds to the tcgen05 kernel."""
    torch.manual_seed(1)
    co = 8
    w = torch.randn(co, 3, 7, 7, dtype=torch.float64)
    x = torch.randn(2, 3, 32, 48, dtype=torch.float64)
    ref = F.conv2d(x, w, stride=2, padding=3)                         # [2, co, 16, 24]
    n, c, h, wd = x.shape
    x2 = x.view(n, c, h // 2, 2, wd // 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(n, 12, h // 2, wd // 2)
    x2 = torch.cat([x2, x2.new_zeros(n, 4, h // 2, wd // 2)], 1)      # [n, 16, h/2, w/2]
    ws = wd // 8
    rows = x2.new_zeros(n, h // 2, (ws + 1) * 4, 16)                  # NHWC rows, 2 zero pixels of left padding
    rows[:, :, 2:2 + wd // 2] = x2.permute(0, 2, 3, 1)
    sp = rows.reshape(n, h // 2, ws + 1, 64).permute(0, 3, 1, 2)      # super pixels as channels-first [n,64,h/2,ws+1]
    got = F.conv2d(F.pad(sp, (0, 0, 2, 1)), stem_superpixel_weight(w))   # [n, 4*co, h/2, ws]
    got = got.view(n, 4, co, h // 2, ws).permute(0, 2, 3, 4, 1).reshape(n, co, h // 2, wd // 2)
    assert torch.allclose(got, ref, atol=1e-12)


def test_precision_switch_and_lr_groups():
    with pytest.warns(UserWarning):
        m = unipose(dataset="MPII", num_classes=16)
    m.set_precision("bf16")
    assert m.backbone.layer3[5].precision == "bf16"
    with pytest.raises(ValueError):
        m.set_precision("int8")
    n1 = sum(p.numel() for p in m.get_1x_lr_params())
    n10 = sum(p.numel() for p in m.get_10x_lr_params())
    assert n1 == 42500160 and n1 + n10 == 47547313


def test_act_reshaped_aliases_storage():
    """The stem's output is written as [n, h, w/4, 256] and read by the max-pool as [n, h, w, 64]: same bytes."""
    from unipose_b200 import ops
    a = ops.Act(2, 4, 8, 64, ops.UP_FP16, torch.device("cpu"))
    a.t.copy_(torch.arange(a.t.numel(), dtype=torch.float32).view_as(a.t) % 251)
    b = a.reshaped(4, 2, 256)
    assert (b.n, b.h, b.w, b.c) == (2, 4, 2, 256) and b.t.data_ptr() == a.t.data_ptr()
    assert torch.equal(b.t.reshape(-1), a.t.reshape(-1)) and b.plane_stride == a.plane_stride
    with pytest.raises(AssertionError):
        a.reshaped(4, 8, 32)


def test_kernel_table_tool_reads_the_committed_ncu_capture():
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "kernel_table.py"),
                          os.path.join(root, "profiles", "wasp_block_ncu_metrics_r1.csv")],
                         capture_output=True, text=True, check=True).stdout
    assert "conv_tcgen05_kernel<1>" in out and "global_avgpool_kernel<0>" in out and "total" in out


def test_stem_weight_regroupings_as_gathers_match_their_slice_copy_definitions():
    """The stem's filter regroupings run as one gather through a cached index map; the map must reproduce the
    slice-copy definitions exactly (including the structural zeros)."""
    from unipose_b200.model.modules.backbone import resnet as R
    w = torch.randn(64, 3, 7, 7)
    assert torch.equal(R.stem_s2d_weight(w), R._stem_s2d_loops(w))
    assert torch.equal(R.stem_superpixel_weight(w), R._stem_superpixel_loops(w))
    w2 = torch.randn(8, 3, 7, 7)              # another co: its own index map
    assert torch.equal(R.stem_superpixel_weight(w2), R._stem_superpixel_loops(w2))
    assert R.stem_window_weight(w).shape == (64, 64, 4, 1)
