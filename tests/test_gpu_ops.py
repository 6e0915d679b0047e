"""GPU parity tests of the individual kernels through the C-ABI (ctypes): tcgen05 implicit-GEMM conv in
every configuration the model uses, and the bandwidth kernels, each against an fp64/fp32 torch evaluation
of the same operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _cases():
    from tools import gpu_ladder
    return list(range(len(gpu_ladder.CASES)))


@pytest.mark.parametrize("idx", _cases())
def test_conv_case(idx):
    from tools import gpu_ladder
    res = gpu_ladder.run_case(idx)
    assert res["ok"], res


@pytest.mark.parametrize("prec", ["bf16", "fp16", "fp32"])
def test_layout_roundtrip_and_pools(prec):
    from unipose_b200 import ops
    dev = torch.device("cuda:0")
    mode = ops.mode_of(prec)
    torch.manual_seed(0)
    x = torch.randn(3, 70, 23, 30, device=dev)
    a = ops.Act(3, 23, 30, 128, mode, dev)
    ops.nchw_to_act(x, a)
    back = ops.act_to_nchw(a, 70)
    tol = {"bf16": 2 ** -8, "fp16": 2 ** -11, "fp32": 2 ** -16}[prec]
    assert (back - x).abs().max() <= tol * x.abs().max()
    assert float(a.to_float()[..., 70:].abs().max()) == 0.0     # padded channels are zero
    xq = back                                                    # what the kernels see
    # max-pool 3/2/1 into a channel slice of a wider buffer
    y = ops.Act(3, 12, 15, 192, mode, dev, zero=True)
    ops.maxpool3x3s2(a, ops.View(y, coff=64, c=128))
    ref = F.max_pool2d(xq, 3, 2, 1)
    got = y.to_float()[..., 64:64 + 70].permute(0, 3, 1, 2)
    assert (got - ref).abs().max() <= tol * ref.abs().max()
    assert float(y.to_float()[..., :64].abs().max()) == 0.0
    # bilinear align_corners up-sample
    u = ops.Act(3, 46, 61, 128, mode, dev)
    ops.upsample_bilinear_ac(a, u)
    ref = F.interpolate(xq, size=(46, 61), mode="bilinear", align_corners=True)
    got = u.to_float()[..., :70].permute(0, 3, 1, 2)
    assert (got - ref).abs().max() <= 2 * tol * ref.abs().max() + 1e-6
    # global average pool + broadcast
    g = ops.Act(3, 1, 1, 128, mode, dev)
    ops.global_avgpool(a, g)
    ref = xq.mean(dim=(2, 3))
    got = g.to_float()[:, 0, 0, :70]
    assert (got - ref).abs().max() <= tol * ref.abs().max() + 1e-6
    bc = ops.Act(3, 5, 7, 128, mode, dev)
    ops.broadcast_hw(g, bc)
    assert torch.equal(bc.t[:, :, 2, 3], g.t[:, :, 0, 0])


def test_pack_input_s2d_matches_definition():
    from unipose_b200 import ops
    dev = torch.device("cuda:0")
    x = torch.randn(2, 3, 32, 48, device=dev)
    a = ops.Act(2, 16, 24 + 3, 16, ops.UP_SPLIT, dev, zero=True)
    ops.pack_input_s2d(x, a, wpad_left=2)
    ref = x.view(2, 3, 16, 2, 24, 2).permute(0, 2, 4, 3, 5, 1).reshape(2, 16, 24, 12)
    got = a.to_float()
    assert (got[:, :, 2:26, :12] - ref).abs().max() < 1e-4
    assert float(got[..., 12:].abs().max()) == 0.0
    assert float(got[:, :, :2].abs().max()) == 0.0 and float(got[:, :, 26:].abs().max()) == 0.0


@pytest.mark.parametrize("prec,tol", [("bf16", 2 ** -7), ("fp32", 1e-4)])
def test_stem_as_overlapping_window_conv(prec, tol):
    """7x7/s2 stem (resnet.py:61) = 4-row conv over 64-element windows of the row-padded space-to-depth image."""
    from unipose_b200 import ops
    from unipose_b200.model.modules.backbone.resnet import stem_window_weight
    dev = torch.device("cuda:0")
    mode = ops.mode_of(prec)
    torch.manual_seed(3)
    n, h, w = 3, 64, 96
    x = torch.randn(n, 3, h, w, device=dev)
    wt = torch.randn(64, 3, 7, 7, device=dev) / (3 * 49) ** 0.5
    x2 = ops.Act(n, h // 2, w // 2 + 3, 16, mode, dev, zero=True)
    ops.pack_input_s2d(x, x2, wpad_left=2)
    pc = ops.make_packed_conv(stem_window_weight(wt), mode, cout=64, cin=64)
    y = ops.Act(n, h // 2, w // 2, 64, mode, dev)
    ops.conv2d(x2, pc, y, pad=(2, 0), relu=False, ho=h // 2, wo=w // 2, x_window=(w // 2, 64))
    if prec == "bf16":
        xq, wq = x.bfloat16().double(), wt.bfloat16().double()
    else:
        xq, wq = x.double(), wt.double()
    ref = F.conv2d(xq, wq, stride=2, padding=3)
    got = y.to_float().permute(0, 3, 1, 2).double()
    assert (got - ref).abs().max() <= tol * ref.abs().max() + tol * ref.abs().mean()


def test_fp32_nchw_bilinear_and_avgpool():
    from unipose_b200 import _lib, ops
    dev = torch.device("cuda:0")
    x = torch.randn(2, 5, 12, 12, device=dev)
    got = ops.upsample_bilinear_ac_nchw(x, (96, 96))
    ref = F.interpolate(x, size=(96, 96), mode="bilinear", align_corners=True)
    assert (got - ref).abs().max() < 1e-5
    cm = torch.rand(2, 1, 368, 368, device=dev)
    out = torch.zeros(2, 15, 46, 46, device=dev)
    _lib.call("up_avgpool9s8p1_f32", ops._ptr(cm), ops._ptr(out), 2, 1, 368, 368, 46, 46, 15, 14, ops._stream())
    ref = F.avg_pool2d(cm, 9, 8, 1)
    assert (out[:, 14:15] - ref).abs().max() < 1e-6
    assert float(out[:, :14].abs().max()) == 0.0


def test_mse_and_adam_match_torch():
    from unipose_b200 import _lib, ops
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    pred = torch.randn(4, 17, 48, 48, device=dev)
    tgt = torch.rand(4, 17, 48, 48, device=dev)
    loss = torch.zeros(1, device=dev)
    grad = torch.empty_like(pred)
    scratch = torch.zeros(1, dtype=torch.float64, device=dev)
    _lib.call("up_mse_fwd_bwd", ops._ptr(pred), ops._ptr(tgt), ops._ptr(loss), ops._ptr(grad), ops._ptr(scratch),
              pred.numel(), 1.0, ops._stream())
    p2 = pred.clone().requires_grad_(True)
    ref = F.mse_loss(p2, tgt)
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-6 * float(ref)
    assert (grad - p2.grad).abs().max() < 1e-9
    # Adam: three steps against torch.optim.Adam
    p = torch.randn(1000, device=dev)
    ref_p = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], lr=1e-3)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for step in range(1, 4):
        g = torch.randn(1000, device=dev)
        ref_p.grad = g.clone()
        opt.step()
        _lib.call("up_adam_step", ops._ptr(p), ops._ptr(g), ops._ptr(m), ops._ptr(v), p.numel(), 1e-3, 0.9, 0.999,
                  1e-8, step, ops._stream())
    assert (p - ref_p.detach()).abs().max() < 1e-6
