"""Projection shortcut inside conv3's GEMM (UP_FLAG_PROJ, csrc/conv_tcgen05.cu): out = ReLU(W3 t2 + Wd x_strided + shift),
the first bottleneck of layer2 / layer3 / layer4 (resnet.py:22-42 with `downsample`, :75-79) as one launch instead of
downsample + conv3-with-residual.  Op level against fp64 torch on the same 16-bit operands, model level against the
unfused plan and the CPU oracle."""
import warnings

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import unipose_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,ho,wo,planes,cin_x,stride,prec", [
    (4, 48, 48, 128, 256, 2, "fp16"),      # layer2 block 0 (MPII 384^2)
    (4, 24, 24, 256, 512, 2, "bf16"),      # layer3 block 0
    (4, 24, 24, 512, 1024, 1, "fp16"),     # layer4 block 0 (output_stride 16)
    (3, 23, 23, 128, 256, 2, "fp16"),      # odd batch (no CTA pair), partial tiles
    (2, 12, 20, 64, 64, 1, "bf16"),        # layer1 block 0 shape, one projection slice, non-square map
])
def test_conv3_with_projection_matches_fp64(n, ho, wo, planes, cin_x, stride, prec):
    from unipose_b200 import ops
    dev = torch.device("cuda:0")
    mode = ops.mode_of(prec)
    dt = torch.float16 if prec == "fp16" else torch.bfloat16
    g = torch.Generator(device="cpu").manual_seed(1234 + planes + stride)
    cout = 4 * planes
    t2 = ops.Act(n, ho, wo, planes, mode, dev)
    x = ops.Act(n, ho * stride, wo * stride, cin_x, mode, dev)
    t2.t.copy_(torch.randn(t2.t.shape, generator=g).to(dt))
    x.t.copy_(torch.randn(x.t.shape, generator=g).to(dt))
    w3 = (torch.randn(cout, planes, generator=g) / planes ** 0.5).to(dt)
    wd = (torch.randn(cout, cin_x, generator=g) / cin_x ** 0.5).to(dt)
    shift = torch.randn(cout, generator=g)
    pt = cin_x // planes
    wbuf = torch.empty((1, 1 + pt, cout, planes), dtype=dt, device=dev)
    wbuf[0, 0].copy_(w3)
    for j in range(pt):
        wbuf[0, 1 + j].copy_(wd[:, j * planes:(j + 1) * planes])
    pc = ops.PackedConv(wbuf, torch.ones(cout, device=dev), shift.to(dev), 1, 1, cout, planes, cout, planes, mode)
    y = ops.Act(n, ho, wo, cout, mode, dev)
    ops.conv2d(t2, pc, y, relu=True, proj=(x, stride))
    torch.cuda.synchronize()
    t2f = t2.t[0].double().cpu()
    xs = x.t[0].double().cpu()[:, ::stride, ::stride]
    ref = torch.relu(t2f @ w3.double().t() + xs @ wd.double().t() + shift.double())
    got = y.t[0].double().cpu()
    tol = 2 ** -10 if prec == "fp16" else 2 ** -7           # one rounding of the 16-bit output
    err = float((got - ref).abs().max() / ref.abs().max())
    print("conv3+proj %s planes %d stride %d: max-rel %.3g" % (prec, planes, stride, err))
    assert err < tol, err


def _model(precision, seed=0, **kw):
    from unipose_b200.model.unipose import unipose
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = unipose(dataset="MPII", num_classes=16, precision=precision, **kw)
    sd = O.synth_state_dict(16, seed=seed)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd


@pytest.mark.parametrize("n,size,precision,os_", [(4, 384, "fp16", 16), (3, 368, "fp16", 16), (2, 256, "bf16", 8)])
def test_fused_projection_matches_unfused_plan_and_oracle(n, size, precision, os_, monkeypatch):
    kw = {"output_stride": os_}
    m, sd = _model(precision, seed=21, **kw)
    x = O.synth_input(n, size, size, seed=21)
    with torch.no_grad():
        ref = O.unipose_forward(x, sd, output_stride=os_).numpy()
    outs = {}
    for fuse in ("1", "0"):
        monkeypatch.setenv("UNIPOSE_B200_PROJ_FUSE", fuse)
        m._plans.clear()
        outs[fuse] = m(x.cuda()).cpu().numpy()
        names = [nm for nm, f, s in m.plan_for(x.cuda()).ops if f is not None]
        assert names.count("bottleneck.conv3+proj") == (3 if fuse == "1" else 0), names
        assert names.count("bottleneck.downsample") == (0 if fuse == "1" else 3), names
    scale = float(np.abs(ref).max())
    e_f = float(np.abs(outs["1"] - ref).max() / scale)
    e_u = float(np.abs(outs["0"] - ref).max() / scale)
    print("projection in conv3 %dx%d^2 %s os%d: max-rel %.3g (unfused %.3g)" % (n, size, precision, os_, e_f, e_u))
    bound = 5e-3 if precision == "fp16" else 3e-2
    assert e_f < bound and e_f < 2.0 * e_u + 1e-3, (e_f, e_u)


def test_fused_projection_follows_parameter_updates():
    """The summed shift (bn3 + downsample BN) and both filters are re-derived when parameters change in place."""
    m, sd = _model("fp16", seed=3)
    x = O.synth_input(2, 128, 128, seed=3)
    a = m(x.cuda()).clone()
    with torch.no_grad():
        bn = m.backbone.layer2[0].downsample[1]
        bn.bias.add_(0.5)
        m.backbone.layer3[0].downsample[0].weight.mul_(1.25)
    b = m(x.cuda()).clone()
    sd2 = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    with torch.no_grad():
        ref = O.unipose_forward(x, sd2).numpy()
    assert not torch.equal(a, b)
    err = float(np.abs(b.cpu().numpy() - ref).max() / np.abs(ref).max())
    assert err < 5e-3, err
