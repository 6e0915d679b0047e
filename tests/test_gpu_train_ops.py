"""GPU parity of the backward kernels through the C-ABI against torch autograd (fp64) on the same operands:
tcgen05 wgrad (MN-major operands), dgrad (= forward kernel on flipped/transposed filters, zero insertion for
stride 2), train-mode BatchNorm forward/backward, max-pool / bilinear / GAP adjoints."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

WGRAD_CASES = [
    # n, h, w, cin, cout, k, stride, dil, groups, prec
    (2, 16, 16, 64, 64, 1, 1, 1, 1, "bf16"),
    (2, 24, 24, 256, 256, 3, 1, 1, 1, "bf16"),
    (4, 24, 24, 256, 256, 3, 1, 12, 1, "bf16"),
    (2, 48, 48, 128, 128, 3, 2, 1, 1, "bf16"),
    (2, 48, 48, 256, 512, 1, 2, 1, 1, "bf16"),
    (3, 23, 23, 64, 128, 3, 1, 1, 1, "fp16"),
    (2, 24, 24, 1024, 256, 1, 1, 1, 1, "bf16"),
    (2, 24, 24, 512, 256, 1, 1, 1, 2, "bf16"),
    (2, 24, 24, 320, 256, 3, 1, 1, 1, "bf16"),
    (2, 24, 24, 128, 128, 3, 1, 2, 1, "fp32"),
    (32, 24, 24, 256, 256, 3, 1, 1, 1, "bf16"),
    (2, 46, 46, 16, 128, 11, 1, 1, 1, "bf16"),
    (2, 48, 48, 256, 17, 1, 1, 1, 1, "bf16"),
]


def _quant(t, prec):
    if prec == "bf16":
        return t.bfloat16().double()
    if prec == "fp16":
        return t.half().double()
    hi = t.bfloat16()
    lo = (t - hi.float()).bfloat16()
    return hi.double() + lo.double()


@pytest.mark.parametrize("case", WGRAD_CASES)
def test_wgrad_vs_autograd(case):
    from unipose_b200 import ops
    n, h, w, cin, cout, k, stride, dil, groups, prec = case
    dev = torch.device("cuda:0")
    mode = ops.mode_of(prec)
    torch.manual_seed(7)
    pad = dil * (k - 1) // 2
    ho = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1
    wo = (w + 2 * pad - dil * (k - 1) - 1) // stride + 1
    cin_pad, cout_pad = ops.round_up(cin, 16), ops.round_up(cout, 64)
    x = torch.randn(n, cin, h, w, device=dev)
    dz = torch.randn(n, cout, ho, wo, device=dev)
    if groups > 1:
        cg = cin // groups
        xa = ops.Act(n * groups, h, w, cg, mode, dev)
        for g in range(groups):
            ops.nchw_to_act(x[:, g * cg:(g + 1) * cg].contiguous(), ops.View(xa, n_off=g * n, n=n))
        xin = ops.View(xa, n_off=0, n=n)
    else:
        xa = ops.Act(n, h, w, cin_pad, mode, dev)
        ops.nchw_to_act(x, xa)
        xin = xa
    dza = ops.Act(n, ho, wo, cout_pad, mode, dev)
    ops.nchw_to_act(dz, dza)
    wt = torch.zeros(cout, cin, k, k, device=dev)
    pc = ops.make_packed_conv(wt, mode, cout=cout_pad, cin=cin_pad)
    d = ops.conv_desc(xin, pc, ho, wo, stride=stride, dil=dil, pad=(pad, pad), x_groups=groups,
                      x_group_nstride=n if groups > 1 else 0)
    scratch = torch.empty(ops.wgrad_scratch_bytes(d) // 4, dtype=torch.float32, device=dev)
    dw = torch.full((cout, cin, k, k), float("nan"), device=dev)
    ops.conv2d_wgrad(d, xin, dza, dw, scratch)
    torch.cuda.synchronize()
    xq = _quant(x, prec).requires_grad_(False)
    wq = torch.zeros(cout, cin, k, k, device=dev, dtype=torch.double, requires_grad=True)
    y = F.conv2d(xq, wq, stride=stride, dilation=dil, padding=pad)
    (ref,) = torch.autograd.grad(y, wq, _quant(dz, prec))
    err = (dw.double() - ref).abs().max()
    tol = {"bf16": 2e-5, "fp16": 2e-5, "fp32": 2e-4}[prec]   # fp32 accumulation of exact 16-bit products
    assert torch.isfinite(dw).all()
    assert err <= tol * ref.abs().max() + 1e-6, (float(err), float(ref.abs().max()))
    # accumulate=1 doubles the result
    ops.conv2d_wgrad(d, xin, dza, dw, scratch, accumulate=True)
    assert (dw.double() - 2 * ref).abs().max() <= 2 * tol * ref.abs().max() + 2e-6


@pytest.mark.parametrize("stride,k,dil,prec", [(1, 3, 1, "bf16"), (1, 3, 6, "bf16"), (2, 3, 1, "bf16"),
                                               (2, 1, 1, "bf16"), (1, 1, 1, "fp32")])
def test_dgrad_is_forward_kernel_on_flipped_filter(stride, k, dil, prec):
    from unipose_b200 import ops
    dev = torch.device("cuda:0")
    mode = ops.mode_of(prec)
    torch.manual_seed(5)
    n, h, w, cin, cout = 2, 24, 24, 128, 64
    pad = dil * (k - 1) // 2
    ho, wo = h // stride, w // stride
    wt = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    dz = torch.randn(n, cout, ho, wo, device=dev)
    dza = ops.Act(n, ho, wo, cout, mode, dev)
    ops.nchw_to_act(dz, dza)
    src = dza
    if stride == 2:
        src = ops.Act(n, h, w, cout, mode, dev)
        ops.zero_insert2x(dza, src)
    wt_t = wt.flip(2, 3).transpose(0, 1).contiguous()            # [cin, cout, k, k]
    pc = ops.make_packed_conv(wt_t, mode, cout=cin, cin=cout)
    dx = ops.Act(n, h, w, cin, mode, dev)
    ops.conv2d(src, pc, dx, dil=dil, pad=dil * (k - 1) - pad, ho=h, wo=w)
    xq = torch.zeros(n, cin, h, w, device=dev, dtype=torch.double, requires_grad=True)
    y = F.conv2d(xq, _quant(wt, prec), stride=stride, dilation=dil, padding=pad)
    (ref,) = torch.autograd.grad(y, xq, _quant(dz, prec))
    got = dx.to_float().permute(0, 3, 1, 2).double()
    tol = {"bf16": 2 ** -7, "fp32": 1e-4}[prec]
    assert (got - ref).abs().max() <= tol * ref.abs().max()


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
@pytest.mark.parametrize("c,relu,res", [(64, True, False), (256, True, True), (2048, False, False)])
def test_batchnorm_train_forward_backward(prec, c, relu, res):
    from unipose_b200 import ops
    dev = torch.device("cuda:0")
    mode = ops.mode_of(prec)
    torch.manual_seed(11)
    n, h, w = 4, 12, 10
    z = torch.randn(n, c, h, w, device=dev) * 2 + 0.5
    r = torch.randn(n, c, h, w, device=dev)
    dy = torch.randn(n, c, h, w, device=dev)
    bn = torch.nn.BatchNorm2d(c).to(dev)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.2)
    za, ra, dya = (ops.Act(n, h, w, c, mode, dev) for _ in range(3))
    ops.nchw_to_act(z, za)
    ops.nchw_to_act(r, ra)
    ops.nchw_to_act(dy, dya)
    zq = za.to_float().permute(0, 3, 1, 2).double().requires_grad_(True)
    rq = ra.to_float().permute(0, 3, 1, 2).double().requires_grad_(True)
    dyq = dya.to_float().permute(0, 3, 1, 2).double()
    sums = torch.zeros(ops.bn_work_doubles(c), dtype=torch.float64, device=dev)
    scale, shift, mean, invstd = (torch.empty(c, device=dev) for _ in range(4))
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    ops.bn_stats(za, c, sums)
    ops.bn_finalize(sums, n * h * w, bn, scale, shift, mean, invstd, c, c)
    # the two-launch form the training plan uses: same numbers, bit for bit (same reduction order)
    bn2 = torch.nn.BatchNorm2d(c).to(dev)
    bn2.load_state_dict(bn.state_dict())
    with torch.no_grad():
        bn2.running_mean.copy_(rm0)
        bn2.running_var.copy_(rv0)
    sums2 = torch.zeros_like(sums)
    scale2, shift2, mean2, invstd2 = (torch.empty(c, device=dev) for _ in range(4))
    ops.bn_stats_finalize(za, sums2, bn2, scale2, shift2, mean2, invstd2, c, c)
    for a, b in ((scale, scale2), (shift, shift2), (mean, mean2), (invstd, invstd2),
                 (bn.running_mean, bn2.running_mean), (bn.running_var, bn2.running_var), (sums[:2 * c], sums2[:2 * c])):
        assert torch.equal(a, b)
    ya = ops.Act(n, h, w, c, mode, dev)
    ops.scale_shift_act(za, ya, scale, shift, relu=relu, residual=ra if res else None)
    # reference: F.batch_norm in training mode on the same (quantised) input, fp64
    rm, rv = rm0.double(), rv0.double()
    wd = bn.weight.detach().double().requires_grad_(True)
    bd = bn.bias.detach().double().requires_grad_(True)
    ref = F.batch_norm(zq, rm, rv, wd, bd, training=True, momentum=0.1, eps=bn.eps)
    if res:
        ref = ref + rq
    if relu:
        ref = ref.relu()
    got = ya.to_float().permute(0, 3, 1, 2).double()
    tol = {"bf16": 2 ** -7, "fp32": 3e-5}[prec]
    assert (got - ref.detach()).abs().max() <= tol * ref.abs().max() + 1e-5
    assert (bn.running_mean.double() - rm).abs().max() < 1e-5 and (bn.running_var.double() - rv).abs().max() < 1e-4
    # backward
    grads = torch.autograd.grad(ref, (zq, wd, bd) + ((rq,) if res else ()), dyq)
    gz, gw, gb = grads[0], grads[1], grads[2]
    gr = grads[3] if res else None
    ref = ref.detach()
    dza = ops.Act(n, h, w, c, mode, dev)
    dra = ops.Act(n, h, w, c, mode, dev) if res else None
    dgamma, dbeta = torch.empty(c, device=dev), torch.empty(c, device=dev)
    ops.bn_bwd(dya, ya, za, dza, dra, mean, invstd, bn.weight.detach(), sums, c, relu, dgamma, dbeta)
    gotz = dza.to_float().permute(0, 3, 1, 2).double()
    # the ReLU gate uses the kernel's own (rounded) forward output; compare where the reference is not at the kink
    safe = (ref.abs() > 1e-2) | (not relu)
    tolb = {"bf16": 3e-2, "fp32": 2e-4}[prec]
    assert ((gotz - gz).abs() * safe).max() <= tolb * gz.abs().max()
    if prec == "fp32":
        assert (dgamma.double() - gw).abs().max() <= 1e-3 * gw.abs().max() + 1e-4
        assert (dbeta.double() - gb).abs().max() <= 1e-3 * gb.abs().max() + 1e-4
    if res:
        gotr = dra.to_float().permute(0, 3, 1, 2).double()
        assert ((gotr - gr).abs() * safe).max() <= tolb * gr.abs().max()


def test_pool_upsample_gap_adjoints():
    from unipose_b200 import ops
    dev = torch.device("cuda:0")
    mode = ops.UP_SPLIT
    torch.manual_seed(2)
    n, c, h, w = 2, 64, 24, 20
    x = torch.randn(n, c, h, w, device=dev)
    xa = ops.Act(n, h, w, c, mode, dev)
    ops.nchw_to_act(x, xa)
    xq = xa.to_float().permute(0, 3, 1, 2).double().requires_grad_(True)
    # max-pool
    y = F.max_pool2d(xq, 3, 2, 1)
    dy = torch.randn_like(y)
    (ref,) = torch.autograd.grad(y, xq, dy)
    dya = ops.Act(n, y.shape[2], y.shape[3], c, mode, dev)
    ops.nchw_to_act(dy.float(), dya)
    dxa = ops.Act(n, h, w, c, mode, dev)
    ops.maxpool3x3s2_bwd(xa, dya, dxa)
    got = dxa.to_float().permute(0, 3, 1, 2).double()
    assert (got - ref).abs().max() < 1e-4
    ops.maxpool3x3s2_bwd(xa, dya, dxa, accumulate=True)
    assert (dxa.to_float().permute(0, 3, 1, 2).double() - 2 * ref).abs().max() < 2e-4
    # two-pass form (arg-max map in caller scratch), incl. ties: quantised input has many equal neighbours
    for tie in (False, True):
        xt = (x * 2).round() / 2 if tie else x
        xb = ops.Act(n, h, w, c, mode, dev)
        ops.nchw_to_act(xt, xb)
        xqt = xb.to_float().permute(0, 3, 1, 2).double().requires_grad_(True)
        (reft,) = torch.autograd.grad(F.max_pool2d(xqt, 3, 2, 1), xqt, dy)
        idx = torch.empty(n * dya.h * dya.w * c, dtype=torch.uint8, device=dev)
        dxb = ops.Act(n, h, w, c, mode, dev)
        ops.maxpool3x3s2_bwd(xb, dya, dxb, idx=idx)
        assert (dxb.to_float().permute(0, 3, 1, 2).double() - reft).abs().max() < 1e-4
        ops.maxpool3x3s2_bwd(xb, dya, dxb, accumulate=True, idx=idx)
        assert (dxb.to_float().permute(0, 3, 1, 2).double() - 2 * reft).abs().max() < 2e-4
    # bilinear align_corners
    y = F.interpolate(xq, size=(48, 41), mode="bilinear", align_corners=True)
    dy = torch.randn_like(y)
    (ref,) = torch.autograd.grad(y, xq, dy)
    dya = ops.Act(n, 48, 41, c, mode, dev)
    ops.nchw_to_act(dy.float(), dya)
    ops.upsample_bilinear_ac_bwd(dya, dxa)
    assert (dxa.to_float().permute(0, 3, 1, 2).double() - ref).abs().max() < 2e-4
    # global average pool adjoint and zero insertion
    g = torch.randn(n, c, 1, 1, device=dev)
    ga = ops.Act(n, 1, 1, c, mode, dev)
    ops.nchw_to_act(g, ga)
    ops.add_broadcast(ga, dxa, 1.0 / (h * w), accumulate=False)
    assert (dxa.to_float().permute(0, 3, 1, 2) - ga.to_float().permute(0, 3, 1, 2) / (h * w)).abs().max() < 1e-6
    za = ops.Act(n, 2 * h, 2 * w, c, mode, dev)
    ops.zero_insert2x(xa, za)
    zf = za.to_float()
    assert torch.equal(zf[:, ::2, ::2], xa.to_float()) and float(zf[:, 1::2].abs().max()) == 0.0 \
        and float(zf[:, :, 1::2].abs().max()) == 0.0
