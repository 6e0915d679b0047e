"""GPU parity of the training path (train-mode forward + backward through the whole network) against torch
autograd over the CPU oracle graph on identical weights, inputs and dropout masks (unipose.py:113-124)."""
import warnings

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import evaluate_oracle as E
from oracle import unipose_oracle as O

pytestmark = pytest.mark.gpu


def _rel_l2(a, b):
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _setup(n=4, size=96, seed=0, precision="fp32"):
    from unipose_b200.model.unipose import unipose
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = unipose(dataset="MPII", num_classes=16, precision=precision)
    sd = O.synth_state_dict(16, seed=seed)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().train()
    x = O.synth_input(n, size, size, seed=seed)
    hs = size // 8
    target = torch.from_numpy(E.gaussian_heatmaps(n, 16, hs, hs, seed=seed + 3))
    g = torch.Generator().manual_seed(seed + 9)
    masks = []
    for shape, p in (((n, 256, size // 16, size // 16), 0.5), ((n, 256, hs, hs), 0.5), ((n, 256, hs, hs), 0.1)):
        masks.append((torch.rand(shape, generator=g) >= p).float() / (1.0 - p))
    return m, sd, x, target, masks


def _oracle_step(sd, x, target, masks, dtype=torch.float32):
    sd = {k: (v.clone().to(dtype).requires_grad_(True) if v.is_floating_point() and "running" not in k
              else (v.clone().to(dtype) if v.is_floating_point() else v.clone())) for k, v in sd.items()}
    heat = O.unipose_forward(x.to(dtype), sd, training=True, dropout_masks=[m.to(dtype) for m in masks])
    loss = F.mse_loss(heat, target.to(dtype))
    loss.backward()
    return heat.detach(), loss.detach(), sd


def _cos(a, b):
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-300))


CHECK = ["backbone.conv1.weight", "backbone.bn1.weight", "backbone.layer1.0.conv1.weight",
         "backbone.layer1.0.downsample.0.weight", "backbone.layer2.0.conv2.weight", "backbone.layer3.11.conv2.weight",
         "backbone.layer3.22.bn3.bias", "backbone.layer4.2.conv2.weight", "wasp.aspp1.atrous_conv.weight",
         "wasp.aspp3.atrous_conv.weight", "wasp.conv2.weight", "wasp.global_avg_pool.1.weight",
         "wasp.global_avg_pool.2.weight", "wasp.conv1.weight", "wasp.bn1.weight", "decoder.conv1.weight",
         "decoder.last_conv.0.weight", "decoder.last_conv.4.weight", "decoder.last_conv.5.bias",
         "decoder.last_conv.8.weight", "decoder.last_conv.8.bias"]


def test_train_step_matches_oracle_autograd():
    m, sd, x, target, masks = _setup()
    from unipose_b200 import train
    heat = train.forward_train(m, x.cuda(), dropout_masks=[t.cuda() for t in masks])
    assert heat.requires_grad and heat.shape == (4, 17, 12, 12)
    loss = F.mse_loss(heat, target.cuda())
    loss.backward()
    ref_heat, ref_loss, ref_sd = _oracle_step(sd, x, target, masks)                        # the reference's fp32
    h64, l64, sd64 = _oracle_step(sd, x, target, masks, dtype=torch.float64)               # ground truth
    assert _rel_l2(heat, ref_heat) < 1e-3 and _rel_l2(heat, h64) < 1e-3
    assert abs(float(loss.detach()) - float(ref_loss)) < 2e-3 * float(ref_loss)
    params = dict(m.named_parameters())
    # Back-propagation through ~100 ReLU/BatchNorm layers of this synthetic net is ill-conditioned: the reference's
    # OWN fp32 gradients differ from the fp64 ones by ~1e-2 (ReLU gates flipping), so parity is judged against that
    # noise floor.  The bf16x3 split arithmetic rounds operands at 2^-17 (fp32: 2^-24), so its seed error is
    # ~20-30x the fp32 one already at the head (measured 1.6e-4 vs 7.4e-6 on last_conv.8.weight) and the same
    # conditioning amplifies both: the bound is 40x the reference's fp32-vs-fp64 error, never above 0.1, cosine
    # similarity > 0.995 - while the per-kernel tests (test_gpu_train_ops.py) hold the tight bounds.
    report = {}
    for k in CHECK:
        assert params[k].grad is not None, k
        ours = _rel_l2(params[k].grad, sd64[k].grad)
        floor = _rel_l2(ref_sd[k].grad, sd64[k].grad)
        report[k] = (ours, floor, _cos(params[k].grad, sd64[k].grad))
    print("grad rel-L2 (ours vs fp64, reference-fp32 vs fp64, cosine):",
          {k: "%.1e / %.1e / %.5f" % v for k, v in report.items()})
    for k, (ours, floor, cos) in report.items():
        assert ours <= min(max(40.0 * floor, 2e-3), 0.1), (k, ours, floor)
        assert cos > 0.995, (k, report[k])
    assert report["decoder.last_conv.8.weight"][0] < 1e-3 and report["decoder.last_conv.8.bias"][0] < 1e-3
    # dead parameters of the reference stay without gradient (decoder.py:20-21)
    assert params["decoder.conv2.weight"].grad is None and params["decoder.bn2.weight"].grad is None
    # running statistics were updated like torch's (momentum 0.1, unbiased variance)
    bufs = dict(m.named_buffers())
    for k in ("backbone.bn1.running_mean", "backbone.layer3.5.bn2.running_var", "wasp.bn1.running_var",
              "decoder.last_conv.1.running_mean"):
        assert _rel_l2(bufs[k], ref_sd[k]) < 2e-3, k
    assert int(bufs["backbone.bn1.num_batches_tracked"]) == 1


def test_train_step_config3_geometry_384_bs4():
    """BASELINE.json configs[2] geometry (384x384 -> 24x24 maps: image-pair tiles, CTA pairs and tap skipping in the
    dgrad / wgrad kernels, 48x48 decoder maps) at batch 4: loss, heat-maps and gradients from the stem to the head
    against torch autograd over the CPU oracle (fp32 = the reference's arithmetic, fp64 = ground truth)."""
    m, sd, x, target, masks = _setup(n=4, size=384, seed=3)
    from unipose_b200 import train
    heat = train.forward_train(m, x.cuda(), dropout_masks=[t.cuda() for t in masks])
    assert heat.shape == (4, 17, 48, 48)
    loss = F.mse_loss(heat, target.cuda())
    loss.backward()
    ref_heat, ref_loss, ref_sd = _oracle_step(sd, x, target, masks)
    h64, l64, sd64 = _oracle_step(sd, x, target, masks, dtype=torch.float64)
    assert _rel_l2(heat, h64) < 1e-3
    assert abs(float(loss.detach()) - float(l64)) < 2e-3 * float(l64)
    params = dict(m.named_parameters())
    report = {}
    for k in CHECK:
        ours = _rel_l2(params[k].grad, sd64[k].grad)
        floor = _rel_l2(ref_sd[k].grad, sd64[k].grad)
        report[k] = (ours, floor, _cos(params[k].grad, sd64[k].grad))
    print("384^2 bs4 grad rel-L2 (ours vs fp64, reference-fp32 vs fp64, cosine):",
          {k: "%.1e / %.1e / %.5f" % v for k, v in report.items()})
    for k, (ours, floor, cos) in report.items():
        assert ours <= min(max(40.0 * floor, 2e-3), 0.1), (k, ours, floor)
        assert cos > 0.995, (k, report[k])
    assert report["decoder.last_conv.8.weight"][0] < 1e-3


def test_reference_training_loop_runs_unchanged():
    """optimizer.zero_grad(); heat = model(x); loss = MSELoss(heat, target); loss.backward(); optimizer.step()."""
    m, sd, x, target, masks = _setup(n=2, size=96, seed=1, precision="bf16")
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    crit = torch.nn.MSELoss().cuda()
    losses = []
    xc, tc = x.cuda(), target.cuda()
    for _ in range(3):
        opt.zero_grad()
        heat = m(xc)
        loss = crit(heat, tc)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert all(np.isfinite(losses)), losses
    assert losses[-1] < losses[0], losses


def test_fused_train_step_matches_manual_adam():
    from unipose_b200 import train
    m, sd, x, target, masks = _setup(n=2, size=96, seed=2, precision="fp32")
    for mod in m.modules():     # dropout off so that both paths see the same network
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    ts = train.TrainStep(m, lr=1e-3)
    before = {k: v.detach().clone() for k, v in m.named_parameters()}
    loss = ts.step(x.cuda(), target.cuda())
    torch.cuda.synchronize()
    # oracle: same step with torch.optim.Adam on the CPU graph
    sd2 = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
           for k, v in sd.items()}
    heat = O.unipose_forward(x, sd2, training=True)
    ref_loss = F.mse_loss(heat, target)
    ref_loss.backward()
    live = [k for k, v in sd2.items() if getattr(v, "grad", None) is not None]
    opt = torch.optim.Adam([sd2[k] for k in live], lr=1e-3)
    opt.step()
    assert abs(float(loss) - float(ref_loss)) < 2e-3 * float(ref_loss)
    after = dict(m.named_parameters())
    for k in ("backbone.layer3.11.conv2.weight", "wasp.conv2.weight", "decoder.last_conv.8.bias"):
        # Adam's first step moves every weight by ~lr * sign(grad): compare the update direction
        upd = (after[k].detach().cpu() - before[k].cpu())
        ref_upd = (sd2[k].detach() - sd[k])
        agree = (torch.sign(upd) == torch.sign(ref_upd)).float().mean()
        assert agree > 0.97, (k, float(agree))
    assert torch.equal(after["decoder.conv2.weight"].detach().cpu(), before["decoder.conv2.weight"].cpu())


def test_fused_train_steps_parameter_deltas_match_torch_adam():
    """Three fused steps (MSE + backward + Adam on flat buffers, the third one replayed from CUDA graphs) against three
    steps of torch.optim.Adam on the CPU oracle graph: the accumulated parameter UPDATES themselves, not just their signs
    (after several steps m / sqrt(v) no longer saturates at +-lr, so the deltas carry the gradient magnitudes)."""
    from unipose_b200 import train
    m, sd, x, target, masks = _setup(n=2, size=96, seed=5, precision="fp32")
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    ts = train.TrainStep(m, lr=1e-5)      # small steps: the test is about the update vector, not about chaotic dynamics
    before = {k: v.detach().clone().cpu() for k, v in m.named_parameters()}
    xc, tc = x.cuda(), target.cuda()
    losses = [float(ts.step(xc, tc)) for _ in range(3)]
    torch.cuda.synchronize()
    assert ts.graph is not None
    sd2 = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
           for k, v in sd.items()}
    live = None
    opt = None
    ref_losses = []
    for _ in range(3):
        heat = O.unipose_forward(x, sd2, training=True)
        loss = F.mse_loss(heat, target)
        if opt is not None:
            opt.zero_grad()
        loss.backward()
        if opt is None:
            live = [k for k, v in sd2.items() if getattr(v, "grad", None) is not None]
            opt = torch.optim.Adam([sd2[k] for k in live], lr=1e-5)
        opt.step()
        ref_losses.append(float(loss))
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) < 1e-2 * abs(b), (losses, ref_losses)   # 1.2184, 0.9615, 0.8548 vs 1.2185, 0.9635, 0.8610 measured
    after = dict(m.named_parameters())
    report = {}
    for k in ("decoder.last_conv.8.weight", "decoder.last_conv.8.bias", "decoder.last_conv.4.weight", "wasp.conv2.weight",
              "wasp.aspp3.atrous_conv.weight", "backbone.layer4.2.conv2.weight", "backbone.layer3.11.conv2.weight"):
        upd = after[k].detach().cpu() - before[k]
        ref_upd = sd2[k].detach() - sd[k]
        report[k] = (_rel_l2(upd, ref_upd), _cos(upd, ref_upd))
    print("3-step Adam parameter deltas (rel-L2, cosine vs torch.optim.Adam on the oracle):",
          {k: "%.2e / %.4f" % v for k, v in report.items()})
    # Adam divides by sqrt(v): elements whose gradient is at the noise level get O(1) relative changes in their update,
    # so the bound is on the update VECTOR: head tight, deep layers within the gradient noise documented above
    # measured on B200: head 2.0e-2 / 2.6e-4 (cosine 0.9998 / 1.0000), decoder 0.11 (0.994), WASP 0.19-0.22 (0.98),
    # layer4 0.30 (0.954), layer3 0.35 (0.937) - Adam's element-wise 1 / sqrt(v) turns the gradient noise of the deep
    # layers (tests above) into O(1) changes of the elements whose gradient is at the noise level
    assert report["decoder.last_conv.8.weight"][0] < 3e-2 and report["decoder.last_conv.8.bias"][0] < 2e-3, report
    for k, (rel, cos) in report.items():
        assert rel < 0.45 and cos > 0.90, (k, rel, cos)
    assert torch.equal(after["decoder.conv2.weight"].detach().cpu(), before["decoder.conv2.weight"])


def test_train_step_graph_replay_matches_eager(monkeypatch):
    """From the third step on forward + loss + backward replay as one CUDA graph: same losses and weights as eager."""
    from unipose_b200 import train
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("UNIPOSE_B200_TRAIN_GRAPH", flag)
        m, sd, x, target, masks = _setup(n=2, size=96, seed=4, precision="bf16")
        for mod in m.modules():     # dropout off: the two runs must see the same network
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        ts = train.TrainStep(m, lr=1e-4)
        xc, tc = x.cuda(), target.cuda()
        losses = [float(ts.step(xc, tc)) for _ in range(5)]
        assert (ts.graph is not None) == (flag == "1")
        out[flag] = (losses, dict(m.named_parameters())["wasp.conv1.weight"].detach().clone())
    assert out["0"][0] == out["1"][0], out
    assert torch.equal(out["0"][1], out["1"][1])
    assert out["1"][0][-1] < out["1"][0][0]
