"""The fused bottleneck tail (csrc/bneck_tail.cu: 3x3 conv + 1x1 expansion + residual + ReLU in one launch, layer1 /
layer2) against the layer-wise plan and the CPU oracle, incl. partial tiles (368 -> 92x92 / 46x46 maps) and small maps."""
import warnings

import numpy as np
import pytest
import torch

from oracle import unipose_oracle as O

pytestmark = pytest.mark.gpu


def _model(precision, seed=0):
    from unipose_b200.model.unipose import unipose
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = unipose(dataset="MPII", num_classes=16, precision=precision)
    sd = O.synth_state_dict(16, seed=seed)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd


def _tails(names):
    return names.count("bottleneck.tail") + names.count("bottleneck.tail+conv1")


def _run(m, x, tail, monkeypatch):
    monkeypatch.setenv("UNIPOSE_B200_BNECK_TAIL", "1" if tail else "0")
    m._plans.clear()
    out = m(x.cuda())
    torch.cuda.synchronize()
    names = [n for n, f, s in m.plan_for(x.cuda()).ops if f is not None]
    # layer1 x3 (block 0 with its projection shortcut inside the kernel) + layer2 blocks 1..3; the other projection
    # shortcuts ride in their block's conv3 (UP_FLAG_PROJ): no downsample launch is left
    assert _tails(names) == (6 if tail else 0), names
    assert names.count("bottleneck.downsample") == 0 and names.count("bottleneck.conv3+proj") == (3 if tail else 4), names
    return out.cpu().numpy()


@pytest.mark.parametrize("n,size,precision", [(4, 384, "fp16"), (3, 368, "fp16"), (1, 96, "fp16"), (2, 256, "bf16"),
                                              (2, 512, "fp16")])
def test_bneck_tail_matches_layerwise_and_oracle(n, size, precision, monkeypatch):
    m, sd = _model(precision, seed=13)
    x = O.synth_input(n, size, size, seed=13)
    with torch.no_grad():
        ref = O.unipose_forward(x, sd).numpy()
    got = _run(m, x, True, monkeypatch)
    assert np.array_equal(got, m(x.cuda()).cpu().numpy())          # graph replay: same bits
    base = _run(m, x, False, monkeypatch)
    scale = float(np.abs(ref).max())
    e_tail = float(np.abs(got - ref).max() / scale)
    e_base = float(np.abs(base - ref).max() / scale)
    print("bottleneck tail %dx%d^2 %s: max-rel %.3g (layer-wise %.3g)" % (n, size, precision, e_tail, e_base))
    bound = 5e-3 if precision == "fp16" else 3e-2
    assert e_tail < bound and e_base < bound, (e_tail, e_base)
    assert e_tail < 2.0 * e_base + 1e-3


def test_projection_shortcut_inside_the_tail_matches_the_separate_launch(monkeypatch):
    """layer1 block 0: downsample(x) accumulated in the tail kernel vs written by its own launch and read back."""
    m, sd = _model("fp16", seed=5)
    x = O.synth_input(2, 368, 368, seed=5)        # 92x92 maps: partial tiles
    with torch.no_grad():
        ref = O.unipose_forward(x, sd).numpy()
    monkeypatch.setenv("UNIPOSE_B200_PROJ_FUSE", "0")     # layers 2-4 keep their downsample launches in both runs
    monkeypatch.setenv("UNIPOSE_B200_BNECK_TAIL_PROJ", "0")
    m._plans.clear()
    sep = m(x.cuda()).cpu().numpy()
    names = [n for n, f, s in m.plan_for(x.cuda()).ops if f is not None]
    assert _tails(names) == 6 and names.count("bottleneck.downsample") == 4
    monkeypatch.setenv("UNIPOSE_B200_BNECK_TAIL_PROJ", "1")
    m._plans.clear()
    fused = m(x.cuda()).cpu().numpy()
    names = [n for n, f, s in m.plan_for(x.cuda()).ops if f is not None]
    assert _tails(names) == 6 and names.count("bottleneck.downsample") == 3
    scale = float(np.abs(ref).max())
    e_f, e_s = float(np.abs(fused - ref).max() / scale), float(np.abs(sep - ref).max() / scale)
    print("projection in the tail: max-rel %.3g (separate launch %.3g)" % (e_f, e_s))
    assert e_f < 5e-3 and e_f < 2.0 * e_s + 1e-3


@pytest.mark.parametrize("n,size,precision", [(4, 384, "fp16"), (3, 368, "fp16"), (1, 96, "bf16"), (2, 256, "fp16"),
                                              (5, 160, "fp16")])
def test_next_conv1_inside_the_tail_matches_its_own_launch(n, size, precision, monkeypatch):
    """layer1: the following block's conv1 + bn1 + ReLU computed by the tail kernel from the on-chip output tile
    (64 -> 64 twice, 64 -> 128 into layer2 block 0) vs the separate conv1 launches."""
    m, sd = _model(precision, seed=17)
    x = O.synth_input(n, size, size, seed=17)
    with torch.no_grad():
        ref = O.unipose_forward(x, sd).numpy()
    outs, conv1s = {}, {}
    for fuse in ("0", "1"):
        monkeypatch.setenv("UNIPOSE_B200_TAIL_CONV1", fuse)
        m._plans.clear()
        outs[fuse] = m(x.cuda()).cpu().numpy()
        if fuse == "1":
            assert np.array_equal(outs[fuse], m(x.cuda()).cpu().numpy())      # graph replay: same bits
        names = [nm for nm, f, s in m.plan_for(x.cuda()).ops if f is not None]
        assert names.count("bottleneck.tail+conv1") == (3 if fuse == "1" else 0), names
        conv1s[fuse] = names.count("bottleneck.conv1")
    assert conv1s["1"] == conv1s["0"] - 3, conv1s
    scale = float(np.abs(ref).max())
    e_f = float(np.abs(outs["1"] - ref).max() / scale)
    e_s = float(np.abs(outs["0"] - ref).max() / scale)
    print("next conv1 in the tail %dx%d^2 %s: max-rel %.3g (own launch %.3g)" % (n, size, precision, e_f, e_s))
    bound = 5e-3 if precision == "fp16" else 3e-2
    assert e_f < bound and e_f < 2.0 * e_s + 1e-3, (e_f, e_s)
