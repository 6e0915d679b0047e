"""The fused WASP chain kernel (csrc/wasp_chain.cu: the whole block as one persistent launch) against the CPU oracle
of wasp.forward (model/modules/wasp.py:66-90) and against the layer-wise plan it replaces, on the shapes the
network produces: 24x24 (config 2, image-pair tiles), 32x32 (config 5, single-image tiles), 23x23 (368x368 video
frames: partial tiles), output_stride=8 dilations, the waspVideo variant, and a forced multi-wave launch."""
import os
import warnings

import numpy as np
import pytest
import torch

from oracle import unipose_oracle as O

pytestmark = pytest.mark.gpu


def _wasp(video, output_stride, precision, seed):
    from unipose_b200.model.modules.wasp import build_wasp
    from unipose_b200.model.modules.waspVideo import build_wasp as build_wasp_video
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = (build_wasp_video if video else build_wasp)("resnet", output_stride, torch.nn.BatchNorm2d)
    m.precision = precision
    sd = {k[len("wasp."):]: v for k, v in O.synth_state_dict(16, video=video, seed=seed, output_stride=output_stride).items()
          if k.startswith("wasp.")}
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), {"wasp." + k: v for k, v in sd.items()}


def _run(m, x, chain, monkeypatch):
    monkeypatch.setenv("UNIPOSE_B200_WASP_CHAIN", "1" if chain else "0")
    m._plans.clear()
    out = m(x.cuda())
    torch.cuda.synchronize()
    plan = next(iter(m._plans.values()))
    names = [n for n, f, s in plan.ops if f is not None]
    assert ("wasp.chain" in names) == chain, names
    return out.cpu().numpy()


@pytest.mark.parametrize("n,hw,os_,video,precision", [
    (4, 24, 16, False, "fp16"), (32, 24, 16, False, "fp16"), (4, 24, 16, False, "bf16"), (2, 32, 16, False, "fp16"),
    (4, 23, 16, True, "fp16"), (2, 32, 8, False, "fp16"), (2, 16, 16, False, "fp16")])
def test_chain_matches_oracle_and_layerwise_plan(n, hw, os_, video, precision, monkeypatch):
    m, sd = _wasp(video, os_, precision, seed=21)
    g = torch.Generator().manual_seed(5)
    # post-ReLU backbone features as the 16-bit kernels see them
    x = (torch.randn(n, 2048, hw, hw, generator=g).clamp_min_(0) * 0.5)
    x = x.half().float() if precision == "fp16" else x.bfloat16().float()
    with torch.no_grad():
        ref = O.wasp_forward(x, sd, output_stride=os_, video=video).numpy()
    got = _run(m, x, True, monkeypatch)
    again = m(x.cuda()).cpu().numpy()          # CUDA-graph replay with self re-armed counters: same bits
    assert np.array_equal(again, got)
    base = _run(m, x, False, monkeypatch)
    scale = float(np.abs(ref).max())
    e_chain = float(np.abs(got - ref).max() / scale)
    e_base = float(np.abs(base - ref).max() / scale)
    print("WASP %dx%dx%d os%d %s%s: chain max-rel %.3g, layer-wise %.3g" % (n, hw, hw, os_, precision,
                                                                            " video" if video else "", e_chain, e_base))
    bound = 4e-3 if precision == "fp16" else 3e-2
    assert e_chain < bound and e_base < bound, (e_chain, e_base)
    assert e_chain < 1.5 * e_base + 1e-3
    assert np.array_equal(got, _run(m, x, True, monkeypatch))     # a freshly built plan reproduces the same bits


def test_chain_multi_wave(monkeypatch):
    """More image groups than co-resident CTA pairs: the kernel walks them in waves (forced with 9 clusters)."""
    m, sd = _wasp(False, 16, "fp16", seed=22)
    g = torch.Generator().manual_seed(6)
    x = (torch.randn(8, 2048, 24, 24, generator=g).clamp_min_(0) * 0.5).half().float()
    one = _run(m, x, True, monkeypatch)
    monkeypatch.setenv("UP_CHAIN_MAX_CLUSTERS", "9")
    waves = _run(m, x, True, monkeypatch)
    assert np.array_equal(one, waves)
