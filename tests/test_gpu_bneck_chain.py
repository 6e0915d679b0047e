"""The fused bottleneck-run kernel (csrc/bneck_chain.cu: layer3 blocks 1..22 as one persistent launch) against the
layer-wise plan it replaces and the CPU oracle, on the map sizes the network produces (24x24 image-pair tiles,
32x32 / 16x16 single-image tiles, 48x48 at output_stride 8 with dilation 2)."""
import warnings

import numpy as np
import pytest
import torch

from oracle import unipose_oracle as O

pytestmark = pytest.mark.gpu


def _model(precision, output_stride=16, seed=0):
    from unipose_b200.model.unipose import unipose
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = unipose(dataset="MPII", num_classes=16, precision=precision, output_stride=output_stride)
    sd = O.synth_state_dict(16, seed=seed, output_stride=output_stride)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd


def _run(m, x, chain, monkeypatch):
    monkeypatch.setenv("UNIPOSE_B200_BNECK_CHAIN", "1" if chain else "0")
    m._plans.clear()
    out = m(x.cuda())
    torch.cuda.synchronize()
    names = [n for n, f, s in m.plan_for(x.cuda()).ops if f is not None]
    assert any(n.startswith("bottleneck.chain") for n in names) == chain, names
    return out.cpu().numpy()


@pytest.mark.parametrize("n,size,os_,precision", [(4, 384, 16, "fp16"), (2, 512, 16, "fp16"), (2, 256, 16, "fp16"),
                                                  (4, 384, 16, "bf16"), (2, 384, 8, "fp16"), (4, 368, 16, "fp16")])
def test_bneck_chain_matches_layerwise_and_oracle(n, size, os_, precision, monkeypatch):
    m, sd = _model(precision, os_, seed=11)
    x = O.synth_input(n, size, size, seed=11)
    with torch.no_grad():
        ref = O.unipose_forward(x, sd, output_stride=os_).numpy()
    got = _run(m, x, True, monkeypatch)
    again = m(x.cuda()).cpu().numpy()              # graph replay, self re-armed counters
    assert np.array_equal(got, again)
    base = _run(m, x, False, monkeypatch)
    scale = float(np.abs(ref).max())
    e_chain = float(np.abs(got - ref).max() / scale)
    e_base = float(np.abs(base - ref).max() / scale)
    print("bottleneck chain %dx%d^2 os%d %s: max-rel %.3g (layer-wise %.3g)" % (n, size, os_, precision, e_chain, e_base))
    bound = 5e-3 if precision == "fp16" else 3e-2
    assert e_chain < bound and e_base < bound, (e_chain, e_base)
    assert e_chain < 2.0 * e_base + 1e-3
