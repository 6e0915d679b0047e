"""Host logic of the inference plans, without a GPU: a plan can be EMITTED on CPU tensors (buffers, launch list, weight
tables; nothing is launched), so the fusion decisions of the backbone are checked here - which launches a
configuration produces, and that the switches documented in INTEGRATION.md select the layer-wise forms."""
import warnings

import pytest
import torch

from unipose_b200 import _lib, engine
from unipose_b200.model.unipose import unipose


def _names(monkeypatch, precision="fp16", shape=(2, 3, 128, 128), env=(), **kw):
    for k, v in env:
        monkeypatch.setenv(k, v)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = unipose(dataset="MPII", num_classes=16, precision=precision, **kw).eval()
    plan = m._build_plan(shape, torch.device("cpu"))
    return [n for n, f, s in plan.ops if f is not None], plan, m


def _needs_lib():
    try:
        _lib.load()
    except Exception as e:      # pragma: no cover - the driver builds the library before the CPU suite
        pytest.skip("libunipose_b200.so not built: %s" % e)


def test_default_fp16_plan_fuses_tails_projections_and_next_conv1(monkeypatch):
    _needs_lib()
    names, plan, m = _names(monkeypatch)
    # layer1: conv1 of block 0, then three tail kernels that also produce the next block's conv1 (layer1.1, layer1.2,
    # layer2.0); layer2 blocks 1..3: conv1 + tail; the first block of layers 2-4: conv2 + (conv3 + projection)
    assert names[:3] == ["pack_input_s2d", "stem", "maxpool"]
    assert names[3:9] == ["bottleneck.conv1", "bottleneck.tail+conv1", "bottleneck.tail+conv1", "bottleneck.tail+conv1",
                          "bottleneck.conv2", "bottleneck.conv3+proj"]
    assert names.count("bottleneck.tail") == 3 and names.count("bottleneck.conv3+proj") == 3
    assert names.count("bottleneck.downsample") == 0
    # 33 bottlenecks, three of which get their conv1 from the previous tail kernel
    assert names.count("bottleneck.conv1") == 30
    assert names[-6:] == ["decoder.low_conv", "decoder.maxpool", "decoder.upsample", "decoder.conv_a", "decoder.conv_b",
                          "decoder.head"]


@pytest.mark.parametrize("env,tails,tail_c1,proj,down", [
    ((("UNIPOSE_B200_TAIL_CONV1", "0"),), 6, 0, 3, 0),
    ((("UNIPOSE_B200_PROJ_FUSE", "0"),), 3, 3, 0, 3),                       # layer1.0's projection stays in its tail
    ((("UNIPOSE_B200_PROJ_FUSE", "0"), ("UNIPOSE_B200_BNECK_TAIL_PROJ", "0")), 3, 3, 0, 4),
    ((("UNIPOSE_B200_BNECK_TAIL", "0"),), 0, 0, 4, 0),                      # layer1.0 then takes the conv3 + projection GEMM
    ((("UNIPOSE_B200_BNECK_TAIL", "0"), ("UNIPOSE_B200_PROJ_FUSE", "0")), 0, 0, 0, 4),
])
def test_switches_select_the_layerwise_forms(monkeypatch, env, tails, tail_c1, proj, down):
    _needs_lib()
    names, _, _ = _names(monkeypatch, env=env)
    assert names.count("bottleneck.tail") == tails, names
    assert names.count("bottleneck.tail+conv1") == tail_c1, names
    assert names.count("bottleneck.conv3+proj") == proj, names
    assert names.count("bottleneck.downsample") == down, names
    blocks_with_own_conv1 = 33 - tail_c1
    assert names.count("bottleneck.conv1") == blocks_with_own_conv1


def test_parity_mode_plan_is_layerwise(monkeypatch):
    """fp32-grade (bf16 x 3 split) mode: no fused kernels - every bottleneck is conv1, conv2, conv3 (+ downsample)."""
    _needs_lib()
    names, _, _ = _names(monkeypatch, precision="fp32")
    assert not any(n.startswith("bottleneck.tail") or n.endswith("+proj") or n.startswith("bottleneck.chain") for n in names)
    assert names.count("bottleneck.conv1") == names.count("bottleneck.conv2") == names.count("bottleneck.conv3") == 33
    assert names.count("bottleneck.downsample") == 4


def test_output_stride8_still_fuses_three_projections(monkeypatch):
    _needs_lib()
    names, _, _ = _names(monkeypatch, output_stride=8)
    assert names.count("bottleneck.conv3+proj") == 3 and names.count("bottleneck.downsample") == 0


def test_fused_projection_filter_is_packed_into_disjoint_slices(monkeypatch):
    """conv3 + projection: ONE filter buffer [1 + inplanes / planes][cout][planes]; conv3 fills slice 0, the downsample
    filter the others (one pack job per input-channel slice), and the epilogue shift is the sum of both BatchNorm shifts
    (a torch-side pack job that re-runs with every refresh of the weight table)."""
    _needs_lib()
    names, plan, m = _names(monkeypatch)
    blk = m.backbone.layer2[0]
    wd = blk.downsample[0].weight
    jobs = [e for e in plan.weights.entries if e["src"]() is wd]
    planes, cout = blk.conv1.out_channels, 4 * blk.conv1.out_channels
    assert len(jobs) == wd.shape[1] // planes == 2
    assert [e["ci_off"] for e in jobs] == [0, planes] and all(e["cin_slice"] == planes for e in jobs)
    assert all((e["rows"], e["cols"]) == (cout, planes) for e in jobs)
    base = jobs[0]["out"].data_ptr() - cout * planes * jobs[0]["out"].element_size()      # slice 0 = conv3's filter
    c3 = [e for e in plan.weights.entries if e["src"]() is blk.conv3.weight]
    assert len(c3) == 1 and c3[0]["out"].data_ptr() == base
    ptrs = sorted(e["out"].data_ptr() for e in c3 + jobs)
    assert [p - base for p in ptrs] == [i * cout * planes * 2 for i in range(3)]
    assert len(plan.pack_jobs) == 3 and all(j.with_table for j in plan.pack_jobs)    # one shift-sum job per fused projection,
    watched = set(id(t) for t in plan.weights._watched())                            # re-run with every table refresh
    assert id(blk.bn3.bias) in watched and id(blk.downsample[1].running_var) in watched


@pytest.mark.parametrize("freeze", [False, True])
def test_train_plan_has_a_gradient_slot_for_every_live_parameter(freeze):
    """The training plan (forward list + reverse tape) also emits on CPU tensors.  Every parameter the reference's
    backward reaches gets a slot in the flat gradient; the only parameters without one are the reference's own dead
    ones (`Decoder.conv2` / `bn2`, defined at model/modules/decoder.py:20-21 but commented out of forward(), :43-45 -
    autograd leaves their .grad None and Adam skips them).  Frozen BatchNorm layers keep their affine gradients
    (eval-mode BatchNorm still trains gamma / beta) but drop out of the statistics update list."""
    _needs_lib()
    from unipose_b200 import train
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = unipose(dataset="MPII", num_classes=16, precision="bf16", freeze_bn=freeze).train()
        if freeze:
            m.freeze_bn()
    tp = train.build_image_train_plan(m, (2, 3, 96, 96), torch.device("cpu"), "bf16", flat=True)
    have = set(id(p) for p in tp.params)
    missing = sorted(n for n, p in m.named_parameters() if p.requires_grad and id(p) not in have)
    assert missing == ["decoder.bn2.bias", "decoder.bn2.weight", "decoder.conv2.weight"]
    live = sum(p.numel() for n, p in m.named_parameters() if p.requires_grad and not n.startswith(("decoder.conv2", "decoder.bn2")))
    assert tp.flat_g.numel() == live == 47_022_513           # 188 MB of fp32 gradient per step
    assert len(set(id(p) for p in tp.params)) == len(tp.params)
    n_bn = sum(1 for mod in m.modules() if isinstance(mod, torch.nn.BatchNorm2d)) - 1        # decoder.bn2 never runs
    assert len(tp.bn_modules) == (0 if freeze else n_bn)
    assert len(tp.fwd) > 300 and len(tp.bwd) > len(tp.fwd)


def test_video_plans_share_the_trunk_emission_and_split_at_the_recurrence(monkeypatch):
    """UniPose-LSTM: the per-frame plan = trunk (same fused backbone as the image model, waspVideo, decoder writing
    the first K+1 channels of the 15-channel ConvLSTM input) + centre-map pooling + cell + middle CNN; temporal
    batching splits it into a trunk plan over all T*B frames and a per-frame step plan (uniposeLSTM.py:106-147)."""
    _needs_lib()
    from unipose_b200.model import uniposeLSTM
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = uniposeLSTM.unipose(num_classes=13, precision="fp16").eval()
    dev = torch.device("cpu")
    frame = [n for n, f, s in m._build_plan(2, 128, 128, True, dev).ops if f is not None]
    trunk = [n for n, f, s in m._build_trunk_plan(6, 128, 128, dev).ops if f is not None]
    step0 = [n for n, f, s in m._build_step_plan(2, 16, 16, True, dev).ops if f is not None]
    step = [n for n, f, s in m._build_step_plan(2, 16, 16, False, dev).ops if f is not None]
    assert trunk[-1] == "pool_center" and trunk.count("bottleneck.tail+conv1") == 3 and trunk.count("bottleneck.conv3+proj") == 3
    assert step0 == ["lstm_0", "hide_to_nhwc", "middle.conv1", "middle.conv2", "middle.conv3", "middle.conv4", "middle.conv5"]
    assert step == ["lstm"] + step0[1:]
    assert frame == trunk + step0                      # same launches, one plan instead of two


def test_pack_job_refresh_rules():
    """_PackJob: own sources are compared by (storage, version); with_table jobs also re-run whenever the weight table
    has refreshed, and run once on the first call either way."""
    runs = []
    src = torch.zeros(4)
    j = engine._PackJob([src], lambda: runs.append("a"))
    assert j.refresh() and not j.refresh() and not j.refresh(table_changed=True)
    src.add_(1)                                    # in-place update bumps the version counter
    assert j.refresh() and runs == ["a", "a"]
    engine.note_raw_parameter_update()             # raw-pointer kernels (Adam, BN statistics) bump the global epoch
    assert j.refresh() and not j.refresh()
    t = engine._PackJob((), lambda: runs.append("t"), with_table=True)
    assert t.refresh(False) and not t.refresh(False) and t.refresh(True) and t.refresh(True) and not t.refresh(False)
    assert runs.count("t") == 3


def test_param_lookup_follows_replaced_parameters():
    conv = torch.nn.Conv2d(4, 4, 1, bias=False)
    bn = torch.nn.BatchNorm2d(4)
    assert engine._param(conv, "weight") is conv.weight and engine._param(bn, "running_var") is bn.running_var
    conv.weight = torch.nn.Parameter(torch.ones(4, 4, 1, 1))         # e.g. load_state_dict(assign=True)
    assert engine._param(conv, "weight") is conv.weight
    bn.running_mean = torch.ones(4)
    assert engine._param(bn, "running_mean") is bn.running_mean
