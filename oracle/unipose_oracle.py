"""CPU oracle for the UniPose hot path — TEST INFRASTRUCTURE ONLY.

A functional (state_dict in, tensors out) restatement in plain torch CPU ops of what the reference
modules compute.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file; the product (unipose_b200/) never does.

Pinning: `oracle/make_golden.py` runs the REAL reference modules imported from /root/reference on
seeded synthetic weights/inputs and stores their outputs under tests/golden/; tests/test_oracle.py
checks every function below against those fixtures (and, when /root/reference is present, against the
live reference modules).  The reference itself ships no tests or golden vectors, so
"whatever torch 2.11 CPU fp32 produces for the reference graph" is the pin.

Reference citations are file:line under /root/reference.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
BN_EPS = 1e-5       # nn.BatchNorm2d default, used everywhere (resnet.py:11, wasp.py:11, decoder.py:18)
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------------
def batchnorm(x: torch.Tensor, sd: SD, prefix: str, training: bool) -> torch.Tensor:
    """nn.BatchNorm2d: eval -> running stats; train -> batch stats (+ in-place running-stat update)."""
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training=training, momentum=BN_MOMENTUM, eps=BN_EPS)


def conv(x, sd: SD, prefix: str, stride=1, padding=0, dilation=1):
    return F.conv2d(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"), stride=stride, padding=padding,
                    dilation=dilation)


def bottleneck(x, sd: SD, p: str, stride: int, dilation: int, training: bool):
    """Bottleneck.forward — model/modules/backbone/resnet.py:22-42 (ctor :8-20)."""
    out = F.relu(batchnorm(conv(x, sd, p + ".conv1"), sd, p + ".bn1", training))
    out = conv(out, sd, p + ".conv2", stride=stride, padding=dilation, dilation=dilation)
    out = F.relu(batchnorm(out, sd, p + ".bn2", training))
    out = batchnorm(conv(out, sd, p + ".conv3"), sd, p + ".bn3", training)
    if (p + ".downsample.0.weight") in sd:
        x = batchnorm(conv(x, sd, p + ".downsample.0", stride=stride), sd, p + ".downsample.1", training)
    return F.relu(out + x)


def resnet_layout(output_stride: int = 16):
    """(layer name, [(stride, dilation) per block]) — resnet.py:49-58 (strides/dilations), :67-70,
    _make_layer :77-92, _make_MG_unit :94-111 with blocks=[1,2,4]."""
    if output_stride == 16:
        strides, dils = [1, 2, 2, 1], [1, 1, 1, 2]
    elif output_stride == 8:
        strides, dils = [1, 2, 1, 1], [1, 1, 2, 4]
    else:
        raise NotImplementedError
    counts = [3, 4, 23]  # ResNet101: resnet.py:159
    layout = []
    for li in range(3):
        blocks = [(strides[li], dils[li])] + [(1, dils[li])] * (counts[li] - 1)
        layout.append(("layer%d" % (li + 1), blocks))
    mg = [1, 2, 4]
    layout.append(("layer4", [(strides[3], mg[0] * dils[3])] + [(1, m * dils[3]) for m in mg[1:]]))
    return layout


def resnet101_forward(x, sd: SD, prefix: str = "backbone.", output_stride: int = 16, training: bool = False):
    """ResNet.forward — resnet.py:113-124: returns (x, low_level_feat = layer1 output)."""
    x = F.relu(batchnorm(conv(x, sd, prefix + "conv1", stride=2, padding=3), sd, prefix + "bn1", training))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    low = None
    for name, blocks in resnet_layout(output_stride):
        for bi, (s, d) in enumerate(blocks):
            x = bottleneck(x, sd, "%s%s.%d" % (prefix, name, bi), s, d, training)
        if name == "layer1":
            low = x
    return x, low


def wasp_forward(x, sd: SD, prefix: str = "wasp.", output_stride: int = 16, video: bool = False,
                 training: bool = False, dropout_mask: Optional[torch.Tensor] = None):
    """wasp.forward — model/modules/wasp.py:66-90 (ctor :33-64); video variant waspVideo.py:67-91 whose
    global-pool branch has no BatchNorm (waspVideo.py:56-59)."""
    if output_stride == 16:
        dil = [24, 18, 12, 6]
    elif output_stride == 8:
        dil = [48, 36, 24, 12]
    else:
        raise NotImplementedError

    def atrous(t, name, k, d):  # _AtrousModule.forward wasp.py:16-20
        pad = 0 if k == 1 else d
        t = conv(t, sd, prefix + name + ".atrous_conv", padding=pad, dilation=d)
        return F.relu(batchnorm(t, sd, prefix + name + ".bn", training))

    x1 = atrous(x, "aspp1", 1, dil[0])
    x2 = atrous(x1, "aspp2", 3, dil[1])
    x3 = atrous(x2, "aspp3", 3, dil[2])
    x4 = atrous(x3, "aspp4", 3, dil[3])
    # the SAME 1x1 conv2 applied twice to every branch, nothing in between (wasp.py:72-80)
    branches = [conv(conv(t, sd, prefix + "conv2"), sd, prefix + "conv2") for t in (x1, x2, x3, x4)]
    g = F.adaptive_avg_pool2d(x, (1, 1))
    g = conv(g, sd, prefix + "global_avg_pool.1")
    if not video:
        g = batchnorm(g, sd, prefix + "global_avg_pool.2", training)
    g = F.relu(g)
    g = F.interpolate(g, size=x4.shape[2:], mode="bilinear", align_corners=True)
    y = torch.cat(branches + [g], dim=1)
    y = F.relu(batchnorm(conv(y, sd, prefix + "conv1"), sd, prefix + "bn1", training))
    if dropout_mask is not None:  # nn.Dropout(0.5), wasp.py:63,90 — mask supplied by the test, already scaled
        y = y * dropout_mask
    return y


def decoder_forward(x, low, sd: SD, prefix: str = "decoder.", training: bool = False,
                    dropout_masks: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
    """Decoder.forward — model/modules/decoder.py:38-56 (ctor :7-35).  decoder.conv2/bn2 are dead."""
    low = F.relu(batchnorm(conv(low, sd, prefix + "conv1"), sd, prefix + "bn1", training))
    low = F.max_pool2d(low, kernel_size=3, stride=2, padding=1)
    x = F.interpolate(x, size=low.shape[2:], mode="bilinear", align_corners=True)
    y = torch.cat((x, low), dim=1)
    y = F.relu(batchnorm(conv(y, sd, prefix + "last_conv.0", padding=1), sd, prefix + "last_conv.1", training))
    if dropout_masks is not None:
        y = y * dropout_masks[0]
    y = F.relu(batchnorm(conv(y, sd, prefix + "last_conv.4", padding=1), sd, prefix + "last_conv.5", training))
    if dropout_masks is not None:
        y = y * dropout_masks[1]
    return conv(y, sd, prefix + "last_conv.8")


def unipose_forward(x, sd: SD, output_stride: int = 16, stride: int = 8, training: bool = False,
                    video: bool = False, dropout_masks=None):
    """unipose.forward — model/unipose.py:27-38."""
    feat, low = resnet101_forward(x, sd, "backbone.", output_stride, training)
    m = dropout_masks or (None, None, None)
    y = wasp_forward(feat, sd, "wasp.", output_stride, video, training, m[0])
    y = decoder_forward(y, low, sd, "decoder.", training, None if m[1] is None else (m[1], m[2]))
    if stride != 8:
        y = F.interpolate(y, size=x.shape[2:], mode="bilinear", align_corners=True)
    return y


# --------------------------------------------------------------------------------------------
# video variant
# --------------------------------------------------------------------------------------------
def lstm0_forward(x, sd: SD, prefix: str = "lstm_0."):
    """LSTM_0.forward — model/uniposeLSTM.py:16-24."""
    g = torch.tanh(conv(x, sd, prefix + "conv_g_lstm", padding=1))
    i = torch.sigmoid(conv(x, sd, prefix + "conv_i_lstm", padding=1))
    o = torch.sigmoid(conv(x, sd, prefix + "conv_o_lstm", padding=1))
    cell = torch.tanh(g * i)
    return cell, o * cell


def lstm_forward(x, h_prev, c_prev, sd: SD, prefix: str = "lstm."):
    """LSTM.forward — model/uniposeLSTM.py:40-64."""
    def gate(name):
        return conv(x, sd, prefix + "conv_%sx_lstm" % name, padding=1) + \
            conv(h_prev, sd, prefix + "conv_%sh_lstm" % name, padding=1)
    g = torch.tanh(gate("g"))
    o = torch.sigmoid(gate("o"))
    i = torch.sigmoid(gate("i"))
    f = torch.sigmoid(gate("f"))
    cell = f * c_prev + i * g
    return cell, o * torch.tanh(cell)


def middle_cnn_forward(h, sd: SD):
    """conv1..conv5 with ReLU after each, incl. the last — model/uniposeLSTM.py:120-124."""
    y = F.relu(conv(h, sd, "conv1", padding=5))
    y = F.relu(conv(y, sd, "conv2", padding=5))
    y = F.relu(conv(y, sd, "conv3", padding=5))
    y = F.relu(conv(y, sd, "conv4"))
    return F.relu(conv(y, sd, "conv5"))


def unipose_lstm_forward(inp, centermap, it: int, prev_heat, prev_hide, prev_cell, sd: SD,
                         output_stride: int = 16, training: bool = False, dropout_masks=None):
    """uniposeLSTM.unipose.forward — model/uniposeLSTM.py:98-147, generalised over the batch dim
    (the reference hard-codes batch 1 / 46x46 via torch.zeros(1,15,46,46).cuda(), :99-104).
    training=True: BatchNorm on batch statistics, dropout masks supplied by the test (uniposeLSTM.py:102 .train())."""
    frame = inp[:, it]
    heat = unipose_forward(frame, sd, output_stride, stride=8, video=True, training=training,
                           dropout_masks=dropout_masks)
    cm = F.avg_pool2d(centermap[:, it], kernel_size=9, stride=8, padding=1)
    cat = torch.cat((heat, cm), dim=1)
    if it == 0:
        cell, hide = lstm0_forward(cat, sd)
    else:
        cell, hide = lstm_forward(cat, prev_hide, prev_cell, sd)
    return middle_cnn_forward(hide, sd), cell, hide


# --------------------------------------------------------------------------------------------
# architecture spec + deterministic "trained-like" synthetic weights
# --------------------------------------------------------------------------------------------
def param_specs(num_classes: int, video: bool = False, output_stride: int = 16):
    """[(state_dict key, shape, kind)] in the reference's registration order.
    kind in {conv, bn_w, bn_b, bn_rm, bn_rv, bn_nbt, bias}."""
    specs = []

    def add_conv(key, co, ci, k, bias=False):
        specs.append((key + ".weight", (co, ci, k, k), "conv"))
        if bias:
            specs.append((key + ".bias", (co,), "bias"))

    def add_bn(key, c):
        specs.extend([(key + ".weight", (c,), "bn_w"), (key + ".bias", (c,), "bn_b"),
                      (key + ".running_mean", (c,), "bn_rm"), (key + ".running_var", (c,), "bn_rv"),
                      (key + ".num_batches_tracked", (), "bn_nbt")])

    # backbone (resnet.py:61-70)
    add_conv("backbone.conv1", 64, 3, 7)
    add_bn("backbone.bn1", 64)
    inplanes = 64
    for (name, blocks), planes in zip(resnet_layout(output_stride), [64, 128, 256, 512]):
        for bi, (s, _d) in enumerate(blocks):
            p = "backbone.%s.%d" % (name, bi)
            add_conv(p + ".conv1", planes, inplanes, 1)
            add_bn(p + ".bn1", planes)
            add_conv(p + ".conv2", planes, planes, 3)
            add_bn(p + ".bn2", planes)
            add_conv(p + ".conv3", planes * 4, planes, 1)
            add_bn(p + ".bn3", planes * 4)
            if bi == 0 and (s != 1 or inplanes != planes * 4):
                add_conv(p + ".downsample.0", planes * 4, inplanes, 1)
                add_bn(p + ".downsample.1", planes * 4)
            inplanes = planes * 4
    # wasp (wasp.py:46-62 / waspVideo.py:49-66)
    add_conv("wasp.aspp1.atrous_conv", 256, 2048, 1)
    add_bn("wasp.aspp1.bn", 256)
    for i in (2, 3, 4):
        add_conv("wasp.aspp%d.atrous_conv" % i, 256, 256, 3)
        add_bn("wasp.aspp%d.bn" % i, 256)
    add_conv("wasp.global_avg_pool.1", 256, 2048, 1)
    if not video:
        add_bn("wasp.global_avg_pool.2", 256)
    add_conv("wasp.conv1", 256, 1280, 1)
    add_conv("wasp.conv2", 256, 256, 1)
    add_bn("wasp.bn1", 256)
    # decoder (decoder.py:17-30)
    add_conv("decoder.conv1", 48, 256, 1)
    add_bn("decoder.bn1", 48)
    add_conv("decoder.conv2", 256, 2048, 1)   # dead parameters (decoder.py:20-21, never used in forward)
    add_bn("decoder.bn2", 256)
    add_conv("decoder.last_conv.0", 256, 304, 3)
    add_bn("decoder.last_conv.1", 256)
    add_conv("decoder.last_conv.4", 256, 256, 3)
    add_bn("decoder.last_conv.5", 256)
    add_conv("decoder.last_conv.8", num_classes + 1, 256, 1, bias=True)
    if video:  # model/uniposeLSTM.py:81-89
        for g in "gio":
            add_conv("lstm_0.conv_%s_lstm" % g, 15, 15, 3, bias=True)
        for g in ("gx", "ix", "ox", "fx", "gh", "ih", "oh", "fh"):
            add_conv("lstm.conv_%s_lstm" % g, 15, 15, 3, bias=True)
        add_conv("conv1", 128, 15, 11, bias=True)
        add_conv("conv2", 128, 128, 11, bias=True)
        add_conv("conv3", 128, 128, 11, bias=True)
        add_conv("conv4", 128, 128, 1, bias=True)
        add_conv("conv5", 14, 128, 1, bias=True)
    return specs


def synth_state_dict(num_classes: int, video: bool = False, seed: int = 0, output_stride: int = 16,
                     dtype=torch.float32) -> SD:
    """Deterministic synthetic weights with trained-like statistics (no checkpoint is available offline):
    He-normal convs, non-trivial BatchNorm affine + running stats (so BN-folding bugs show), and a small
    gamma on every residual-branch output BN so activations stay O(1) through 33 blocks."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd: SD = {}
    for key, shape, kind in param_specs(num_classes, video, output_stride):
        if kind == "conv":
            fan_in = shape[1] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
        elif kind == "bias":
            t = torch.randn(shape, generator=g) * 0.05
        elif kind == "bn_w":
            t = 0.75 + 0.5 * torch.rand(shape, generator=g)
            if key.endswith(".bn3.weight"):
                t = t * 0.3
        elif kind == "bn_b":
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == "bn_rm":
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == "bn_rv":
            t = 0.6 + 0.8 * torch.rand(shape, generator=g)
        else:  # num_batches_tracked
            sd[key] = torch.zeros((), dtype=torch.long)
            continue
        sd[key] = t.to(dtype)
    return sd


def synth_input(n: int, h: int, w: int, seed: int = 0) -> torch.Tensor:
    """MPII-style normalised image: (uint8 - 128) / 256  (utils/mpii_data.py:184-185)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000 + seed)
    u8 = torch.randint(0, 256, (n, 3, h, w), generator=g)
    return (u8.float() - 128.0) / 256.0
