"""UniPose-LSTM video model — drop-in for model/uniposeLSTM.py of the reference (LSTM_0 :9-24, LSTM :27-64,
unipose :67-147): same constructors / forward signatures / state_dict keys.

    heat, cell, hide = model(input[B,T,3,H,W], centermap[B,T,1,H,W], iter, previous, previousHide, previousCell)

Per frame: trunk (backbone -> waspVideo -> decoder) writes its K+1 heat-maps straight into the 15-channel
fp32 buffer that the 9x9/8 centre-map pooling completes; one fused ConvLSTM-cell kernel (all gate convs +
sigmoid/tanh + state update); the 11x11 / 1x1 "middle CNN" runs on the tcgen05 conv kernel with bias+ReLU
epilogues.  The reference hard-codes batch 1 and 46x46 (uniposeLSTM.py:99-104); here the batch dimension is free.
"""
import torch
import torch.nn as nn

from .. import engine, ops
from ..plan_module import PlanModule
from .modules.backbone import build_backbone
from .modules.decoder import build_decoder
from .modules.waspVideo import build_wasp


def _nchw_state(t, b, c, h, w, what):
    t = t.detach()
    if t.dim() == 3:
        t = t.unsqueeze(0)
    if tuple(t.shape) != (b, c, h, w):
        raise ValueError('unipose_b200: %s must have shape %s (got %s)' % (what, (b, c, h, w), tuple(t.shape)))
    return t.float().contiguous()


class LSTM_0(PlanModule):
    def __init__(self, inplanes, planes, kernel_size, padding):
        super().__init__()
        assert kernel_size == 3 and padding == 1 and planes <= 16
        self.conv_g_lstm = nn.Conv2d(inplanes, planes, kernel_size=kernel_size, padding=padding)
        self.conv_i_lstm = nn.Conv2d(inplanes, planes, kernel_size=kernel_size, padding=padding)
        self.conv_o_lstm = nn.Conv2d(inplanes, planes, kernel_size=kernel_size, padding=padding)

    def _gates(self):
        return (self.conv_g_lstm, self.conv_i_lstm, self.conv_o_lstm)

    def stacked(self):
        """(weights [3, c, cin, 3, 3], biases [3, c]) in gate order g, i, o - what up_convlstm_cell0_fwd reads."""
        return (torch.stack([c.weight.detach().float() for c in self._gates()]).contiguous(),
                torch.stack([c.bias.detach().float() for c in self._gates()]).contiguous())

    def sources(self):
        return [p for c in self._gates() for p in (c.weight, c.bias)]

    def launch(self, x, cell, hide, stacked=None, gates=None):
        """x fp32 NCHW (contiguous, CUDA) -> cell, hide written in place.  `stacked`: pre-stacked gate weights (plans
        keep them in static buffers refreshed by a pack job, so a captured graph never bakes a temporary's address)."""
        w3, b3 = stacked if stacked is not None else self.stacked()
        b, cin, h, w = x.shape
        ops._lib.call("up_convlstm_cell0_fwd", ops._ptr(x), ops._ptr(w3), ops._ptr(b3), ops._ptr(cell), ops._ptr(hide),
                      b, cin, w3.shape[1], h, w, ops._ptr(gates), ops._stream())

    def forward(self, x):
        ops.require_cuda(x, "LSTM_0 input")
        x = x.detach().float().contiguous()
        planes = self.conv_g_lstm.out_channels
        cell = torch.empty((x.shape[0], planes, x.shape[2], x.shape[3]), device=x.device)
        hide = torch.empty_like(cell)
        with torch.cuda.device(x.device):
            self.launch(x, cell, hide)
        return cell, hide


class LSTM(PlanModule):
    def __init__(self, inplanes, planes, kernel_size, padding):
        super().__init__()
        assert kernel_size == 3 and padding == 1 and planes <= 16
        for g in ('gx', 'ix', 'ox', 'fx'):
            setattr(self, 'conv_%s_lstm' % g, nn.Conv2d(inplanes, planes, kernel_size=kernel_size, padding=padding))
        for g in ('gh', 'ih', 'oh', 'fh'):
            setattr(self, 'conv_%s_lstm' % g, nn.Conv2d(planes, planes, kernel_size=kernel_size, padding=padding))

    def _stacked(self, suffix):
        convs = [getattr(self, 'conv_%s%s_lstm' % (g, suffix)) for g in 'giof']
        return (torch.stack([c.weight.detach().float() for c in convs]).contiguous(),
                torch.stack([c.bias.detach().float() for c in convs]).contiguous())

    def stacked(self):
        return self._stacked('x') + self._stacked('h')

    def sources(self):
        return [p for sfx in 'xh' for g in 'giof' for p in (getattr(self, 'conv_%s%s_lstm' % (g, sfx)).weight,
                                                            getattr(self, 'conv_%s%s_lstm' % (g, sfx)).bias)]

    def launch(self, x, h_prev, c_prev, cell, hide, stacked=None, gates=None):
        wx, bx, wh, bh = stacked if stacked is not None else self.stacked()
        b, cin, h, w = x.shape
        ops._lib.call("up_convlstm_cell_fwd", ops._ptr(x), ops._ptr(h_prev), ops._ptr(c_prev), ops._ptr(wx),
                      ops._ptr(bx), ops._ptr(wh), ops._ptr(bh), ops._ptr(cell), ops._ptr(hide), b, cin, wx.shape[1],
                      h, w, ops._ptr(gates), ops._stream())

    def forward(self, x, prevHide, prevCell):
        ops.require_cuda(x, "LSTM input")
        x = x.detach().float().contiguous()
        b, _, h, w = x.shape
        planes = self.conv_gx_lstm.out_channels
        hp = _nchw_state(prevHide, b, planes, h, w, 'prevHide')
        cp = _nchw_state(prevCell, b, planes, h, w, 'prevCell')
        cell = torch.empty_like(hp)
        hide = torch.empty_like(hp)
        with torch.cuda.device(x.device):
            self.launch(x, hp, cp, cell, hide)
        return cell, hide


class unipose(PlanModule):
    def __init__(self, backbone='resnet', output_stride=16, num_classes=21, sync_bn=True, freeze_bn=False,
                 stride=8, precision=None):
        super().__init__()
        self.stride = stride
        self.precision = precision
        self.BatchNorm = nn.BatchNorm2d
        self.pool_center = nn.AvgPool2d(kernel_size=9, stride=8, padding=1)
        self.backbone = build_backbone(backbone, output_stride, self.BatchNorm)
        self.wasp = build_wasp(backbone, output_stride, self.BatchNorm)
        self.decoder = build_decoder("Penn_Action", num_classes, backbone, self.BatchNorm)
        self.lstm_0 = LSTM_0(15, 15, 3, 1)
        self.lstm = LSTM(15, 15, 3, 1)
        # "middle CNN" (uniposeLSTM.py:85-89)
        self.conv1 = nn.Conv2d(15, 128, kernel_size=11, padding=5)
        self.conv2 = nn.Conv2d(128, 128, kernel_size=11, padding=5)
        self.conv3 = nn.Conv2d(128, 128, kernel_size=11, padding=5)
        self.conv4 = nn.Conv2d(128, 128, kernel_size=1, padding=0)
        self.conv5 = nn.Conv2d(128, 14, kernel_size=1, padding=0)
        if freeze_bn:
            self.freeze_bn()

    # ------------------------------------------------------------------------------------------------------------
    # kernel plans
    # ------------------------------------------------------------------------------------------------------------
    def _emit_trunk(self, b, frames, cmaps, n, h, w):
        """backbone -> waspVideo -> decoder for `n` frames; the K+1 heat-maps land in channels [0, K+1) of the fp32
        [n, 15, h/8, w/8] buffer that the 9x9/8 centre-map pooling completes (torch.cat of uniposeLSTM.py:116)."""
        k1 = self.decoder.num_out
        lstm_c = self.lstm_0.conv_g_lstm.in_channels    # 15 = K+1 + centre map
        hs, ws = h // 8, w // 8
        x, low = self.backbone._emit_image(b, frames)
        x = self.wasp._emit(b, x)
        cat = b.tensor((n, lstm_c, hs, ws), zero=True)
        self.decoder._emit(b, x, low, out=cat, out_c_total=lstm_c)
        b.add(lambda: ops._lib.call("up_avgpool9s8p1_f32", ops._ptr(cmaps), ops._ptr(cat), n, 1, h, w, hs, ws, lstm_c,
                                    k1, ops._stream()), "pool_center")
        return cat

    def _emit_recurrent(self, b, plan, cat, b_, hs, ws, first):
        """ConvLSTM cell (LSTM_0 for the first frame) + the 11x11 / 1x1 "middle CNN" (uniposeLSTM.py:118-124)."""
        planes = self.lstm_0.conv_g_lstm.out_channels
        cell = b.tensor((b_, planes, hs, ws))
        hide = b.tensor((b_, planes, hs, ws))
        cell_mod = self.lstm_0 if first else self.lstm
        gate_w = [torch.empty_like(t) for t in cell_mod.stacked()]      # static: refreshed in place by a pack job

        def restack():
            for dst, src in zip(gate_w, cell_mod.stacked()):
                dst.copy_(src)
        plan.pack_jobs.append(engine._PackJob(cell_mod.sources(), restack))
        if first:
            b.add(lambda: self.lstm_0.launch(cat, cell, hide, stacked=gate_w), "lstm_0")
        else:
            hp = plan.static_input((b_, planes, hs, ws))
            cp = plan.static_input((b_, planes, hs, ws))
            b.add(lambda: self.lstm.launch(cat, hp, cp, cell, hide, stacked=gate_w), "lstm")
        hin = b.act(b_, hs, ws, 16, zero=True)
        b.add(lambda: ops.nchw_to_act(hide, hin), "hide_to_nhwc")
        a1 = b.act(b_, hs, ws, 128)
        b.conv(hin, b.packed_conv(self.conv1, None, cin_pad=16), a1, "middle.conv1", pad=5, relu=True)
        a2 = b.act(b_, hs, ws, 128)
        b.conv(a1, b.packed_conv(self.conv2, None), a2, "middle.conv2", pad=5, relu=True)
        a3 = b.act(b_, hs, ws, 128)
        b.conv(a2, b.packed_conv(self.conv3, None), a3, "middle.conv3", pad=5, relu=True)
        a4 = b.act(b_, hs, ws, 128)
        b.conv(a3, b.packed_conv(self.conv4, None), a4, "middle.conv4", relu=True)
        heat = b.tensor((b_, self.conv5.out_channels, hs, ws))
        b.conv(a4, b.packed_conv(self.conv5, None, nchw_out=True), heat, "middle.conv5", relu=True)
        return heat, cell, hide

    def _build_plan(self, b_, h, w, first, device):
        """One frame end to end (the reference's per-call work): trunk + recurrent part."""
        plan = engine.Plan(device, self._precision())
        b = plan.builder
        frame = plan.static_input((b_, 3, h, w))
        cmap = plan.static_input((b_, 1, h, w))
        cat = self._emit_trunk(b, frame, cmap, b_, h, w)
        plan.finalize(list(self._emit_recurrent(b, plan, cat, b_, h // 8, w // 8, first)))
        return plan

    def _build_trunk_plan(self, n, h, w, device):
        """Temporal batching: the trunk of ALL T frames of a clip as one batch of B*T images - it does
        not depend on the recurrence (uniposeLSTM.py:106-118)."""
        plan = engine.Plan(device, self._precision())
        b = plan.builder
        frames = plan.static_input((n, 3, h, w))
        cmaps = plan.static_input((n, 1, h, w))
        plan.finalize([self._emit_trunk(b, frames, cmaps, n, h, w)])
        return plan

    def _build_step_plan(self, b_, hs, ws, first, device):
        plan = engine.Plan(device, self._precision())
        b = plan.builder
        cat = plan.static_input((b_, self.lstm_0.conv_g_lstm.in_channels, hs, ws))
        plan.finalize(list(self._emit_recurrent(b, plan, cat, b_, hs, ws, first)))
        return plan

    def _clip_features(self, input, centermap):
        """[T*B, 15, h/8, w/8] trunk output of the whole clip (frame-major), cached per (input, centermap) tensor."""
        b_, t_, _c, h, w = input.shape
        key = (input.data_ptr(), input._version, centermap.data_ptr(), centermap._version, tuple(input.shape),
               self._precision(), engine._RAW_UPDATE_EPOCH[0],
               tuple(p._version for p in (self.conv1.weight, self.decoder.last_conv[8].weight, self.backbone.conv1.weight)))
        cache = self.__dict__.get("_clip_cache")
        if cache is not None and cache[0] == key:
            return cache[1]
        pkey = ("trunk", b_ * t_, h, w, self._precision(), input.device.index)
        plan = self._plans.get(pkey)
        if plan is None:
            plan = self._build_trunk_plan(b_ * t_, h, w, input.device)
            self._plans[pkey] = plan
        frames = input.detach().float().transpose(0, 1).reshape(t_ * b_, 3, h, w)
        cmaps = centermap.detach().float().transpose(0, 1).reshape(t_ * b_, 1, h, w)
        feats = plan.run(frames, cmaps)[0]
        object.__setattr__(self, "_clip_cache", (key, feats))
        return feats

    def forward(self, input, centermap, iter, previous, previousHide, previousCell):
        import os
        self._check_inputs([input, centermap])
        if self.training:
            # train mode: the frame's training plan, chained to the previous frames by autograd through (cell, hide)
            from .. import train
            return train.forward_train_video(self, input, centermap, int(iter), previousHide, previousCell)
        self._bn_eval_only()
        b_, t_, _c, h, w = input.shape
        first = (iter == 0)
        planes = self.lstm_0.conv_g_lstm.out_channels
        states = ()
        if not first:
            states = (_nchw_state(previousHide, b_, planes, h // 8, w // 8, 'previousHide'),
                      _nchw_state(previousCell, b_, planes, h // 8, w // 8, 'previousCell'))
        if os.environ.get("UNIPOSE_B200_TEMPORAL_BATCH", "1") != "0" and t_ > 1:
            # the clip's trunk runs once (first call with this input tensor); every call then only advances the
            # ConvLSTM + middle CNN by one frame - the reference's per-frame signature is a view onto the cached trunk
            feats = self._clip_features(input, centermap)
            key = ("step", b_, h, w, first, self._precision(), input.device.index)
            plan = self._plans.get(key)
            if plan is None:
                plan = self._build_step_plan(b_, h // 8, w // 8, first, input.device)
                self._plans[key] = plan
            outs = plan.run(feats[iter * b_:(iter + 1) * b_], *states)
        else:
            key = (b_, h, w, first, self._precision(), input.device.index)
            plan = self._plans.get(key)
            if plan is None:
                plan = self._build_plan(b_, h, w, first, input.device)
                self._plans[key] = plan
            outs = plan.run(input[:, iter].detach().float(), centermap[:, iter].detach().float(), *states)
        heat, cell, hide = [o.clone() for o in outs]
        return heat, cell, hide

    def freeze_bn(self):
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()

    def _lr_params(self, roots):
        for root in roots:
            for m in root.modules():
                if isinstance(m, (nn.Conv2d, nn.BatchNorm2d)):
                    for p in m.parameters(recurse=False):
                        if p.requires_grad:
                            yield p

    def get_1x_lr_params(self):
        return self._lr_params([self.backbone])

    def get_10x_lr_params(self):
        return self._lr_params([self.wasp, self.decoder])
