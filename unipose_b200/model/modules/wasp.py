"""WASP "waterfall" atrous module on tcgen05 implicit-GEMM convs.

Mirrors model/modules/wasp.py of the reference (_AtrousModule :6-31, wasp :33-104, build_wasp :106-107).
Kernel plan for x [N, h, w, 2048] (all NHWC 16-bit, BN folded into the conv epilogues in eval mode):

    aspp1 1x1 2048->256          -> S[0:N]          (S stacks the four cascade outputs along n)
    aspp2/3/4 3x3 d18/12/6       -> S[N:2N], S[2N:3N], S[3N:4N]   (taps that fall outside the map are skipped)
    GAP -> 1x1 -> BN -> ReLU -> broadcast                -> S[4N:5N]   (side stream, overlaps the cascade)
    conv1' 1x1 1280->256 + bn1 + ReLU reads S as a 5-group K-split (no concat buffer) -> out, where
    conv1' = conv1 with the shared conv2 applied twice folded into its weights (eval only; the training plan in
    train.py keeps conv2 as two launches over M = 4N*h*w)
"""
import torch
import torch.nn as nn

from ... import ops
from ...ops import View
from ...plan_module import PlanModule


class _AtrousModule(PlanModule):
    def __init__(self, inplanes, planes, kernel_size, padding, dilation, BatchNorm):
        super().__init__()
        self.atrous_conv = nn.Conv2d(inplanes, planes, kernel_size=kernel_size, stride=1, padding=padding,
                                     dilation=dilation, bias=False)
        self.bn = BatchNorm(planes)
        self._out_channels = (planes,)
        _kaiming_init(self)

    def _emit(self, b, x, y=None):
        conv = self.atrous_conv
        if y is None:
            y = b.act(x.n, x.h, x.w, conv.out_channels)
        b.conv(x, b.packed_conv(conv, self.bn), y, "wasp.atrous", dil=conv.dilation[0], pad=conv.padding[0], relu=True)
        return y


def _kaiming_init(module):
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            torch.nn.init.kaiming_normal_(m.weight)
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.data.fill_(1)
            m.bias.data.zero_()


class wasp(PlanModule):
    _gap_has_bn = True

    def __init__(self, backbone, output_stride, BatchNorm):
        super().__init__()
        inplanes = self._inplanes(backbone)
        if output_stride == 16:
            dilations = [24, 18, 12, 6]
        elif output_stride == 8:
            dilations = [48, 36, 24, 12]
        else:
            raise NotImplementedError
        self.aspp1 = _AtrousModule(inplanes, 256, 1, padding=0, dilation=dilations[0], BatchNorm=BatchNorm)
        self.aspp2 = _AtrousModule(256, 256, 3, padding=dilations[1], dilation=dilations[1], BatchNorm=BatchNorm)
        self.aspp3 = _AtrousModule(256, 256, 3, padding=dilations[2], dilation=dilations[2], BatchNorm=BatchNorm)
        self.aspp4 = _AtrousModule(256, 256, 3, padding=dilations[3], dilation=dilations[3], BatchNorm=BatchNorm)
        gap = [nn.AdaptiveAvgPool2d((1, 1)), nn.Conv2d(inplanes, 256, 1, stride=1, bias=False)]
        if self._gap_has_bn:
            gap.append(nn.BatchNorm2d(256))
        gap.append(nn.ReLU())
        self.global_avg_pool = nn.Sequential(*gap)
        self.conv1 = nn.Conv2d(1280, 256, 1, bias=False)
        self.conv2 = nn.Conv2d(256, 256, 1, bias=False)
        self.bn1 = BatchNorm(256)
        self.dropout = nn.Dropout(0.5)
        self._out_channels = (256,)
        _kaiming_init(self)

    @staticmethod
    def _inplanes(backbone):
        return 2048

    def _folded_conv1_weight(self, w1):
        """Inference-only algebra: conv1(cat(conv2(conv2(x_i)), x5)) == conv1'(cat(x_i, x5)) with
        W1'_i = W1_i . W2 . W2 (fp32 on the host side of the pack job) - the two shared-1x1 launches over 4N*h*w
        pixels disappear (wasp.py:72-88).  Training keeps the unfolded graph (train.py)."""
        w2 = self.conv2.weight.detach().float()[:, :, 0, 0]
        w22 = w2 @ w2
        w1m = w1.float()[:, :, 0, 0]
        parts = [w1m[:, i * 256:(i + 1) * 256] @ w22 for i in range(4)] + [w1m[:, 1024:]]
        return torch.cat(parts, dim=1)[:, :, None, None].contiguous()

    # ---- the whole block as one persistent kernel (fp16 / bf16 inference) ------------------------------------
    def _chain_desc(self, b, x):
        import ctypes
        import os
        from ... import _lib
        if os.environ.get("UNIPOSE_B200_WASP_CHAIN", "1") == "0" or b.mode == ops.UP_SPLIT:
            return None
        d = _lib.UpWaspChainDesc()
        d.n, d.h, d.w, d.cin = x.n, x.h, x.w, x.c
        for i, a in enumerate((self.aspp2, self.aspp3, self.aspp4)):
            d.dil[i] = a.atrous_conv.dilation[0]
        d.dtype = b.mode
        d.conv1_cin = 1280
        if _lib.load().up_wasp_chain_supported(ctypes.byref(d)) != 0:
            return None          # e.g. odd batch, tiny maps: the layer-wise plan below handles every shape
        return d

    def _emit_chain(self, b, x, d):
        """aspp1..4, the folded conv1 and the pooling branch as ONE launch of up_wasp_chain_fwd (csrc/wasp_chain.cu)."""
        import ctypes
        from ... import _lib
        n, h, w = x.n, x.h, x.w
        S = b.act(4 * n, h, w, 256)          # x1..x4: halo sources of the cascade
        out = b.act(n, h, w, 256)
        pcs = [b.packed_conv(a.atrous_conv, a.bn) for a in (self.aspp1, self.aspp2, self.aspp3, self.aspp4)]
        pc1 = b.packed_conv(self.conv1, self.bn1, weight_fn=self._folded_conv1_weight, extra_sources=[self.conv2.weight])
        gap_bn = self.global_avg_pool[2] if self._gap_has_bn else None
        gap_t, shift_gap = b.packed_transposed(self.global_avg_pool[1], gap_bn)
        pool_t, _ = b.packed_transposed(self.conv1, None, ci_off=1024, cin_slice=256, weight_fn=self._folded_conv1_weight,
                                        fold=pc1.fold, fold_period=256, extra_sources=[self.conv2.weight])
        ws_bytes = int(_lib.load().up_wasp_chain_workspace_bytes(ctypes.byref(d)))
        ws = b.tensor((ws_bytes,), dtype=torch.uint8, zero=True)      # counters start at zero; the kernel re-arms them
        wts = _lib.UpWaspChainWeights()
        for i, pc in enumerate(pcs):
            wts.aspp[i] = pc.w.data_ptr()
            wts.shift[i] = pc.shift.data_ptr()
        wts.conv1, wts.shift1 = pc1.w.data_ptr(), pc1.shift.data_ptr()
        wts.gap_t, wts.shift_gap = gap_t.data_ptr(), shift_gap.data_ptr()
        wts.conv1_pool_t = pool_t.data_ptr()
        keep = (pcs, pc1, gap_t, shift_gap, pool_t, ws)

        def launch(keep=keep):
            _lib.call("up_wasp_chain_fwd", ctypes.byref(d), ctypes.byref(wts), x.ptr(), S.ptr(), out.ptr(), ops._ptr(ws),
                      ws_bytes, ops._stream())
        b.add(launch, "wasp.chain")
        return out

    def _emit(self, b, x):
        d = self._chain_desc(b, x) if isinstance(x, ops.Act) else None
        if d is not None:
            return self._emit_chain(b, x, d)
        n, h, w = x.n, x.h, x.w
        S = b.act(5 * n, h, w, 256)      # the four cascade outputs + the broadcast pooling branch, stacked along n
        branch = [View(S, n_off=i * n, n=n) for i in range(5)]
        # image-level branch on the side stream: GAP -> 1x1 (-> BN) -> ReLU -> bilinear from 1x1 == broadcast
        b.fork()
        g = b.act(n, 1, 1, x.c)
        b.add(lambda: ops.global_avgpool(x, g), "wasp.gap", side=True)
        g2 = b.act(n, 1, 1, 256)
        gap_bn = self.global_avg_pool[2] if self._gap_has_bn else None
        b.conv(g, b.packed_conv(self.global_avg_pool[1], gap_bn), g2, "wasp.gap_conv", side=True, relu=True)
        b.add(lambda: ops.broadcast_hw(g2, branch[4]), "wasp.gap_broadcast", side=True)
        # the waterfall: 1x1, then three dilated 3x3 feeding each other (wasp.py:67-70)
        self.aspp1._emit(b, x, branch[0])
        self.aspp2._emit(b, branch[0], branch[1])
        self.aspp3._emit(b, branch[1], branch[2])
        self.aspp4._emit(b, branch[2], branch[3])
        b.join()
        out = b.act(n, h, w, 256)
        pc1 = b.packed_conv(self.conv1, self.bn1, weight_fn=self._folded_conv1_weight,
                            extra_sources=[self.conv2.weight])
        b.conv(branch[0], pc1, out, "wasp.conv1", relu=True, x_groups=5, x_group_nstride=n)
        return out   # nn.Dropout(0.5) is the identity in eval mode (wasp.py:90)


def build_wasp(backbone, output_stride, BatchNorm):
    return wasp(backbone, output_stride, BatchNorm)
