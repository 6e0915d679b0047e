"""Heat-map decoder: mirrors model/modules/decoder.py of the reference (Decoder :6-64, build_decoder :66-67).

Plan: low-level 1x1 (256->48, stored as 64 channels) + BN + ReLU @H/4 -> max-pool 3/2/1 written straight into
channels [256,320) of the concat buffer; bilinear (align_corners) up-sample of x written into channels [0,256);
3x3 (304->256) + BN + ReLU; 3x3 + BN + ReLU; 1x1 (+bias) whose epilogue stores the fp32 NCHW heat-maps.
decoder.conv2 / bn2 are dead parameters of the reference (decoder.py:20-21): kept for the state_dict, never used.
"""
import torch
import torch.nn as nn

from ... import ops
from ...ops import View
from ...plan_module import PlanModule


class Decoder(PlanModule):
    def __init__(self, dataset, num_classes, backbone, BatchNorm):
        super().__init__()
        low_level_inplanes = 256   # resnet layer1 output
        self.conv1 = nn.Conv2d(low_level_inplanes, 48, 1, bias=False)
        self.bn1 = BatchNorm(48)
        self.conv2 = nn.Conv2d(2048, 256, 1, bias=False)
        self.bn2 = BatchNorm(256)
        self.last_conv = nn.Sequential(
            nn.Conv2d(304, 256, kernel_size=3, stride=1, padding=1, bias=False), BatchNorm(256), nn.ReLU(),
            nn.Dropout(0.5),
            nn.Conv2d(256, 256, kernel_size=3, stride=1, padding=1, bias=False), BatchNorm(256), nn.ReLU(),
            nn.Dropout(0.1),
            nn.Conv2d(256, num_classes + 1, kernel_size=1, stride=1))
        self.num_out = num_classes + 1
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                torch.nn.init.kaiming_normal_(m.weight)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _emit(self, b, x, low, out=None, out_c_total=None):
        n = x.n
        lo = b.act(n, low.h, low.w, 64)
        b.conv(low, b.packed_conv(self.conv1, self.bn1, cout_pad=64), lo, "decoder.low_conv", relu=True)
        hc, wc = (low.h - 1) // 2 + 1, (low.w - 1) // 2 + 1
        cat = b.act(n, hc, wc, 320)
        b.add(lambda: ops.maxpool3x3s2(lo, View(cat, coff=256, c=64)), "decoder.maxpool")
        b.add(lambda: ops.upsample_bilinear_ac(x, View(cat, coff=0, c=256)), "decoder.upsample")
        d1 = b.act(n, hc, wc, 256)
        b.conv(cat, b.packed_conv(self.last_conv[0], self.last_conv[1], cin_pad=320), d1, "decoder.conv_a", pad=1,
               relu=True)
        d2 = b.act(n, hc, wc, 256)
        b.conv(d1, b.packed_conv(self.last_conv[4], self.last_conv[5]), d2, "decoder.conv_b", pad=1, relu=True)
        if out is None:
            out = b.tensor((n, self.num_out, hc, wc))
        b.conv(d2, b.packed_conv(self.last_conv[8], None, nchw_out=True), out, "decoder.head",
               cout_valid=self.num_out, out_c_total=out_c_total)
        return out


def build_decoder(dataset, num_classes, backbone, BatchNorm):
    return Decoder(dataset, num_classes, backbone, BatchNorm)
