"""Run a few representative conv launches standalone (for `ncu --set full -k regex:conv_tcgen05`).

    python tools/prof_convs.py [--reps 3] [--prec fp16] [--only name]
"""
from __future__ import annotations

import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from unipose_b200 import ops  # noqa: E402

# name, n, h, w, cin, cout, k, stride, dil, residual, cin_pad_override
CONVS = [
    ("l1_conv2_3x3_64", 32, 96, 96, 64, 64, 3, 1, 1, False),
    ("l1_conv3_1x1_64_256_res", 32, 96, 96, 64, 256, 1, 1, 1, True),
    ("l3_conv1_1x1_1024_256", 32, 24, 24, 1024, 256, 1, 1, 1, False),
    ("l3_conv2_3x3_256", 32, 24, 24, 256, 256, 3, 1, 1, False),
    ("l3_conv3_1x1_256_1024_res", 32, 24, 24, 256, 1024, 1, 1, 1, True),
    ("l4_conv2_3x3_512_d4", 32, 24, 24, 512, 512, 3, 1, 4, False),
    ("wasp_aspp3_3x3_d12", 32, 24, 24, 256, 256, 3, 1, 12, False),
    ("dec_conv_a_3x3_320_256", 32, 48, 48, 320, 256, 3, 1, 1, False),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--prec", default="fp16")
    ap.add_argument("--only", default="")
    ap.add_argument("--time", action="store_true", help="print CUDA-event timings instead of running once")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    mode = ops.mode_of(args.prec)
    for name, n, h, w, cin, cout, k, stride, dil, res in CONVS:
        if args.only and args.only not in name:
            continue
        x = ops.Act(n, h, w, cin, mode, dev)
        x.t.copy_(torch.randn(x.t.shape, device=dev))
        wt = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
        pc = ops.make_packed_conv(wt, mode, cout=cout, cin=cin)
        pad = dil * (k - 1) // 2
        ho, wo = h // stride, w // stride
        y = ops.Act(n, ho, wo, cout, mode, dev)
        r = None
        if res:
            r = ops.Act(n, ho, wo, cout, mode, dev)
            r.t.copy_(torch.randn(r.t.shape, device=dev))
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        times = []
        for _ in range(args.reps):
            flush.zero_()
            a = torch.cuda.Event(enable_timing=True)
            b = torch.cuda.Event(enable_timing=True)
            a.record()
            ops.conv2d(x, pc, y, stride=stride, dil=dil, pad=pad, relu=True, residual=r, ho=ho, wo=wo)
            b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b) * 1e3)
        flops = 2.0 * n * ho * wo * cout * cin * k * k
        byts = 2.0 * (n * h * w * cin + n * ho * wo * cout * (2 if res else 1))
        best = min(times)
        print("%-28s %8.1f us  %7.1f TFLOP/s  %6.0f GB/s (algorithmic)" % (name, best, flops / best / 1e6, byts / best / 1e3),
              flush=True)


if __name__ == "__main__":
    main()
