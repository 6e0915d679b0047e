"""Bring-up ladder for the tcgen05 conv kernel: each case runs in its own subprocess (a trap or a
watchdog in one case must not poison the CUDA context of the next) and is compared with an fp64
torch conv of the same (quantised) operands.

    python tools/gpu_ladder.py            # all cases
    python tools/gpu_ladder.py --case 3   # one case, in-process
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = [
    # name, dict(n,h,w,cin,cout,k,stride,dil,prec, extras)
    dict(name="1x1_c64_ident", n=1, h=16, w=16, cin=64, cout=64, k=1, prec="bf16", ident=True),
    dict(name="1x1_c64", n=1, h=16, w=16, cin=64, cout=64, k=1, prec="bf16"),
    dict(name="1x1_c256_n256", n=2, h=24, w=24, cin=256, cout=256, k=1, prec="bf16"),
    dict(name="1x1_k1024_many_tiles", n=8, h=24, w=24, cin=1024, cout=512, k=1, prec="bf16"),
    dict(name="3x3_d1", n=2, h=16, w=16, cin=64, cout=64, k=3, prec="bf16"),
    dict(name="3x3_d6_wasp", n=4, h=24, w=24, cin=256, cout=256, k=3, dil=6, prec="bf16"),
    dict(name="3x3_d18_wasp_skip", n=4, h=24, w=24, cin=256, cout=256, k=3, dil=18, prec="bf16"),
    dict(name="3x3_s2", n=2, h=48, w=48, cin=128, cout=128, k=3, stride=2, prec="bf16"),
    dict(name="1x1_s2", n=2, h=48, w=48, cin=256, cout=512, k=1, stride=2, prec="bf16"),
    dict(name="stem_like_ck16", n=2, h=32, w=32, cin=16, cout=64, k=4, pad=2, out_hw=(32, 32), prec="bf16"),
    dict(name="res_relu", n=2, h=24, w=24, cin=256, cout=1024, k=1, prec="bf16", residual=True, relu=True),
    dict(name="nchw_head", n=2, h=48, w=48, cin=256, cout=17, k=1, prec="bf16", nchw=True, bias=True),
    dict(name="fp16_3x3", n=2, h=24, w=24, cin=128, cout=128, k=3, dil=2, prec="fp16", relu=True),
    dict(name="split_1x1", n=2, h=24, w=24, cin=256, cout=256, k=1, prec="fp32"),
    dict(name="split_3x3_res", n=2, h=24, w=24, cin=128, cout=128, k=3, dil=2, prec="fp32", residual=True, relu=True),
    dict(name="odd_23x23", n=3, h=23, w=23, cin=64, cout=128, k=3, prec="bf16", relu=True),
    dict(name="groups_concat", n=2, h=24, w=24, cin=512, cout=256, k=1, prec="bf16", groups=2),
    dict(name="k11_video", n=1, h=46, w=46, cin=16, cout=128, k=11, prec="bf16", relu=True, bias=True),
    dict(name="big_layer3_3x3", n=32, h=24, w=24, cin=256, cout=256, k=3, prec="bf16", relu=True),
    # N = 512 tiles (one 512-column accumulator, two UMMAs per k-step): layer4 conv1 shapes at batch 32
    dict(name="wide_n512_k2048", n=32, h=24, w=24, cin=2048, cout=512, k=1, prec="fp16", relu=True),
    dict(name="wide_n512_k1024_bf16", n=32, h=24, w=24, cin=1024, cout=512, k=1, prec="bf16", relu=True),
    dict(name="wide_n1024_two_ntiles", n=16, h=24, w=24, cin=1024, cout=1024, k=1, prec="fp16"),
    dict(name="wide_n512_3x3_d2_layer4", n=32, h=24, w=24, cin=512, cout=512, k=3, dil=2, prec="fp16", relu=True),
    dict(name="wide_n512_3x3_d4_odd_map", n=32, h=23, w=23, cin=512, cout=512, k=3, dil=4, prec="bf16", relu=True),
    dict(name="layer4_conv3_res_bs32", n=32, h=24, w=24, cin=512, cout=2048, k=1, prec="fp16", residual=True, relu=True),
]


def run_case(idx: int) -> dict:
    import torch
    import torch.nn.functional as F

    from unipose_b200 import ops

    c = CASES[idx]
    torch.manual_seed(1234 + idx)
    dev = torch.device("cuda:0")
    n, h, w, cin, cout, k = c["n"], c["h"], c["w"], c["cin"], c["cout"], c["k"]
    stride, dil = c.get("stride", 1), c.get("dil", 1)
    pad = c.get("pad", dil * (k - 1) // 2)
    mode = ops.mode_of(c["prec"])
    groups = c.get("groups", 1)
    cin_pad = ops.round_up(cin, 16)
    cout_pad = ops.round_up(cout, 32 if c.get("nchw") else 64)

    x = torch.randn(n, cin, h, w, device=dev)
    wt = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    if c.get("ident"):
        wt = torch.zeros(cout, cin, 1, 1, device=dev)
        for i in range(min(cin, cout)):
            wt[i, i, 0, 0] = 1.0
    scale = torch.rand(cout, device=dev) + 0.5
    shift = torch.randn(cout, device=dev) * 0.1
    if c.get("ident"):
        scale.fill_(1.0)
        shift.zero_()
    if c.get("bias") or c.get("residual"):
        scale.fill_(1.0)   # with a residual the add happens before scale/shift: BN scale is folded into the weights

    # ---- device operands ----
    if groups > 1:
        cg = cin // groups
        xa = ops.Act(n * groups, h, w, cg, mode, dev)
        for g in range(groups):
            ops.nchw_to_act(x[:, g * cg:(g + 1) * cg].contiguous(), ops.View(xa, n_off=g * n, n=n))
        xin = ops.View(xa, n_off=0, n=n)
    else:
        xa = ops.Act(n, h, w, cin_pad, mode, dev)
        ops.nchw_to_act(x, xa)
        xin = xa
    sc = torch.zeros(cout_pad, device=dev)
    sh = torch.zeros(cout_pad, device=dev)
    sc[:cout] = scale
    sh[:cout] = shift
    pc = ops.make_packed_conv(wt, mode, scale=sc, shift=sh, cout=cout_pad, cin=cin_pad)
    if "out_hw" in c:
        ho, wo = c["out_hw"]
    else:
        ho = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1
        wo = (w + 2 * pad - dil * (k - 1) - 1) // stride + 1
    res = None
    res_f = None
    if c.get("residual"):
        res_f = torch.randn(n, cout, ho, wo, device=dev)
        res = ops.Act(n, ho, wo, cout_pad, mode, dev)
        ops.nchw_to_act(res_f, res)
    if c.get("nchw"):
        y = torch.full((n, cout, ho, wo), float("nan"), device=dev)
    else:
        y = ops.Act(n, ho, wo, cout_pad, mode, dev, zero=True)
    ops.conv2d(xin, pc, y, stride=stride, dil=dil, pad=pad, relu=c.get("relu", False), residual=res, ho=ho, wo=wo,
               x_groups=groups, x_group_nstride=n if groups > 1 else 0)
    torch.cuda.synchronize()

    # ---- reference on the operands the kernel actually saw ----
    if groups > 1:
        xq = torch.cat([ops.View(xa, n_off=g * n, n=n).act.to_float()[g * n:(g + 1) * n] for g in range(groups)], dim=3)
        xq = xq.permute(0, 3, 1, 2).double()
    else:
        xq = xa.to_float()[..., :cin].permute(0, 3, 1, 2).double()
    wq = pc.w.float().sum(0) if mode == ops.UP_SPLIT else pc.w.float()[0]
    wq = wq.view(k, k, cout_pad, cin_pad)[:, :, :cout, :cin].permute(2, 3, 0, 1).double()
    if "out_hw" in c:
        # asymmetric padding: pad top/left = pad, bottom/right implied
        pb = (ho - 1) * stride + dil * (k - 1) + 1 - h - pad
        pr = (wo - 1) * stride + dil * (k - 1) + 1 - w - pad
        xp = F.pad(xq, (pad, max(pr, 0), pad, max(pb, 0)))
        ref = F.conv2d(xp, wq, stride=stride, dilation=dil)
    else:
        ref = F.conv2d(xq, wq, stride=stride, dilation=dil, padding=pad)
    ref = ref * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)
    if res is not None:
        ref = ref + res.to_float()[..., :cout].permute(0, 3, 1, 2).double()
    if c.get("relu"):
        ref = ref.clamp_min(0)
    got = y.double() if c.get("nchw") else y.to_float()[..., :cout].permute(0, 3, 1, 2).double()

    err = (got - ref).abs()
    tol_rel = {"bf16": 2.0 ** -7, "fp16": 2.0 ** -10, "fp32": 1e-4}[c["prec"]]
    if c.get("nchw"):
        tol_rel = 1e-5
    tol = tol_rel * ref.abs() + tol_rel * ref.abs().mean()
    bad = err > tol
    out = dict(name=c["name"], ok=bool(not bad.any() and torch.isfinite(got).all()), max_abs=float(err.max()),
               ref_absmean=float(ref.abs().mean()), n_bad=int(bad.sum()), numel=bad.numel(),
               got_absmean=float(got.abs().mean()), nonfinite=int((~torch.isfinite(got)).sum()))
    if bad.any():
        idxs = bad.nonzero()[:12].tolist()
        out["first_bad"] = [(i, float(got[tuple(i)]), float(ref[tuple(i)])) for i in idxs]
        # structure of the failure: which channels / rows are wrong
        out["bad_per_channel_head"] = bad.sum(dim=(0, 2, 3))[:16].tolist()
        out["bad_per_row_head"] = bad.sum(dim=(0, 1, 3))[:16].tolist()
        out["bad_per_col_head"] = bad.sum(dim=(0, 1, 2))[:16].tolist()
    if not c.get("nchw"):
        padc = y.to_float()[..., cout:]
        if padc.numel():
            out["pad_channels_absmax"] = float(padc.abs().max())
    return out


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", type=int, default=-1)
    ap.add_argument("--out", default="gpurun_out/ladder.jsonl")
    ap.add_argument("--timeout", type=int, default=120)
    args = ap.parse_args()
    if args.case >= 0:
        print("LADDER_RESULT " + json.dumps(run_case(args.case)))
        return 0
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    n_ok = 0
    with open(args.out, "w") as f:
        for i, c in enumerate(CASES):
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", str(i)], capture_output=True,
                                   text=True, timeout=args.timeout)
                line = [l for l in r.stdout.splitlines() if l.startswith("LADDER_RESULT ")]
                if line:
                    res = json.loads(line[-1][len("LADDER_RESULT "):])
                else:
                    res = dict(name=c["name"], ok=False, rc=r.returncode, stdout=r.stdout[-1500:], stderr=r.stderr[-2500:])
            except subprocess.TimeoutExpired as e:
                res = dict(name=c["name"], ok=False, timeout=True, stdout=(e.stdout or b"")[-1000:].decode("utf8", "replace")
                           if isinstance(e.stdout, bytes) else str(e.stdout)[-1000:])
            n_ok += bool(res.get("ok"))
            f.write(json.dumps(res) + "\n")
            f.flush()
            print(("PASS " if res.get("ok") else "FAIL ") + json.dumps(res)[:1800], flush=True)
    print("ladder: %d / %d passed" % (n_ok, len(CASES)))
    return 0 if n_ok == len(CASES) else 1


if __name__ == "__main__":
    sys.exit(main())
