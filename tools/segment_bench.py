"""Where does the forward step go?  Times contiguous segments of the compiled plan (stem, layer1..4, WASP, decoder),
each as its own CUDA graph, L2 flushed between replays; also every single launch as a 1-op graph (hot L2 off: flushed).

    python tools/segment_bench.py [--prec fp16] [--batch 32] [--size 384] [--per-op]
"""
from __future__ import annotations

import argparse
import os
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402


def time_range(plan, lo, hi, flush, reps=10):
    ops_backup = plan.ops
    plan.ops = ops_backup[lo:hi]
    torch.cuda.synchronize()
    plan._launch_all()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        plan._launch_all()
    plan.ops = ops_backup
    ts = []
    for _ in range(reps):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--prec", default="fp16")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=384)
    ap.add_argument("--per-op", action="store_true")
    args = ap.parse_args()
    from unipose_b200.model.unipose import unipose
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = unipose(dataset="MPII", num_classes=16, precision=args.prec).cuda().eval()
    x = torch.randn(args.batch, 3, args.size, args.size, device="cuda")
    plan = m.plan_for(x)
    plan.run(x)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    names = [n for n, _f, _s in plan.ops]
    # segment boundaries: stem = everything before the first block; a block starts at "bottleneck.conv1", a fused run
    # "bottleneck.chain[n]" counts n blocks; 3 + 4 + 23 + 3 blocks
    import re
    starts, nblk = [], 0
    bounds = {0: None, 3: None, 7: None, 30: None}
    for i, n in enumerate(names):
        # a block whose conv1 ran inside the previous block's tail kernel starts at its next own launch
        if n == "bottleneck.conv1" or n.startswith("bottleneck.chain") or (i > 0 and names[i - 1] == "bottleneck.tail+conv1"):
            if nblk in bounds and bounds[nblk] is None:
                bounds[nblk] = i
            nblk += int(re.search(r"\[(\d+)\]", n).group(1)) if n.startswith("bottleneck.chain") else 1
    assert nblk == 33, nblk
    last_backbone = max(i for i, n in enumerate(names) if n.startswith("bottleneck."))
    segs = [("stem", 0, bounds[0]), ("layer1", bounds[0], bounds[3]), ("layer2", bounds[3], bounds[7]),
            ("layer3", bounds[7], bounds[30]), ("layer4", bounds[30], last_backbone + 1),
            ("wasp+decoder", last_backbone + 1, len(names))]
    total = time_range(plan, 0, len(names), flush)
    print("whole plan: %d ops, %.1f us" % (len(names), total))
    acc = 0.0
    for name, lo, hi in segs:
        t = time_range(plan, lo, hi, flush)
        acc += t
        print("%-14s ops %3d..%3d  %8.1f us  %5.1f%%  (%.1f us/op)" % (name, lo, hi, t, 100 * t / total, t / (hi - lo)))
    print("sum of segments %.1f us" % acc)
    if args.per_op:
        for i, n in enumerate(names):
            if plan.ops[i][1] is None:
                continue
            if plan.ops[i][2]:
                print("%3d %-28s (side stream)" % (i, n))
                continue
            print("%3d %-28s %7.1f us" % (i, n, time_range(plan, i, i + 1, flush, reps=5)))


if __name__ == "__main__":
    main()
