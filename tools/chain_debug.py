"""Debug aid for the fused WASP chain: repeated launches with poisoned intermediates (S stack, pooling scratch) must
reproduce the same bits; prints where they differ."""
import os, sys, warnings
os.environ.setdefault("UNIPOSE_B200_GRAPH", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from unipose_b200 import engine, ops, synth
from unipose_b200.model.unipose import unipose

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 24
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    m = unipose(dataset="MPII", num_classes=16, precision="fp16")
synth.trained_like_init_(m, 0)
m = m.cuda().eval()
plan = engine.Plan(torch.device("cuda:0"), "fp16", use_graph=False)
b = plan.builder
x = b.act(n, hw, hw, 2048)
torch.manual_seed(0)
x.t.copy_(torch.randn(x.t.shape, device="cuda").clamp_min_(0) * 0.5)
out = m.wasp._emit(b, x)
plan.finalize([])
bufs = [a for a in plan.buffers if isinstance(a, ops.Act) and a is not x and a is not out]
raw = [t for t in plan.buffers if isinstance(t, torch.Tensor) and t.dtype == torch.uint8]
print("ops:", [nm for nm, f, s in plan.ops if f is not None], "S bufs", [tuple(a.t.shape) for a in bufs], "ws", [t.numel() for t in raw])
ref = None
for it in range(8):
    if it >= 2:
        for a in bufs:
            a.t.fill_(float("nan"))
        out.t.fill_(float("nan"))
        for t in raw:     # poison the pooling scratch behind the counters (first 4 KB hold the counters)
            t[8192:].fill_(0xFF)
    plan.run()
    torch.cuda.synchronize()
    o = out.t.float().cpu().numpy()
    if ref is None:
        ref = o
        print("run 0: finite", np.isfinite(o).all(), "absmax", np.abs(o).max())
        continue
    d = np.abs(o - ref)
    bad = np.argwhere(~(d == 0))
    print("run %d: finite %s, differing %d, max diff %.3g" % (it, np.isfinite(o).all(), len(bad), np.nanmax(d)))
    if len(bad):
        imgs = np.bincount(bad[:, 1], minlength=n)
        print("   per image:", imgs.tolist(), " first:", bad[:5].tolist())
