"""In-kernel phase timeline of the fused WASP chain (UP_DEBUG_TIMING=1 -> globaltimer stamps per CTA)."""
import os, sys, warnings, ctypes
os.environ["UP_DEBUG_TIMING"] = "1"
os.environ.setdefault("UNIPOSE_B200_GRAPH", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from unipose_b200 import _lib, engine, synth
from unipose_b200.model.unipose import unipose

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 24
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    m = unipose(dataset="MPII", num_classes=16, precision="fp16")
synth.trained_like_init_(m, 0)
m = m.cuda().eval()
plan = engine.Plan(torch.device("cuda:0"), "fp16", use_graph=False)
b = plan.builder
x = b.act(n, hw, hw, 2048)
x.t.copy_(torch.randn(x.t.shape, device="cuda").clamp_min_(0) * 0.5)
m.wasp._emit(b, x)
plan.finalize([])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
names = {0: "entry", 1: "deps", 2: "gap_sums", 28: "pool_gemv1", 29: "pool_gemv2", 30: "final_epi", 31: "exit"}
for s in range(4):
    for j, nm in enumerate(["dep_ok", "main_issued", "acc_ready", "epi_done", "stored", "gemm2_issued"]):
        names[4 + 6 * s + j] = "s%d_%s" % (s, nm)
for rep in range(3):
    flush.zero_()
    plan.run()
    torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (160 * 32))()
_lib.call("up_debug_chain_timing", buf)
t = np.array(buf, dtype=np.float64).reshape(160, 32)
act = t[:, 0] > 0
t = t[act]
t0 = t[:, 0].min()
print("CTAs", len(t))
for k in sorted(names):
    col = t[:, k]
    col = col[col > 0]
    if len(col) == 0:
        continue
    r = (col - t0) / 1e3
    print("%-16s n=%3d  min %7.2f  med %7.2f  max %7.2f us" % (names[k], len(r), r.min(), np.median(r), r.max()))
