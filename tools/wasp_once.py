"""Runs the WASP block plan (config 2 geometry) a few times without a CUDA graph - target for ncu captures."""
import os, sys, warnings
os.environ.setdefault("UNIPOSE_B200_GRAPH", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unipose_b200 import engine, synth
from unipose_b200.model.unipose import unipose
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    m = unipose(dataset="MPII", num_classes=16, precision=prec)
synth.trained_like_init_(m, 0)
m = m.cuda().eval()
plan = engine.Plan(torch.device("cuda:0"), prec, use_graph=False)
b = plan.builder
x = b.act(32, 24, 24, 2048)
x.t.copy_(torch.randn(x.t.shape, device="cuda").clamp_min_(0))
m.wasp._emit(b, x)
plan.finalize([])
for _ in range(3):
    plan.run()
torch.cuda.synchronize()
print("wasp plan launches:", plan.launches)
