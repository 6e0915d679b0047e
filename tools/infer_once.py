"""Runs the inference plan of BASELINE.json configs[1] (384x384, batch 32) a few times without a CUDA graph - the
target of ncu launch lists / metric passes (every launch is then a separate kernel node)."""
import os, sys, warnings
os.environ.setdefault("UNIPOSE_B200_GRAPH", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unipose_b200 import synth
from unipose_b200.model.unipose import unipose
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    m = unipose(dataset="MPII", num_classes=16, precision=prec)
synth.trained_like_init_(m, 0)
m = m.cuda().eval()
x = synth.mpii_like_input(32, 384, 384, seed=0).cuda()
for _ in range(reps):
    m.forward_static(x)
torch.cuda.synchronize()
plan = m.plan_for(x)
print("launches per forward:", plan.launches, [n for n, f, s in plan.ops if f is not None][:6], "...")
