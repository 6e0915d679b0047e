"""BASELINE.json configs[3]: PennAction UniPose-LSTM, 5-frame window, 13 joints, batch 8, 368x368, one GPU.
Times the reference's call pattern (one forward per frame, states fed back: uniposeLSTM.py:124-125)."""
import argparse, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unipose_b200 import synth
from unipose_b200.model import uniposeLSTM

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--frames", type=int, default=5)
ap.add_argument("--size", type=int, default=368)
ap.add_argument("--precision", default="fp16")
ap.add_argument("--clips", type=int, default=10)
ap.add_argument("--warmup", type=int, default=2)
a = ap.parse_args()
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    m = uniposeLSTM.unipose(num_classes=13, precision=a.precision)
synth.trained_like_init_(m, 0)
m = m.cuda().eval()
B, T, S = a.batch, a.frames, a.size
inp = torch.randn(B, T, 3, S, S, device="cuda")
cm = torch.rand(B, T, 1, S, S, device="cuda")


def clip():
    inp.add_(0)          # a new clip: bumps the tensor version so the cached trunk output is recomputed
    heat = hide = cell = None
    for it in range(T):
        heat, cell, hide = m(inp, cm, it, heat, hide, cell)
    return heat


for _ in range(a.warmup):
    clip()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.clips):
    clip()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.clips
print("temporal batching %s | " % os.environ.get("UNIPOSE_B200_TEMPORAL_BATCH", "1"), end="")
print("video clip (%d frames, batch %d, %dx%d, %s): %.2f ms -> %.1f frames/s" % (T, B, S, S, a.precision, ms, B * T / ms * 1e3))
