"""Time the fused training step (config 3 per-GPU shape: MPII 384x384, batch 32, bf16) on one GPU."""
import argparse, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unipose_b200 import synth, train
from unipose_b200.model.unipose import unipose

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--size", type=int, default=384)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=4)
a = ap.parse_args()
# under torchrun (BASELINE.json configs[2]: batch 32 per GPU, one NCCL all-reduce of the flat gradient per step)
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl")
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    m = unipose(dataset="MPII", num_classes=16, precision=a.precision)
synth.trained_like_init_(m, 0)
m = m.cuda().train()
torch.manual_seed(100 + rank)     # every rank trains on its own shard of the global batch
x = synth.mpii_like_input(a.batch, a.size, a.size).cuda() + 0.01 * torch.randn(a.batch, 3, a.size, a.size, device="cuda")
t = torch.rand(a.batch, 17, a.size // 8, a.size // 8, device="cuda")
ts = train.TrainStep(m)
for _ in range(a.warmup):     # eager warm-up, then the graph capture of forward + loss + backward
    loss = ts.step(x, t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    loss = ts.step(x, t)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
if world > 1:
    from unipose_b200 import parallel
    ms = parallel.max_over_ranks([ms], "cuda")[0]
    # data-parallel invariant: identical parameters on every rank after the all-reduced Adam steps
    chk = ts.flat_p.double().sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert float(hi - lo) == 0.0, "parameters diverged across ranks: %r vs %r" % (float(lo), float(hi))
    if rank != 0:
        dist.destroy_process_group()
        sys.exit(0)
print("train step (%d GPU%s, batch %d per GPU): %.2f ms  -> %.1f frames/s  (loss %.5f, fwd ops %d, bwd ops %d, mem %.1f GB)" % (
    world, "s" if world > 1 else "", a.batch, ms, world * a.batch / ms * 1e3, float(loss), len(ts.plan.fwd), len(ts.plan.bwd), torch.cuda.max_memory_allocated() / 2**30))
