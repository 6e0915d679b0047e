"""Where does the training step go?  One eager step (no CUDA graphs) with a CUDA-event pair around every C-ABI call,
aggregated per entry point (in-order single stream: the sum is the step minus launch gaps and torch-side ops)."""
import collections, os, sys, warnings
os.environ["UNIPOSE_B200_TRAIN_GRAPH"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unipose_b200 import _lib, synth, train
from unipose_b200.model.unipose import unipose

with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    m = unipose(dataset="MPII", num_classes=16, precision="bf16")
synth.trained_like_init_(m, 0)
m = m.cuda().train()
x = synth.mpii_like_input(32, 384, 384).cuda()
t = torch.rand(32, 17, 48, 48, device="cuda")
ts = train.TrainStep(m)
for _ in range(2):
    ts.step(x, t)
torch.cuda.synchronize()
records = []
rec_args = []
rec_snap = []
orig = _lib.call


def timed(name, *args):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    orig(name, *args)
    b.record()
    records.append((name, a, b))
    rec_args.append(args)
    snap = None
    if name in ("up_conv2d_fwd", "up_conv2d_wgrad"):
        d = args[0]._obj        # ctypes.byref(desc): the descriptor is reused by the caller, so copy the geometry now
        snap = (d.n, d.h, d.w, d.ho, d.wo, d.cin, d.cout, d.kh, d.kw, d.stride, d.dil, d.flags)
    rec_snap.append(snap)


_lib.call = timed
train.ops._lib.call = timed
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ts.step(x, t)
e1.record()
torch.cuda.synchronize()
agg = collections.OrderedDict()
for name, a, b in records:
    d = agg.setdefault(name, [0, 0.0])
    d[0] += 1
    d[1] += a.elapsed_time(b)
tot = sum(v[1] for v in agg.values())
print("eager step %.2f ms wall (events), %.2f ms inside %d C-ABI calls" % (e0.elapsed_time(e1), tot, len(records)))
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-28s %5d calls %8.3f ms %5.1f%%" % (name, n, ms, 100 * ms / tot))

# per layer shape for the streaming BatchNorm kernels: effective bandwidth over the ALGORITHMIC bytes (2 B per element
# per tensor touched)
SHAPE = {"up_bn_stats_finalize": lambda a: (a[1], a[2], 1),
         "up_scale_shift_act": lambda a: (a[6], a[7], 2 + (a[2] is not None) + (a[3] is not None)),
         "up_bn_bwd": lambda a: (a[9], a[11], 5 + 2 * (a[12] & 1) + (a[4] is not None))}
if "--shapes" in sys.argv:
    for kname in SHAPE:
        by = collections.OrderedDict()
        for (name, a, b), args in zip(records, rec_args):
            if name != kname:
                continue
            npix, c, tensors = SHAPE[kname](args)
            d = by.setdefault((npix, c, tensors), [0, 0.0])
            d[0] += 1
            d[1] += a.elapsed_time(b)
        print("\n%s" % kname)
        for (npix, c, tensors), (n, ms) in sorted(by.items(), key=lambda kv: -kv[1][1]):
            gb = npix * c * 2 * tensors / 1e9
            print("  npix %8d c %5d tensors %d : %3d calls %7.1f us each %7.3f ms total  %6.0f GB/s" %
                  (npix, c, tensors, n, 1e3 * ms / n, ms, gb / (ms / n * 1e-3)))

# convolutions by geometry: executed FLOPs (2 * n*ho*wo * cout * cin*kh*kw) and TFLOP/s
if "--convs" in sys.argv:
    for kname in ("up_conv2d_fwd", "up_conv2d_wgrad"):
        by = collections.OrderedDict()
        for (name, a, b), args in zip(records, rec_snap):
            if name != kname:
                continue
            d = by.setdefault(args, [0, 0.0])
            d[0] += 1
            d[1] += a.elapsed_time(b)
        print("\n%s   (n h w -> ho wo | cin cout k stride dil | flags)" % kname)
        for key, (n, ms) in sorted(by.items(), key=lambda kv: -kv[1][1]):
            nb, h, w, ho, wo, cin, cout, kh, kw, stride, dil, flags = key
            fl = 2.0 * nb * ho * wo * cout * cin * kh * kw
            print("  %3d %3dx%3d -> %3dx%3d | %4d -> %4d k%d s%d d%2d | f%3d : %3d calls %7.1f us each %7.3f ms  %6.0f TFLOP/s" %
                  (nb, h, w, ho, wo, cin, cout, kh, stride, dil, flags, n, 1e3 * ms / n, ms, fl / (ms / n * 1e-3) / 1e12))
