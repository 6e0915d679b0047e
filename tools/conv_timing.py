"""Phase timeline of single conv launches (UP_DEBUG_TIMING=1): where do the microseconds of a kernel go?"""
import ctypes, os, sys
os.environ["UP_DEBUG_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from unipose_b200 import ops, _lib
NAMES = ["entry", "prologue", "deps", "first_full", "mma_issued", "acc_ready", "epi_math", "last_store", "stores_done", "exit", "kb0_issued", "kb3_issued", "kb2_full", "kb2_commit", "epi_t1", "epi_t2"]
CASES = [("tiny_1x1_64", 32, 24, 24, 64, 64, 1, 1, False, {}),
         ("l1_conv1_1x1_256_64", 32, 96, 96, 256, 64, 1, 1, False, {}),
         ("l1_conv2_3x3_64", 32, 96, 96, 64, 64, 3, 1, False, {}),
         ("l1_conv2_3x3_64 stages4", 32, 96, 96, 64, 64, 3, 1, False, {"UP_DEBUG_STAGES": "4"}),
         ("l1_conv3_64_256_res", 32, 96, 96, 64, 256, 1, 1, True, {}),
         ("l2_conv2_3x3_128", 32, 48, 48, 128, 128, 3, 1, False, {}),
         ("l2_conv3_128_512_res", 32, 48, 48, 128, 512, 1, 1, True, {}),
         ("l3_conv1_1x1_1024_256", 32, 24, 24, 1024, 256, 1, 1, False, {}),
         ("flat l3_conv1 (144x1x128)", 144, 1, 128, 1024, 256, 1, 1, False, {}),
         ("flat l1_conv1 (2304x1x128)", 2304, 1, 128, 256, 64, 1, 1, False, {}),
         ("l3_conv1 bsplit2", 32, 24, 24, 1024, 256, 1, 1, False, {"UP_DEBUG_BSPLIT": "2"}),
         ("l3_conv1 bsplit4", 32, 24, 24, 1024, 256, 1, 1, False, {"UP_DEBUG_BSPLIT": "4"}),
         ("l3_conv1 stages3", 32, 24, 24, 1024, 256, 1, 1, False, {"UP_DEBUG_STAGES": "3"}),
         ("l3_conv1 stages2", 32, 24, 24, 1024, 256, 1, 1, False, {"UP_DEBUG_STAGES": "2"}),
         ("l3_conv1 blockn128", 32, 24, 24, 1024, 256, 1, 1, False, {"UP_DEBUG_BLOCKN": "128"}),
         ("l3_conv1 pair", 32, 24, 24, 1024, 256, 1, 1, False, {"UP_PAIR": "1", "UP_CLUSTER": "2"}),
         ("l3_conv2_3x3_256", 32, 24, 24, 256, 256, 3, 1, False, {}),
         ("l3_conv2 pair", 32, 24, 24, 256, 256, 3, 1, False, {"UP_PAIR": "1", "UP_CLUSTER": "2"}),
         ("l3_conv3_256_1024_res", 32, 24, 24, 256, 1024, 1, 1, True, {}),
         ("l4_conv2_3x3_512_d4", 32, 24, 24, 512, 512, 3, 4, False, {}),
         ("l4_conv2 pair", 32, 24, 24, 512, 512, 3, 4, False, {"UP_PAIR": "1", "UP_CLUSTER": "2"})]
dev = torch.device("cuda:0"); mode = ops.mode_of("fp16")
for name, n, h, w, cin, cout, k, dil, res, env in CASES:
    for kk_, vv_ in env.items():
        os.environ[kk_] = vv_
    x = ops.Act(n, h, w, cin, mode, dev); x.t.normal_()
    pc = ops.make_packed_conv(torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** .5, mode, cout=cout, cin=cin)
    y = ops.Act(n, h, w, cout, mode, dev)
    r = None
    if res:
        r = ops.Act(n, h, w, cout, mode, dev); r.t.normal_()
    for it in range(3):
        torch.cuda.synchronize()
        ops.conv2d(x, pc, y, dil=dil, relu=True, residual=r)
        torch.cuda.synchronize()
    buf = np.zeros(160 * 16, dtype=np.uint64)
    _lib.call("up_debug_conv_timing", buf.ctypes.data_as(ctypes.c_void_p))
    t = buf.reshape(160, 16).astype(np.int64)
    live = t[:, 0] > 0
    t0 = t[live, 0].min()
    rel = (t[live][:, :16] - t0) / 1e3
    for kk_ in env:
        del os.environ[kk_]
    print("%-24s CTAs %3d | " % (name, live.sum()) + "  ".join("%s %.1f/%.1f" % (NAMES[i], np.median(rel[:, i]), rel[:, i].max()) for i in range(16) if np.median(rel[:, i]) > -1e6))
