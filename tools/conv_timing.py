"""Phase timeline of single conv launches (UP_DEBUG_TIMING=1): where do the microseconds of a kernel go?"""
import ctypes, os, sys
os.environ["UP_DEBUG_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from unipose_b200 import ops, _lib
NAMES = ["entry", "prologue", "deps", "first_full", "mma_issued", "acc_ready", "epi_math", "last_store", "stores_done", "exit", "kb0_issued", "kb3_issued", "kb2_full", "kb2_commit", "epi_t1", "epi_t2"]
CASES = [
    ("l1_conv2_3x3_64", 32, 96, 96, 64, 64, 3, 1, False, {"UP_PAIR": "0"}),
    ("l1_conv2 pair", 32, 96, 96, 64, 64, 3, 1, False, {"UP_PAIR": "1"}),
    ("l1_conv3_64_256_res", 32, 96, 96, 64, 256, 1, 1, True, {"UP_PAIR": "0"}),
    ("l1_conv3 pair", 32, 96, 96, 64, 256, 1, 1, True, {"UP_PAIR": "1"}),
    ("l2_conv1_512_128", 32, 48, 48, 512, 128, 1, 1, False, {"UP_PAIR": "0"}),
    ("l2_conv1 pair", 32, 48, 48, 512, 128, 1, 1, False, {"UP_PAIR": "1"}),
    ("l2_conv2_3x3_128", 32, 48, 48, 128, 128, 3, 1, False, {"UP_PAIR": "0"}),
    ("l2_conv2 pair", 32, 48, 48, 128, 128, 3, 1, False, {"UP_PAIR": "1"}),
    ("l2_conv3_128_512_res", 32, 48, 48, 128, 512, 1, 1, True, {"UP_PAIR": "0"}),
    ("l2_conv3 pair", 32, 48, 48, 128, 512, 1, 1, True, {"UP_PAIR": "1"}),
    ("l3_conv1_1x1_1024_256", 32, 24, 24, 1024, 256, 1, 1, False, {"UP_PAIR": "0"}),
    ("l3_conv1 pair", 32, 24, 24, 1024, 256, 1, 1, False, {"UP_PAIR": "1"}),
    ("l3_conv2_3x3_256", 32, 24, 24, 256, 256, 3, 1, False, {"UP_PAIR": "0"}),
    ("l3_conv2 pair", 32, 24, 24, 256, 256, 3, 1, False, {"UP_PAIR": "1"}),
    ("l3_conv3_256_1024_res", 32, 24, 24, 256, 1024, 1, 1, True, {"UP_PAIR": "0"}),
    ("l3_conv3 pair", 32, 24, 24, 256, 1024, 1, 1, True, {"UP_PAIR": "1"}),
    ("l4_conv1_1024_512", 32, 24, 24, 1024, 512, 1, 1, False, {"UP_PAIR": "0"}),
    ("l4_conv1 pair", 32, 24, 24, 1024, 512, 1, 1, False, {"UP_PAIR": "1"}),
    ("l4_conv2_3x3_512_d4", 32, 24, 24, 512, 512, 3, 4, False, {"UP_PAIR": "0"}),
    ("l4_conv2 pair", 32, 24, 24, 512, 512, 3, 4, False, {"UP_PAIR": "1"}),
    ("l4_conv3_512_2048_res", 32, 24, 24, 512, 2048, 1, 1, True, {"UP_PAIR": "0"}),
    ("l4_conv3 pair", 32, 24, 24, 512, 2048, 1, 1, True, {"UP_PAIR": "1"}),
    ("aspp1_2048_256", 32, 24, 24, 2048, 256, 1, 1, False, {"UP_PAIR": "0"}),
    ("aspp1 pair", 32, 24, 24, 2048, 256, 1, 1, False, {"UP_PAIR": "1"}),
    ("aspp2_3x3_256_d6", 32, 24, 24, 256, 256, 3, 6, False, {"UP_PAIR": "0"}),
    ("aspp2 pair", 32, 24, 24, 256, 256, 3, 6, False, {"UP_PAIR": "1"}),
    ("dec_a_3x3_320_256", 32, 48, 48, 320, 256, 3, 1, False, {"UP_PAIR": "0"}),
    ("dec_a pair", 32, 48, 48, 320, 256, 3, 1, False, {"UP_PAIR": "1"}),
]
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
