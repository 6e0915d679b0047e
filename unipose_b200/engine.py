"""Execution plans: a module's forward is compiled ONCE per (input shapes, precision, train/eval) into a
flat list of C-ABI kernel launches over pre-allocated NHWC buffers with pre-packed weights, optionally
captured in a CUDA graph.  torch provides device memory and streams only.

    plan = Plan(device, precision)
    b = plan.builder
    x = b.input_image(n, 3, h, w)          # static fp32 NCHW staging tensor
    ...module._emit(b, ...)...              # records ops + weight-pack jobs
    plan.finalize(outputs)
    outs = plan.run(input)                  # refreshes packed weights if parameters changed, launches
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops
from .ops import Act, PackedConv, View, as_view, round_up

_BN_EPS_DEFAULT = 1e-5


def default_precision() -> str:
    """fp32 = bf16x3 split ("parity") mode, the reference's numerics; bf16 / fp16 are the throughput modes."""
    return os.environ.get("UNIPOSE_B200_PRECISION", "fp32")


def _versions(tensors: Sequence[torch.Tensor]) -> Tuple[int, ...]:
    return tuple((t.data_ptr(), t._version) for t in tensors)


class _PackJob:
    """Keeps one packed weight (+ scale/shift) in sync with its source parameters."""

    def __init__(self, sources: Sequence[torch.Tensor], fn: Callable[[], None]):
        self.sources = list(sources)
        self.fn = fn
        self.seen = None

    def refresh(self) -> bool:
        v = _versions(self.sources)
        if v != self.seen:
            self.fn()
            self.seen = v
            return True
        return False


class Builder:
    def __init__(self, plan: "Plan"):
        self.plan = plan
        self.device = plan.device
        self.mode = plan.mode

    # ---- buffers ----
    def act(self, n, h, w, c, zero=False) -> Act:
        a = Act(n, h, w, c, self.mode, self.device, zero=zero)
        self.plan.buffers.append(a)
        return a

    def tensor(self, shape, dtype=torch.float32, zero=False) -> torch.Tensor:
        t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.device)
        self.plan.buffers.append(t)
        return t

    def add(self, fn: Callable[[], None], name: str = "", side: bool = False) -> None:
        """Record a launch.  side=True puts it on the plan's side stream (between fork() and join()), so an
        independent branch (e.g. WASP's image-level pooling branch) overlaps the main chain inside the CUDA graph."""
        self.plan.ops.append((name, fn, side))

    def fork(self) -> None:
        self.plan.ops.append(("fork", None, False))

    def join(self) -> None:
        self.plan.ops.append(("join", None, False))

    # ---- conv (+ folded eval BatchNorm / bias) ----
    def packed_conv(self, conv: nn.Conv2d, bn: Optional[nn.BatchNorm2d], cin_pad: Optional[int] = None,
                    cout_pad: Optional[int] = None, weight_fn: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                    nchw_out: bool = False, extra_sources: Sequence[torch.Tensor] = ()) -> PackedConv:
        """Allocate packed buffers for `conv` (+`bn` folded as eval-mode scale/shift) and register the job that
        (re)fills them from the live parameters."""
        w0 = conv.weight if weight_fn is None else weight_fn(conv.weight.detach())
        co_r, ci_r, kh, kw = w0.shape
        cout = cout_pad or round_up(co_r, 32 if nchw_out else 64)
        cin = cin_pad or round_up(ci_r, 16)
        planes = 2 if self.mode == ops.UP_SPLIT else 1
        dt = torch.float16 if self.mode == ops.UP_FP16 else torch.bfloat16
        wbuf = torch.empty((planes, kh * kw, cout, cin), dtype=dt, device=self.device)
        scale = torch.zeros(cout, dtype=torch.float32, device=self.device)
        shift = torch.zeros(cout, dtype=torch.float32, device=self.device)
        pc = PackedConv(wbuf, scale, shift, kh, kw, cout, cin, co_r, ci_r, self.mode)

        c_bn = bn.num_features if bn is not None else 0
        fold_scale = torch.empty(c_bn, dtype=torch.float32, device=self.device) if bn is not None else None
        fold_shift = torch.empty(c_bn, dtype=torch.float32, device=self.device) if bn is not None else None

        def fill():
            w = conv.weight.detach().float()
            scale.zero_()
            scale[:co_r] = 1.0
            shift.zero_()
            if bn is not None:
                # eval-mode BatchNorm: the per-channel scale is folded into the canonical filter
                # (w' = w * gamma/sqrt(var+eps)), the shift stays in the epilogue.  The residual of a bottleneck is
                # added inside the tensor-core pipeline BEFORE the epilogue, so the epilogue itself must not scale.
                ops._lib.call("up_bn_fold", ops._ptr(bn.weight.detach().float().contiguous()),
                              ops._ptr(bn.bias.detach().float().contiguous()),
                              ops._ptr(bn.running_mean.float().contiguous()),
                              ops._ptr(bn.running_var.float().contiguous()), float(bn.eps), ops._ptr(fold_scale),
                              ops._ptr(fold_shift), c_bn, c_bn, ops._stream())
                w = w * fold_scale.view(-1, 1, 1, 1)
                # a weight_fn may replicate the output channels (the stem's 4 pixels per super pixel): tile the shift
                shift[:co_r] = fold_shift.repeat(co_r // c_bn)
            elif conv.bias is not None:
                shift[:co_r] = conv.bias.detach().float().repeat(co_r // conv.bias.numel())
            if weight_fn is not None:
                w = weight_fn(w)
            w = w.contiguous()
            ops._lib.call("up_pack_conv_weight", ops._ptr(w), ops._ptr(wbuf), co_r, ci_r, kh, kw, cout, cin, self.mode,
                          kh * kw * cout * cin, ops._stream())

        srcs = [conv.weight] + ([conv.bias] if conv.bias is not None else []) + list(extra_sources)
        if bn is not None:
            srcs += [bn.weight, bn.bias, bn.running_mean, bn.running_var]
        self.plan.pack_jobs.append(_PackJob(srcs, fill))
        return pc

    def conv(self, x, pc: PackedConv, y, name: str = "conv", side: bool = False, **kw) -> None:
        self.add(lambda: ops.conv2d(x, pc, y, **kw), name, side=side)


class Plan:
    def __init__(self, device, precision: str, use_graph: Optional[bool] = None):
        self.device = torch.device(device)
        self.precision = precision
        self.mode = ops.mode_of(precision)
        self.buffers: List = []
        self.ops: List[Tuple[str, Optional[Callable[[], None]], bool]] = []
        self.side_stream: Optional[torch.cuda.Stream] = None
        self.pack_jobs: List[_PackJob] = []
        self.inputs: List[torch.Tensor] = []
        self.outputs: List[torch.Tensor] = []
        self.builder = Builder(self)
        if use_graph is None:
            use_graph = os.environ.get("UNIPOSE_B200_GRAPH", "1") != "0"
        self.use_graph = use_graph
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches = 0

    def static_input(self, shape) -> torch.Tensor:
        t = torch.zeros(shape, dtype=torch.float32, device=self.device)
        self.inputs.append(t)
        return t

    def finalize(self, outputs: Sequence[torch.Tensor]) -> None:
        self.outputs = list(outputs)

    def refresh_weights(self) -> bool:
        changed = False
        for j in self.pack_jobs:
            changed |= j.refresh()
        return changed

    def _launch_all(self) -> None:
        main = torch.cuda.current_stream(self.device)
        for name, fn, side in self.ops:
            if fn is None:
                if self.side_stream is None:
                    self.side_stream = torch.cuda.Stream(device=self.device)
                if name == "fork":
                    self.side_stream.wait_stream(main)
                else:
                    main.wait_stream(self.side_stream)
            elif side:
                with torch.cuda.stream(self.side_stream):
                    fn()
            else:
                fn()

    def run(self, *inputs: torch.Tensor) -> List[torch.Tensor]:
        assert len(inputs) == len(self.inputs)
        self.refresh_weights()  # repacks in place (same addresses), so a captured graph stays valid
        for dst, src in zip(self.inputs, inputs):
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        if self.use_graph:
            if self.graph is None:
                # warm-up launch outside capture (lazy CUDA init, function attributes), then capture
                self._launch_all()
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch_all()
                self.graph = g
            self.graph.replay()
        else:
            self._launch_all()
        self.launches = sum(1 for _n, fn, _s in self.ops if fn is not None)
        return self.outputs
