"""Base class of the nn.Module mirrors: parameters live in ordinary nn.Conv2d / nn.BatchNorm2d containers
(identical state_dict keys, shapes and init as the reference); `forward` never calls them but emits a
kernel plan (engine.Plan) and runs it.  There is no CPU / eager fallback."""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch
import torch.nn as nn

from . import engine, ops
from .ops import Act


class PlanModule(nn.Module):
    """Sub-classes implement `_emit(builder, *acts) -> Act | tuple[Act] | Tensor` in terms of NHWC buffers.

    Calling the module with fp32 NCHW CUDA tensors (the reference's calling convention) converts the
    inputs to NHWC 16-bit, runs the emitted plan and converts the outputs back to fp32 NCHW.
    """

    precision: Optional[str] = None     # None -> engine.default_precision()
    _out_channels: Sequence[int] = ()   # real channel count of every Act output (for the NCHW conversion)

    def __init__(self):
        super().__init__()
        object.__setattr__(self, "_plans", {})

    # -- helpers -------------------------------------------------------------------------------
    def _precision(self) -> str:
        return self.precision or engine.default_precision()

    def set_precision(self, precision: Optional[str]) -> "PlanModule":
        """fp32 (bf16x3 split, reference-grade numerics) | bf16 | fp16; applies to all sub-modules."""
        if precision is not None:
            ops.mode_of(precision)
        for m in self.modules():
            if isinstance(m, PlanModule):
                m.precision = precision
                m._plans.clear()
        return self

    def _plan_cache(self) -> Dict:
        return self._plans

    # Plans survive .train() / .eval() flips: inference plans are only used when every BatchNorm is in eval mode (their
    # packed weights follow the parameters through engine.WeightTable), training plans are keyed by the set of frozen
    # BatchNorm layers - so an epoch loop that alternates training and validation rebuilds nothing.

    def _check_inputs(self, tensors: Sequence[torch.Tensor]) -> None:
        for t in tensors:
            ops.require_cuda(t, "%s input" % type(self).__name__)
        dev = tensors[0].device
        for p in self.parameters():
            if p.device != dev:
                raise RuntimeError("unipose_b200: parameters on %s but input on %s (call .cuda())" % (p.device, dev))
            break

    def _bn_eval_only(self) -> None:
        bns = self.__dict__.get("_bn_list")
        if bns is None:
            bns = [m for m in self.modules() if isinstance(m, nn.BatchNorm2d)]
            object.__setattr__(self, "_bn_list", bns)
        for m in bns:
            if m.training:
                raise NotImplementedError(
                    "unipose_b200: train-mode BatchNorm (batch statistics) is only supported through the training "
                    "step API (unipose_b200.train); call .eval() or freeze_bn() for inference")

    # -- generic NCHW fp32 entry point ------------------------------------------------------------
    def _input_channels_pad(self, idx: int, c: int) -> int:
        return ops.round_up(c, 64) if c > 16 else 16

    def _forward_nchw(self, *inputs: torch.Tensor):
        self._check_inputs(inputs)
        self._bn_eval_only()
        key = (tuple(tuple(t.shape) for t in inputs), self._precision(), inputs[0].device.index)
        plan = self._plans.get(key)
        if plan is None:
            plan = engine.Plan(inputs[0].device, self._precision())
            b = plan.builder
            acts = []
            for i, t in enumerate(inputs):
                n, c, h, w = t.shape
                st = plan.static_input((n, c, h, w))
                a = b.act(n, h, w, self._input_channels_pad(i, c), zero=True)
                b.add(lambda st=st, a=a: ops.nchw_to_act(st, a), "nchw_to_nhwc")
                acts.append(a)
            outs = self._emit(b, *acts)
            single = not isinstance(outs, (tuple, list))
            outs = [outs] if single else list(outs)
            finals = []
            for o, c_real in zip(outs, list(self._out_channels) + [None] * len(outs)):
                if isinstance(o, torch.Tensor):
                    finals.append(o)
                else:
                    dst = b.tensor((o.n, c_real, o.h, o.w))
                    b.add(lambda o=o, c_real=c_real, dst=dst: ops.act_to_nchw(o, c_real, dst), "nhwc_to_nchw")
                    finals.append(dst)
            plan.finalize(finals)
            plan.single = single
            self._plans[key] = plan
        outs = plan.run(*[t.detach().float() for t in inputs])
        outs = [o.clone() for o in outs]
        return outs[0] if plan.single else tuple(outs)

    def forward(self, *inputs):
        return self._forward_nchw(*inputs)

    def _emit(self, b: engine.Builder, *acts: Act):  # pragma: no cover - abstract
        raise NotImplementedError
