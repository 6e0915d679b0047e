"""Training path (forward in .train() mode + backward) of the image model on the B200 kernels.

`model(input)` in train mode returns heat-maps that carry a grad_fn: the reference's own training loop
(unipose.py:113-124: optimizer.zero_grad(); heat = model(x); loss = MSELoss(heat, target); loss.backward();
optimizer.step()) runs unchanged.  One torch.autograd.Function spans the whole network; its forward / backward
replay two static kernel plans:

  forward   per conv+BN unit: tcgen05 conv (raw output z) -> up_bn_stats_finalize (batch statistics, running
            stats update) -> up_scale_shift_act (normalise + residual + ReLU + dropout mask)
  backward  reverse tape: up_bn_bwd (ReLU gate + BatchNorm backward, dgamma/dbeta, residual gradient) ->
            tcgen05 wgrad (MN-major operands) -> dgrad = the forward conv kernel on the flipped/transposed filter
            (gradient accumulation across branches rides on its residual input)

`TrainStep` adds the fused tail used by the bench / multi-GPU training: up_mse_fwd_bwd, one NCCL all-reduce of the
flat gradient buffer, up_adam_step.
"""
from __future__ import annotations

import os

import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import engine, ops
from .ops import Act, PackedConv, View, as_view, round_up


def _region(v: View):
    return (id(v.act), v.n_off, v.n_off + v.n, v.coff, v.coff + v.c)


class TrainPlan:
    """Forward + backward op lists over static buffers for one (input shape, precision)."""

    def __init__(self, device, precision: str):
        self.device = torch.device(device)
        self.precision = precision
        self.mode = ops.mode_of(precision)
        self.fwd: List = []
        self.bwd: List = []
        # indices of backward ops that may run on the side stream: the weight gradients.  wgrad(L) only needs dz(L) and
        # the saved input x(L); nothing on the critical path bn_bwd(L) -> dgrad(L) -> bn_bwd(L-1) -> ... reads its
        # result, so it overlaps the (HBM-bound) BatchNorm backward and the heads/tails of the dgrad kernels.
        self.bwd_side = set()
        self.side_stream: Optional[torch.cuda.Stream] = None
        self.overlap_wgrad = os.environ.get("UNIPOSE_B200_WGRAD_OVERLAP", "1") != "0"
        self.tape: List = []
        self.buffers: List = []
        self.weights = engine.WeightTable(self.device, self.mode, always=True)   # refilled every step
        self.fwd_counter = 0
        self.grad_ready: Dict[int, int] = {}   # id(param) -> index into self.bwd after which its gradient is final
        self.grads: Dict[int, Act] = {}    # id(forward Act) -> gradient Act
        self.written: List[Tuple] = []     # gradient regions that already hold a value
        self.pgrad: Dict[int, torch.Tensor] = {}
        self.pgrad_written = set()
        self.params: List[nn.Parameter] = []
        self.masks: List[Tuple[Act, float]] = []
        self.bn_modules: List[nn.BatchNorm2d] = []
        self.scratch_bytes = 0
        self.scratch_main_bytes = 0
        self.live_params: List[nn.Parameter] = []
        self._live_ids = set()
        self.flat_g: Optional[torch.Tensor] = None
        self.scratch: Optional[torch.Tensor] = None
        self.scratch_main: Optional[torch.Tensor] = None
        self.input: Optional[torch.Tensor] = None
        self.heat: Optional[torch.Tensor] = None
        self.dheat: Optional[torch.Tensor] = None

    # ---- buffers -------------------------------------------------------------------------------
    def act(self, n, h, w, c, zero=False) -> Act:
        a = Act(n, h, w, c, self.mode, self.device, zero=zero)
        self.buffers.append(a)
        return a

    def tensor(self, shape, dtype=torch.float32, zero=True) -> torch.Tensor:
        t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.device)
        self.buffers.append(t)
        return t

    def grad_act(self, a: Act) -> Act:
        g = self.grads.get(id(a))
        if g is None:
            g = self.act(a.n, a.h, a.w, a.c)
            self.grads[id(a)] = g
        return g

    def grad_view(self, v) -> View:
        v = as_view(v)
        return View(self.grad_act(v.act), coff=v.coff, c=v.c, n_off=v.n_off, n=v.n)

    def is_written(self, gv: View) -> bool:
        a, n0, n1, c0, c1 = _region(gv)
        for (b, m0, m1, d0, d1) in self.written:
            if a == b and n0 < m1 and m0 < n1 and c0 < d1 and d0 < c1:
                return True
        return False

    def mark_written(self, gv: View) -> None:
        self.written.append(_region(gv))

    def has_grad(self, v) -> bool:
        v = as_view(v)
        g = self.grads.get(id(v.act))
        if g is None:
            return False
        return self.is_written(View(g, coff=v.coff, c=v.c, n_off=v.n_off, n=v.n))

    def _grad_final(self, p: nn.Parameter) -> None:
        """The op just appended to self.bwd is (so far) the last writer of p's gradient."""
        self.grad_ready[id(p)] = len(self.bwd)

    def param_grad(self, p: nn.Parameter) -> torch.Tensor:
        g = self.pgrad.get(id(p))
        if g is None:
            g = torch.zeros_like(p, dtype=torch.float32, memory_format=torch.contiguous_format)
            self.pgrad[id(p)] = g
            self.params.append(p)
        return g

    # ---- weights ---------------------------------------------------------------------------------
    def packed(self, conv: nn.Conv2d, cout: int, cin: int, *, weight_fn=None, transpose: bool = False, ci_off: int = 0,
               cin_slice: Optional[int] = None, use_bias: bool = False) -> PackedConv:
        """Packed filter of `conv`, refilled every step by the plan's weight table (ONE launch for all filters).
        transpose=True is the dgrad layout of the input-channel slice [ci_off, ci_off+cin_slice): rows = those input
        channels (padded to `cout`), cols = the conv's output channels (padded to `cin`), taps flipped."""
        w0 = conv.weight.detach() if weight_fn is None else weight_fn(conv.weight.detach())
        co_r, ci_r, kh, kw = w0.shape
        planes = 2 if self.mode == ops.UP_SPLIT else 1
        dt = torch.float16 if self.mode == ops.UP_FP16 else torch.bfloat16
        wbuf = torch.empty((planes, kh * kw, cout, cin), dtype=dt, device=self.device)
        scale = torch.empty(cout, dtype=torch.float32, device=self.device)
        shift = torch.empty(cout, dtype=torch.float32, device=self.device)
        rows_real = (cin_slice if cin_slice is not None else ci_r - ci_off) if transpose else co_r
        cols_real = co_r if transpose else ci_r
        pc = PackedConv(wbuf, scale, shift, kh, kw, cout, cin, rows_real, cols_real, self.mode)
        bias = (lambda: conv.bias) if (use_bias and conv.bias is not None) else None
        self.weights.add_epilogue(scale, shift, rows_real, bias=bias)
        self.weights.add_pack(lambda: conv.weight, wbuf, cout, cin, transpose=transpose, ci_off=ci_off,
                              cin_slice=cin_slice, weight_fn=weight_fn)
        return pc

    # ---- ops with adjoints --------------------------------------------------------------------------
    def conv_unit(self, x, conv: nn.Conv2d, bn: Optional[nn.BatchNorm2d], *, relu: bool, residual=None, stride=1,
                  dil=1, pad=None, mask=None, x_groups=1, x_group_nstride=0, x_window=None, weight_fn=None,
                  weight_fn_inv=None, cin_pad=None, cout_pad=None, x_needs_grad=True, out=None, nchw_out=None,
                  nchw_grad=None):
        """conv (+ train-mode BatchNorm) (+ residual) (+ ReLU) (+ dropout mask); records its backward.
        bn=None: conv (+ bias) (+ ReLU) - the video model's middle CNN and pooling branch.  nchw_out: the conv writes
        fp32 NCHW straight into (the first channels of) that tensor; its gradient is read from `nchw_grad` (default: the
        plan's dheat), a tensor of the same shape."""
        xv = as_view(x)
        w_src = (lambda: conv.weight.detach()) if weight_fn is None else (lambda: weight_fn(conv.weight.detach()))
        w0 = w_src()
        co_r, ci_r, kh, kw = w0.shape
        cout = cout_pad or round_up(co_r, 64)
        cin = cin_pad or round_up(ci_r, 16)
        if pad is None:
            pad = (dil * (kh - 1) // 2, dil * (kw - 1) // 2)
        if isinstance(pad, int):
            pad = (pad, pad)
        h_in, w_in = xv.h, (x_window[0] if x_window is not None else xv.w)
        ho = (h_in + 2 * pad[0] - dil * (kh - 1) - 1) // stride + 1
        wo = (w_in + 2 * pad[1] - dil * (kw - 1) - 1) // stride + 1
        if x_window is not None:
            ho, wo = h_in, w_in          # the stem: explicit output size, asymmetric padding
        n = xv.n
        pc = self.packed(conv, cout, cin, weight_fn=weight_fn, use_bias=bn is None)
        kw_conv = dict(stride=stride, dil=dil, pad=pad, ho=ho, wo=wo, x_groups=x_groups,
                       x_group_nstride=x_group_nstride, x_window=x_window)
        z = out if (bn is None and out is not None) else self.act(n, ho, wo, cout)
        if bn is None and out is not None:
            z = as_view(out)
            assert (z.n, z.h, z.w, z.c) == (n, ho, wo, cout)
        for prm in ([conv.weight] + ([conv.bias] if (conv.bias is not None and bn is None) else []) +
                    ([bn.weight, bn.bias] if bn is not None else [])):
            if id(prm) not in self._live_ids:
                self._live_ids.add(id(prm))
                self.live_params.append(prm)
        rec = dict(x=xv, conv=conv, bn=bn, relu=relu, residual=residual, mask=mask, pc=pc, z=z, stride=stride, dil=dil,
                   pad=pad, ho=ho, wo=wo, x_groups=x_groups, x_group_nstride=x_group_nstride, x_window=x_window,
                   weight_fn=weight_fn, weight_fn_inv=weight_fn_inv, x_needs_grad=x_needs_grad, cout=cout, cin=cin,
                   co_r=co_r, ci_r=ci_r, kh=kh, kw=kw, nchw_out=nchw_out, nchw_grad=nchw_grad)
        if nchw_out is not None:
            # network head: conv + bias (+ ReLU) straight into the fp32 NCHW heat-map tensor
            assert bn is None and mask is None and residual is None
            self.fwd.append(lambda: ops.conv2d(xv, pc, nchw_out, cout_valid=co_r, out_c_total=nchw_out.shape[1],
                                               relu=relu, **kw_conv))
            rec["y"] = None
            self.tape.append(lambda: self._conv_unit_bwd(rec))
            return nchw_out
        if bn is None:
            if mask is not None or residual is not None:
                raise NotImplementedError("conv without BatchNorm supports bias (+ ReLU) epilogues in training")
            self.fwd.append(lambda: ops.conv2d(xv, pc, z, relu=relu, **kw_conv))
            y = z
        else:
            self.fwd.append(lambda: ops.conv2d(xv, pc, z, **kw_conv))
        if bn is None:
            pass
        else:
            c = cout
            sums = self.tensor((ops.bn_work_doubles(c),), dtype=torch.float64)   # sums + scratch + partial rows
            scale, shift, mean, invstd = (self.tensor((c,)) for _ in range(4))
            y = out if out is not None else self.act(n, ho, wo, cout)
            count = n * ho * wo
            frozen = not bn.training      # freeze_bn(): running statistics, no update (model/unipose.py:40-43)
            if frozen:
                self.fwd.append(lambda: ops.bn_eval_prepare(bn, scale, shift, mean, invstd, co_r, c))
            else:
                self.bn_modules.append(bn)
                zv = ops.as_view(z)
                assert count == zv.n * zv.h * zv.w
                self.fwd.append(lambda: ops.bn_stats_finalize(z, sums, bn, scale, shift, mean, invstd, co_r, c))
            self.fwd.append(lambda: ops.scale_shift_act(z, y, scale, shift, relu=relu, residual=residual, mask=mask))
            rec.update(sums=sums, mean=mean, invstd=invstd, frozen=frozen)
        rec["y"] = y
        self.tape.append(lambda: self._conv_unit_bwd(rec))
        return y

    def _conv_unit_bwd(self, r) -> None:
        xv, conv, bn, pc, z = r["x"], r["conv"], r["bn"], r["pc"], r["z"]
        n, ho, wo, cout = z.n, z.h, z.w, z.c
        # ---- gradient w.r.t. the conv output z ----
        if r["nchw_out"] is not None:
            dz = self.act(n, ho, wo, cout, zero=True)
            dheat = r["nchw_grad"] if r["nchw_grad"] is not None else self.dheat
            if r["relu"]:
                gated = self.tensor(tuple(dheat.shape), zero=False)
                yout = r["nchw_out"]
                self.bwd.append(lambda src=dheat, yout=yout, gated=gated: torch.mul(src, (yout > 0).to(src.dtype), out=gated))
                dheat = gated
            self.bwd.append(lambda src=dheat, dz=dz: ops.nchw_to_act(src, dz))
            if conv.bias is not None:
                bsums = self.tensor((ops.bn_work_doubles(cout),), dtype=torch.float64)
                gb = self.param_grad(conv.bias)
                self.bwd.append(lambda: ops.bn_stats(dz, cout, bsums))
                self.bwd.append(lambda: gb.copy_(bsums[:r["co_r"]]))
                self._grad_final(conv.bias)
        else:
            y = r["y"]
            if not self.has_grad(y):
                return
            dy = self.grad_view(y)
            if r["mask"] is not None:
                tmp = self.act(n, ho, wo, cout)
                mask = r["mask"]
                self.bwd.append(lambda dy=dy, tmp=tmp, mask=mask: ops.ew(dy, tmp, m=mask, op=1))
                dy = as_view(tmp)
            if bn is None:
                dz = dy          # plain conv: dz is the incoming gradient itself ...
                if r["relu"]:    # ... gated by the ReLU of the fused epilogue
                    tmp = self.act(n, ho, wo, cout)
                    self.bwd.append(lambda dy=dy, tmp=tmp, y=y: ops.ew(dy, tmp, m=y, op=2))
                    dz = as_view(tmp)
                if conv.bias is not None:
                    bsums = self.tensor((ops.bn_work_doubles(cout),), dtype=torch.float64)
                    gb = self.param_grad(conv.bias)
                    dzb = dz
                    self.bwd.append(lambda: ops.bn_stats(dzb, cout, bsums))
                    self.bwd.append(lambda: gb.copy_(bsums[:r["co_r"]]))
                    self._grad_final(conv.bias)
            else:
                dz = self.act(n, ho, wo, cout)
                dres, dres_tmp, rv = None, None, None
                if r["residual"] is not None:
                    rv = self.grad_view(r["residual"])
                    if self.is_written(rv):
                        dres_tmp = self.act(n, ho, wo, cout)
                        dres = dres_tmp
                    else:
                        dres = rv
                        self.mark_written(rv)
                dgamma, dbeta = self.param_grad(bn.weight), self.param_grad(bn.bias)
                sums, mean, invstd, relu, co_r = r["sums"], r["mean"], r["invstd"], r["relu"], r["co_r"]
                frozen = r["frozen"]
                self.bwd.append(lambda: ops.bn_bwd(dy, y, z, dz, dres, mean, invstd, bn.weight.detach(), sums, co_r,
                                                   relu, dgamma, dbeta, frozen=frozen))
                self._grad_final(bn.weight)
                self._grad_final(bn.bias)
                if dres_tmp is not None:
                    self.bwd.append(lambda: ops.ew(dres_tmp, rv, accumulate=True))
        dzv = as_view(dz)
        # ---- weight gradient ----
        d = ops.conv_desc(xv, pc, ho, wo, stride=r["stride"], dil=r["dil"], pad=r["pad"], x_groups=r["x_groups"],
                          x_group_nstride=r["x_group_nstride"], x_window=r["x_window"])
        if r["weight_fn"] is None:
            self.scratch_bytes = max(self.scratch_bytes, ops.wgrad_scratch_bytes(d))
        else:       # stays on the main stream (torch ops follow it): its own split scratch, the side stream owns the other
            self.scratch_main_bytes = max(self.scratch_main_bytes, ops.wgrad_scratch_bytes(d))
        gw = self.param_grad(conv.weight)
        acc = id(conv.weight) in self.pgrad_written
        self.pgrad_written.add(id(conv.weight))
        if r["weight_fn"] is None:
            self.bwd_side.add(len(self.bwd))
            self.bwd.append(lambda: ops.conv2d_wgrad(d, xv, dzv, gw, self.scratch, accumulate=acc))
        else:
            gtmp = self.tensor((r["co_r"], r["ci_r"], r["kh"], r["kw"]))
            inv = r["weight_fn_inv"]
            self.bwd.append(lambda: ops.conv2d_wgrad(d, xv, dzv, gtmp, self.scratch_main, accumulate=False))
            if acc:
                self.bwd.append(lambda: gw.add_(inv(gtmp)))
            else:
                self.bwd.append(lambda: gw.copy_(inv(gtmp)))
        self._grad_final(conv.weight)
        # ---- input gradient: the forward kernel on the flipped / transposed filter ----
        if not r["x_needs_grad"]:
            return
        kh, kw, dil, stride = r["kh"], r["kw"], r["dil"], r["stride"]
        pad_t = (dil * (kh - 1) - r["pad"][0], dil * (kw - 1) - r["pad"][1])
        src = dzv
        if stride == 2:
            up = self.act(n, 2 * ho, 2 * wo, cout)
            self.bwd.append(lambda: ops.zero_insert2x(dzv, up))
            src = as_view(up)
        groups = r["x_groups"]
        cg_real = r["ci_r"] // groups
        cg_pad = r["cin"] // groups
        wfn = r["weight_fn"]
        for g in range(groups):
            pct = self.packed(conv, cout=cg_pad if groups > 1 else xv.c, cin=cout, weight_fn=wfn, transpose=True,
                              ci_off=g * cg_real, cin_slice=cg_real)
            xg = View(self.grad_act(xv.act), coff=xv.coff, c=xv.c, n_off=xv.n_off + g * r["x_group_nstride"], n=xv.n)
            res = xg if self.is_written(xg) else None
            self.mark_written(xg)
            self.bwd.append(lambda src=src, pct=pct, xg=xg, res=res: ops.conv2d(
                src, pct, xg, dil=dil, pad=pad_t, ho=xv.h, wo=xv.w, residual=res))

    def maxpool(self, x, y) -> None:
        xv, yv = as_view(x), as_view(y)
        self.fwd.append(lambda: ops.maxpool3x3s2(xv, yv))

        def bwd():
            if not self.has_grad(yv):
                return
            dy, dx = self.grad_view(yv), self.grad_view(xv)
            acc = self.is_written(dx)
            self.mark_written(dx)
            idx = self.tensor((xv.n * ((xv.h - 1) // 2 + 1) * ((xv.w - 1) // 2 + 1) * xv.c,), dtype=torch.uint8)
            self.bwd.append(lambda: ops.maxpool3x3s2_bwd(xv, dy, dx, accumulate=acc, idx=idx))
        self.tape.append(bwd)

    def upsample(self, x, y) -> None:
        xv, yv = as_view(x), as_view(y)
        self.fwd.append(lambda: ops.upsample_bilinear_ac(xv, yv))

        def bwd():
            if not self.has_grad(yv):
                return
            dy, dx = self.grad_view(yv), self.grad_view(xv)
            acc = self.is_written(dx)
            self.mark_written(dx)
            self.bwd.append(lambda: ops.upsample_bilinear_ac_bwd(dy, dx, accumulate=acc))
        self.tape.append(bwd)

    def global_avgpool(self, x, g) -> None:
        xv, gv = as_view(x), as_view(g)
        self.fwd.append(lambda: ops.global_avgpool(xv, gv))

        def bwd():
            if not self.has_grad(gv):
                return
            dg, dx = self.grad_view(gv), self.grad_view(xv)
            acc = self.is_written(dx)
            self.mark_written(dx)
            self.bwd.append(lambda: ops.add_broadcast(dg, dx, 1.0 / (xv.h * xv.w), accumulate=acc))
        self.tape.append(bwd)

    def broadcast_hw(self, g, y) -> None:
        gv, yv = as_view(g), as_view(y)
        self.fwd.append(lambda: ops.broadcast_hw(gv, yv))

        def bwd():
            if not self.has_grad(yv):
                return
            dy, dg = self.grad_view(yv), self.grad_view(gv)
            assert not self.is_written(dg)
            self.mark_written(dg)
            self.bwd.append(lambda: ops.global_sumpool(dy, dg))
        self.tape.append(bwd)

    # ---- assembly --------------------------------------------------------------------------------
    def finalize(self, flat: bool = False) -> None:
        """Emit the backward ops (reverse tape).  flat=True: every live parameter's gradient is a view of ONE flat
        fp32 buffer (self.flat_g) - what the all-reduce and the fused Adam step operate on."""
        if flat:
            # laid out in REVERSE registration order = the order in which the backward pass finishes the gradients, so
            # that contiguous buckets of the buffer become final one after the other (bucketed all-reduce, TrainStep)
            total = sum(q.numel() for q in self.live_params)
            self.flat_g = torch.zeros(total, dtype=torch.float32, device=self.device)
            off = 0
            self.flat_offsets: Dict[int, Tuple[int, int]] = {}
            for q in reversed(self.live_params):
                self.pgrad[id(q)] = self.flat_g[off:off + q.numel()].view(q.shape)
                self.flat_offsets[id(q)] = (off, q.numel())
                self.params.append(q)
                off += q.numel()
        for t in reversed(self.tape):
            t()
        self.scratch = torch.empty(max(self.scratch_bytes // 4, 1), dtype=torch.float32, device=self.device)
        self.scratch_main = torch.empty(max(self.scratch_main_bytes // 4, 1), dtype=torch.float32, device=self.device)

    def run_forward(self, x: torch.Tensor, masks: Optional[Sequence[torch.Tensor]] = None) -> torch.Tensor:
        with torch.cuda.device(self.device):     # C-ABI launches use the current device's stream
            return self._run_forward(x, masks)

    def _run_forward(self, x, masks):
        self.input.copy_(x)
        self.weights.refresh()
        self._fill_masks(masks)
        self.fwd_counter += 1
        for op in self.fwd:
            op()
        for bn in self.bn_modules:
            if bn.num_batches_tracked is not None:
                bn.num_batches_tracked.add_(1)
        engine.note_raw_parameter_update()      # running statistics changed behind torch's version counters
        return self.heat

    def run_backward(self, dheat: torch.Tensor) -> None:
        with torch.cuda.device(self.device):
            self.dheat.copy_(dheat)
            self.run_backward_ops(0, len(self.bwd))

    def run_backward_ops(self, lo: int, hi: int) -> None:
        """Backward ops [lo, hi) in order; the weight gradients go to the side stream (forked behind the op that
        produced their dz, joined at the end of the range - so a range is also capturable as one CUDA graph)."""
        if not (self.overlap_wgrad and any(i in self.bwd_side for i in range(lo, hi))):
            for op in self.bwd[lo:hi]:
                op()
            return
        if self.side_stream is None:
            self.side_stream = torch.cuda.Stream(device=self.device)
        main, side = torch.cuda.current_stream(self.device), self.side_stream
        for i in range(lo, hi):
            if i in self.bwd_side:
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    self.bwd[i]()
            else:
                self.bwd[i]()
        main.wait_stream(side)

    def _fill_masks(self, given) -> None:
        for i, (act, p) in enumerate(self.masks):
            if given is not None:
                ops.nchw_to_act(given[i].float().contiguous(), act)
                continue
            keep = (torch.rand((act.n, act.h, act.w, act.c), device=self.device) >= p)
            s = 1.0 / (1.0 - p)
            if self.mode == ops.UP_SPLIT:
                # hi / lo planes of the scale as host scalars (no host->device copy: the step is graph-captured)
                hi = float(torch.tensor(s, dtype=torch.bfloat16))
                lo = float(torch.tensor(s - hi, dtype=torch.bfloat16))
                act.t[0] = keep.to(act.t.dtype) * hi
                act.t[1] = keep.to(act.t.dtype) * lo
            else:
                act.t[0] = keep.to(act.t.dtype) * s


# ----------------------------------------------------------------------------------------------------
# network emission (mirrors the eval-mode plan of model/unipose.py, with train-mode BatchNorm + dropout)
# ----------------------------------------------------------------------------------------------------
def _stem_window_weight_inv(g: torch.Tensor) -> torch.Tensor:
    """Adjoint of resnet.stem_window_weight: [co,64,4,1] gradient -> [co,3,7,7]."""
    co = g.shape[0]
    g2 = g.reshape(co, 4, 16, 4).permute(0, 2, 3, 1)      # [co, ch16, kh', kw']
    out = g.new_zeros((co, 3, 7, 7))
    for a in range(4):
        for ph in range(2):
            kh = 2 * (a - 2) + ph + 3
            if not 0 <= kh < 7:
                continue
            for bb in range(4):
                for pw in range(2):
                    kw = 2 * (bb - 2) + pw + 3
                    if 0 <= kw < 7:
                        c0 = (ph * 2 + pw) * 3
                        out[:, :, kh, kw] = g2[:, c0:c0 + 3, a, bb]
    return out


def _emit_trunk(tp: TrainPlan, model, n: int, h: int, w: int, head_out: torch.Tensor, head_grad=None) -> None:
    """backbone -> WASP -> decoder in train mode (model/unipose.py:27-38 / model/uniposeLSTM.py:108-113) over tp.input
    [n,3,h,w]; the head conv writes fp32 NCHW into the first channels of `head_out`, its gradient is read from
    `head_grad` (default tp.dheat)."""
    from .model.modules.backbone.resnet import stem_window_weight
    bb, wasp, dec = model.backbone, model.wasp, model.decoder
    # ---- backbone ----
    x2 = tp.act(n, h // 2, w // 2 + 3, 16, zero=True)
    tp.fwd.append(lambda: ops.pack_input_s2d(tp.input, x2, wpad_left=2))
    stem = tp.conv_unit(x2, bb.conv1, bb.bn1, relu=True, pad=(2, 0), x_window=(w // 2, 64), cin_pad=64,
                        weight_fn=stem_window_weight, weight_fn_inv=_stem_window_weight_inv, x_needs_grad=False)
    x = tp.act(n, h // 4, w // 4, 64)
    tp.maxpool(stem, x)
    low = None
    for name in ("layer1", "layer2", "layer3", "layer4"):
        for blk in getattr(bb, name):
            s, d = blk.stride, blk.dilation
            t1 = tp.conv_unit(x, blk.conv1, blk.bn1, relu=True)
            t2 = tp.conv_unit(t1, blk.conv2, blk.bn2, relu=True, stride=s, dil=d, pad=d)
            res = x
            if blk.downsample is not None:
                res = tp.conv_unit(x, blk.downsample[0], blk.downsample[1], relu=False, stride=s, pad=0)
            x = tp.conv_unit(t2, blk.conv3, blk.bn3, relu=True, residual=res)
        if name == "layer1":
            low = x
    feat = x
    hh, ww = feat.h, feat.w

    # ---- WASP (wasp.py:66-90; waspVideo.py:67-91: no BatchNorm in the pooling branch) ----
    S = tp.act(4 * n, hh, ww, 256)
    br = [View(S, n_off=i * n, n=n) for i in range(4)]
    aspp = [wasp.aspp1, wasp.aspp2, wasp.aspp3, wasp.aspp4]
    src = feat
    for i, a in enumerate(aspp):
        c = a.atrous_conv
        tp.conv_unit(src, c, a.bn, relu=True, dil=c.dilation[0], pad=c.padding[0], out=br[i])
        src = br[i]
    T = tp.conv_unit(S, wasp.conv2, None, relu=False)
    U = tp.act(5 * n, hh, ww, 256)
    tp.conv_unit(T, wasp.conv2, None, relu=False, out=View(U, n_off=0, n=4 * n))
    g = tp.act(n, 1, 1, feat.c)
    tp.global_avgpool(feat, g)
    gap_seq = wasp.global_avg_pool
    gap_bn = gap_seq[2] if isinstance(gap_seq[2], nn.BatchNorm2d) else None
    g2 = tp.conv_unit(g, gap_seq[1], gap_bn, relu=True)
    tp.broadcast_hw(g2, View(U, n_off=4 * n, n=n))
    mask_w = tp.act(n, hh, ww, 256)
    tp.masks.append((mask_w, wasp.dropout.p))
    wout = tp.conv_unit(View(U, n_off=0, n=n), wasp.conv1, wasp.bn1, relu=True, x_groups=5, x_group_nstride=n,
                        mask=mask_w)

    # ---- decoder (decoder.py:38-56) ----
    lo = tp.conv_unit(low, dec.conv1, dec.bn1, relu=True, cout_pad=64)
    hc, wc = (low.h - 1) // 2 + 1, (low.w - 1) // 2 + 1
    cat = tp.act(n, hc, wc, 320, zero=True)
    tp.maxpool(lo, View(cat, coff=256, c=64))
    tp.upsample(wout, View(cat, coff=0, c=256))
    lc = dec.last_conv
    m1 = tp.act(n, hc, wc, 256)
    m2 = tp.act(n, hc, wc, 256)
    tp.masks.append((m1, lc[3].p))
    tp.masks.append((m2, lc[7].p))
    d1 = tp.conv_unit(cat, lc[0], lc[1], relu=True, pad=1, cin_pad=320, mask=m1)
    d2 = tp.conv_unit(d1, lc[4], lc[5], relu=True, pad=1, mask=m2)
    tp.conv_unit(d2, lc[8], None, relu=False, nchw_out=head_out, nchw_grad=head_grad)


def build_image_train_plan(model, shape, device, precision: str, flat: bool = False) -> TrainPlan:
    n, _, h, w = shape
    if h % 16 or w % 16:
        raise ValueError("unipose_b200: input height/width must be multiples of 16")
    tp = TrainPlan(device, precision)
    tp.input = torch.zeros(shape, dtype=torch.float32, device=device)
    tp.heat = torch.zeros((n, model.decoder.num_out, h // 8, w // 8), dtype=torch.float32, device=device)
    tp.dheat = torch.zeros_like(tp.heat)
    _emit_trunk(tp, model, n, h, w, tp.heat)
    tp.finalize(flat=flat)
    return tp


def build_video_frame_train_plan(model, b_: int, h: int, w: int, first: bool, device, precision: str) -> TrainPlan:
    """One frame of model/uniposeLSTM.unipose.forward (:98-147) in train mode with everything its backward needs:
    trunk -> cat(heat-maps, pooled centre map) -> LSTM_0 / LSTM (activated gates saved) -> 11x11 / 1x1 middle CNN with
    bias + ReLU epilogues.  The frames of a clip are chained by torch autograd through (cell, hide) - the reference
    runs ONE backward over all 5 frames (uniposeLSTM.py:116-132)."""
    if h % 16 or w % 16:
        raise ValueError("unipose_b200: input height/width must be multiples of 16")
    tp = TrainPlan(device, precision)
    hs, ws = h // 8, w // 8
    k1 = model.decoder.num_out
    cell_mod = model.lstm_0 if first else model.lstm
    lstm_c = model.lstm_0.conv_g_lstm.in_channels
    planes = model.lstm_0.conv_g_lstm.out_channels
    ng = 3 if first else 4

    def f32(*shape):
        return torch.zeros(shape, dtype=torch.float32, device=device)
    tp.input = f32(b_, 3, h, w)
    tp.cmap = f32(b_, 1, h, w)
    tp.cat, tp.dcat = f32(b_, lstm_c, hs, ws), f32(b_, lstm_c, hs, ws)
    tp.heat = f32(b_, model.conv5.out_channels, hs, ws)
    tp.dheat = torch.zeros_like(tp.heat)
    tp.cell, tp.hide = f32(b_, planes, hs, ws), f32(b_, planes, hs, ws)
    tp.hp, tp.cp = f32(b_, planes, hs, ws), f32(b_, planes, hs, ws)
    tp.dhide_ext, tp.dcell_ext = f32(b_, planes, hs, ws), f32(b_, planes, hs, ws)
    tp.dhide_mid, tp.dhide_tot = f32(b_, planes, hs, ws), f32(b_, planes, hs, ws)
    tp.dhp, tp.dcp = f32(b_, planes, hs, ws), f32(b_, planes, hs, ws)
    gates, dpre = f32(b_, ng, planes, hs, ws), f32(b_, ng, planes, hs, ws)

    # ---- trunk: heat-maps into channels [0, K+1) of the LSTM input, centre map into the last one ----
    _emit_trunk(tp, model, b_, h, w, tp.cat, head_grad=tp.dcat)
    tp.fwd.append(lambda: ops._lib.call("up_avgpool9s8p1_f32", ops._ptr(tp.cmap), ops._ptr(tp.cat), b_, 1, h, w, hs, ws,
                                        lstm_c, k1, ops._stream()))

    # ---- ConvLSTM cell ----
    gate_w = [torch.empty_like(t) for t in cell_mod.stacked()]
    gate_g = [torch.zeros_like(t) for t in gate_w]

    def restack():
        for dst, src in zip(gate_w, cell_mod.stacked()):
            dst.copy_(src)
    tp.fwd.append(restack)
    if first:
        tp.fwd.append(lambda: model.lstm_0.launch(tp.cat, tp.cell, tp.hide, stacked=gate_w, gates=gates))
    else:
        tp.fwd.append(lambda: model.lstm.launch(tp.cat, tp.hp, tp.cp, tp.cell, tp.hide, stacked=gate_w, gates=gates))
    lstm_params = cell_mod.sources()
    for prm in lstm_params:
        if id(prm) not in tp._live_ids:
            tp._live_ids.add(id(prm))
            tp.live_params.append(prm)

    def lstm_bwd():
        tp.bwd.append(lambda: torch.add(tp.dhide_mid, tp.dhide_ext, out=tp.dhide_tot))
        null = None
        if first:
            args = (tp.cat, null, null, gates, tp.cell, tp.dcell_ext, tp.dhide_tot, gate_w[0], null, tp.dcat, null, null,
                    gate_g[0], gate_g[1], null, null, dpre)
        else:
            args = (tp.cat, tp.hp, tp.cp, gates, tp.cell, tp.dcell_ext, tp.dhide_tot, gate_w[0], gate_w[2], tp.dcat, tp.dhp,
                    tp.dcp, gate_g[0], gate_g[1], gate_g[2], gate_g[3], dpre)
        tp.bwd.append(lambda: ops._lib.call("up_convlstm_cell_bwd", *[ops._ptr(a) for a in args], b_, lstm_c, planes, hs, ws,
                                            ops._stream()))
        # the stacked gradients back to the reference's per-gate parameters
        if first:
            convs = [(c, 0, i) for i, c in enumerate(model.lstm_0._gates())]
        else:
            convs = [(getattr(model.lstm, 'conv_%s%s_lstm' % (g, sfx)), 0 if sfx == 'x' else 2, i)
                     for sfx in 'xh' for i, g in enumerate('giof')]
        for conv, base, i in convs:
            gw, gb = tp.param_grad(conv.weight), tp.param_grad(conv.bias)
            tp.bwd.append(lambda gw=gw, gb=gb, base=base, i=i: (gw.copy_(gate_g[base][i]), gb.copy_(gate_g[base + 1][i])))
            tp._grad_final(conv.weight)
            tp._grad_final(conv.bias)
    tp.tape.append(lstm_bwd)

    # ---- middle CNN: conv 11x11 x3, 1x1 x2, ReLU after every one including the last (uniposeLSTM.py:120-124) ----
    hin = tp.act(b_, hs, ws, 64, zero=True)
    tp.fwd.append(lambda: ops.nchw_to_act(tp.hide, hin))

    def hin_bwd():
        if tp.has_grad(hin):
            ghin = tp.grad_view(hin)
            tp.bwd.append(lambda: ops.act_to_nchw(ghin, planes, tp.dhide_mid))
        else:
            tp.bwd.append(lambda: tp.dhide_mid.zero_())
    tp.tape.append(hin_bwd)
    a1 = tp.conv_unit(hin, model.conv1, None, relu=True, pad=5, cin_pad=64)
    a2 = tp.conv_unit(a1, model.conv2, None, relu=True, pad=5)
    a3 = tp.conv_unit(a2, model.conv3, None, relu=True, pad=5)
    a4 = tp.conv_unit(a3, model.conv4, None, relu=True)
    tp.conv_unit(a4, model.conv5, None, relu=True, nchw_out=tp.heat)
    tp.finalize(flat=False)
    return tp


class _TrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan: TrainPlan, masks, x, *params):
        ctx.plan = plan
        heat = plan.run_forward(x.detach().float(), masks)
        ctx.counter = plan.fwd_counter
        return heat.clone()

    @staticmethod
    def backward(ctx, dheat):
        plan: TrainPlan = ctx.plan
        if plan.fwd_counter != ctx.counter:
            # the plan keeps ONE set of saved activations per input shape: a second forward overwrote them
            raise RuntimeError("unipose_b200: backward() after another forward() of the same shape - the training plan "
                               "keeps one set of activations (call backward before the next forward)")
        plan.run_backward(dheat.detach().float().contiguous())
        grads = tuple(plan.pgrad[id(p)].clone() for p in plan.params)
        return (None, None, None) + grads


def forward_train(model, input: torch.Tensor, dropout_masks=None) -> torch.Tensor:
    """Train-mode forward of model/unipose.unipose; the result carries a grad_fn whose backward runs the B200 plan.
    `dropout_masks`: optional three pre-scaled fp32 NCHW masks (wasp, decoder 0.5, decoder 0.1) for parity tests."""
    if getattr(model, "stride", 8) != 8:
        raise NotImplementedError("unipose_b200: training supports stride=8 outputs (the reference's training setting)")
    frozen_sig = tuple(m.training for m in model.modules() if isinstance(m, nn.BatchNorm2d))
    key = ("train", tuple(input.shape), model._precision(), input.device.index, frozen_sig)
    plan = model._plans.get(key)
    if plan is None:
        plan = build_image_train_plan(model, tuple(input.shape), input.device, model._precision())
        model._plans[key] = plan
    return _TrainFn.apply(plan, dropout_masks, input, *plan.params)


class _VideoFrameFn(torch.autograd.Function):
    """One frame of the video model; (cell, hide) carry the autograd chain from frame to frame."""

    @staticmethod
    def forward(ctx, plan: TrainPlan, masks, frame, cmap, hide_prev, cell_prev, *params):
        with torch.cuda.device(plan.device):
            plan.cmap.copy_(cmap.detach().float())
            if hide_prev is not None:
                plan.hp.copy_(hide_prev.detach().float().reshape(plan.hp.shape))
                plan.cp.copy_(cell_prev.detach().float().reshape(plan.cp.shape))
        plan.run_forward(frame.detach().float(), masks)
        ctx.plan = plan
        ctx.counter = plan.fwd_counter
        ctx.recurrent = hide_prev is not None
        return plan.heat.clone(), plan.cell.clone(), plan.hide.clone()

    @staticmethod
    def backward(ctx, dheat, dcell, dhide):
        plan: TrainPlan = ctx.plan
        if plan.fwd_counter != ctx.counter:
            raise RuntimeError("unipose_b200: backward() after another forward() of the same frame slot - the training "
                               "plan keeps one set of activations per frame index")
        with torch.cuda.device(plan.device):
            plan.dcell_ext.copy_(dcell.detach().float())
            plan.dhide_ext.copy_(dhide.detach().float())
        plan.run_backward(dheat.detach().float().contiguous())
        grads = tuple(plan.pgrad[id(p)].clone() for p in plan.params)
        rec = (plan.dhp.clone(), plan.dcp.clone()) if ctx.recurrent else (None, None)
        return (None, None, None, None) + rec + grads


def forward_train_video(model, input, centermap, it: int, prev_hide, prev_cell, dropout_masks=None):
    """Train-mode forward of model/uniposeLSTM.unipose for frame `it` of the clip `input` [B,T,3,H,W]; the returned
    (heat, cell, hide) carry grad_fns, so the reference's loop - five calls, summed MSE, one backward
    (uniposeLSTM.py:116-132) - runs unchanged."""
    b_, _t, _c, h, w = input.shape
    first = (it == 0)
    frozen_sig = tuple(m.training for m in model.modules() if isinstance(m, nn.BatchNorm2d))
    key = ("train_video", b_, h, w, int(it), model._precision(), input.device.index, frozen_sig)
    plan = model._plans.get(key)
    if plan is None:
        plan = build_video_frame_train_plan(model, b_, h, w, first, input.device, model._precision())
        model._plans[key] = plan
    hp = cp = None
    if not first:
        planes = model.lstm_0.conv_g_lstm.out_channels
        hs, ws = h // 8, w // 8

        def state(t, what):
            if t.dim() == 3:
                t = t.unsqueeze(0)
            if tuple(t.shape) != (b_, planes, hs, ws):
                raise ValueError("unipose_b200: %s must have shape %s (got %s)" % (what, (b_, planes, hs, ws), tuple(t.shape)))
            return t
        hp, cp = state(prev_hide, "previousHide"), state(prev_cell, "previousCell")
    return _VideoFrameFn.apply(plan, dropout_masks, input[:, it], centermap[:, it], hp, cp, *plan.params)


# ----------------------------------------------------------------------------------------------------
# fused training step: forward, MSE, backward, (all-reduce), Adam  -  unipose.py:107-124
# ----------------------------------------------------------------------------------------------------
class TrainStep:
    """One optimisation step of the reference's training loop (unipose.py:107-124) on flat fp32 buffers.

    `step(input, target)` = forward (train-mode BN, dropout) -> nn.MSELoss (up_mse_fwd_bwd) -> backward -> NCCL
    all-reduce of the gradient (data parallel over the batch; BatchNorm stays per-GPU as in the reference, which has
    no SyncBN) -> Adam (up_adam_step) on the flat parameter buffer.  Parameters the reference leaves without gradient
    (decoder.conv2 / bn2) are never touched.

    The flat gradient buffer is laid out in the order the backward pass completes the gradients and cut into
    ~25 MB buckets; the backward op list is split where a bucket becomes final, each segment replays
    as its own CUDA graph, and the bucket's all-reduce is enqueued on a side stream right behind it - the collective
    of bucket k overlaps the backward of buckets k+1.. (torch.distributed / NCCL over NVLink)."""

    BUCKET_BYTES = 25 << 20

    def __init__(self, model, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, bucket_bytes: Optional[int] = None):
        self.model, self.lr, self.betas, self.eps = model, lr, betas, eps
        self.plan: Optional[TrainPlan] = None
        self.t = 0
        self.bucket_bytes = int(bucket_bytes or os.environ.get("UNIPOSE_B200_BUCKET_BYTES", self.BUCKET_BYTES))
        self.comm_enabled = True          # bench: switch the collective off to measure what it costs

    # ---- setup ------------------------------------------------------------------------------------------------
    def _setup(self, x: torch.Tensor) -> None:
        from . import parallel
        m = self.model
        self.plan = build_image_train_plan(m, tuple(x.shape), x.device, m._precision(), flat=True)
        p = self.plan
        self.flat_g = p.flat_g
        self.flat_p = torch.empty_like(self.flat_g)
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        off = 0
        with torch.no_grad():
            for q in p.params:        # the parameters become views of the flat buffer (state_dict() still works)
                nq = q.numel()
                self.flat_p[off:off + nq].copy_(q.detach().reshape(-1))
                q.data = self.flat_p[off:off + nq].view(q.shape)
                off += nq
        world = parallel.world()[1]
        if world > 1:
            # what DistributedDataParallel does at construction: every rank starts from rank 0's parameters / buffers
            import torch.distributed as dist
            dist.broadcast(self.flat_p, 0)
            for b in m.buffers():
                if b.is_floating_point() or b.dtype == torch.long:
                    dist.broadcast(b, 0)
        self.loss = torch.zeros(1, dtype=torch.float32, device=x.device)
        self.loss_scratch = torch.zeros(1, dtype=torch.float64, device=x.device)
        self._make_buckets()
        self.comm_stream = torch.cuda.Stream(device=x.device) if world > 1 else None

    def _make_buckets(self) -> None:
        """Buckets = contiguous ranges of the flat gradient; a bucket is final once backward op `ready` has run."""
        p = self.plan
        items = []      # (offset, numel, ready index) in layout order
        for q in p.params:
            off, nq = p.flat_offsets[id(q)]
            items.append((off, nq, p.grad_ready.get(id(q), len(p.bwd))))
        from . import parallel
        self.buckets = parallel.make_buckets(items, self.bucket_bytes, len(p.bwd))

    # ---- one step ---------------------------------------------------------------------------------------------
    def _fwd_loss(self, world: int) -> None:
        p = self.plan
        heat = p.run_forward(p.input)
        ops._lib.call("up_mse_fwd_bwd", ops._ptr(heat), ops._ptr(self.target_buf), ops._ptr(self.loss), ops._ptr(p.dheat),
                      ops._ptr(self.loss_scratch), heat.numel(), 1.0 / world, ops._stream())

    def _bwd_segment(self, k: int) -> None:
        lo = self.buckets[k - 1][2] if k > 0 else 0
        self.plan.run_backward_ops(lo, self.buckets[k][2])

    def _run_piece(self, k: int, world: int) -> None:
        """piece 0 = forward + loss + backward segment 0; piece k = backward segment k."""
        if k == 0:
            self._fwd_loss(world)
        self._bwd_segment(k)

    def step(self, x: torch.Tensor, target: torch.Tensor, lr: Optional[float] = None) -> torch.Tensor:
        from . import parallel
        if self.plan is None:
            self._setup(x)
            self.target_buf = torch.empty_like(target, dtype=torch.float32)
            self.graphs: Optional[List[torch.cuda.CUDAGraph]] = None
            self.steps_done = 0
            # the pieces replay as CUDA graphs from the third step on; the all-reduces and Adam (step-dependent bias
            # correction) stay eager.  UNIPOSE_B200_TRAIN_GRAPH=0 disables.
            self.use_graph = os.environ.get("UNIPOSE_B200_TRAIN_GRAPH", "1") != "0"
        p = self.plan
        world = parallel.world()[1]
        with torch.cuda.device(p.device):
            p.input.copy_(x)
            self.target_buf.copy_(target)
            if self.use_graph and self.graphs is None and self.steps_done >= 2:
                torch.cuda.synchronize(p.device)
                graphs, pool = [], None
                for k in range(len(self.buckets)):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, pool=pool):
                        self._run_piece(k, world)
                    pool = g.pool()
                    graphs.append(g)
                self.graphs = graphs          # capturing does not execute: fall through to the replays below
            main = torch.cuda.current_stream(p.device)
            works = []
            for k, (lo, hi, _ready) in enumerate(self.buckets):
                if self.graphs is not None:
                    self.graphs[k].replay()
                else:
                    self._run_piece(k, world)
                if world > 1 and self.comm_enabled:
                    # bucket k is final: its all-reduce runs on the side stream while the next segments compute
                    import torch.distributed as dist
                    self.comm_stream.wait_stream(main)
                    with torch.cuda.stream(self.comm_stream):
                        works.append(dist.all_reduce(self.flat_g[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
            for w in works:
                w.wait()                      # the side stream waits for NCCL ...
            if works:
                main.wait_stream(self.comm_stream)   # ... and Adam for the side stream
            self.steps_done += 1
            engine.note_raw_parameter_update()     # Adam / BatchNorm statistics move behind torch's version counters
            self.t += 1
            # gradients were pre-scaled by 1/world in the MSE kernel: the sum is the global-batch mean gradient
            ops._lib.call("up_adam_step", ops._ptr(self.flat_p), ops._ptr(self.flat_g), ops._ptr(self.exp_avg),
                          ops._ptr(self.exp_avg_sq), self.flat_p.numel(), float(self.lr if lr is None else lr),
                          float(self.betas[0]), float(self.betas[1]), float(self.eps), self.t, ops._stream())
        return self.loss

    @property
    def graph(self):
        """Truthy once the step replays CUDA graphs (kept for callers of the one-graph version)."""
        return self.graphs

    def comm_stats(self) -> dict:
        return {"buckets": len(self.buckets) if self.plan is not None else None,
                "bucket_mb": [round((hi - lo) * 4 / 2 ** 20, 1) for lo, hi, _ in self.buckets] if self.plan else None}

    def time_allreduce_alone(self, reps: int = 5) -> float:
        """ms of all bucket all-reduces back to back with nothing else on the GPU (max over ranks is the caller's job)."""
        from . import parallel
        if parallel.world()[1] <= 1 or self.plan is None:
            return 0.0
        import torch.distributed as dist
        scratch = torch.zeros_like(self.flat_g)
        for _ in range(2):
            for lo, hi, _r in self.buckets:
                dist.all_reduce(scratch[lo:hi])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            for lo, hi, _r in self.buckets:
                dist.all_reduce(scratch[lo:hi])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
