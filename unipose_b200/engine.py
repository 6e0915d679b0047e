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


# Bumped whenever a kernel updates parameters / running statistics through raw pointers (up_adam_step,
# up_bn_finalize): torch's per-tensor version counters do not see those writes, the packed-weight caches must.
_RAW_UPDATE_EPOCH = [0]


def note_raw_parameter_update() -> None:
    _RAW_UPDATE_EPOCH[0] += 1


def _versions(tensors: Sequence[torch.Tensor]) -> Tuple[int, ...]:
    return (_RAW_UPDATE_EPOCH[0],) + tuple((t.data_ptr(), t._version) for t in tensors)


class _PackJob:
    """Keeps one derived buffer in sync with its source tensors (torch-side jobs: e.g. stacked ConvLSTM gates).
    with_table=True: the job derives from buffers the plan's WeightTable fills (e.g. the sum of two epilogue shifts) - it
    re-runs whenever the table has refreshed, which also covers parameters that were REPLACED (the table looks its
    tensors up through the modules at every call; `sources` are then only what the table does not watch)."""

    def __init__(self, sources: Sequence[torch.Tensor], fn: Callable[[], None], with_table: bool = False):
        self.sources = list(sources)
        self.fn = fn
        self.with_table = with_table
        self.seen = None

    def refresh(self, table_changed: bool = False) -> bool:
        v = _versions(self.sources) if self.sources else ()
        if v != self.seen or (self.with_table and table_changed):
            self.fn()
            self.seen = v
            return True
        return False


def _param(mod: nn.Module, name: str):
    """mod.<name> for a registered parameter / buffer without nn.Module.__getattr__'s slow path (the version scan below
    touches ~600 tensors on every forward call)."""
    t = mod._parameters.get(name)
    if t is None:
        t = mod._buffers.get(name)
    return t if t is not None else getattr(mod, name)


class WeightTable:
    """Every packed filter and every epilogue constant of a plan, refreshed by TWO kernel launches
    (up_epilogue_consts, up_pack_conv_weights) driven by job tables in device memory - instead of ~8 small launches
    per layer.  Entries whose filter needs a host-side transform (`weight_fn`: the stem's super-pixel regrouping,
    WASP's folded conv1) first materialise the transformed OIHW filter with torch ops.

    `always=True` (training plans: the weights change every step through raw-pointer kernels) skips the version
    check; the tables are then uploaded once and only re-uploaded when a parameter's storage moves."""

    def __init__(self, device, mode: int, always: bool = False):
        self.device = torch.device(device)
        self.mode = mode
        self.always = always
        self.entries: List[dict] = []
        self.epi: List[dict] = []
        self._seen = None
        self._ptr_key = None
        self._d_pack: Optional[torch.Tensor] = None
        self._d_epi: Optional[torch.Tensor] = None
        self._total_tiles = 0
        self._max_c = 0
        self._keep: List[torch.Tensor] = []

    # ---- registration -----------------------------------------------------------------------------------
    def add_pack(self, src: Callable[[], torch.Tensor], out: torch.Tensor, rows: int, cols: int, *, transpose=False,
                 ci_off: int = 0, cin_slice: Optional[int] = None, row_scale: Optional[torch.Tensor] = None,
                 scale_period: int = 1, weight_fn=None, watch: Sequence[torch.Tensor] = ()) -> None:
        """src() -> OIHW fp32 tensor (a parameter's .detach()); weight_fn (optional) maps it to the filter to pack."""
        self.entries.append(dict(src=src, out=out, rows=rows, cols=cols, transpose=bool(transpose), ci_off=ci_off,
                                 cin_slice=cin_slice, row_scale=row_scale, scale_period=scale_period,
                                 weight_fn=weight_fn, watch=list(watch)))

    def add_epilogue(self, scale: torch.Tensor, shift: torch.Tensor, cout_real: int, *, bn=None, bias=None,
                     fold_scale: Optional[torch.Tensor] = None, fold_into_weights: bool = True) -> None:
        self.epi.append(dict(scale=scale, shift=shift, cout_real=cout_real, bn=bn, bias=bias, fold_scale=fold_scale,
                             fold_into_weights=fold_into_weights))

    # ---- refresh ---------------------------------------------------------------------------------------------
    def _watched(self) -> List[torch.Tensor]:
        ts = []
        for e in self.entries:
            ts.append(e["src"]())
            ts.extend(e["watch"])
        for j in self.epi:
            if j["bn"] is not None:
                bn = j["bn"]
                ts.extend([_param(bn, "weight"), _param(bn, "bias"), _param(bn, "running_mean"), _param(bn, "running_var")])
            if j["bias"] is not None:
                ts.append(j["bias"]())
        return ts

    def _f32(self, t: torch.Tensor) -> torch.Tensor:
        t = t.detach()
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.float().contiguous()
            self._keep.append(t)          # the kernel reads it asynchronously
        return t

    def _upload(self, structs, holder: str) -> torch.Tensor:
        import ctypes as C
        arr = (type(structs[0]) * len(structs))(*structs)
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        dev = getattr(self, holder)
        if dev is None or dev.numel() != host.numel():
            dev = torch.empty(host.numel(), dtype=torch.uint8, device=self.device)
            setattr(self, holder, dev)
        dev.copy_(host)                   # H2D memcpy (pageable: synchronous w.r.t. the host buffer)
        return dev

    def _build_tables(self, sources: List[torch.Tensor]) -> None:
        from ._lib import UpEpilogueJob, UpPackJob
        lib = ops._lib.load()
        packs, tiles = [], 0
        import ctypes as C
        for e, w in zip(self.entries, sources):
            co_r, ci_t, kh, kw = w.shape
            j = UpPackJob()
            j.w, j.out = w.data_ptr(), e["out"].data_ptr()
            j.row_scale = 0 if e["row_scale"] is None else e["row_scale"].data_ptr()
            j.plane_stride = kh * kw * e["rows"] * e["cols"]
            j.tile_start = tiles
            j.kh, j.kw, j.rows, j.cols = kh, kw, e["rows"], e["cols"]
            j.cout_real, j.cin_total, j.ci_off = co_r, ci_t, e["ci_off"]
            j.cin_slice = (ci_t - e["ci_off"]) if e["cin_slice"] is None else e["cin_slice"]
            j.scale_period, j.dtype, j.transpose = e["scale_period"], self.mode, int(e["transpose"])
            t = int(lib.up_pack_job_tiles(C.byref(j)))
            assert t > 0, "bad pack job"
            tiles += t
            packs.append(j)
        self._total_tiles = tiles
        if packs:
            self._upload(packs, "_d_pack")
        epis, self._max_c = [], 0
        for e in self.epi:
            j = UpEpilogueJob()
            bn = e["bn"]
            if bn is not None:
                j.gamma, j.beta = self._f32(bn.weight).data_ptr(), self._f32(bn.bias).data_ptr()
                j.mean, j.var = self._f32(bn.running_mean).data_ptr(), self._f32(bn.running_var).data_ptr()
                j.eps, j.c_bn = float(bn.eps), bn.num_features
                j.fold_scale = 0 if e["fold_scale"] is None else e["fold_scale"].data_ptr()
                j.fold_into_weights = int(e["fold_into_weights"])
            elif e["bias"] is not None:
                b = self._f32(e["bias"]())
                j.bias, j.bias_len = b.data_ptr(), b.numel()
            j.scale, j.shift = e["scale"].data_ptr(), e["shift"].data_ptr()
            j.cout_real, j.cout = e["cout_real"], e["scale"].numel()
            self._max_c = max(self._max_c, j.cout, j.c_bn)
            epis.append(j)
        if epis:
            self._upload(epis, "_d_epi")

    def refresh(self) -> bool:
        if not self.entries and not self.epi:
            return False
        watched = self._watched()
        if not self.always:
            v = _versions(watched)
            if v == self._seen:
                return False
            self._seen = v
        # sources of the pack kernel: the parameter itself, or the torch-side transform of it
        self._keep = []
        sources = []
        for e in self.entries:
            w = e["src"]()
            if e["weight_fn"] is not None:
                # torch-side transform into a PERSISTENT buffer: the job table (and a captured CUDA graph) keeps its address
                t = e["weight_fn"](w.detach().float())
                if e.get("tmp") is None or e["tmp"].shape != t.shape:
                    e["tmp"] = torch.empty_like(t, memory_format=torch.contiguous_format)
                e["tmp"].copy_(t)
                w = e["tmp"]
            else:
                w = self._f32(w)
            sources.append(w)
        key = tuple(t.data_ptr() for t in sources) + tuple(t.data_ptr() for t in watched)
        if key != self._ptr_key:
            self._build_tables(sources)
            self._ptr_key = key
        st = ops._stream()
        if self.epi:
            ops._lib.call("up_epilogue_consts", ops._ptr(self._d_epi), len(self.epi), self._max_c, st)
        if self.entries:
            ops._lib.call("up_pack_conv_weights", ops._ptr(self._d_pack), len(self.entries), self._total_tiles, st)
        return True


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
                    nchw_out: bool = False, extra_sources: Sequence[torch.Tensor] = (), wbuf=None, scale=None,
                    shift=None) -> PackedConv:
        """Allocate packed buffers for `conv` (+`bn` folded as eval-mode scale/shift) - or use the given `wbuf` / `scale`
        / `shift` views of larger buffers (fused multi-layer kernels read their filters from one contiguous array) - and register them in the plan's
        weight table, which (re)fills them from the live parameters with two table-driven launches.

        Eval-mode BatchNorm: the per-channel scale is folded into the filter (w' = w * gamma/sqrt(var+eps)), the
        shift stays in the epilogue.  The residual of a bottleneck is added inside the tensor-core pipeline BEFORE the
        epilogue, so the epilogue itself must not scale.  A weight_fn may replicate the output channels (the stem's
        4 pixels per super pixel): the BatchNorm vectors then repeat with period bn.num_features."""
        w0 = conv.weight if weight_fn is None else weight_fn(conv.weight.detach())
        co_r, ci_r, kh, kw = w0.shape
        cout = cout_pad or round_up(co_r, 32 if nchw_out else 64)
        cin = cin_pad or round_up(ci_r, 16)
        planes = 2 if self.mode == ops.UP_SPLIT else 1
        dt = torch.float16 if self.mode == ops.UP_FP16 else torch.bfloat16
        if wbuf is None:
            wbuf = torch.empty((planes, kh * kw, cout, cin), dtype=dt, device=self.device)
        else:
            assert wbuf.dtype == dt and wbuf.numel() == planes * kh * kw * cout * cin and wbuf.is_contiguous()
        if scale is None:
            scale = torch.empty(cout, dtype=torch.float32, device=self.device)
        if shift is None:
            shift = torch.empty(cout, dtype=torch.float32, device=self.device)
        assert scale.numel() == cout and shift.numel() == cout and scale.is_contiguous() and shift.is_contiguous()
        wt = self.plan.weights
        fold = torch.empty(bn.num_features, dtype=torch.float32, device=self.device) if bn is not None else None
        pc = PackedConv(wbuf, scale, shift, kh, kw, cout, cin, co_r, ci_r, self.mode, fold)
        bias = (lambda: conv.bias) if (bn is None and conv.bias is not None) else None
        wt.add_epilogue(scale, shift, co_r, bn=bn, bias=bias, fold_scale=fold)
        wt.add_pack(lambda: _param(conv, "weight"), wbuf, cout, cin, row_scale=fold,
                    scale_period=bn.num_features if bn is not None else 1, weight_fn=weight_fn,
                    watch=list(extra_sources))
        return pc

    def packed_transposed(self, conv: nn.Conv2d, bn: Optional[nn.BatchNorm2d] = None, *, ci_off: int = 0,
                          cin_slice: Optional[int] = None, weight_fn=None, fold: Optional[torch.Tensor] = None,
                          fold_period: int = 1, extra_sources: Sequence[torch.Tensor] = ()):
        """TRANSPOSED packing [planes, taps, cin_slice_pad, cout_pad] of (a slice of) `conv`'s filter, optionally scaled
        per output channel by an eval BatchNorm (`bn`: its own fold, or `fold`: one computed elsewhere in the plan).
        Returns (packed, shift) - shift is the fp32 [cout_pad] BatchNorm shift (zeros without `bn`)."""
        w0 = conv.weight if weight_fn is None else weight_fn(conv.weight.detach())
        co_r, ci_t, kh, kw = w0.shape
        ci_r = (ci_t - ci_off) if cin_slice is None else cin_slice
        rows, cols = round_up(ci_r, 16), round_up(co_r, 64)
        planes = 2 if self.mode == ops.UP_SPLIT else 1
        dt = torch.float16 if self.mode == ops.UP_FP16 else torch.bfloat16
        wbuf = torch.empty((planes, kh * kw, rows, cols), dtype=dt, device=self.device)
        scale = torch.empty(cols, dtype=torch.float32, device=self.device)
        shift = torch.empty(cols, dtype=torch.float32, device=self.device)
        wt = self.plan.weights
        period = fold_period
        if bn is not None:
            fold = torch.empty(bn.num_features, dtype=torch.float32, device=self.device)
            period = bn.num_features
        wt.add_epilogue(scale, shift, co_r, bn=bn, fold_scale=fold if bn is not None else None)
        wt.add_pack(lambda: _param(conv, "weight"), wbuf, rows, cols, transpose=True, ci_off=ci_off, cin_slice=ci_r, row_scale=fold,
                    scale_period=period, weight_fn=weight_fn, watch=list(extra_sources))
        return wbuf, shift

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
        self.weights = WeightTable(self.device, self.mode)
        self.inputs: List[torch.Tensor] = []
        self.outputs: List[torch.Tensor] = []
        self.builder = Builder(self)
        if use_graph is None:
            use_graph = os.environ.get("UNIPOSE_B200_GRAPH", "1") != "0"
        self.use_graph = use_graph
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches = 0

    def static_input(self, shape, dtype=torch.float32) -> torch.Tensor:
        t = torch.zeros(shape, dtype=dtype, device=self.device)
        self.inputs.append(t)
        return t

    def finalize(self, outputs: Sequence[torch.Tensor]) -> None:
        self.outputs = list(outputs)

    def refresh_weights(self) -> bool:
        changed = table_changed = self.weights.refresh()
        for j in self.pack_jobs:
            changed |= j.refresh(table_changed)
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
        # every C-ABI launch goes to the current stream of the CURRENT device: make that the plan's device
        with torch.cuda.device(self.device):
            return self._run(*inputs)

    def _run(self, *inputs: torch.Tensor) -> List[torch.Tensor]:
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
