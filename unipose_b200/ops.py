"""Tensor-level wrappers over the C-ABI: torch owns device memory and the stream, nothing else.

`Act` is an NHWC 16-bit activation buffer ([planes, N, H, W, C]; planes == 2 in the fp32-grade
"split" mode).  Every function launches on torch's current CUDA stream and returns immediately.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import (UP_BF16, UP_FLAG_OUT_NCHW_F32, UP_FLAG_PROJ, UP_FLAG_RELU, UP_FLAG_RESIDUAL, UP_FP16, UP_SPLIT,
                   UpConvDesc)

PRECISIONS = {"bf16": UP_BF16, "fp16": UP_FP16, "fp32": UP_SPLIT}


def mode_of(precision: str) -> int:
    try:
        return PRECISIONS[precision]
    except KeyError:
        raise ValueError("precision must be one of %s (got %r)" % (sorted(PRECISIONS), precision))


def _torch_dtype(mode: int) -> torch.dtype:
    return torch.float16 if mode == UP_FP16 else torch.bfloat16


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError("unipose_b200: %s must live on a CUDA device (there is no CPU path)" % what)


class Act:
    """NHWC activation buffer; `.t` has shape [planes, n, h, w, c]."""

    def __init__(self, n: int, h: int, w: int, c: int, mode: int, device, zero: bool = False):
        assert c % 8 == 0
        self.n, self.h, self.w, self.c, self.mode = n, h, w, c, mode
        planes = 2 if mode == UP_SPLIT else 1
        alloc = torch.zeros if zero else torch.empty
        self.t = alloc((planes, n, h, w, c), dtype=_torch_dtype(mode), device=device)

    @property
    def plane_stride(self) -> int:
        return self.n * self.h * self.w * self.c

    def ptr(self, n_off: int = 0) -> ctypes.c_void_p:
        return ctypes.c_void_p(self.t.data_ptr() + 2 * n_off * self.h * self.w * self.c)

    def reshaped(self, h: int, w: int, c: int) -> "Act":
        """Alias of the same storage with another (h, w, c) factorisation of an image's elements."""
        assert h * w * c == self.h * self.w * self.c and c % 8 == 0
        a = object.__new__(Act)
        a.n, a.h, a.w, a.c, a.mode = self.n, h, w, c, self.mode
        a.t = self.t.view(self.t.shape[0], self.n, h, w, c)
        return a

    def to_float(self) -> torch.Tensor:
        """[n, h, w, c] fp32 (hi + lo in split mode) - for tests."""
        f = self.t.float()
        return f.sum(0) if self.mode == UP_SPLIT else f[0]


@dataclass
class View:
    """Channel slice [coff, coff + c) of images [n_off, n_off + n) of an Act."""
    act: Act
    coff: int = 0
    c: int = -1
    n_off: int = 0
    n: int = -1

    def __post_init__(self):
        if self.c < 0:
            self.c = self.act.c - self.coff
        if self.n < 0:
            self.n = self.act.n - self.n_off

    @property
    def h(self):
        return self.act.h

    @property
    def w(self):
        return self.act.w

    def ptr(self):
        return self.act.ptr(self.n_off)


def as_view(a) -> View:
    return a if isinstance(a, View) else View(a)


@dataclass
class PackedConv:
    """Pre-packed conv weights [planes, kh*kw, cout, cin] + fp32 per-channel scale / shift."""
    w: torch.Tensor
    scale: torch.Tensor
    shift: torch.Tensor
    kh: int
    kw: int
    cout: int
    cin: int
    cout_real: int
    cin_real: int
    mode: int
    fold: Optional[torch.Tensor] = None    # eval BatchNorm scale folded into the filter rows (fp32 [bn.num_features])

    @property
    def plane_stride(self) -> int:
        return self.kh * self.kw * self.cout * self.cin


def pack_conv_weight(w_oihw: torch.Tensor, mode: int, cout: Optional[int] = None, cin: Optional[int] = None):
    """OIHW fp32 -> packed 16-bit [planes, taps, cout, cin] (zero padded) via up_pack_conv_weight."""
    require_cuda(w_oihw, "weight")
    w_oihw = w_oihw.detach().contiguous().float()
    co_r, ci_r, kh, kw = w_oihw.shape
    cout = cout or round_up(co_r, 32)
    cin = cin or round_up(ci_r, 16)
    planes = 2 if mode == UP_SPLIT else 1
    out = torch.empty((planes, kh * kw, cout, cin), dtype=_torch_dtype(mode), device=w_oihw.device)
    _lib.call("up_pack_conv_weight", _ptr(w_oihw), _ptr(out), co_r, ci_r, kh, kw, cout, cin, mode,
              kh * kw * cout * cin, _stream())
    return out, (kh, kw, cout, cin, co_r, ci_r)


def bn_fold(gamma, beta, mean, var, eps: float, c: int):
    c_real = gamma.numel()
    scale = torch.empty(c, dtype=torch.float32, device=gamma.device)
    shift = torch.empty(c, dtype=torch.float32, device=gamma.device)
    _lib.call("up_bn_fold", _ptr(gamma.detach().float().contiguous()), _ptr(beta.detach().float().contiguous()),
              _ptr(mean.float().contiguous()), _ptr(var.float().contiguous()), float(eps), _ptr(scale), _ptr(shift),
              c_real, c, _stream())
    return scale, shift


def make_packed_conv(w_oihw, mode, scale=None, shift=None, bias=None, cout=None, cin=None) -> PackedConv:
    """Pack weights and build the epilogue constants: (scale, shift) given, or scale=1 / shift=bias(0)."""
    w, (kh, kw, co, ci, co_r, ci_r) = pack_conv_weight(w_oihw, mode, cout, cin)
    dev = w.device
    if scale is None:
        scale = torch.zeros(co, dtype=torch.float32, device=dev)
        scale[:co_r] = 1.0
        shift = torch.zeros(co, dtype=torch.float32, device=dev)
        if bias is not None:
            shift[:co_r] = bias.detach().float()
    return PackedConv(w, scale.contiguous(), shift.contiguous(), kh, kw, co, ci, co_r, ci_r, mode)


def conv2d(x, pc: PackedConv, y, *, stride: int = 1, dil: int = 1, pad=None, relu: bool = False,
           residual=None, ho: Optional[int] = None, wo: Optional[int] = None, x_groups: int = 1,
           x_group_nstride: int = 0, cout_valid: Optional[int] = None, out_c_total: Optional[int] = None,
           x_window=None, proj=None) -> None:
    """y = epilogue(conv(x, w)).  `y` is an Act/View (NHWC 16-bit) or an fp32 NCHW tensor [n, cout_valid, ho, wo].
    proj = (x2, stride): UP_FLAG_PROJ - the 1x1 / `stride` projection of the second input x2 (all of its view's
    channels) extends K; `pc.w` then holds 1 + x2.c / pc.cin filter slices (see include/unipose_b200.h)."""
    xv = as_view(x)
    if pad is None:
        pad = (dil * (pc.kh - 1) // 2, dil * (pc.kw - 1) // 2)
    if isinstance(pad, int):
        pad = (pad, pad)
    d = UpConvDesc()
    d.n, d.h, d.w = xv.n, xv.h, xv.w
    if x_window is not None:
        # overlapping channel windows over a row-padded buffer: (logical width, channels per window)
        d.w, d.x_cextent = x_window
        d.x_wpitch = xv.w
    if ho is None:
        ho = (d.h + 2 * pad[0] - dil * (pc.kh - 1) - 1) // stride + 1
        wo = (d.w + 2 * pad[1] - dil * (pc.kw - 1) - 1) // stride + 1
    d.ho, d.wo = ho, wo
    d.cin, d.cout = pc.cin, pc.cout
    d.kh, d.kw, d.stride, d.dil = pc.kh, pc.kw, stride, dil
    d.pad_h, d.pad_w = pad
    d.x_cstride, d.x_coff = xv.act.c, xv.coff
    d.x_groups, d.x_group_nstride = x_groups, x_group_nstride
    d.dtype = pc.mode
    d.x_plane_stride = xv.act.plane_stride
    d.w_plane_stride = pc.plane_stride
    flags = UP_FLAG_RELU if relu else 0
    rptr = ctypes.c_void_p(0)
    if residual is not None:
        rv = as_view(residual)
        flags |= UP_FLAG_RESIDUAL
        d.r_cstride, d.r_coff, d.r_plane_stride = rv.act.c, rv.coff, rv.act.plane_stride
        rptr = rv.ptr()
    if proj is not None:
        assert residual is None, "UP_FLAG_PROJ excludes UP_FLAG_RESIDUAL"
        pv, pstride = as_view(proj[0]), int(proj[1])
        assert (pv.n, pv.h, pv.w) == (xv.n, ho * pstride, wo * pstride), ((pv.n, pv.h, pv.w), (xv.n, ho, wo, pstride))
        assert pc.w.numel() == (1 + pv.c // pc.cin) * pc.cout * pc.cin and pv.c % pc.cin == 0
        flags |= UP_FLAG_PROJ
        d.r_cstride, d.r_coff, d.proj_cin, d.proj_stride = pv.act.c, pv.coff, pv.c, pstride
        rptr = pv.ptr()
    if isinstance(y, torch.Tensor):
        flags |= UP_FLAG_OUT_NCHW_F32
        d.cout_valid = cout_valid if cout_valid is not None else pc.cout_real
        d.out_c_total = out_c_total or d.cout_valid
        assert y.dtype == torch.float32 and y.is_contiguous() and tuple(y.shape) == (xv.n, d.out_c_total, ho, wo), \
            (tuple(y.shape), (xv.n, d.out_c_total, ho, wo))
        yptr = _ptr(y)
    else:
        yv = as_view(y)
        assert (yv.n, yv.h, yv.w) == (xv.n, ho, wo) and yv.c == pc.cout, ((yv.n, yv.h, yv.w, yv.c), (xv.n, ho, wo, pc.cout))
        d.y_cstride, d.y_coff, d.y_plane_stride = yv.act.c, yv.coff, yv.act.plane_stride
        yptr = yv.ptr()
    d.flags = flags
    _lib.call("up_conv2d_fwd", ctypes.byref(d), xv.ptr(), _ptr(pc.w), _ptr(pc.scale), _ptr(pc.shift), rptr, yptr,
              ctypes.c_void_p(0), _stream())


# ---------------------------------------------------------------------------------------------
# bandwidth kernels
# ---------------------------------------------------------------------------------------------
def pack_input_s2d(x_nchw: torch.Tensor, y: Act, wpad_left: int = 0) -> None:
    """fp32 NCHW image -> 2x2 space-to-depth NHWC16; `y` may have padded rows (row pitch >= w/2 + wpad_left s2d
    pixels) and may group k s2d pixels per buffer pixel (y.c = 16*k, e.g. the stem's 64-element super pixels)."""
    n, c, h, w = x_nchw.shape
    assert c == 3 and (y.n, y.h) == (n, h // 2) and y.c % 16 == 0
    wpitch = y.w * (y.c // 16)
    assert wpitch >= w // 2 + wpad_left
    _lib.call("up_pack_input_s2d", _ptr(x_nchw), y.ptr(), n, h, w, y.mode, y.plane_stride, wpitch, wpad_left, _stream())


def pack_input_u8_s2d(x_nhwc_u8: torch.Tensor, y: Act, wpad_left: int = 0, mean: float = 128.0, std: float = 256.0) -> None:
    """uint8 HWC images [n, h, w, 3] -> normalised 2x2 space-to-depth NHWC16 (same layout rules as pack_input_s2d)."""
    n, h, w, c = x_nhwc_u8.shape
    assert c == 3 and x_nhwc_u8.dtype == torch.uint8 and x_nhwc_u8.is_contiguous()
    assert (y.n, y.h) == (n, h // 2) and y.c % 16 == 0
    wpitch = y.w * (y.c // 16)
    assert wpitch >= w // 2 + wpad_left
    _lib.call("up_pack_input_u8_s2d", _ptr(x_nhwc_u8), y.ptr(), n, h, w, y.mode, y.plane_stride, wpitch, wpad_left,
              float(mean), float(std), _stream())


def nchw_to_act(x: torch.Tensor, y, c_real: Optional[int] = None) -> None:
    yv = as_view(y)
    n, c, h, w = x.shape
    x = x.contiguous()
    assert (yv.n, yv.h, yv.w) == (n, h, w) and x.dtype == torch.float32
    _lib.call("up_nchw_f32_to_nhwc", _ptr(x), yv.ptr(), n, c, h, w, yv.c, yv.act.c, yv.coff, yv.act.mode,
              yv.act.plane_stride, _stream())


def act_to_nchw(x, c_real: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    xv = as_view(x)
    if out is None:
        out = torch.empty((xv.n, c_real, xv.h, xv.w), dtype=torch.float32, device=xv.act.t.device)
    _lib.call("up_nhwc_to_nchw_f32", xv.ptr(), _ptr(out), xv.n, c_real, xv.h, xv.w, xv.act.c, xv.coff, xv.act.mode,
              xv.act.plane_stride, _stream())
    return out


def maxpool3x3s2(x, y) -> None:
    xv, yv = as_view(x), as_view(y)
    assert xv.c == yv.c and xv.n == yv.n
    _lib.call("up_maxpool3x3s2", xv.ptr(), yv.ptr(), xv.n, xv.h, xv.w, xv.c, xv.act.c, xv.coff, yv.act.c, yv.coff,
              xv.act.mode, xv.act.plane_stride, yv.act.plane_stride, _stream())


def upsample_bilinear_ac(x, y) -> None:
    xv, yv = as_view(x), as_view(y)
    assert xv.c == yv.c and xv.n == yv.n
    _lib.call("up_upsample_bilinear_ac", xv.ptr(), yv.ptr(), xv.n, xv.h, xv.w, yv.h, yv.w, xv.c, xv.act.c, xv.coff,
              yv.act.c, yv.coff, xv.act.mode, xv.act.plane_stride, yv.act.plane_stride, _stream())


def global_avgpool(x, y) -> None:
    xv, yv = as_view(x), as_view(y)
    assert xv.c == yv.c and xv.n == yv.n and yv.h == 1 and yv.w == 1
    _lib.call("up_global_avgpool", xv.ptr(), yv.ptr(), xv.n, xv.h, xv.w, xv.c, xv.act.c, xv.coff, yv.act.c, yv.coff,
              xv.act.mode, xv.act.plane_stride, yv.act.plane_stride, _stream())


def global_sumpool(x, y) -> None:
    xv, yv = as_view(x), as_view(y)
    assert xv.c == yv.c and xv.n == yv.n and yv.h == 1 and yv.w == 1
    _lib.call("up_global_sumpool", xv.ptr(), yv.ptr(), xv.n, xv.h, xv.w, xv.c, xv.act.c, xv.coff, yv.act.c, yv.coff,
              xv.act.mode, xv.act.plane_stride, yv.act.plane_stride, _stream())


def broadcast_hw(x, y) -> None:
    xv, yv = as_view(x), as_view(y)
    assert xv.c == yv.c and xv.n == yv.n and xv.h == 1 and xv.w == 1
    _lib.call("up_broadcast_hw", xv.ptr(), yv.ptr(), xv.n, yv.h, yv.w, xv.c, xv.act.c, xv.coff, yv.act.c, yv.coff,
              xv.act.mode, xv.act.plane_stride, yv.act.plane_stride, _stream())


def upsample_bilinear_ac_nchw(x: torch.Tensor, size) -> torch.Tensor:
    n, c, h, w = x.shape
    out = torch.empty((n, c, size[0], size[1]), dtype=torch.float32, device=x.device)
    _lib.call("up_upsample_bilinear_ac_nchw_f32", _ptr(x.contiguous()), _ptr(out), n, c, h, w, size[0], size[1],
              _stream())
    return out


# ---------------------------------------------------------------------------------------------
# training kernels
# ---------------------------------------------------------------------------------------------
def _uview(a) -> "_lib.UpView":
    v = as_view(a)
    return _lib.UpView(v.ptr(), v.act.c, v.coff, v.act.plane_stride)


def _vref(a):
    return ctypes.byref(_uview(a)) if a is not None else None


def conv_desc(x, pc: PackedConv, ho: int, wo: int, *, stride=1, dil=1, pad=(0, 0), x_groups=1, x_group_nstride=0,
              x_window=None) -> UpConvDesc:
    """Forward descriptor of a layer (geometry only) - what up_conv2d_wgrad takes."""
    xv = as_view(x)
    d = UpConvDesc()
    d.n, d.h, d.w = xv.n, xv.h, xv.w
    if x_window is not None:
        d.w, d.x_cextent = x_window
        d.x_wpitch = xv.w
    d.ho, d.wo = ho, wo
    d.cin, d.cout = pc.cin, pc.cout
    d.kh, d.kw, d.stride, d.dil = pc.kh, pc.kw, stride, dil
    d.pad_h, d.pad_w = pad if not isinstance(pad, int) else (pad, pad)
    d.x_cstride, d.x_coff = xv.act.c, xv.coff
    d.x_groups, d.x_group_nstride = x_groups, x_group_nstride
    d.dtype = pc.mode
    d.x_plane_stride = xv.act.plane_stride
    return d


def wgrad_scratch_bytes(d: UpConvDesc) -> int:
    return int(_lib.load().up_conv2d_wgrad_scratch_bytes(ctypes.byref(d)))


def conv2d_wgrad(d: UpConvDesc, x, dz, dw: torch.Tensor, scratch: torch.Tensor, accumulate: bool = False) -> None:
    """dw[cout_real, cin_real, kh, kw] (+)= wgrad(x, dz); dz is a dense [n, ho, wo, cout_pad] Act (or an image
    sub-range of one)."""
    xv = as_view(x)
    dz = as_view(dz)
    assert dz.coff == 0 and dz.c == dz.act.c == d.cout and (dz.n, dz.h, dz.w) == (d.n, d.ho, d.wo), \
        ((dz.n, dz.h, dz.w, dz.c), (d.n, d.ho, d.wo, d.cout))
    assert dw.dtype == torch.float32 and dw.is_contiguous()
    d.y_cstride, d.y_coff, d.y_plane_stride = dz.c, 0, dz.act.plane_stride
    _lib.call("up_conv2d_wgrad", ctypes.byref(d), xv.ptr(), dz.ptr(), _ptr(dw), dw.shape[0], dw.shape[1],
              _ptr(scratch), scratch.numel() * scratch.element_size(), 1 if accumulate else 0, _stream())


def bn_work_doubles(c: int) -> int:
    """Size (in doubles) of the work buffer every BatchNorm reduction of `c` channels needs (sums, scratch, rows)."""
    return int(_lib.load().up_bn_work_doubles(int(c)))


def bn_stats(z, c: int, sums: torch.Tensor) -> None:
    zv = as_view(z)
    assert sums.numel() >= bn_work_doubles(c), "bn_stats: work buffer smaller than ops.bn_work_doubles(c)"
    _lib.call("up_bn_stats", _vref(zv), zv.n * zv.h * zv.w, c, zv.act.mode, _ptr(sums), _stream())


def bn_finalize(sums, count, bn, scale, shift, save_mean, save_invstd, c_real: int, c: int, update_running=True):
    rm = bn.running_mean if (update_running and bn.running_mean is not None) else None
    rv = bn.running_var if (update_running and bn.running_var is not None) else None
    mom = 0.1 if bn.momentum is None else float(bn.momentum)
    _lib.call("up_bn_finalize", _ptr(sums), count, _ptr(bn.weight.detach()), _ptr(bn.bias.detach()), _ptr(rm), _ptr(rv),
              mom, float(bn.eps), _ptr(scale), _ptr(shift), _ptr(save_mean), _ptr(save_invstd), c_real, c, _stream())


def bn_stats_finalize(z, sums, bn, scale, shift, save_mean, save_invstd, c_real: int, c: int, update_running=True):
    """bn_stats + bn_finalize over the pixels of z (two launches instead of three, same numbers)."""
    zv = as_view(z)
    assert sums.numel() >= bn_work_doubles(c), "bn_stats_finalize: work buffer smaller than ops.bn_work_doubles(c)"
    rm = bn.running_mean if (update_running and bn.running_mean is not None) else None
    rv = bn.running_var if (update_running and bn.running_var is not None) else None
    mom = 0.1 if bn.momentum is None else float(bn.momentum)
    _lib.call("up_bn_stats_finalize", _vref(zv), zv.n * zv.h * zv.w, c, zv.act.mode, _ptr(sums),
              _ptr(bn.weight.detach()), _ptr(bn.bias.detach()), _ptr(rm), _ptr(rv), mom, float(bn.eps), _ptr(scale),
              _ptr(shift), _ptr(save_mean), _ptr(save_invstd), c_real, _stream())


def scale_shift_act(z, y, scale, shift, *, relu: bool, residual=None, mask=None) -> None:
    zv = as_view(z)
    _lib.call("up_scale_shift_act", _vref(zv), _vref(y), _vref(residual), _vref(mask), _ptr(scale), _ptr(shift),
              zv.n * zv.h * zv.w, zv.c, 1 if relu else 0, zv.act.mode, _stream())


def bn_eval_prepare(bn, scale, shift, save_mean, save_invstd, c_real: int, c: int) -> None:
    """Frozen (eval-mode) BatchNorm of a training step: epilogue constants + backward statistics from running stats."""
    _lib.call("up_bn_eval_prepare", _ptr(bn.weight.detach()), _ptr(bn.bias.detach()), _ptr(bn.running_mean),
              _ptr(bn.running_var), float(bn.eps), _ptr(scale), _ptr(shift), _ptr(save_mean), _ptr(save_invstd), c_real, c,
              _stream())


def bn_bwd(dy, y, z, dz, dres, save_mean, save_invstd, gamma, sums, c_real: int, relu: bool, dgamma, dbeta,
           frozen: bool = False) -> None:
    dv = as_view(dy)
    npix = dv.n * dv.h * dv.w
    assert sums.numel() >= bn_work_doubles(dv.c), "bn_bwd: work buffer smaller than ops.bn_work_doubles(c)"
    _lib.call("up_bn_bwd", _vref(dv), _vref(y) if relu else None, _vref(z), _vref(dz), _vref(dres),
              _ptr(save_mean), _ptr(save_invstd), _ptr(gamma), _ptr(sums), npix, c_real, dv.c,
              (1 if relu else 0) | (2 if frozen else 0), dv.act.mode, _ptr(dgamma), _ptr(dbeta), _stream())


def ew(a, out, *, m=None, op: int = 0, accumulate: bool = False) -> None:
    """out (+)= a | a*m | a*(m>0)   (op 0 / 1 / 2)."""
    av = as_view(a)
    _lib.call("up_ew_mul", _vref(av), _vref(m), _vref(out), av.n * av.h * av.w, av.c, op, 1 if accumulate else 0,
              av.act.mode, _stream())


def maxpool3x3s2_bwd(x, dy, dx, accumulate: bool = False, idx: Optional[torch.Tensor] = None) -> None:
    """`idx`: optional uint8 scratch of n*ho*wo*c elements -> two-pass (arg-max map + gather) form."""
    xv = as_view(x)
    if idx is not None:
        ho, wo = (xv.h - 1) // 2 + 1, (xv.w - 1) // 2 + 1
        assert idx.dtype == torch.uint8 and idx.numel() >= xv.n * ho * wo * xv.c
    _lib.call("up_maxpool3x3s2_bwd", _vref(xv), _vref(dy), _vref(dx), xv.n, xv.h, xv.w, xv.c, 1 if accumulate else 0,
              xv.act.mode, _ptr(idx), _stream())


def upsample_bilinear_ac_bwd(dy, dx, accumulate: bool = False) -> None:
    dyv, dxv = as_view(dy), as_view(dx)
    _lib.call("up_upsample_bilinear_ac_bwd", _vref(dyv), _vref(dxv), dxv.n, dxv.h, dxv.w, dyv.h, dyv.w, dxv.c,
              1 if accumulate else 0, dxv.act.mode, _stream())


def add_broadcast(g, dx, mult: float, accumulate: bool) -> None:
    dxv = as_view(dx)
    _lib.call("up_add_broadcast", _vref(g), _vref(dxv), dxv.n, dxv.h * dxv.w, dxv.c, float(mult),
              1 if accumulate else 0, dxv.act.mode, _stream())


def zero_insert2x(x, y) -> None:
    xv = as_view(x)
    _lib.call("up_zero_insert2x", _vref(xv), _vref(y), xv.n, xv.h, xv.w, xv.c, xv.act.mode, _stream())
