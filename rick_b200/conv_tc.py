"""Python binding of ``rick_conv_tc`` -- the tcgen05 / TMEM / TMA implicit-GEMM convolution (rick_b200/csrc/conv_tc.cu).

Tensors at this level are physical NHWC: activations ``(B, H, W, C)`` contiguous fp32, weights packed per filter tap as
``(k*k, Cout, Cin)``.  ``rick_b200.conv`` adapts the module layer (logical NCHW, channels-last memory) to it.
"""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int, c_void_p
from typing import Optional

import torch

from . import _lib


class ConvPhase(ctypes.Structure):
    _fields_ = [("n_taps", c_int), ("dy", c_int * 9), ("dx", c_int * 9), ("widx", c_int * 9), ("out_y0", c_int),
                ("out_x0", c_int), ("rows", c_int), ("cols", c_int)]


class ConvGeom(ctypes.Structure):
    _fields_ = [("batch", c_int), ("in_h", c_int), ("in_w", c_int), ("cin", c_int), ("cout", c_int), ("out_h", c_int),
                ("out_w", c_int), ("in_stride", c_int), ("out_stride", c_int), ("n_weight_taps", c_int),
                ("n_phases", c_int), ("phase", ConvPhase * 4)]


class ConvEpilogue(ctypes.Structure):
    _fields_ = [("demod", c_void_p), ("noise", c_void_p), ("noise_weight", c_void_p), ("bias", c_void_p),
                ("s_next", c_void_p), ("out2", c_void_p), ("act", c_int), ("alpha", c_float), ("scale", c_float)]


_lib._OPTIONAL["rick_conv_tc"] = (c_int, [c_void_p, c_void_p, c_void_p, ctypes.POINTER(ConvGeom),
                                          ctypes.POINTER(ConvEpilogue), c_void_p])


_lib._OPTIONAL["rick_blur_nhwc"] = (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                            ctypes.POINTER(ConvEpilogue), c_void_p])
_lib._OPTIONAL["rick_to_rgb_nhwc"] = (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                              c_int, c_void_p])


def supported(cin: int, cout: int) -> bool:
    return cin % 32 == 0 and cout % 128 == 0


def pack_weight(w: torch.Tensor) -> torch.Tensor:
    """(Cout, Cin, kh, kw) -> (kh*kw, Cout, Cin): one K-major GEMM operand per filter tap."""
    co, ci, kh, kw = w.shape
    return w.permute(2, 3, 0, 1).reshape(kh * kw, co, ci).contiguous()


def geom_conv(batch, h, w, cin, cout, k, stride=1, pad=0) -> ConvGeom:
    """Cross-correlation as F.conv2d: out[y, x] = sum w[ky, kx] * in[y*stride + ky - pad, x*stride + kx - pad]."""
    g = ConvGeom()
    oh, ow = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    g.batch, g.in_h, g.in_w, g.cin, g.cout, g.out_h, g.out_w = batch, h, w, cin, cout, oh, ow
    g.in_stride, g.out_stride, g.n_weight_taps, g.n_phases = stride, 1, k * k, 1
    ph = g.phase[0]
    ph.n_taps = k * k
    for ky in range(k):
        for kx in range(k):
            t = ky * k + kx
            ph.dy[t], ph.dx[t], ph.widx[t] = ky - pad, kx - pad, t
    ph.out_y0 = ph.out_x0 = 0
    ph.rows, ph.cols = oh, ow
    return g


def geom_conv_transpose_s2(batch, h, w, cin, cout, k=3) -> ConvGeom:
    """F.conv_transpose2d(stride=2, padding=0) as 4 polyphase sub-convolutions: output (2h+1, 2w+1) for k = 3.
    out[2m+a, 2n+b] = sum_{ty, tx} W[a+2ty, b+2tx] * in[m-ty, n-tx]."""
    assert k == 3
    g = ConvGeom()
    oh, ow = (h - 1) * 2 + k, (w - 1) * 2 + k
    g.batch, g.in_h, g.in_w, g.cin, g.cout, g.out_h, g.out_w = batch, h, w, cin, cout, oh, ow
    g.in_stride, g.out_stride, g.n_weight_taps, g.n_phases = 1, 2, k * k, 4
    i = 0
    for a in (0, 1):
        for b in (0, 1):
            ph = g.phase[i]
            i += 1
            t = 0
            for ty in ((0, 1) if a == 0 else (0,)):
                for tx in ((0, 1) if b == 0 else (0,)):
                    ph.dy[t], ph.dx[t], ph.widx[t] = -ty, -tx, (a + 2 * ty) * k + (b + 2 * tx)
                    t += 1
            ph.n_taps = t
            ph.out_y0, ph.out_x0 = a, b
            ph.rows, ph.cols = (h + 1 if a == 0 else h), (w + 1 if b == 0 else w)
    return g


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def conv_tc_nhwc(xm: torch.Tensor, wt: torch.Tensor, geom: ConvGeom, demod=None, noise=None, noise_weight=None,
                 bias=None, act: bool = False, alpha: float = 0.2, scale: float = 2 ** 0.5, s_next=None,
                 want_out2: bool = False):
    """Launch the kernel.  Returns ``out`` (B, OH, OW, Cout) or ``(out, out2)`` when ``want_out2``."""
    for name, t in (("xm", xm), ("wt", wt), ("demod", demod), ("noise", noise), ("noise_weight", noise_weight),
                    ("bias", bias), ("s_next", s_next)):
        if t is not None and (not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous()):
            raise RuntimeError(f"rick_b200.conv_tc: {name} must be a contiguous float32 CUDA tensor")
    if tuple(xm.shape) != (geom.batch, geom.in_h, geom.in_w, geom.cin):
        raise RuntimeError(f"rick_b200.conv_tc: xm shape {tuple(xm.shape)} does not match the geometry")
    if tuple(wt.shape) != (geom.n_weight_taps, geom.cout, geom.cin):
        raise RuntimeError(f"rick_b200.conv_tc: wt shape {tuple(wt.shape)} does not match the geometry")
    out = torch.empty((geom.batch, geom.out_h, geom.out_w, geom.cout), dtype=torch.float32, device=xm.device)
    out2 = torch.empty_like(out) if want_out2 else None
    ep = ConvEpilogue(_ptr(demod), _ptr(noise), _ptr(noise_weight), _ptr(bias), _ptr(s_next), _ptr(out2), int(act),
                      float(alpha), float(scale))
    with torch.cuda.device(xm.device):
        st = _lib.lib().rick_conv_tc(out.data_ptr(), xm.data_ptr(), wt.data_ptr(), ctypes.byref(geom),
                                     ctypes.byref(ep), torch.cuda.current_stream().cuda_stream)
    _lib.check(st, "rick_conv_tc")
    return (out, out2) if want_out2 else out


def blur_nhwc(x: torch.Tensor, taps: torch.Tensor, pad, demod=None, noise=None, noise_weight=None, bias=None,
              act: bool = False, alpha: float = 0.2, scale: float = 2 ** 0.5, s_next=None, want_out2: bool = False,
              flip: bool = False):
    """4x4 FIR over NHWC ``x`` with the StyledConv epilogue fused (see rick_blur_nhwc in include/rick_b200.h)."""
    b, h, w, c = x.shape
    oh, ow = h + pad[0] + pad[1] - 3, w + pad[0] + pad[1] - 3
    out = torch.empty((b, oh, ow, c), dtype=torch.float32, device=x.device)
    out2 = torch.empty_like(out) if want_out2 else None
    ep = ConvEpilogue(_ptr(demod), _ptr(noise), _ptr(noise_weight), _ptr(bias), _ptr(s_next), _ptr(out2), int(act),
                      float(alpha), float(scale))
    with torch.cuda.device(x.device):
        st = _lib.lib().rick_blur_nhwc(out.data_ptr(), x.data_ptr(), taps.data_ptr(), b, h, w, c, pad[0], pad[1],
                                       int(flip), ctypes.byref(ep), torch.cuda.current_stream().cuda_stream)
    _lib.check(st, "rick_blur_nhwc")
    return (out, out2) if want_out2 else out


def to_rgb_nhwc(y: torch.Tensor, wmod: torch.Tensor, bias: torch.Tensor, skip: Optional[torch.Tensor]):
    """(B, H, W, C) activation -> (B, 3, H, W) image: sum_c y * wmod[b, o, c] + bias[o] (+ skip)."""
    b, h, w, c = y.shape
    rgb = torch.empty((b, 3, h, w), dtype=torch.float32, device=y.device)
    with torch.cuda.device(y.device):
        st = _lib.lib().rick_to_rgb_nhwc(rgb.data_ptr(), y.data_ptr(), wmod.data_ptr(), bias.data_ptr(), _ptr(skip), b,
                                         h, w, c, torch.cuda.current_stream().cuda_stream)
    _lib.check(st, "rick_to_rgb_nhwc")
    return rgb
