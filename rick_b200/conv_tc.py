"""Python binding of ``rick_conv_tc`` -- the tcgen05 / TMEM / TMA implicit-GEMM convolution (rick_b200/csrc/conv_tc.cu).

Tensors at this level are physical NHWC: activations ``(B, H, W, C)`` contiguous fp32, weights packed per filter tap as
``(k*k, Cout, Cin)``.  ``rick_b200.conv`` adapts the module layer (logical NCHW, channels-last memory) to it.
"""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int, c_int64, c_void_p
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


class ConvWeight(ctypes.Structure):
    _fields_ = [("ptr", c_void_p), ("stride_m", c_int64), ("stride_k", c_int64), ("stride_tap", c_int64)]


class WgradGeom(ctypes.Structure):
    _fields_ = [("batch", c_int), ("g_h", c_int), ("g_w", c_int), ("cout", c_int), ("x_h", c_int), ("x_w", c_int),
                ("cin", c_int), ("n_taps", c_int), ("gy", c_int * 9), ("gx", c_int * 9), ("xy", c_int * 9),
                ("xx", c_int * 9), ("g_stride", c_int), ("x_stride", c_int), ("rows", c_int), ("cols", c_int)]


_lib._OPTIONAL["rick_conv_tc"] = (c_int, [c_void_p, c_void_p, c_void_p, ctypes.POINTER(ConvGeom),
                                          ctypes.POINTER(ConvEpilogue), c_void_p])
_lib._OPTIONAL["rick_conv_tc_w"] = (c_int, [c_void_p, c_void_p, ctypes.POINTER(ConvWeight), ctypes.POINTER(ConvGeom),
                                            ctypes.POINTER(ConvEpilogue), c_void_p, c_int64, c_void_p])
_lib._OPTIONAL["rick_conv_tc_workspace"] = (c_int64, [ctypes.POINTER(ConvGeom)])
_lib._OPTIONAL["rick_conv_tc_plan"] = (ctypes.c_int, [ctypes.POINTER(ConvGeom)] + [ctypes.POINTER(ctypes.c_int)] * 5)


def launch_plan(geom: "ConvGeom") -> dict:
    """The tile plan ``rick_conv_tc_w`` would use (host logic only; works without a GPU)."""
    tw, th = (ctypes.c_int * 4)(), (ctypes.c_int * 4)()
    nb, ks, tot = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.check(_lib.lib().rick_conv_tc_plan(ctypes.byref(geom), tw, th, ctypes.byref(nb), ctypes.byref(ks), ctypes.byref(tot)),
               "rick_conv_tc_plan")
    return {"tiles": [(tw[i], th[i]) for i in range(geom.n_phases)], "samples_per_tile": nb.value, "ksplit": ks.value,
            "total_tiles": tot.value}
_lib._OPTIONAL["rick_conv_wgrad_workspace"] = (c_int64, [ctypes.POINTER(WgradGeom)])
_lib._OPTIONAL["rick_conv_wgrad_tc"] = (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                                                ctypes.POINTER(WgradGeom), c_void_p, c_float, c_void_p])


_lib._OPTIONAL["rick_blur_nhwc"] = (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                            ctypes.POINTER(ConvEpilogue), c_void_p])
_lib._OPTIONAL["rick_to_rgb_nhwc"] = (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                              c_int, c_void_p])


SPLIT_K = True      # divide the K loop of small-map convolutions over idle SMs (tests switch it off to compare)


def supported(cin: int, cout: int) -> bool:
    """Channel counts rick_conv_tc accepts (GEMM-K = cin in blocks of 32; cout in tiles of 128, a partial tile rides on
    TMA zero fill -- the 64 / 32-channel layers of the 512 / 1024 px generators included)."""
    return cin % 32 == 0 and cout % 32 == 0


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


def geom_conv_dgrad(batch, h, w, cin, cout, k, stride=1, pad=0) -> ConvGeom:
    """Data gradient of ``F.conv2d(x (B,cin,h,w), W (cout,cin,k,k), stride, pad)`` as a convolution over the output
    gradient g (B, oh, ow, cout):  gx[iy, ix, ci] = sum_{ky,kx,co} W[co,ci,ky,kx] * g[(iy+pad-ky)/s, (ix+pad-kx)/s, co].
    In the returned geometry GEMM-M = cin (``geom.cout``), GEMM-K = cout (``geom.cin``); ``widx`` indexes the layer's own
    tap order ky*k + kx, so the forward weight memory is used as is (see ``weight_operand``)."""
    oh, ow = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    g = ConvGeom()
    g.batch, g.in_h, g.in_w, g.cin, g.cout, g.out_h, g.out_w = batch, oh, ow, cout, cin, h, w
    g.n_weight_taps = k * k
    if stride == 1:
        g.in_stride, g.out_stride, g.n_phases = 1, 1, 1
        ph = g.phase[0]
        ph.n_taps = k * k
        for ky in range(k):
            for kx in range(k):
                t = ky * k + kx
                ph.dy[t], ph.dx[t], ph.widx[t] = pad - ky, pad - kx, t
        ph.out_y0 = ph.out_x0 = 0
        ph.rows, ph.cols = h, w
        return g
    assert stride == 2
    # polyphase: input row iy = 2m + a - pad receives taps ky = a + 2 ty from g row m - ty
    g.in_stride, g.out_stride = 1, 2
    i = 0
    for a in range(2):
        for b in range(2):
            kys = [ky for ky in range(k) if ky % 2 == a]
            kxs = [kx for kx in range(k) if kx % 2 == b]
            y0, x0 = a - pad, b - pad                     # first output row / column of this phase (m = 0)
            m0, n0 = (0 if y0 >= 0 else (-y0 + 1) // 2), (0 if x0 >= 0 else (-x0 + 1) // 2)
            rows = len(range(y0 + 2 * m0, h, 2))
            cols = len(range(x0 + 2 * n0, w, 2))
            if rows <= 0 or cols <= 0 or not kys or not kxs:     # nothing lands on these pixels (their gradient is zero)
                continue
            ph = g.phase[i]
            i += 1
            t = 0
            for ky in kys:
                for kx in kxs:
                    ph.dy[t], ph.dx[t], ph.widx[t] = m0 - (ky - a) // 2, n0 - (kx - b) // 2, ky * k + kx
                    t += 1
            ph.n_taps = t
            ph.out_y0, ph.out_x0 = y0 + 2 * m0, x0 + 2 * n0
            ph.rows, ph.cols = rows, cols
    g.n_phases = i
    return g


def geom_wgrad(batch, h, w, cin, cout, k, stride=1, pad=0, transposed=False) -> WgradGeom:
    """Weight gradient of ``F.conv2d(x (B,cin,h,w), W (cout,cin,k,k), stride, pad)`` or, with ``transposed``, of
    ``F.conv_transpose2d(x, W^T, stride=2)`` (x (B,cin,h,w) -> (B,cout,2h+1,2w+1) for k = 3)."""
    g = WgradGeom()
    g.batch, g.cin, g.cout, g.n_taps = batch, cin, cout, k * k
    g.x_h, g.x_w = h, w
    if transposed:
        assert stride == 2 and pad == 0
        g.g_h, g.g_w = (h - 1) * 2 + k, (w - 1) * 2 + k
        g.g_stride, g.x_stride = 2, 1
        g.rows, g.cols = h, w
        for ky in range(k):
            for kx in range(k):
                t = ky * k + kx
                g.gy[t], g.gx[t], g.xy[t], g.xx[t] = ky, kx, 0, 0
        return g
    g.g_h, g.g_w = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    g.g_stride, g.x_stride = 1, stride
    g.rows, g.cols = g.g_h, g.g_w
    for ky in range(k):
        for kx in range(k):
            t = ky * k + kx
            g.gy[t], g.gx[t], g.xy[t], g.xx[t] = 0, 0, ky - pad, kx - pad
    return g


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def weight_operand(w: torch.Tensor, transpose: bool = False) -> ConvWeight:
    """Describe a (Cout, Cin, k, k) weight stored channels-last ([Cout][k][k][Cin] in memory) to the kernel.
    ``transpose=False``: GEMM-M = Cout (forward).  ``transpose=True``: GEMM-M = Cin, GEMM-K = Cout (data gradient) -- the
    same memory read as an MN-major operand."""
    co, ci, kh, kw = w.shape
    sc, si, sy, sx = w.stride()
    if si != 1 or (kw > 1 and sx != ci) or (kh > 1 and sy != kw * ci) or sc < kh * kw * ci:
        raise RuntimeError("rick_b200.conv_tc: weight must be stored channels-last ([Cout][kh][kw][Cin])")
    tap = ci if kh * kw > 1 else 4
    if transpose:
        return ConvWeight(w.data_ptr(), 1, sc, tap)
    return ConvWeight(w.data_ptr(), sc, 1, tap)


def _covers_partially(geom: ConvGeom) -> bool:
    """A polyphase geometry that does not write every output pixel (e.g. the data gradient of a 1x1 stride-2 conv)."""
    covered = sum(geom.phase[i].rows * geom.phase[i].cols for i in range(geom.n_phases))
    return covered < geom.out_h * geom.out_w


_WS_CACHE: dict = {}


def conv_wgrad_tc(g: torch.Tensor, x: torch.Tensor, geom: WgradGeom, like: torch.Tensor, scale: float = 1.0):
    """rick_conv_wgrad_tc: ``g`` (B, g_h, g_w, Cout) and ``x`` (B, x_h, x_w, Cin) physical NHWC; returns dW with the shape
    and strides of ``like`` (the (Cout, Cin, k, k) weight)."""
    for name, t, shp in (("g", g, (geom.batch, geom.g_h, geom.g_w, geom.cout)),
                         ("x", x, (geom.batch, geom.x_h, geom.x_w, geom.cin))):
        if tuple(t.shape) != shp or not t.is_contiguous() or t.dtype != torch.float32 or not t.is_cuda:
            raise RuntimeError(f"rick_b200.conv_tc: {name} {tuple(t.shape)} does not match the geometry / is not NHWC fp32")
    lib = _lib.lib()
    nbytes = int(lib.rick_conv_wgrad_workspace(ctypes.byref(geom)))
    if nbytes < 0:
        raise RuntimeError("rick_b200.conv_tc: unsupported weight-gradient geometry")
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=g.device)
    dw = torch.empty_strided(like.shape, like.stride(), dtype=torch.float32, device=g.device)
    co, ci, kh, kw = like.shape
    sc, si, sy, sx = like.stride()
    if kh * kw > 1 and sy != kw * sx:
        raise RuntimeError("rick_b200.conv_tc: weight taps must be evenly strided")
    with torch.cuda.device(g.device):
        st = lib.rick_conv_wgrad_tc(dw.data_ptr(), sc, si, sx if kh * kw > 1 else 4, g.data_ptr(), x.data_ptr(),
                                    ctypes.byref(geom), ws.data_ptr(), float(scale), torch.cuda.current_stream().cuda_stream)
    _lib.check(st, "rick_conv_wgrad_tc")
    return dw


def conv_tc_nhwc(xm: torch.Tensor, wt: torch.Tensor, geom: ConvGeom, demod=None, noise=None, noise_weight=None,
                 bias=None, act: bool = False, alpha: float = 0.2, scale: float = 2 ** 0.5, s_next=None,
                 want_out2: bool = False, transpose_weight: bool = False):
    """Launch the kernel.  ``xm`` is a physical-NHWC activation (B, H, W, C).  ``wt`` is either the packed
    ``(taps, Cout, Cin)`` form of :func:`pack_weight` or a ``(Cout, Cin, k, k)`` weight stored channels-last, which is
    read in place (``transpose_weight=True``: as the (Cin x Cout) operand of the data gradient).
    Returns ``out`` (B, OH, OW, Cout) or ``(out, out2)`` when ``want_out2``."""
    for name, t in (("xm", xm), ("demod", demod), ("noise", noise), ("noise_weight", noise_weight),
                    ("bias", bias), ("s_next", s_next)):
        if t is not None and (not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous()):
            raise RuntimeError(f"rick_b200.conv_tc: {name} must be a contiguous float32 CUDA tensor")
    if tuple(xm.shape) != (geom.batch, geom.in_h, geom.in_w, geom.cin):
        raise RuntimeError(f"rick_b200.conv_tc: xm shape {tuple(xm.shape)} does not match the geometry")
    if not wt.is_cuda or wt.dtype != torch.float32:
        raise RuntimeError("rick_b200.conv_tc: weight must be a float32 CUDA tensor")
    if wt.dim() == 3:
        if tuple(wt.shape) != (geom.n_weight_taps, geom.cout, geom.cin) or not wt.is_contiguous() or transpose_weight:
            raise RuntimeError(f"rick_b200.conv_tc: packed weight {tuple(wt.shape)} does not match the geometry")
        wd = ConvWeight(wt.data_ptr(), geom.cin, 1, geom.cin * geom.cout)
    else:
        m, k = (wt.shape[1], wt.shape[0]) if transpose_weight else (wt.shape[0], wt.shape[1])
        if (m, k, wt.shape[2] * wt.shape[3]) != (geom.cout, geom.cin, geom.n_weight_taps):
            raise RuntimeError(f"rick_b200.conv_tc: weight {tuple(wt.shape)} does not match the geometry")
        wd = weight_operand(wt, transpose_weight)
    out = torch.empty((geom.batch, geom.out_h, geom.out_w, geom.cout), dtype=torch.float32, device=xm.device)
    if geom.out_stride > 1 and _covers_partially(geom):
        out.zero_()
    out2 = torch.empty_like(out) if want_out2 else None
    ep = ConvEpilogue(_ptr(demod), _ptr(noise), _ptr(noise_weight), _ptr(bias), _ptr(s_next), _ptr(out2), int(act),
                      float(alpha), float(scale))
    lib = _lib.lib()
    ws_bytes = int(lib.rick_conv_tc_workspace(ctypes.byref(geom))) if SPLIT_K else 0
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xm.device) if ws_bytes > 0 else None     # split-K partial sums
    with torch.cuda.device(xm.device):
        st = lib.rick_conv_tc_w(out.data_ptr(), xm.data_ptr(), ctypes.byref(wd), ctypes.byref(geom), ctypes.byref(ep),
                                _ptr(ws), max(ws_bytes, 0), torch.cuda.current_stream().cuda_stream)
    _lib.check(st, "rick_conv_tc_w")
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
