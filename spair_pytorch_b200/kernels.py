"""ctypes binding of libspair_b200.so (the C-ABI declared in include/spair_b200.h).

Every function here is a thin, allocation-free call into one ``extern "C"`` entry point: it
validates the tensors, passes raw device pointers + the current CUDA stream, and raises on a
non-zero return code.  There is NO CPU implementation and no alternative backend: if the
library is missing it is built with nvcc (``_build.py``); if that fails, or a tensor is not a
CUDA fp32 tensor, the call raises.  PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

from . import _build

_c_float_p = ctypes.c_void_p
_LIB = None
_LOCK = threading.Lock()


class BoxGeom(ctypes.Structure):
    """``spair_box_geom`` of include/spair_b200.h (reference models.py:339-374)."""
    _fields_ = [(n, ctypes.c_float) for n in
                ("yx_scale", "yx_min", "hw_scale", "hw_min", "anchor", "img_h", "img_w", "cell_ratio_y", "cell_ratio_x")]


class SweepDims(ctypes.Structure):
    """``spair_sweep_dims`` of include/spair_b200.h."""
    _fields_ = [(n, ctypes.c_int) for n in
                ("B", "HW", "Hc", "Wc", "F", "A", "P", "C", "Ih", "Iw", "G", "ipc", "n_wavefronts", "max_cells", "n_nb")]


class SweepMLP(ctypes.Structure):
    """``spair_sweep_mlp`` of include/spair_b200.h."""
    _fields_ = [("wt", ctypes.c_void_p * 3), ("b", ctypes.c_void_p * 3), ("k", ctypes.c_int * 3), ("n", ctypes.c_int * 3),
                ("x", ctypes.c_void_p), ("ld_x", ctypes.c_int), ("h0", ctypes.c_void_p), ("h1", ctypes.c_void_p),
                ("y", ctypes.c_void_p)]


class SweepPack(ctypes.Structure):
    """``spair_sweep_pack`` of include/spair_b200.h."""
    _fields_ = [("w", ctypes.c_void_p), ("n", ctypes.c_int), ("k", ctypes.c_int), ("fwd", ctypes.c_void_p),
                ("bwd", ctypes.c_void_p)]


class SweepMLPBwd(ctypes.Structure):
    """``spair_sweep_mlp_bwd`` of include/spair_b200.h."""
    _fields_ = [("w", ctypes.c_void_p * 3), ("k", ctypes.c_int * 3), ("n", ctypes.c_int * 3),
                ("h0", ctypes.c_void_p), ("h1", ctypes.c_void_p), ("y", ctypes.c_void_p),
                ("dx", ctypes.c_void_p), ("ld_dx", ctypes.c_int), ("dh0", ctypes.c_void_p), ("dh1", ctypes.c_void_p),
                ("dy", ctypes.c_void_p)]


class SpairKernelError(RuntimeError):
    pass


_P, _I, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
_SIGNATURES = {
    "spair_abi_version": [],
    "spair_base_grid": [_I, _P],
    "spair_context_gather_fwd": [_P] * 6 + [_P, _I, _P, _I] + [_I] * 5 + [_P, _I, _P, _I, _P, _I, _P],
    "spair_context_grad_gather": [_P, _I, _P, _I, _P, _I, _I, _P, _I, _P, _P, _I, _I, _I, _I, _I, _P, _I, _P],
    "spair_box_head_fwd": [_P, _I, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _I, _P, _I, _P, _I, _I, _P, _I, _P],
    "spair_box_head_bwd": [_P, _I, _P, _P, _I, _I, _I, _I, _P, _P, _P, _I, _P, _I, _P, _I, _P, _I, _P, _P, _P, _I, _I, _P,
                           _I, _P, _P],
    "spair_normal_head_fwd": [_P, _I, _I, _P, _P, _I, _I, _I, _I, _F, _P, _P, _P, _I, _P, _I, _P, _I, _I, _P, _I, _P],
    "spair_normal_head_bwd": [_P, _I, _I, _P, _P, _I, _I, _I, _I, _F, _P, _P, _I, _P, _I, _P, _I, _P, _P, _P, _I, _I,
                              _P, _I, _P, _P],
    "spair_pres_head_fwd": [_P, _I, _P, _P, _I, _I, _I, _P, _P],
    "spair_pres_head_bwd": [_P, _I, _P, _P, _I, _I, _I, _P, _P, _I, _P, _P, _P],
    "spair_glimpse_fwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P],
    "spair_glimpse_bwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P, _P, _P],
    "spair_paste_fwd": [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P],
    "spair_paste_bwd": [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "spair_render_num_tiles": [_I, _I, _I],
    "spair_render_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _F, _F, _I, _P, _P, _P, _P, _P],
    "spair_render_bwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _F, _F, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "spair_kl_fwd": [_P] * 6 + [_I, _I, _I, _P, _P, _P, _P],
    "spair_kl_bwd": [_P] * 8 + [_I, _I, _I, _P, _P, _P, _P],
    "spair_relu_bwd": [_P, _I, _P, _I, _I, _I, _P],
    "spair_stem_bwd_ctas": [],
    "spair_stem_conv_fwd": [_P, _P, _P] + [_I] * 12 + [_P, _P],
    "spair_stem_conv_bwd": [_P, _P, _P] + [_I] * 12 + [_P, _P, _P, _P],
    "spair_broadcast_rows": [_P, _I, _I, _P, _P],
    "spair_gemm_block_n": [_I, _I],
    "spair_gemm_splits": [_I, _I, _I],
    "spair_gemm3x": [_P, _I, _I, _P, _I, _I, _P, _I, _I, _I, _I, _P, _I, _I, _F, _F, _F, _P, _I, _P, _I, _P, _P, _P],
    "spair_split_tf32": [_P, _P, _P, _I, _P],
    "spair_conv_gemm3x": [_P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P, _I, _I, _P, _I, _P, _I, _P, _I, _P, _P, _P],
    "spair_conv_dgrad3x": [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "spair_im2col_nhwc": [_P, _I, _I, _I, _I, _I, _I, _P, _P],
    "spair_col2im_nhwc": [_P, _I, _I, _I, _I, _I, _I, _P, _P],
    "spair_transpose_batched": [_P, _I, _I, _I, _P, _P],
    "spair_colsum_chunks": [_I, _I],
    "spair_relu_bwd_colsum": [_P, _I, _P, _I, _I, _I, _P, _P, _P],
    "spair_sweep_max_rows": [],
    "spair_sweep_pack_weights": [_P, _I, _P],
    "spair_sweep_fwd": [_P] * 23 + [_P],
    "spair_sweep_bwd": [_P] * 23 + [_P],
    "spair_sweep_tc_stream_floats": [_P, _P, _I, _I],
    "spair_sweep_tc_pack": [_P, _P, _P, _I, _P, _P, _P],
    "spair_sweep_fwd_tc": [_P] * 24 + [_P],
    "spair_sweep_bwd_tc": [_P] * 24 + [_P],
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


RENDER_MAX_TEXELS, RENDER_MAX_CHANNELS = 1024, 4   # spair_render_fwd/bwd: G*G <= 1024, C <= 4 (csrc/render.cu)
NUM_SMS = 148          # kSMs of csrc/common.cuh (B200: 2 dies x 74 SMs)
MAX_NEIGHBOURS = 12   # SPAIR_MAX_NEIGHBOURS of include/spair_b200.h (N_LOOKBACK <= 2)
ABI_VERSION = 5       # SPAIR_ABI_VERSION of include/spair_b200.h this binding was written against


def lib() -> ctypes.CDLL:
    """Loads (building first if stale or missing) the kernel library.  Raises if unavailable."""
    global _LIB
    if _LIB is None:
        with _LOCK:
            if _LIB is None:
                path = _build.LIB_PATH
                if _build.is_stale():
                    try:
                        path = _build.build()
                    except Exception as e:  # keep a prebuilt library if nvcc is absent on this box
                        if not os.path.exists(path):
                            raise SpairKernelError("libspair_b200.so is missing and could not be built: %s" % e) from e
                handle = ctypes.CDLL(path)
                for name, argtypes in _SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.argtypes = argtypes
                    fn.restype = ctypes.c_int
                if handle.spair_abi_version() != ABI_VERSION:
                    raise SpairKernelError("libspair_b200.so ABI version %d, binding expects %d: rebuild with "
                                           "`python __graft_entry__.py`" % (handle.spair_abi_version(), ABI_VERSION))
                _LIB = handle
    return _LIB


def require_cuda(t, what="input"):
    """The product path has no CPU implementation: refuse anything that is not a CUDA tensor."""
    if not t.is_cuda:
        raise SpairKernelError("%s must be a CUDA tensor: the SPAIR per-cell pipeline runs on sm_100a kernels and "
                               "has no CPU implementation" % what)


def base_grid(n: int) -> torch.Tensor:
    """Host helper: the n-point normalised base grid the kernels use (see spair_base_grid)."""
    out = torch.empty(n, dtype=torch.float32)
    _check(lib().spair_base_grid(n, out.data_ptr()), "spair_base_grid")
    return out


# kernels launched per C-ABI call (render_bwd = prep + object kernel; paste_bwd's memset is not a kernel)
_LAUNCHES_PER_CALL = {"spair_render_bwd": 2, "spair_base_grid": 0, "spair_sweep_max_rows": 0, "spair_stem_bwd_ctas": 0,
                      "spair_sweep_tc_stream_floats": 0,
                      "spair_stem_conv_bwd": 2}
LAUNCH_COUNT = 0


def launch_count() -> int:
    """Number of libspair_b200 kernel launches issued by this process so far."""
    return LAUNCH_COUNT


def _check(code: int, name: str) -> None:
    global LAUNCH_COUNT
    LAUNCH_COUNT += _LAUNCHES_PER_CALL.get(name, 1)
    if code != 0:
        if code < 0:
            raise SpairKernelError("%s rejected its arguments (SPAIR_ERR_INVALID)" % name)
        raise SpairKernelError("%s failed: CUDA error %d" % (name, code))


def _ptr(t, name="tensor"):
    """Device pointer of a CUDA fp32 tensor (None -> NULL).  No CPU path exists."""
    if t is None:
        return None
    if not t.is_cuda:
        raise SpairKernelError("%s must live on a CUDA device: the SPAIR kernels have no CPU implementation" % name)
    if t.dtype != torch.float32:
        raise SpairKernelError("%s must be float32, got %s" % (name, t.dtype))
    return t.data_ptr()


def _iptr(t, name="index tensor"):
    if t is None:
        return None
    if not t.is_cuda or t.dtype != torch.int32 or not t.is_contiguous():
        raise SpairKernelError("%s must be a contiguous CUDA int32 tensor" % name)
    return t.data_ptr()


def _ld(t):
    """Row stride (leading dimension) of a 2-D view whose rows are contiguous."""
    if t is None:
        return 0
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise SpairKernelError("expected a 2-D tensor with unit column stride, got shape %s stride %s"
                               % (tuple(t.shape), t.stride()))
    return t.stride(0)


def _contig(t, name):
    if t is not None and not t.is_contiguous():
        raise SpairKernelError("%s must be contiguous" % name)
    return t


def _stream():
    return torch.cuda.current_stream().cuda_stream


def parallel_branches(device, stream_pool, thunks):
    """Runs ``thunks[0]`` on the current stream of ``device`` and the others on side streams forked from / joined to it with
    events (CUDA-graph capture records them as parallel branches).  ``stream_pool(device, n)`` returns n side streams."""
    cur = torch.cuda.current_stream(device)
    fork = torch.cuda.Event()
    fork.record(cur)
    sides = stream_pool(device, len(thunks) - 1)
    for fn, st in zip(thunks[1:], sides):
        st.wait_event(fork)
        with torch.cuda.stream(st):
            fn()
    thunks[0]()
    for st in sides:
        cur.wait_stream(st)


def _device_guarded(fn):
    """Every launch wrapper below passes raw pointers and ``torch.cuda.current_stream()`` of the CURRENT device to a
    ``<<<>>>`` launch.  If the tensors live on another device (``SPAIR(...).to('cuda:1')`` without ``set_device``) that
    would launch on the wrong GPU, so the wrapper switches to the device of its first CUDA tensor argument for the call."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        for a in args:
            if torch.is_tensor(a) and a.is_cuda:
                if a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kwargs)
                break
        return fn(*args, **kwargs)
    return wrapped


def _offsets_array(offsets):
    flat = [int(v) for pair in offsets for v in pair]
    return (ctypes.c_int * len(flat))(*flat), len(offsets)


# ----------------------------------------------------------------------------------------
# L0 context
# ----------------------------------------------------------------------------------------
def context_gather_fwd(feat, box, attr, depth, pres, edge, cells, offsets, dsts):
    """feat [B,F,Hc,Wc]; box/attr/depth/pres image-major; cells int32 [n]; dsts: up to three 2-D row views."""
    B, F, Hc, Wc = feat.shape
    A = attr.shape[-1]
    arr, n_nb = _offsets_array(offsets)
    d = list(dsts) + [None] * (3 - len(dsts))
    for t in (feat, box, attr, depth, pres, edge):
        _contig(t, "context input")
    _check(lib().spair_context_gather_fwd(_ptr(feat), _ptr(box), _ptr(attr), _ptr(depth), _ptr(pres), _ptr(edge),
                                          _iptr(cells), cells.numel(), arr, n_nb, B, F, Hc, Wc, A,
                                          _ptr(d[0]), _ld(d[0]), _ptr(d[1]), _ld(d[1]), _ptr(d[2]), _ld(d[2]), _stream()),
           "spair_context_gather_fwd")


def context_grad_gather(dxs, col0, cells, wf_pos, offsets, B, Hc, Wc, A, out):
    arr, n_nb = _offsets_array(offsets)
    d = list(dxs) + [None] * (3 - len(dxs))
    _check(lib().spair_context_grad_gather(_ptr(d[0]), _ld(d[0]), _ptr(d[1]), _ld(d[1]), _ptr(d[2]), _ld(d[2]), col0,
                                           _iptr(cells), cells.numel(), _iptr(wf_pos), arr, n_nb, B, Hc, Wc, A,
                                           _ptr(out), _ld(out), _stream()), "spair_context_grad_gather")


# ----------------------------------------------------------------------------------------
# heads
# ----------------------------------------------------------------------------------------
def box_head_fwd(y, eps, cells, B, HW, Wc, geom, box, z_where, dmean, dstd, xdsts, n_pt, pt_dst):
    x = list(xdsts) + [None] * (2 - len(xdsts))
    _check(lib().spair_box_head_fwd(_ptr(y), _ld(y), _ptr(_contig(eps, "eps")), _iptr(cells), cells.numel(), B, HW, Wc,
                                    ctypes.byref(geom), _ptr(_contig(box, "box")), _ptr(_contig(z_where, "z_where")),
                                    _ptr(dmean), _ptr(dstd), dmean.shape[-1],
                                    _ptr(x[0]), _ld(x[0]), _ptr(x[1]), _ld(x[1]), n_pt, _ptr(pt_dst), _ld(pt_dst),
                                    _stream()), "spair_box_head_fwd")


def box_head_bwd(y, eps, cells, B, HW, Wc, geom, wheel, d_boxes, d_zw_local, d_zw_img, d_dmean, d_dstd, ld_dist,
                 n_pt, d_pt_src, d_y):
    d = list(d_boxes) + [None] * (3 - len(d_boxes))
    _check(lib().spair_box_head_bwd(_ptr(y), _ld(y), _ptr(eps), _iptr(cells), cells.numel(), B, HW, Wc,
                                    ctypes.byref(geom), _ptr(wheel), _ptr(d[0]), _ld(d[0]), _ptr(d[1]), _ld(d[1]),
                                    _ptr(d[2]), _ld(d[2]), _ptr(d_zw_local), _ld(d_zw_local), _ptr(d_zw_img),
                                    _ptr(d_dmean), _ptr(d_dstd), ld_dist, n_pt, _ptr(d_pt_src), _ld(d_pt_src),
                                    _ptr(d_y), _stream()), "spair_box_head_bwd")


def normal_head_fwd(y, W, eps, cells, B, HW, squash, scale, out, dmean_ptr_view, dstd_ptr_view, ld_dist, xdsts, n_pt,
                    pt_dst):
    """dmean_ptr_view / dstd_ptr_view: views of the [B,HW,D] maps starting at this head's first column."""
    x = list(xdsts) + [None] * (2 - len(xdsts))
    _check(lib().spair_normal_head_fwd(_ptr(y), _ld(y), W, _ptr(_contig(eps, "eps")), _iptr(cells), cells.numel(), B, HW,
                                       int(squash), float(scale), _ptr(_contig(out, "out")), _ptr(dmean_ptr_view),
                                       _ptr(dstd_ptr_view), ld_dist, _ptr(x[0]), _ld(x[0]), _ptr(x[1]), _ld(x[1]),
                                       n_pt, _ptr(pt_dst), _ld(pt_dst), _stream()), "spair_normal_head_fwd")


def normal_head_bwd(y, W, eps, cells, B, HW, squash, scale, wheel, d_outs, d_out_img, d_dmean_view, d_dstd_view,
                    ld_dist, n_pt, d_pt_src, d_y):
    d = list(d_outs) + [None] * (3 - len(d_outs))
    _check(lib().spair_normal_head_bwd(_ptr(y), _ld(y), W, _ptr(eps), _iptr(cells), cells.numel(), B, HW, int(squash),
                                       float(scale), _ptr(wheel), _ptr(d[0]), _ld(d[0]), _ptr(d[1]), _ld(d[1]),
                                       _ptr(d[2]), _ld(d[2]), _ptr(d_out_img), _ptr(d_dmean_view), _ptr(d_dstd_view),
                                       ld_dist, n_pt, _ptr(d_pt_src), _ld(d_pt_src), _ptr(d_y), _stream()),
           "spair_normal_head_bwd")


def pres_head_fwd(y, u, cells, B, HW, pres):
    _check(lib().spair_pres_head_fwd(_ptr(y), _ld(y), _ptr(_contig(u, "u")), _iptr(cells), cells.numel(), B, HW,
                                     _ptr(_contig(pres, "pres")), _stream()), "spair_pres_head_fwd")


def pres_head_bwd(y, u, cells, B, HW, wheel, d_local, d_img, d_y):
    _check(lib().spair_pres_head_bwd(_ptr(y), _ld(y), _ptr(u), _iptr(cells), cells.numel(), B, HW, _ptr(wheel),
                                     _ptr(d_local), _ld(d_local), _ptr(d_img), _ptr(d_y), _stream()),
           "spair_pres_head_bwd")


def relu_bwd(dh, h):
    _check(lib().spair_relu_bwd(_ptr(dh), _ld(dh), _ptr(h), _ld(h), dh.shape[0], dh.shape[1], _stream()),
           "spair_relu_bwd")


# ----------------------------------------------------------------------------------------
# backbone stem: ZeroPad2d + Conv2d(C -> 128, 4x4, stride s) + bias + ReLU
# ----------------------------------------------------------------------------------------
def stem_supported(C: int, Cout: int, k) -> bool:
    return C in (1, 3) and Cout == 128 and tuple(k) == (4, 4)


def stem_conv_fwd(x, w, bias, stride: int, pad_t: int, pad_l: int, Ho: int, Wo: int, y, channels_last=False):
    """y [B,Cout,Ho,Wo], or [B,Ho,Wo,Cout] with ``channels_last``."""
    B, C, Ih, Iw = x.shape
    for t in (x, w, bias, y):
        _contig(t, "stem tensor")
    _check(lib().spair_stem_conv_fwd(_ptr(x), _ptr(w), _ptr(bias), B, C, Ih, Iw, w.shape[0], w.shape[2], stride, pad_t, pad_l,
                                     Ho, Wo, int(channels_last), _ptr(y), _stream()), "spair_stem_conv_fwd")


def broadcast_rows(row, rows: int, out):
    _check(lib().spair_broadcast_rows(_ptr(_contig(row, "row")), rows, row.numel(), _ptr(_contig(out, "out")), _stream()),
           "spair_broadcast_rows")


def stem_bwd_workspace(C: int, Cout: int, device) -> torch.Tensor:
    return torch.empty(lib().spair_stem_bwd_ctas() * Cout * (C * 16 + 4), device=device, dtype=torch.float32)


def stem_conv_bwd(x, y, dy, w_shape, stride: int, pad_t: int, pad_l: int, ws, d_w, d_bias, channels_last=False):
    B, C, Ih, Iw = x.shape
    Ho, Wo = (y.shape[1], y.shape[2]) if channels_last else (y.shape[2], y.shape[3])
    for t in (x, y, dy, ws, d_w, d_bias):
        _contig(t, "stem tensor")
    _check(lib().spair_stem_conv_bwd(_ptr(x), _ptr(y), _ptr(dy), B, C, Ih, Iw, w_shape[0], w_shape[2], stride, pad_t, pad_l,
                                     Ho, Wo, int(channels_last), _ptr(ws), _ptr(d_w), _ptr(d_bias), _stream()), "spair_stem_conv_bwd")


# ----------------------------------------------------------------------------------------
# tcgen05 3xTF32 GEMM (csrc/gemm.cu)
# ----------------------------------------------------------------------------------------
GEMM_EPI_NONE, GEMM_EPI_RELU, GEMM_EPI_TEXEL = 0, 1, 2
_GEMM_WS = {}


def gemm_supported(*mats) -> bool:
    """TMA needs 16-byte aligned rows: every operand a 2-D fp32 view with unit column stride and a row pitch that is a
    multiple of 4 floats."""
    return all(t.dim() == 2 and t.stride(1) == 1 and t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0 for t in mats)


def _gemm_workspace(device, floats):
    """Split-K partial products: one growing buffer per (device, stream) so that GEMMs forked to side streams never share it."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _GEMM_WS.get(key)
    if ws is None or ws.numel() < floats:
        ws = torch.empty(floats, device=device, dtype=torch.float32)
        _GEMM_WS[key] = ws
    return ws


_KINK_WS = {}
KINK_CAP = 1 << 20     # entries of the ReLU sign fix-up list (4 MB); overflowing entries simply keep the 3xTF32 value


def _kink_workspace(device):
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _KINK_WS.get(key)
    if ws is None:
        ws = torch.zeros(1 + KINK_CAP, device=device, dtype=torch.int32)
        _KINK_WS[key] = ws
    return ws


class SplitWeight:
    """A weight matrix with its TF32 hi / lo planes (``spair_split_tf32``), made once per step: the GEMM kernels then load the
    planes of their B operand instead of splitting every tile of it (half of the splitter work, identical results).
    ``w`` is the fp32 matrix itself (a view is fine as long as hi / lo are asked for the same view)."""

    def __init__(self, w):
        self.w = _contig(w.detach(), "weight")
        planes = torch.empty((2,) + tuple(self.w.shape), device=self.w.device, dtype=torch.float32)
        self.hi, self.lo = planes[0], planes[1]
        with torch.cuda.device(self.w.device):
            _check(lib().spair_split_tf32(_ptr(self.w), _ptr(self.hi), _ptr(self.lo), self.w.numel(), _stream()), "spair_split_tf32")

    @property
    def shape(self):
        return self.w.shape


def gemm3x(A, a_kmajor, B, b_kmajor, out, bias=None, epilogue=GEMM_EPI_NONE, period=2, scales=(1.0, 1.0, 0.0), splits=None,
           exact_relu=True):
    """out[M,N] = op(A) . op(B) (+ bias) on the tensor cores at fp32 accuracy (see spair_gemm3x in include/spair_b200.h).
    a_kmajor: A is [M,K] (else [K,M]); b_kmajor: B is [N,K] (else [K,N]); B may be a ``SplitWeight``.  ``exact_relu``: with
    the ReLU epilogue, outputs within rounding error of the kink are re-evaluated in float64 (two launches instead of one)."""
    B_hi = B_lo = None
    if isinstance(B, SplitWeight):
        B, B_hi, B_lo = B.w, B.hi, B.lo
    M, N = out.shape
    Kd = A.shape[1] if a_kmajor else A.shape[0]
    assert (A.shape[0] if a_kmajor else A.shape[1]) == M and (B.shape[0] if b_kmajor else B.shape[1]) == N
    assert (B.shape[1] if b_kmajor else B.shape[0]) == Kd
    if splits is None:
        splits = lib().spair_gemm_splits(M, N, Kd) if epilogue == GEMM_EPI_NONE else 1
    ws = _gemm_workspace(out.device, splits * M * N) if splits > 1 else None
    fix = exact_relu and epilogue == GEMM_EPI_RELU and a_kmajor and b_kmajor and splits == 1 and M * N < (1 << 32)
    kink = _kink_workspace(out.device) if fix else None
    _check(lib().spair_gemm3x(_ptr(A), _ld(A), int(a_kmajor), _ptr(B), _ld(B), int(b_kmajor), _ptr(out), _ld(out), M, N, Kd,
                              _ptr(bias), epilogue, period, float(scales[0]), float(scales[1]), float(scales[2]), _ptr(ws),
                              splits, kink.data_ptr() if fix else None, KINK_CAP if fix else 0, _ptr(B_hi), _ptr(B_lo), _stream()),
           "spair_gemm3x")
    if splits > 1 or fix:
        global LAUNCH_COUNT
        LAUNCH_COUNT += 1      # the fixed-order split-K reduction / the ReLU sign fix-up


def transpose_batched(x, out):
    """x [B,R,C] -> out [B,C,R] (both contiguous)."""
    B, R, C = x.shape
    _check(lib().spair_transpose_batched(_ptr(_contig(x, "x")), B, R, C, _ptr(_contig(out, "out")), _stream()),
           "spair_transpose_batched")


def relu_bwd_colsum(g, y, out):
    """In place g *= (y > 0) (skipped when y is None), out[c] = sum_r g[r, c].  Falls back to torch for column counts or
    row pitches that are not multiples of 4."""
    rows, cols = g.shape
    if cols % 4 or g.stride(0) % 4 or g.data_ptr() % 16 or (y is not None and (y.stride(0) % 4 or y.data_ptr() % 16)):
        if y is not None:
            relu_bwd(g, y)
        torch.sum(g, 0, out=out)
        return
    ws = _gemm_workspace(g.device, lib().spair_colsum_chunks(rows, cols) * cols)
    _check(lib().spair_relu_bwd_colsum(_ptr(g), _ld(g), _ptr(y), _ld(y), rows, cols, _ptr(ws), _ptr(_contig(out, "out")), _stream()),
           "spair_relu_bwd_colsum")
    global LAUNCH_COUNT
    LAUNCH_COUNT += 1


def conv_supported(x, k: int, stride: int) -> bool:
    """Shapes the TMA-im2col convolution (spair_conv_gemm3x) covers: channels-last fp32, C % 32 == 0, k, stride <= 8."""
    return x.dim() == 4 and x.is_contiguous() and x.shape[3] % 32 == 0 and 1 <= k <= 8 and 1 <= stride <= 8 and x.data_ptr() % 16 == 0


def conv_fwd(x, k: int, stride: int, wr, bias, out, relu: bool, exact_relu=True):
    """out[B*Ho*Wo, Cout] = act(patches(x) . wr^T + bias): x [B,H,W,C] channels-last, wr [Cout, k*k*C] in (kh, kw, c) order."""
    B, H, W, C = x.shape
    w_hi = w_lo = None
    if isinstance(wr, SplitWeight):
        wr, w_hi, w_lo = wr.w, wr.hi, wr.lo
    fix = relu and exact_relu and out.numel() < (1 << 32)
    kink = _kink_workspace(out.device) if fix else None
    _check(lib().spair_conv_gemm3x(_ptr(_contig(x, "x")), B, H, W, C, k, stride, 1, _ptr(wr), _ld(wr), _ptr(out), _ld(out),
                                   wr.shape[0], _ptr(bias), GEMM_EPI_RELU if relu else GEMM_EPI_NONE, None, 1,
                                   kink.data_ptr() if fix else None, KINK_CAP if fix else 0, _ptr(w_hi), _ptr(w_lo), _stream()),
           "spair_conv_gemm3x")
    if fix:
        global LAUNCH_COUNT
        LAUNCH_COUNT += 1


def conv_wgrad(x, k: int, stride: int, dy, d_wr):
    """d_wr[Cout, k*k*C] = dy^T . patches(x): dy [B*Ho*Wo, Cout]; deterministic split over the pixels."""
    B, H, W, C = x.shape
    Cout, KK = d_wr.shape
    splits = lib().spair_gemm_splits(Cout, KK, dy.shape[0])
    ws = _gemm_workspace(d_wr.device, splits * Cout * KK) if splits > 1 else None
    _check(lib().spair_conv_gemm3x(_ptr(_contig(x, "x")), B, H, W, C, k, stride, 2, _ptr(dy), _ld(dy), _ptr(d_wr), _ld(d_wr),
                                   Cout, None, GEMM_EPI_NONE, _ptr(ws), splits, None, 0, None, None, _stream()), "spair_conv_gemm3x")
    if splits > 1:
        global LAUNCH_COUNT
        LAUNCH_COUNT += 1


def conv_dgrad_supported(k: int, stride: int, Cout: int) -> bool:
    return k % stride == 0 and k // stride <= 8 and Cout % 32 == 0


def pack_dgrad_weights(w, stride: int):
    """w [Cout, Cin, k, k] -> the stride*stride class weights of spair_conv_dgrad3x, [s*s, Cin, T*T*Cout]."""
    Cout, Cin, k, _ = w.shape
    T = k // stride
    classes = []
    for py in range(stride):
        for px in range(stride):
            sub = w[:, :, py::stride, px::stride].flip(2, 3)               # [Cout, Cin, T, T] indexed [T-1-a][T-1-a']
            classes.append(sub.permute(1, 2, 3, 0).reshape(Cin, T * T * Cout))
    return torch.stack(classes).contiguous()


def conv_dgrad(dy, B: int, k: int, stride: int, wc, dx):
    """dx [B,H,W,Cin] = input gradient of the k x k / stride convolution; dy [B*Ho*Wo, Cout] channels-last rows."""
    _, H, W, Cin = dx.shape
    Cout = dy.shape[1]
    n_cls = stride * stride
    wc_hi = wc_lo = None
    if isinstance(wc, SplitWeight):
        wc, wc_hi, wc_lo = wc.w, wc.hi, wc.lo
    _check(lib().spair_conv_dgrad3x(_ptr(_contig(dy, "dy")), B, H, W, Cin, k, stride, Cout, _ptr(_contig(wc, "wc")),
                                    _ptr(_contig(dx, "dx")), _ptr(wc_hi), _ptr(wc_lo), _stream()), "spair_conv_dgrad3x")
    global LAUNCH_COUNT
    LAUNCH_COUNT += n_cls - 1


def im2col_nhwc(x, k: int, stride: int, col):
    """x [B,H,W,C] channels-last -> col [B*Ho*Wo, k*k*C] (see spair_im2col_nhwc)."""
    B, H, W, C = x.shape
    _check(lib().spair_im2col_nhwc(_ptr(_contig(x, "x")), B, H, W, C, k, stride, _ptr(_contig(col, "col")), _stream()),
           "spair_im2col_nhwc")


def col2im_nhwc(dcol, k: int, stride: int, dx):
    """Adjoint of im2col_nhwc: d_col [B*Ho*Wo, k*k*C] -> dx [B,H,W,C]."""
    B, H, W, C = dx.shape
    _check(lib().spair_col2im_nhwc(_ptr(_contig(dcol, "dcol")), B, H, W, C, k, stride, _ptr(_contig(dx, "dx")), _stream()),
           "spair_col2im_nhwc")


# ----------------------------------------------------------------------------------------
# fused forward sweep
# ----------------------------------------------------------------------------------------
def sweep_max_rows() -> int:
    return lib().spair_sweep_max_rows()


class PackedSweepWeights:
    """Both packed layouts of a list of ``nn.Linear`` weights [N, K] in ONE flat buffer, filled by one launch
    (``spair_sweep_pack_weights``).  ``fwd[i]`` / ``bwd[i]`` are views; keep this object alive while kernels use them."""

    def __init__(self, weights):
        for w in weights:
            require_cuda(w, "weight")
        self.shapes = [(int(w.shape[0]), int(w.shape[1])) for w in weights]
        sizes = [(((k + 3) // 4) * 4 * n, ((n + 3) // 4) * 4 * k) for n, k in self.shapes]
        self.flat = torch.empty(sum(a + b for a, b in sizes), device=weights[0].device, dtype=torch.float32)
        self.fwd, self.bwd = [], []
        off = 0
        for a, b in sizes:
            self.fwd.append(self.flat[off:off + a])
            self.bwd.append(self.flat[off + a:off + a + b])
            off += a + b
        self._sources = [_contig(w.detach(), "weight") for w in weights]
        arr = (SweepPack * len(weights))()
        for i, w in enumerate(self._sources):
            arr[i].w, arr[i].n, arr[i].k = _ptr(w), self.shapes[i][0], self.shapes[i][1]
            arr[i].fwd, arr[i].bwd = _ptr(self.fwd[i]), _ptr(self.bwd[i])
        with torch.cuda.device(weights[0].device):
            _check(lib().spair_sweep_pack_weights(arr, len(weights), _stream()), "spair_sweep_pack_weights")


class PackedSweepWeightsTC:
    """The weight streams of the tensor-core sweeps (``spair_sweep_tc_pack``): all 12 ``nn.Linear`` weights [N, K] (order
    box0 .. obj2) split into TF32 hi / lo, swizzled and laid out in the order the kernels consume them, one stream per
    direction, filled by one launch.  Keep this object alive while kernels use the streams."""

    def __init__(self, weights, forward=True, backward=True):
        assert len(weights) == 12, "four MLPs of two hidden layers + output"
        for w in weights:
            require_cuda(w, "weight")
        self.shapes = [(int(w.shape[0]), int(w.shape[1])) for w in weights]
        n = (ctypes.c_int * 12)(*[s[0] for s in self.shapes])
        k = (ctypes.c_int * 12)(*[s[1] for s in self.shapes])
        nf = lib().spair_sweep_tc_stream_floats(n, k, 12, 0) if forward else 0
        nb = lib().spair_sweep_tc_stream_floats(n, k, 12, 1) if backward else 0
        if nf < 0 or nb < 0 or nf + nb == 0:
            raise SpairKernelError("spair_sweep_tc_stream_floats rejected the layer list")
        self.flat = torch.empty(nf + nb, device=weights[0].device, dtype=torch.float32)
        self.fwd = self.flat[:nf] if forward else None       # a direction that is not asked for is not packed
        self.bwd = self.flat[nf:] if backward else None
        self._sources = [_contig(w.detach(), "weight") for w in weights]
        ptrs = (ctypes.c_void_p * 12)(*[_ptr(w) for w in self._sources])
        with torch.cuda.device(weights[0].device):
            _check(lib().spair_sweep_tc_pack(ptrs, n, k, 12, _ptr(self.fwd), _ptr(self.bwd), _stream()), "spair_sweep_tc_pack")


def sweep_mlp_desc(packed, first, biases, X, H0, H1, Y) -> SweepMLP:
    """packed: PackedSweepWeights; layers first .. first+2 are the two hidden layers and the output layer."""
    m = SweepMLP()
    for i, b in enumerate(biases):
        # tensor-core sweep: the unpacked weight (re-evaluation of ReLU pre-activations near zero), else the fwd-packed one
        m.wt[i] = _ptr(packed._sources[first + i]) if isinstance(packed, PackedSweepWeightsTC) else _ptr(packed.fwd[first + i])
        m.b[i] = _ptr(_contig(b, "bias"))
        m.n[i], m.k[i] = packed.shapes[first + i]
    m.x, m.ld_x, m.h0, m.h1, m.y = _ptr(X, "X"), _ld(X), _ptr(_contig(H0, "H0")), _ptr(_contig(H1, "H1")), \
        _ptr(_contig(Y, "Y"))
    return m


def sweep_fwd(dims: SweepDims, order, starts, offsets, image, feat, edge, eps_where, eps_attr, eps_depth, u_pres, geom,
              mlps, box, z_where, attr, depth, pres, dmean, dstd, tc_stream=None):
    arr, _ = _offsets_array(offsets)
    for t in (image, feat, edge, eps_where, eps_attr, eps_depth, u_pres, box, z_where, attr, depth, pres, dmean, dstd):
        _contig(t, "sweep tensor")
    args = (ctypes.byref(dims), _iptr(order), _iptr(starts), arr, _ptr(image), _ptr(feat), _ptr(edge),
            _ptr(eps_where), _ptr(eps_attr), _ptr(eps_depth), _ptr(u_pres), ctypes.byref(geom),
            ctypes.byref(mlps[0]), ctypes.byref(mlps[1]), ctypes.byref(mlps[2]), ctypes.byref(mlps[3]),
            _ptr(box), _ptr(z_where), _ptr(attr), _ptr(depth), _ptr(pres), _ptr(dmean), _ptr(dstd))
    if tc_stream is not None:
        _check(lib().spair_sweep_fwd_tc(*args, _ptr(tc_stream), _stream()), "spair_sweep_fwd_tc")
    else:
        _check(lib().spair_sweep_fwd(*args, _stream()), "spair_sweep_fwd")


def sweep_mlp_bwd_desc(packed, first, H0, H1, Y, dX, dH0, dH1, dY) -> SweepMLPBwd:
    """packed: PackedSweepWeights; layers first .. first+2 are the two hidden layers and the output layer."""
    m = SweepMLPBwd()
    for i in range(3):
        m.w[i] = None if isinstance(packed, PackedSweepWeightsTC) else _ptr(packed.bwd[first + i])
        m.n[i], m.k[i] = packed.shapes[first + i]
    m.h0, m.h1, m.y = _ptr(_contig(H0, "H0")), _ptr(_contig(H1, "H1")), _ptr(_contig(Y, "Y"))
    m.dx, m.ld_dx = _ptr(dX, "dX"), _ld(dX)
    m.dh0, m.dh1, m.dy = _ptr(_contig(dH0, "dH0")), _ptr(_contig(dH1, "dH1")), _ptr(_contig(dY, "dY"))
    return m


def sweep_bwd(dims: SweepDims, order, starts, wf_pos, offsets, image, z_where, eps_where, eps_attr, eps_depth, u_pres, wheel,
              geom, mlps, d_zw, d_attr, d_depth, d_pres, d_dmean, d_dstd, tc_stream=None):
    arr, _ = _offsets_array(offsets)
    for t in (image, z_where, eps_where, eps_attr, eps_depth, u_pres, d_zw, d_attr, d_depth, d_pres, d_dmean, d_dstd):
        _contig(t, "sweep tensor")
    args = (ctypes.byref(dims), _iptr(order), _iptr(starts), _iptr(wf_pos), arr, _ptr(image),
            _ptr(z_where), _ptr(eps_where), _ptr(eps_attr), _ptr(eps_depth), _ptr(u_pres), _ptr(wheel),
            ctypes.byref(geom), ctypes.byref(mlps[0]), ctypes.byref(mlps[1]), ctypes.byref(mlps[2]),
            ctypes.byref(mlps[3]), _ptr(d_zw), _ptr(d_attr), _ptr(d_depth), _ptr(d_pres), _ptr(d_dmean), _ptr(d_dstd))
    if tc_stream is not None:
        _check(lib().spair_sweep_bwd_tc(*args, _ptr(tc_stream), _stream()), "spair_sweep_bwd_tc")
    else:
        _check(lib().spair_sweep_bwd(*args, _stream()), "spair_sweep_bwd")


# ----------------------------------------------------------------------------------------
# G glimpse / paste
# ----------------------------------------------------------------------------------------
def glimpse_fwd(image, z_where, cells, B, HW, Gh, Gw, out):
    """cells None: row r samples image r with z_where[r] (plain stn); else wavefront rows."""
    _, C, Ih, Iw = image.shape
    _check(lib().spair_glimpse_fwd(_ptr(_contig(image, "image")), _ptr(_contig(z_where, "z_where")), _iptr(cells),
                                   0 if cells is None else cells.numel(), B, HW, C, Ih, Iw, Gh, Gw, _ptr(out), _ld(out),
                                   _stream()), "spair_glimpse_fwd")


def glimpse_bwd(image, z_where, cells, B, HW, Gh, Gw, d_out, d_zw_local, d_image=None):
    _, C, Ih, Iw = image.shape
    _check(lib().spair_glimpse_bwd(_ptr(_contig(image, "image")), _ptr(_contig(z_where, "z_where")), _iptr(cells),
                                   0 if cells is None else cells.numel(), B, HW, C, Ih, Iw, Gh, Gw, _ptr(d_out),
                                   _ld(d_out), _ptr(_contig(d_zw_local, "d_zw")), _ptr(_contig(d_image, "d_image")),
                                   _stream()), "spair_glimpse_bwd")


def paste_fwd(image, z_where, Oh, Ow, out):
    n, C, Gh, Gw = image.shape
    _check(lib().spair_paste_fwd(_ptr(_contig(image, "image")), _ptr(_contig(z_where, "z_where")), n, C, Gh, Gw, Oh, Ow,
                                 _ptr(_contig(out, "out")), _stream()), "spair_paste_fwd")


def paste_bwd(image, z_where, Oh, Ow, d_out, d_image, d_z_where):
    n, C, Gh, Gw = image.shape
    _check(lib().spair_paste_bwd(_ptr(_contig(image, "image")), _ptr(_contig(z_where, "z_where")), n, C, Gh, Gw, Oh, Ow,
                                 _ptr(_contig(d_out, "d_out")), _ptr(_contig(d_image, "d_image")),
                                 _ptr(_contig(d_z_where, "d_z_where")), _stream()), "spair_paste_bwd")


# ----------------------------------------------------------------------------------------
# R render
# ----------------------------------------------------------------------------------------
def render_num_tiles(B, Ih, Iw):
    return lib().spair_render_num_tiles(B, Ih, Iw)


def render_fwd(logits, z_where, z_depth, z_pres, B, HW, C, G, Ih, Iw, scales, recon, denom, target, bce_partial, decoded=False):
    _check(lib().spair_render_fwd(_ptr(_contig(logits, "logits")), _ptr(_contig(z_where, "z_where")),
                                  _ptr(_contig(z_depth, "z_depth")), _ptr(_contig(z_pres, "z_pres")), B, HW, C, G, Ih, Iw,
                                  float(scales[0]), float(scales[1]), float(scales[2]), int(decoded), _ptr(_contig(recon, "recon")),
                                  _ptr(_contig(denom, "denom")), _ptr(_contig(target, "target")),
                                  _ptr(_contig(bce_partial, "bce_partial")), _stream()), "spair_render_fwd")


def render_bwd(logits, z_where, z_depth, z_pres, B, HW, C, G, Ih, Iw, scales, recon, denom, d_recon, target, bce_scale,
               gs_ws, d_logits, d_z_where, d_z_depth, d_z_pres, decoded=False):
    _check(lib().spair_render_bwd(_ptr(_contig(logits, "logits")), _ptr(_contig(z_where, "z_where")),
                                  _ptr(_contig(z_depth, "z_depth")), _ptr(_contig(z_pres, "z_pres")), B, HW, C, G, Ih, Iw,
                                  float(scales[0]), float(scales[1]), float(scales[2]), int(decoded), _ptr(recon), _ptr(denom),
                                  _ptr(_contig(d_recon, "d_recon")), _ptr(_contig(target, "target")), _ptr(bce_scale),
                                  _ptr(gs_ws), _ptr(_contig(d_logits, "d_logits")), _ptr(_contig(d_z_where, "d_z_where")),
                                  _ptr(_contig(d_z_depth, "d_z_depth")), _ptr(_contig(d_z_pres, "d_z_pres")), _stream()),
           "spair_render_bwd")


# ----------------------------------------------------------------------------------------
# K KL
# ----------------------------------------------------------------------------------------
def kl_fwd(dmean, dstd, pres, prior_mean, prior_std, count_dist0, B, HW, A, kl_map, p_z, kl_sums):
    for t in (dmean, dstd, pres, prior_mean, prior_std, count_dist0, kl_map, p_z, kl_sums):
        _contig(t, "kl tensor")
    _check(lib().spair_kl_fwd(_ptr(dmean), _ptr(dstd), _ptr(pres), _ptr(prior_mean), _ptr(prior_std), _ptr(count_dist0),
                              B, HW, A, _ptr(kl_map), _ptr(p_z), _ptr(kl_sums), _stream()), "spair_kl_fwd")


def kl_bwd(dmean, dstd, pres, prior_mean, prior_std, kl_map, p_z, d_sums, B, HW, A, d_dmean, d_dstd, d_pres):
    for t in (dmean, dstd, pres, prior_mean, prior_std, p_z, d_sums, d_dmean, d_dstd, d_pres):
        _contig(t, "kl tensor")
    _check(lib().spair_kl_bwd(_ptr(dmean), _ptr(dstd), _ptr(pres), _ptr(prior_mean), _ptr(prior_std), _ptr(kl_map),
                              _ptr(p_z), _ptr(d_sums), B, HW, A, _ptr(d_dmean), _ptr(d_dstd), _ptr(d_pres), _stream()),
           "spair_kl_bwd")


# device guard on every launch wrapper (see _device_guarded)
for _name in ("context_gather_fwd", "context_grad_gather", "box_head_fwd", "box_head_bwd", "normal_head_fwd", "normal_head_bwd",
              "pres_head_fwd", "pres_head_bwd", "relu_bwd", "stem_conv_fwd", "broadcast_rows", "stem_conv_bwd", "sweep_fwd",
              "sweep_bwd", "gemm3x", "conv_fwd", "conv_wgrad", "conv_dgrad", "im2col_nhwc", "col2im_nhwc", "transpose_batched", "relu_bwd_colsum", "glimpse_fwd", "glimpse_bwd", "paste_fwd", "paste_bwd", "render_fwd", "render_bwd", "kl_fwd", "kl_bwd"):
    globals()[_name] = _device_guarded(globals()[_name])
del _name
