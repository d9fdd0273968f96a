"""ctypes binding of libmpb200.so (the C ABI declared in include/mpb200.h).

The library is built in-tree by `build()` (nvcc, sm_100a) and loaded lazily.  There is no fallback: if the shared
object is missing or a call fails, a RuntimeError is raised -- the product path never routes around the CUDA code.
"""
from __future__ import annotations

import ctypes
import glob
import os
import subprocess
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libmpb200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-shared",
              "-Xcompiler", "-fPIC"]

ACT_NONE, ACT_RELU, ACT_RELU_TANH, ACT_SIGMOID, ACT_TANH = 0, 1, 2, 3, 4
PREC_SPLIT_BF16, PREC_F16X2, PREC_F16_Q8 = 0, 1, 2   # mp_conv_desc.prec
FMT_NATIVE, FMT_SPLIT_BF16, FMT_F16, FMT_F16_Q8 = 0, 1, 2, 3   # mp_conv_desc.out_fmt / res_fmt
ABI_VERSION = 4


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


STAMP_PATH = LIB_PATH + ".stamp"      # sha256 of every source, header and flag the library was built from
OBJ_DIR = os.path.join(HERE, "build")


def _headers():
    return sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h")))


def _digest(paths) -> str:
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in paths:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def source_digest() -> str:
    return _digest(sources() + _headers())


def needs_build() -> bool:
    """True when the library is missing or was built from different sources / flags (content hash, not mtime: a prebuilt
    binary that travelled to another box is used only if it matches the sources that travelled with it)."""
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH)):
        return True
    with open(STAMP_PATH) as f:
        return f.read().strip() != source_digest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a (cross-compiles without a GPU) and link libmpb200.so.  Objects are
    cached per source under build/ keyed on the content of the source, the headers and the flags, and compiled in
    parallel."""
    if not force and not needs_build():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    cflags = [f for f in NVCC_FLAGS if f != "-shared"]
    hdrs = _headers()

    def compile_one(src):
        key = _digest([src] + hdrs)[:24]
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + "." + key + ".o")
        if force or not os.path.exists(obj):
            for old in glob.glob(os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".*.o")):
                os.remove(old)
            cmd = [nvcc] + cflags + ["-c", "-o", obj, src]
            if verbose:
                print(" ".join(cmd))
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc link failed:\n{r.stdout}\n{r.stderr}")
    with open(STAMP_PATH, "w") as f:
        f.write(source_digest() + "\n")
    return LIB_PATH


class ConvDesc(Structure):
    """mirror of `mp_conv_desc` (include/mpb200.h)"""
    _fields_ = [
        ("in_hi", c_void_p), ("in_lo", c_void_p), ("w_hi", c_void_p), ("w_lo", c_void_p), ("bias", c_void_p),
        ("res_f32", c_void_p), ("res_hi", c_void_p), ("res_lo", c_void_p),
        ("out_f32", c_void_p), ("out_hi", c_void_p), ("out_lo", c_void_p), ("stats", c_void_p),
        ("N", c_int), ("D", c_int), ("H", c_int), ("W", c_int), ("Cin", c_int), ("Cout", c_int),
        ("KD", c_int), ("KH", c_int), ("KW", c_int), ("Cout_pad", c_int), ("gn_groups", c_int), ("act", c_int),
        ("stride", c_int), ("in_c_off", c_int), ("in_C", c_int), ("out_c_off", c_int), ("out_C", c_int),
        ("prec", c_int),
        ("Cin2", c_int), ("in2_C", c_int), ("in2_c_off", c_int), ("stride2", c_int),
        ("in2_hi", c_void_p), ("in2_lo", c_void_p),
        ("out_fmt", c_int), ("res_fmt", c_int), ("corr_scale", c_float),
        ("out_q8_scale", c_float), ("res_q8_scale", c_float), ("acc_scale", c_float),
    ]


_P = c_void_p
_SIGNATURES = {
    "mp_abi_version": (c_int, []),
    "mp_last_error": (c_char_p, []),
    "mp_device_supported": (c_int, []),
    "mp_nchw_to_cl": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int64, _P]),
    "mp_cl_to_nchw": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int64, _P]),
    "mp_nchw_to_cl_pad16": (c_int, [_P, _P, _P, c_int, c_int, c_int64, _P]),
    "mp_split": (c_int, [_P, _P, _P, c_int64, _P]),
    "mp_avgpool2_cl": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "mp_upsample2x_linear_cl": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "mp_upsample2x_bilinear_hq": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, _P]),
    "mp_upsample_nearest_cl": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "mp_gn_stats": (c_int, [_P, _P, c_int, c_int64, c_int, c_int, _P]),
    "mp_gn_finalize": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int64, c_int, c_int, c_float, _P]),
    "mp_affine_act_cl": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int64, c_int, c_int, _P]),
    "mp_maxpool3x3s2_cl": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "mp_global_avgpool_cl": (c_int, [_P, _P, _P, _P, c_int, c_int64, c_int, _P]),
    "mp_im2col3x3_f16": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "mp_maxpool3x3s2_cl_f16": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P]),
    "mp_stem3x3_relu_maxpool_f16": (c_int, [_P, _P, _P, _P, c_float, _P, c_int, c_int, c_int, c_int, _P]),
    "mp_global_avgpool_cl_f16": (c_int, [_P, _P, c_int, c_int64, c_int, _P]),
    "mp_release_caches": (c_int, []),
    "mp_jpeg_info": (c_int, [_P, c_size_t, POINTER(c_int), POINTER(c_int)]),
    "mp_decode_jpeg_frames": (c_int, [_P, _P, c_int, _P, c_int, c_int, _P]),
    "mp_conv_tc": (c_int, [POINTER(ConvDesc), _P]),
    "mp_conv_simt": (c_int, [POINTER(ConvDesc), _P]),
    "mp_conv_tc_supported": (c_int, [POINTER(ConvDesc)]),
    "mp_grid_sample3d": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "mp_apply_warping_field": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "mp_gather_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "mp_grid_sample3d_ws": (c_int, [_P, _P, _P, _P, c_size_t, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "mp_apply_warping_field_ws": (c_int, [_P, _P, _P, _P, c_size_t, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                          c_int, _P]),
    "mp_gs_brick_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "mp_grid_sample3d_brick": (c_int, [_P, _P, _P, _P, c_size_t, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                       c_int, _P]),
    "mp_apply_warping_field_brick": (c_int, [_P, _P, _P, _P, c_size_t, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                             c_int, c_int, _P]),
    "mp_gs_brick_tune": (c_int, [POINTER(c_int), c_int]),
    "mp_conv_wgrad": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "mp_upsample2x_linear_backward_cl": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "mp_maxpool3x3s2_forward_idx": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "mp_maxpool3x3s2_backward": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "mp_im2col_rgb_split": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "mp_pack_conv_weights": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "mp_conv_wgrad_tc_supported": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "mp_conv_wgrad_tc": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "mp_bias_grad": (c_int, [_P, _P, c_int64, c_int, _P]),
    "mp_group_norm_backward": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int64, c_int, c_int, c_float, _P]),
    "mp_apply_warping_field_backward": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "mp_warp_field": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P]),
    "mp_warp_fused_cl": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_int, _P]),
    "mp_gn_relu_conv3x3_head": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "mp_frames_u8_to_f32": (c_int, [_P, _P, c_int, c_int, c_int, c_float, c_float, _P]),
    "mp_frames_f32_to_u8": (c_int, [_P, _P, c_int, c_int, c_int, c_float, c_float, c_int, _P]),
    "mp_blur_subsample": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
}

EXPORTS = tuple(_SIGNATURES.keys())
_lib = None


def load():
    """Load libmpb200.so; raises RuntimeError (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the CUDA extension is mandatory; there is no CPU or ATen fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.mp_abi_version() != ABI_VERSION:
        raise RuntimeError("libmpb200.so ABI version mismatch")
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().mp_last_error()
        raise RuntimeError(f"libmpb200 {what} failed: {msg.decode() if msg else status}")
