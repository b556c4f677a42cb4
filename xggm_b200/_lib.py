"""ctypes binding of libxggm_b200.so (the C ABI in include/xggm_b200.h).

There is no fallback: if the shared library is missing, or the device is not a
compute-capability-10.x GPU, every entry point raises.
"""
import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libxggm_b200.so")

_vp, _i, _f, _d, _ll, _u64 = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_longlong, C.c_uint64

# name -> argtypes (return type is int unless listed in _RESTYPES)
SIGNATURES = {
    "xggm_abi_version": [],
    "xggm_strerror": [_i],
    "xggm_last_cuda_error": [],
    "xggm_launch_count": [],
    "xggm_prof_enable": [_i],
    "xggm_prof_read": [_vp, _vp, _vp],
    "xggm_debug_timeline": [_vp],
    "xggm_set_device": [_i],
    "xggm_device_check": [_i],
    "xggm_set_precision": [_i],
    "xggm_get_precision": [],
    "xggm_linear_work_bytes": [_i, _i, _i],
    "xggm_linear_fwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp],
    "xggm_linear_bwd_input": [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "xggm_linear_bwd_weight": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "xggm_adj_apply_work_bytes": [_i, _i, _i],
    "xggm_adj_apply_fwd": [_vp, _vp, _vp, _i, _i, _i, _f, _vp, _f, _vp, _vp],
    "xggm_adj_apply_bwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _f, _i, _vp, _vp],
    "xggm_layernorm_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp],
    "xggm_layernorm_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "xggm_gelu_ln_drop_fwd": [_vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _i, _i, _f, _i, _vp],
    "xggm_gelu_ln_drop_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _i, _i, _vp],
    "xggm_adj_regen_work_bytes": [_i, _i, _i],
    "xggm_adj_regen_fwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "xggm_adj_regen_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp],
    "xggm_planes_bytes": [_ll],
    "xggm_adj_regen_fwd_ex": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp],
    "xggm_adj_regen_bwd_ex": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp],
    "xggm_gnn_fwd_ex": [_i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "xggm_gnn_bwd_ex": [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp,
                        _i, _vp, _i, _i, _i, _i, _vp, _vp],
    "xggm_weight_planes_bytes": [_i, _i],
    "xggm_weight_planes_build": [_vp, _vp, _vp, _vp, _i, _vp],
    "xggm_linear_fwd_ex": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp],
    "xggm_linear_bwd_input_ex": [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp],
    "xggm_feat_noise_ex": [_vp, _vp, _d, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "xggm_feat_noise_philox": [_vp, _vp, _d, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "xggm_gnn_saved_floats": [_i, _i, _i, _i, _i],
    "xggm_gnn_work_floats": [_i, _i, _i, _i, _i],
    "xggm_gnn_fwd": [_i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "xggm_gnn_bwd": [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp,
                     _i, _i, _i, _i, _i, _vp],
    "xggm_gat_attn_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _i, _vp],
    "xggm_gat_attn_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _i, _vp],
    "xggm_gelu_fwd": [_vp, _vp, _ll, _vp],
    "xggm_gelu_bwd": [_vp, _vp, _vp, _ll, _vp],
    "xggm_mask_scale": [_vp, _vp, _f, _vp, _ll, _vp],
    "xggm_avg2_drop": [_vp, _vp, _vp, _f, _vp, _ll, _vp],
    "xggm_strip_diag": [_vp, _vp, _i, _i, _vp],
    "xggm_triu_scatter_fwd": [_vp, _vp, _i, _i, _vp],
    "xggm_triu_scatter_bwd": [_vp, _vp, _i, _i, _vp],
    "xggm_edge_noise": [_vp, _vp, _d, _vp, _vp, _i, _i, _vp],
    "xggm_feat_noise": [_vp, _vp, _d, _vp, _vp, _i, _i, _i, _i, _vp],
    "xggm_sum_nodes": [_vp, _vp, _i, _i, _i, _vp],
    "xggm_score_mse_fwd": [_vp, _vp, _d, _vp, _ll, _vp],
    "xggm_score_mse_bwd": [_vp, _vp, _vp, _d, _vp, _ll, _vp],
    "xggm_sym_kl_fwd": [_vp, _vp, _vp, _i, _i, _vp],
    "xggm_sym_kl_bwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "xggm_fuse_readout_fwd": [_vp, _vp, _vp, _i, _i, _i, _vp],
    "xggm_fuse_readout_bwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "xggm_node_tail_fwd": [_vp, _vp, _vp, _vp, _d, _d, _d, _vp, _vp, _i, _i, _i, _vp],
    "xggm_node_tail_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _d, _d, _d, _vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "xggm_bce_logits_fwd": [_vp, _vp, _d, _vp, _ll, _vp],
    "xggm_bce_logits_bwd": [_vp, _vp, _vp, _d, _vp, _ll, _vp],
    "xggm_grad_sumsq": [_vp, _ll, _vp, _i, _vp],
    "xggm_bertadam_step": [_vp, _vp, _vp, _vp, _ll, _d, _d, _d, _d, _d, _vp, _d, _vp],
    "xggm_bertadam_step_ex": [_vp, _vp, _vp, _vp, _ll, _d, _d, _d, _d, _d, _vp, _d, _vp, _vp],
    "xggm_visn_tail_supported": [_i, _i],
    "xggm_visn_tail_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp],
    "xggm_visn_tail_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp],
    "xggm_dp_bertadam_step": [_vp, _vp, _vp, _ll, _vp, _vp, _vp, _i, _d, _d, _d, _d, _d, _d, _vp, _vp, _vp],
    "xggm_sigmoid_fwd": [_vp, _vp, _ll, _vp],
    "xggm_sigmoid_bwd": [_vp, _vp, _vp, _ll, _vp],
    "xggm_keep_mask": [_vp, _ll, _f, _u64, _u64, _vp, _vp],
}
_RESTYPES = {"xggm_launch_count": C.c_ulonglong, "xggm_strerror": C.c_char_p, "xggm_last_cuda_error": C.c_char_p,
             "xggm_planes_bytes": _ll, "xggm_gnn_saved_floats": _ll, "xggm_gnn_work_floats": _ll, "xggm_linear_work_bytes": _ll, "xggm_adj_regen_work_bytes": _ll, "xggm_adj_apply_work_bytes": _ll,
             "xggm_weight_planes_bytes": _ll}

ABI_VERSION = 4
_lib = None
_checked_devices = set()


def kernel_launches():
    """CUDA kernels launched by the library so far (counted inside the .so)."""
    return int(load().xggm_launch_count())


def gemm_profile(on=None):
    """on=True/False: start/stop per-launch timing of the projection GEMMs.
    on=None: read -> (total_ms, launches, flops)."""
    lib = load()
    if on is not None:
        lib.xggm_prof_enable(int(on))
        return None
    ms, n, fl = C.c_double(), C.c_longlong(), C.c_double()
    rc = lib.xggm_prof_read(C.byref(ms), C.byref(n), C.byref(fl))
    if rc != 0:
        _raise(rc)
    return ms.value, n.value, fl.value


def load():
    """Load the shared library (idempotent).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m xggm_b200.build` "
            "(nvcc, sm_100a).  xggm_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    if lib.xggm_abi_version() != ABI_VERSION:
        raise RuntimeError("libxggm_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


class PhiloxSpec(C.Structure):
    """xggm_philox_t of include/xggm_b200.h."""
    _fields_ = [("seed", C.c_uint64), ("stream0", C.c_uint64), ("dev_epoch", C.c_void_p)]


class LrSchedule(C.Structure):
    """xggm_lr_schedule_t of include/xggm_b200.h."""
    _fields_ = [("step_dev", C.c_void_p), ("ticket_dev", C.c_void_p), ("warmup", C.c_double),
                ("t_total", C.c_longlong), ("schedule", C.c_int), ("advance", C.c_int)]


DP_MAX_RANKS, DP_MAX_RANGES, DP_CTL_BYTES = 16, 8, 512


class DpPeers(C.Structure):
    """xggm_dp_peers_t of include/xggm_b200.h."""
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("grad", C.c_void_p * DP_MAX_RANKS),
                ("param", C.c_void_p * DP_MAX_RANKS), ("ctl", C.c_void_p * DP_MAX_RANKS),
                ("grad_multicast", C.c_void_p), ("param_multicast", C.c_void_p)]


PRECISIONS = {"fp32": 0, "bf16": 1, "fp32_simt": 2}


def set_precision(mode):
    """Projection engine: 'fp32' (tcgen05, split-bf16 x3, default), 'bf16' (tcgen05, single pass),
    'fp32_simt' (exact fp32 FMA kernel).  Process-wide."""
    rc = load().xggm_set_precision(PRECISIONS[mode])
    if rc != 0:
        _raise(rc)


def get_precision():
    code = load().xggm_get_precision()
    return {v: k for k, v in PRECISIONS.items()}[code]


def _raise(code):
    lib = load()
    msg = lib.xggm_strerror(code).decode()
    if code == -2:
        msg += ": " + lib.xggm_last_cuda_error().decode()
    raise RuntimeError(f"xggm_b200: {msg}")


def check_device(t):
    """The kernels exist only as sm_100a SASS: refuse anything else."""
    if not t.is_cuda:
        raise RuntimeError("xggm_b200: tensors must live on a CUDA (B200) device; there is no CPU path")
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if idx != torch.cuda.current_device():
        # the call would enqueue on the CURRENT device's stream with another device's pointers
        raise RuntimeError(f"xggm_b200: tensor lives on cuda:{idx} but the current device is "
                           f"cuda:{torch.cuda.current_device()}; wrap the call in torch.cuda.device({idx})")
    if idx not in _checked_devices:
        rc = load().xggm_device_check(idx)
        if rc != 0:
            _raise(rc)
        _checked_devices.add(idx)
    return idx


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def f32(t, name="tensor"):
    """Validate a dense fp32 CUDA tensor (the ABI's only tensor type)."""
    if t.dtype != torch.float32:
        raise RuntimeError(f"xggm_b200: {name} must be float32, got {t.dtype}")
    check_device(t)
    return t if t.is_contiguous() else t.contiguous()


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def call(name, *args):
    """Invoke an int-returning entry point on the current torch stream."""
    lib = load()
    # autograd runs backward on its own threads: bind the library's runtime to torch's device
    rc = lib.xggm_set_device(torch.cuda.current_device())
    if rc == 0:
        rc = getattr(lib, name)(*args, stream())
    if rc != 0:
        _raise(rc)


def ptr_table(tensors):
    """Host array of device pointers (NULL for None)."""
    arr = (C.c_void_p * max(1, len(tensors)))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr
