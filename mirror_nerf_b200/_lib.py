"""ctypes binding of libmnrf.so (include/mnrf.h).  There is no fallback: if the library is missing, or a
call fails, this raises -- the product never silently computes on the CPU or through plain PyTorch ops."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MNRF_LIB: another build flavour of the same library (tools/tc_trace.py uses the -DMNRF_TC_TRACE build); never a fallback
LIB_PATH = os.environ.get("MNRF_LIB") or os.path.join(_HERE, "lib", "libmnrf.so")

IMPL_TC3, IMPL_TC2, IMPL_TC1, IMPL_FP32 = 3, 2, 1, 0
IMPL_BY_NAME = {"tc3": IMPL_TC3, "tc2": IMPL_TC2, "tc1": IMPL_TC1, "fp32": IMPL_FP32}
NUM_PARAM_TENSORS = 32
NUM_HASH_PARAM_TENSORS = 12
RAW_STRIDE = 8

c_float_p = C.c_void_p  # device or host pointers are passed as integers
c_int = C.c_int


class CompositeOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "weights", "opacity", "rgb", "depth", "mirror_mask", "pred_normal", "surface_normal",
        "surface_normal_grad", "normal_dif", "x_surface")]


class LevelCfg(C.Structure):
    _fields_ = [("n_samples", c_int), ("n_importance", c_int), ("use_disp", c_int), ("perturb", C.c_float),
                ("noise_std", C.c_float), ("white_back", c_int), ("test_time", c_int), ("compute_normal", c_int),
                ("rerun_coarse_on_fine", c_int), ("impl", c_int), ("early_termination_eps", C.c_float),
                ("no_fused_composite", c_int), ("dir_source", C.c_void_p), ("stats", C.c_void_p)]


class LevelRng(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("perturb_u", "noise_coarse", "u_pdf", "noise_fine")]


class LevelOut(C.Structure):
    _fields_ = [("z_coarse", C.c_void_p), ("coarse", CompositeOut), ("normal_coarse", C.c_void_p),
                ("z_fine", C.c_void_p), ("fine", CompositeOut), ("normal_fine", C.c_void_p)]


class TraceCfg(C.Structure):
    _fields_ = [("level", LevelCfg), ("max_recursive_level", c_int), ("only_trace_rays_in_mirrors", c_int),
                ("trace_ray_times", c_int), ("normal_noise_std", C.c_float), ("noise_seed", C.c_uint64)]


class TraceOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "rgb", "rgb_direct", "rgb_reflect", "depth", "depth_reflect", "opacity", "mirror_mask", "surface_normal",
        "x_surface", "reflect_direction", "level_rays")]


class TrainCfg(C.Structure):
    _fields_ = [("S", c_int), ("compute_normal", c_int), ("white_back", c_int), ("noise_std", C.c_float),
                ("detach_density_for_mask_loss", c_int), ("detach_density_for_normal_loss", c_int)]


TRAIN_GRAD_FIELDS = ("rgb", "depth", "opacity", "mirror_mask", "surface_normal", "surface_normal_grad", "normal_dif",
                     "x_surface", "weights", "pred_normal", "normal")


class TrainGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in TRAIN_GRAD_FIELDS]


_SIGS = {
    "mnrf_last_error": (C.c_char_p, []),
    "mnrf_abi_version": (c_int, []),
    "mnrf_launch_count": (C.c_int64, []),
    "mnrf_profile_enable": (c_int, [c_int]),
    "mnrf_profile_collect": (c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "mnrf_debug_set_trace": (c_int, [C.c_void_p, C.c_int64]),
    "mnrf_debug_set_tc_schedule": (c_int, [c_int]),
    "mnrf_macs_full": (C.c_int64, []),
    "mnrf_macs_sigma_only": (C.c_int64, []),
    "mnrf_field_create": (c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]),
    "mnrf_hash_field_create": (c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int64, C.c_float,
                                       C.POINTER(C.c_float), C.POINTER(c_int), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                       C.c_void_p]),
    "mnrf_field_update": (c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]),
    "mnrf_field_destroy": (None, [C.c_void_p]),
    "mnrf_field_has_normal": (c_int, [C.c_void_p]),
    "mnrf_field_has_mirror": (c_int, [C.c_void_p]),
    "mnrf_field_eval_rays": (c_int, [C.c_void_p, c_int, c_float_p, c_float_p, c_int, c_int, c_int, c_float_p,
                                     c_float_p, c_float_p, C.c_void_p]),
    "mnrf_field_eval_points": (c_int, [C.c_void_p, c_int, c_float_p, c_int, c_int, c_float_p, c_float_p, c_float_p,
                                       c_float_p, c_float_p, c_float_p, C.c_void_p]),
    "mnrf_generate_rays": (c_int, [c_int, c_int, C.c_float, C.POINTER(C.c_float), C.c_float, C.c_float, c_float_p,
                                   C.c_void_p]),
    "mnrf_embed": (c_int, [c_float_p, c_int, c_int, c_float_p, C.c_void_p]),
    "mnrf_coarse_z": (c_int, [c_float_p, c_int, c_float_p, c_int, c_int, C.c_float, c_float_p, c_float_p, C.c_void_p]),
    "mnrf_searchsorted_right": (c_int, [c_float_p, c_int, c_int, c_float_p, c_int, c_int, C.c_void_p, C.c_void_p]),
    "mnrf_sample_pdf": (c_int, [c_float_p, c_float_p, c_int, c_int, c_int, c_float_p, c_int, c_float_p, c_float_p,
                                C.c_void_p, c_float_p, C.c_void_p]),
    "mnrf_sample_pdf_bins": (c_int, [c_float_p, c_float_p, c_int, c_int, c_int, c_float_p, c_int, c_float_p,
                                     C.c_void_p, c_float_p, C.c_void_p]),
    "mnrf_composite": (c_int, [c_float_p, c_float_p, c_float_p, c_int, c_float_p, c_float_p, c_float_p, C.c_float,
                               c_int, c_int, c_int, C.POINTER(CompositeOut), C.c_void_p]),
    "mnrf_level_workspace_bytes": (C.c_int64, [c_int, C.POINTER(LevelCfg)]),
    "mnrf_level_workspace_bytes_for": (C.c_int64, [C.c_void_p, C.c_void_p, c_int, C.POINTER(LevelCfg), c_int]),
    "mnrf_render_level": (c_int, [C.c_void_p, C.c_void_p, c_float_p, c_int, C.POINTER(LevelCfg), C.POINTER(LevelRng),
                                  c_float_p, c_float_p, C.c_void_p, C.c_int64, C.POINTER(LevelOut), C.c_void_p]),
    "mnrf_render_level_host": (c_int, [C.c_void_p, C.c_void_p, c_float_p, c_int, C.POINTER(LevelCfg), c_float_p,
                                       c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p,
                                       C.c_void_p]),
    "mnrf_recursive_workspace_bytes": (C.c_int64, [C.c_void_p, C.c_void_p, c_int, C.POINTER(TraceCfg), C.c_int64]),
    "mnrf_render_recursive": (c_int, [C.c_void_p, C.c_void_p, c_float_p, c_int, C.POINTER(TraceCfg), c_float_p, c_float_p,
                                      c_float_p, C.c_void_p, C.c_int64, C.POINTER(TraceOut), C.c_void_p]),
    "mnrf_train_fwd_workspace_bytes": (C.c_int64, [c_int, c_int, c_int]),
    "mnrf_train_bwd_workspace_bytes": (C.c_int64, [c_int, c_int, c_int]),
    "mnrf_field_train_fwd_workspace_bytes": (C.c_int64, [C.c_void_p, c_int, c_int, c_int]),
    "mnrf_field_train_bwd_workspace_bytes": (C.c_int64, [C.c_void_p, c_int, c_int, c_int]),
    "mnrf_train_pass_fwd": (c_int, [C.c_void_p, c_float_p, c_float_p, c_float_p, c_int, C.POINTER(TrainCfg), C.c_void_p,
                                    C.c_int64, C.POINTER(CompositeOut), c_float_p, C.c_void_p]),
    "mnrf_train_pass_bwd": (c_int, [C.c_void_p, c_float_p, c_float_p, c_float_p, c_int, C.POINTER(TrainCfg), C.c_void_p,
                                    C.c_int64, C.c_void_p, C.c_int64, C.POINTER(TrainGrads), c_float_p,
                                    C.POINTER(C.c_void_p), c_float_p, c_float_p, C.c_void_p]),
    "mnrf_train_set_gemm": (c_int, [c_int]),
    "mnrf_debug_gemm_bench": (c_int, [C.c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, C.POINTER(C.c_float)]),
    "mnrf_adam_step": (c_int, [c_float_p, c_float_p, c_float_p, c_float_p, C.c_int64, C.c_float, C.c_float, C.c_float,
                               C.c_float, C.c_float, c_int, C.c_float, C.c_void_p]),
    "mnrf_peer_shard": (c_int, [C.c_int64, c_int, c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mnrf_peer_allreduce_adam": (c_int, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), c_int, c_int, c_float_p, c_float_p,
                                         C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, c_int, C.c_void_p]),
    "mnrf_axpy": (c_int, [c_float_p, c_float_p, C.c_int64, C.c_float, C.c_void_p]),
    "mnrf_reflect_rays": (c_int, [c_float_p, c_float_p, c_float_p, c_float_p, c_int, C.c_float, c_float_p, c_float_p,
                                  C.c_void_p, C.c_void_p]),
    "mnrf_compact_rows": (c_int, [c_float_p, c_float_p, c_int, c_int, c_float_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mnrf_axpy_rows": (c_int, [c_float_p, c_float_p, C.c_void_p, c_int, c_int, C.c_float, C.c_float, C.c_void_p]),
    "mnrf_blend_reflection": (c_int, [c_float_p, c_float_p, c_float_p, c_float_p, C.c_void_p, c_int, c_float_p,
                                      c_float_p, c_float_p, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(sorted(_SIGS))

_lib = None


class MnrfError(RuntimeError):
    pass


def load():
    """Load libmnrf.so (once).  Raises if it has not been built: run __graft_entry__.build() or
    `make -C mirror_nerf_b200/csrc`."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MnrfError(f"{LIB_PATH} not found: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                        "(there is no CPU / PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    if lib.mnrf_abi_version() != 2:
        raise MnrfError("libmnrf.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().mnrf_last_error().decode("utf-8", "replace")
        raise MnrfError(f"{what} failed (rc={rc}): {msg}")


def launch_count():
    return int(load().mnrf_launch_count())
