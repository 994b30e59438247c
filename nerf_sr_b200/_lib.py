"""ctypes binding of libnsr_b200.so (include/nsr.h).  This is the stub a
reference maintainer would add; there is no fallback: if the
library is missing the import fails loudly."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NSR_LIB_PATH") or os.path.join(_HERE, "libnsr_b200.so")
ABI_VERSION = 1

NSR_OK = 0
STATUS_NAMES = {0: "NSR_OK", 1: "NSR_ERR_INVALID_ARG", 2: "NSR_ERR_UNSUPPORTED", 3: "NSR_ERR_NOT_PACKED",
                4: "NSR_ERR_WORKSPACE", 5: "NSR_ERR_CUDA", 6: "NSR_ERR_NO_DEVICE"}
PRECISIONS = {"fp32_simt": 0, "bf16x3": 1, "fp16x3": 2, "bf16": 3}


class NsrConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("precision", C.c_int32),
                ("D", C.c_int32), ("W", C.c_int32), ("skips_mask", C.c_uint32), ("no_dir", C.c_int32),
                ("color_activation", C.c_int32), ("deg_pos", C.c_int32), ("deg_dir", C.c_int32),
                ("no_xyz", C.c_int32), ("no_logscale", C.c_int32), ("n_coarse", C.c_int32),
                ("n_importance", C.c_int32), ("lindisp", C.c_int32), ("white_bkgd", C.c_int32),
                ("sigma_activation", C.c_int32), ("gamma_correct", C.c_int32), ("noise_std", C.c_float),
                ("viewdir_offset", C.c_int32), ("reserved", C.c_int32 * 8)]


class NsrRng(C.Structure):
    _fields_ = [("u_coarse", C.c_void_p), ("noise_coarse", C.c_void_p), ("u_fine", C.c_void_p),
                ("noise_fine", C.c_void_p)]


class NsrOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("coarse_comp_rgbs", "coarse_depth", "coarse_opacity", "coarse_weights", "fine_comp_rgbs",
                 "fine_depth", "fine_opacity", "fine_weights", "z_fine")]


class NsrOutGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("coarse_comp_rgbs", "coarse_depth", "coarse_opacity", "fine_comp_rgbs", "fine_depth", "fine_opacity")]


class NsrLossTerms(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("s", C.c_int32), ("lambda_mse", C.c_float), ("lambda_var", C.c_float),
                ("lambda_depth_var", C.c_float), ("far_plane", C.c_float), ("lambda_hr", C.c_float), ("reserved", C.c_int32 * 5)]


class NsrRayGen(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("H", C.c_int32), ("W", C.c_int32), ("s", C.c_int32), ("focal", C.c_float),
                ("ndc", C.c_int32), ("near_plane", C.c_float), ("far_plane", C.c_float), ("use_pixel_centers", C.c_int32),
                ("unified_dir", C.c_int32), ("reserved", C.c_int32 * 6)]


class NsrLrOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("coarse_rgb", "coarse_depth", "fine_rgb", "fine_depth")]


class NsrPassOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("comp_rgbs", "depth", "opacity", "weights", "raw")]


# name -> (restype, argtypes); mirrors include/nsr.h one to one
SIGNATURES = {
    "nsr_abi_version": (C.c_int, []),
    "nsr_create": (C.c_int, [C.POINTER(NsrConfig), C.POINTER(C.c_void_p)]),
    "nsr_destroy": (C.c_int, [C.c_void_p]),
    "nsr_last_error": (C.c_char_p, [C.c_void_p]),
    "nsr_param_count": (C.c_int, [C.c_void_p]),
    "nsr_param_numel": (C.c_int64, [C.c_void_p, C.c_int]),
    "nsr_pack_weights": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    "nsr_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64]),
    "nsr_render": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.POINTER(NsrRng),
                             C.POINTER(NsrOutputs), C.c_void_p, C.c_size_t, C.c_void_p]),
    "nsr_frame_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64, C.c_int]),
    "nsr_render_frame": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_float), C.POINTER(NsrRayGen),
                                   C.c_int, C.POINTER(NsrRng), C.POINTER(NsrOutputs), C.POINTER(NsrLrOutputs), C.c_void_p,
                                   C.c_size_t, C.c_void_p]),
    "nsr_render_pass": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int,
                                  C.c_void_p, C.POINTER(NsrPassOutputs), C.c_void_p, C.c_size_t, C.c_void_p]),
    "nsr_sample_coarse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsr_resample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsr_posenc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "nsr_box_average": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nsr_generate_rays": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_float, C.c_int,
                                    C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "nsr_lr_metrics": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsr_assemble_frame": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float,
                                     C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsr_render_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nsr_render_pose_host": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                       C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    # training (scope row f-1)
    "nsr_grad_numel": (C.c_int64, [C.c_void_p]),
    "nsr_train_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64]),
    "nsr_render_train": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.POINTER(NsrRng), C.POINTER(NsrOutputs),
                                   C.c_void_p, C.c_size_t, C.c_void_p]),
    "nsr_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.POINTER(NsrRng), C.POINTER(NsrOutGrads),
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "nsr_lr_loss_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsr_generate_rays_ex": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(NsrRayGen), C.c_void_p, C.c_void_p]),
    "nsr_loss_epilogue": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(NsrLossTerms),
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsr_clip_coef": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p]),
    "nsr_adam_step": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_float, C.c_void_p]),
    "nsr_debug_train_layout": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "nsr_debug_relu_bits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "nsr_debug_pack_image": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nsr_debug_unpack_image": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nsr_debug_dx": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_int64, C.c_void_p]),
    "nsr_debug_dw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]),
    "nsr_debug_set_trace": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nsr_debug_set_flags": (C.c_int, [C.c_void_p, C.c_int]),
    "nsr_launch_count": (C.c_int64, [C.c_void_p]),
    # data-parallel gradient all-reduce over peer-mapped memory (nsr_comm.cu)
    "nsr_comm_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_void_p)]),
    "nsr_comm_destroy": (C.c_int, [C.c_void_p]),
    "nsr_comm_buffer": (C.c_void_p, [C.c_void_p]),
    "nsr_comm_buffer_floats": (C.c_int64, [C.c_void_p]),
    "nsr_comm_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nsr_comm_connect_ipc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "nsr_comm_connect_ptrs": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int]),
    "nsr_comm_allreduce_mean": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nsr_debug_frame_schedule": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64)]),
    "nsr_debug_frame_lr_in_kernel": (C.c_int, [C.c_void_p, C.c_int64, C.c_int]),
    "nsr_debug_kernel_clock": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  nerf_sr_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)           # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.nsr_abi_version() != ABI_VERSION:
        raise ImportError(f"libnsr_b200 ABI {lib.nsr_abi_version()} != binding ABI {ABI_VERSION}")
    _lib = lib
    return lib


class NsrError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{STATUS_NAMES.get(code, code)}: {msg}")
        self.code = code
