"""ctypes binding of libddp_b200.so (the C ABI declared in include/ddp_b200.h).

There is deliberately no fallback: if the shared library is missing or no sm_100 device is
present, calls fail loudly.
"""
import ctypes
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libddp_b200.so")

ABI_VERSION = 1
TASK_SEG, TASK_DEPTH = 0, 1
SCHEDULE_COSINE, SCHEDULE_LINEAR = 0, 1
DIFFUSION_DDIM, DIFFUSION_DDPM = 0, 1
GEMM_FP32, GEMM_TC_3XF16, GEMM_TC_F16 = 0, 1, 2
GEMM_MODES = {"fp32": GEMM_FP32, "tc_3xf16": GEMM_TC_3XF16, "tc_f16": GEMM_TC_F16}

TAP_HEAD_IN, TAP_VALUE, TAP_SAMPLING, TAP_GATHERED, TAP_LN1, TAP_LAYER_OUT, TAP_LOGITS, TAP_STATE, \
    TAP_TEMB, TAP_FILM = range(10)

EXPORTS = ["ddp_abi_version", "ddp_create", "ddp_destroy", "ddp_last_error", "ddp_weight_count",
           "ddp_weight_name", "ddp_set_weight", "ddp_commit_weights", "ddp_set_schedule",
           "ddp_get_schedule", "ddp_set_ddpm_schedule", "ddp_set_step_noise", "ddp_set_uncertainty_outputs", "ddp_plan", "ddp_sample", "ddp_sample_host", "ddp_sample_host_ex", "ddp_sample_host_submit", "ddp_sample_host_wait", "ddp_head_forward", "ddp_resize_argmax", "ddp_tail_probs", "ddp_probs_argmax", "ddp_add_tap",
           "ddp_set_state_override", "ddp_clear_debug", "ddp_last_launch_count", "ddp_profile_enable",
           "ddp_profile_collect", "ddp_kernel_class_name", "ddp_graph_replays", "ddp_graph_captures", "ddp_graph_last_fallback",
           "ddp_neck_create", "ddp_neck_destroy", "ddp_neck_last_error", "ddp_neck_weight_count", "ddp_neck_weight_name",
           "ddp_neck_set_weight", "ddp_neck_commit_weights", "ddp_neck_plan", "ddp_neck_forward",
           "ddp_neck_last_launch_count",
           "ddp_bev_create", "ddp_bev_destroy", "ddp_bev_last_error", "ddp_bev_weight_count", "ddp_bev_weight_name",
           "ddp_bev_set_weight", "ddp_bev_commit_weights", "ddp_bev_set_schedule", "ddp_bev_plan", "ddp_bev_sample",
           "ddp_bev_last_launch_count"]
NECK_STAGE_FPN, NECK_STAGE_MERGE = 1, 2
K_COUNT = 14


class DDPConfig(ctypes.Structure):
    _fields_ = [("abi_version", ctypes.c_int32), ("task", ctypes.c_int32), ("num_classes", ctypes.c_int32),
                ("timesteps", ctypes.c_int32), ("time_difference", ctypes.c_int32),
                ("noise_schedule", ctypes.c_int32), ("diffusion", ctypes.c_int32),
                ("accumulation", ctypes.c_int32), ("learned_sinusoidal_dim", ctypes.c_int32),
                ("num_layers", ctypes.c_int32), ("gemm_mode", ctypes.c_int32),
                ("sample_range_lo", ctypes.c_float), ("bit_scale", ctypes.c_float),
                ("min_depth", ctypes.c_float), ("max_depth", ctypes.c_float)]


class DDPNeckConfig(ctypes.Structure):
    _fields_ = [("abi_version", ctypes.c_int32), ("stages", ctypes.c_int32), ("num_levels", ctypes.c_int32),
                ("in_channels", ctypes.c_int32 * 4), ("out_channels", ctypes.c_int32), ("num_groups", ctypes.c_int32),
                ("eps", ctypes.c_float)]


class DDPBevConfig(ctypes.Structure):
    _fields_ = [("abi_version", ctypes.c_int32), ("timesteps", ctypes.c_int32), ("time_difference", ctypes.c_int32),
                ("noise_schedule", ctypes.c_int32), ("diffusion", ctypes.c_int32),
                ("learned_sinusoidal_dim", ctypes.c_int32), ("num_layers", ctypes.c_int32),
                ("feat_channels", ctypes.c_int32), ("gemm_mode", ctypes.c_int32), ("bit_scale", ctypes.c_float),
                ("threshold", ctypes.c_float)]


class DDPError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libddp_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load the shared library (once).  Raises ImportError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m ddp_b200.build` "
            "(nvcc, sm_100a).  ddp_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, cp, i32, i64 = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int64
    fp = ctypes.POINTER(ctypes.c_float)
    lib.ddp_abi_version.restype = i32
    lib.ddp_create.argtypes = [ctypes.POINTER(DDPConfig), ctypes.POINTER(vp)]
    lib.ddp_destroy.argtypes = [vp]
    lib.ddp_destroy.restype = None
    lib.ddp_last_error.argtypes = [vp]
    lib.ddp_last_error.restype = cp
    lib.ddp_weight_count.argtypes = [vp]
    lib.ddp_weight_name.argtypes = [vp, i32, ctypes.POINTER(i64)]
    lib.ddp_weight_name.restype = cp
    lib.ddp_set_weight.argtypes = [vp, cp, vp, i64]
    lib.ddp_commit_weights.argtypes = [vp]
    lib.ddp_set_schedule.argtypes = [vp, i32, fp, fp, fp, fp, fp]
    lib.ddp_get_schedule.argtypes = [vp, fp, fp, fp, fp, fp]
    lib.ddp_set_ddpm_schedule.argtypes = [vp, i32, fp, fp, fp, ctypes.POINTER(ctypes.c_int32)]
    lib.ddp_set_step_noise.argtypes = [vp, vp]
    lib.ddp_set_uncertainty_outputs.argtypes = [vp, vp, vp]
    lib.ddp_plan.argtypes = [vp, i32, i32, i32, i32, ctypes.POINTER(ctypes.c_size_t)]
    lib.ddp_sample.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
    lib.ddp_sample_host.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
    lib.ddp_sample_host_ex.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp, ctypes.c_size_t, vp]
    lib.ddp_sample_host_submit.argtypes = [vp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t, vp, ctypes.POINTER(i64)]
    lib.ddp_sample_host_wait.argtypes = [vp, i64]
    lib.ddp_head_forward.argtypes = [vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
    lib.ddp_resize_argmax.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.ddp_tail_probs.argtypes = [vp, vp] + [i32] * 13 + [vp, vp]
    lib.ddp_probs_argmax.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
    lib.ddp_add_tap.argtypes = [vp, i32, i32, i32, vp]
    lib.ddp_set_state_override.argtypes = [vp, i32, vp]
    lib.ddp_clear_debug.argtypes = [vp]
    lib.ddp_last_launch_count.argtypes = [vp]
    lib.ddp_last_launch_count.restype = i64
    lib.ddp_graph_replays.argtypes = [vp]
    lib.ddp_graph_replays.restype = i64
    lib.ddp_graph_captures.argtypes = [vp]
    lib.ddp_graph_captures.restype = i64
    lib.ddp_graph_last_fallback.argtypes = [vp]
    lib.ddp_graph_last_fallback.restype = cp
    lib.ddp_profile_enable.argtypes = [vp, i32]
    lib.ddp_profile_collect.argtypes = [vp, fp, ctypes.POINTER(i64), i32]
    lib.ddp_kernel_class_name.argtypes = [i32]
    lib.ddp_kernel_class_name.restype = cp
    i32p = ctypes.POINTER(ctypes.c_int32)
    lib.ddp_neck_create.argtypes = [ctypes.POINTER(DDPNeckConfig), ctypes.POINTER(vp)]
    lib.ddp_neck_destroy.argtypes = [vp]
    lib.ddp_neck_destroy.restype = None
    lib.ddp_neck_last_error.argtypes = [vp]
    lib.ddp_neck_last_error.restype = cp
    lib.ddp_neck_weight_count.argtypes = [vp]
    lib.ddp_neck_weight_name.argtypes = [vp, i32, ctypes.POINTER(i64)]
    lib.ddp_neck_weight_name.restype = cp
    lib.ddp_neck_set_weight.argtypes = [vp, cp, vp, i64]
    lib.ddp_neck_commit_weights.argtypes = [vp]
    lib.ddp_neck_plan.argtypes = [vp, i32, i32p, i32p, ctypes.POINTER(ctypes.c_size_t)]
    lib.ddp_neck_forward.argtypes = [vp, ctypes.POINTER(vp), vp, ctypes.POINTER(vp), vp, ctypes.c_size_t, vp]
    lib.ddp_neck_last_launch_count.argtypes = [vp]
    lib.ddp_neck_last_launch_count.restype = i64
    lib.ddp_bev_create.argtypes = [ctypes.POINTER(DDPBevConfig), ctypes.POINTER(vp)]
    lib.ddp_bev_destroy.argtypes = [vp]
    lib.ddp_bev_destroy.restype = None
    lib.ddp_bev_last_error.argtypes = [vp]
    lib.ddp_bev_last_error.restype = cp
    lib.ddp_bev_weight_count.argtypes = [vp]
    lib.ddp_bev_weight_name.argtypes = [vp, i32, ctypes.POINTER(i64)]
    lib.ddp_bev_weight_name.restype = cp
    lib.ddp_bev_set_weight.argtypes = [vp, cp, vp, i64]
    lib.ddp_bev_commit_weights.argtypes = [vp]
    lib.ddp_bev_set_schedule.argtypes = [vp, i32, fp, fp, fp, fp, fp]
    lib.ddp_bev_plan.argtypes = [vp, i32, i32, i32, i32, i32, i32, fp, fp, ctypes.POINTER(ctypes.c_size_t)]
    lib.ddp_bev_sample.argtypes = [vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
    lib.ddp_bev_last_launch_count.argtypes = [vp]
    lib.ddp_bev_last_launch_count.restype = i64
    if lib.ddp_abi_version() != ABI_VERSION:
        raise ImportError(f"libddp_b200.so ABI {lib.ddp_abi_version()} != binding {ABI_VERSION}; rebuild")
    _lib = lib
    return lib
