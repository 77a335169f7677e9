"""ctypes binding of libclimsim_b200.so (the C ABI declared in include/climsim_b200.h).

There is no CPU fallback: if the shared library is missing ``load()`` raises, and every engine entry point raises
``CsbError`` when the library reports a failure (e.g. CSB_ENODEV on a machine without an sm_100 GPU).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CSB_LIB_PATH") or os.path.join(HERE, "libclimsim_b200.so")    # override: A/B runs of two builds
MAX_LAYERS = 24

OK, EINVAL, ENODEV, ENOMEM, ECUDA, ESTATE, EUNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
ACT = {"none": 0, "linear": 0, "relu": 1, "elu": 2, "leakyrelu": 3}
DTYPE = {"fp32": 0, "f32": 0, "float32": 0, "bf16": 1, "bfloat16": 1, "tf32": 2, "tf32x3": 3}
LOSS = {"mse": 0, "mae": 1, "huber": 2}
OPT = {"adam_keras": 0, "adam": 0, "adam_torch": 1, "sgd": 2, "radam": 3, "rmsprop": 4}
FWD_NORMALIZE_IN, FWD_DENORM_OUT, FWD_KEEP_ACTIVATIONS, TRAIN_FUSED_OPT = 1, 2, 4, 8
BATCH_METRICS_SCRATCH = 2048
HSR_NO_OPT = 16
FWD_TRAINING = 32
IPC_HANDLE_BYTES = 64


class MlpCfg(C.Structure):
    _fields_ = [("in_dim", C.c_int32), ("n_layers", C.c_int32), ("units", C.c_int32 * MAX_LAYERS),
                ("act", C.c_int32 * MAX_LAYERS), ("alpha", C.c_float * MAX_LAYERS),
                ("layernorm", C.c_int32 * MAX_LAYERS), ("head_relu_from", C.c_int32), ("dtype", C.c_int32),
                ("loss", C.c_int32), ("max_batch", C.c_int64)]


class CnnCfg(C.Structure):
    _fields_ = [("depth", C.c_int32), ("width", C.c_int32), ("kernel", C.c_int32), ("in_ch", C.c_int32), ("out_ch", C.c_int32),
                ("out_lin", C.c_int32), ("levels", C.c_int32), ("act", C.c_int32), ("pre_out_act", C.c_int32), ("dtype", C.c_int32),
                ("loss", C.c_int32), ("max_batch", C.c_int64)]


class CsbError(RuntimeError):
    def __init__(self, code: int, where: str, detail: str):
        super().__init__(f"{where} failed: {detail} (code {code})")
        self.code = code


# every symbol the header declares: name -> (restype, argtypes)
_F, _P, _VP = C.POINTER(C.c_float), C.POINTER, C.c_void_p
SIGNATURES = {
    "csb_version": (C.c_int, []),
    "csb_strerror": (C.c_char_p, [C.c_int]),
    "csb_last_error": (C.c_char_p, []),
    "csb_device_info": (C.c_int, [_P(C.c_int), _P(C.c_int), _P(C.c_int), _P(C.c_size_t)]),
    "csb_mlp_create": (C.c_int, [_P(MlpCfg), _P(_VP)]),
    "csb_mlp_destroy": (C.c_int, [_VP]),
    "csb_mlp_param_count": (C.c_size_t, [_VP]),
    "csb_mlp_set_params": (C.c_int, [_VP, _VP]),
    "csb_mlp_get_params": (C.c_int, [_VP, _VP]),
    "csb_mlp_get_grads": (C.c_int, [_VP, _VP]),
    "csb_mlp_set_params_device": (C.c_int, [_VP, _VP, _VP]),
    "csb_mlp_get_params_device": (C.c_int, [_VP, _VP, _VP]),
    "csb_mlp_get_grads_device": (C.c_int, [_VP, _VP, _VP]),
    "csb_mlp_get_opt_state": (C.c_int, [_VP, _VP, _VP, _P(C.c_int64)]),
    "csb_mlp_set_opt_state": (C.c_int, [_VP, _VP, _VP, C.c_int64]),
    "csb_mlp_set_norm": (C.c_int, [_VP, _VP, _VP, _VP, _VP]),
    "csb_mlp_set_input_transform": (C.c_int, [_VP, _VP, _VP, _VP, _VP]),
    "csb_mlp_set_output_mask": (C.c_int, [_VP, _VP]),
    "csb_mlp_set_dropout": (C.c_int, [_VP, C.c_float, C.c_uint32]),
    "csb_mlp_debug_dropout_mask": (C.c_int, [_VP, C.c_int, _VP, C.c_int64, _VP]),
    "csb_mlp_forward": (C.c_int, [_VP, _VP, _VP, C.c_int64, C.c_uint32, _VP]),
    "csb_mlp_forward_host": (C.c_int, [_VP, _VP, _VP, C.c_int64, C.c_uint32, _VP]),
    "csb_mlp_backward": (C.c_int, [_VP, _VP, _VP, C.c_int64, _VP]),
    "csb_mlp_train_step": (C.c_int, [_VP, _VP, _VP, C.c_int64, C.c_float, C.c_uint32, _VP, _VP]),
    "csb_mlp_grad_buffer": (C.c_int, [_VP, _P(_VP), _P(C.c_size_t)]),
    "csb_mlp_apply_opt": (C.c_int, [_VP, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _VP]),
    "csb_mlp_train_step_host": (C.c_int, [_VP, _VP, _VP, C.c_int64, C.c_float, C.c_uint32, C.c_int, C.c_float,
                                          C.c_float, C.c_float, C.c_float, C.c_float, _P(C.c_float), _VP]),
    "csb_mlp_stage_host_batch": (C.c_int, [_VP, _VP, _VP, C.c_int64, _VP, _P(_VP), _P(_VP)]),
    "csb_mlp_release_staged": (C.c_int, [_VP, _VP]),
    "csb_mlp_train_step_host_async": (C.c_int, [_VP, _VP, _VP, C.c_int64, C.c_float, C.c_uint32, C.c_int, C.c_float,
                                                C.c_float, C.c_float, C.c_float, C.c_float, _VP, _VP]),
    "csb_mlp_launch_count": (C.c_int64, [_VP]),
    "csb_mlp_profile": (C.c_int, [_VP, C.c_int]),
    "csb_mlp_profile_read": (C.c_int, [_VP, _P(C.c_double), _P(C.c_int64), C.c_int]),
    "csb_profile_kind_count": (C.c_int, []),
    "csb_profile_kind_name": (C.c_char_p, [C.c_int]),
    "csb_cnn_create": (C.c_int, [_P(CnnCfg), _P(_VP)]),
    "csb_cnn_destroy": (C.c_int, [_VP]),
    "csb_cnn_param_count": (C.c_size_t, [_VP]),
    "csb_cnn_set_params": (C.c_int, [_VP, _VP]),
    "csb_cnn_get_params": (C.c_int, [_VP, _VP]),
    "csb_cnn_get_grads": (C.c_int, [_VP, _VP]),
    "csb_cnn_set_loss_weights": (C.c_int, [_VP, _VP]),
    "csb_cnn_set_dropout": (C.c_int, [_VP, C.c_float, C.c_uint32]),
    "csb_cnn_debug_read_hidden": (C.c_int, [_VP, C.c_int, C.c_int, _VP, C.c_int64, _VP]),
    "csb_cnn_forward": (C.c_int, [_VP, _VP, _VP, C.c_int64, _VP]),
    "csb_cnn_train_step": (C.c_int, [_VP, _VP, _VP, C.c_int64, C.c_float, _VP, _VP]),
    "csb_cnn_backward": (C.c_int, [_VP, _VP, _VP, C.c_int64, _VP]),
    "csb_cnn_set_params_device": (C.c_int, [_VP, _VP, _VP]),
    "csb_cnn_get_grads_device": (C.c_int, [_VP, _VP, _VP]),
    "csb_cnn_grad_buffer": (C.c_int, [_VP, _P(_VP), _P(C.c_size_t)]),
    "csb_cnn_apply_opt": (C.c_int, [_VP, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _VP]),
    "csb_cnn_launch_count": (C.c_int64, [_VP]),
    "csb_normalize": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int64, C.c_int32, _VP]),
    "csb_reshape_input_for_cnn": (C.c_int, [_VP, _VP, C.c_int64, _VP]),
    "csb_reshape_target_for_cnn": (C.c_int, [_VP, _VP, C.c_int64, _VP]),
    "csb_reshape_target_from_cnn": (C.c_int, [_VP, _VP, C.c_int64, _VP]),
    "csb_eval_metrics": (C.c_int, [_VP, _VP, _VP, C.c_int64, C.c_int32, _VP, _VP, C.c_double, _VP, _VP, C.c_double, C.c_double, C.c_double,
                                   C.c_int, _VP, _VP, _VP]),
    "csb_test_gemm_tn": (C.c_int, [_VP, _VP, _VP, C.c_int, C.c_int, C.c_int, C.c_int, _VP]),
    "csb_test_gemm_tn_tf32": (C.c_int, [_VP, _VP, _VP, C.c_int, C.c_int, C.c_int, C.c_int, _VP]),
    "csb_test_linear_fwd": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _VP]),
    "csb_test_set_debug": (None, [C.c_int]),
    "csb_test_set_stats": (None, [_VP]),
    "csb_eval_crps": (C.c_int, [_VP, _VP, C.c_int, C.c_int64, C.c_int, C.c_int, _VP, _VP, _VP]),
    "csb_batch_metrics": (C.c_int, [_VP, _VP, C.c_int64, C.c_int32, _VP, _VP, _VP]),
    "csb_hsr_train_step": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int64, C.c_int, C.c_uint32, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                     C.c_float, C.c_float, _VP, _VP, _VP]),
    "csb_mlp_dp_export": (C.c_int, [_VP, _VP]),
    "csb_mlp_dp_attach": (C.c_int, [_VP, C.c_int, C.c_int, _VP]),
    "csb_mlp_dp_debug": (C.c_int, [_VP, _VP]),
    "csb_mlp_dp_step": (C.c_int, [_VP, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _VP, _VP]),
    "csb_gather_rows": (C.c_int, [_VP, _VP, _VP, C.c_int64, C.c_int, C.c_int64, _VP]),
    "csb_gather_rows_check": (C.c_int, [_VP]),
    "csb_test_gemm_nt": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, C.c_int, C.c_int, C.c_int, _VP]),
    "csb_test_gemm_nt_cg": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), _VP]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (building is ``python -m climsim_b200.build`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: run `python -m climsim_b200.build` (needs nvcc). "
                          "climsim_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        if name.startswith("csb_test_") and os.environ.get("CSB_LIB_PATH") and not hasattr(lib, name):
            continue                     # an older build loaded for an A/B run may lack newer self-test hooks
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(code: int, where: str) -> None:
    if code != OK:
        lib = load()
        detail = lib.csb_last_error().decode() or lib.csb_strerror(code).decode()
        raise CsbError(code, where, detail)


def current_stream_ptr() -> int:
    import torch
    return int(torch.cuda.current_stream().cuda_stream)
