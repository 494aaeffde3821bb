"""ctypes binding of libgradtts_b200.so (the C ABI in include/gradtts_b200.h).

There is no fallback: if the library is missing or the call fails, a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgradtts_b200.so")

FLAG_FP32 = 1
FLAG_SDE = 2

_c = ctypes
_vp, _i, _sz, _d, _f, _cp, _l = _c.c_void_p, _c.c_int, _c.c_size_t, _c.c_double, _c.c_float, _c.c_char_p, _c.c_long

# name -> (restype, argtypes); must list every function declared in include/gradtts_b200.h
SIGNATURES = {
    "gtts_version": (_i, []),
    "gtts_last_error": (_cp, []),
    "gtts_sm100_device_count": (_i, []),
    "gtts_mas_workspace_bytes": (_sz, [_i, _i, _i]),
    "gtts_mas_maximum_path": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp, _vp]),
    "gtts_mas_maximum_path_c": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _sz, _vp, _vp]),
    "gtts_mas_maximum_path_host": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _i]),
    "gtts_decoder_create": (_i, [_c.POINTER(_vp), _i, _i, _i, _d, _d, _d, _i]),
    "gtts_decoder_destroy": (None, [_vp]),
    "gtts_decoder_set_param": (_i, [_vp, _cp, _vp, _sz]),
    "gtts_decoder_set_option": (_i, [_vp, _cp, _i]),
    "gtts_decoder_reverse_diffusion": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "gtts_decoder_estimator": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "gtts_decoder_estimator_vjp": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "gtts_decoder_estimator_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "gtts_decoder_get_param_grad": (_i, [_vp, _cp, _vp, _sz, _vp]),
    "gtts_decoder_param_grad_slot": (_i, [_vp, _cp, _c.POINTER(_sz), _c.POINTER(_sz)]),
    "gtts_decoder_get_param_grads_flat": (_i, [_vp, _vp, _sz, _vp]),
    "gtts_decoder_reverse_diffusion_host": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i]),
    "gtts_decoder_profile_step": (_i, [_vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "gtts_decoder_launches_last_call": (_l, [_vp]),
    "gtts_decoder_cache_info": (_i, [_vp, _vp, _i]),
    "gtts_encoder_create": (_i, [_c.POINTER(_vp), _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i]),
    "gtts_encoder_destroy": (None, [_vp]),
    "gtts_encoder_set_param": (_i, [_vp, _cp, _vp, _sz]),
    "gtts_encoder_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "gtts_encoder_check_tokens": (_i, [_vp, _vp]),
    "gtts_encoder_launches_last_call": (_l, [_vp]),
    "gtts_vocoder_create": (_i, [_c.POINTER(_vp), _i, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _i]),
    "gtts_vocoder_destroy": (None, [_vp]),
    "gtts_vocoder_set_param": (_i, [_vp, _cp, _vp, _sz]),
    "gtts_vocoder_set_option": (_i, [_vp, _cp, _c.c_longlong]),
    "gtts_vocoder_hop": (_i, [_vp]),
    "gtts_vocoder_forward": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "gtts_vocoder_forward_host": (_i, [_vp, _vp, _vp, _i, _i, _i]),
    "gtts_vocoder_profile": (_i, [_vp, _i, _i, _i, _vp, _sz, _vp]),
    "gtts_vocoder_launches_last_call": (_l, [_vp]),
    "gtts_vocoder_cache_info": (_i, [_vp, _vp, _i]),
    "gtts_align_log_prior": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "gtts_align_outputs": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "gtts_score_loss_workspace_bytes": (_sz, []),
    "gtts_forward_diffusion": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _d, _d, _vp]),
    "gtts_score_loss": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _vp, _i, _i, _i, _d, _d, _vp]),
    "gtts_test_conv_apply": (_i, [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "gtts_test_attn_xk": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "gtts_test_attn_fold": (_i, [_vp, _vp, _vp, _f, _vp, _i, _i, _i, _i, _vp]),
    "gtts_test_issue_microbench": (_i, [_i, _i, _i, _i, _i, _i, _vp, _vp]),
    "gtts_test_conv": (_i, [_i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
}

_lib = None


def load():
    """Load the shared library (building is the job of __graft_entry__.build / build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python grad-tts_b200/build.py` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().gtts_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"gradtts_b200: {what} failed (code {rc}): {last_error()}")


def require_cuda_tensor(t, name, dtype=None):
    import torch
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"gradtts_b200: {name} is on {t.device}; this path runs on sm_100a CUDA devices only "
                           "(no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must have dtype {dtype}, got {t.dtype}")
