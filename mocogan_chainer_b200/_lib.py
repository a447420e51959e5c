"""ctypes binding of libmcg.so (include/mcg.h).  There is no fallback: if the shared object is missing or a call
fails, an exception is raised."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCG_LIB") or os.path.join(_HERE, "libmcg.so")   # MCG_LIB: A/B builds of the same ABI

F32, BF16, U8 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3
IMPL_SIMT, IMPL_TC = 0, 1


class ConvGeom(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "N", "Cin", "Cout", "Ti", "Hi", "Wi", "To", "Ho", "Wo", "kT", "kH", "kW", "sT", "sH", "sW", "pT", "pH", "pW")]

    def key(self):
        return tuple(getattr(self, n) for n, _ in self._fields_)


class McgError(RuntimeError):
    pass


_p, _i, _ll, _f, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t
_G = C.POINTER(ConvGeom)

# name -> (restype, argtypes); must list every symbol include/mcg.h declares (tests/test_abi.py checks this)
SIGNATURES = {
    "mcg_version": (_i, []),
    "mcg_last_error": (C.c_char_p, []),
    "mcg_launch_count": (_ll, []),
    "mcg_conv_workspace_bytes": (_sz, [_G, _i]),
    "mcg_conv_fprop": (_i, [_G, _p, _p, _p, _p, _i, _i, _i, _p, _sz, _p]),
    "mcg_conv_dgrad": (_i, [_G, _p, _p, _p, _p, _i, _i, _i, _i, _p, _sz, _p]),
    "mcg_conv_wgrad": (_i, [_G, _p, _p, _p, _i, _i, _p, _sz, _p]),
    "mcg_colreduce_workspace_bytes": (_sz, [_ll, _i]),
    "mcg_bn_stats": (_i, [_p, _ll, _i, _i, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "mcg_colsum": (_i, [_p, _ll, _i, _i, _p, _i, _p, _sz, _p]),
    "mcg_affine_act_noise": (_i, [_p, _ll, _i, _ll, _i, _p, _p, _i, _f, _f, _p, _ll, _ll, _ll, _p, _i, _p, _i, _p]),
    "mcg_pack_video": (_i, [_p, _i, _i, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _ll, _p, _f, _p, _ll, _ll, _ll, _p, _i, _p,
                            _i, _p]),
    "mcg_act_bn_bwd_reduce": (_i, [_p, _p, _ll, _i, _i, _p, _p, _p, _p, _i, _f, _p, _p, _p, _p, _p, _sz, _p]),
    "mcg_act_bn_bwd_apply": (_i, [_p, _p, _ll, _i, _i, _p, _p, _p, _p, _p, _i, _f, _i, _p, _p, _p, _i, _p]),
    "mcg_tanh_bwd_video": (_i, [_p, _p, _i, _p, _i, _i, _i, _i, _i, _p, _p, _i, _p]),
    "mcg_video_to_uint8": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _i, _p]),
    "mcg_gru_forward": (_i, [C.POINTER(_p), _p, _i, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "mcg_gru_backward": (_i, [C.POINTER(_p), C.POINTER(_p), _p, _i, _p, _p, _p, _i, _i, _i, _i, _p]),
    "mcg_loss_dis": (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "mcg_loss_gen": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "mcg_adam_step": (_i, [_p, _p, _p, _p, _p, _ll, _f, _f, _f, _f, _f, _f, _p, _p]),
    "mcg_cast_f32_to_bf16": (_i, [_p, _p, _ll, _p]),
    "mcg_cast_bf16_to_f32": (_i, [_p, _p, _ll, _p]),
    "mcg_fill_zero": (_i, [_p, _sz, _p]),
    "mcg_pad_channels": (_i, [_p, _p, _ll, _i, _i, _i, _p]),
    "mcg_step_state_init": (_i, [_p, C.c_ulonglong, _p]),
    "mcg_step_advance": (_i, [_p, _i, _p]),
    "mcg_randn": (_i, [_p, _ll, _f, _p, _i, _p]),
    "mcg_randint": (_i, [_p, _ll, _i, _p, _i, _p]),
    "mcg_int_add": (_i, [_p, _i, _p]),
    "mcg_tc_error_flag": (_i, [_i]),
    "mcg_set_tc_sm_limit": (_i, [_i]),
    "mcg_get_tc_sm_limit": (_i, []),
}

_lib = None


def load():
    """Loads libmcg.so (once).  Raises McgError if it has not been built: there is no CPU/torch fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise McgError("libmcg.so is missing at %s — run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(the CUDA extension is mandatory; there is no fallback path)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().mcg_last_error().decode("utf-8", "replace")
        raise McgError("%s failed (rc=%d): %s" % (what or "libmcg call", rc, msg))
