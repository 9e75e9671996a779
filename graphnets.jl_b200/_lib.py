"""ctypes binding of libgnb200.so - the C ABI declared in include/gnb200.h.

There is no CPU fallback: if the library is missing the import fails, and every compute entry
point fails with the library's own error when no B200 is present."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# GNB_LIB_VARIANT selects a test-only build of the library (build.py VARIANTS); the product path never sets it
_VARIANT = os.environ.get("GNB_LIB_VARIANT", "")
LIB_PATH = os.path.join(HERE, "libgnb200_%s.so" % _VARIANT if _VARIANT else "libgnb200.so")

GNB_OK, GNB_ERR_INVALID, GNB_ERR_CUDA, GNB_ERR_OOM, GNB_ERR_UNSUPPORTED, GNB_ERR_TIMEOUT = 0, -1, -2, -3, -4, -5
PREC_FP32, PREC_BF16, PREC_AUTO = 0, 2, 3
ADJ_F32, ADJ_U8, ADJ_I32, ADJ_BITS = 0, 1, 2, 3
LAYER_BLOCK, LAYER_CORE = 0, 1
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "auto": PREC_AUTO}

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)


class BlockParams(C.Structure):
    _fields_ = [("in_e", C.c_int32), ("in_n", C.c_int32), ("in_g", C.c_int32),
                ("out_e", C.c_int32), ("out_n", C.c_int32), ("out_g", C.c_int32),
                ("We", C.c_void_p), ("be", C.c_void_p), ("Wn", C.c_void_p), ("bn", C.c_void_p),
                ("Wg", C.c_void_p), ("bg", C.c_void_p)]


class FfnParams(C.Structure):
    _fields_ = [("W1", C.c_void_p), ("b1", C.c_void_p), ("W2", C.c_void_p), ("b2", C.c_void_p)]


class LnParams(C.Structure):
    _fields_ = [("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float), ("eps_mode", C.c_int32)]


class CoreParams(C.Structure):
    _fields_ = [("block", BlockParams), ("ffn", FfnParams * 3), ("ln1", LnParams * 3), ("ln2", LnParams * 3)]


class Layer(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("block", BlockParams), ("core", CoreParams)]


# name -> (restype, argtypes); must list every symbol include/gnb200.h declares
SIGNATURES = {
    "gnb_version": (C.c_int, []),
    "gnb_last_error": (C.c_char_p, []),
    "gnb_ctx_create": (C.c_void_p, [C.c_int, C.POINTER(C.c_int)]),
    "gnb_ctx_destroy": (C.c_int, [C.c_void_p]),
    "gnb_ctx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "gnb_sync": (C.c_int, [C.c_void_p]),
    "gnb_ctx_launch_count": (C.c_int64, [C.c_void_p]),
    "gnb_graph_lower": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, i32p, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(C.c_void_p)]),
    "gnb_graph_from_coo": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, i32p, i32p, C.c_int, C.c_int,
                                     C.POINTER(C.c_void_p)]),
    "gnb_graph_destroy": (C.c_int, [C.c_void_p]),
    "gnb_graph_counts": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), i32p, i32p]),
    "gnb_graph_export_host": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_void_p] * 7),
    "gnb_pad_edges": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "gnb_unpad_edges": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "gnb_pad_nodes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "gnb_unpad_nodes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "gnb_collapse_edges": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "gnb_model_create": (C.c_int, [C.c_void_p, C.POINTER(Layer), C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "gnb_model_destroy": (C.c_int, [C.c_void_p]),
    "gnb_model_out_dims": (C.c_int, [C.c_void_p, i32p, i32p, i32p]),
    "gnb_model_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_void_p] * 6 + [C.c_int]),
    "gnb_model_forward_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_void_p] * 6 + [C.c_int]),
    "gnb_block_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(BlockParams)] + [C.c_void_p] * 6 + [C.c_int]),
    "gnb_core_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(CoreParams)] + [C.c_void_p] * 6 + [C.c_int]),
    "gnb_corelist_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(CoreParams), C.c_int] + [C.c_void_p] * 6
                             + [C.c_int]),
}


SIGNATURES["gnb_logit_cross_entropy"] = (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p])
SIGNATURES["gnb_logit_cross_entropy_bwd"] = (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_void_p])


class LinSrc(C.Structure):
    _fields_ = [("x", C.c_void_p), ("d", C.c_int), ("ldx", C.c_int), ("W", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("eps", C.c_float), ("eps_mode", C.c_int)]


class LinAdd(C.Structure):
    _fields_ = [("a", C.c_void_p), ("idx", C.c_void_p), ("lda", C.c_int)]


class LinArgs(C.Structure):
    _fields_ = [("R", C.c_int64), ("Nout", C.c_int), ("ldw", C.c_int), ("nsrc", C.c_int), ("src", LinSrc * 3), ("bias", C.c_void_p),
                ("nadd", C.c_int), ("add", LinAdd * 4), ("relu", C.c_int), ("out", C.c_void_p), ("ldo", C.c_int), ("precision", C.c_int)]


_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
SIGNATURES["gnb_op_linear"] = (C.c_int, [_vp, C.POINTER(LinArgs)])
SIGNATURES["gnb_op_segsum"] = (C.c_int, [_vp, _vp, _i, _vp, _i64, _vp, _vp])
SIGNATURES["gnb_op_layernorm"] = (C.c_int, [_vp, _vp, _i64, _i, _vp, _vp, _f, _i, _vp])
SIGNATURES["gnb_op_layernorm_bwd"] = (C.c_int, [_vp, _vp, _vp, _i64, _i, _vp, _f, _i, _vp, _vp])
SIGNATURES["gnb_op_wgrad"] = (C.c_int, [_vp, _vp, _i, _i, _vp, _vp, _i, _i, _i64, _vp, _i, _i])
SIGNATURES["gnb_op_colsum"] = (C.c_int, [_vp, _vp, _i, _i, _i64, _vp])
SIGNATURES["gnb_op_relu_mask"] = (C.c_int, [_vp, _vp, _vp, _i64])
SIGNATURES["gnb_op_gather_add"] = (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i])
SIGNATURES["gnb_op_transpose"] = (C.c_int, [_vp, _vp, _i, _i, _i, _vp])
SIGNATURES["gnb_op_adamw"] = (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _f, _i])
SIGNATURES["gnb_graph_device_index"] = (C.c_int, [_vp] + [C.POINTER(C.c_void_p)] * 7)


class ProfEntry(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_int64), ("ms", C.c_double),
                ("alg_bytes", C.c_double), ("alg_flops", C.c_double)]


SIGNATURES["gnb_ctx_set_profiling"] = (C.c_int, [C.c_void_p, C.c_int])
SIGNATURES["gnb_debug_tc_timing"] = (C.c_int, [C.POINTER(C.c_ulonglong), C.c_int])
SIGNATURES["gnb_ctx_profile_read"] = (C.c_int, [C.c_void_p, C.POINTER(ProfEntry), C.c_int, C.POINTER(C.c_int)])


class GnbError(RuntimeError):
    pass


class GnbTimeout(GnbError):
    """GNB_ERR_TIMEOUT: a kernel watchdog fired; the context is still usable, that forward's results are not."""


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libgnb200.so is not built (%s). Run `python graphnets.jl_b200/build.py` "
            "(or __graft_entry__.build()); there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc):
    """Map ABI error codes onto the reference's error behaviour: invalid input is an
    AssertionError (the reference uses @assert everywhere, src/checks.jl)."""
    if rc == GNB_OK:
        return
    msg = (lib.gnb_last_error() or b"").decode()
    if rc == GNB_ERR_INVALID:
        raise AssertionError(msg)
    if rc == GNB_ERR_OOM:
        raise MemoryError(msg)
    if rc == GNB_ERR_TIMEOUT:
        raise GnbTimeout("gnb200 error %d: %s" % (rc, msg))
    raise GnbError("gnb200 error %d: %s" % (rc, msg))
