"""ctypes binding of the C-ABI kernel library (include/tcct_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  Importing this module without
the built `.so` raises; calling any op without a CUDA device raises."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libtcct_b200.so")

_T = {"p": ctypes.c_void_p, "i": ctypes.c_int, "l": ctypes.c_longlong, "f": ctypes.c_float, "d": ctypes.c_double}

# name -> argument signature (p pointer, i int, l long long, f float, d double); every entry returns int status.
SIGNATURES = {
    "tcct_pack_weights": "piip",
    "tcct_conv2d_nhwc": "pp l pp iiiiiii pp pi p",
    "tcct_gemm_px": "pp l pp lii pp i pi p",
    "tcct_gemm_tma": "pppp lii pp i pi p",
    "tcct_gemm_tma_gelu": "ppppp lii p",
    "tcct_gemm_tma_dgelu": "pppp lii p",
    "tcct_conv2d_tma": "pppp iiiii pi p",
    "tcct_conv2d_tma_slice": "pii pp pii i iiiii pi p",
    "tcct_wgrad": "pppp iiiiiii iii i p",
    "tcct_wgrad_slice": "pi ppp iiiiii iii i p",
    "tcct_gate_fuse_fwd": "pppp iiiiii p",
    "tcct_gate_fuse_bwd": "pppp iiiiii p",
    "tcct_conv2d_nhwc_slice": "pi p l pp iiiiiii p pi p",
    "tcct_wgrad_tma": "pppp iiiiii pp p",
    "tcct_wgrad_gemm_tma": "pppp liii pp p",
    "tcct_wgrad_tma_partial": "ppp iiiiii pp p",
    "tcct_wgrad_gemm_tma_partial": "ppp lii pp p",
    "tcct_wgrad_reduce_batch": "pi p",
    "tcct_stats_nhwc": "plipp",
    "tcct_bn_finalize": "pdppffpppipip",
    "tcct_bn_act2_fwd": "ppippiipli p",
    "tcct_bn_act2_fwd_bn": "ppippiipli p",
    "tcct_bn_act2_bwd": "ppip ppip i p p pp pppp li p",
    "tcct_maxpool2_fwd": "ppiiiip",
    "tcct_maxpool2_bwd": "pppiiiip",
    "tcct_dwconv3_fwd": "pppp iiii ii pp",
    "tcct_dwconv3_bwd": "pppppp iiii ii p",
    "tcct_layernorm_fwd": "ppppp lif p",
    "tcct_layernorm_bwd": "ppppppp li p",
    "tcct_metapool_fwd": "pppp iii p",
    "tcct_ln_metapool_fwd": "pppppp ppp iii f p",
    "tcct_ln_metapool_bwd": "pppppp pp ppppp iii p",
    "tcct_metapool_bwd": "ppp iii p",
    "tcct_resize_nhwc_fwd": "ppp iiiiiii f i p",
    "tcct_resize_nhwc_bwd": "pp iiiiiii f p",
    "tcct_resize_nchw_fwd": "pp iiiii p",
    "tcct_resize_nchw_bwd": "pp iiiii p",
    "tcct_l2norm32_fwd": "pplfp",
    "tcct_l2norm32_bwd": "ppplfp",
    "tcct_norm_add3_fwd": "pppp iiiiiii f p",
    "tcct_stem_conv_fwd": "pppp iiii pp",
    "tcct_stem_conv_wgrad": "pppp iiii p",
    "tcct_head_fwd": "pppp iii p",
    "tcct_head_bwd": "pppppp iii p",
    "tcct_onehot_to_index": "pp iii p",
    "tcct_index64_to_u8": "pplp",
    "tcct_dice_fwd": "pp iiii ppp p",
    "tcct_score_sums": "pp iiii p p",
    "tcct_dice_multi_fwd": "pppp pp p iiii p ppp p",
    "tcct_dice_multi_bwd": "pppp pp p iiii p pp pppp p",
    "tcct_dice_bwd": "pp iii pp f p i p",
    "tcct_prep_pair": "pp iiiiiiii pp p",
    "tcct_post_labels": "p iiiiiiii p p",
    "tcct_prep_augment": "ppp iiiiiiiiii pp p",
    "tcct_argmax_nchw": "pp iii p",
    "tcct_soft_argmax": "pp iii f p",
    "tcct_boundary_positions": "pp iiii f p",
    "tcct_label_counts": "pp iii p p",
    "tcct_sqnorm": "plpp",
    "tcct_adamw_step": "pppp l pp fffff f p",
    "tcct_scale_per_sample": "ppp li p",
    "tcct_breg_forward": "pppp pppp pp pp pp ppp i iiii ppp p",
    "tcct_breg_backward": "ppppppppppp i iiii ppppp pppppppppp p",
    "tcct_fpolar_forward": "pppp iiii pppp p",
    "tcct_fpolar_backward": "ppp iiii ppp p",
}
INT_FUNCS = ("tcct_pack_entry_size", "tcct_abi_version", "tcct_device_arch", "tcct_aug_params_size", "tcct_reduce_job_size")
# int f(int H, int W, int Cin, int Cout, int KH, int KW)
SHAPE_FUNCS = ("tcct_conv_tma_supported", "tcct_wgrad_tma_supported")
# int f(long long M, int K, int N)
GEMM_SHAPE_FUNCS = ("tcct_gemm_tma_supported", "tcct_wgrad_gemm_tma_supported")
# workspace-size queries returning long long
LL_FUNCS = {"tcct_wgrad_tma_ws_floats": "iiiii", "tcct_wgrad_gemm_tma_ws_floats": "lii", "tcct_breg_ws_floats": "iiii", "tcct_breg_bwd_ws_floats": "iiii", "tcct_fpolar_ws_words": "l",
            "tcct_fpolar_fws_bytes": "", "tcct_launch_count": "", "tcct_route_count": "i", "tcct_dice_multi_sums_doubles": "i"}
ROUTES = {"conv_tma": 0, "wgrad_tma": 1, "gemm_tma": 2, "wgrad_gemm_tma": 3}


def route_counts():
    """{route name: launches so far} of the tcgen05 + TMA kernels."""
    return {k: int(_lib.tcct_route_count(v)) for k, v in ROUTES.items()}


class BnSrc(ctypes.Structure):
    """tcct_bn_src of include/tcct_b200.h."""
    _fields_ = [("stats", ctypes.c_void_p), ("count", ctypes.c_double), ("gamma", ctypes.c_void_p), ("beta", ctypes.c_void_p),
                ("eps", ctypes.c_float), ("momentum", ctypes.c_float), ("running_mean", ctypes.c_void_p),
                ("running_var", ctypes.c_void_p), ("num_batches", ctypes.c_void_p), ("update_running", ctypes.c_int),
                ("coef", ctypes.c_void_p)]


class ReduceJob(ctypes.Structure):
    """tcct_reduce_job of include/tcct_b200.h."""
    _fields_ = [("ws", ctypes.c_void_p), ("dw", ctypes.c_void_p), ("kind", ctypes.c_int), ("nparts", ctypes.c_int),
                ("p0", ctypes.c_int), ("p1", ctypes.c_int), ("p2", ctypes.c_int), ("reserved", ctypes.c_int)]


class TcctError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "tcct_b200: %s is missing - build it with `python -m tcct_b200.build` (nvcc, sm_100a). "
            "There is no fallback path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.tcct_last_error.restype = ctypes.c_char_p
    lib.tcct_last_error.argtypes = []
    for name in INT_FUNCS:
        fn = getattr(lib, name)
        fn.restype = ctypes.c_int
        fn.argtypes = []
    for name in SHAPE_FUNCS:
        fn = getattr(lib, name)
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_int] * 6
    for name in GEMM_SHAPE_FUNCS:
        fn = getattr(lib, name)
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_longlong, ctypes.c_int, ctypes.c_int]
    for name, sig in LL_FUNCS.items():
        fn = getattr(lib, name)
        fn.restype = ctypes.c_longlong
        fn.argtypes = [_T[c] for c in sig]
    return lib


_lib = _load()
_bound = {}


def _bind(name):
    sig = SIGNATURES[name].replace(" ", "")
    fn = getattr(_lib, name)
    fn.restype = ctypes.c_int
    fn.argtypes = [_T[c] for c in sig]

    def call(*args):
        if len(args) != len(sig):
            raise TypeError("%s expects %d arguments, got %d" % (name, len(sig), len(args)))
        rc = fn(*args)
        if rc != 0:
            raise TcctError("%s failed (%d): %s" % (name, rc, _lib.tcct_last_error().decode()))
    call.__name__ = name
    return call


def __getattr__(name):
    key = name if name.startswith("tcct_") else "tcct_" + name
    if key in SIGNATURES:
        if key not in _bound:
            _bound[key] = _bind(key)
        return _bound[key]
    if key in INT_FUNCS or key in LL_FUNCS or key in SHAPE_FUNCS or key in GEMM_SHAPE_FUNCS:
        return getattr(_lib, key)
    raise AttributeError(name)


def exported_symbols():
    """Every symbol include/tcct_b200.h declares."""
    return sorted(list(SIGNATURES) + list(INT_FUNCS) + list(LL_FUNCS) + list(SHAPE_FUNCS) + list(GEMM_SHAPE_FUNCS) + ["tcct_last_error"])
