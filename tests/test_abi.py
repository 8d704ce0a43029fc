"""CPU: the C-ABI library loads without a GPU and exports exactly the symbols include/tcct_b200.h declares; the
ctypes signature table of tcct_b200/_lib.py agrees with the header's parameter lists; pure host-side queries work."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tcct_b200.h")


def header_decls():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"(?:^|\n)\s*(const char\*|long long|int)\s+(tcct_\w+)\s*\(([^;]*?)\)\s*;", src):
        ret, name, params = m.group(1), m.group(2), " ".join(m.group(3).split())
        decls[name] = (ret, [] if params in ("void", "") else [p.strip() for p in params.split(",")])
    return decls


def code_of(param):
    if "*" in param:
        return "p"
    t = param.rsplit(" ", 1)[0].strip()
    return {"int": "i", "long long": "l", "float": "f", "double": "d"}[t.replace("const ", "")]


@pytest.fixture(scope="module")
def lib():
    from tcct_b200.build import build
    return ctypes.CDLL(build())


def test_header_symbols_are_exported(lib):
    decls = header_decls()
    assert len(decls) >= 50
    missing = [n for n in decls if not hasattr(lib, n)]
    assert not missing, missing


def test_exported_symbols_are_declared():
    from tcct_b200.build import LIB
    out = subprocess.run(["nm", "-D", "--defined-only", LIB], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("tcct_")}
    internal = {"tcct_set_error", "tcct_count_launch", "tcct_num_sms"}
    exported = {e for e in exported if not e.startswith("_Z") and e not in internal}
    decls = header_decls()
    assert exported - set(decls) == set(), sorted(exported - set(decls))


def test_ctypes_table_matches_header():
    import tcct_b200._lib as L
    decls = header_decls()
    assert sorted(L.exported_symbols()) == sorted(decls)
    for name, sig in L.SIGNATURES.items():
        ret, params = decls[name]
        assert ret == "int", name
        assert "".join(code_of(p) for p in params) == sig.replace(" ", ""), name
    for name, sig in L.LL_FUNCS.items():
        ret, params = decls[name]
        assert ret == "long long" and "".join(code_of(p) for p in params) == sig, name
    for name in L.GEMM_SHAPE_FUNCS:
        ret, params = decls[name]
        assert ret == "int" and "".join(code_of(p) for p in params) == "lii", name


def test_host_side_queries(lib):
    lib.tcct_abi_version.restype = ctypes.c_int
    assert lib.tcct_abi_version() >= 1
    lib.tcct_breg_ws_floats.restype = ctypes.c_longlong
    assert lib.tcct_breg_ws_floats(2, 5, 8, 4) == 6 * 64 + 4 * 2 * 4 + 8 + 16 * 32 + 2 * 16 * 4 + 16 * 32 // 4
    lib.tcct_conv_tma_supported.restype = ctypes.c_int
    assert lib.tcct_conv_tma_supported(64, 64, 32, 32, 3, 3) == 0          # lines shorter than 128 pixels
    assert lib.tcct_conv_tma_supported(256, 256, 64, 32, 3, 3) == 0
    lib.tcct_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.tcct_last_error(), bytes)


def test_ops_refuse_cpu_tensors():
    import torch
    from tcct_b200 import ops as O
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        O.MaxPool2Fn.apply(torch.zeros(1, 4, 4, 4))


def test_torch_custom_ops_are_registered():
    """The TORCH_LIBRARY shim (csrc_torch/torch_ops.cpp) over the C ABI loads without a GPU, registers its schemas under
    torch.ops.tcct_b200 and has NO CPU kernels (a CPU tensor is refused by the dispatcher: no fallback path)."""
    import torch
    from tcct_b200 import torch_ops
    torch_ops.build()
    ns = torch_ops.load()
    for name in ("conv2d_tma", "conv2d_wgrad_tma", "dice_multi_fwd", "dice_multi_bwd", "argmax_labels", "soft_argmax", "boundary_positions",
                 "score_sums", "route_count", "ln_metapool_fwd", "ln_metapool_bwd", "clip_adamw_step", "gate_fuse_fwd", "gate_fuse_bwd",
                 "prep_augment"):
        assert hasattr(ns, name), name
    assert "Tensor lab, float w_aux" in str(ns.dice_multi_fwd.default._schema)
    assert ns.route_count(0) >= 0
    with pytest.raises((NotImplementedError, RuntimeError)):
        ns.argmax_labels(torch.zeros(1, 2, 4, 4))
