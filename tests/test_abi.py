"""CPU checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol
include/vtc_b200.h declares, and rejects bad arguments without touching a GPU."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "vtc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vtc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_the_binding_binds():
    from vtc_b200 import _ffi

    declared = _declared_symbols()
    assert declared, "no declarations parsed"
    assert sorted(_ffi.SIGNATURES) == declared


def test_library_exports_every_declared_symbol(lib):
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} not exported by libvtc_b200.so"
    assert lib.vtc_abi_version() == 3
    assert lib.vtc_strerror(0) == b"ok"
    assert b"workspace" in lib.vtc_strerror(-3)


def test_argument_counts_match_header(lib):
    from vtc_b200 import _ffi

    text = open(os.path.join(ROOT, "include", "vtc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name, (_, argtypes) in _ffi.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, text, flags=re.S)
        assert m, name
        args = m.group(1).strip()
        n = 0 if args in ("void", "") else args.count(",") + 1
        assert n == len(argtypes), f"{name}: header has {n} parameters, binding {len(argtypes)}"


def test_workspace_sizes_are_sane(lib):
    from vtc_b200 import _ffi

    small = lib.vtc_workspace_bytes(_ffi.OP_SIM_RANK, 1000, 1000, 512, _ffi.PREC_BF16)
    big = lib.vtc_workspace_bytes(_ffi.OP_SIM_RANK, 100000, 100000, 512, _ffi.PREC_BF16)
    exact = lib.vtc_workspace_bytes(_ffi.OP_SIM_RANK, 100000, 100000, 512, _ffi.PREC_EXACT)
    assert 0 < small < big < exact < 2 * 1024 ** 3
    # operands dominate: 2 * 100k * 512 * 2 B (bf16) and 3x that for the split
    assert big > 2 * 100000 * 512 * 2
    assert exact > 2 * 100000 * 1536 * 2
    assert lib.vtc_workspace_bytes(99, 1, 1, 1, 0) == 0
    assert lib.vtc_workspace_bytes(_ffi.OP_SIM_RANK, -1, 1, 1, 0) == 0


def test_invalid_arguments_are_rejected_before_any_launch(lib):
    from vtc_b200 import _ffi

    before = lib.vtc_launch_count()
    assert lib.vtc_row_norms(None, 4, 8, 8, _ffi.F32, None, None, None) == -1
    assert lib.vtc_sim_rank(None, None, 4, 4, 8, _ffi.F32, None, 0, 0, _ffi.METRIC_L2,
                            _ffi.PREC_EXACT, None, None, 0, None, None, 0, None) == -1
    # zero queries is a valid no-op even with NULL pointers
    assert lib.vtc_sim_rank(None, None, 0, 4, 8, _ffi.F32, None, 0, 0, _ffi.METRIC_L2,
                            _ffi.PREC_EXACT, None, None, 0, None, None, 0, None) == 0
    assert lib.vtc_topk_merge(None, None, 2, 4, 3, None, None, None) == -1
    # prepared ranking: outputs / prepared inputs are required, and the rows must be the canonical
    # values of the mode (bf16 rows <-> bf16 mode), VTC_ERR_UNSUPPORTED_SHAPE otherwise
    import ctypes
    buf = (ctypes.c_double * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.vtc_rank_prepare(None, 4, 8, _ffi.F32, _ffi.PREC_EXACT, p, None, None) == -1
    assert lib.vtc_rank_prepare(p, 4, 8, _ffi.F32, _ffi.PREC_EXACT, None, None, None) == -1
    assert lib.vtc_rank_prepare(p, 4, 8, _ffi.F32, _ffi.PREC_BF16, p, None, None) == -2
    assert lib.vtc_rank_prepare(p, 4, 8, _ffi.BF16, _ffi.PREC_EXACT, p, None, None) == -2
    # chunked ranking: each cached quantity is handed in or asked for (never neither), and the
    # brute-force mode has nothing to cache
    assert lib.vtc_sim_rank_prepared(p, p, 4, 4, 8, _ffi.F32, None, 0, 0, _ffi.METRIC_L2,
                                     _ffi.PREC_EXACT, p, None, None, None, p, None, 0, p, None, 0,
                                     None) == -1
    assert lib.vtc_sim_rank_prepared(p, p, 4, 4, 8, _ffi.F32, None, 0, 0, _ffi.METRIC_L2,
                                     _ffi.PREC_EXACT, None, None, p, None, p, None, 0, p, None, 0,
                                     None) == -1
    assert lib.vtc_sim_rank_prepared(p, p, 4, 4, 8, _ffi.F32, None, 0, 0, _ffi.METRIC_L2,
                                     _ffi.PREC_BRUTE, p, None, p, None, p, None, 0, p, None, 0,
                                     None) == -2
    assert lib.vtc_rank_eval(None, None, 4, 4, 8, _ffi.F32, None, _ffi.METRIC_L2, _ffi.PREC_EXACT,
                             None, 0, None, None, None, None, None, 0, None) == -1
    assert lib.vtc_rank_eval(p, p, 4, 4, 8, _ffi.F32, None, _ffi.METRIC_L2, _ffi.PREC_EXACT, None, 9,
                             p, p, None, None, None, 0, None) == -1
    assert lib.vtc_cam_attn_core(None, 6, 4, 512, 8, None, None) == -1
    # round-2 entry points: the one-call CAM backward, clip_loss on a materialised sim, the backward's
    # precision argument
    assert lib.vtc_cam_backward(None, None, None, None, None, None, 6, 4, 512, 8, 2, None, 0, None, 0,
                                1.0, None, None, _ffi.PREC_EXACT, None, None, None, None, 0, None) == -1
    assert lib.vtc_cam_backward_workspace_bytes(6, 256, 512, _ffi.PREC_BF16) > 0
    assert lib.vtc_cam_backward_workspace_bytes(0, 256, 512, _ffi.PREC_BF16) == 0
    assert lib.vtc_infonce_dense_fwd(None, 4, 4, None, None, None, None, None) == -1
    assert lib.vtc_infonce_dense_bwd(p, 4, 2, p, p, p, p, 4, None) == -1      # ld < n
    assert lib.vtc_infonce_bwd(p, p, 4, 8, _ffi.F32, 7, p, p, p, p, p, p, p, None, 0, None) == -1
    # the launch trace is off by default: ending one that was never begun reports zero launches
    buf8 = ctypes.create_string_buffer(64)
    assert lib.vtc_trace_end(buf8, 64) == 0
    assert lib.vtc_launch_count() == before


def test_product_path_fails_loudly_without_cuda():
    """No CPU fallback: CPU tensors are refused."""
    import numpy as np
    import torch

    from vtc_b200 import VtcError, ops
    from vtc_b200.model import RecallAtK, clip_loss

    x = torch.randn(4, 8)
    with pytest.raises(VtcError):
        ops.normalize(x)
    with pytest.raises(VtcError):
        ops.sim_rank(x, x)
    with pytest.raises(VtcError):
        clip_loss((x, x, x @ x.t()), {})
    if not torch.cuda.is_available():
        with pytest.raises(VtcError):
            RecallAtK("a", "b", [1]).compute(np.zeros((4, 8), np.float32), np.zeros((4, 8), np.float32))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under vtc_b200/ or scripts/ may import or call
    it; the only users are tests/, __graft_entry__.smoke() and the CPU-baseline leg of bench.py."""
    for top in ("vtc_b200", "scripts"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".sh")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                    assert "vtc_oracle" not in src.replace("oracle/vtc_oracle", ""), f
    bench = open(os.path.join(ROOT, "bench.py")).read()
    uses = [m.start() for m in re.finditer(r"^\s*(from|import)\s+oracle\b", bench, flags=re.M)]
    assert len(uses) == 1  # inside cpu_reference_pairs_per_s only
    fn = bench.index("def cpu_reference_pairs_per_s")
    assert fn < uses[0] < bench.index("def run_reference")


def test_header_is_plain_c_and_a_c_program_links_against_the_library(tmp_path):
    """The boundary is a C ABI: include/vtc_b200.h must compile as C11 (no torch / C++ types) and a
    plain C translation unit must link against libvtc_b200.so and call it."""
    import shutil
    import subprocess

    from vtc_b200 import build as vbuild

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    inc = os.path.join(root, "include")
    subprocess.run([gcc, "-std=c11", "-Wall", "-Werror", "-fsyntax-only", "-x", "c",
                    os.path.join(inc, "vtc_b200.h")], check=True)
    lib = vbuild.build()
    src = tmp_path / "abi.c"
    src.write_text(r"""
#include <stdio.h>
#include "vtc_b200.h"
int main(void) {
  size_t ws = vtc_workspace_bytes(VTC_OP_SIM_RANK, 1000, 1000, 512, VTC_PREC_BF16);
  printf("%d %zu %s\n", vtc_abi_version(), ws, vtc_strerror(VTC_ERR_INVALID_ARG));
  /* argument validation happens before any CUDA call: no GPU needed */
  return vtc_sim_rank(0, 0, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0) == VTC_ERR_INVALID_ARG ? 0 : 1;
}
""")
    exe = tmp_path / "abi"
    subprocess.run([gcc, "-std=c11", "-I", inc, str(src), "-o", str(exe), lib,
                    "-Wl,-rpath," + os.path.dirname(lib)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    ver, ws, msg = r.stdout.split(None, 2)
    assert int(ver) >= 1 and int(ws) > 0 and "invalid" in msg


def test_integration_md_ctypes_stub_matches_the_library(lib):
    """INTEGRATION.md shows the ctypes binding a maintainer of the reference would write for the
    evaluation path (the replacement of model/metric.py:137-161): its argtypes must have the arity
    and kinds of the real entry points, and the enum literals it passes must be the header's."""
    import ctypes

    from vtc_b200 import _ffi

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    block = text[text.index("import ctypes, torch"):]
    block = block[:block.index("```")]
    ns = {}
    stub = "\n".join(ln for ln in block.splitlines()
                     if ln.startswith(("P, i64", "lib.vtc_")) and "CDLL" not in ln)

    class FakeFn:
        pass

    class FakeLib:
        def __getattr__(self, name):
            fn = FakeFn()
            object.__setattr__(self, name, fn)
            return fn

    fake = FakeLib()
    exec("import ctypes\n" + stub, {"lib": fake, "ctypes": ctypes}, ns)
    for name in ("vtc_workspace_bytes", "vtc_sim_rank", "vtc_rank_finalize"):
        want = _ffi.SIGNATURES[name][1]
        got = getattr(fake, name).argtypes
        assert len(got) == len(want), name
        for g, w in zip(got, want):
            assert ctypes.sizeof(g) == ctypes.sizeof(w), (name, g, w)
    # the literals in the example call: F32 = 0, L2 = 1, EXACT = 0, OP_SIM_RANK = 0
    assert (_ffi.F32, _ffi.METRIC_L2, _ffi.PREC_EXACT, _ffi.OP_SIM_RANK) == (0, 1, 0, 0)
    assert "N, M, D, 0, None, 0, 0, 1, 0," in block
