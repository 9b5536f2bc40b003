"""The drop-in boundary: libplba.so loads without a GPU, exports every entry point include/plba.h declares, the ctypes
mirrors match the C layouts, and the library refuses to work (loudly) when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from pl_slam_plucker_b200 import _lib, abi

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HDR = os.path.join(ROOT, "include", "plba.h")


def _declared_functions():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|void|char)\s*\*?\s*(plba_\w+)\s*\(", src, flags=re.M)
    return sorted(set(n for n in names if not n.endswith("_fn")))


def test_header_compiles_as_plain_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "plba.h"\n#include <stdio.h>\nint main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(plba_problem), sizeof(plba_options), '
                 'sizeof(plba_trace_rec), sizeof(plba_result), sizeof(plba_timing), sizeof(plba_scene_spec)); return 0;}\n')
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    mirrors = [abi.plba_problem, abi.plba_options, abi.plba_trace_rec, abi.plba_result, abi.plba_timing, abi.plba_scene_spec]
    assert sizes == [C.sizeof(m) for m in mirrors]


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    decl = _declared_functions()
    assert len(decl) >= 20
    for name in decl:
        assert hasattr(L, name), name
    assert set(_lib.EXPORTS) == set(decl)
    assert L.plba_version() == 1


def test_defaults_are_the_reference_config():
    o = abi.plba_options()
    _lib.load().plba_default_options(abi.PROFILE_G, C.byref(o))
    d = abi.Options(abi.PROFILE_G)
    for f, _ in abi.plba_options._fields_:
        assert getattr(o, f) == getattr(d.c, f), f
    assert (o.lambda_lba_lm, o.lambda_lba_k, o.max_iters_lba) == (1e-5, 10.0, 15)       # src/slamConfig.cpp:65-67
    assert (o.iters_stage1, o.iters_stage2, o.chi2_gate) == (5, 10, 5.991)              # src/mapHandler.cpp:6122,6152,6129


def test_no_cpu_fallback_without_cuda():
    """On a box without a GPU the product path must fail loudly (never route to the oracle or the emulation)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from pl_slam_plucker_b200 import solver
    with pytest.raises(solver.LBAError):
        solver.LBASolver(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pl_slam_plucker_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("the CPU oracle", "").replace("CPU oracle", "") or "import oracle" not in txt, f
                assert "from oracle" not in txt and "import oracle" not in txt and "libplba_oracle" not in txt and "emu_lib" not in txt, f
