"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports every
symbol include/textboost_b200.h declares; with no GPU every compute entry point fails loudly (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "textboost_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    from textboost_b200 import _cabi
    lib = ctypes.CDLL(built_lib)
    declared = header_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    # the ctypes binding covers exactly the declared surface
    assert sorted(_cabi.exported_symbols()) == declared
    assert _cabi.lib().tb_version() == 1


def test_both_precision_builds_export_the_same_surface(built_lib):
    """libtextboost_b200.so (fp16) and libtextboost_b200_bf16.so (-DTB_BF16) are the same ABI; tb_storage_dtype() tells
    them apart, and the loader refuses a library that does not match the process's precision policy."""
    from textboost_b200 import build, precision
    fp16, bf16 = ctypes.CDLL(build.VARIANTS["fp16"][0]), ctypes.CDLL(build.VARIANTS["bf16"][0])
    for name in header_symbols():
        assert hasattr(bf16, name), f"{name} missing from the bf16 build"
    assert fp16.tb_storage_dtype() == 0 and bf16.tb_storage_dtype() == 1
    assert precision.POLICY.name == os.environ.get("TEXTBOOST_B200_PRECISION", "fp16")
    with pytest.raises(NotImplementedError):
        precision.set_policy("no")
    with pytest.raises(ValueError):
        precision.set_policy("fp8")
    code = ("import os; os.environ['TB_LIB'] = %r\n"
            "from textboost_b200 import _cabi\n"
            "try:\n    _cabi.lib()\nexcept RuntimeError as e:\n    print('REFUSED', e)\n") % build.VARIANTS["bf16"][0]
    import subprocess
    import sys
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True,
                       env={k: v for k, v in os.environ.items() if k != "TEXTBOOST_B200_PRECISION"})
    assert "REFUSED" in r.stdout and "precision" in r.stdout, r.stdout + r.stderr


def test_header_is_plain_c(tmp_path):
    """The header compiles as C (no C++/torch types cross the boundary)."""
    import subprocess
    c = tmp_path / "t.c"
    c.write_text('#include "textboost_b200.h"\nint main(void){ tb_epilogue e; (void)e; return TB_ABI_VERSION - 1; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c),
                    "-o", str(tmp_path / "t.o")], check=True)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_calls_fail_loudly_without_gpu(built_lib):
    from textboost_b200 import _cabi
    with pytest.raises(RuntimeError, match="no CPU fallback|no fallback"):
        _cabi.call("tb_check_device")
    # a compute entry point: returns TB_E_ARCH before touching any pointer
    ep = _cabi.Epilogue()
    rc = _cabi.lib().tb_gemm_f16(None, 64, None, 64, None, 64, 128, 64, 64, ctypes.byref(ep), None)
    assert rc == -3 and "fallback" in _cabi.last_error()


def test_sass_uses_blackwell_tensor_and_tma_paths(built_lib):
    """tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG in the shipped SASS (B200_PROFILING.md)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", built_lib], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass and "UTMALDG" in sass
    assert "HMMA.16816" not in sass  # no legacy mma.sync path
    # the bf16 build: same tensor / TMA paths, bf16 conversions instead of fp16 ones in the epilogues and SIMT kernels
    from textboost_b200 import build
    sass16 = subprocess.run([cuobjdump, "-sass", build.VARIANTS["bf16"][0]], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass16 and "LDTM" in sass16 and "UTMALDG" in sass16 and "HMMA.16816" not in sass16
    assert "F2FP.BF16" in sass16 and "F2FP.BF16" not in sass and "F2FP.F16" in sass
