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
