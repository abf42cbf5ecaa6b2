"""CPU: the C-ABI library builds, loads, and exports every symbol include/svlm_b200.h declares;
compute entry points refuse to run without an sm_100 device (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from streaming_vlm_b200 import _abi


@pytest.fixture(scope="module")
def lib():
    _abi.build()
    return _abi.lib()


def test_exports_match_header(lib):
    header = open(_abi.HEADER).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(svlm_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_abi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert lib.svlm_version() == 1


def test_sass_is_blackwell_native():
    import subprocess
    sass = subprocess.run(["cuobjdump", "-sass", _abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "STTM"):
        assert mnemonic in sass, f"{mnemonic} missing: the attention/GEMM kernels must use tcgen05 + TMA"
    import re
    assert re.search(r"(?<!UTC)HMMA\.", sass) is None, "legacy mma.sync found"
    assert "UTCHMMA.2CTA" in sass, "the CTA-pair GEMM must issue tcgen05.mma.cta_group::2"


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback(lib):
    arena = _abi.Arena(ctypes.c_void_p(1024), 1, 1, 128, 128)
    rc = lib.svlm_kv_compact_rope(ctypes.byref(arena), 0, None, 0, None, 0, None, None, 0, None, None)
    assert rc == -3          # SVLM_ENODEV
    assert b"no CUDA device" in lib.svlm_last_error() or b"CUDA" in lib.svlm_last_error()
    with pytest.raises(RuntimeError):
        _abi.gemm_bf16(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16), None,
                       torch.zeros(8, 8, dtype=torch.bfloat16))
