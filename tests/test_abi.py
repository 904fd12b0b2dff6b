"""The C-ABI shared library loads and exports every symbol include/mixdq_b200.h declares.
No compute calls (CPU-only)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from mixdq_b200 import build, _lib
    build.build()
    return _lib.load()


def declared_symbols():
    text = (ROOT / "include" / "mixdq_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mixdq_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(lib):
    from mixdq_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 14
    assert sorted(_lib.ABI_SYMBOLS) == syms
    for s in syms:
        assert hasattr(lib, s), s


def test_no_torch_or_python_dependency():
    """plain C ABI: the library must not link libtorch / libpython."""
    import subprocess
    from mixdq_b200 import _lib
    out = subprocess.run(["ldd", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    # library names only: the load addresses ldd prints are random hex and may contain "c10"
    names = [line.split()[0] for line in out.splitlines() if line.strip()]
    assert not any(("torch" in n) or ("python" in n) or ("c10" in n) for n in names), names


def test_version_and_strerror(lib):
    assert lib.mixdq_abi_version() == 2
    assert lib.mixdq_strerror(0) == b"success"
    # the reference's alignment message (qlinear.cc:130-133)
    assert lib.mixdq_strerror(-2) == \
        b"Int8 kernel with input or output alignment not to 4 is not supported."
    assert lib.mixdq_strerror(-4) == b"CUDA kernel failed"   # qlinear.cc:134
    assert lib.mixdq_quant_dynamic_ws_bytes() >= 8200


def test_argument_validation_without_gpu(lib):
    """Argument checks run before any CUDA call, so they are testable on a CPU box."""
    assert lib.mixdq_gemm_w8a8_f16(None, 0, None, None, None, None, None, 0, 4, 4, 4, None, None) == -1
    buf = ctypes.create_string_buffer(64)
    p = ctypes.addressof(buf)
    # K % 4 != 0 -> alignment error, as the reference
    assert lib.mixdq_gemm_w8a8_f16(p, 6, p, p, p, None, p, 8, 1, 8, 6, None, None) == -2
    assert lib.mixdq_conv_w8a8_f16(p, 6, p, p, None, p, p, None, p, 1, 4, 4, 6, 8, 1, 1, 1, 0,
                                   None, None) == -2
    # pad > 0 needs wsum_krs + zp
    assert lib.mixdq_conv_w8a8_f16(p, 8, p, p, None, p, None, None, p, 1, 4, 4, 8, 8, 3, 3, 1, 1,
                                   None, None) == -1
    assert lib.mixdq_quant_i8_static(p, -1, p, p, p, None) == -1
    assert lib.mixdq_quant_i8_static(None, 0, p, p, None, None) == 0   # empty input is a no-op
    assert lib.mixdq_gemm_w8a8_f16(None, 8, p, p, p, None, None, 8, 0, 8, 8, None, None) == 0


def test_ops_fail_loudly_without_library(monkeypatch, tmp_path):
    from mixdq_b200 import _lib
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setenv("MIXDQ_B200_LIB", str(tmp_path / "missing.so"))
    with pytest.raises(_lib.MixdqLibraryError):
        _lib.load()
