"""CPU-side checks of the C-ABI boundary: the shared library loads without a GPU and exports every symbol that
include/oasr_b200.h declares; the ctypes table in olmoasr_b200/_lib.py covers exactly that set; argument validation
fails with error codes (no compute calls are made here)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "oasr_b200.h").read_text()
    return sorted(set(re.findall(r"OASR_API\s+(?:const\s+char\*|int)\s+(oasr_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from olmoasr_b200 import _lib

    names = _declared()
    assert len(names) >= 25
    h = _lib.lib()
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/oasr_b200.h but not exported by {_lib.lib_path()}"
    assert sorted(_lib.exported_symbols()) == names, "ctypes signature table and header disagree"
    assert h.oasr_abi_version() == 1


def test_header_cites_the_reference_interface():
    text = (ROOT / "include" / "oasr_b200.h").read_text()
    for cite in ("olmoasr/model.py", "train_timestamps.py", "whisper.audio.log_mel_spectrogram"):
        assert cite in text


def test_argument_validation_returns_error_codes_without_a_gpu():
    from olmoasr_b200 import _lib

    h = _lib.lib()
    rc = h.oasr_gemm_bf16(None, 0, 0, None, 0, 0, None, 0, None, None, None, 0, 0, 0, 0, 0, 1, 0, None)
    assert rc == -1 and b"empty problem" in h.oasr_last_error()
    rc = h.oasr_layernorm_fwd(None, None, None, None, None, None, 4, 12, ctypes.c_float(1e-5), None)
    assert rc == -1 and b"multiple of 8" in h.oasr_last_error()
    rc = h.oasr_attention_fwd(None, 0, None, 0, None, 0, None, 0, None, 1, 1, 1, 1, 80, 0, None, ctypes.c_float(1.0), None)
    assert rc == -1 and b"head_dim" in h.oasr_last_error()
    with pytest.raises(_lib.OasrError):
        _lib.check(rc, "attention")


def test_gemm_sm_budget_is_process_state_and_returns_the_previous_value():
    from olmoasr_b200 import kernels as K

    prev = K.set_gemm_sm_budget(132)
    assert K.set_gemm_sm_budget(0) == 132
    assert K.set_gemm_sm_budget(-5) == 0 and K.set_gemm_sm_budget(prev) == 0   # negative clamps to "all SMs"


def test_product_package_never_imports_the_oracle():
    for path in (ROOT / "olmoasr_b200").rglob("*.py"):
        src = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{path} imports the oracle"
