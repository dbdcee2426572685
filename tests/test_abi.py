"""CPU-side checks of the drop-in boundary: libmprg.so builds, loads, exports every symbol that
include/mprg.h declares, and refuses to create a context without a CUDA device (no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from make_prg_b200 import build

    path = build.build_library()
    return ctypes.CDLL(str(path))


def declared_symbols():
    text = (REPO / "include" / "mprg.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mprg_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/mprg.h but not exported"


def test_binding_covers_header():
    from make_prg_b200 import _lib

    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from make_prg_b200 import _lib, device

    with pytest.raises(_lib.MprgError):
        device.Context(0)
