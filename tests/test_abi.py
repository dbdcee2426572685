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


def build_c_caller(tmp_path):
    """tests/c_abi/lanes.c: a plain C99 + pthreads caller of include/mprg.h (the header must be valid C)."""
    import subprocess

    from make_prg_b200 import build

    lib = build.build_library()
    exe = tmp_path / "lanes"
    subprocess.run(["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I", str(REPO / "include"),
                    str(REPO / "tests" / "c_abi" / "lanes.c"), "-o", str(exe), str(lib), "-lpthread",
                    f"-Wl,-rpath,{lib.parent}"], check=True)
    return exe


def test_c_caller_compiles_links_and_fails_loudly_without_device(tmp_path):
    import subprocess

    import torch

    exe = build_c_caller(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present (tests/test_gpu_host_api.py runs the program)")
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 77 and "MPRG_E_NO_DEVICE" in out.stdout
