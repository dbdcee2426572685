"""CPU check of the DEVICE KMeans arithmetic: make_prg_b200/csrc/kmeans.cu's device functions are
compiled as plain C++ through tests/hostemu/cuda_shim.h (one "thread") and must reproduce, label for
label and with bit-identical inertia, what scikit-learn returned inside the unmodified reference
(tests/golden/kmeans_cases.npz).  This pins the operation order (numpy einsum / OpenBLAS dgemm, dgemv,
ddot rounding models, sequential Elkan sums) without a GPU; the -m gpu tests run the same cases on
the real kernel.  Test infrastructure only -- libmprg.so has no CPU path."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

from helpers import kmeans_cases

HERE = Path(__file__).resolve().parent / "hostemu"


@pytest.fixture(scope="module")
def emu():
    lib_path = HERE / "libkmemu.so"
    src = [HERE / "kmeans_host.cpp", HERE / "cuda_shim.h", HERE.parent.parent / "make_prg_b200" / "csrc" / "kmeans.cu"]
    if not lib_path.exists() or any(s.stat().st_mtime > lib_path.stat().st_mtime for s in src):
        subprocess.run(["g++", "-O1", "-ffp-contract=off", "-mfma", "-shared", "-fPIC", "-o", str(lib_path),
                        str(HERE / "kmeans_host.cpp")], check=True)
    lib = ctypes.CDLL(str(lib_path))
    rand = np.random.RandomState(2).random_sample(400)

    def run(X, K):
        X = np.ascontiguousarray(X, np.float64)
        labels = np.zeros(X.shape[0], np.int32)
        inertia = ctypes.c_double()
        lib.emu_kmeans(rand.ctypes.data_as(ctypes.c_void_p), X.ctypes.data_as(ctypes.c_void_p),
                       X.shape[0], X.shape[1], K, labels.ctypes.data_as(ctypes.c_void_p),
                       ctypes.byref(inertia))
        return labels, inertia.value

    return run


def test_device_arithmetic_reproduces_sklearn_on_reference_problems(emu):
    n = 0
    for X, K, labels, inertia in kmeans_cases():
        got, got_inertia = emu(X, K)
        assert np.array_equal(got, labels), (X.shape, K)
        assert got_inertia == inertia, (X.shape, K)
        n += 1
    assert n > 500


def test_tie_case_from_alsB(emu):
    """count matrix == I3, K = 2 (from amira alsB): all pairings tie; scikit-learn pairs rows 0 and 1
    because of how OpenBLAS dgemv rounds the first centre's self-distance."""
    labels, inertia = emu(np.eye(3), 2)
    assert labels[0] == labels[1] != labels[2]
    assert inertia == 1.0
