"""pytest configuration: the `gpu` marker, import paths and shared helpers.

`-m "not gpu"` runs here without a GPU (oracle vs golden vectors / the compiled reference, the host
front-end, C-ABI symbol checks, gloo sharding); `-m gpu` holds the parity tests proper, which call
the CUDA path through the C ABI and compare it with the oracle bit for bit.
"""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

PKG = "fx8010-emulator-core_b200"

# The program translator compiles in a background thread by default and switches kernels when NVRTC is done: the suites
# written for the interpreter kernels pin it off so that the kernel under test does not depend on timing;
# tests/test_gpu_translate.py selects its modes explicitly (fx8010_gpu_set_option overrides the environment).
os.environ.setdefault("FX8010_TRANSLATE", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    pkg_dir = os.path.join(ROOT, PKG)
    if not (os.path.exists(os.path.join(pkg_dir, "libfx8010_gpu.so")) and os.path.exists(os.path.join(pkg_dir, "libfx8010_host.so"))):
        subprocess.run(["make", "-s", "-C", pkg_dir, "all"], check=True)
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"], check=True)


@pytest.fixture(scope="session")
def fx():
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle
    return pyoracle


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({4: np.uint32, 8: np.uint64}[a.dtype.itemsize])


def assert_bits_equal(a, b, what=""):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    ba, bb = bits(a), bits(b)
    if not np.array_equal(ba, bb):
        idx = np.argwhere(ba != bb)
        first = tuple(idx[0])
        raise AssertionError(f"{what}: {len(idx)} of {ba.size} elements differ; first at {first}: "
                             f"{a[first]!r} (0x{int(ba[first]):x}) vs {b[first]!r} (0x{int(bb[first]):x})")
