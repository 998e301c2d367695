"""The C-ABI libraries load and export every symbol include/*.h declares; without a GPU every
compute entry point fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"FX8010_API\s+[\w\s\*]+?\b(fx8010_\w+)\s*\(", src)))


def test_gpu_header_symbols_exported(fx):
    names = declared("fx8010_gpu.h")
    assert len(names) >= 25 and "fx8010_gpu_process_batch" in names and "fx8010_gpu_create" in names
    lib = ctypes.CDLL(fx.GPU_SO)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fx8010_gpu.h but not exported"
    assert set(names) == set(fx.GPU_SYMBOLS), "python binding table out of sync with the header"


def test_multi_header_symbols_exported(fx):
    names = declared("fx8010_multi.h")
    assert "fx8010_multi_create" in names and "fx8010_multi_process_batch_host" in names
    lib = ctypes.CDLL(fx.GPU_SO)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fx8010_multi.h but not exported"
    assert set(names) == set(fx.MULTI_SYMBOLS)


def test_multi_shard_ranges_match_the_python_glue(fx):
    """Contiguous ranges [g * N / G, (g + 1) * N / G): every instance in exactly one shard, same split as the per-rank
    split bench.py uses under torchrun (fx.shard_range)."""
    L = fx.gpu_lib()
    for n, G in [(65536, 8), (262144, 8), (4096, 3), (7, 7), (1000, 6), (5, 2)]:
        covered = 0
        for g in range(G):
            lo, hi = ctypes.c_int(), ctypes.c_int()
            L.fx8010_multi_shard_range(n, g, G, ctypes.byref(lo), ctypes.byref(hi))
            assert (lo.value, hi.value) == fx.shard_range(n, g, G)
            assert lo.value == covered
            covered = hi.value
        assert covered == n


def test_host_header_symbols_exported(fx):
    names = declared("fx8010_host.h")
    assert len(names) >= 25
    fx.gpu_lib()
    lib = ctypes.CDLL(fx.HOST_SO)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fx8010_host.h but not exported"
    assert set(names) == set(fx.HOST_SYMBOLS)


def test_struct_layouts_match_header(fx):
    assert ctypes.sizeof(fx.CInstr) == 24 and ctypes.sizeof(fx.CReg) == 16
    assert ctypes.sizeof(fx.CDims) == 32 and ctypes.sizeof(fx.CLaunchInfo) == 48


def test_no_cpu_fallback(fx):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is for machines without one")
    with pytest.raises(fx.FxError) as e:
        fx.Gpu(16, 1)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)
    p = fx.Program("input in_l 0\noutput out_l 0\nmacs out_l, 0, in_l, 0.5\nend")
    assert p.loaded
    with pytest.raises(fx.FxError):
        p.process(np.zeros((4, 1), np.float32))
    with pytest.raises(fx.FxError):
        p.process_block(np.zeros((1, 4, 1), np.float32))


def test_product_never_links_the_oracle(fx):
    """Nothing under the product package may reference oracle/ (checked on sources and on the .so deps)."""
    pkg = os.path.join(ROOT, "fx8010-emulator-core_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".h", ".py")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in text and "pyoracle" not in text and "fx_oracle_" not in text, f
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
    import subprocess
    for so in (fx.GPU_SO, fx.HOST_SO):
        deps = subprocess.run(["ldd", so], capture_output=True, text=True).stdout
        assert "oracle" not in deps and "fx8010_ref" not in deps
