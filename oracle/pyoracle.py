"""oracle/pyoracle.py — TEST INFRASTRUCTURE ONLY.

ctypes bindings for the two CPU checkers:

  * ``Oracle``    — oracle/liboracle.so, the plain-C restatement (oracle/fx8010_oracle.c)
  * ``Reference`` — oracle/_ref/libfx8010_ref.so, the UNMODIFIED reference sources compiled
                    with oracle/ref_harness.cpp (only when that binary exists)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.environ.get("FX_REF_SO", os.path.join(HERE, "_ref", "libfx8010_ref.so"))

# enum values shared with include/fx8010_gpu.h
OPC = dict(macs=0, macsn=1, macw=2, macwn=3, macints=4, macintw=5, acc3=6, macmv=7, andxor=8,
           tstneg=9, limit=10, limitn=11, log=12, exp=13, interp=14, skip=15, idelay=16,
           xdelay=17, end=18)
REG_INPUT, REG_OUTPUT, REG_READ, REG_WRITE = 3, 4, 8, 9


class CInstr(C.Structure):
    _fields_ = [("opcode", C.c_int32), ("r", C.c_int32), ("a", C.c_int32), ("x", C.c_int32),
                ("y", C.c_int32), ("has_input", C.c_uint8), ("has_output", C.c_uint8),
                ("has_noise", C.c_uint8), ("reserved", C.c_uint8)]


class CReg(C.Structure):
    _fields_ = [("type", C.c_int32), ("init_value", C.c_float), ("io_index", C.c_int32),
                ("is_noise", C.c_int32)]


class CImage(C.Structure):
    _fields_ = [("instrs", C.POINTER(CInstr)), ("n_instrs", C.c_int32),
                ("regs", C.POINTER(CReg)), ("n_regs", C.c_int32),
                ("itram_size", C.c_int32), ("xtram_size", C.c_int32),
                ("log_tables", C.POINTER(C.c_double)), ("exp_tables", C.POINTER(C.c_double))]


@dataclass
class Image:
    """Decoded program image in Python form (what loadFile leaves in the reference object)."""
    instrs: list            # (opcode, r, a, x, y, has_input, has_output, has_noise)
    regs: list              # (type, init_value, io_index, name)
    itram_size: int = 0
    xtram_size: int = 0
    controls: list = field(default_factory=list)
    tables: np.ndarray | None = None      # [2][32][64] float64 (LOG, EXP)

    def reg_index(self, name: str) -> int:
        for i, r in enumerate(self.regs):
            if r[3] == name:
                return i
        return -1

    def to_c(self):
        """Returns (CImage, keepalive)."""
        ins = (CInstr * len(self.instrs))()
        for k, t in enumerate(self.instrs):
            ins[k].opcode, ins[k].r, ins[k].a, ins[k].x, ins[k].y = [int(v) for v in t[:5]]
            ins[k].has_input, ins[k].has_output, ins[k].has_noise = int(t[5]), int(t[6]), int(t[7])
        regs = (CReg * len(self.regs))()
        for k, t in enumerate(self.regs):
            regs[k].type = int(t[0]); regs[k].init_value = float(np.float32(t[1]))
            regs[k].io_index = int(t[2]); regs[k].is_noise = 1 if t[3] == "noise" else 0
        tabs = self.tables if self.tables is not None else build_tables()
        tabs = np.ascontiguousarray(tabs, dtype=np.float64)
        img = CImage(ins, len(self.instrs), regs, len(self.regs), int(self.itram_size), int(self.xtram_size),
                     tabs[0].ctypes.data_as(C.POINTER(C.c_double)), tabs[1].ctypes.data_as(C.POINTER(C.c_double)))
        return img, (ins, regs, tabs)


# ------------------------------------------------------------------------------------------------
def build_oracle(force: bool = False) -> str:
    """Compile liboracle.so (and _ref when /root/reference is present) through oracle/Makefile."""
    if force or not os.path.exists(ORACLE_SO) or \
            os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "fx8010_oracle.c")):
        subprocess.run(["make", "-s", "-C", HERE, "all"], check=True)
    return ORACLE_SO


_olib = None


def olib():
    global _olib
    if _olib is None:
        build_oracle()
        L = C.CDLL(ORACLE_SO)
        L.fx_oracle_create.restype = C.c_void_p
        L.fx_oracle_create.argtypes = [C.POINTER(CImage), C.c_int, C.c_int]
        L.fx_oracle_destroy.argtypes = [C.c_void_p]
        L.fx_oracle_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        for name, rt in [("registers", C.POINTER(C.c_float)), ("acc", C.POINTER(C.c_double)),
                         ("lfsr", C.POINTER(C.c_uint32)), ("out_latch", C.POINTER(C.c_float)),
                         ("tram_ptrs", C.POINTER(C.c_int32)), ("counts", C.POINTER(C.c_ulonglong))]:
            f = getattr(L, "fx_oracle_" + name); f.restype = rt; f.argtypes = [C.c_void_p]
        L.fx_oracle_get_tram.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.fx_oracle_set_tram.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.fx_oracle_runtime_flags.argtypes = [C.c_void_p]; L.fx_oracle_runtime_flags.restype = C.c_uint
        L.fx_oracle_build_tables.argtypes = [C.c_void_p, C.c_void_p]
        L.fx_oracle_table_eval.restype = C.c_float
        L.fx_oracle_table_eval.argtypes = [C.c_void_p, C.c_int, C.c_float, C.POINTER(C.c_uint)]
        _olib = L
    return _olib


def build_tables() -> np.ndarray:
    t = np.zeros((2, 32, 64), dtype=np.float64)
    olib().fx_oracle_build_tables(t[0].ctypes.data, t[1].ctypes.data)
    return t


def table_hash(tab: np.ndarray) -> str:
    """Word-wise FNV-1a variant from SURVEY.md §8c."""
    h = 0xcbf29ce484222325
    for w in np.ascontiguousarray(tab, dtype=np.float64).view(np.uint64).ravel():
        h = ((h ^ int(w)) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


class Oracle:
    """N instances of one decoded program on the C restatement."""

    def __init__(self, image: Image, n_instances: int = 1, n_channels: int = 1):
        self.image, self.n, self.c = image, n_instances, n_channels
        cimg, self._keep = image.to_c()
        self.h = olib().fx_oracle_create(C.byref(cimg), n_instances, n_channels)
        if not self.h:
            raise ValueError("oracle rejected the program image")

    def __del__(self):
        if getattr(self, "h", None):
            olib().fx_oracle_destroy(self.h); self.h = None

    def process(self, x: np.ndarray | None, n_samples: int | None = None, threads: int = 1) -> np.ndarray:
        """x: [C][S][N] float32 (or None with n_samples) -> out [C][S][N]."""
        if x is not None:
            x = np.ascontiguousarray(x, dtype=np.float32).reshape(self.c, -1, self.n)
            n_samples = x.shape[1]
        out = np.zeros((self.c, n_samples, self.n), dtype=np.float32)
        rc = olib().fx_oracle_process(self.h, x.ctypes.data if x is not None else None, out.ctypes.data,
                                      n_samples, threads)
        assert rc == 0
        return out

    def _view(self, name, shape, dtype):
        p = getattr(olib(), "fx_oracle_" + name)(self.h)
        return np.ctypeslib.as_array(p, shape=shape).view(dtype)

    @property
    def registers(self): return self._view("registers", (len(self.image.regs), self.n), np.float32)
    @property
    def acc(self): return self._view("acc", (self.n,), np.float64)
    @property
    def lfsr(self): return self._view("lfsr", (2, self.n), np.uint32)
    @property
    def out_latch(self): return self._view("out_latch", (self.c, self.n), np.float32)
    @property
    def tram_ptrs(self): return self._view("tram_ptrs", (4, self.n), np.int32)
    @property
    def counts(self): return self._view("counts", (self.n,), np.uint64)
    @property
    def flags(self): return int(olib().fx_oracle_runtime_flags(self.h))

    def set_register(self, name_or_index, values):
        idx = name_or_index if isinstance(name_or_index, int) else self.image.reg_index(name_or_index)
        if idx < 0:
            return 1
        self.registers[idx, :] = np.asarray(values, dtype=np.float32)
        return 0

    def tram(self, which: int, instance: int) -> np.ndarray:
        size = self.image.itram_size if which == 0 else self.image.xtram_size
        out = np.zeros(size, dtype=np.float32)
        rc = olib().fx_oracle_get_tram(self.h, which, instance, out.ctypes.data)
        return out if rc == 0 else np.zeros(0, dtype=np.float32)

    def set_tram(self, which: int, instance: int, values) -> int:
        v = np.ascontiguousarray(values, dtype=np.float32)
        return int(olib().fx_oracle_set_tram(self.h, which, instance, v.ctypes.data))


# ------------------------------------------------------------------------------------------------
_rlib = None


def have_reference() -> bool:
    return os.path.exists(REF_SO)


def rlib():
    global _rlib
    if _rlib is None:
        L = C.CDLL(REF_SO)
        L.ref_create.restype = C.c_void_p; L.ref_create.argtypes = [C.c_int]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_sizeof.restype = C.c_size_t
        L.ref_load_file.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_set_register.argtypes = [C.c_void_p, C.c_char_p, C.c_float]
        L.ref_get_register.argtypes = [C.c_void_p, C.c_char_p]; L.ref_get_register.restype = C.c_float
        for n in ("ref_get_instruction_counter", "ref_get_ready", "ref_get_channels", "ref_num_registers",
                  "ref_num_instructions", "ref_itram_size", "ref_xtram_size", "ref_num_errors", "ref_num_controls"):
            getattr(L, n).argtypes = [C.c_void_p]
        L.ref_register_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float),
                                        C.POINTER(C.c_int), C.c_char_p, C.c_int]
        L.ref_register_values.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_set_register_index.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.ref_instruction_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.ref_tram_pointers.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.ref_tram_read.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.ref_accumulator.argtypes = [C.c_void_p]; L.ref_accumulator.restype = C.c_double
        L.ref_lfsr.argtypes = [C.c_void_p, C.POINTER(C.c_uint)]
        L.ref_tables.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_error_info.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        L.ref_control_name.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        L.ref_metadata.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
        L.ref_bench.restype = C.c_double
        L.ref_bench.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p,
                                C.POINTER(C.c_ulonglong), C.POINTER(C.c_int)]
        _rlib = L
    return _rlib


class Reference:
    """One object of the unmodified reference class (Klangraum::FX8010)."""

    def __init__(self, text: str | None = None, channels: int = 1, path: str | None = None, raw: bytes | None = None):
        self.L = rlib()
        self.channels = channels
        self.h = self.L.ref_create(channels)
        self.loaded = None
        if raw is not None or text is not None or path is not None:
            self.loaded = self.load(text=text, path=path, raw=raw)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_destroy(self.h); self.h = None

    def load(self, text: str | None = None, path: str | None = None, raw: bytes | None = None) -> bool:
        if path is None:
            data = raw if raw is not None else text.encode()
            with tempfile.NamedTemporaryFile("wb", suffix=".da", delete=False) as f:
                f.write(data); tmp = f.name
            try:
                return bool(self.L.ref_load_file(self.h, tmp.encode()))
            finally:
                os.unlink(tmp)
        return bool(self.L.ref_load_file(self.h, path.encode()))

    def process(self, x: np.ndarray) -> np.ndarray:
        """x: [S][C] (or [S] when 1 channel) -> out [S][C]."""
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, self.channels)
        out = np.zeros_like(x)
        self.L.ref_process(self.h, x.ctypes.data, out.ctypes.data, x.shape[0])
        return out

    def set_register(self, name: str, v: float) -> int: return self.L.ref_set_register(self.h, name.encode(), float(v))
    def get_register(self, name: str) -> float: return float(self.L.ref_get_register(self.h, name.encode()))
    @property
    def instruction_counter(self) -> int: return self.L.ref_get_instruction_counter(self.h)
    @property
    def ready(self) -> bool: return bool(self.L.ref_get_ready(self.h))

    def registers(self):
        out = []
        for i in range(self.L.ref_num_registers(self.h)):
            t, v, io = C.c_int(), C.c_float(), C.c_int(); nb = C.create_string_buffer(256)
            self.L.ref_register_info(self.h, i, C.byref(t), C.byref(v), C.byref(io), nb, 256)
            out.append((t.value, np.float32(v.value), io.value, nb.value.decode()))
        return out

    def register_values(self) -> np.ndarray:
        out = np.zeros(self.L.ref_num_registers(self.h), dtype=np.float32)
        self.L.ref_register_values(self.h, out.ctypes.data)
        return out

    def set_register_index(self, i: int, v: float): self.L.ref_set_register_index(self.h, i, float(v))

    def instructions(self):
        out = []
        for i in range(self.L.ref_num_instructions(self.h)):
            f = (C.c_int * 8)()
            self.L.ref_instruction_info(self.h, i, f)
            out.append(tuple(int(v) for v in f))
        return out

    def tables(self) -> np.ndarray:
        t = np.zeros((2, 32, 64), dtype=np.float64)
        self.L.ref_tables(self.h, t.ctypes.data)
        return t

    def image(self) -> Image:
        return Image(self.instructions(), self.registers(), self.L.ref_itram_size(self.h),
                     self.L.ref_xtram_size(self.h), self.controls(), self.tables())

    def tram_pointers(self) -> np.ndarray:
        p = (C.c_int * 4)(); self.L.ref_tram_pointers(self.h, p)
        return np.array(list(p), dtype=np.int32)

    def tram(self, which: int, n: int) -> np.ndarray:
        out = np.zeros(n, dtype=np.float32)
        self.L.ref_tram_read(self.h, which, out.ctypes.data, n)
        return out

    @property
    def accumulator(self) -> float: return float(self.L.ref_accumulator(self.h))

    def lfsr(self) -> np.ndarray:
        p = (C.c_uint * 2)(); self.L.ref_lfsr(self.h, p)
        return np.array(list(p), dtype=np.uint32)

    def errors(self):
        out = []
        for i in range(self.L.ref_num_errors(self.h)):
            b = C.create_string_buffer(512)
            row = self.L.ref_error_info(self.h, i, b, 512)
            out.append((b.value.decode("latin-1"), row))
        return out

    def controls(self):
        out = []
        for i in range(self.L.ref_num_controls(self.h)):
            b = C.create_string_buffer(256); self.L.ref_control_name(self.h, i, b, 256)
            out.append(b.value.decode())
        return out

    def metadata(self):
        out = {}
        for k in ("name", "copyright", "created", "engine", "comment", "guid"):
            b = C.create_string_buffer(1024)
            if self.L.ref_metadata(self.h, k.encode(), b, 1024):
                out[k] = b.value.decode("latin-1")
        return out


def reference_bench(text: str, channels: int, n_threads: int, n_samples: int, x: np.ndarray | None,
                    controls: dict | None = None, want_out: bool = False):
    """Runs ref_bench: one reference object per thread.  x: [T][S][C] or None.
    controls: {name: array[T]}.  Returns (seconds, executed_instructions, out or None)."""
    L = rlib()
    with tempfile.NamedTemporaryFile("w", suffix=".da", delete=False) as f:
        f.write(text); tmp = f.name
    try:
        names = list((controls or {}).keys())
        arr = (C.c_char_p * max(1, len(names)))(*[n.encode() for n in names]) if names else None
        vals = None
        if names:
            vals = np.ascontiguousarray(np.stack([np.asarray(controls[n], dtype=np.float32) for n in names], axis=1))
        if x is not None:
            x = np.ascontiguousarray(x, dtype=np.float32).reshape(n_threads, n_samples, channels)
        out = np.zeros((n_threads, n_samples, channels), dtype=np.float32) if want_out else None
        total = C.c_ulonglong(0); ok = C.c_int(0)
        secs = L.ref_bench(tmp.encode(), channels, n_threads, n_samples,
                           x.ctypes.data if x is not None else None, len(names), arr,
                           vals.ctypes.data if vals is not None else None,
                           out.ctypes.data if out is not None else None, C.byref(total), C.byref(ok))
        if not ok.value:
            raise ValueError("reference failed to load the program")
        return secs, int(total.value), out
    finally:
        os.unlink(tmp)
