"""Developer tool (GPU box): find the first instruction at which the CUDA path and the oracle part
ways for one program.  usage: python tests/debug_bisect.py random:<seed> | file:<path>"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import progs  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

fx = importlib.import_module("fx8010-emulator-core_b200")


def state_diff(text, n, ch, x):
    prog = fx.Program(text, channels=ch)
    if not prog.loaded:
        return None
    img = po.Image(prog.instructions(), prog.registers(), prog.itram_size, prog.xtram_size, prog.controls(), prog.tables())
    try:
        orc = po.Oracle(img, n, ch)
    except ValueError:
        return None
    gpu = fx.Gpu(n, ch); gpu.load_program(prog)
    yg = gpu.process_host(x); yo = orc.process(x)
    rg, ro = gpu.registers(), orc.registers
    bad = np.argwhere(rg.view(np.uint32) != ro.view(np.uint32))
    out_bad = np.argwhere(yg.view(np.uint32) != yo.view(np.uint32))
    acc, lfsr, latch, ptrs = gpu.scalars()
    other = []
    if not np.array_equal(acc.view(np.uint64), orc.acc.view(np.uint64)): other.append("acc")
    if not np.array_equal(lfsr, orc.lfsr): other.append("lfsr")
    if not np.array_equal(ptrs, orc.tram_ptrs): other.append("ptrs")
    if not np.array_equal(gpu.counts(), orc.counts): other.append("counts")
    names = [r[3] for r in prog.registers()]
    gpu.close()
    return bad, out_bad, other, names, rg, ro


def main():
    kind, arg = sys.argv[1].split(":", 1)
    if kind == "random":
        seed = int(arg)
        rng = np.random.default_rng(1000 + seed)
        n = [128, 100, 37, 64][seed % 4]; ch = 1 + seed % 2
        text = progs.random_program(rng, 60 + 10 * seed, channels=ch, xtram=(seed % 3 == 0), read_offsets=(seed % 4 == 1))
    else:
        text = open(arg).read(); n, ch = 64, 1
    rng = np.random.default_rng(99)
    S = 6
    x = (1.8 * rng.random((ch, S, n)) - 0.9).astype(np.float32)
    lines = text.split("\n")
    first_instr = next(i for i, l in enumerate(lines) if l.split(" ")[0] in progs.opcode_histogram("\n".join(["macs x"])) or l.split(" ")[0] in
                       ("macs", "macsn", "macw", "macwn", "macints", "macintw", "acc3", "macmv", "andxor", "tstneg", "limit", "limitn", "log", "exp", "interp", "skip", "idelay", "xdelay"))
    body_end = len(lines) - 1
    for k in range(first_instr + 1, body_end + 1):
        t = "\n".join(lines[:k] + ["end"])
        r = state_diff(t, n, ch, x)
        if r is None:
            continue
        bad, out_bad, other, names, rg, ro = r
        if len(bad) or len(out_bad) or other:
            print(f"first divergence with {k - first_instr} instructions; last = {lines[k - 1]!r}")
            print("context:"); print("\n".join(lines[max(first_instr, k - 8):k]))
            for reg, inst in bad[:6]:
                print(f"  reg {names[reg]}[{inst}]: gpu {rg[reg, inst]!r} oracle {ro[reg, inst]!r}")
            print("  outputs differing:", len(out_bad), "other:", other)
            return
    print("no divergence found")


if __name__ == "__main__":
    main()
