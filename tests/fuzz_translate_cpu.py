"""Developer tool (no GPU needed): differential fuzz of the program translator's code generator — random programs of both
generator families, 1 / 2 / 4 instances per thread, stateless and stateful, translated source compiled as plain C++
(tests/translate_host.py) against the oracle, all state compared.
usage: fuzz_translate_cpu.py [seconds]   (prints the first failing case, exit code 1)"""
import importlib, os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import progs
from oracle import pyoracle as po
import test_translate as TT
import conftest
from conftest import assert_bits_equal as exact_equal

fx = importlib.import_module("fx8010-emulator-core_b200")


def nan_tolerant_equal(a, b, what=""):
    a, b = np.array(a, copy=True), np.array(b, copy=True)
    if a.dtype.kind == "f":
        a[np.isnan(a)] = np.nan; b[np.isnan(b)] = np.nan
    exact_equal(a, b, what)


budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
t0, seed, n_ok, n_declined, n_sl, n_delay = time.time(), 0, 0, 0, 0, 0
while time.time() - t0 < budget:
    seed += 1
    rng = np.random.default_rng(991000 + seed)
    os.environ["FX8010_TR_K"] = ["1", "2", "4"][seed % 3]
    kind = seed % 4
    ch = 2 if seed % 7 == 3 else 1
    if kind == 0:
        text = progs.random_flow_program(rng, int(rng.integers(2, 18)), channels=ch, tram=["", "i", "x", "ix", ""][seed % 5], size=int(rng.choice([5, 64, 70, 129])))
    elif kind == 1:
        text = progs.random_program(rng, int(rng.integers(8, 100)), channels=ch, xtram=bool(seed % 2), read_offsets=(seed % 4 == 1), skip=bool(seed % 5))
    elif kind == 2:
        text = progs.random_program(rng, int(rng.integers(8, 80)), safe=False, skip=bool(seed % 2), wild_tables=True)
        ch = 1
    else:                                  # stateless candidates: no SKIP / TRAM / noise
        text = progs.random_program(rng, int(rng.integers(3, 40)), channels=ch, skip=False, tram=False, noise=False,
                                    ops=progs.SAT_OPS + progs.TABLE_OPS + ["limit", "limitn", "tstneg", "andxor"] + progs.WRAP_OPS)
    n = int(rng.choice([4, 8, 12]))
    blocks = [int(b) for b in rng.choice([1, 2, 7, 8, 31, 33], size=int(rng.integers(1, 4)))]
    TT.assert_bits_equal = nan_tolerant_equal if kind == 2 else exact_equal
    prog = fx.Program(text, channels=ch)
    if not prog.loaded:
        continue
    src, _ = fx.translate_source(prog, ch, instances=n)
    if src is None:
        n_declined += 1
        continue
    if "fx_translated_sl" in src and "#define FXT_NTR 0" not in src:
        n_delay += 1              # a delay line cut along time: needs the stretch length (tests/test_translate.py covers those with known spans)
        continue
    n_sl += "fx_translated_sl" in src
    try:
        ctl = {nm: rng.random(n).astype(np.float32) for nm in prog.controls()}
        TT.check(po, text, n, blocks, rng, channels=ch, controls=ctl, what=f"fuzz {seed}")
        n_ok += 1
    except Exception:
        print("FAILED seed", seed, "K", os.environ["FX8010_TR_K"], "n", n, "blocks", blocks, "channels", ch)
        print(text)
        traceback.print_exc()
        sys.exit(1)
print(f"translator fuzz (CPU check of the generated source): {n_ok} cases passed ({n_sl} on the streaming kernel), {n_declined} declined by the translator, {n_delay} delay lines skipped, {time.time() - t0:.0f} s")
