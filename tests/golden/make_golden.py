"""Generates tests/golden/*.json|npz from the UNMODIFIED reference (oracle/_ref/libfx8010_ref.so,
compiled from /root/reference by oracle/Makefile).  Run in the authoring container:

    python tests/golden/make_golden.py

The reference ships no test vectors of its own, so these files — outputs of the
reference itself on seeded inputs — are what pins the oracle and the front-end on machines where
/root/reference does not exist.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import progs  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def frontend_record(text: str, channels: int = 1):
    r = po.Reference(text, channels=channels)
    return {"text": text, "channels": channels, "loaded": bool(r.loaded), "ready": bool(r.ready),
            "registers": [[int(t), int(np.float32(v).view(np.uint32)), int(io), n] for t, v, io, n in r.registers()],
            "instructions": [list(i) for i in r.instructions()],
            "errors": [[d, int(row)] for d, row in r.errors()],
            "controls": r.controls(), "metadata": r.metadata(),
            "itram": int(r.L.ref_itram_size(r.h)), "xtram": int(r.L.ref_xtram_size(r.h))}


def exec_programs():
    rng = np.random.default_rng(progs.SEED)
    cases = {"cfg1a_testcode": (progs.CFG1A_TESTCODE, 1), "cfg1b_logtube": (progs.CFG1B_LOGTUBE, 1),
             "cfg2_log_gain": (progs.CFG2_LOG_GAIN, 1), "cfg3_delay_100": (progs.cfg3_delay(100), 1),
             "cfg3_delay_1000": (progs.cfg3_delay(1000), 1), "cfg4_onepole": (progs.CFG4_ONEPOLE, 1),
             "cfg5_allops": (progs.cfg5_allops(), 1)}
    for k, v in progs.SNIPPETS.items():
        cases["snippet_" + k] = (v, 1)
    for s in range(8):
        cases[f"random_{s}"] = (progs.random_program(np.random.default_rng(500 + s), 60, channels=1 + s % 2, xtram=(s % 2 == 0)), 1 + s % 2)
    for s in range(3):
        cases[f"unsafe_{s}"] = (progs.random_program(np.random.default_rng(600 + s), 60, safe=False), 1)
    cases["two_channel_quirk"] = ("static a\ninput in_l 0\ninput in_r 1\noutput out_l 0\noutput out_r 1\nmacs out_r, 0, in_r, 1.0\n"
                                  "macs out_l, in_r, in_l, 0.5\nmacs a, 0.1, 0.5, in_r\nend", 2)
    cases["end_skipped_wrap"] = (progs.END_SKIPPED_WRAP, 1)
    return cases, rng


def main():
    assert po.have_reference(), "oracle/_ref/libfx8010_ref.so missing: run make -C oracle"
    # ---- front-end fixtures
    fe = {name: frontend_record(text) for name, text in progs.FRONTEND_CASES.items()}
    fe["testcode_da_shipped_2ch"] = frontend_record(progs.FRONTEND_CASES["io_out_of_range"], channels=2)
    frng = np.random.default_rng(4242)
    for k in range(300):
        fe[f"fuzz_{k:03d}"] = frontend_record(progs.fuzz_source(frng))
    ref_path = "/root/reference/source/testcode.da"
    if os.path.exists(ref_path):
        rec = frontend_record(open(ref_path, "rb").read().decode("latin-1"))
        rec["text"] = None            # the shipped file itself is not copied into this repo
        rec["note"] = "decoded image of the reference's shipped source/testcode.da (text not stored)"
        fe["shipped_testcode_da"] = rec
    json.dump(fe, open(os.path.join(HERE, "frontend.json"), "w"), indent=0, sort_keys=True)

    # ---- tables
    tabs = po.Reference("end").tables()
    # ---- execution fixtures
    cases, rng = exec_programs()
    meta, arrays = {}, {"tables": tabs}
    for name, (text, ch) in cases.items():
        r = po.Reference(text, channels=ch)
        assert r.loaded, (name, r.errors())
        S = 384 if not name.startswith("cfg3_delay_1000") else 2300
        if name.startswith("cfg5"):
            S = 48
        x = (1.8 * rng.random((S, ch)) - 0.9).astype(np.float32)
        if name.startswith("cfg1b") or name.startswith("cfg2"):
            x[:12, 0] = np.array([1, -1, 0, -0.0, 1 / 63, -1 / 63, 0.99999994, 0.5, -0.5, 0.0159, 0.9375, 1e-30], dtype=np.float32)
        if name.startswith("cfg3"):
            x *= 0.25 / 0.9; x[0] = 1.0
        ctl = {}
        for c in r.controls():
            ctl[c] = float(np.float32(rng.random()))
            r.set_register(c, ctl[c])
        y = r.process(x)
        meta[name] = {"text": text, "channels": ch, "controls": ctl, "instruction_counter": r.instruction_counter,
                      "tram_ptrs": [int(v) for v in r.tram_pointers()], "accumulator_bits": int(np.float64(r.accumulator).view(np.uint64)),
                      "lfsr": [int(v) for v in r.lfsr()]}
        arrays[name + "__x"] = x
        arrays[name + "__y"] = y
        arrays[name + "__regs"] = r.register_values()
        its = r.L.ref_itram_size(r.h)
        uses_i = any(i[0] == 16 for i in r.instructions())
        if uses_i and its:
            arrays[name + "__itram"] = r.tram(0, its)
    json.dump(meta, open(os.path.join(HERE, "exec.json"), "w"), indent=0, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "exec.npz"), **arrays)
    print("golden:", len(fe), "front-end records,", len(meta), "execution records")


if __name__ == "__main__":
    main()
