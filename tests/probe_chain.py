"""Developer tool (GPU box): per-sample latency of one-instruction recurrences, N x 1024 samples."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
fx = importlib.import_module("fx8010-emulator-core_b200")
N = int(os.environ.get("PROBE_N", 65536)); S = 1024
progs = {
  "macs_rec":   "input in_l 0\noutput out_l 0\nmacs out_l, out_l, in_l, 0.001\nend",
  "interp_rec": "input in_l 0\ncontrol c = 0.1\noutput out_l 0\ninterp out_l, out_l, c, in_l\nend",
  "limit_rec":  "input in_l 0\noutput out_l 0\nlimit out_l, out_l, in_l, 0.5\nend",
  "macs2_rec":  "static a\ninput in_l 0\noutput out_l 0\nmacs a, a, in_l, 0.001\nmacs out_l, a, in_l, 0.5\nend",
  "macs4_rec":  "static a\nstatic b\nstatic c\ninput in_l 0\noutput out_l 0\nmacs a, a, in_l, 0.001\nmacs b, a, in_l, 0.5\nmacs c, b, a, 0.5\nmacs out_l, c, in_l, 0.5\nend",
  "macs8_rec":  "static a\nstatic b\nstatic c\ninput in_l 0\noutput out_l 0\n" + "macs a, a, in_l, 0.001\nmacs b, a, in_l, 0.5\nmacs c, b, a, 0.5\nmacs a, c, in_l, 0.5\n" * 2 + "macs out_l, a, b, 0.5\nend",
  "noin_rec":   "static a = 0.5\noutput out_l 0\nmacs out_l, out_l, a, 0.001\nend",
}
x = torch.rand(1, S, N, device="cuda") - 0.5
y = torch.empty_like(x)
for name, text in progs.items():
    p = fx.Program(text); assert p.loaded, p.errors()
    g = fx.Gpu(N, 1); g.load_program(p)
    st = torch.cuda.Stream()
    for _ in range(3): g.process_device(x, y, S, st.cuda_stream)
    g.synchronize(st.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(5): g.process_device(x, y, S, st.cuda_stream)
        e1.record(st)
    g.synchronize(st.cuda_stream)
    us = e0.elapsed_time(e1) * 1e3 / 5
    info = g.launch_info()
    n_instr = len(p.instructions()) - 1
    print(f"{name:12s} {us:9.1f} us/block  {us*1e-6*1.965e9/S:8.0f} cycles/sample  {us*1e-6*1.965e9/S/n_instr:7.0f} cycles/instr  grid {info.last_grid}x{info.last_block} variant 0x{info.kernel_variant:x}")
    g.close()
