"""Developer tool (GPU box): run one BASELINE config a few times for ncu. usage: probe_cfg.py cfgN instances samples [launches [itramsize]]"""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import progs
import bench
fx = importlib.import_module("fx8010-emulator-core_b200")
cfg, N, S = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
L = int(sys.argv[4]) if len(sys.argv) > 4 else 4
text = bench.workload(cfg)[0]
if cfg == "cfg3" and len(sys.argv) > 5:
    text = progs.cfg3_delay(int(sys.argv[5]))
p = fx.Program(text); assert p.loaded, p.errors()
g = fx.Gpu(N, 1); g.load_program(p)
rng = np.random.default_rng(1)
for name, v in bench.controls_for(cfg, p, N, rng).items():
    g.set_controls(p.reg_index(name), v)
amp = 0.9 if cfg == "cfg5" else 0.5
xs = [torch.from_numpy(progs.sine_bank(N, S, rng, amp_lo=amp, amp_hi=amp)).cuda() for _ in range(2)]
ys = [torch.empty_like(xs[0]) for _ in range(2)]
for i in range(L): g.process_device(xs[i % 2], ys[i % 2], S, None)
g.synchronize(None); print("ok", hex(g.launch_info().kernel_variant), g.launch_info().last_grid, g.launch_info().last_block)
