"""Developer tool (GPU box): run ONE program a few times (for ncu). usage: probe_one.py <progs attr or text file> N"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import progs
fx = importlib.import_module("fx8010-emulator-core_b200")
name, N = sys.argv[1], int(sys.argv[2]); S = 1024
text = getattr(progs, name) if hasattr(progs, name) else open(name).read()
if callable(text): text = text()
p = fx.Program(text); assert p.loaded, p.errors()
g = fx.Gpu(N, 1); g.load_program(p)
x = torch.rand(1, S, N, device="cuda") - 0.5; y = torch.empty_like(x)
for _ in range(4): g.process_device(x, y, S, None)
g.synchronize(None); print("ok", hex(g.launch_info().kernel_variant))
