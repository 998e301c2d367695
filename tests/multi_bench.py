"""Developer tool (GPU box): the single-process multi-GPU executor (include/fx8010_multi.h) end to end with page-locked
host buffers — cfg2 weak (4 096 instances per GPU) and cfg4 strong (65 536 in total), host-side gather into one buffer.
usage: multi_bench.py G   -> one JSON line"""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import progs
fx = importlib.import_module("fx8010-emulator-core_b200")
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1
S = 1024
out = {"gpus": G, "what": "fx8010_multi_process_batch_host_async per step + one fx8010_multi_synchronize, pinned buffers, median of 3 runs of 20 steps"}
rng = np.random.default_rng(5)
for name, text, n, ctl in (("cfg2_weak", progs.CFG2_LOG_GAIN, 4096 * G, "volume"), ("cfg4_strong", progs.CFG4_ONEPOLE, 65536, "filter_cutoff")):
    prog = fx.Program(text)
    m = fx.MultiGpu(list(range(G)), n, 1)
    m.load_program(prog)
    m.set_controls(prog.reg_index(ctl), (0.001 + 0.998 * rng.random(n)).astype(np.float32))
    bufs = [(fx.pinned_array((1, S, n)), fx.pinned_array((1, S, n))) for _ in range(3)]
    x = progs.sine_bank(n, S, rng)
    for (pi, _), (po_, _) in bufs:
        pi[0] = x
    ts = []
    for rep in range(4):
        t0 = time.perf_counter()
        for i in range(20):
            (pi, _), (po_, _) = bufs[i % 3]
            m.process_host(pi, out=po_, wait=False)
        m.synchronize()
        ts.append((time.perf_counter() - t0) / 20)
    dt = float(np.median(ts[1:]))
    out[name] = {"instances": n, "ms_per_step": 1e3 * dt, "instance_samples_per_s": n * S / dt, "pcie_gbs_each_way_total": 4 * n * S / dt / 1e9}
    if name == "cfg4_strong":       # the parameter sweep driven by ONE input signal: [channel][sample] in, [channel][sample][instance] out
        xb = np.ascontiguousarray(x[:, :1].T)          # [1][S]
        ts = []
        for rep in range(4):
            t0 = time.perf_counter()
            for i in range(20):
                m.process_host_broadcast(xb, out=bufs[i % 3][1][0], wait=False)
            m.synchronize()
            ts.append((time.perf_counter() - t0) / 20)
        dt = float(np.median(ts[1:]))
        out["cfg4_strong_broadcast_input"] = {"instances": n, "ms_per_step": 1e3 * dt, "instance_samples_per_s": n * S / dt, "pcie_gbs_d2h_total": 4 * n * S / dt / 1e9}
    m.close()
print(json.dumps(out))
