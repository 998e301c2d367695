#!/usr/bin/env python
"""bench.py — the reference's headline metric on B200: instance-samples/s (and DSP instructions/s)
of the batched FX8010 interpreter, next to the reference's own CPU interpreter.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg3|cfg4|cfg5]

One "step" = one fx8010_gpu_process_batch call = one block of 1 024 samples for every instance of
this rank (BASELINE.json configs[1]: MACS gain + LOG waveshaper, 4 096 instances per GPU).  Instances
shard across ranks with no collective on the data path (weak scaling: every rank runs the same
per-GPU workload).  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
PKG = "fx8010-emulator-core_b200"

import progs  # noqa: E402  (program texts + stimuli shared with the tests)

BLOCK = 1024          # samples per step (48 kHz blocks of 1 024 samples)
L2_BYTES = 126 << 20


def workload(name: str):
    """(program text, instances per GPU, algorithmic bytes per instance-sample, label)."""
    if name == "cfg2":
        return progs.CFG2_LOG_GAIN, 4096, 8, "configs[1]: MACS gain + LOG waveshaper, 4096 instances x 1024-sample blocks per GPU"
    if name == "cfg1":
        return progs.CFG1A_TESTCODE, 4096, 8, "configs[0] batched: shipped testcode.da (MACS gain), 4096 instances x 1024-sample blocks per GPU"
    if name == "cfg3":
        return progs.cfg3_delay(1000), 16384, 16, "configs[2]: idelay feedback delay line (itramsize 1000), 16384 instances x 1024-sample blocks per GPU"
    if name == "cfg4":
        return progs.CFG4_ONEPOLE, 65536, 8, "configs[3]: INTERP one-pole low-pass bank, 65536 instances x 1024-sample blocks per GPU"
    if name == "cfg5":
        return progs.cfg5_allops(), 32768, 8, "configs[4]: 512-instruction all-opcode program, 32768 instances x 1024-sample blocks per GPU"
    raise SystemExit(f"unknown config {name}")


def controls_for(name: str, prog, n: int, rng):
    if name in ("cfg1", "cfg2"):
        return {"volume": rng.random(n).astype(np.float32)}
    if name == "cfg4":
        return {"filter_cutoff": (0.001 + 0.998 * np.arange(n) / max(1, n - 1)).astype(np.float32)}
    if name == "cfg5":
        return {f"k{i}": rng.random(n).astype(np.float32) for i in range(4)}
    return {}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [t.strip() for t in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# IEEE operations one executed DSP instruction needs: (FP32/INT-pipe ops, FP64-pipe ops).
# Dispatch, operand fetch and CCR are excluded — this is the arithmetic the semantics require.
OP_COST = {"macs": (4, 0), "macsn": (4, 0), "macints": (4, 0), "acc3": (4, 0), "macw": (5, 0), "macwn": (5, 0), "macintw": (5, 0),
           "macmv": (2, 1), "andxor": (10, 0), "tstneg": (7, 0), "limit": (2, 0), "limitn": (2, 0), "log": (2, 5), "exp": (2, 5),
           "interp": (6, 3), "skip": (4, 0), "idelay": (0, 0), "xdelay": (0, 0), "end": (0, 0)}


def compute_roofline(text: str, executed_per_step: float, n_inst: int, step_s: float, sm_mhz: float, bytes_per: int, hbm_gbs: float):
    """SURVEY.md §8d for compute-bound programs: t_roofline = max(bytes / BW_HBM, sum FP32 ops / P32, sum FP64 ops / P64)
    with the sums from the program's opcode histogram scaled by the executed fraction (skipped instructions do no
    arithmetic), P32 = 128 and P64 = 64 lane-ops per clock and SM at the SM clock seen during the run (nominal pipe
    widths, not microbenchmarked)."""
    ops = [ln.split()[0] for ln in text.lower().splitlines() if ln.split() and ln.split()[0] in OP_COST]
    f32 = sum(OP_COST[o][0] for o in ops)
    f64 = sum(OP_COST[o][1] for o in ops)
    frac = executed_per_step / (len(ops) * float(n_inst) * BLOCK)          # executed / issued
    clk = (sm_mhz or 1965.0) * 1e6
    t32 = f32 * frac * n_inst * BLOCK / (148 * 128 * clk)
    t64 = f64 * frac * n_inst * BLOCK / (148 * 64 * clk)
    thbm = bytes_per * n_inst * BLOCK / (hbm_gbs * 1e9)
    t_roof = max(t32, t64, thbm)
    return {"bound": "fp32-pipe" if t_roof == t32 else ("fp64-pipe" if t_roof == t64 else "hbm"),
            "t_roofline_us": 1e6 * t_roof, "t_measured_us": 1e6 * step_s, "frac": t_roof / step_s,
            "fp32_ops_per_sample": f32 * frac, "fp64_ops_per_sample": f64 * frac, "executed_fraction": frac,
            "peaks": "nominal: 148 SMs x 128 FP32 / 64 FP64 lane-ops per clock at the SM clock sampled during the run"}


def ncu_traffic(cfg: str):
    """dram bytes per launch of the interpreter kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(cfg)
        except Exception:
            return None
    return None


def cpu_reference_rate(text: str, cfg: str, target_seconds: float, threads: int):
    """The reference's own interpreter (oracle/_ref, unmodified sources) — or the oracle port when
    that binary is absent — one instance per thread on all host cores.  Returns a dict."""
    from oracle import pyoracle as po
    rng = np.random.default_rng(progs.SEED)
    probe = 100_000
    if po.have_reference():
        ctl = controls_for(cfg, None, threads, rng)

        def run(ns):
            x = progs.sine_bank(threads, min(ns, 48000), rng).T.copy()            # [T][S]
            if ns > x.shape[1]:
                x = np.tile(x, (1, (ns + x.shape[1] - 1) // x.shape[1]))[:, :ns].copy()
            secs, instr, _ = po.reference_bench(text, 1, threads, ns, x.reshape(threads, ns, 1), ctl)
            return secs, instr
        secs, _ = run(probe)
        ns = int(max(probe, min(2_000_000_000 // max(1, threads), probe * target_seconds / max(secs, 1e-6))))
        secs, instr = run(ns)
        return {"value": threads * ns / secs, "unit": "instance-samples/s", "cores": threads, "kind": "reference",
                "dsp_instr_per_s": instr / secs,
                "sample": f"{threads} threads x 1 instance x {ns} samples, one process() call per sample (reference main.cpp loop), {secs:.2f} s"}
    prog_mod = importlib.import_module(PKG)
    prog = prog_mod.Program(text)
    img = po.Image(prog.instructions(), prog.registers(), prog.itram_size, prog.xtram_size, prog.controls(), prog.tables())
    n = threads * 8
    orc = po.Oracle(img, n, 1)
    x = progs.sine_bank(n, 4096, rng).reshape(1, 4096, n)
    t0 = time.perf_counter(); orc.process(x, threads=threads); secs = time.perf_counter() - t0
    reps = max(1, int(target_seconds / max(secs, 1e-6)))
    t0 = time.perf_counter()
    for _ in range(reps):
        orc.process(x, threads=threads)
    secs = time.perf_counter() - t0
    return {"value": n * 4096 * reps / secs, "unit": "instance-samples/s", "cores": threads, "kind": "port",
            "dsp_instr_per_s": None, "sample": f"oracle port, {n} instances x {4096 * reps} samples on {threads} threads, {secs:.2f} s"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--instances", type=int, default=0, help="override instances per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    text, n_inst, bytes_per, label = workload(args.config)
    if args.instances:
        n_inst = args.instances
    instr_per_sample = sum(progs.opcode_histogram(text).values())

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        per_step = max(0.2, min(1.0, 60.0 / max(1, args.steps + args.warmup)))
        base = cpu_reference_rate(text, args.config, per_step, threads)     # calibrates the step size
        from oracle import pyoracle as po
        # timed: K steps of the same bounded sample
        rng = np.random.default_rng(progs.SEED)
        ns = max(1024, int(base["value"] / threads * per_step))
        vals = []
        if po.have_reference():
            ctl = controls_for(args.config, None, threads, rng)
            x = progs.sine_bank(threads, min(ns, 48000), rng).T.copy()
            if ns > x.shape[1]:
                x = np.tile(x, (1, (ns + x.shape[1] - 1) // x.shape[1]))[:, :ns].copy()
            x = x.reshape(threads, ns, 1)
            total_s, total_instr = 0.0, 0
            for i in range(args.warmup + args.steps):
                secs, instr, _ = po.reference_bench(text, 1, threads, ns, x, ctl)
                if i >= args.warmup:
                    total_s += secs; total_instr += instr
            value = threads * ns * args.steps / total_s
            kind, ips = "reference", total_instr / total_s
        else:
            value, kind, ips, total_s = base["value"], "port", None, per_step * args.steps
        line = {"impl": "reference", "metric": "instance_samples_per_s", "value": value, "unit": "instance-samples/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
                "dsp_instr_per_s": ips,
                "config": {"workload": label, "program_instructions_per_sample": instr_per_sample},
                "cpu_baseline": {"value": value, "unit": "instance-samples/s", "cores": threads, "kind": kind,
                                 "sample": f"per step: {threads} threads x 1 instance x {ns} samples, one process() per sample"},
                "e2e": {"value": value, "unit": "instance-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------- our arm (GPU)
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    fx = importlib.import_module(PKG)
    rng = np.random.default_rng(progs.SEED + rank)

    prog = fx.Program(text)
    assert prog.loaded, prog.errors()
    gpu = fx.Gpu(n_inst, 1, local_rank)
    gpu.load_program(prog)
    for name, v in controls_for(args.config, prog, n_inst, rng).items():
        gpu.set_controls(prog.reg_index(name), v)

    # inputs: rotate over enough buffer pairs that a step never finds its data in L2
    block_bytes = 4 * n_inst * BLOCK
    n_bufs = max(2, -(-2 * L2_BYTES // (2 * block_bytes)) + 1)
    if args.config == "cfg3":
        host_in = [progs.impulse_noise(n_inst, BLOCK, rng) for _ in range(2)]
    else:
        host_in = [progs.sine_bank(n_inst, BLOCK, rng, start=b * BLOCK) for b in range(2)]
    d_in = [torch.from_numpy(host_in[b % 2]).cuda() for b in range(n_bufs)]
    d_out = [torch.empty_like(d_in[0]) for _ in range(n_bufs)]
    stream = torch.cuda.Stream()
    st = stream.cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        gpu.process_device(d_in[i % n_bufs], d_out[i % n_bufs], BLOCK, st)
    barrier()
    launches0 = gpu.launch_info().kernel_launches
    count0 = gpu.count_total()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for i in range(args.steps):
            gpu.process_device(d_in[i % n_bufs], d_out[i % n_bufs], BLOCK, st)
        e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = gpu.launch_info().kernel_launches - launches0
    # keep the clock sampler running over a longer window of the same work so it sees load
    if rank == 0:
        t_end = time.time() + 0.4
        while time.time() < t_end:
            for i in range(50):
                gpu.process_device(d_in[i % n_bufs], d_out[i % n_bufs], BLOCK, st)
            gpu.synchronize(st)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    total_samples = float(n_inst) * BLOCK * args.steps * world
    value = total_samples / (ms * 1e-3)
    info = gpu.launch_info()

    # executed DSP instructions (END included, skipped excluded): from the device counters
    torch.cuda.synchronize()
    g2 = fx.Gpu(n_inst, 1, local_rank); g2.load_program(prog)
    for name, v in controls_for(args.config, prog, n_inst, np.random.default_rng(progs.SEED + rank)).items():
        g2.set_controls(prog.reg_index(name), v)
    g2.process_device(d_in[0], d_out[0], BLOCK, st); g2.synchronize(st)
    instr_per_step = g2.count_total()
    g2.close()

    # ---- end to end through the C ABI with HOST buffers (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        pin = [fx.pinned_array((1, BLOCK, n_inst)) for _ in range(2)]
        pout = [fx.pinned_array((1, BLOCK, n_inst)) for _ in range(2)]
        for b in range(2):
            pin[b][0][...] = host_in[b].reshape(1, BLOCK, n_inst)
        e2e_steps = max(10, min(args.steps, 200))
        n_host = 4                                       # page-locked in/out pairs the steps rotate over
        pin += [fx.pinned_array((1, BLOCK, n_inst)) for _ in range(n_host - 2)]
        pout += [fx.pinned_array((1, BLOCK, n_inst)) for _ in range(n_host - 2)]
        for b in range(2, n_host):
            pin[b][0][...] = host_in[b % 2].reshape(1, BLOCK, n_inst)

        def e2e_run(wait):
            for i in range(3):
                gpu.process_host_ptr(pin[i % n_host][0].ctypes.data, pout[i % n_host][0].ctypes.data, BLOCK)
            barrier()
            t0 = time.perf_counter()
            for i in range(e2e_steps):
                gpu.process_host_ptr(pin[i % n_host][0].ctypes.data, pout[i % n_host][0].ctypes.data, BLOCK, wait=wait)
            gpu.synchronize(None)                        # every step's output has reached host memory
            barrier()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
            return dt
        dt_sync = e2e_run(True)                          # one blocking call per step
        dt = e2e_run(False)                              # queued calls: copies of consecutive steps overlap
        e2e = {"value": float(n_inst) * BLOCK * e2e_steps * world / dt, "unit": "instance-samples/s",
               "h2d_bytes_per_step": block_bytes, "d2h_bytes_per_step": block_bytes, "steps": e2e_steps,
               "ms_per_step": 1e3 * dt / e2e_steps, "host_buffers": "pinned (fx8010_gpu_host_alloc)",
               "api": "fx8010_gpu_process_batch_host_async per step + one fx8010_gpu_synchronize",
               "blocking_call_value": float(n_inst) * BLOCK * e2e_steps * world / dt_sync,
               "blocking_call_ms_per_step": 1e3 * dt_sync / e2e_steps}

    if rank == 0:
        peak, peak_src = measured_peak()
        alg_bytes = bytes_per * n_inst * BLOCK
        launch_ms = ms / max(1, launches)
        achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
        line = {"metric": "instance_samples_per_s", "value": value, "unit": "instance-samples/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
                "dsp_instr_per_s": instr_per_step * world * args.steps / (ms * 1e-3),
                "config": {"workload": label, "instances_per_gpu": n_inst, "block_samples": BLOCK,
                           "program_instructions_per_sample": instr_per_sample,
                           "l2": f"rotating {n_bufs} input/output buffer pairs ({2 * n_bufs * block_bytes >> 20} MiB > 126 MiB L2)",
                           "kernel": {"grid": info.last_grid, "block": info.last_block, "time_split": info.last_time_split,
                                      "smem_bytes": info.last_smem_bytes, "instances_per_thread": (info.kernel_variant >> 8) & 0xff, "samples_per_batch": info.kernel_variant >> 16}},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": ncu_traffic(args.config), "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_us": 1e3 * launch_ms},
                "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e}
        if args.config == "cfg5":        # compute-bound program: the arithmetic roofline of SURVEY.md §8d beside the HBM one
            line["compute_roofline"] = compute_roofline(text, instr_per_step, n_inst, 1e-3 * ms / args.steps,
                                                        (clocks or {}).get("sm_mhz"), bytes_per, peak)
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_reference_rate(text, args.config, 10.0, os.cpu_count() or 1)
        print(json.dumps(line))
    gpu.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
