"""Developer tool (GPU box, G GPUs): the box's host<->device copy ceiling with all G GPUs copying at once — page-locked
buffers, H2D and D2H concurrently on two streams per GPU, one host thread per GPU (what the e2e path of bench.py and the
multi-GPU executor are bound by).  usage: pcie_ceiling.py G -> one JSON line"""
import json, sys, threading, time
import torch
G = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
MB = 16 * 1024 * 1024          # one cfg2 block per direction
res = {}
def run(g, both, n_iter, out):
    torch.cuda.set_device(g)
    hin = torch.empty(MB // 4, dtype=torch.float32).pin_memory(); hout = torch.empty_like(hin).pin_memory()
    din = torch.empty(MB // 4, dtype=torch.float32, device=f"cuda:{g}"); dout = torch.empty_like(din)
    s1, s2 = torch.cuda.Stream(g), torch.cuda.Stream(g)
    for _ in range(3):
        with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)
    torch.cuda.synchronize(g)
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(n_iter):
        with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
        if both:
            with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)
    torch.cuda.synchronize(g)
    out[g] = MB * n_iter / (time.perf_counter() - t0) / 1e9
line = {"gpus": G, "bytes_per_copy": MB}
for label, gs, both in (("one_gpu_h2d_only", [0], False), ("one_gpu_both_ways", [0], True), ("all_gpus_h2d_only", list(range(G)), False), ("all_gpus_both_ways", list(range(G)), True)):
    out = {}
    barrier = threading.Barrier(len(gs))
    th = [threading.Thread(target=run, args=(g, both, 60, out)) for g in gs]
    [t.start() for t in th]; [t.join() for t in th]
    line[label] = {"gbs_each_way_per_gpu_mean": sum(out.values()) / len(out), "gbs_each_way_total": sum(out.values())}
print(json.dumps(line))
