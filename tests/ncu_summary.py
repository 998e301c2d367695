"""Developer tool: summarise an .ncu-rep (raw + source pages) the way profiles/*.md reports it.
usage: python tests/ncu_summary.py gpurun_out/prof.ncu-rep [--sass]"""
import collections
import csv
import subprocess
import sys


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep = sys.argv[1]
    rows = page(rep, "raw")
    h, d = rows[0], rows[2]
    want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
            'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_active.avg.per_cycle_active',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__average_warp_latency_per_inst_issued.ratio',
            'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
            'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'lts__t_bytes.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
            'sm__inst_executed_pipe_lsu.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fp64.sum']
    for w in want:
        if w in h:
            print(f"{w:72s} {rows[1][h.index(w)]:16s} {d[h.index(w)]}")
    print("-- stall samples")
    for i, x in enumerate(h):
        if 'pcsamp_warps_issue_stalled' in x and 'not_issued' not in x and float(d[i] or 0) > 0:
            print(f"   {x.replace('smsp__pcsamp_warps_issue_stalled_', ''):28s} {d[i]}")
    rows = page(rep, "source", ("--print-source", "sass"))
    hdr, out = None, []
    for r in rows:
        if r and r[0] == "Address":
            hdr = r; continue
        if hdr and len(r) > 5:
            out.append(r)
    ci, si, st = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
    tot = sum(int(r[ci] or 0) for r in out); tots = sum(int(r[st] or 0) for r in out)
    print(f"-- {tot} warp instructions, {tots} stall samples; by execution count:")
    b, bs, nb = collections.Counter(), collections.Counter(), collections.Counter()
    for r in out:
        c = int(r[ci] or 0)
        if c:
            b[c] += c; bs[c] += int(r[st] or 0); nb[c] += 1
    for c, t in sorted(b.items(), key=lambda x: -x[1])[:10]:
        print(f"   x{c:8d}: {nb[c]:4d} sass, {t:9d} ({100 * t / tot:4.1f}%), stall samples {bs[c]:5d} ({100 * bs[c] / max(1, tots):4.1f}%)")
    if "--sass" in sys.argv:
        top = [c for c, _ in sorted(b.items(), key=lambda x: -x[1])[:int(sys.argv[sys.argv.index("--sass") + 1]) if sys.argv[-1].isdigit() else 3]]
        for r in out:
            c = int(r[ci] or 0)
            if c in top:
                print(f"{c:8d} {int(r[st] or 0):4d} {r[si][:100]}")


if __name__ == "__main__":
    main()
