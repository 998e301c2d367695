"""GPU parity at the sizes and through the calls bench.py times (BASELINE.json configs at their stated instance
counts, device buffers, back-to-back launches), the launch-overlap rules, and constant-memory residency.

The oracle cannot run 65 536 instances for thousands of samples in seconds, so it runs a SAMPLE of instances
(first, last, middle of the device's range, the rest random — BASELINE.md §3): instances are independent, so feeding
the oracle the sampled columns of the same input reproduces those instances exactly.  Everything is compared bit for bit.
"""
import numpy as np
import pytest

import progs
from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu


def sample_instances(n, k, rng):
    base = {0, 1, n - 1, n - 2, n // 2, n // 2 - 1, n // 3, (2 * n) // 3}
    base = {i for i in base if 0 <= i < n}
    while len(base) < min(k, n):
        base.add(int(rng.integers(0, n)))
    return np.array(sorted(base))


class Sampled:
    """A GPU handle over n instances next to an oracle over the sampled ones."""

    def __init__(self, fx, po, text, n, idx, controls=None):
        self.prog = fx.Program(text)
        assert self.prog.loaded, self.prog.errors()
        self.img = po.Image(self.prog.instructions(), self.prog.registers(), self.prog.itram_size, self.prog.xtram_size,
                            self.prog.controls(), self.prog.tables())
        self.idx, self.n = idx, n
        self.orc = po.Oracle(self.img, len(idx), 1)
        self.gpu = fx.Gpu(n, 1)
        self.gpu.load_program(self.prog)
        for name, v in (controls or {}).items():
            r = self.prog.reg_index(name)
            self.gpu.set_controls(r, v)
            self.orc.set_register(r, np.ascontiguousarray(v[idx]))

    def check_state(self, what):
        regs = self.gpu.registers()
        assert_bits_equal(regs[:, self.idx], self.orc.registers, what + " registers")
        acc, lfsr, latch, ptrs = self.gpu.scalars()
        assert_bits_equal(acc[self.idx], self.orc.acc, what + " accumulator")
        assert_bits_equal(lfsr[:, self.idx], self.orc.lfsr, what + " lfsr")
        assert_bits_equal(latch[:, self.idx], self.orc.out_latch, what + " latch")
        assert_bits_equal(ptrs[:, self.idx], self.orc.tram_ptrs, what + " tram pointers")
        assert_bits_equal(self.gpu.counts()[self.idx], self.orc.counts, what + " counters")
        d = self.gpu.dims()
        for which, size in ((0, d.itram_size), (1, d.xtram_size)):
            if size:
                for j in (0, len(self.idx) - 1):
                    assert_bits_equal(self.gpu.tram(which, int(self.idx[j])), self.orc.tram(which, j), f"{what} tram{which}[{self.idx[j]}]")

    def close(self):
        self.gpu.close()


def _run_device_blocks(torch, S, x_blocks, n_steps, how, n_rot=4):
    """n_steps blocks through device buffers; block i reads x_blocks[i % len]; outputs rotate over n_rot buffers and are
    checked from copies made on the same stream.  Returns per-step output samples [step][S_block][len(idx)]."""
    st = torch.cuda.Stream()
    s = x_blocks[0].shape[0]
    d_x = [torch.from_numpy(x).cuda() for x in x_blocks]
    d_y = [torch.empty_like(d_x[0]) for _ in range(n_rot)]
    d_idx = torch.from_numpy(S.idx.astype(np.int64)).cuda()
    keep = []
    torch.cuda.synchronize()
    if how == "per_call":
        for i in range(n_steps):
            S.gpu.process_device(d_x[i % len(d_x)], d_y[i % n_rot], s, st.cuda_stream)
            with torch.cuda.stream(st):
                keep.append(d_y[i % n_rot].index_select(1, d_idx))
    else:                        # fused: groups of n_rot blocks per fx8010_gpu_process_blocks call
        for i0 in range(0, n_steps, n_rot):
            m = min(n_rot, n_steps - i0)
            S.gpu.process_blocks([d_x[(i0 + j) % len(d_x)] for j in range(m)], [d_y[j] for j in range(m)], s, st.cuda_stream)
            with torch.cuda.stream(st):
                keep += [d_y[j].index_select(1, d_idx) for j in range(m)]
    S.gpu.synchronize(st.cuda_stream)
    torch.cuda.synchronize()
    return [k.cpu().numpy() for k in keep]


@pytest.mark.parametrize("how", ["per_call", "per_call_exclusive", "blocks"])
def test_cfg2_timed_geometry(fx, po, how):
    """cfg2 exactly as bench.py runs it: 4 096 instances, 1 024-sample blocks, device buffers, launches back to back."""
    import torch
    rng = np.random.default_rng(progs.SEED)
    n, s, steps = 4096, 1024, 9
    idx = sample_instances(n, 96, rng)
    S = Sampled(fx, po, progs.CFG2_LOG_GAIN, n, idx, {"volume": rng.random(n).astype(np.float32)})
    try:
        if how == "per_call_exclusive":
            S.gpu.set_option(fx.OPT_STREAM_EXCLUSIVE, 1)
        xs = [progs.sine_bank(n, s, rng, start=b * s) for b in range(3)]
        ys = _run_device_blocks(torch, S, xs, steps, "per_call" if how.startswith("per_call") else "blocks")
        info = S.gpu.launch_info()
        if how == "blocks":
            assert info.last_fused_blocks >= 1 and info.kernel_launches < steps + 20
        for i in range(steps):
            yo = S.orc.process(np.ascontiguousarray(xs[i % 3][:, idx]).reshape(1, s, len(idx)))
            assert_bits_equal(ys[i][None], yo, f"cfg2 {how} step {i}")
        S.check_state(f"cfg2 {how}")
    finally:
        S.close()


@pytest.mark.parametrize("size,blocks", [(100, 3), (1000, 3), (8192, 10), (65536, 66)])
def test_cfg3_full_size(fx, po, size, blocks):
    """cfg3 at BASELINE's 16 384 instances for every ring size of SURVEY.md §8d; enough blocks that the delayed signal
    comes back through the feedback path (the 65 536-slot rings are 4.3 GB of HBM)."""
    import torch
    rng = np.random.default_rng(progs.SEED + size)
    n, s = 16384, 1024
    idx = sample_instances(n, 48, rng)
    S = Sampled(fx, po, progs.cfg3_delay(size), n, idx)
    try:
        xs = [progs.impulse_noise(n, s, rng) for _ in range(2)]
        ys = _run_device_blocks(torch, S, xs, blocks, "per_call", n_rot=2)
        for i in range(blocks):
            yo = S.orc.process(np.ascontiguousarray(xs[i % 2][:, idx]).reshape(1, s, len(idx)))
            assert_bits_equal(ys[i][None], yo, f"cfg3 S={size} block {i}")
        S.check_state(f"cfg3 S={size}")
    finally:
        S.close()


@pytest.mark.parametrize("n", [65536, 8192])
def test_cfg4_full_size(fx, po, n):
    """cfg4 at 65 536 instances (one GPU) and at 8 192 (its share on each of 8 GPUs)."""
    import torch
    rng = np.random.default_rng(progs.SEED + n)
    s = 1024
    idx = sample_instances(n, 96, rng)
    cutoff = (0.001 + 0.998 * np.arange(n) / (n - 1)).astype(np.float32)
    S = Sampled(fx, po, progs.CFG4_ONEPOLE, n, idx, {"filter_cutoff": cutoff})
    try:
        xs = [progs.sine_bank(n, s, rng, start=b * s) for b in range(2)]
        ys = _run_device_blocks(torch, S, xs, 3, "per_call", n_rot=2)
        for i in range(3):
            yo = S.orc.process(np.ascontiguousarray(xs[i % 2][:, idx]).reshape(1, s, len(idx)))
            assert_bits_equal(ys[i][None], yo, f"cfg4 n={n} block {i}")
        S.check_state(f"cfg4 n={n}")
    finally:
        S.close()


def test_cfg5_full_size(fx, po):
    """cfg5 at its per-GPU size (32 768 instances), 2 x 64 samples."""
    import torch
    rng = np.random.default_rng(progs.SEED)
    n, s = 32768, 64
    idx = sample_instances(n, 64, rng)
    ctl = {f"k{i}": rng.random(n).astype(np.float32) for i in range(4)}
    S = Sampled(fx, po, progs.cfg5_allops(), n, idx, ctl)
    try:
        xs = [progs.sine_bank(n, s, rng, start=b * s, amp_lo=0.9, amp_hi=0.9) for b in range(2)]
        ys = _run_device_blocks(torch, S, xs, 2, "per_call", n_rot=2)
        for i in range(2):
            yo = S.orc.process(np.ascontiguousarray(xs[i][:, idx]).reshape(1, s, len(idx)))
            assert_bits_equal(ys[i][None], yo, f"cfg5 block {i}")
        S.check_state("cfg5")
    finally:
        S.close()


# ---- launch overlap (programmatic dependent launch with a postponed wait) ----------------------------------

@pytest.mark.parametrize("n_rot", [2, 3, 4, 9])
def test_rotating_buffers_overlap_rule(fx, po, n_rot):
    """A launch may postpone its wait only if its buffers are disjoint from EVERY launch since the last one that waited at
    its start: rotating n_rot buffer pairs, exactly every n_rot-th launch waits at its start.  Results are checked from
    the final contents of every buffer (a write-after-write race between launch i and i - n_rot would leave stale data)."""
    import torch
    rng = np.random.default_rng(77 + n_rot)
    n, s, steps = 4096, 1024, 3 * n_rot + 2
    idx = sample_instances(n, 32, rng)
    S = Sampled(fx, po, progs.CFG2_LOG_GAIN, n, idx, {"volume": rng.random(n).astype(np.float32)})
    try:
        st = torch.cuda.Stream()
        xs = [progs.sine_bank(n, s, rng, start=b * s) for b in range(steps)]
        d_x = [torch.from_numpy(x).cuda() for x in xs]
        d_y = [torch.zeros_like(d_x[0]) for _ in range(n_rot)]
        torch.cuda.synchronize()
        # without the promise every launch waits at its start
        S.gpu.process_device(d_x[0], d_y[0], s, st.cuda_stream)
        S.gpu.process_device(d_x[1], d_y[1 % n_rot], s, st.cuda_stream)
        assert S.gpu.launch_info().last_late_wait == 0
        S.gpu.synchronize(st.cuda_stream)
        S.gpu.set_option(fx.OPT_STREAM_EXCLUSIVE, 1)
        late = []
        for i in range(steps):
            S.gpu.process_device(d_x[i], d_y[i % n_rot], s, st.cuda_stream)
            late.append(S.gpu.launch_info().last_late_wait)
        S.gpu.synchronize(st.cuda_stream)
        assert late == [0 if i % n_rot == 0 else 1 for i in range(steps)], late
        for j in range(n_rot):
            last = max(i for i in range(steps) if i % n_rot == j)
            yo = po.Oracle(S.img, len(idx), 1)
            yo.set_register(S.prog.reg_index("volume"), S.orc.registers[S.prog.reg_index("volume")].copy())
            ref = yo.process(np.ascontiguousarray(xs[last][:, idx]).reshape(1, s, len(idx)))
            assert_bits_equal(d_y[j].cpu().numpy()[:, idx][None], ref, f"buffer {j} (launch {last})")
    finally:
        S.close()


def test_outputs_chained_as_inputs(fx, po):
    """Launch i reads what launch i - 1 wrote (ring of three buffers): never a postponed wait, results = f applied i times."""
    import torch
    rng = np.random.default_rng(78)
    n, s, steps = 4096, 1024, 7
    idx = sample_instances(n, 32, rng)
    S = Sampled(fx, po, progs.CFG2_LOG_GAIN, n, idx, {"volume": (0.5 + 0.5 * rng.random(n)).astype(np.float32)})
    try:
        S.gpu.set_option(fx.OPT_STREAM_EXCLUSIVE, 1)
        st = torch.cuda.Stream()
        x = progs.sine_bank(n, s, rng)
        d = [torch.from_numpy(x).cuda(), torch.zeros(s, n, device="cuda"), torch.zeros(s, n, device="cuda")]
        torch.cuda.synchronize()
        for i in range(steps):
            S.gpu.process_device(d[i % 3], d[(i + 1) % 3], s, st.cuda_stream)
            assert S.gpu.launch_info().last_late_wait == 0
        S.gpu.synchronize(st.cuda_stream)
        y = np.ascontiguousarray(x[:, idx]).reshape(1, s, len(idx))
        for i in range(steps):
            y = S.orc.process(y)
        assert_bits_equal(d[steps % 3].cpu().numpy()[:, idx][None], y, "chained result")
    finally:
        S.close()


def test_stream_switch_is_ordered(fx, po):
    """One handle driven from two streams without host synchronisation: the library orders the launches on the device."""
    import torch
    rng = np.random.default_rng(79)
    n, s = 8192, 512
    idx = sample_instances(n, 32, rng)
    cutoff = (0.001 + 0.998 * rng.random(n)).astype(np.float32)
    S = Sampled(fx, po, progs.CFG4_ONEPOLE, n, idx, {"filter_cutoff": cutoff})
    try:
        sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
        xs = [progs.sine_bank(n, s, rng, start=b * s) for b in range(4)]
        d_x = [torch.from_numpy(x).cuda() for x in xs]
        d_y = [torch.empty_like(d_x[0]) for _ in range(4)]
        torch.cuda.synchronize()
        for i in range(4):                                  # a recurrence: block i needs the state block i - 1 left
            S.gpu.process_device(d_x[i], d_y[i], s, (sa if i % 2 == 0 else sb).cuda_stream)
        S.gpu.synchronize(sa.cuda_stream); S.gpu.synchronize(sb.cuda_stream)
        for i in range(4):
            yo = S.orc.process(np.ascontiguousarray(xs[i][:, idx]).reshape(1, s, len(idx)))
            assert_bits_equal(d_y[i].cpu().numpy()[:, idx][None], yo, f"block {i}")
        S.check_state("stream switch")
    finally:
        S.close()


# ---- constant-memory residency ---------------------------------------------------------------------------------

def test_many_live_handles_share_the_constant_arena(fx, po):
    """The reference places no limit on live FX8010 objects (include/FX8010.h:51).  Ten handles with different programs,
    three of them long enough that the 64 KiB arena cannot hold all at once, launched in turn: evicted programs
    are uploaded again transparently and every result stays exact."""
    rng = np.random.default_rng(80)
    n, s = 64, 16
    texts = [progs.CFG2_LOG_GAIN, progs.CFG4_ONEPOLE, progs.CFG1A_TESTCODE, progs.cfg3_delay(100), progs.CFG1B_LOGTUBE,
             progs.random_program(rng, 40), progs.random_program(rng, 24, xtram=True),
             progs.cfg5_allops(), progs.cfg5_allops(seed=5), progs.cfg5_allops(seed=6)]
    pairs = []
    try:
        for t in texts:
            prog = fx.Program(t)
            assert prog.loaded, prog.errors()
            img = po.Image(prog.instructions(), prog.registers(), prog.itram_size, prog.xtram_size, prog.controls(), prog.tables())
            g = fx.Gpu(n, 1)
            g.load_program(prog)
            pairs.append((prog, po.Oracle(img, n, 1), g))
        for rnd in range(3):
            for j, (prog, orc, g) in enumerate(pairs):
                x = (1.8 * rng.random((1, s, n)) - 0.9).astype(np.float32)
                assert_bits_equal(g.process_host(x), orc.process(x), f"round {rnd} handle {j}")
        for j, (prog, orc, g) in enumerate(pairs):
            assert_bits_equal(g.registers(), orc.registers, f"handle {j} registers")
            assert_bits_equal(g.counts(), orc.counts, f"handle {j} counters")
    finally:
        for _, _, g in pairs:
            g.close()


# ---- two corner cases of the latch / preload rules ----------------------------------------------------------------

LOG_Y_INPUT = "static a\nstatic b = 0.5\ninput in_l 0\noutput out_l 0\nlog a, b, 3, in_l\nmacs out_l, 0, a, 1.0\nend"
NOP_OUTPUT_R = ("static a = 0.25\nstatic t\ninput in_l 0\noutput out_l 0\nitramsize 16 \nmacs t, 0, in_l, 0.5\nidelay write, t, at, 0\n"
                "macs out_l, 0, in_l, 0.5\nidelay read, out_l, at, 3\nidelay out_l, a, at, 0\nend")
NOP_OUTPUT_R_SKIP = NOP_OUTPUT_R.replace("idelay out_l, a, at, 0", "skip ccr, ccr, 2, 1\nmacs t, t, a, a\nidelay out_l, a, at, 0")


@pytest.mark.parametrize("text", [LOG_Y_INPUT, NOP_OUTPUT_R, NOP_OUTPUT_R_SKIP])
def test_unread_input_preload_and_noop_latch_refresh(fx, po, text):
    """LOG/EXP never read Y, but an INPUT register there is still preloaded (source/FX8010.cpp:1059) and keeps the sample;
    an IDELAY whose R is an OUTPUT register does nothing, yet the latch is refreshed from R afterwards (:1229-1233)."""
    from test_gpu_parity import run_case
    rng = np.random.default_rng(81)
    run_case(fx, po, text, 96, [40, 9], rng, what="corner")


# ---- producer / consumer pairs of the instruction-major kernel -----------------------------------------------------

PAIR_PROGS = {
    "product_x_then_rw": "static a\nstatic b\ninput in_l 0\ncontrol g = 0.5\ncontrol o = 0.1\noutput out_l 0\nexp a, in_l, 7, 0\nmacs b, o, a, g\nmacsn out_l, b, 0.5, 0.25\nend",
    "product_y_macsn": "static a\ninput in_l 0\ncontrol g = 0.7\noutput out_l 0\ninterp a, 0.25, g, in_l\nmacsn out_l, 0.1, g, a\nend",
    "addend": "static a\ninput in_l 0\ncontrol g = 0.7\noutput out_l 0\nlimit a, in_l, 0.3, g\nmacs out_l, a, g, 0.5\nend",
    "addend_macsn": "static a\ninput in_l 0\noutput out_l 0\ntstneg a, in_l, 0.5, 0\nmacsn out_l, a, 0.5, 0.5\nend",
    "macints_consumer": "static a\ninput in_l 0\ncontrol g = 0.9\noutput out_l 0\nmacw a, in_l, in_l, 0.75\nmacints out_l, 0, a, g\nend",
    "inside_delay_line": "static a\nstatic rd\nstatic t\ninput in_l 0\noutput out_l 0\nitramsize 300 \nidelay read, rd, at, 0\nmacs a, in_l, rd, 0.5\n"
                         "idelay write, a, at, 0\nlog t, rd, 2, 0\nmacs out_l, 0.1, t, 0.5\nend",
    "after_recurrence": "static t\nstatic s\ninput in_l 0\ncontrol c = 0.2\noutput out_l 0\ninterp s, s, c, in_l\nexp t, s, 5, 0\nmacs out_l, 0, t, 0.9\nend",
    "two_pairs": "static a\nstatic b\nstatic d\ninput in_l 0\ncontrol g = 0.6\noutput out_l 0\nlog a, in_l, 3, 0\nmacs b, 0, a, g\nacc3 d, b, in_l, 0.1\nmacsn out_l, 0.05, d, g\nend",
}


@pytest.mark.parametrize("name", sorted(PAIR_PROGS))
@pytest.mark.parametrize("mode", ["auto", "nopairs", "M2", "K1"])
def test_producer_consumer_pairs(fx, po, name, mode, monkeypatch):
    """A result with exactly one reader — the next instruction, a MACS/MACSN with two batch-constant operands — is
    forwarded in a register instead of going through shared memory; same bits either way."""
    from test_gpu_parity import run_case
    if mode == "nopairs":
        monkeypatch.setenv("FX8010_NO_PAIRS", "1")
    elif mode == "M2":
        monkeypatch.setenv("FX8010_TUNE_M", "2")
    elif mode == "K1":
        monkeypatch.setenv("FX8010_TUNE_K", "1")
    rng = np.random.default_rng(90)
    n = 200
    ctl = {k: rng.random(n).astype(np.float32) for k in ("g", "o", "c") if f"control {k} " in PAIR_PROGS[name]}
    info = run_case(fx, po, PAIR_PROGS[name], n, [70, 31, 1], rng, controls=ctl, what=f"{name} {mode}")
    assert info.kernel_variant & 8, "instruction-major kernel expected"
    assert bool(info.kernel_variant & 64) == (mode != "nopairs"), "pair fusion flag"


# ---- the multi-GPU executor (include/fx8010_multi.h) ---------------------------------------------------------------

def _multi_case(fx, po, devices, text, n, s_blocks, controls, rng, what):
    prog = fx.Program(text)
    assert prog.loaded, prog.errors()
    img = po.Image(prog.instructions(), prog.registers(), prog.itram_size, prog.xtram_size, prog.controls(), prog.tables())
    orc = po.Oracle(img, n, 1)
    m = fx.MultiGpu(devices, n, 1)
    try:
        assert [hi - lo for lo, hi in m.shards()] == [fx.shard_range(n, g, len(devices))[1] - fx.shard_range(n, g, len(devices))[0] for g in range(len(devices))]
        m.load_program(prog)
        for name, v in controls.items():
            m.set_controls(prog.reg_index(name), v)
            orc.set_register(prog.reg_index(name), v)
        pin_in, own_i = fx.pinned_array((1, max(s_blocks), n))
        pin_out, own_o = fx.pinned_array((1, max(s_blocks), n))
        for b, s in enumerate(s_blocks):
            x = (1.8 * rng.random((1, s, n)) - 0.9).astype(np.float32)
            pin_in[0, :s] = x[0]
            y = m.process_host(pin_in[:, :s].copy(), out=None)          # pageable buffers
            yo = orc.process(x)
            assert_bits_equal(y, yo, f"{what} block {b}")
        # page-locked buffers, queued calls: two blocks in flight, then one synchronize
        xa = (1.8 * rng.random((1, s_blocks[0], n)) - 0.9).astype(np.float32)
        xb = (1.8 * rng.random((1, s_blocks[0], n)) - 0.9).astype(np.float32)
        pa, oa = fx.pinned_array(xa.shape); pb, ob = fx.pinned_array(xb.shape)
        ya, oya = fx.pinned_array(xa.shape); yb, oyb = fx.pinned_array(xb.shape)
        pa[...] = xa; pb[...] = xb
        m.process_host(pa, out=ya, wait=False)
        m.process_host(pb, out=yb, wait=False)
        m.synchronize()
        assert_bits_equal(np.array(ya), orc.process(xa), f"{what} queued block a")
        assert_bits_equal(np.array(yb), orc.process(xb), f"{what} queued block b")
        assert_bits_equal(m.registers(), orc.registers, f"{what} registers")
        assert m.count_total() == int(orc.counts.sum())
        assert m.flags() == orc.flags
        for name in controls:
            assert_bits_equal(m.get_register(prog.reg_index(name)), orc.registers[prog.reg_index(name)], f"{what} {name}")
    finally:
        m.close()


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4", "random"])
def test_multi_executor_shards_on_one_device(fx, po, name):
    """Three shards with uneven ranges (1 000 instances) on device 0: the plumbing of the multi-GPU executor — ranges,
    per-shard threads, strided scatter/gather into ONE host buffer — checked bit for bit, whatever the number of GPUs."""
    rng = np.random.default_rng(101)
    n = 1000
    text, ctl = {"cfg2": (progs.CFG2_LOG_GAIN, {"volume": rng.random(n).astype(np.float32)}),
                 "cfg3": (progs.cfg3_delay(100), {}),
                 "cfg4": (progs.CFG4_ONEPOLE, {"filter_cutoff": (0.001 + 0.998 * rng.random(n)).astype(np.float32)}),
                 "random": (progs.random_program(rng, 40, xtram=True), {})}[name]
    _multi_case(fx, po, [0, 0, 0], text, n, [130, 57], ctl, rng, f"multi x3 {name}")


def test_multi_executor_two_gpus(fx, po):
    """cfg4's shape on two GPUs: 65 536 instances -> 32 768 per device, outputs gathered into one page-locked buffer."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (runs in the multi-GPU tier)")
    rng = np.random.default_rng(102)
    n = 65536
    idx = sample_instances(n, 64, rng)
    prog = fx.Program(progs.CFG4_ONEPOLE)
    img = po.Image(prog.instructions(), prog.registers(), prog.itram_size, prog.xtram_size, prog.controls(), prog.tables())
    orc = po.Oracle(img, len(idx), 1)
    cutoff = (0.001 + 0.998 * np.arange(n) / (n - 1)).astype(np.float32)
    m = fx.MultiGpu([0, 1], n, 1)
    try:
        m.load_program(prog)
        m.set_controls(prog.reg_index("filter_cutoff"), cutoff)
        orc.set_register(prog.reg_index("filter_cutoff"), cutoff[idx].copy())
        pin, o1 = fx.pinned_array((1, 1024, n)); pout, o2 = fx.pinned_array((1, 1024, n))
        for b in range(2):
            pin[0] = progs.sine_bank(n, 1024, rng, start=b * 1024)
            m.process_host(pin, out=pout)
            yo = orc.process(np.ascontiguousarray(pin[0][:, idx]).reshape(1, 1024, len(idx)))
            assert_bits_equal(np.ascontiguousarray(pout[0][:, idx])[None], yo, f"2 GPUs block {b}")
        assert_bits_equal(m.registers()[:, idx], orc.registers, "2 GPUs registers")
    finally:
        m.close()


def test_facade_multi_device_constructor(fx, po):
    """Klangraum::FX8010(channels, instances, devices): the reference's class surface over several shards (two on device 0
    here): setRegisterValue(s), processBlock, the counters and the per-sample process() keep their meaning."""
    rng = np.random.default_rng(103)
    n = 600
    prog = fx.Program(progs.CFG2_LOG_GAIN, instances=n, devices=[0, 0])
    assert prog.loaded
    img = po.Image(prog.instructions(), prog.registers(), prog.itram_size, prog.xtram_size, prog.controls(), prog.tables())
    orc = po.Oracle(img, n, 1)
    vol = rng.random(n).astype(np.float32)
    assert prog.set_register_values("volume", vol) == 0
    orc.set_register("volume", vol)
    x = progs.sine_bank(n, 96, rng).reshape(1, 96, n)
    assert_bits_equal(prog.process_block(x), orc.process(x), "facade multi processBlock")
    assert prog.set_register("volume", 0.25) == 0
    orc.set_register("volume", np.full(n, 0.25, np.float32))
    y1 = prog.process(np.full((3, 1), 0.5, np.float32))           # per-sample legacy call: instance 0, same input for all
    yo = orc.process(np.full((1, 3, n), 0.5, np.float32))
    assert_bits_equal(y1[:, 0], yo[0, :, 0], "facade multi process()")
    assert prog.instruction_counter_total == int(orc.counts.sum())
    assert prog.instruction_counter == int(orc.counts[0])
    assert prog.get_register("volume") == np.float32(0.25)


@pytest.mark.parametrize("shards", [0, 3])
def test_broadcast_input(fx, po, shards):
    """One input signal for all instances (a parameter sweep, BASELINE.json configs[3]): the host hands over [channel][sample]
    only; same results as feeding every instance a copy.  shards = 0: one handle; 3: the multi-GPU executor."""
    rng = np.random.default_rng(104)
    n, s = 777, 300
    prog = fx.Program(progs.CFG4_ONEPOLE)
    img = po.Image(prog.instructions(), prog.registers(), prog.itram_size, prog.xtram_size, prog.controls(), prog.tables())
    orc = po.Oracle(img, n, 1)
    cutoff = (0.001 + 0.998 * rng.random(n)).astype(np.float32)
    orc.set_register("filter_cutoff", cutoff)
    g = fx.MultiGpu([0] * shards, n, 1) if shards else fx.Gpu(n, 1)
    try:
        g.load_program(prog)
        g.set_controls(prog.reg_index("filter_cutoff"), cutoff)
        for b in range(2):
            x = (1.8 * rng.random((1, s)) - 0.9).astype(np.float32)
            y = g.process_host_broadcast(x)
            yo = orc.process(np.repeat(x[:, :, None], n, axis=2))
            assert_bits_equal(y, yo, f"broadcast block {b}")
        assert_bits_equal(g.registers(), orc.registers, "broadcast registers")
    finally:
        g.close()
