"""GPU parity of the program translator (FX8010_OPT_TRANSLATE): the NVRTC-compiled straight-line kernel against the
oracle, bit for bit, through the C ABI — and against the interpreter kernel it replaces."""
import time

import numpy as np
import pytest

import progs
from conftest import assert_bits_equal
from test_gpu_parity import compare_state, make_pair

pytestmark = pytest.mark.gpu


def run_translated(fx, po, text, n, blocks, rng, channels=1, controls=None, amp=0.9, what="case", mode=2, device_path=False):
    prog, img, orc, gpu = make_pair(fx, po, text, n, channels)
    try:
        gpu.set_option(fx.OPT_TRANSLATE, mode)
        for name, vals in (controls or {}).items():
            idx = prog.reg_index(name)
            gpu.set_controls(idx, vals)
            orc.set_register(idx, vals)
        start = 0
        for s in blocks:
            x = (2 * amp * rng.random((channels, s, n)) - amp).astype(np.float32)
            yo = orc.process(x)
            yg = gpu.process_host(x)
            assert_bits_equal(yg, yo, f"{what} outputs of block at {start}")
            start += s
        compare_state(gpu, orc, img, what, sorted({0, n - 1, n // 2}))
        return gpu.translate_status(), gpu.launch_info()
    finally:
        gpu.close()


@pytest.mark.parametrize("seed", range(12))
def test_random_programs_translated(fx, po, seed):
    rng = np.random.default_rng(3000 + seed)
    ch = 1 + seed % 2
    text = progs.random_program(rng, 32 + 9 * seed, channels=ch, xtram=bool(seed % 3 == 0))
    st, info = run_translated(fx, po, text, 96 + 37 * seed, [65, 31], rng, channels=ch, what=f"translated random {seed}")
    assert st["state"] == 2, st
    assert info.kernel_variant & 128


@pytest.mark.parametrize("seed", range(4))
def test_random_programs_unsafe_translated(fx, po, seed):
    rng = np.random.default_rng(4000 + seed)
    text = progs.random_program(rng, 70, safe=False, wild_tables=True)
    st, _ = run_translated(fx, po, text, 200, [40], rng, what=f"translated unsafe {seed}")
    assert st["state"] == 2, st


def test_cfg5_translated(fx, po):
    rng = np.random.default_rng(progs.SEED)
    n = 512
    ctl = {f"k{i}": rng.random(n).astype(np.float32) for i in range(4)}
    st, info = run_translated(fx, po, progs.cfg5_allops(), n, [64, 32], rng, controls=ctl, what="cfg5 translated")
    assert st["state"] == 2, st
    assert info.kernel_variant & 128
    assert st["regs_per_thread"] > 0


def test_noise_macmv_tram_translated(fx, po):
    rng = np.random.default_rng(11)
    text = "\n".join(["static noise", "static a", "static d", "static m", "input in_l 0", "output out_l 0", "itramsize 37 ", "xtramsize 50 ",
                      "macs a, in_l, noise, 0.125", "macmv m, a, in_l, 0.5", "macmv m, m, a, 0.25", "macs a, a, 0, 0",
                      "idelay write, a, at, 0", "idelay read, d, at, 36", "xdelay write, d, at, 3", "xdelay read, m, at, 40",
                      "skip ccr, ccr, 2, 1", "macs d, d, m, 0.5", "macs out_l, d, a, 0.5", "end"])
    st, _ = run_translated(fx, po, text, 333, [90, 45], rng, what="translated noise/macmv/tram")
    assert st["state"] == 2, st


def test_ineligible_program_keeps_the_interpreter(fx, po):
    rng = np.random.default_rng(5)
    text = "\n".join(["static n = 1", "static a", "input in_l 0", "output out_l 0", "macs n, 0, 1, 1", "skip ccr, ccr, 2, n",
                      "macs a, in_l, 0.5, 0.5", "macs out_l, a, 0, 0", "end"])
    st, info = run_translated(fx, po, text, 64, [33], rng, what="ineligible")
    assert st["state"] == -1 and "SKIP" in st["message"], st
    assert not (info.kernel_variant & 128)


def test_folded_register_changed_by_the_host(fx, po):
    """A never-written static is an immediate of the translated kernel; when the host overwrites it the kernel is rebuilt
    (once: the register is not folded again) and results keep matching."""
    rng = np.random.default_rng(6)
    n = 128
    text = "\n".join(["static g = 0.5", "static a", "input in_l 0", "output out_l 0", "macs a, 0, in_l, g", "skip ccr, ccr, 6, 1",
                      "macsn a, a, g, g", "macs out_l, a, g, 0.25", "end"])
    prog, img, orc, gpu = make_pair(fx, po, text, n)
    try:
        gpu.set_option(fx.OPT_TRANSLATE, 2)
        x = (1.8 * rng.random((1, 50, n)) - 0.9).astype(np.float32)
        assert_bits_equal(gpu.process_host(x), orc.process(x), "before")
        assert gpu.translate_status()["state"] == 2
        g = prog.reg_index("g")
        gpu.set_controls(g, np.float32(0.75), broadcast=True); orc.set_register(g, np.full(n, 0.75, np.float32))
        assert_bits_equal(gpu.process_host(x), orc.process(x), "after a broadcast change")
        assert gpu.translate_status()["state"] == 2
        v = rng.random(n).astype(np.float32)
        gpu.set_controls(g, v); orc.set_register(g, v)
        assert_bits_equal(gpu.process_host(x), orc.process(x), "after a per-instance change")
        compare_state(gpu, orc, img, "folded register")
    finally:
        gpu.close()


def test_background_compile_switches_over(fx, po):
    """Mode 1: the first launches run on the interpreter while NVRTC works in a thread; later launches use the translated
    kernel; the output stream is the same bits throughout."""
    rng = np.random.default_rng(7)
    n = 256
    text = progs.random_program(rng, 90)
    prog, img, orc, gpu = make_pair(fx, po, text, n)
    try:
        gpu.set_option(fx.OPT_TRANSLATE, 1)
        seen = set()
        t0 = time.time()
        while time.time() - t0 < 60:
            x = (1.8 * rng.random((1, 16, n)) - 0.9).astype(np.float32)
            assert_bits_equal(gpu.process_host(x), orc.process(x), "stream")
            seen.add(bool(gpu.launch_info().kernel_variant & 128))
            if True in seen:
                break
            time.sleep(0.02)
        assert True in seen, gpu.translate_status()
        compare_state(gpu, orc, img, "background")
    finally:
        gpu.close()


def test_translated_matches_interpreter_device_path(fx, po):
    """Same program, same inputs: translated kernel vs interpreter kernel through process_device at 20 000 instances."""
    import torch
    rng = np.random.default_rng(8)
    n, s = 20000, 48
    text = progs.cfg5_allops(n_instr=200)
    prog = fx.Program(text)
    ctl = {f"k{i}": rng.random(n).astype(np.float32) for i in range(4)}
    outs = []
    for mode in (0, 2):
        gpu = fx.Gpu(n, 1)
        gpu.load_program(prog)
        gpu.set_option(fx.OPT_TRANSLATE, mode)
        for name, v in ctl.items():
            gpu.set_controls(prog.reg_index(name), v)
        x = torch.from_numpy((1.8 * np.random.default_rng(9).random((s, n)) - 0.9).astype(np.float32)).cuda()
        y = torch.empty_like(x)
        st = torch.cuda.current_stream().cuda_stream
        gpu.process_device(x, y, s, st); gpu.process_device(x, y, s, st)
        gpu.synchronize(st)
        outs.append((y.cpu().numpy(), gpu.registers(), gpu.counts(), bool(gpu.launch_info().kernel_variant & 128)))
        gpu.close()
    assert outs[0][3] is False and outs[1][3] is True
    assert_bits_equal(outs[0][0], outs[1][0], "outputs")
    assert_bits_equal(outs[0][1], outs[1][1], "registers")
    assert_bits_equal(outs[0][2], outs[1][2], "counters")


# ---- stateless programs: the translated streaming kernel (fx_translated_sl) ----------------------------------------

@pytest.mark.parametrize("n", [4096, 1000, 332])
def test_cfg2_translated_streaming(fx, po, n):
    rng = np.random.default_rng(progs.SEED)
    vol = rng.random(n).astype(np.float32)
    st, info = run_translated(fx, po, progs.CFG2_LOG_GAIN, n, [1024, 40, 3], rng, controls={"volume": vol}, amp=0.99, what=f"cfg2 translated n={n}")
    assert st["state"] == 2, st
    assert info.kernel_variant & 128 and info.kernel_variant & 4


def test_stateless_odd_instance_count_keeps_the_interpreter(fx, po):
    rng = np.random.default_rng(2)
    n = 333                                      # not a multiple of 4: no 16-byte accesses
    vol = rng.random(n).astype(np.float32)
    st, info = run_translated(fx, po, progs.CFG2_LOG_GAIN, n, [100], rng, controls={"volume": vol}, what="cfg2 n=333")
    assert not (info.kernel_variant & 128)


@pytest.mark.parametrize("name", ["cfg1", "logtube", "two_channels"])
def test_stateless_programs_translated(fx, po, name):
    rng = np.random.default_rng(31)
    n = 512
    ch = 2 if name == "two_channels" else 1
    text = {"cfg1": progs.CFG1A_TESTCODE, "logtube": progs.CFG1B_LOGTUBE,
            "two_channels": "\n".join(["input in_l 0", "input in_r 1", "output out_l 0", "output out_r 1", "static t", "control g = 0.5",
                                       "macs t, in_l, in_r, g", "log t, t, 7, 0", "macs out_l, t, in_r, 0.25", "macsn out_r, in_r, in_l, g", "end"])}[name]
    prog = fx.Program(text, channels=ch)
    ctl = {nm: rng.random(n).astype(np.float32) for nm in prog.controls()}
    st, info = run_translated(fx, po, text, n, [257, 64, 1], rng, channels=ch, controls=ctl, what=name)
    assert st["state"] == 2 and (info.kernel_variant & 128), (st, hex(info.kernel_variant))


def test_fused_blocks_rotating_buffers_translated(fx, po):
    """fx8010_gpu_process_blocks on the translated streaming kernel: 7 blocks over 3 rotating buffer pairs in one call, then
    one block per call with FX8010_OPT_STREAM_EXCLUSIVE (launches overlap through the postponed wait)."""
    import torch
    rng = np.random.default_rng(41)
    n, s = 4096, 256
    prog, img, orc, gpu = make_pair(fx, po, progs.CFG2_LOG_GAIN, n)
    try:
        gpu.set_option(fx.OPT_TRANSLATE, 2)
        vol = rng.random(n).astype(np.float32)
        gpu.set_controls(prog.reg_index("volume"), vol); orc.set_register(prog.reg_index("volume"), vol)
        xs = [(1.98 * rng.random((s, n)) - 0.99).astype(np.float32) for _ in range(3)]
        d_in = [torch.from_numpy(x).cuda() for x in xs]
        d_out = [torch.zeros_like(d_in[0]) for _ in range(3)]
        st = torch.cuda.current_stream().cuda_stream
        order = [b % 3 for b in range(7)]
        gpu.process_blocks([d_in[b] for b in order], [d_out[b] for b in order], s, st)
        gpu.synchronize(st)
        assert gpu.launch_info().kernel_variant & 128
        want = {}
        for b in order:
            want[b] = orc.process(xs[b].reshape(1, s, n))[0]
        for b in range(3):
            assert_bits_equal(d_out[b].cpu().numpy(), want[b], f"fused block buffer {b}")
        gpu.set_option(fx.OPT_STREAM_EXCLUSIVE, 1)
        for b in order:
            gpu.process_device(d_in[b], d_out[b], s, st)
            want[b] = orc.process(xs[b].reshape(1, s, n))[0]
        gpu.synchronize(st)
        for b in range(3):
            assert_bits_equal(d_out[b].cpu().numpy(), want[b], f"per-call block buffer {b}")
        compare_state(gpu, orc, img, "fused translated")
    finally:
        gpu.close()


@pytest.mark.parametrize("seed", range(8))
def test_random_skip_free_programs_four_instances_per_thread(fx, po, seed, monkeypatch):
    """SKIP-free programs with carried state (noise, MACMV, TRAM, cross-instruction recurrences): 4 / 2 / 1 instances per thread
    (16- / 8- / 4-byte cp.async input ring)."""
    monkeypatch.setenv("FX8010_TR_K", ["4", "2", "1", "4"][seed % 4])
    rng = np.random.default_rng(5000 + seed)
    ch = 1 + seed % 2
    text = progs.random_program(rng, 28 + 11 * seed, channels=ch, skip=False, xtram=bool(seed % 2))
    n = 256 + 64 * seed
    st, info = run_translated(fx, po, text, n, [70, 33, 1], rng, channels=ch, what=f"translated K=4 random {seed}")
    assert st["state"] == 2, st
    assert info.kernel_variant & 128
    assert (info.kernel_variant >> 8) & 0xff == [4, 2, 1, 4][seed % 4]


def test_cfg4_translated_serial_kernel(fx, po, monkeypatch):
    monkeypatch.setenv("FX8010_TR_RECUR", "1")
    rng = np.random.default_rng(51)
    n = 2048
    ctl = {"filter_cutoff": (0.001 + 0.998 * np.arange(n) / (n - 1)).astype(np.float32)}
    st, info = run_translated(fx, po, progs.CFG4_ONEPOLE, n, [300, 100, 37], rng, controls=ctl, amp=0.99, what="cfg4 translated")
    assert st["state"] == 2 and (info.kernel_variant & 128), (st, hex(info.kernel_variant))


def test_if_converted_skips(fx, po, monkeypatch):
    """Developer switch FX8010_TR_IFCONV=1: SKIPs over pure register instructions as predicates instead of branches."""
    monkeypatch.setenv("FX8010_TR_IFCONV", "1")
    rng = np.random.default_rng(61)
    n = 512
    ctl = {f"k{i}": rng.random(n).astype(np.float32) for i in range(4)}
    st, info = run_translated(fx, po, progs.cfg5_allops(n_instr=160), n, [48, 17], rng, controls=ctl, what="if-converted cfg5-like")
    assert st["state"] == 2 and (info.kernel_variant & 128)


# ---- TRAM READ streams of the translated serial kernel (prefetched through the cp.async ring when the delay allows) ----------

from test_gpu_parity import TRAM_CASES, _tram_prog


@pytest.mark.parametrize("name", sorted(TRAM_CASES))
@pytest.mark.parametrize("ring", ["32", "8"])
def test_tram_streams_translated(fx, po, name, ring, monkeypatch):
    """The delay-line programs of test_tram_instruction_major on the translated serial kernel (FX8010_TR_RECUR=3 routes them there):
    delays shorter and longer than the prefetch depth, write-before-read taps, offsets, per-instance offsets, both rings."""
    monkeypatch.setenv("FX8010_TR_RECUR", "3")
    monkeypatch.setenv("FX8010_TR_RING", ring)
    rng = np.random.default_rng(31)
    n = 136
    text = _tram_prog(**TRAM_CASES[name])
    ctl = {"dly": rng.integers(70, 390, n).astype(np.float32)} if name == "ctl_off" else None
    prog, img, orc, gpu = make_pair(fx, po, text, n)
    try:
        gpu.set_option(fx.OPT_TRANSLATE, 2)
        for nm, v in (ctl or {}).items():
            gpu.set_controls(prog.reg_index(nm), v); orc.set_register(prog.reg_index(nm), v)
        x = progs.impulse_noise(n, 700, rng)
        start = 0
        for k in [1, 150, 33, 64, 2, 450]:
            xb = x[start:start + k].reshape(1, k, n)
            assert_bits_equal(gpu.process_host(xb), orc.process(xb), f"tram {name} block at {start}")
            start += k
        compare_state(gpu, orc, img, f"tram {name}", (0, n // 2, n - 1))
        assert gpu.launch_info().kernel_variant & 128, gpu.translate_status()
    finally:
        gpu.close()


def test_tram_streams_per_instance_pointers(fx, po, monkeypatch):
    """Pointers that differ per instance (set through fx8010_gpu_set_scalars): some threads of a warp can prefetch, others cannot."""
    monkeypatch.setenv("FX8010_TR_RECUR", "3")
    rng = np.random.default_rng(77)
    n, size = 96, 120
    text = _tram_prog(size=size, order="wr", roff="0", woff="0")
    prog, img, orc, gpu = make_pair(fx, po, text, n)
    try:
        gpu.set_option(fx.OPT_TRANSLATE, 2)
        acc, lfsr, latch, ptrs = gpu.scalars()
        ptrs = ptrs.copy()
        ptrs[0] = rng.integers(0, size, n)        # iw
        ptrs[1] = rng.integers(0, size, n)        # ir: read-after-write distances from 0 to size - 1 across the instances
        gpu.set_scalars(ptrs=ptrs)
        orc.tram_ptrs[:] = ptrs
        x = progs.impulse_noise(n, 400, rng)
        for a, k in ((0, 130), (130, 270)):
            xb = x[a:a + k].reshape(1, k, n)
            assert_bits_equal(gpu.process_host(xb), orc.process(xb), f"per-instance pointers, block at {a}")
        compare_state(gpu, orc, img, "per-instance pointers", tuple(range(0, n, 7)))
        assert gpu.launch_info().kernel_variant & 128
    finally:
        gpu.close()


def test_kernel_cache_is_shared_between_handles(fx, po):
    """A second handle with the same program (and the same folded values) finds the loaded kernel: no second compilation."""
    rng = np.random.default_rng(81)
    text = progs.random_program(rng, 120)
    prog = fx.Program(text)
    n = 128
    x = torch_free_input(rng, 24, n)
    first = None
    for k in range(2):
        gpu = fx.Gpu(n, 1)
        gpu.load_program(prog)
        gpu.set_option(fx.OPT_TRANSLATE, 2)
        t0 = time.time()
        y = gpu.process_host(x)
        dt = time.time() - t0
        assert gpu.translate_status()["state"] == 2
        if first is None:
            first = (y, dt)
        else:
            assert_bits_equal(y, first[0], "second handle")
            assert dt < max(0.5 * first[1], 0.25), (dt, first[1])      # (no second compilation: NVRTC takes 0.3 s and more for this program)
        gpu.close()


def torch_free_input(rng, s, n):
    return (1.8 * rng.random((1, s, n)) - 0.9).astype(np.float32)


def test_facade_set_translation(fx, po):
    """Klangraum::FX8010::setTranslation through the facade's C view: a SKIP program on 64 instances, translated before the first block."""
    rng = np.random.default_rng(91)
    n, s = 64, 40
    text = progs.random_program(rng, 60)
    facade = fx.Program(text, channels=1, instances=n)
    assert facade.loaded
    facade.set_translation(2)
    img = po.Image(facade.instructions(), facade.registers(), facade.itram_size, facade.xtram_size, facade.controls(), facade.tables())
    orc = po.Oracle(img, n, 1)
    x = (1.8 * rng.random((1, s, n)) - 0.9).astype(np.float32)
    assert_bits_equal(facade.process_block(x), orc.process(x), "facade block")
    with pytest.raises(ValueError):
        facade.set_translation(7)
    facade.close()


@pytest.mark.parametrize("name", sorted(TRAM_CASES))
def test_delay_lines_time_cut_on_the_streaming_kernel(fx, po, name):
    """Default routing with the translator on: the launches fx8010_gpu.cu cuts along time (independent periods over a stretch) run on
    the translated streaming kernel with TRAM, everything else of these programs on the instruction-major kernel; same bits either way."""
    rng = np.random.default_rng(31)
    n = 136
    text = _tram_prog(**TRAM_CASES[name])
    ctl = {"dly": rng.integers(70, 390, n).astype(np.float32)} if name == "ctl_off" else None
    prog, img, orc, gpu = make_pair(fx, po, text, n)
    try:
        gpu.set_option(fx.OPT_TRANSLATE, 2)
        for nm, v in (ctl or {}).items():
            gpu.set_controls(prog.reg_index(nm), v); orc.set_register(prog.reg_index(nm), v)
        x = progs.impulse_noise(n, 1900, rng)
        start, used = 0, False
        for k in [1, 150, 33, 64, 2, 450, 1200]:
            xb = x[start:start + k].reshape(1, k, n)
            assert_bits_equal(gpu.process_host(xb), orc.process(xb), f"tram {name} block at {start}")
            used = used or bool(gpu.launch_info().kernel_variant & 128)
            start += k
        compare_state(gpu, orc, img, f"tram {name}", (0, n // 2, n - 1))
        if name in ("s1000", "xdelay", "wr_off", "s100", "s66"):
            assert used, (name, gpu.translate_status())
    finally:
        gpu.close()


@pytest.mark.parametrize("size", [1000, 8192])
def test_cfg3_full_width_translated(fx, po, size):
    rng = np.random.default_rng(33)
    n = 2048
    st, info = run_translated(fx, po, progs.cfg3_delay(size), n, [1024, 1024, 300], rng, amp=0.9, what=f"cfg3 {size} translated")
    assert st["state"] == 2, st


def test_delay_line_background_compile_switches_over(fx, po):
    """Mode 1 on a delay line: the instruction-major kernel runs the time-cut launches until NVRTC is done, then the translated streaming
    kernel takes over in mid-stream; the ring positions it gets from the host continue where the other kernel left them."""
    rng = np.random.default_rng(97)
    n = 256
    # (a gain nobody else uses: the generated source — the kernel cache's key — does not depend on the ring size, and a cached kernel
    #  would be in use from the first launch on)
    prog, img, orc, gpu = make_pair(fx, po, progs.cfg3_delay(300).replace("0.5", "0.4921875"), n)
    try:
        gpu.set_option(fx.OPT_TRANSLATE, 1)
        seen = set()
        t0 = time.time()
        while time.time() - t0 < 60:
            k = int(rng.integers(1, 300))
            x = (1.8 * rng.random((1, k, n)) - 0.9).astype(np.float32)
            assert_bits_equal(gpu.process_host(x), orc.process(x), "delay line stream")
            seen.add(bool(gpu.launch_info().kernel_variant & 128))
            if True in seen:
                break
            time.sleep(0.02)
        for k in (299, 7, 300):
            x = (1.8 * rng.random((1, k, n)) - 0.9).astype(np.float32)
            assert_bits_equal(gpu.process_host(x), orc.process(x), "delay line stream, after the switch")
        assert seen == {False, True}, (seen, gpu.translate_status())
        compare_state(gpu, orc, img, "delay line background")
    finally:
        gpu.close()
