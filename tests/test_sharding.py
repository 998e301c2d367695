"""Multi-GPU path on CPU: world_size-2 gloo run of the sharding + host-side gather logic
(contiguous instance ranges per rank, no collective on the data path, outputs gathered
on the host).  The per-rank compute stand-in is the oracle — this test checks the plumbing
(partition, per-rank controls/inputs, gather order, max-over-ranks timing), not the kernels."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_total, s, out_path):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib
    import torch
    import torch.distributed as dist
    import progs
    from oracle import pyoracle as po
    fx = importlib.import_module("fx8010-emulator-core_b200")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = fx.shard_range(n_total, rank, world)
    rng = np.random.default_rng(progs.SEED)
    cutoff = (0.001 + 0.998 * np.arange(n_total) / (n_total - 1)).astype(np.float32)
    x = progs.sine_bank(n_total, s, rng)                                  # identical on every rank, sliced by range
    prog = fx.Program(progs.CFG4_ONEPOLE)
    img = po.Image(prog.instructions(), prog.registers(), 0, 0, prog.controls(), prog.tables())
    orc = po.Oracle(img, hi - lo, 1)
    orc.set_register("filter_cutoff", cutoff[lo:hi])
    y = orc.process(np.ascontiguousarray(x[:, lo:hi]).reshape(1, s, hi - lo))
    dist.barrier()
    t = torch.tensor([float(rank + 1)]); dist.all_reduce(t, op=dist.ReduceOp.MAX)      # max-over-ranks timing reduction
    assert t.item() == world
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((lo, hi, y), gathered, dst=0)                     # host-side gather of the output slices
    if rank == 0:
        full = np.zeros((1, s, n_total), dtype=np.float32)
        for a, b, part in gathered:
            full[:, :, a:b] = part
        np.save(out_path, full)
    dist.destroy_process_group()


def test_shard_range_partitions():
    import importlib
    fx = importlib.import_module("fx8010-emulator-core_b200")
    for n in (1, 7, 4096, 65536, 262144):
        for w in (1, 2, 4, 8):
            r = [fx.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_two_rank_gloo_matches_single_process(tmp_path):
    import importlib
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import progs
    from oracle import pyoracle as po
    fx = importlib.import_module("fx8010-emulator-core_b200")
    n_total, s = 50, 64
    out = str(tmp_path / "gathered.npy")
    mp.start_processes(_worker, args=(2, _free_port(), n_total, s, out), nprocs=2, join=True, start_method="spawn")
    rng = np.random.default_rng(progs.SEED)
    cutoff = (0.001 + 0.998 * np.arange(n_total) / (n_total - 1)).astype(np.float32)
    x = progs.sine_bank(n_total, s, rng)
    prog = fx.Program(progs.CFG4_ONEPOLE)
    img = po.Image(prog.instructions(), prog.registers(), 0, 0, prog.controls(), prog.tables())
    orc = po.Oracle(img, n_total, 1)
    orc.set_register("filter_cutoff", cutoff)
    want = orc.process(x.reshape(1, s, n_total))
    got = np.load(out)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
