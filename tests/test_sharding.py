"""CPU tier: the N>1 path -- time-sharding with guard overlap and the hit-list merge -- under
torch.distributed (gloo, world_size 2).  The compute behind each shard is the oracle in stateless
mode (test-only injection; on the GPU box bench.py runs the same sharding over the CUDA blocks)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_excerpt
from oracle import oracle as O
import gr_bluetooth_b200  # noqa: F401
from gr_bluetooth_b200 import sharding


def test_shard_calls_cover_everything_once():
    for n in (0, 1, 7, 64, 1601):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                f, c = sharding.shard_calls(n, world, r)
                seen += list(range(f, f + c))
            assert seen == list(range(n))


def test_shard_span_has_guard():
    S, H = 62500, 395001
    lo, hi = sharding.shard_span(100, 10, S, H)
    assert lo == 100 * S - (H - 1) and hi == 109 * S + 1
    x = np.arange(1000, dtype=np.float32).astype(np.complex64)
    seg = sharding.extract_span(x, -5, 10)
    assert np.all(seg[:5] == 0) and np.array_equal(seg[5:], x[:10])
    seg = sharding.extract_span(x, 995, 1005)
    assert np.array_equal(seg[:5], x[995:]) and np.all(seg[5:] == 0)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ex = load_excerpt("keyboard1", "stateless")
    P = O.Plan(ex["fs"], ex["fc"])

    def process(span, first, n):
        # span[0] is absolute sample first*S-(H-1)
        return P.run(span, first_call=first, num_calls=n, stateless=True,
                     iq_first=first * P.S - (P.H - 1), n_total=len(ex["iq"]))["hits"]

    merged = sharding.run_sharded(ex["iq"], process, P.S, P.H, world, rank, dist, batch=5)
    if rank == 0:
        q.put(merged.tobytes())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process():
    ex = load_excerpt("keyboard1", "stateless")
    P = O.Plan(ex["fs"], ex["fc"])
    single = P.run(ex["iq"], stateless=True)["hits"]
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = np.frombuffer(q.get(timeout=300), dtype=O.HIT_DTYPE)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(single) >= 3
    assert np.array_equal(merged, single)


def test_one_stream_generator_is_range_independent():
    """bench.py --gpus N: every rank generates ITS slots of one synthetic stream (synth.generate_range); any two ranges must
    agree sample for sample where they overlap, bursts that straddle a range boundary included, and so must the truths."""
    from gr_bluetooth_b200 import synth
    fs, fc, S = 8e6, 2476.5e6, 5000
    a, ta = synth.generate_range(fs, fc, 0, 24, occupancy=0.3)
    b, tb = synth.generate_range(fs, fc, 7, 9, occupancy=0.3)
    c, tc = synth.generate_range(fs, fc, 16, 8, occupancy=0.3)
    assert np.array_equal(a[7 * S:16 * S], b) and np.array_equal(a[16 * S:], c)
    assert [t for t in ta if 7 <= t["slot"] < 16] == tb and [t for t in ta if t["slot"] >= 16] == tc
    assert len(tb) > 10
    i16, _ = synth.generate_range(fs, fc, 7, 9, occupancy=0.3, as_int16=True)
    assert i16.dtype == np.int16 and np.array_equal(i16, np.round(b.view(np.float32)).astype(np.int16))
