"""Multi-process host logic of the utterance sharding, world size 2 over gloo on CPU:
ranks own disjoint contiguous utterance ranges, no data-path collective; only the timing max / coverage check
uses a collective, like bench.py does with NCCL."""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_utt, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = importlib.import_module("audio-visual-speech-enhancement_b200.engine")
    lo, hi = eng.shard_range(n_utt, rank, world)
    cover = torch.zeros(n_utt, dtype=torch.int32)
    cover[lo:hi] += 1
    dist.all_reduce(cover)                      # test-only check that the shards tile the corpus exactly once
    t = torch.tensor([float(hi - lo) * 1e-3])   # stand-in for this rank's CUDA-event time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)    # bench.py: max over ranks
    # the product-level driver's optional final gather (engine.CorpusDriver.gather): rows come back in corpus order
    drv = eng.CorpusDriver(None, launch=7)
    assert (drv.rank, drv.world_size) == (rank, world) and drv.shard(n_utt) == (lo, hi)
    mine = torch.arange(lo, hi, dtype=torch.float32).unsqueeze(1).expand(hi - lo, 3).contiguous()
    full = drv.gather(mine, n_utt, dst=0)
    gathered_ok = True
    if rank == 0:
        gathered_ok = full.shape == (n_utt, 3) and bool((full[:, 0] == torch.arange(n_utt, dtype=torch.float32)).all())
    else:
        gathered_ok = full is None
    flag = torch.tensor([1 if gathered_ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        q.put((bool((cover == 1).all()) and bool(flag[0] == 1), float(t[0]), hi - lo))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_utt", [100000, 1001, 3])
def test_two_ranks_tile_the_corpus(n_utt):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_utt, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, tmax, n0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok
    assert n0 == (n_utt + 1) // 2
    assert abs(tmax - n0 * 1e-3) < 1e-5 * max(1.0, n0 * 1e-3)


def test_shard_range_properties():
    eng = importlib.import_module("audio-visual-speech-enhancement_b200.engine")
    for n in (0, 1, 7, 1000, 100000):
        for w in (1, 2, 4, 8):
            spans = [eng.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
