"""N > 1 host logic on CPU: two gloo ranks shard an image batch, "caption" their shard and all-gather the result
(the GPU path runs the same code over NCCL in bench.py)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_caption(ids, T):
    """Deterministic per-image 'caption' that depends only on the global image id (what Philox keyed by row_ids gives)."""
    tok = (ids.view(-1, 1) * 7 + torch.arange(T).view(1, -1)).to(torch.int32)
    ln = (ids % T + 1).to(torch.int32)
    return tok, ln


def _worker(rank, world, port, n_items, T, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import clipcap_b200 as cc
    ids = cc.sharding.global_row_ids(n_items, rank, world)
    tok, ln = _fake_caption(ids, T)
    all_tok, all_len = cc.sharding.gather_captions(tok, ln, n_items)
    torch.save((all_tok, all_len), os.path.join(out_dir, "r%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [8, 7, 1])
def test_two_ranks_gather_in_global_order(tmp_path, n_items):
    world, T = 2, 5
    mp.spawn(_worker, args=(world, _free_port(), n_items, T, str(tmp_path)), nprocs=world, join=True)
    want_tok, want_len = _fake_caption(torch.arange(n_items), T)
    for r in range(world):
        tok, ln = torch.load(os.path.join(str(tmp_path), "r%d.pt" % r))
        assert torch.equal(tok, want_tok) and torch.equal(ln, want_len)


def test_shard_ranges_partition_the_batch():
    import clipcap_b200 as cc
    for n in (0, 1, 5, 16, 16384):
        for world in (1, 2, 3, 8):
            spans = [cc.sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        cc.sharding.shard_range(4, 2, 2)
