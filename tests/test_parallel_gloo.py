"""world_size-2 gloo test of the multi-GPU host logic (no GPU): batch sharding, the single weight-blob
broadcast, per-rank seeds, and the final token gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from paella_b200 import parallel as P
    lo, hi = P.shard_range(total, rank, world)
    # rank 0 "packs" the weights; everyone else starts from zeros; ONE broadcast
    blob = torch.arange(4096, dtype=torch.int64).to(torch.uint8) if rank == 0 else torch.zeros(4096, dtype=torch.uint8)
    P.broadcast_blob(blob, src=0)
    ok_blob = bool((blob == torch.arange(4096, dtype=torch.int64).to(torch.uint8)).all())
    # each rank "samples" its shard: token grid filled with the global sample index
    toks = torch.arange(lo, hi, dtype=torch.int64)[:, None, None].expand(hi - lo, 2, 2).contiguous()
    sizes = [P.shard_range(total, r, world)[1] - P.shard_range(total, r, world)[0] for r in range(world)]
    full = P.gather_tokens(toks, sizes)
    ok_gather = bool((full[:, 0, 0] == torch.arange(total)).all())
    # a receiver whose blob differs from the source's (short / skipped broadcast) must be caught on EVERY rank
    bad = blob.clone()
    if rank == 1:
        bad[1000] ^= 1
    caught = False
    try:
        P.assert_same_across_ranks(P.blob_checksum(bad), "blob checksum")
    except RuntimeError:
        caught = True
    P.assert_same_across_ranks(P.blob_checksum(blob), "blob checksum")      # and the good one passes
    ret[rank] = (lo, hi, ok_blob, ok_gather, P.rank_seed(1234, rank), caught)
    dist.destroy_process_group()


def test_two_rank_shard_broadcast_gather():
    world, total = 2, 7
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), total, ret), nprocs=world, join=True)
    assert ret[0][:2] == (0, 4) and ret[1][:2] == (4, 7)
    assert all(ret[r][2] and ret[r][3] for r in range(world))
    assert ret[0][4] != ret[1][4]
    assert ret[0][5] and ret[1][5]          # the corrupted receiver was detected by both ranks


def test_blob_checksum_sees_every_byte():
    from paella_b200.parallel import blob_checksum
    g = torch.Generator().manual_seed(0)
    b = torch.randint(0, 256, (4099,), dtype=torch.uint8, generator=g)       # not a multiple of 8: the tail bytes count too
    ref = int(blob_checksum(b))
    for pos in (0, 7, 8, 2049, 4095, 4098):
        c = b.clone()
        c[pos] = (int(c[pos]) + 1) % 256
        assert int(blob_checksum(c)) != ref, pos


def test_shard_range_partitions():
    from paella_b200.parallel import shard_range
    for n in (0, 1, 7, 64, 512):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
