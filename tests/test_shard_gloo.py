"""CPU, world_size 2 over gloo: the multi-GPU path is block-range sharding with no data-path collective; the
only exchange is the all_gather of shard sizes.  Concatenating the shards must reproduce the 1-rank stream."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H

sys.path.insert(0, os.path.join(H.ROOT, "7bgzf_b200"))
import shard


def test_partition_covers_everything():
    for nb in (0, 1, 2, 7, 16449, 1000003):
        for world in (1, 2, 4, 8):
            ranges = [shard.block_range(nb, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == nb
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            assert max(e - s for s, e in ranges) - min(e - s for s, e in ranges) <= 1
    assert shard.byte_range(100000, 0xFF00, 1, 2) == (0xFF00, 100000)
    assert shard.output_offsets([5, 7, 9]) == ([0, 5, 12], 21)


def _worker(rank, world, port, data, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, e = shard.byte_range(len(data), H.BLOCK, rank, world)
    part = H.emul_stream(data[s:e], 6, eof=False)            # stand-in for the GPU codec: same algorithm, same bytes
    sizes = shard.gather_sizes(len(part))
    dist.barrier()
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                   # the bench's max-over-ranks timing reduction
    q.put((rank, part, sizes, t.item()))
    dist.destroy_process_group()


def test_two_ranks_concatenate_to_single_rank_stream():
    data = H.synth("fastq", 5 * H.BLOCK + 1234)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, data, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    sizes = res[0][2]
    assert sizes == res[1][2] == [len(res[0][1]), len(res[1][1])]
    assert res[0][3] == res[1][3] == 2.0
    offs, total = shard.output_offsets(sizes)
    whole = bytearray(total)
    for rank, part, _, _ in res:
        whole[offs[rank] : offs[rank] + len(part)] = part
    assert bytes(whole) + H.EOF_BLOCK == H.emul_stream(data, 6)
