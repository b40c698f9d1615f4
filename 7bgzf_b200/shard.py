"""Multi-GPU sharding of the BGZF hot path: blocks are independent, so rank g of G takes the
contiguous block range [B*g/G, B*(g+1)/G); there is no data-path collective.  The only exchange is the host-side
gather of per-rank output sizes (host-known offsets for concatenation) — done here with torch.distributed
all_gather on a tiny tensor (gloo on CPU, nccl on GPU)."""


def block_range(nblocks, rank, world):
    return nblocks * rank // world, nblocks * (rank + 1) // world


def byte_range(nbytes, block_size, rank, world):
    nblocks = (nbytes + block_size - 1) // block_size
    b0, b1 = block_range(nblocks, rank, world)
    return min(b0 * block_size, nbytes), min(b1 * block_size, nbytes)


def output_offsets(sizes):
    """exclusive prefix of per-rank compressed sizes: where each rank's shard lands in the concatenated stream"""
    offs, acc = [], 0
    for s in sizes:
        offs.append(acc)
        acc += s
    return offs, acc


def gather_sizes(local_size, device="cpu"):
    """all ranks learn every rank's shard size; returns a python list (world_size == 1: no communication)"""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return [int(local_size)]
    t = torch.tensor([int(local_size)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(x.item()) for x in out]
