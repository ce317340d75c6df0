"""Multi-GPU host logic: one process per GPU under torch.distributed.

  * Proof-level sharding (BASELINE configs 2 and 4): proofs are independent
    given the immutable key, so rank r proves its slice and no data-path
    collective exists; an all_gather of the 192-byte proofs is offered for
    callers that want the whole batch on every rank.
  * Base-split single MSM (BASELINE config 3): bases and scalars are
    range-partitioned, each GPU reduces its range to one projective partial
    (XYZZ, 192 bytes), the partials are all-gathered (NCCL over NVLink on
    GPUs, gloo in the CPU tests) and added.  Elliptic-curve addition is not an
    NCCL reduction operator, hence gather + local add.

The reference has no multi-device path at all; bellperson
proves serially, one create_random_proof at a time
(masp_primitives/src/transaction/components/sapling/builder.rs:941-1140).
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced slice [lo, hi) of n items for `rank`."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def longest_first_queue(counts):
    """Work order for a mixed transaction batch (BASELINE config 4): Spend
    proofs first, then Convert, then Output (largest domain first), so the
    per-GPU queues drain evenly.  counts: dict circuit -> number of proofs.
    Returns a list of (circuit, index)."""
    order = sorted(counts, key=lambda c: -{"spend": 3, "convert": 2, "output": 1}.get(c, 0))
    return [(c, i) for c in order for i in range(counts[c])]


def _gather_bytes(buf, device):
    """all_gather of equal-length byte strings -> list of bytes per rank."""
    world = dist.get_world_size()
    t = torch.frombuffer(bytearray(buf), dtype=torch.uint8).to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return [bytes(o.cpu().numpy()) for o in outs]


def split_msm_g1(pv, bases, scalars, n, device="cpu"):
    """sum_i s_i * B_i with the bases split by index range over the ranks.
    Every rank returns the same 96-byte uncompressed result."""
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_range(n, rank, world)
    local = pv.G1Bases(bases[96 * lo:96 * hi], hi - lo) if hi > lo else None
    partial = local.msm_partial(scalars[32 * lo:32 * hi]) if local else bytes(192)
    parts = _gather_bytes(partial, device)
    return pv.g1_sum_partials(parts)


def prove_sharded(pv, params, assignments, r_s, s_s, gather=True, device="cpu"):
    """Each rank proves its contiguous slice of the batch.  With gather, every
    rank receives all proofs in batch order; otherwise only its own slice."""
    rank, world = dist.get_rank(), dist.get_world_size()
    n = len(assignments)
    lo, hi = shard_range(n, rank, world)
    mine = pv.create_proof_batch(assignments[lo:hi], params, r_s[lo:hi], s_s[lo:hi])
    if not gather:
        return mine
    width = max(shard_range(n, k, world)[1] - shard_range(n, k, world)[0] for k in range(world))
    padded = b"".join(mine) + bytes(192 * (width - len(mine)))
    chunks = _gather_bytes(padded, device)
    out = []
    for k in range(world):
        klo, khi = shard_range(n, k, world)
        out += [chunks[k][192 * i:192 * (i + 1)] for i in range(khi - klo)]
    return out
