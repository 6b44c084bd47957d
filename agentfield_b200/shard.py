"""Sharding of the pending-credential queue over the GPUs of one box.

Sign / verify / HMAC / SHA-256 units are independent: any partition works and there is NO collective on
the data path.  The Merkle append is the one exchange step: leaves must be split into CONTIGUOUS,
2^k-ALIGNED ranges (not round-robin) so each GPU's local RFC 6962 root is a complete subtree of the
global tree; the 32-byte subtree roots are all-gathered and folded redundantly on every rank.

torch.distributed is the plumbing (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""


def shard_range(n, rank, world):
    """Contiguous block partition of n independent units: [lo, hi) for `rank`."""
    lo = n * rank // world
    hi = n * (rank + 1) // world
    return lo, hi


def round_robin_indices(n, rank, world):
    """Item i -> GPU i mod G (how the control plane deals its pending queue; north_star)."""
    return range(rank, n, world)


def merkle_block(n, world):
    """Smallest power of two B with world * B >= n: rank g owns leaves [g*B, min((g+1)*B, n))."""
    b = 1
    while b * world < n:
        b *= 2
    return b


def merkle_shard_range(n, rank, world):
    b = merkle_block(n, world)
    lo = min(rank * b, n)
    hi = min((rank + 1) * b, n)
    return lo, hi


def allgather_roots(local_root, group=None, device=None):
    """All-gather one 32-byte subtree root per rank; returns a list of `world` byte strings (ranks whose
    range is empty contribute None).  `local_root` may be None for an empty shard."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    send = torch.zeros(33, dtype=torch.uint8, device=device)
    if local_root is not None:
        send[0] = 1
        send[1:] = torch.frombuffer(bytearray(local_root), dtype=torch.uint8).to(device)
    recv = [torch.zeros(33, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    out = []
    for t in recv:
        t = t.cpu()
        out.append(bytes(t[1:].tolist()) if int(t[0]) else None)
    return out
