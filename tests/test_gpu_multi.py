"""Multi-GPU checks (need >= 2 GPUs: `gpurun --gpus 2`): the one exchange step of the path — the audit Merkle-root
all-gather — through NCCL, both via torch.distributed and via the library's own dlopen'ed NCCL (afc_comm_*, what the Go
host uses), plus round-robin sharding of independent verify units with no collective."""
import os
import socket

import numpy as np
import pytest

from oracle import merkle as OM

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, q):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import agentfield_b200 as afb
    from agentfield_b200 import _abi, shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = afb.Context(rank)
    rng = np.random.default_rng(0xAF04)
    leaves = [rng.integers(0, 256, 96, dtype=np.uint8).tobytes() for _ in range(n)]
    lo, hi = shard.merkle_shard_range(n, rank, world)
    aud = afb.Auditor(ctx)
    local_root, size = aud.append(leaves[lo:hi])
    assert size == hi - lo
    # (1) all-gather through torch.distributed (NCCL over NVLink)
    roots = shard.allgather_roots(local_root)
    folded = afb.fold_roots(np.frombuffer(b"".join(roots), dtype=np.uint8), ctx)
    # (2) the library's own NCCL path
    lib = _abi.load()
    uid = np.zeros(128, dtype=np.uint8)
    if rank == 0:
        _abi.check(lib.afc_comm_unique_id(uid.ctypes.data))
    t = torch.from_numpy(uid).cuda(); dist.broadcast(t, 0); uid = t.cpu().numpy()
    _abi.check(lib.afc_comm_init(ctx.handle, world, rank, uid.ctypes.data), ctx.handle)
    allr = np.zeros((world, 32), dtype=np.uint8)
    lr = np.frombuffer(local_root, dtype=np.uint8).copy()
    _abi.check(lib.afc_comm_allgather_roots(ctx.handle, lr.ctypes.data, allr.ctypes.data), ctx.handle)
    folded2 = afb.fold_roots(allr, ctx)
    _abi.check(lib.afc_comm_destroy(ctx.handle))
    # (3) independent units, round-robin over ranks, no collective: every rank verifies its share; results gathered only to check
    nv = 4096
    seeds = rng.integers(0, 256, (nv, 32), dtype=np.uint8)
    msgs = rng.integers(0, 256, (nv, 512), dtype=np.uint8)
    mine = np.arange(rank, nv, world)
    off = np.arange(len(mine) + 1, dtype=np.uint64) * 512
    sigs = ctx.sign_packed(seeds[mine].copy(), msgs[mine].reshape(-1).copy(), off)
    sigs[::5, 2] ^= 1
    ok = ctx.verify_packed(ctx.pubkeys(seeds[mine].copy()), sigs, msgs[mine].reshape(-1).copy(), off)
    q.put((rank, folded.hex(), folded2.hex(), OM.root(leaves).hex(), int(ok.sum()), len(mine) - len(range(0, len(mine), 5))))
    dist.destroy_process_group()


def test_merkle_root_allgather_and_round_robin_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    world = 2
    ps = [mpc.Process(target=_worker, args=(r, world, port, 5000, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=300) for _ in ps]
    for p in ps:
        p.join(120)
    for rank, f1, f2, full, okc, expc in res:
        assert f1 == full and f2 == full, rank
        assert okc == expc


def test_single_process_drives_several_gpus():
    """How the Go control plane uses the library (SURVEY.md §8e "Replicas"): ONE process, one afc_ctx per device, batches
    dealt round-robin from concurrent threads, results scattered back by index."""
    import threading
    import torch
    import agentfield_b200 as afb
    from oracle import c_oracle as CO
    g = torch.cuda.device_count()
    if g < 2:
        pytest.skip("needs 2 GPUs")
    ctxs = [afb.Context(d) for d in range(g)]
    rng = np.random.default_rng(0xAF66)
    n = 6000
    seeds = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    msgs = rng.integers(0, 256, (n, 512), dtype=np.uint8)
    off = np.arange(n + 1, dtype=np.uint64) * 512
    sigs = CO.ed25519_sign_batch(seeds, msgs.reshape(-1), off, 8)
    sigs[::4, 7] ^= 1
    pks = CO.ed25519_pubkey_batch(seeds, 8)
    expect = CO.ed25519_verify_batch(pks, sigs, msgs.reshape(-1), off, 8)
    ok = np.zeros(n, dtype=np.uint8)
    batch = 500

    def worker(d):
        for b0 in range(d * batch, n, g * batch):            # batch i -> GPU i mod G
            b1 = min(n, b0 + batch)
            o = np.arange(b1 - b0 + 1, dtype=np.uint64) * 512
            ok[b0:b1] = ctxs[d].verify_packed(pks[b0:b1].copy(), sigs[b0:b1].copy(), msgs[b0:b1].reshape(-1).copy(), o)
    ts = [threading.Thread(target=worker, args=(d,)) for d in range(g)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert (ok == expect).all() and ok.sum() == n - len(range(0, n, 4))
    for c in ctxs:
        c.close()
