"""Host-side mirror of the Go `internal/audit` interface this build adds:

    type Auditor interface{ Append(leaves [][]byte) (root [32]byte, size uint64, err error); Root() ... }

A tamper-evident RFC 6962 Merkle log over issued credentials.  The reference has no such structure —
its chain check is a stub (internal/cli/vc_verification_enhanced.go:531-534) and the workflow roll-up
signs a list of VC *IDs* (internal/services/vc_service.go:525-632) — so the format is specified here:
leaf = the stored vc_document bytes (or any caller-chosen record), leaf hash = SHA-256(0x00 || leaf),
node = SHA-256(0x01 || l || r), split at the largest power of two < n.

Multi-GPU: rank g appends the contiguous, 2^k-aligned leaf range it owns; the 32-byte
subtree roots are all-gathered (NCCL over NVLink — through torch.distributed here, or the library's own
afc_comm_* for the Go host) and every rank folds the top levels redundantly with `fold_roots`.
"""
import ctypes as C

import numpy as np

from . import _abi
from .crypto import default_context, pack


class Auditor:
    def __init__(self, ctx=None):
        self.ctx = ctx or default_context()
        self._lib = _abi.load()
        h = C.c_void_p()
        _abi.check(self._lib.afc_merkle_new(self.ctx.handle, C.byref(h)), self.ctx.handle)
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            self._lib.afc_merkle_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def append(self, leaves):
        """Append(leaves) -> (root, size)."""
        buf, off = pack(leaves)
        return self.append_packed(buf, off)

    def append_packed(self, buf, off):
        root = np.zeros(32, dtype=np.uint8)
        size = C.c_uint64()
        _abi.check(self._lib.afc_merkle_append(self.handle, _abi.ptr(buf), _abi.ptr(off), len(off) - 1, _abi.ptr(root), C.byref(size)),
                   self.ctx.handle)
        return root.tobytes(), size.value

    def append_hashes(self, hashes):
        """Append already-hashed nodes (n x 32) as leaves of this tree."""
        hs = np.ascontiguousarray(hashes, dtype=np.uint8).reshape(-1, 32)
        root = np.zeros(32, dtype=np.uint8)
        size = C.c_uint64()
        _abi.check(self._lib.afc_merkle_append_hashes(self.handle, _abi.ptr(hs), hs.shape[0], _abi.ptr(root), C.byref(size)), self.ctx.handle)
        return root.tobytes(), size.value

    def root(self):
        root = np.zeros(32, dtype=np.uint8)
        size = C.c_uint64()
        _abi.check(self._lib.afc_merkle_root(self.handle, _abi.ptr(root), C.byref(size)), self.ctx.handle)
        return root.tobytes(), size.value

    # device-resident variants (torch tensors on the ctx's GPU)
    def append_dev(self, d_leaves, d_off, n, stream=None):
        _abi.check(self._lib.afc_merkle_append_dev(self.handle, _abi.ptr(d_leaves), _abi.ptr(d_off), n, self.ctx._stream(stream)), self.ctx.handle)

    def append_hashes_dev(self, d_hashes, n, stream=None):
        _abi.check(self._lib.afc_merkle_append_hashes_dev(self.handle, _abi.ptr(d_hashes), n, self.ctx._stream(stream)), self.ctx.handle)

    def root_dev(self, d_root32, stream=None):
        _abi.check(self._lib.afc_merkle_root_dev(self.handle, _abi.ptr(d_root32), self.ctx._stream(stream)), self.ctx.handle)

    # checkpoint / resume: leaf count + frontier hashes
    def save(self):
        st = np.zeros(_abi.MERKLE_STATE_BYTES, dtype=np.uint8)
        _abi.check(self._lib.afc_merkle_save(self.handle, _abi.ptr(st)), self.ctx.handle)
        return st.tobytes()

    def load(self, state):
        st = np.frombuffer(state, dtype=np.uint8).copy()
        if st.size != _abi.MERKLE_STATE_BYTES:
            raise ValueError("bad Merkle state length")
        _abi.check(self._lib.afc_merkle_load(self.handle, _abi.ptr(st)), self.ctx.handle)


def fold_roots(subtree_roots, ctx=None):
    """Root of the tree whose consecutive, equally sized 2^k-leaf blocks have the given roots (the
    last block may be a partial block's own MTH): MTH over the block roots taken as nodes."""
    a = Auditor(ctx)
    try:
        root, _ = a.append_hashes(subtree_roots)
        return root
    finally:
        a.close()


class MerkleTree:
    """afc_merkle_tree: every level kept on the device over a fixed set of leaf hashes, for bulk audit paths (N4)."""

    def __init__(self, leaf_hashes, ctx=None):
        self.ctx = ctx or default_context()
        self._lib = _abi.load()
        h = C.c_void_p()
        if hasattr(leaf_hashes, "data_ptr"):                # torch tensor on the ctx's GPU
            n = leaf_hashes.numel() // 32
            _abi.check(self._lib.afc_merkle_tree_build_dev(self.ctx.handle, _abi.ptr(leaf_hashes), n, C.byref(h)), self.ctx.handle)
        else:
            lh = np.ascontiguousarray(leaf_hashes, dtype=np.uint8).reshape(-1, 32)
            n = lh.shape[0]
            _abi.check(self._lib.afc_merkle_tree_build(self.ctx.handle, _abi.ptr(lh) if n else None, n, C.byref(h)), self.ctx.handle)
        self.handle, self.n = h, n
        root = np.zeros(32, dtype=np.uint8)
        nl, depth = C.c_uint64(), C.c_uint32()
        _abi.check(self._lib.afc_merkle_tree_root(self.handle, _abi.ptr(root), C.byref(nl), C.byref(depth)), self.ctx.handle)
        self.root, self.depth = root.tobytes(), depth.value

    def close(self):
        if getattr(self, "handle", None):
            self._lib.afc_merkle_tree_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def inclusion_proofs(self, indices):
        """-> list of audit paths (each a list of 32-byte nodes, leaf-to-root order, RFC 6962 §2.1.1)."""
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        m = idx.shape[0]
        out = np.zeros((m, self.depth, 32), dtype=np.uint8)
        lens = np.zeros(m, dtype=np.uint32)
        _abi.check(self._lib.afc_merkle_tree_inclusion_proofs(self.handle, _abi.ptr(idx), m, _abi.ptr(out), _abi.ptr(lens)), self.ctx.handle)
        return [[out[i, k].tobytes() for k in range(int(lens[i]))] for i in range(m)]

    def consistency_proof(self, first):
        """-> RFC 6962 §2.1.2 proof (list of 32-byte nodes) that this tree extends the tree of its first `first` leaves."""
        out = np.zeros((2 * max(self.depth, 1) + 2, 32), dtype=np.uint8)
        k = C.c_uint32()
        _abi.check(self._lib.afc_merkle_tree_consistency_proof(self.handle, int(first), _abi.ptr(out), C.byref(k)), self.ctx.handle)
        return [out[i].tobytes() for i in range(k.value)]


def verify_consistency_batch(first_sizes, first_roots, second_size, second_root, proofs, ctx=None):
    """Bulk check that the log at (second_size, second_root) extends each earlier checkpoint (first_sizes[i], first_roots[i])
    (RFC 9162 §2.1.4.2); proofs[i] = list of 32-byte nodes (RFC 6962 §2.1.2)."""
    ctx = ctx or default_context()
    lib = _abi.load()
    m = len(first_sizes)
    fs = np.ascontiguousarray(first_sizes, dtype=np.uint64)
    fr = np.frombuffer(b"".join(first_roots), dtype=np.uint8).copy() if m else np.zeros(32, dtype=np.uint8)
    off = np.zeros(m + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(p) for p in proofs])
    flat = np.frombuffer(b"".join(b"".join(p) for p in proofs), dtype=np.uint8).copy() if off[-1] else np.zeros(32, dtype=np.uint8)
    sr = np.frombuffer(second_root, dtype=np.uint8).copy()
    ok = np.zeros(m, dtype=np.uint8)
    _abi.check(lib.afc_merkle_verify_consistency_batch(ctx.handle, _abi.ptr(fs), _abi.ptr(fr), second_size, _abi.ptr(sr), _abi.ptr(flat),
                                                       _abi.ptr(off), m, _abi.ptr(ok)), ctx.handle)
    return ok.astype(bool)


def verify_inclusion_batch(leaf_hashes, indices, tree_size, proofs, root, ctx=None):
    """Bulk offline audit: ok[i] = path i leads from leaf_hashes[i] at position indices[i] to `root` (RFC 9162 §2.1.3.2)."""
    ctx = ctx or default_context()
    lib = _abi.load()
    if isinstance(leaf_hashes, (list, tuple)):
        leaf_hashes = np.frombuffer(b"".join(leaf_hashes), dtype=np.uint8)
    lh = np.ascontiguousarray(leaf_hashes, dtype=np.uint8).reshape(-1, 32).copy()
    m = lh.shape[0]
    idx = np.ascontiguousarray(indices, dtype=np.uint64)
    off = np.zeros(m + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(p) for p in proofs])
    flat = np.frombuffer(b"".join(b"".join(p) for p in proofs), dtype=np.uint8).copy() if off[-1] else np.zeros(32, dtype=np.uint8)
    rt = np.frombuffer(root, dtype=np.uint8).copy()
    ok = np.zeros(m, dtype=np.uint8)
    _abi.check(lib.afc_merkle_verify_inclusion_batch(ctx.handle, _abi.ptr(lh), _abi.ptr(idx), tree_size, _abi.ptr(flat), _abi.ptr(off), _abi.ptr(rt), m,
                                                     _abi.ptr(ok)), ctx.handle)
    return ok.astype(bool)
