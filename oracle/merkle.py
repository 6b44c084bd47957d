"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by agentfield_b200/ (the product).

RFC 6962 §2.1 Merkle Tree Hash — the audit-log append format this build specifies (the
reference has no Merkle tree: its chain check is a stub that returns true,
control-plane/internal/cli/vc_verification_enhanced.go:531-534).

  MTH({})        = SHA-256()
  MTH({d0})      = SHA-256(0x00 || d0)
  MTH(D[0:n])    = SHA-256(0x01 || MTH(D[0:k]) || MTH(D[k:n])),  k = largest power of two < n

Pinned by the Certificate-Transparency reference roots for the 8 standard leaves
(tests/golden/rfc6962.json).
"""
import hashlib


def leaf_hash(d: bytes) -> bytes:
    return hashlib.sha256(b"\x00" + d).digest()


def node_hash(l: bytes, r: bytes) -> bytes:
    return hashlib.sha256(b"\x01" + l + r).digest()


def root_from_leaf_hashes(hs):
    """MTH over precomputed leaf hashes (iterative; equals the recursive RFC definition)."""
    n = len(hs)
    if n == 0:
        return hashlib.sha256(b"").digest()
    # frontier algorithm: stack of (height, hash) complete subtrees, left to right
    stack = []
    for h in hs:
        node = (0, h)
        while stack and stack[-1][0] == node[0]:
            lh = stack.pop()
            node = (node[0] + 1, node_hash(lh[1], node[1]))
        stack.append(node)
    acc = stack.pop()[1]
    while stack:
        acc = node_hash(stack.pop()[1], acc)
    return acc


def root(leaves):
    return root_from_leaf_hashes([leaf_hash(d) for d in leaves])


def root_recursive(leaves):
    n = len(leaves)
    if n == 0:
        return hashlib.sha256(b"").digest()
    if n == 1:
        return leaf_hash(leaves[0])
    k = 1
    while k * 2 < n:
        k *= 2
    return node_hash(root_recursive(leaves[:k]), root_recursive(leaves[k:]))


def frontier(hs):
    """Compact append state: list of (height, hash) of the complete subtrees, left to right."""
    stack = []
    for h in hs:
        node = (0, h)
        while stack and stack[-1][0] == node[0]:
            lh = stack.pop()
            node = (node[0] + 1, node_hash(lh[1], node[1]))
        stack.append(node)
    return stack


def root_from_frontier(stack):
    if not stack:
        return hashlib.sha256(b"").digest()
    acc = stack[-1][1]
    for _, h in reversed(stack[:-1]):
        acc = node_hash(h, acc)
    return acc


def inclusion_proof(hs, m):
    """RFC 6962 §2.1.1 audit path for leaf index m over leaf hashes hs."""
    def mth(lo, hi):
        return root_from_leaf_hashes(hs[lo:hi])

    def path(m, lo, hi):
        n = hi - lo
        if n == 1:
            return []
        k = 1
        while k * 2 < n:
            k *= 2
        if m < k:
            return path(m, lo, lo + k) + [mth(lo + k, hi)]
        return path(m - k, lo + k, hi) + [mth(lo, lo + k)]

    return path(m, 0, len(hs))


def verify_inclusion(leaf_h, m, n, proof, root_h):
    """RFC 9162 §2.1.3.2 verification algorithm."""
    if m >= n:
        return False
    fn, sn = m, n - 1
    r = leaf_h
    for p in proof:
        if sn == 0:
            return False
        if (fn & 1) or fn == sn:
            r = node_hash(p, r)
            if not (fn & 1):
                while fn and not (fn & 1):
                    fn >>= 1
                    sn >>= 1
        else:
            r = node_hash(r, p)
        fn >>= 1
        sn >>= 1
    return sn == 0 and r == root_h


def consistency_proof(hs, m):
    """RFC 6962 §2.1.2: PROOF(m, D[n]) over leaf hashes hs (n = len(hs)), 0 < m <= n, by the recursive definition."""
    n = len(hs)
    if not 0 < m <= n:
        raise ValueError("need 0 < m <= n")

    def mth(lo, hi):
        return root_from_leaf_hashes(hs[lo:hi])

    def sub(m, lo, hi, b):
        n = hi - lo
        if m == n:
            return [] if b else [mth(lo, hi)]
        k = 1
        while k * 2 < n:
            k *= 2
        if m <= k:
            return sub(m, lo, lo + k, b) + [mth(lo + k, hi)]
        return sub(m - k, lo + k, hi, False) + [mth(lo, lo + k)]

    return sub(m, 0, n, True)


def verify_consistency(first, second, first_root, second_root, proof):
    """RFC 9162 §2.1.4.2: the tree of `second` leaves with root second_root extends the tree of `first` leaves with
    root first_root (iterative algorithm — independent of the recursive construction above)."""
    if not 0 < first <= second:
        return False
    if first == second:
        return not proof and first_root == second_root
    path = list(proof)
    if first & (first - 1) == 0:
        path = [first_root] + path
    if not path:
        return False
    fn, sn = first - 1, second - 1
    while fn & 1:
        fn >>= 1
        sn >>= 1
    fr = sr = path[0]
    for c in path[1:]:
        if sn == 0:
            return False
        if (fn & 1) or fn == sn:
            fr = node_hash(c, fr)
            sr = node_hash(c, sr)
            while fn and not (fn & 1):
                fn >>= 1
                sn >>= 1
        else:
            sr = node_hash(sr, c)
        fn >>= 1
        sn >>= 1
    return sn == 0 and fr == first_root and sr == second_root
