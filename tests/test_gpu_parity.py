"""GPU parity tests — the CUDA path, through the C ABI (libafcrypto.so), against the oracle and the golden fixtures.
Bit-exact bar: every byte / index must match (integer work; no tolerance).  Nothing here reads /root/reference."""
import hashlib
import hmac

import numpy as np
import pytest

from conftest import golden
from oracle import c_oracle as CO, go_ed25519 as G, merkle as OM

pytestmark = pytest.mark.gpu


def _pack(msgs):
    from agentfield_b200 import pack
    return pack(msgs)


def test_native_library_is_what_runs(ctx):
    import agentfield_b200 as afb
    assert afb.LIB_PATH.endswith("libafcrypto.so")
    info = ctx.device_info()
    assert info["sm_count"] >= 100
    before = ctx.launch_count()
    buf, off = _pack([b"abc"])
    assert ctx.sha256_packed(buf, off)[0].tobytes() == hashlib.sha256(b"abc").digest()
    assert ctx.launch_count() > before                      # a kernel really launched


def test_ptx_field_arithmetic_selftest(ctx):
    """Inline-PTX carry-chain field arithmetic vs the portable versions, on the device (random + boundary operands)."""
    assert ctx.selftest(3000) == 0


# ----------------------------------------------------------------------------- golden fixtures
def test_golden_fips180_sha256(ctx):
    es = [e for e in golden("fips180.json") if e["alg"] == "sha256"]
    msgs = [bytes.fromhex(e["repeat"]) * e["count"] if "repeat" in e else bytes.fromhex(e["msg"]) for e in es]
    buf, off = _pack(msgs)
    out = ctx.sha256_packed(buf, off)
    for e, o in zip(es, out):
        assert o.tobytes().hex() == e["digest"], e["source"]


def test_golden_rfc4231_hmac(ctx):
    from agentfield_b200 import MAC
    es = golden("rfc4231.json")
    tags = MAC(ctx).hmac_sha256_batch([bytes.fromhex(e["key"]) for e in es], [bytes.fromhex(e["msg"]) for e in es])
    for e, t in zip(es, tags):
        assert t.hex() == e["tag"], e["name"]


def test_golden_rfc8032_sign_verify_pubkey(ctx):
    from agentfield_b200 import Signer, Verifier
    es = golden("rfc8032.json")
    seeds = [bytes.fromhex(e["seed"]) for e in es]
    msgs = [bytes.fromhex(e["msg"]) for e in es]
    assert [s.hex() for s in Signer(ctx).sign_batch(seeds, msgs)] == [e["sig"] for e in es]
    assert [p.hex() for p in Signer(ctx).public_keys(seeds)] == [e["pk"] for e in es]
    ok = Verifier(ctx).verify_batch([bytes.fromhex(e["pk"]) for e in es], msgs, [bytes.fromhex(e["sig"]) for e in es])
    assert all(ok)


def test_golden_ed25519_edge_set_go_rules(ctx):
    """S = L, S + L, sig[63] high bits, non-canonical R / A, small-order A, A off-curve, x = 0 with sign bit, bit flips:
    the GPU must accept/reject exactly as Go does."""
    from agentfield_b200 import Verifier
    es = golden("ed25519_edge.json")
    ok = Verifier(ctx).verify_batch([bytes.fromhex(e["pk"]) for e in es], [bytes.fromhex(e["msg"]) for e in es],
                                    [bytes.fromhex(e["sig"]) for e in es])
    for e, o in zip(es, ok):
        assert o == e["valid"], e["name"]
    assert sum(ok) >= 10


def test_golden_rfc6962_roots(ctx):
    from agentfield_b200 import Auditor
    g = golden("rfc6962.json")
    leaves = [bytes.fromhex(x) for x in g["leaves"]]
    a = Auditor(ctx)
    assert a.root() == (bytes.fromhex(g["empty_root"]), 0)
    for n in range(1, 9):                                   # one leaf at a time: exercises every frontier merge
        root, size = a.append([leaves[n - 1]])
        assert size == n and root.hex() == g["roots"][n - 1]
    for s in g["synthetic"]:
        if s["leaves"] is not None:
            b = Auditor(ctx)
            root, size = b.append([bytes.fromhex(x) for x in s["leaves"]])
            assert size == s["n"] and root.hex() == s["root"]
            b.close()
    a.close()


def test_golden_reference_flow(ctx):
    """derivePrivateKey -> did:key, hashData, and a VC-shaped canonical message (reference_flow.json)."""
    from agentfield_b200 import Hasher, Signer, Verifier
    g = golden("reference_flow.json")
    master = bytes.fromhex(g["master_seed"])
    seeds = Hasher(ctx).sha256_batch([master + d["path"].encode() for d in g["derivations"]])
    assert [s.hex() for s in seeds] == [d["seed"] for d in g["derivations"]]
    assert [p.hex() for p in Signer(ctx).public_keys(seeds)] == [d["pk"] for d in g["derivations"]]
    digs = Hasher(ctx).sha256_batch([bytes.fromhex(h["marshalled"]) for h in g["hash_data"]])
    import base64
    assert [base64.urlsafe_b64encode(d).rstrip(b"=").decode() for d in digs] == [h["hash"] for h in g["hash_data"]]
    vc = g["vc"]
    msg = vc["canonical"].encode()
    assert Signer(ctx).sign_batch([bytes.fromhex(vc["seed"])], [msg])[0].hex() == vc["sig"]
    assert Verifier(ctx).verify_batch([bytes.fromhex(vc["pk"])], [msg], [bytes.fromhex(vc["sig"])]) == [True]


# ----------------------------------------------------------------------------- randomised parity vs the oracle
def test_sha256_hmac_ragged_lengths_and_alignments(ctx):
    rng = np.random.default_rng(0xAF03)
    lens = list(range(0, 200)) + [255, 256, 257, 511, 512, 513, 1000, 1300, 4096, 70000] + [int(x) for x in rng.integers(0, 3000, 300)]
    msgs = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in lens]        # packed back to back => every alignment
    buf, off = _pack(msgs)
    out = ctx.sha256_packed(buf, off)
    assert (out == CO.sha256_batch(buf, off, 4)).all()
    assert out[3].tobytes() == hashlib.sha256(msgs[3]).digest()
    klens = [int(x) for x in rng.integers(0, 200, len(msgs))]
    klens[:8] = [0, 1, 63, 64, 65, 131, 4096, 32]
    keys = [rng.integers(0, 256, k, dtype=np.uint8).tobytes() for k in klens]
    from agentfield_b200 import pack32
    kb, ko = pack32(keys)
    tags = ctx.hmac_sha256_packed(kb, ko, buf, off)
    assert (tags == CO.hmac_sha256_batch(kb, ko, buf, off, 4)).all()
    assert tags[6].tobytes() == hmac.new(keys[6], msgs[6], hashlib.sha256).digest()


def test_streaming_sha256_payload_hasher(ctx):
    """H2 (FilePayloadStore.SaveFromReader, payload_store.go:45-97): 24 concurrent streams of 0 B ... 64 MiB advance 32 KiB at a
    time (plus ragged final chunks) through afc_sha256_update_batch; every digest equals hashlib's over the whole stream, states
    survive a save / restore between rounds, and a malformed state or a ragged middle chunk is refused."""
    import hashlib
    from agentfield_b200 import AfcError, PayloadHasher, _abi
    rng = np.random.default_rng(0xAF69)
    sizes = [0, 1, 63, 64, 65, 32767, 32768, 32769, 100_000, 1_000_000, 5_000_000, 64 << 20] + [int(x) for x in rng.integers(1, 300_000, 12)]
    datas = [rng.integers(0, 256, s_, dtype=np.uint8).tobytes() if s_ < (8 << 20) else (rng.integers(0, 256, 1 << 20, dtype=np.uint8).tobytes() * (s_ >> 20))
             for s_ in sizes]
    ph = PayloadHasher(len(datas), ctx)
    pos, done, digests = [0] * len(datas), [False] * len(datas), [None] * len(datas)
    CH = 32 * 1024                                              # the reference's copy buffer (payload_store.go:154)
    rounds = 0
    while not all(done):
        chunks, final = [], []
        for i, d in enumerate(datas):
            if done[i]:
                chunks.append(b""); final.append(False); continue
            step = CH * (64 if len(d) > (8 << 20) else 1)       # the 64 MiB stream moves 2 MiB per round to keep the test short
            if len(d) - pos[i] > step:
                chunks.append(d[pos[i]:pos[i] + step]); final.append(False); pos[i] += step
            else:
                chunks.append(d[pos[i]:]); final.append(True); pos[i] = len(d)
        out = ph.update(chunks, final)
        for i, f in enumerate(final):
            if f:
                digests[i], done[i] = out[i], True
        rounds += 1
        if rounds == 3:                                         # states are plain bytes: persist and restore between rounds
            saved = ph.states.copy(); ph = PayloadHasher(len(datas), ctx); ph.states[:] = saved
    assert [d.hex() for d in digests] == [hashlib.sha256(d).hexdigest() for d in datas]
    ph = PayloadHasher(2, ctx)
    with pytest.raises(ValueError):
        ph.update([b"x" * 65, b""])
    ph.states[1, 0] = 0                                         # not a state any more
    with pytest.raises(AfcError):
        ph.update([b"", b""], [True, True])


def test_empty_batches(ctx):
    from agentfield_b200 import Auditor
    z8, off0 = np.zeros(1, np.uint8), np.zeros(1, np.uint64)
    assert ctx.sha256_packed(z8, off0).shape == (0, 32)
    assert ctx.verify_packed(np.zeros((0, 32), np.uint8), np.zeros((0, 64), np.uint8), z8, off0).shape == (0,)
    assert ctx.sign_packed(np.zeros((0, 32), np.uint8), z8, off0).shape == (0, 64)
    a = Auditor(ctx)
    assert a.append([]) == (hashlib.sha256(b"").digest(), 0)
    a.close()


def test_argument_validation_through_the_abi(ctx):
    """Bad arguments come back as AFC_EINVAL instead of being dereferenced: offsets that run backwards (a slip of the caller's
    packing code would otherwise become an out-of-bounds copy), key indices past the key set, NULL result buffers."""
    from agentfield_b200 import AfcError, _abi
    lib, H = _abi.load(), ctx.handle
    buf = np.zeros(4096, dtype=np.uint8)
    out = np.zeros((3, 64), dtype=np.uint8)
    bad_off = np.array([0, 512, 256, 1024], dtype=np.uint64)
    seeds = np.zeros((3, 32), dtype=np.uint8)
    assert lib.afc_sha256_batch(H, _abi.ptr(buf), _abi.ptr(bad_off), 3, _abi.ptr(out)) == _abi.AFC_EINVAL
    assert lib.afc_ed25519_sign_batch(H, _abi.ptr(seeds), _abi.ptr(buf), _abi.ptr(bad_off), 3, _abi.ptr(out)) == _abi.AFC_EINVAL
    good_off = np.array([0, 256, 512, 1024], dtype=np.uint64)
    bad_koff = np.array([0, 32, 16, 64], dtype=np.uint32)
    assert lib.afc_hmac_sha256_batch(H, _abi.ptr(buf), _abi.ptr(bad_koff), _abi.ptr(buf), _abi.ptr(good_off), 3, _abi.ptr(out)) == _abi.AFC_EINVAL
    exp = ctx.expand(seeds)
    with pytest.raises(AfcError):
        ctx.sign_expanded_packed(exp, np.array([0, 1, 3], dtype=np.uint32), buf, good_off)          # index 3 of 3 keys
    assert lib.afc_ed25519_verify_batch(H, None, None, _abi.ptr(buf), _abi.ptr(good_off), 3, None) == _abi.AFC_EINVAL
    assert lib.afc_sha256_batch(None, _abi.ptr(buf), _abi.ptr(good_off), 3, _abi.ptr(out)) == _abi.AFC_EINVAL
    assert (ctx.sha256_packed(buf, good_off) == CO.sha256_batch(buf, good_off)).all()              # and the context still works


def test_ed25519_random_parity_ragged(ctx):
    rng = np.random.default_rng(0xAF02)
    n = 3000
    seeds = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    lens = rng.integers(0, 1400, n)
    lens[:6] = [0, 1, 63, 64, 65, 1300]
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    buf = rng.integers(0, 256, int(off[-1]) + 1, dtype=np.uint8)
    sigs = ctx.sign_packed(seeds, buf, off)
    assert (sigs == CO.ed25519_sign_batch(seeds, buf, off, 8)).all()
    pks = ctx.pubkeys(seeds)
    assert (pks == CO.ed25519_pubkey_batch(seeds, 8)).all()
    # corrupt a third of the items in different places (sig R, sig S, pk, message)
    s2, p2, b2 = sigs.copy(), pks.copy(), buf.copy()
    idx = np.arange(n)
    s2[idx % 9 == 0, 5] ^= 0x01
    s2[idx % 9 == 3, 40] ^= 0x80
    p2[idx % 9 == 6, 1] ^= 0x04
    for i in idx[(idx % 9 == 7) & (lens > 0)]:
        b2[int(off[i])] ^= 0x40
    ok = ctx.verify_packed(p2, s2, b2, off)
    exp = CO.ed25519_verify_batch(p2, s2, b2, off, 8)
    assert (ok == exp).all()
    assert 0.55 * n < ok.sum() < 0.7 * n


def test_constant_time_and_fast_signing_agree_with_the_oracle(ctx):
    """E1: the default constant-time fixed-base multiplication (k_ed_sign_ct / k_ed_expand_ct) and the fast gather path produce the
    same signatures and public keys — RFC 8032's and the oracle's — from seeds and from expanded keys, ragged messages."""
    rng = np.random.default_rng(0xAF71)
    n = 3000
    seeds = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    lens = rng.integers(0, 700, n)
    off = np.zeros(n + 1, dtype=np.uint64); off[1:] = np.cumsum(lens)
    buf = rng.integers(0, 256, int(off[-1]) + 1, dtype=np.uint8)
    want = CO.ed25519_sign_batch(seeds, buf, off, 8)
    want_pk = CO.ed25519_pubkey_batch(seeds, 8)
    assert ctx.sign_mode() == "constant-time"                      # the default
    try:
        for ct in (True, False):
            ctx.sign_configure(ct)
            assert (ctx.sign_packed(seeds, buf, off) == want).all(), ct
            assert (ctx.pubkeys(seeds) == want_pk).all(), ct
            exp = ctx.expand(seeds[:64])
            assert (exp[:, 64:] == want_pk[:64]).all()
            ki = rng.integers(0, 64, n).astype(np.uint32)
            assert (ctx.sign_expanded_packed(exp, ki, buf, off) == CO.ed25519_sign_batch(seeds[ki].copy(), buf, off, 8)).all(), ct
        g = golden("rfc8032.json")
        from agentfield_b200 import Signer
        ctx.sign_configure(True)
        assert [s_.hex() for s_ in Signer(ctx).sign_batch([bytes.fromhex(e["seed"]) for e in g], [bytes.fromhex(e["msg"]) for e in g])] == [e["sig"] for e in g]
    finally:
        ctx.sign_configure(True)


def test_expanded_key_cache_sign(ctx):
    rng = np.random.default_rng(0xAF11)
    nk, n = 16, 500
    seeds = rng.integers(0, 256, (nk, 32), dtype=np.uint8)
    exp96 = ctx.expand(seeds)
    assert (exp96[:, 64:] == CO.ed25519_pubkey_batch(seeds)).all()
    ki = (np.arange(n) % nk).astype(np.uint32)
    msgs = rng.integers(0, 256, (n, 512), dtype=np.uint8)
    off = np.arange(n + 1, dtype=np.uint64) * 512
    sigs = ctx.sign_expanded_packed(exp96, ki, msgs.reshape(-1), off)
    assert (sigs == CO.ed25519_sign_batch(seeds[ki], msgs.reshape(-1), off, 8)).all()


def test_merkle_incremental_appends_match_oracle(ctx):
    """Random (size, n) appends: every frontier merge / orphan case of the level schedule."""
    from agentfield_b200 import Auditor
    rng = np.random.default_rng(0xAF04)
    for trial in range(6):
        a = Auditor(ctx)
        leaves = []
        for step in range(10):
            k = int(rng.integers(1, 70)) if trial < 4 else int(rng.integers(1, 3000))
            new = [rng.integers(0, 256, int(rng.integers(0, 150)), dtype=np.uint8).tobytes() for _ in range(k)]
            leaves += new
            root, size = a.append(new)
            assert size == len(leaves) and root == OM.root(leaves), (trial, step, len(leaves))
        st = a.save()                                       # checkpoint / resume
        b = Auditor(ctx)
        b.load(st)
        more = [b"resume-%d" % i for i in range(37)]
        assert b.append(more) == a.append(more) == (OM.root(leaves + more), len(leaves) + 37)
        a.close(); b.close()


def test_merkle_fold_of_aligned_shard_roots(ctx):
    from agentfield_b200 import Auditor, fold_roots, shard
    rng = np.random.default_rng(0xAF14)
    for n in (8, 1000, 4096, 5000):
        leaves = [rng.integers(0, 256, 96, dtype=np.uint8).tobytes() for _ in range(n)]
        full = OM.root(leaves)
        for world in (2, 8):
            roots = []
            for r in range(world):
                lo, hi = shard.merkle_shard_range(n, r, world)
                if hi > lo:
                    a = Auditor(ctx)
                    roots.append(a.append(leaves[lo:hi])[0])
                    a.close()
            assert fold_roots(np.frombuffer(b"".join(roots), dtype=np.uint8), ctx) == full, (n, world)


def test_keyed_verify_equals_generic_verify_and_go_rules(ctx):
    """N1 identity cache: per-key tables must give bit-identical accept/reject decisions, edge keys included."""
    from agentfield_b200 import KeySet
    es = golden("ed25519_edge.json") + [dict(e, valid=True) for e in golden("rfc8032.json")]
    pks = sorted({e["pk"] for e in es})
    ks = KeySet([bytes.fromhex(p) for p in pks], ctx)
    assert ks.info()["n_keys"] == len(pks)
    ok = ks.verify_batch([bytes.fromhex(e["pk"]) for e in es], [bytes.fromhex(e["msg"]) for e in es], [bytes.fromhex(e["sig"]) for e in es])
    for e, o in zip(es, ok):
        assert o == e["valid"], e["name"]
    ks.close()
    # random keys, ragged messages, corruption in every field; out-of-range key index -> 0
    rng = np.random.default_rng(0xAF12)
    nk, n = 64, 12000                                     # >= 4096: the batch is bucketed by key index before the table-driven kernel
    seeds = rng.integers(0, 256, (nk, 32), dtype=np.uint8)
    kpks = ctx.pubkeys(seeds)
    ki = rng.integers(0, nk, n).astype(np.uint32)
    lens = rng.integers(0, 900, n)
    off = np.zeros(n + 1, dtype=np.uint64); off[1:] = np.cumsum(lens)
    buf = rng.integers(0, 256, int(off[-1]) + 1, dtype=np.uint8)
    sigs = ctx.sign_packed(seeds[ki].copy(), buf, off)
    idx = np.arange(n)
    sigs[idx % 7 == 0, 9] ^= 2
    sigs[idx % 7 == 3, 50] ^= 0x40
    ks = KeySet([bytes(p) for p in kpks], ctx)
    got = ks.verify_packed(ki, sigs, buf, off)
    exp = ctx.verify_packed(kpks[ki].copy(), sigs, buf, off)
    assert (got == exp).all() and (exp == CO.ed25519_verify_batch(kpks[ki].copy(), sigs, buf, off, 8)).all()
    ki2 = ki.copy(); ki2[::11] = nk + 5
    got2 = ks.verify_packed(ki2, sigs, buf, off)
    assert not got2[::11].any() and (np.delete(got2, np.arange(0, n, 11)) == np.delete(exp, np.arange(0, n, 11))).all()
    ks.close()


def test_transparent_key_cache_paths(ctx):
    """afc_ed25519_verify_batch with the issuer-key cache: every decision path must give the oracle's bitmap — distinct keys
    (all cold: generic kernel), repeated keys (tables built, then reused), a full cache (least recently used tables evicted),
    more hot keys than the cache holds (some hot, the rest cold IN THE SAME CALL), hot + cold mixes, Go-rule edge keys through
    the table path, and cache disabled."""
    rng = np.random.default_rng(0xAF77)

    def batch(nk, n, seed_off=0, msg_len=200, extra_distinct=0):
        seeds = np.random.default_rng(1000 + seed_off).integers(0, 256, (nk + extra_distinct, 32), dtype=np.uint8)
        kp = ctx.pubkeys(seeds)
        ki = rng.integers(0, nk, n + extra_distinct)
        if extra_distinct:
            ki[rng.permutation(n + extra_distinct)[:extra_distinct]] = nk + np.arange(extra_distinct)     # keys used exactly once
        n = n + extra_distinct
        msgs = rng.integers(0, 256, (n, msg_len), dtype=np.uint8)
        off = np.arange(n + 1, dtype=np.uint64) * msg_len
        sigs = ctx.sign_packed(seeds[ki].copy(), msgs.reshape(-1), off)
        sigs[::13, 4] ^= 8
        pks = kp[ki].copy()
        pks[5::2500, 2] ^= 1                                 # a few corrupted (mostly off-curve / unknown) keys
        return pks, sigs, msgs.reshape(-1), off

    def check(args):
        pks, sigs, buf, off = args
        got = ctx.verify_packed(pks, sigs, buf, off)
        assert (got == CO.ed25519_verify_batch(pks, sigs, buf, off, 8)).all()
        st = ctx.keycache_stats()
        if st["max_keys"]:
            assert st["last_hot"] + st["last_cold"] == len(off) - 1, st
        assert ctx.keycache_info()["last_mode"] == (1 if st["last_hot"] else 0)
        return st

    try:
        ctx.keycache_configure(16)
        st = check(batch(3000, 3000, 1))                      # ~all keys distinct: everything cold
        assert st["last_hot"] == 0 and st["last_built"] == 0 and st["cached_keys"] == 0
        s1 = check(batch(10, 6000, 2))                        # 10 keys (+ a few corrupted ones: used once, cold), heavy reuse
        assert s1["last_built"] == 10 and s1["cached_keys"] == 10 and 0 < s1["last_cold"] <= 3
        b = batch(10, 6000, 2)
        b[0][5::2500, 2] ^= 1                                 # same 10 keys, no corrupted ones this time
        s2 = check(b)                                         # reuse: nothing new to build, nothing cold
        assert s2["last_built"] == 0 and s2["last_cold"] == 0 and s2["cached_keys"] == 10
        s3 = check(batch(12, 9000, 3))                        # 12 other keys: 6 free ids + 6 least recently used tables evicted
        assert s3["last_built"] == 12 and s3["last_evicted"] == 6 and s3["cached_keys"] == 16
        s4 = check(batch(40, 20000, 4))                       # 40 hot keys > capacity 16: 16 get tables, 24 stay cold in the same call
        assert s4["last_built"] == 16 and s4["last_evicted"] == 16 and s4["cached_keys"] == 16
        assert 0.4 * 20000 < s4["last_cold"] < 0.8 * 20000 and s4["last_hot"] > 0.2 * 20000
        s5 = check(batch(8, 8000, 5, extra_distinct=1500))    # 8 hot issuers among 1500 one-off keys
        assert s5["last_built"] == 8 and s5["last_cold"] >= 1500 and s5["last_hot"] >= 7000
        # LRU order: capacity 4; A = {a, b, c, d}; B touches {a, b}; C brings {e, f}: c and d go, a and b stay
        ctx.keycache_configure(4)
        seeds = np.random.default_rng(77).integers(0, 256, (6, 32), dtype=np.uint8)
        kp = ctx.pubkeys(seeds)

        def call(which, n=600):
            ki = np.array(which)[rng.integers(0, len(which), n)]
            msgs = rng.integers(0, 256, (n, 64), dtype=np.uint8)
            off = np.arange(n + 1, dtype=np.uint64) * 64
            sigs = ctx.sign_packed(seeds[ki].copy(), msgs.reshape(-1), off)
            sigs[::9, 40] ^= 1
            return check((kp[ki].copy(), sigs, msgs.reshape(-1), off))
        assert call([0, 1, 2, 3])["last_built"] == 4
        assert call([0, 1])["last_built"] == 0
        sc = call([4, 5])
        assert sc["last_built"] == 2 and sc["last_evicted"] == 2
        assert call([0, 1, 4, 5])["last_built"] == 0          # the recently used pair survived
        sd = call([2, 3])
        assert sd["last_built"] == 2 and sd["last_evicted"] == 2 and sd["total_evicted"] == 4
        ctx.keycache_configure(256)
        es = golden("ed25519_edge.json")
        rep = 120
        pks = np.frombuffer(b"".join(bytes.fromhex(e["pk"]) for e in es) * rep, dtype=np.uint8).reshape(-1, 32).copy()
        sigs = np.frombuffer(b"".join(bytes.fromhex(e["sig"]) for e in es) * rep, dtype=np.uint8).reshape(-1, 64).copy()
        from agentfield_b200 import pack
        buf, off = pack([bytes.fromhex(e["msg"]) for e in es] * rep)
        got = ctx.verify_packed(pks, sigs, buf, off)
        st = ctx.keycache_stats()
        assert st["last_hot"] == len(es) * rep and st["last_cold"] == 0          # every edge key went through a table
        assert (got.reshape(rep, -1) == np.array([e["valid"] for e in es], dtype=np.uint8)).all()
        ctx.keycache_configure(0)
        st = check(batch(10, 6000, 5))
        assert st["last_hot"] == 0 and ctx.keycache_info()["max_keys"] == 0
    finally:
        ctx.keycache_configure(4096)


@pytest.mark.parametrize("knob", ["AFC_DYN_GRID_CAP=1", "AFC_VERIFY_DYNAMIC=0", "AFC_VERIFY_QUAD=1", "AFC_VERIFY_QUAD=2"])
def test_alternative_verify_kernels_same_results(knob):
    """The table-driven kernels that are not the default — the static split (AFC_VERIFY_DYNAMIC=0) and the four-lane experiment
    (AFC_VERIFY_QUAD=1|2: hashing before / fused), DESIGN.md §4 — must give the results of the default ones: the Go edge set, the
    key-set path and every regime of the transparent cache are re-run in a fresh process with the knob set (read once per process).
    AFC_DYN_GRID_CAP=1 runs the default, counter-driven kernels on ONE CTA: a warp then goes through several groups of 32 tiles
    with 12 000 credentials, the path a full grid only takes beyond 2.4 M credentials per call."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    name, value = knob.split("=")
    env = dict(os.environ)
    env[name] = value
    sel = "test_golden_ed25519_edge_set_go_rules or test_keyed_verify_equals_generic_verify_and_go_rules or test_transparent_key_cache_paths or test_ed25519_random_parity_ragged"
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider", "-k", sel],
                       cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "4 passed" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


def test_expanded_key_cache_derivation_matches_reference_flow(ctx):
    from agentfield_b200 import ExpandedKeys
    g = golden("reference_flow.json")
    cache = ExpandedKeys(ctx)
    dids = cache.derive(bytes.fromhex(g["master_seed"]), [d["path"] for d in g["derivations"]])
    assert dids == [d["did"] for d in g["derivations"]]
    vc = g["vc"]
    assert cache.sign_batch([dids[1]], [vc["canonical"].encode()])[0].hex() == vc["sig"]


def test_config1_issue_and_verify_1000_vcs_reference_flow(ctx):
    """BASELINE.json configs[0]: issue + verify 1 000 synthetic agent-action VCs (canonical form padded to a uniform 1536 B;
    the reference's own document cannot be as small as the 512 B the config names — SURVEY.md §8a row V1).  The GPU path (services.VCService over
    the key cache) must produce byte-identical vc_document / signature to the CPU restatement of the reference flow
    (oracle/ref_vc.py, OpenSSL signatures) and accept / reject exactly like it."""
    import json
    from test_host_logic import _vc_requests
    from agentfield_b200 import ExpandedKeys, go_json
    from agentfield_b200.services import VCService, generate_webhook_signature_batch
    from oracle import ref_vc, go_hash as H
    rng = np.random.default_rng(0xAF01)
    master = rng.integers(0, 256, 32, dtype=np.uint8).tobytes()
    paths = ["m/44'/0'"] + ["m/44'/%d'/%d'" % (1237 + a, 0) for a in range(16)] + ["m/44'/%d'/0'/0'/0'" % (1237 + a) for a in range(16)] + \
            ["m/44'/%d'/0'/1'/0'" % (1237 + a) for a in range(16)]
    cache = ExpandedKeys(ctx)
    dids = cache.derive(master, paths)
    seeds = {d: H.derive_seed(master, p) for d, p in zip(dids, paths)}
    assert all(d == H.did_key(CO.pubkey(seeds[d])) for d in dids[:5])
    reqs = _vc_requests(1000, dids[1:], rng, pad_to=1536)     # a real execution VC is ~1.25 KB (5 DIDs, 4 hashes): 512 B is unreachable
    svc = VCService(cache, ctx)
    issued = svc.generate_execution_vc_batch(reqs)
    canon_lens = set()
    for r, got in zip(reqs, issued):
        exp = ref_vc.generate_execution_vc(r, seeds[r["caller_did"]])
        assert got["vc_document"] == exp["vc_document"] and got["signature"] == exp["signature"]
        canon_lens.add(len(exp["canonical"]))
    assert canon_lens == {1536}
    # the same flow with the canonical bytes assembled on the device (values -> documents -> signatures, no host marshalling)
    issued_dev = VCService(cache, ctx, canonical_on_device=True).generate_execution_vc_batch(reqs)
    assert [(v["vc_document"], v["signature"]) for v in issued_dev] == [(v["vc_document"], v["signature"]) for v in issued]
    ok = svc.verify_vc_batch(issued)
    assert all(ok) and all(ref_vc.verify_vc(v["vc_document"], cache.public_key(v["doc"]["issuer"])) for v in issued[:50])
    # tamper: change a field of the parsed document / swap signatures
    bad = [dict(v, doc=json.loads(json.dumps(v["doc"]))) for v in issued[:100]]
    for i, v in enumerate(bad):
        if i % 2:
            v["doc"]["credentialSubject"]["execution"]["status"] = "tampered"
        else:
            v["proof"] = dict(v["proof"], proofValue=issued[(i + 1) % 100]["proof"]["proofValue"])
    assert not any(svc.verify_vc_batch(bad))
    # webhook header shape
    bodies = [go_json.webhook_payload({"event": "execution.completed", "execution_id": r["execution_id"], "workflow_id": r["workflow_id"],
                                       "status": r["status"], "target": "node.fn", "type": "reasoner", "duration_ms": r["duration_ms"],
                                       "result": None, "error_message": r["error_message"], "timestamp": r["timestamp"]}) for r in reqs[:200]]
    secrets = ["secret-%d" % (i % 5) for i in range(200)]
    assert generate_webhook_signature_batch(secrets, bodies, ctx) == [H.webhook_signature(s, b) for s, b in zip(secrets, bodies)]


def test_go_cases_signatures_tags_and_device_canonical_form(ctx):
    """tests/golden/go_cases.json through the GPU: Ed25519 signatures over the canonical bytes of execution and workflow VCs,
    verification of the stored documents, webhook HMAC headers, and the three device templates filled by the kernels."""
    import go_cases_util as U
    from agentfield_b200 import Signer, Verifier, canonical as CA, go_json as GJ
    from agentfield_b200.services import generate_webhook_signature_batch
    g = golden("go_cases.json")
    cases = [(c, c["expect_canonical"].encode("utf-8")) for c in g["execution_vcs"] + g["workflow_vcs"]]
    sigs = Signer(ctx).sign_batch([bytes.fromhex(c["seed"]) for c, _ in cases], [m for _, m in cases])
    assert [s.hex() for s in sigs] == [c["expect_sig"] for c, _ in cases]
    pks = [bytes.fromhex(c["pk"]) for c, _ in cases]
    assert Signer(ctx).public_keys([bytes.fromhex(c["seed"]) for c, _ in cases]) == pks
    assert all(Verifier(ctx).verify_batch(pks, [m for _, m in cases], sigs))
    assert not any(Verifier(ctx).verify_batch(pks, [m + b" " for _, m in cases], sigs))
    assert generate_webhook_signature_batch([c["secret"] for c in g["webhooks"]], [c["expect_body"].encode("utf-8") for c in g["webhooks"]], ctx) == \
        [c["expect_header"] for c in g["webhooks"]]
    # device templates
    docs = [U.execution_doc(c) for c in g["execution_vcs"]]
    t0, t1 = CA.vc_document_template(False, ctx), CA.vc_document_template(True, ctx)
    assert t0.fill([CA.vc_document_values(d) for d in docs]) == [c["expect_canonical"].encode("utf-8") for c in g["execution_vcs"]]
    proofs = [U.proof(c["issuer"], bytes.fromhex(c["expect_sig"]), c["proof_created"]) for c in g["execution_vcs"]]
    assert t1.fill([CA.vc_document_values(d, p) for d, p in zip(docs, proofs)]) == [c["expect_stored"].encode("utf-8") for c in g["execution_vcs"]]
    wdocs = [U.workflow_doc(c) for c in g["workflow_vcs"]]
    w0, w1 = CA.workflow_vc_document_template(False, ctx), CA.workflow_vc_document_template(True, ctx)
    assert w0.fill([CA.workflow_vc_document_values(d) for d in wdocs]) == [c["expect_canonical"].encode("utf-8") for c in g["workflow_vcs"]]
    wproofs = [U.proof(c["issuer_did"], bytes.fromhex(c["expect_sig"]), c["proof_created"]) for c in g["workflow_vcs"]]
    assert w1.fill([CA.workflow_vc_document_values(d, p) for d, p in zip(wdocs, wproofs)]) == [c["expect_stored"].encode("utf-8") for c in g["workflow_vcs"]]
    wt = CA.JsonTemplate(*CA.webhook_payload_template_parts(), ctx)
    assert wt.fill([CA.webhook_payload_values(U.webhook(c)) for c in g["webhooks"]]) == [c["expect_body"].encode("utf-8") for c in g["webhooks"]]


def test_workflow_vc_issue_verify_and_comprehensive_audit(ctx):
    """The workflow-level credential path (vc_service.go:525-718, 1386-1644): roll 64 workflows of 0..40 execution VCs up into
    signed WorkflowVCDocuments on the GPU — byte-identical to the CPU restatement (oracle/ref_vc.py, OpenSSL signatures) —
    verify them from their stored bytes, and run the comprehensive audit of one chain (every signature in one GPU batch):
    clean chain valid with score 100, each kind of tampering reported the way the reference reports it."""
    import json
    from test_host_logic import _vc_requests
    from agentfield_b200 import ExpandedKeys
    from agentfield_b200.services import VCService
    from oracle import ref_vc, go_hash as H
    rng = np.random.default_rng(0xAF61)
    master = rng.integers(0, 256, 32, dtype=np.uint8).tobytes()
    paths = ["m/44'/0'"] + ["m/44'/%d'/0'" % (2000 + a) for a in range(8)]
    cache = ExpandedKeys(ctx)
    dids = cache.derive(master, paths)
    seeds = {d: H.derive_seed(master, p) for d, p in zip(dids, paths)}
    svc = VCService(cache, ctx)
    workflows, oracle_in = [], []
    statuses = ["succeeded", "completed", "failed", "running", "timeout", "pending", "cancelled"]
    for w in range(64):
        m = int(rng.integers(0, 41)) if w else 0
        reqs = _vc_requests(m, dids[1:], rng, pad_to=0)
        issued = svc.generate_execution_vc_batch(reqs) if m else []
        evs = [{"vc_id": r["vc_id"], "execution_id": r["execution_id"], "workflow_id": "wf-%d" % w, "session_id": "sess-%d" % w,
                "status": statuses[int(rng.integers(0, 3 if w % 2 else len(statuses)))], "created_at": "2026-09-21T%02d:%02d:00Z" % (rng.integers(0, 24), rng.integers(0, 60)),
                "issuer_did": r["caller_did"]} for r in reqs]
        wf = {"workflow_id": "wf-%d" % w, "execution_vcs": evs, "root_did": dids[0], "vc_id": "vc-%d" % (10 ** 18 + w), "workflow_vc_id": "vc-%d" % (2 * 10 ** 18 + w),
              "issuance_date": "2026-09-22T00:00:00Z", "snapshot_time": "2026-09-22T00:00:00Z", "proof_created": "2026-09-22T00:00:01Z"}
        workflows.append(wf)
    got = svc.generate_workflow_vc_batch(workflows)
    for wf, g_ in zip(workflows, got):
        evs = wf["execution_vcs"]
        from agentfield_b200.services import normalize_execution_status
        status = ref_vc.determine_workflow_status([normalize_execution_status(e["status"]) for e in evs])
        times = [e["created_at"] for e in evs]
        issuer = evs[0]["issuer_did"] if evs else dids[0]
        o = ref_vc.generate_workflow_vc({"workflow_id": wf["workflow_id"], "session_id": evs[0]["session_id"] if evs else "",
                                         "component_vc_ids": [e["vc_id"] for e in evs], "status": status,
                                         "start_time": min(times) if evs else wf["snapshot_time"],
                                         "end_time": max(times) if evs and status in ("succeeded", "failed", "cancelled", "timeout") else None,
                                         "snapshot_time": wf["snapshot_time"], "issuer_did": issuer, "vc_id": wf["vc_id"],
                                         "issuance_date": wf["issuance_date"], "proof_created": wf["proof_created"]}, seeds[issuer])
        assert g_["vc_document"] == o["vc_document"] and g_["signature"] == o["signature"] and g_["status"] == status
        assert g_["total_steps"] == len(evs) and g_["document_size_bytes"] == len(o["vc_document"])
    assert all(svc.verify_workflow_vc_batch(got))
    assert all(ref_vc.verify_workflow_vc(g_["vc_document"], cache.public_key(g_["issuer_did"])) for g_ in got[:10])
    tampered = [dict(g_, vc_document=g_["vc_document"].replace(b'"snapshotTime":"2026-09-22', b'"snapshotTime":"2026-09-23', 1)) for g_ in got]
    assert not any(svc.verify_workflow_vc_batch(tampered))
    # comprehensive audit of one chain
    reqs = _vc_requests(200, dids[1:], rng, pad_to=0)
    for r in reqs:
        r["workflow_id"], r["session_id"] = "wf-audit", "sess-audit"
    issued = svc.generate_execution_vc_batch(reqs)

    def record(r, v):
        return {"vc_id": r["vc_id"], "execution_id": r["execution_id"], "workflow_id": r["workflow_id"], "session_id": r["session_id"],
                "issuer_did": r["caller_did"], "target_did": r.get("target_did", ""), "caller_did": r["caller_did"], "vc_document": v["vc_document"],
                "signature": v["signature"], "input_hash": v["input_hash"], "output_hash": v["output_hash"], "status": r["status"]}
    comps = [record(r, v) for r, v in zip(reqs, issued)]
    wfvc = svc.generate_workflow_vc_batch([{"workflow_id": "wf-audit", "execution_vcs": [dict(c, created_at="2026-09-21T10:00:00Z") for c in comps],
                                            "root_did": dids[0], "vc_id": "vc-9", "workflow_vc_id": "vc-10", "issuance_date": "2026-09-22T00:00:00Z",
                                            "snapshot_time": "2026-09-22T00:00:00Z", "proof_created": "2026-09-22T00:00:01Z"}])[0]
    res = svc.verify_workflow_vc_comprehensive({"component_vcs": comps, "workflow_vc": wfvc})
    assert res["valid"] and res["overall_score"] == 100.0 and not res["critical_issues"] and not res["warnings"]
    assert res["security_analysis"]["key_validation"] and res["security_analysis"]["security_score"] == 100.0
    # (a) a document altered after signing  (b) metadata that disagrees with the document  (c) a workflow VC signed by someone else
    bad = [dict(c) for c in comps]
    bad[3]["vc_document"] = bad[3]["vc_document"].replace(b'"durationMs":', b'"durationMs":1', 1)
    bad[7]["execution_id"] = "exec-other"
    bad[11]["signature"] = comps[12]["signature"]
    bad[20]["vc_document"] = b"{not json"
    res = svc.verify_workflow_vc_comprehensive({"component_vcs": bad, "workflow_vc": dict(wfvc, vc_document=tampered[0]["vc_document"])})
    kinds = sorted(i["type"] for i in res["critical_issues"])
    assert kinds == sorted(["signature_verification_failed", "execution_id_mismatch", "signature_mismatch", "parse_error",
                            "workflow_signature_verification_failed"]), kinds
    assert not res["valid"] and not res["security_analysis"]["key_validation"] and not res["integrity_checks"]["field_consistency"]
    assert sorted(res["security_analysis"]["tamper_evidence"]) == ["execution_id_inconsistency", "signature_inconsistency"]
    assert res["security_analysis"]["security_score"] == 60.0 and [w["type"] for w in res["warnings"]] == ["tamper_evidence"] * 2
    assert res["overall_score"] == max(0.0, ((100.0 - 25.0 * 5 - 5.0 * 2) + 60.0) / 2.0)


def test_export_bundle_offline_audit(ctx):
    """N4 end to end: issue a workflow of 300 credentials among 1 500 other log entries, export the bundle (EnhancedVCChain shape of
    internal/cli/vc.go:78-97 + the audit_log section), verify it offline in three GPU batches — and see every kind of tampering
    reported: an edited credential, a credential dropped from / never in the log, a forged tree head, a rewritten history."""
    import copy
    import json
    from test_host_logic import _vc_requests
    from agentfield_b200 import ExpandedKeys, export_bundle, verify_bundle
    from agentfield_b200.services import VCService
    from oracle import merkle as OM, go_hash as H
    rng = np.random.default_rng(0xAF64)
    master = rng.integers(0, 256, 32, dtype=np.uint8).tobytes()
    cache = ExpandedKeys(ctx)
    dids = cache.derive(master, ["m/44'/0'"] + ["m/44'/%d'/0'" % (3000 + a) for a in range(6)])
    svc = VCService(cache, ctx)
    reqs = _vc_requests(300, dids[1:], rng, pad_to=0)
    for r in reqs:
        r["workflow_id"], r["session_id"] = "wf-export", "sess-export"
    issued = svc.generate_execution_vc_batch(reqs)
    evs = [{"vc_id": r["vc_id"], "execution_id": r["execution_id"], "workflow_id": r["workflow_id"], "session_id": r["session_id"],
            "issuer_did": r["caller_did"], "target_did": r.get("target_did", ""), "caller_did": r["caller_did"], "vc_document": v["vc_document"],
            "signature": v["signature"], "input_hash": v["input_hash"], "output_hash": v["output_hash"], "status": r["status"],
            "created_at": "2026-09-21T10:00:00Z"} for r, v in zip(reqs, issued)]
    wfvc = svc.generate_workflow_vc_batch([{"workflow_id": "wf-export", "execution_vcs": evs, "root_did": dids[0], "vc_id": "vc-77", "workflow_vc_id": "vc-78",
                                            "issuance_date": "2026-09-22T00:00:00Z", "snapshot_time": "2026-09-22T00:00:00Z",
                                            "proof_created": "2026-09-22T00:00:01Z"}])[0]
    # the audit log: 700 unrelated leaves, then this workflow's documents interleaved with 800 more
    other = [rng.integers(0, 256, int(rng.integers(40, 400)), dtype=np.uint8).tobytes() for _ in range(1500)]
    log, where = list(other[:700]), {}
    for i, e in enumerate(evs):
        where[e["vc_id"]] = len(log); log.append(e["vc_document"])
        log += other[700 + 2 * i: 702 + 2 * i]
    log += other[700 + 2 * len(evs):]
    pks = {d: cache.public_key(d) for d in dids}
    bundle = export_bundle("wf-export", evs, wfvc, pks, log, where, lambda m: cache.sign_batch([dids[0]], [m])[0], dids[0], "2026-09-22T01:00:00Z",
                           checkpoints=(1, 700, 1000, len(log)), ctx=ctx)
    # the exported head is the RFC 6962 root of the whole log (oracle), and the file survives a JSON round trip
    sth = bundle["audit_log"]["signed_tree_head"]
    assert sth["tree_size"] == len(log) and sth["root_hash"] == H.b64url_nopad(OM.root(log))
    text = json.dumps(bundle)
    res = verify_bundle(text, ctx)
    assert res["valid"] and res["signature_valid"] and res["format_valid"] and res["type"] == "workflow", res.get("error")
    assert res["summary"] == {"total_components": 300, "valid_components": 300, "total_dids": len(res["did_resolutions"]),
                              "resolved_dids": len(res["did_resolutions"]), "total_signatures": 302, "valid_signatures": 302}
    assert res["audit_log"]["included"] == 300 and res["audit_log"]["checkpoints_ok"] and len(res["audit_log"]["checkpoints"]) == 4
    # a checkpoint root the auditor already trusts must match
    trusted = {700: OM.root(log[:700])}
    assert verify_bundle(text, ctx, trusted_checkpoints=trusted)["valid"]
    assert not verify_bundle(text, ctx, trusted_checkpoints={700: OM.root(log[:699] + [b"x"])})["valid"]

    def broken(mutate):
        b = copy.deepcopy(bundle); mutate(b); return verify_bundle(json.dumps(b), ctx)
    r1 = broken(lambda b: b["execution_vcs"][5]["vc_document"]["credentialSubject"]["execution"].__setitem__("durationMs", 123456789))
    assert not r1["valid"] and not r1["component_results"][5]["signature_valid"] and r1["audit_log"]["not_included"] == [evs[5]["vc_id"]]
    assert r1["summary"]["valid_signatures"] == 301 and r1["summary"]["valid_components"] == 299
    r2 = broken(lambda b: b["audit_log"]["entries"][9]["inclusion_proof"].__setitem__(0, b["audit_log"]["entries"][10]["inclusion_proof"][0]))
    assert not r2["valid"] and r2["audit_log"]["not_included"] == [evs[9]["vc_id"]] and r2["summary"]["valid_signatures"] == 302
    r3 = broken(lambda b: b["audit_log"]["entries"].pop(0))
    assert not r3["valid"] and r3["audit_log"]["not_included"] == [evs[0]["vc_id"]]
    r4 = broken(lambda b: b["audit_log"]["signed_tree_head"].__setitem__("tree_size", len(log) - 1))
    assert not r4["valid"] and not r4["audit_log"]["tree_head"]["signature_valid"]
    r5 = broken(lambda b: b["audit_log"]["checkpoints"][1].__setitem__("root_hash", b["audit_log"]["checkpoints"][2]["root_hash"]))
    assert not r5["valid"] and not r5["audit_log"]["checkpoints_ok"] and r5["summary"]["valid_signatures"] == 302
    r6 = broken(lambda b: b["did_resolution_bundle"].pop(dids[2]))
    assert not r6["valid"] and r6["summary"]["resolved_dids"] == r6["summary"]["total_dids"] - 1
    r7 = broken(lambda b: b["execution_vcs"][3].__setitem__("execution_id", "exec-forged"))
    assert not r7["valid"] and "Execution ID mismatch" in r7["component_results"][3]["error"]
    r8 = broken(lambda b: b["workflow_vc"]["vc_document"]["credentialSubject"].__setitem__("totalSteps", 299))
    assert not r8["valid"] and not r8["workflow_verification"]["signature_valid"]
    assert not verify_bundle("{not json", ctx)["format_valid"] and not verify_bundle({"no": "workflow"}, ctx)["format_valid"]


def test_ingest_dispatcher_results_and_audit_log(ctx):
    """N2 / configs[4] shape: actions submitted one by one come back with the oracle's signature and tag, and the dispatcher's
    audit log root is the RFC 6962 root over the signatures in ticket order, whatever the batch boundaries were."""
    from agentfield_b200 import Ingest
    rng = np.random.default_rng(0xAF05)
    nk = 8
    seeds = rng.integers(0, 256, (nk, 32), dtype=np.uint8)
    ing = Ingest(ctx.expand(seeds), ctx, batch_max=256, linger_us=300, max_msg=1500, max_key=64, max_body=600)
    n = 3000
    acts, tickets = [], []
    for i in range(n):
        k = int(rng.integers(0, nk))
        msg = rng.integers(0, 256, int(rng.integers(0, 1400)), dtype=np.uint8).tobytes()
        hk = rng.integers(0, 256, int(rng.integers(1, 64)), dtype=np.uint8).tobytes()
        body = rng.integers(0, 256, int(rng.integers(0, 500)), dtype=np.uint8).tobytes()
        acts.append((k, msg, hk, body))
        tickets.append(ing.submit(k, msg, hk, body))
    assert tickets == list(range(n))
    sigs = []
    for (k, msg, hk, body), t in zip(acts, tickets):
        sig, tag = ing.wait(t)
        assert sig == CO.sign(seeds[k].tobytes(), msg) and tag == CO.hmac_sha256(hk, body), t
        sigs.append(sig)
    st = ing.stats()
    assert st["completed"] == n and st["log_size"] == n and st["batches"] >= n // 256 and st["last_error"] == 0
    assert st["log_root"] == OM.root(sigs).hex()
    ing.close()


def test_ingest_soak_sustains_target_rate(ctx):
    """Open-loop Poisson load at the configs[4] rate (100 k actions/s) for a short window on one GPU: nothing dropped,
    everything completed, latency bounded by linger + one batch."""
    from agentfield_b200 import Ingest
    rng = np.random.default_rng(0xAF05)
    ing = Ingest(ctx.expand(rng.integers(0, 256, (64, 32), dtype=np.uint8)), ctx, batch_max=4096, linger_us=500, max_msg=512, max_key=32, max_body=256)
    r = ing.soak(100_000, 3.0, producers=4)
    assert r["completed"] == r["submitted"] == r["log_size"] and r["last_error"] == 0
    assert r["achieved_rate"] > 90_000, r
    assert r["p99_us"] < 20_000, r
    ing.close()


def test_concurrent_callers_share_one_context(ctx):
    """The C ABI is thread-safe and re-entrant (reference callers are concurrent goroutines, one per HTTP request;
    webhook_dispatcher.go:118-121 runs 4 workers): 8 threads hammer one afc_ctx with different operations."""
    import threading
    rng = np.random.default_rng(0xAF55)
    n = 3000
    seeds = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    msgs = rng.integers(0, 256, (n, 300), dtype=np.uint8)
    off = np.arange(n + 1, dtype=np.uint64) * 300
    exp_sig = CO.ed25519_sign_batch(seeds, msgs.reshape(-1), off, 8)
    exp_dig = CO.sha256_batch(msgs.reshape(-1), off, 8)
    pks = CO.ed25519_pubkey_batch(seeds, 8)
    errs = []

    def work(t):
        try:
            for rep in range(3):
                if t % 3 == 0:
                    assert (ctx.sign_packed(seeds, msgs.reshape(-1), off) == exp_sig).all()
                elif t % 3 == 1:
                    assert ctx.verify_packed(pks, exp_sig, msgs.reshape(-1), off).all()
                else:
                    assert (ctx.sha256_packed(msgs.reshape(-1), off) == exp_dig).all()
        except Exception as e:     # noqa: BLE001
            errs.append((t, repr(e)))
    ts = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs


# ----------------------------------------------------------------------------- device-pointer (resident) variants
def test_device_resident_variants_equal_host_variants(ctx):
    import torch
    rng = np.random.default_rng(0xAF21)
    n = 2048
    seeds = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    msgs = rng.integers(0, 256, (n, 512), dtype=np.uint8)
    off = np.arange(n + 1, dtype=np.uint64) * 512
    dev = torch.device("cuda", 0)
    d_seeds = torch.from_numpy(seeds).to(dev)
    d_msgs = torch.from_numpy(msgs.reshape(-1)).to(dev)
    d_off = torch.from_numpy(off.view(np.int64)).to(dev)
    d_sigs = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    d_pks = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    d_ok = torch.empty(n, dtype=torch.uint8, device=dev)
    d_dig = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    ctx.sign_dev(d_seeds, d_msgs, d_off, n, d_sigs)
    d_exp = torch.empty((n, 96), dtype=torch.uint8, device=dev)
    ctx.expand_dev(d_seeds, n, d_exp)
    d_pks.copy_(d_exp[:, 64:])
    ctx.verify_dev(d_pks, d_sigs, d_msgs, d_off, n, d_ok)
    ctx.sha256_dev(d_msgs, d_off, n, d_dig)
    torch.cuda.synchronize()
    sigs = ctx.sign_packed(seeds, msgs.reshape(-1), off)
    assert (d_sigs.cpu().numpy() == sigs).all() and d_ok.cpu().numpy().all()
    assert (d_dig.cpu().numpy() == CO.sha256_batch(msgs.reshape(-1), off, 4)).all()
    d_sigs2 = torch.empty_like(d_sigs)
    ctx.sign_expanded_dev(d_exp, None, d_msgs, d_off, n, d_sigs2)
    torch.cuda.synchronize()
    assert torch.equal(d_sigs, d_sigs2)


def test_merkle_audit_proofs_match_rfc6962(ctx):
    """N4: audit paths read out of the materialised tree == the recursive RFC 6962 definition; bulk verification accepts them and
    rejects a wrong leaf, a wrong index, a truncated path and a wrong root."""
    from agentfield_b200 import MerkleTree, verify_inclusion_batch, Auditor
    rng = np.random.default_rng(0xAF44)
    for n in (1, 2, 3, 7, 8, 100, 1000, 4097):
        hs = [rng.integers(0, 256, 32, dtype=np.uint8).tobytes() for _ in range(n)]
        t = MerkleTree(np.frombuffer(b"".join(hs), dtype=np.uint8), ctx)
        assert t.root == OM.root_from_leaf_hashes(hs)
        a = Auditor(ctx); assert a.append_hashes(np.frombuffer(b"".join(hs), dtype=np.uint8))[0] == t.root; a.close()
        idx = sorted({0, n - 1, n // 2, n // 3, min(n - 1, 5)})
        proofs = t.inclusion_proofs(idx)
        for m, p in zip(idx, proofs):
            assert p == OM.inclusion_proof(hs, m), (n, m)
            assert OM.verify_inclusion(hs[m], m, n, p, t.root)
        ok = verify_inclusion_batch([hs[m] for m in idx], idx, n, proofs, t.root, ctx)
        assert ok.all(), n
        if n > 3:
            assert not verify_inclusion_batch([hs[(m + 1) % n] for m in idx], idx, n, proofs, t.root, ctx).any()
            assert not verify_inclusion_batch([hs[m] for m in idx], [(m + 1) % n for m in idx], n, proofs, t.root, ctx).all()
            assert not verify_inclusion_batch([hs[m] for m in idx], idx, n, [p[:-1] for p in proofs], t.root, ctx).any()
            assert not verify_inclusion_batch([hs[m] for m in idx], idx, n, proofs, bytes(32), ctx).any()
        t.close()


def test_merkle_consistency_proofs_match_rfc6962(ctx):
    """RFC 6962 §2.1.2 consistency proofs read out of the device tree == the published CT vectors and the oracle's recursive
    construction; the device verifier (RFC 9162 §2.1.4.2) accepts them and rejects tampered ones."""
    from agentfield_b200.audit import MerkleTree, verify_consistency_batch
    g = golden("rfc6962.json")
    hs = [OM.leaf_hash(bytes.fromhex(x)) for x in g["leaves"]]
    for c in g["consistency"]:
        t = MerkleTree(np.frombuffer(b"".join(hs[:c["second"]]), dtype=np.uint8), ctx)
        assert [x.hex() for x in t.consistency_proof(c["first"])] == c["proof"], c
        ok = verify_consistency_batch([c["first"]], [bytes.fromhex(g["roots"][c["first"] - 1])], c["second"],
                                      bytes.fromhex(g["roots"][c["second"] - 1]), [[bytes.fromhex(x) for x in c["proof"]]], ctx)
        assert ok.all(), c
        t.close()
    rng = np.random.default_rng(0xAF63)
    for n in (1, 2, 3, 5, 8, 13, 64, 100, 1000, 4097):
        hs = [rng.integers(0, 256, 32, dtype=np.uint8).tobytes() for _ in range(n)]
        t = MerkleTree(np.frombuffer(b"".join(hs), dtype=np.uint8), ctx)
        firsts = sorted({1, n, max(1, n // 2), max(1, n - 1), max(1, n // 3), min(n, 4), min(n, 7)})
        proofs = [t.consistency_proof(m) for m in firsts]
        roots = [OM.root_from_leaf_hashes(hs[:m]) for m in firsts]
        for m, p, r in zip(firsts, proofs, roots):
            assert p == OM.consistency_proof(hs, m), (n, m)
            assert OM.verify_consistency(m, n, r, t.root, p)
        assert verify_consistency_batch(firsts, roots, n, t.root, proofs, ctx).all(), n
        bad_roots = [bytes(32)] * len(firsts)
        got = verify_consistency_batch(firsts, bad_roots, n, t.root, proofs, ctx)
        assert not got.any(), n
        if n > 3:
            # dropping a node, a wrong new root, or a size outside (0, n] must all fail; first == n only passes with an empty proof
            cut = [p[:-1] if p else [bytes(32)] for p in proofs]
            assert not verify_consistency_batch(firsts, roots, n, t.root, cut, ctx).any()
            assert not verify_consistency_batch(firsts, roots, n, bytes(32), proofs, ctx).any()
            assert not verify_consistency_batch([0, n + 1], [roots[0], roots[0]], n, t.root, [[], []], ctx).any()
        t.close()


def test_device_canonical_form_equals_go_json(ctx):
    """afc_json_fill_*_dev: documents assembled on the GPU from the VCDocument template == the host restatement of
    json.Marshal (go_json.vc_document), with and without proof, for values full of characters Go escapes; raw byte strings
    (invalid UTF-8 included) == the byte-level oracle; sizes / offsets consistent; then sign what the GPU assembled."""
    import torch
    from agentfield_b200 import canonical as CA, go_json as GJ
    from oracle import go_json as OJ
    rng = np.random.default_rng(0xAF35)
    alphabet = list("abcXYZ019 -_:/.") + ['"', "\\", "<", ">", "&", "\n", "\t", "\r", "\b", "\f", "\x00", "\x1f", "\x7f", "é", "ß", "中", "😀", "\u2028", "\u2029"]

    def word(lo=0, hi=24):
        return "".join(alphabet[int(i)] for i in rng.integers(0, len(alphabet), int(rng.integers(lo, hi))))

    docs, proofs = [], []
    for i in range(700):
        docs.append({"@context": ["https://www.w3.org/2018/credentials/v1", word()], "type": ["VerifiableCredential", word()] if i % 5 else None,
                     "id": "urn:agentfield:vc:" + word(), "issuer": "did:key:z" + word(1), "issuanceDate": word(),
                     "credentialSubject": {"executionId": word(0, 200 if i % 7 == 0 else 24), "workflowId": word(), "sessionId": word(),
                                           "caller": {"did": word(), "type": word(), "agentNodeDid": word()},
                                           "target": {"did": word(), "agentNodeDid": word(), "functionName": word()},
                                           "execution": {"inputHash": word(), "outputHash": word(), "timestamp": word(),
                                                         "durationMs": int(rng.integers(0, 2**31)), "status": word(),
                                                         "errorMessage": word(1) if i % 3 == 0 else ""},
                                           "audit": {"inputDataHash": word(), "outputDataHash": word(),
                                                     "metadata": {"agentfield_version": "1.0.0", word(1): word()} if i % 4 else None}}})
        proofs.append({"type": "Ed25519Signature2020", "created": word(), "verificationMethod": word() + "#key-1", "proofPurpose": "assertionMethod",
                       "proofValue": word()})
    t0, t1 = CA.vc_document_template(False, ctx), CA.vc_document_template(True, ctx)
    got0 = t0.fill([CA.vc_document_values(d) for d in docs])
    got1 = t1.fill([CA.vc_document_values(d, p) for d, p in zip(docs, proofs)])
    for d, p, a, b in zip(docs, proofs, got0, got1):
        assert a == GJ.vc_document(d)
        assert b == GJ.vc_document(d, p)
    # byte-level: arbitrary (also invalid) UTF-8 through a 2-value template, every output alignment
    pool = [bytes([i]) for i in range(256)] + ["中".encode(), "😀".encode(), b"\xe2\x80\xa8", b"\xed\xa0\x80", b"\xf4\x90\x80\x80"]
    segs, kinds = [b"[", b"|", b"]"], [CA.STRING, CA.RAW]
    t2 = CA.JsonTemplate(segs, kinds, ctx)
    items = [[b"".join(pool[int(i)] for i in rng.integers(0, len(pool), int(rng.integers(0, 70)))),
              b"".join(pool[int(i)] for i in rng.integers(0, len(pool), int(rng.integers(0, 9))))] for _ in range(3000)]
    got = t2.fill(items)
    assert got == [OJ.fill_template(segs, kinds, it) for it in items]
    assert t2.fill([]) == [] and CA.JsonTemplate([b"const"], [], ctx).fill([[], []]) == [b"const", b"const"]
    # the assembled bytes feed the signer without leaving the device
    dev = torch.device("cuda", 0)
    fields, off = t0.pack_values([CA.vc_document_values(d) for d in docs])
    d_out, d_off = t0.fill_dev(torch.from_numpy(fields).to(dev), torch.from_numpy(off.view(np.int64)).to(dev), len(docs))
    seeds = rng.integers(0, 256, (len(docs), 32), dtype=np.uint8)
    d_sigs = torch.empty((len(docs), 64), dtype=torch.uint8, device=dev)
    ctx.sign_dev(torch.from_numpy(seeds).to(dev), d_out, d_off, len(docs), d_sigs)
    torch.cuda.synchronize()
    sigs = d_sigs.cpu().numpy()
    for i in (0, 1, 350, 699):
        assert sigs[i].tobytes() == CO.sign(seeds[i].tobytes(), GJ.vc_document(docs[i]))


def test_device_text_codecs(ctx):
    """base64url (no padding) and lowercase hex of fixed-size records == Go's base64.RawURLEncoding / hex.EncodeToString."""
    import base64
    import torch
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0xAF33)
    for item in (1, 2, 3, 31, 32, 33, 34, 64):
        n = 257
        raw = rng.integers(0, 256, (n, item), dtype=np.uint8)
        d_in = torch.from_numpy(raw).to(dev)
        olen = (item * 4 + 2) // 3
        d_b64 = torch.zeros((n, olen), dtype=torch.uint8, device=dev)
        d_hex = torch.zeros((n, 2 * item), dtype=torch.uint8, device=dev)
        ctx.b64url_encode_dev(d_in, item, n, d_b64)
        ctx.hex_encode_dev(d_in, item, n, d_hex)
        torch.cuda.synchronize()
        b64, hx = d_b64.cpu().numpy(), d_hex.cpu().numpy()
        for i in (0, 1, n - 1):
            assert b64[i].tobytes() == base64.urlsafe_b64encode(raw[i].tobytes()).rstrip(b"="), (item, i)
            assert hx[i].tobytes() == raw[i].tobytes().hex().encode(), (item, i)


# ----------------------------------------------------------------------------- BASELINE.json sizes: properties
def test_full_size_verify_properties_config2(ctx):
    """cfg2 (1 M x 512 B, K = 1024 keys, 1 % corrupted): sign->verify round trip at full size; the ok-bitmap must equal
    the corruption pattern exactly, and a 20 000-item sample must equal the oracle item by item."""
    import torch
    n, K = 1_000_000, 1024
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(0xAF02)
    d_msgs = torch.randint(0, 256, (n, 512), dtype=torch.uint8, device=dev, generator=g)
    d_kseeds = torch.randint(0, 256, (K, 32), dtype=torch.uint8, device=dev, generator=g)
    d_exp = torch.empty((K, 96), dtype=torch.uint8, device=dev)
    ctx.expand_dev(d_kseeds, K, d_exp)
    d_ki = (torch.arange(n, device=dev) % K).to(torch.int32)
    d_off = (torch.arange(n + 1, device=dev, dtype=torch.int64) * 512)
    d_sigs = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    ctx.sign_expanded_dev(d_exp, d_ki, d_msgs.view(-1), d_off, n, d_sigs)
    d_pks = d_exp[:, 64:][d_ki.long()].contiguous()
    # at this size k_ed_sign shares one inversion between 4 credentials per thread (strided by T = n / 4): signatures from every
    # stride position, and from the seed path (8 points per inversion) on a 400 000-item slice, equal the oracle's
    ks_h = d_kseeds.cpu().numpy()
    T4 = (n + 3) // 4
    sample = sorted({0, 1, 31, 32, 127, 128, T4 - 1, T4, T4 + 1, 2 * T4 - 1, 2 * T4, 3 * T4, 3 * T4 + 77, n - 2, n - 1, 123457, 765432})
    sg_h, ms_s = d_sigs[sample].cpu().numpy(), d_msgs[sample].cpu().numpy()
    for j, i in enumerate(sample):
        assert sg_h[j].tobytes() == CO.sign(ks_h[i % K].tobytes(), ms_s[j].tobytes()), i
    ns = 400_000
    d_seeds_full = d_kseeds[d_ki[:ns].long()].contiguous()
    d_sigs2 = torch.empty((ns, 64), dtype=torch.uint8, device=dev)
    ctx.sign_dev(d_seeds_full, d_msgs.view(-1), d_off[:ns + 1], ns, d_sigs2)
    assert torch.equal(d_sigs2, d_sigs[:ns])
    d_exp2 = torch.empty((ns, 96), dtype=torch.uint8, device=dev)
    ctx.expand_dev(d_seeds_full, ns, d_exp2)                      # grouped key expansion == the 1024 keys expanded one per thread
    assert torch.equal(d_exp2, d_exp[d_ki[:ns].long()])
    del d_sigs2, d_seeds_full, d_exp2
    idx = torch.arange(n, device=dev)
    flip_msg = idx % 100 == 0
    flip_s = idx % 100 == 50
    bitpos = (idx // 100) % 4096
    rows, cols = idx[flip_msg], bitpos[flip_msg] // 8
    d_msgs[rows, cols] = d_msgs[rows, cols] ^ (1 << (bitpos[flip_msg] % 8)).to(torch.uint8)
    srows = idx[flip_s]
    d_sigs[srows, 33] = d_sigs[srows, 33] ^ 0x08
    d_ok = torch.empty(n, dtype=torch.uint8, device=dev)
    ctx.verify_dev(d_pks, d_sigs, d_msgs.view(-1), d_off, n, d_ok)
    torch.cuda.synchronize()
    expect = (~(flip_msg | flip_s)).to(torch.uint8)
    assert torch.equal(d_ok, expect)
    assert int(d_ok.sum()) == n - 2 * (n // 100)
    # oracle sample
    m = 20_000
    pk_h, sg_h, ms_h = d_pks[:m].cpu().numpy(), d_sigs[:m].cpu().numpy(), d_msgs[:m].cpu().numpy()
    off_h = np.arange(m + 1, dtype=np.uint64) * 512
    assert (CO.ed25519_verify_batch(pk_h, sg_h, ms_h.reshape(-1), off_h, 8) == d_ok[:m].cpu().numpy()).all()
    # the host-buffer call (chunked H2D / compute / D2H ring, several chunks, a ragged last one) returns the same bitmap,
    # with a cold issuer-key cache (tables built while the first chunk is in flight) and with a warm one
    nh = 300_017
    pk_h, sg_h = d_pks[:nh].cpu().numpy(), d_sigs[:nh].cpu().numpy()
    ms_h = d_msgs[:nh].cpu().numpy().reshape(-1)
    off_h = np.arange(nh + 1, dtype=np.uint64) * 512
    ctx.keycache_clear(); torch.cuda.synchronize()
    for _ in range(2):
        got = ctx.verify_packed(pk_h, sg_h, ms_h, off_h)
        assert (np.asarray(got).astype(np.uint8) == expect[:nh].cpu().numpy()).all()


def test_three_million_credentials_cross_the_group_boundary(ctx):
    """The counter-driven table kernels run groups of up to 32 tiles per warp; a full grid (75 776 threads on a B200) only starts a
    SECOND group beyond 2.4 M credentials in one launch.  3 M short credentials from 2048 keys, 1 in 64 corrupted, through the
    transparent cache and the key-set path: the bitmap must equal the corruption pattern exactly."""
    import torch
    from agentfield_b200 import KeySet
    n, K, L = 3_000_000, 2048, 48
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(0xAF66)
    d_msgs = torch.randint(0, 256, (n, L), dtype=torch.uint8, device=dev, generator=g)
    d_kseeds = torch.randint(0, 256, (K, 32), dtype=torch.uint8, device=dev, generator=g)
    d_exp = torch.empty((K, 96), dtype=torch.uint8, device=dev)
    ctx.expand_dev(d_kseeds, K, d_exp)
    d_ki = ((torch.arange(n, device=dev) * 7919) % K).to(torch.int32)
    d_off = (torch.arange(n + 1, device=dev, dtype=torch.int64) * L)
    d_sigs = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    ctx.sign_expanded_dev(d_exp, d_ki, d_msgs.view(-1), d_off, n, d_sigs)
    idx = torch.arange(n, device=dev)
    bad_m, bad_r = idx % 64 == 5, idx % 64 == 37
    d_msgs[idx[bad_m], 3] ^= 0x10
    d_sigs[idx[bad_r], 7] ^= 0x01
    expect = (~(bad_m | bad_r)).to(torch.uint8)
    d_pks = d_exp[:, 64:][d_ki.long()].contiguous()
    d_ok = torch.zeros(n, dtype=torch.uint8, device=dev)
    for _ in range(2):                                            # cold (tables built inside the call), then cached
        d_ok.zero_()
        ctx.verify_dev(d_pks, d_sigs, d_msgs.view(-1), d_off, n, d_ok)
        torch.cuda.synchronize()
        assert torch.equal(d_ok, expect)
    ks = KeySet([bytes(p) for p in d_exp[:, 64:].cpu().numpy()], ctx)
    d_ok.zero_()
    ks.verify_dev(d_ki, d_sigs, d_msgs.view(-1), d_off, n, d_ok)
    torch.cuda.synchronize()
    assert torch.equal(d_ok, expect)
    ks.close()
    # and a sample against the oracle, so that the signatures themselves are not taken on trust
    sample = [0, 1, 5, 37, 63, 64, n // 2, n - 1, 2_424_833, 2_900_001]
    pk_h, sg_h, ms_h = d_pks[sample].cpu().numpy(), d_sigs[sample].cpu().numpy(), d_msgs[sample].cpu().numpy()
    for j, i in enumerate(sample):
        assert CO.verify(pk_h[j].tobytes(), ms_h[j].tobytes(), sg_h[j].tobytes()) == bool(expect[i].item()), i


def test_full_size_hmac_properties_config3(ctx):
    """cfg3 (10 M x 256 B bodies, 32 B per-message keys), processed resident in 2 M slices: a sampled slice equals the
    oracle; tags are a deterministic function of (key, body) (same inputs twice -> same tags; one flipped bit -> new tag)."""
    import torch
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(0xAF03)
    per, total = 2_000_000, 10_000_000
    d_off = torch.arange(per + 1, device=dev, dtype=torch.int64) * 256
    d_koff = (torch.arange(per + 1, device=dev, dtype=torch.int64) * 32).to(torch.int32)
    xor_all = torch.zeros(32, dtype=torch.uint8, device=dev)
    for sl in range(total // per):
        d_bodies = torch.randint(0, 256, (per, 256), dtype=torch.uint8, device=dev, generator=g)
        d_keys = torch.randint(0, 256, (per, 32), dtype=torch.uint8, device=dev, generator=g)
        d_tags = torch.empty((per, 32), dtype=torch.uint8, device=dev)
        ctx.hmac_sha256_dev(d_keys.view(-1), d_koff, d_bodies.view(-1), d_off, per, d_tags)
        if sl == 0:
            m = 30_000
            exp = CO.hmac_sha256_batch(d_keys[:m].cpu().numpy().reshape(-1), np.arange(m + 1, dtype=np.uint32) * 32,
                                       d_bodies[:m].cpu().numpy().reshape(-1), np.arange(m + 1, dtype=np.uint64) * 256, 8)
            assert (d_tags[:m].cpu().numpy() == exp).all()
            d_tags2 = torch.empty_like(d_tags)
            ctx.hmac_sha256_dev(d_keys.view(-1), d_koff, d_bodies.view(-1), d_off, per, d_tags2)
            assert torch.equal(d_tags, d_tags2)
            d_bodies[::1000, 255] ^= 1
            ctx.hmac_sha256_dev(d_keys.view(-1), d_koff, d_bodies.view(-1), d_off, per, d_tags2)
            diff = (d_tags != d_tags2).any(dim=1)
            assert int(diff.sum()) == len(range(0, per, 1000)) and bool(diff[::1000].all())
        xor_all ^= torch.from_numpy(np.bitwise_xor.reduce(d_tags.cpu().numpy(), axis=0)).to(dev)
    torch.cuda.synchronize()
    assert xor_all.cpu().numpy().any()


def test_full_size_merkle_config4_single_gpu(ctx):
    """cfg4 shape on one GPU: 2^20 and a non-power-of-two leaf count; root == oracle root over the same leaf hashes;
    8 aligned shard roots fold to the same root."""
    import torch
    from agentfield_b200 import Auditor, fold_roots, shard
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(0xAF04)
    for n in (1 << 20, 1_000_003):
        d_leaves = torch.randint(0, 256, (n, 96), dtype=torch.uint8, device=dev, generator=g)
        d_off = torch.arange(n + 1, device=dev, dtype=torch.int64) * 96
        a = Auditor(ctx)
        a.append_dev(d_leaves.view(-1), d_off, n)
        root, size = a.root()
        assert size == n
        d_lh = torch.empty((n, 32), dtype=torch.uint8, device=dev)
        ctx.merkle_leaf_hashes_dev(d_leaves.view(-1), d_off, n, d_lh)
        torch.cuda.synchronize()
        lh = d_lh.cpu().numpy()
        assert lh[5].tobytes() == OM.leaf_hash(d_leaves[5].cpu().numpy().tobytes())
        assert CO.merkle_root_from_hashes(lh) == root
        roots = []
        for r in range(8):
            lo, hi = shard.merkle_shard_range(n, r, 8)
            b = Auditor(ctx)
            b.append_hashes_dev(d_lh[lo:hi], hi - lo)
            roots.append(b.root()[0])
            b.close()
        assert fold_roots(np.frombuffer(b"".join(roots), dtype=np.uint8), ctx) == root
        a.close()


def test_full_size_config4_sign_4M_and_merkle_append(ctx):
    """cfg4 at full size on one GPU: 4 194 304 (2^22) and 4 000 000 credentials of 512 B signed with 1 024 cached keys, every
    signature verified back (table path), signatures appended as audit leaves; the root equals the oracle's RFC 6962 root over
    the same signatures, and a 20 000-signature sample equals the oracle's signatures byte for byte."""
    import torch
    from agentfield_b200 import Auditor
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(0xAF04)
    K = 1024
    rng = np.random.default_rng(0xAF04)
    kseeds = rng.integers(0, 256, (K, 32), dtype=np.uint8)
    d_exp = torch.empty((K, 96), dtype=torch.uint8, device=dev)
    ctx.expand_dev(torch.from_numpy(kseeds).to(dev), K, d_exp)
    for n in (1 << 22, 4_000_000):
        d_msgs = torch.randint(0, 256, (n, 512), dtype=torch.uint8, device=dev, generator=g)
        d_off = torch.arange(n + 1, device=dev, dtype=torch.int64) * 512
        d_ki = (torch.arange(n, device=dev) % K).to(torch.int32)
        d_sigs = torch.empty((n, 64), dtype=torch.uint8, device=dev)
        ctx.sign_expanded_dev(d_exp, d_ki, d_msgs.view(-1), d_off, n, d_sigs)
        d_pks = d_exp[:, 64:][d_ki.long()].contiguous()
        d_ok = torch.empty(n, dtype=torch.uint8, device=dev)
        ctx.verify_dev(d_pks, d_sigs, d_msgs.view(-1), d_off, n, d_ok)
        a = Auditor(ctx)
        a.append_dev(d_sigs.view(-1), torch.arange(n + 1, device=dev, dtype=torch.int64) * 64, n)
        root, size = a.root()
        torch.cuda.synchronize()
        assert size == n and bool(d_ok.all())
        sig_h = d_sigs.cpu().numpy()
        assert CO.merkle_root(sig_h.reshape(-1), np.arange(n + 1, dtype=np.uint64) * 64, 8) == root
        m = 20_000
        exp = CO.ed25519_sign_batch(kseeds[np.arange(m) % K].copy(), d_msgs[:m].cpu().numpy().reshape(-1), np.arange(m + 1, dtype=np.uint64) * 512, 8)
        assert (sig_h[:m] == exp).all()
        a.close()
        del d_msgs, d_sigs, d_pks, d_ok
