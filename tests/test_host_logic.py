"""Host-side logic, no GPU: packing, Go-mirrored argument checks, shard partitions, and the N>1 Merkle
exchange step over torch.distributed `gloo` (world_size 2) with the oracle standing in for the device hashing."""
import os
import socket

import numpy as np
import pytest

from conftest import golden
from oracle import merkle as M


def test_pack_layout():
    from agentfield_b200 import pack, pack32
    buf, off = pack([b"ab", b"", b"cde"])
    assert off.dtype == np.uint64 and off.tolist() == [0, 2, 2, 5] and bytes(buf[:5]) == b"abcde"
    buf, off = pack([])
    assert off.tolist() == [0]
    kb, ko = pack32([b"k1", b"key2"])
    assert ko.dtype == np.uint32 and ko.tolist() == [0, 2, 6]


def test_go_mirrored_argument_errors():
    """ed25519.Verify panics on a bad public-key length, NewKeyFromSeed on a bad seed length — before any packing
    (so these raise even without a GPU)."""
    from agentfield_b200 import Signer, Verifier

    class _NoCtx:
        pass
    v, s = Verifier(ctx=_NoCtx()), Signer(ctx=_NoCtx())
    with pytest.raises(ValueError, match="bad public key length"):
        v.verify_batch([b"\x00" * 31], [b"m"], [b"\x00" * 64])
    with pytest.raises(ValueError, match="bad seed length"):
        s.sign_batch([b"\x00" * 33], [b"m"])
    with pytest.raises(ValueError):
        v.verify_batch([b"\x00" * 32], [b"m", b"n"], [b"\x00" * 64])
    assert v.verify_batch([], [], []) == []


def test_shard_partitions():
    from agentfield_b200 import shard
    for n in (0, 1, 7, 8, 1000, 4_000_000, 1 << 22):
        for world in (1, 2, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = shard.shard_range(n, r, world)
                cover.append((lo, hi))
            assert cover[0][0] == 0 and cover[-1][1] == n and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
            b = shard.merkle_block(n, world)
            assert b & (b - 1) == 0 and b * world >= n and (b == 1 or (b // 2) * world < n)
            mcover = [shard.merkle_shard_range(n, r, world) for r in range(world)]
            assert mcover[0][0] == 0 and mcover[-1][1] == n and all(lo % b == 0 or lo == n for lo, _ in mcover)
            assert sorted(i for r in range(world) for i in shard.round_robin_indices(min(n, 50), r, world)) == list(range(min(n, 50)))


def test_aligned_shard_roots_fold_to_the_global_root():
    """Why shards must be contiguous and 2^k-aligned: MTH over the per-shard roots == MTH over all leaves."""
    from agentfield_b200 import shard
    rng = np.random.default_rng(3)
    for n in (1, 5, 8, 100, 1000, 1025):
        hs = [rng.integers(0, 256, 32, dtype=np.uint8).tobytes() for _ in range(n)]
        full = M.root_from_leaf_hashes(hs)
        for world in (1, 2, 4, 8):
            roots = []
            for r in range(world):
                lo, hi = shard.merkle_shard_range(n, r, world)
                if hi > lo:
                    roots.append(M.root_from_leaf_hashes(hs[lo:hi]))
            assert M.root_from_leaf_hashes(roots) == full, (n, world)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    from agentfield_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0xAF04)
    leaves = [rng.integers(0, 256, 96, dtype=np.uint8).tobytes() for _ in range(n)]
    lo, hi = shard.merkle_shard_range(n, rank, world)
    local = M.root(leaves[lo:hi]) if hi > lo else None          # oracle stands in for the device hashing here
    roots = shard.allgather_roots(local)
    folded = M.root_from_leaf_hashes([r for r in roots if r is not None])
    q.put((rank, folded.hex(), M.root(leaves).hex()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [8, 1000])
def test_merkle_exchange_step_world2_gloo(n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(60)
    assert len(res) == 2 and all(f == full for _, f, full in res) and res[0][1] == res[1][1]


def _vc_requests(n, dids, rng, pad_to=512):
    """Synthetic GenerateExecutionVC inputs; executionId padded so the canonical form is exactly `pad_to` bytes (cfg1)."""
    from oracle import ref_vc, go_hash as H
    reqs = []
    for i in range(n):
        r = {"vc_id": "vc-%019d" % (1789971100759287000 + i), "caller_did": dids[i % len(dids)], "target_did": dids[(i + 1) % len(dids)],
             "agent_node_did": dids[0], "caller_type": "agent", "function_name": "fn_%d" % (i % 7), "issuance_date": "2026-09-21T06:00:00Z",
             "execution_id": "exec_%06d" % i, "workflow_id": "wf_%04d" % (i // 10), "session_id": "sess_%04d" % (i // 100),
             "input": rng.integers(0, 256, int(rng.integers(0, 64)), dtype=np.uint8).tobytes() if i % 5 else None,
             "output": rng.integers(0, 256, int(rng.integers(0, 64)), dtype=np.uint8).tobytes(), "status": "succeeded" if i % 9 else "failed",
             "error_message": None if i % 9 else "boom <&> \u2028 \"x\"", "duration_ms": int(rng.integers(0, 5000)),
             "timestamp": "2026-09-21T06:00:%02dZ" % (i % 60), "proof_created": "2026-09-21T06:00:01Z"}
        if pad_to:
            ih = H.hash_data(H.marshal_data_or_null(r["input"])); oh = H.hash_data(H.marshal_data_or_null(r["output"]))
            cur = len(ref_vc.go_marshal(ref_vc.vc_document(r, ih, oh)))
            if cur < pad_to:
                r["execution_id"] += "x" * (pad_to - cur)
        reqs.append(r)
    return reqs


def test_go_json_restatement_two_constructions_agree():
    """agentfield_b200/go_json.py (string builder) vs oracle/ref_vc.py (json module + escapes): same canonical bytes."""
    from agentfield_b200 import go_json
    from oracle import ref_vc
    rng = np.random.default_rng(0xAF01)
    dids = ["did:key:z%s" % ("A" * 46), "did:key:z%s" % ("B" * 46), "did:key:z%s" % ("C" * 46)]
    for r in _vc_requests(200, dids, rng, pad_to=0):
        doc = ref_vc.vc_document(r, "ih-%s" % r["vc_id"], "oh")
        a = ref_vc.go_marshal(doc)
        d2 = dict(doc); d2["credentialSubject"] = dict(doc["credentialSubject"])
        d2["credentialSubject"]["execution"] = dict(doc["credentialSubject"]["execution"])
        d2["credentialSubject"]["execution"].setdefault("errorMessage", "")
        assert go_json.vc_document(d2) == a
        assert go_json.vc_document(d2, {"type": "T", "created": "c", "verificationMethod": "v", "proofPurpose": "p", "proofValue": "s"}) == \
            ref_vc.go_marshal(dict(doc, proof={"type": "T", "created": "c", "verificationMethod": "v", "proofPurpose": "p", "proofValue": "s"}))
    assert go_json.string("a<b>&c\u2028\x01\b\f\"\\") == '"a\\u003cb\\u003e\\u0026c\\u2028\\u0001\\b\\f\\"\\\\"'
    assert [go_json.number(x) for x in (0.1, 1e-7, 1e21, 42.0, 123456789.125, 3)] == ["0.1", "1e-7", "1e+21", "42", "123456789.125", "3"]
    assert go_json.value({"b": [1, "x", None, True], "a": {"z": 1.5}}) == '{"a":{"z":1.5},"b":[1,"x",null,true]}'
    wp = go_json.webhook_payload({"event": "execution.completed", "execution_id": "e", "workflow_id": "w", "status": "succeeded", "target": "n.fn",
                                  "type": "reasoner", "duration_ms": 12, "result": {"k": "<v>"}, "error_message": None, "timestamp": "t"})
    assert wp == b'{"event":"execution.completed","execution_id":"e","workflow_id":"w","status":"succeeded","target":"n.fn","type":"reasoner","duration_ms":12,"result":{"k":"\\u003cv\\u003e"},"timestamp":"t"}'


def test_vc_document_template_matches_go_json():
    """The constant segments + value order of the device canonical-form template reproduce json.Marshal(VCDocument) when filled
    by the byte-level oracle (no GPU needed): checks the template itself, omitempty errorMessage, null slices, nested metadata."""
    from agentfield_b200 import canonical as CA, go_json as GJ
    from oracle import go_json as OJ
    rng = np.random.default_rng(0xAF36)
    alphabet = list("abcXYZ019 -_:/.") + ['"', "\\", "<", ">", "&", "\n", "\t", "\x00", "\x7f", "é", "中", "😀", " "]

    def word(lo=0, hi=20):
        return "".join(alphabet[int(i)] for i in rng.integers(0, len(alphabet), int(rng.integers(lo, hi))))

    p0, p1 = CA.vc_document_template_parts(False), CA.vc_document_template_parts(True)
    for i in range(300):
        doc = {"@context": [word(), word()], "type": [word()] if i % 5 else None, "id": word(), "issuer": word(), "issuanceDate": word(),
               "credentialSubject": {"executionId": word(), "workflowId": word(), "sessionId": word(),
                                     "caller": {"did": word(), "type": word(), "agentNodeDid": word()},
                                     "target": {"did": word(), "agentNodeDid": word(), "functionName": word()},
                                     "execution": {"inputHash": word(), "outputHash": word(), "timestamp": word(), "durationMs": int(rng.integers(0, 2**40)),
                                                   "status": word(), "errorMessage": word(1) if i % 3 == 0 else ""},
                                     "audit": {"inputDataHash": word(), "outputDataHash": word(), "metadata": {word(1): word(), "k": [1, word()]} if i % 4 else None}}}
        proof = {"type": word(), "created": word(), "verificationMethod": word(), "proofPurpose": word(), "proofValue": word()}
        assert OJ.fill_template(p0[0], p0[1], CA.vc_document_values(doc)) == GJ.vc_document(doc)
        assert OJ.fill_template(p1[0], p1[1], CA.vc_document_values(doc, proof)) == GJ.vc_document(doc, proof)
    # the webhook payload template (HMAC input, W1)
    wp = CA.webhook_payload_template_parts()
    for i in range(100):
        pl = {"event": word(), "execution_id": word(), "workflow_id": word(), "status": word(), "target": word(), "type": word(),
              "duration_ms": int(rng.integers(0, 10**6)) if i % 2 else None, "result": {"a": word(), "b": [1, 2.5, None]} if i % 3 else None,
              "error_message": word() if i % 5 == 0 else None, "timestamp": word()}
        assert OJ.fill_template(wp[0], wp[1], CA.webhook_payload_values(pl)) == GJ.webhook_payload(pl)


def test_go_cases_host_mirror_reproduces_every_expectation():
    """tests/golden/go_cases.json (oracle-made, re-pinnable by baseline/go/gen_golden_test.go): the host mirror must write the
    same bytes — float64 formatting in every range, string escaping incl. invalid UTF-8, omitempty, nil slices, the metadata map
    after json.Unmarshal, an error message cut inside a rune, WorkflowVCDocument, the webhook payload."""
    import struct
    from agentfield_b200 import go_json as GJ
    import go_cases_util as U
    g = golden("go_cases.json")
    for c in g["floats"]:
        x = struct.unpack(">d", bytes.fromhex(c["bits"]))[0]
        assert GJ.number(x) == c["expect"], c
    for c in g["strings"]:
        assert GJ.string_bytes(bytes.fromhex(c["utf8"])) == c["expect"], c
    for c in g["execution_vcs"]:
        doc = U.execution_doc(c)
        assert GJ.vc_document(doc) == c["expect_canonical"].encode("utf-8"), c["id"]
        assert GJ.vc_document(doc, U.proof(c["issuer"], bytes.fromhex(c["expect_sig"]), c["proof_created"])) == c["expect_stored"].encode("utf-8")
    for c in g["workflow_vcs"]:
        doc = U.workflow_doc(c)
        assert GJ.workflow_vc_document(doc) == c["expect_canonical"].encode("utf-8"), c["workflow_id"]
        assert GJ.workflow_vc_document(doc, U.proof(c["issuer_did"], bytes.fromhex(c["expect_sig"]), c["proof_created"])) == c["expect_stored"].encode("utf-8")
    for c in g["webhooks"]:
        assert GJ.webhook_payload(U.webhook(c)) == c["expect_body"].encode("utf-8"), c["execution_id"]


def test_go_cases_device_templates_reproduce_every_expectation():
    """The device canonical-form templates (VCDocument, WorkflowVCDocument, webhook payload) filled by the byte-level oracle of
    the device routine give the same bytes (no GPU needed; the kernels are checked against the same oracle in the GPU tier)."""
    from agentfield_b200 import canonical as CA
    from oracle import go_json as OJ
    import go_cases_util as U
    g = golden("go_cases.json")
    p0, p1 = CA.vc_document_template_parts(False), CA.vc_document_template_parts(True)
    for c in g["execution_vcs"]:
        doc = U.execution_doc(c)
        assert OJ.fill_template(p0[0], p0[1], CA.vc_document_values(doc)) == c["expect_canonical"].encode("utf-8")
        pr = U.proof(c["issuer"], bytes.fromhex(c["expect_sig"]), c["proof_created"])
        assert OJ.fill_template(p1[0], p1[1], CA.vc_document_values(doc, pr)) == c["expect_stored"].encode("utf-8")
    w0, w1 = CA.workflow_vc_document_template_parts(False), CA.workflow_vc_document_template_parts(True)
    for c in g["workflow_vcs"]:
        doc = U.workflow_doc(c)
        assert OJ.fill_template(w0[0], w0[1], CA.workflow_vc_document_values(doc)) == c["expect_canonical"].encode("utf-8")
        pr = U.proof(c["issuer_did"], bytes.fromhex(c["expect_sig"]), c["proof_created"])
        assert OJ.fill_template(w1[0], w1[1], CA.workflow_vc_document_values(doc, pr)) == c["expect_stored"].encode("utf-8")
    wp = CA.webhook_payload_template_parts()
    for c in g["webhooks"]:
        assert OJ.fill_template(wp[0], wp[1], CA.webhook_payload_values(U.webhook(c))) == c["expect_body"].encode("utf-8")


def test_workflow_status_roll_up_and_normalisation():
    """determineWorkflowStatus / NormalizeExecutionStatus / countCompletedSteps (vc_service.go:721-787, pkg/types/status.go:28-75)."""
    from agentfield_b200 import services as S
    from oracle import ref_vc as RV
    assert [S.normalize_execution_status(x) for x in (" Completed ", "OK", "errored", "canceled", "timed_out", "waiting", "in_progress", "", "bogus", "SUCCEEDED")] == \
        ["succeeded", "succeeded", "failed", "cancelled", "timeout", "queued", "running", "unknown", "unknown", "succeeded"]
    assert S.is_terminal_execution_status("done") and not S.is_terminal_execution_status("processing")
    rng = np.random.default_rng(5)
    pool = ["succeeded", "completed", "failed", "error", "timeout", "cancelled", "running", "queued", "pending", "weird", ""]
    for _ in range(300):
        evs = [{"status": pool[int(i)]} for i in rng.integers(0, len(pool), int(rng.integers(0, 6)))]
        assert S.determine_workflow_status(evs) == RV.determine_workflow_status([S.normalize_execution_status(e["status"]) for e in evs])
        assert S.count_completed_steps(evs) == sum(e["status"] in ("succeeded", "completed") for e in evs)
    assert S.determine_workflow_status([]) == "pending"


def test_group_commit_persistence(tmp_path):
    """N2: the reference commits one credential at a time (StoreExecutionVC, internal/storage/local.go:6336-6361); the batch seam
    commits a batch at a time.  Same table, same UPSERT, same rows either way; a re-issued vc_id replaces document / signature /
    status and nothing else; a batch is atomic; and on a real file with synchronous=FULL the group commit is many times faster."""
    import time
    from agentfield_b200.vc_store import ExecutionVCStore
    rng = np.random.default_rng(9)

    def vc(i, wf="wf-1", doc=None):
        return {"vc_id": "vc-%06d" % i, "execution_id": "exec-%06d" % i, "workflow_id": wf, "session_id": "s", "issuer_did": "did:key:zA",
                "target_did": "did:key:zB", "caller_did": "did:key:zA", "vc_document": doc or ('{"n":%d,"pad":"%s"}' % (i, "x" * 1200)).encode(),
                "signature": "sig-%d" % i, "input_hash": "ih", "output_hash": "oh", "status": "succeeded"}
    a, b = ExecutionVCStore(str(tmp_path / "one.db")), ExecutionVCStore(str(tmp_path / "batch.db"))
    n = 400
    vcs = [vc(i) for i in range(n)]
    t0 = time.perf_counter()
    for v in vcs:
        a.store_execution_vc(v)
    t_one = time.perf_counter() - t0
    t0 = time.perf_counter()
    b.store_execution_vcs(vcs)
    t_batch = time.perf_counter() - t0
    assert a.count() == b.count() == n and a.commits == n and b.commits == 1
    ra, rb = a.get_execution_vc("vc-000123"), b.get_execution_vc("vc-000123")
    assert {k: v for k, v in ra.items() if k != "created_at"} == {k: v for k, v in rb.items() if k != "created_at"}
    assert rb["document_size_bytes"] == len(vcs[123]["vc_document"]) and [r["vc_id"] for r in b.workflow_vcs("wf-1")][:3] == ["vc-000000", "vc-000001", "vc-000002"]
    assert t_batch < t_one / 5, (t_one, t_batch)               # one journal sync instead of 400
    # UPSERT: document, signature, status, size are replaced; ids and hashes of the first insert stay
    again = dict(vc(123), vc_document=b'{"changed":true}', signature="sig-new", status="failed", input_hash="OTHER")
    b.store_execution_vcs([again])
    r2 = b.get_execution_vc("vc-000123")
    assert (r2["vc_document"], r2["signature"], r2["status"], r2["document_size_bytes"], r2["input_hash"]) == ('{"changed":true}', "sig-new", "failed", 16, "ih")
    # atomic: a batch with a row that violates the (execution_id, issuer, target) uniqueness leaves nothing behind
    clash = dict(vc(900), execution_id="exec-000005")           # another vc_id for an execution that already has one
    with pytest.raises(Exception):
        b.store_execution_vcs([vc(901), clash, vc(902)])
    assert b.count() == n and b.get_execution_vc("vc-000901") is None
    a.close(); b.close()
