"""Group-commit persistence beside the GPU batch.

Once the cryptography of a batch costs microseconds, what is left of issuing a credential in the reference is its database write:
`VCService.GenerateExecutionVC` ends in `vcStorage.StoreExecutionVC` (control-plane/internal/services/vc_service.go:229-233 ->
internal/services/vc_storage.go:32-59 -> internal/storage/local.go:6336-6361), ONE autocommitted `INSERT ... ON CONFLICT(vc_id) DO
UPDATE` per credential — one journal sync per credential.  The batch seam that feeds the GPU also feeds the database: N credentials,
one transaction, one sync.

`ExecutionVCStore` is the host-side mirror: the same table (migrations/004_create_execution_vcs.sql, + 006 storage_uri), the same
UPSERT text, the same argument order as `StorageProvider.StoreExecutionVC`, in SQLite as the reference's "local" storage mode uses.
`store_execution_vc` is the reference's behaviour (one commit per credential); `store_execution_vcs` is the group commit.  The Go
call-site sketch is in INTEGRATION.md §5 (`StoreExecutionVCs(ctx, []*types.ExecutionVC)` inside one `BeginTx` / `Commit`).

Nothing here touches the GPU; it exists so that `services.VCService(..., store=...)` can hand a whole issued batch over in one call
and so that the difference is measured (tests/test_host_logic.py::test_group_commit_persistence).
"""
import sqlite3
import time

SCHEMA = """
CREATE TABLE IF NOT EXISTS execution_vcs (
    vc_id TEXT PRIMARY KEY,
    execution_id TEXT NOT NULL,
    workflow_id TEXT NOT NULL,
    session_id TEXT NOT NULL,
    issuer_did TEXT NOT NULL,
    target_did TEXT,
    caller_did TEXT NOT NULL,
    vc_document TEXT NOT NULL,
    signature TEXT NOT NULL,
    storage_uri TEXT DEFAULT '',
    document_size_bytes INTEGER DEFAULT 0,
    input_hash TEXT NOT NULL,
    output_hash TEXT NOT NULL,
    status TEXT NOT NULL DEFAULT 'pending',
    parent_vc_id TEXT,
    child_vc_ids TEXT DEFAULT '[]',
    created_at TIMESTAMP DEFAULT CURRENT_TIMESTAMP,
    updated_at TIMESTAMP DEFAULT CURRENT_TIMESTAMP,
    FOREIGN KEY (parent_vc_id) REFERENCES execution_vcs(vc_id) ON DELETE SET NULL
);
CREATE INDEX IF NOT EXISTS idx_execution_vcs_execution_id ON execution_vcs(execution_id);
CREATE INDEX IF NOT EXISTS idx_execution_vcs_workflow_id ON execution_vcs(workflow_id);
CREATE INDEX IF NOT EXISTS idx_execution_vcs_session_id ON execution_vcs(session_id);
CREATE INDEX IF NOT EXISTS idx_execution_vcs_issuer_did ON execution_vcs(issuer_did);
CREATE UNIQUE INDEX IF NOT EXISTS idx_execution_vcs_execution_unique ON execution_vcs(execution_id, issuer_did, target_did);
"""
# internal/storage/local.go:6342-6355, verbatim in meaning: a re-issued credential replaces document, signature and status
UPSERT = """
INSERT INTO execution_vcs (
    vc_id, execution_id, workflow_id, session_id, issuer_did, target_did,
    caller_did, vc_document, signature, storage_uri, document_size_bytes,
    input_hash, output_hash, status, created_at
) VALUES (?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?)
ON CONFLICT(vc_id) DO UPDATE SET
    status = excluded.status,
    vc_document = excluded.vc_document,
    signature = excluded.signature,
    storage_uri = excluded.storage_uri,
    document_size_bytes = excluded.document_size_bytes;
"""
COLUMNS = ("vc_id", "execution_id", "workflow_id", "session_id", "issuer_did", "target_did", "caller_did", "vc_document", "signature",
           "storage_uri", "document_size_bytes", "input_hash", "output_hash", "status", "created_at")


def _row(vc, now):
    """types.ExecutionVC -> the 15 bind values of StoreExecutionVC (vc_storage.go:37-58: document_size_bytes falls back to len(doc))."""
    doc = vc["vc_document"]
    doc = doc.decode("utf-8") if isinstance(doc, (bytes, bytearray)) else doc
    size = vc.get("document_size_bytes") or len(doc.encode("utf-8"))
    return (vc["vc_id"], vc["execution_id"], vc["workflow_id"], vc["session_id"], vc["issuer_did"], vc.get("target_did", ""), vc["caller_did"], doc,
            vc["signature"], vc.get("storage_uri", ""), size, vc["input_hash"], vc["output_hash"], vc["status"], vc.get("created_at") or now)


class ExecutionVCStore:
    def __init__(self, path=":memory:", synchronous="FULL"):
        self.db = sqlite3.connect(path, isolation_level=None)            # autocommit unless a transaction is opened explicitly
        self.db.execute("PRAGMA journal_mode=WAL")
        self.db.execute("PRAGMA synchronous=%s" % synchronous)
        self.db.executescript(SCHEMA)
        self.commits = 0

    def close(self):
        self.db.close()

    def store_execution_vc(self, vc):
        """The reference's path: one statement, one commit (one journal sync) per credential."""
        self.db.execute(UPSERT, _row(vc, time.strftime("%Y-%m-%d %H:%M:%S")))
        self.commits += 1

    def store_execution_vcs(self, vcs):
        """Group commit: the batch that was signed together is persisted together — all rows or none."""
        if not vcs:
            return
        now = time.strftime("%Y-%m-%d %H:%M:%S")
        self.db.execute("BEGIN IMMEDIATE")
        try:
            self.db.executemany(UPSERT, [_row(v, now) for v in vcs])
            self.db.execute("COMMIT")
        except Exception:
            self.db.execute("ROLLBACK")
            raise
        self.commits += 1

    def get_execution_vc(self, vc_id):
        cur = self.db.execute("SELECT %s FROM execution_vcs WHERE vc_id = ?" % ", ".join(COLUMNS), (vc_id,))
        r = cur.fetchone()
        return dict(zip(COLUMNS, r)) if r else None

    def workflow_vcs(self, workflow_id):
        cur = self.db.execute("SELECT %s FROM execution_vcs WHERE workflow_id = ? ORDER BY created_at, vc_id" % ", ".join(COLUMNS), (workflow_id,))
        return [dict(zip(COLUMNS, r)) for r in cur.fetchall()]

    def count(self):
        return self.db.execute("SELECT COUNT(*) FROM execution_vcs").fetchone()[0]
