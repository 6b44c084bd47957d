"""agentfield_b200 — B200-native (sm_100a) drop-in for AgentField's cryptographic identity-and-audit hot
path: batched Ed25519 sign/verify for Verifiable Credentials, HMAC-SHA256 webhook signing, SHA-256
payload hashing and an RFC 6962 Merkle audit log, behind the C ABI in include/afcrypto.h.

Only what the path needs lives here: csrc/ (CUDA kernels + C ABI), _abi.py (ctypes binding), crypto.py /
audit.py / identity.py / services.py (host-side mirrors of the reference's Go seams).  All arithmetic
runs in libafcrypto.so on the GPU; there is no CPU fallback in this package.
"""
from . import _abi
from ._abi import AfcError, LIB_PATH
from .crypto import Context, Hasher, MAC, PayloadHasher, Signer, Verifier, default_context, pack, pack32
from .audit import Auditor, MerkleTree, fold_roots, verify_inclusion_batch, verify_consistency_batch
from .canonical import (JsonTemplate, vc_document_template, vc_document_values, workflow_vc_document_template,
                        workflow_vc_document_values)
from .identity import ExpandedKeys, KeySet, did_key
from .dispatcher import Ingest
from .bundle import export_bundle, verify_bundle

__all__ = ["AfcError", "LIB_PATH", "Context", "Hasher", "MAC", "PayloadHasher", "Signer", "Verifier", "Auditor", "MerkleTree", "fold_roots", "verify_inclusion_batch", "verify_consistency_batch", "JsonTemplate", "vc_document_template", "vc_document_values", "workflow_vc_document_template", "workflow_vc_document_values",
           "default_context", "pack", "pack32", "ExpandedKeys", "KeySet", "did_key", "Ingest", "export_bundle", "verify_bundle"]
__version__ = "0.1.0"
