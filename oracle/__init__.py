"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's identity-and-audit hot path (see the module headers for the
reference file:line each function follows).  Only tests/, bench.py's cpu_baseline / --impl reference
leg and __graft_entry__.smoke() may import this package; agentfield_b200/ (the product) never does.

Parity status: the reference itself pins nothing on this path and its Go
toolchain is absent, so the oracle is pinned against RFC 8032 §7.1, RFC 4231, FIPS 180-4 and RFC 6962 /
Certificate-Transparency vectors (tests/golden/*.json) and cross-checked against OpenSSL and libsodium.
"""
