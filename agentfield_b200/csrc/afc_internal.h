// Internal (hidden-visibility) accessors shared between the translation units of libafcrypto.so.
#pragma once
#include "../../include/afcrypto.h"

#if defined(__GNUC__)
#pragma GCC visibility push(hidden)
#endif
int afc_internal_device(afc_ctx* ctx);
const void* afc_internal_comb(afc_ctx* ctx);
// the constant-time signing table, or nullptr when the context is configured for the fast (variable-time) path
const void* afc_internal_sign_table(afc_ctx* ctx);
void afc_internal_add_launches(afc_ctx* ctx, unsigned long long n);
// cudaHostAlloc from a thread temporarily bound to the CPUs next to ctx's GPU (pages land on that NUMA node)
void* afc_internal_pinned_alloc(afc_ctx* ctx, size_t bytes);
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
