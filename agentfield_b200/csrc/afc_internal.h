// Internal (hidden-visibility) accessors shared between the translation units of libafcrypto.so.
#pragma once
#include "../../include/afcrypto.h"

#if defined(__GNUC__)
#pragma GCC visibility push(hidden)
#endif
int afc_internal_device(afc_ctx* ctx);
const void* afc_internal_comb(afc_ctx* ctx);
void afc_internal_add_launches(afc_ctx* ctx, unsigned long long n);
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
