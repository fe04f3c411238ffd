#include "ocb_common.cuh"
extern "C" int ocb_bpe_reference_height(const ocb_grid*, const ocb_field*, int32_t, const ocb_field*, const ocb_field*, int64_t*, void*, void*) { OCB_FAIL(OCB_ERR_UNSUPPORTED, "bpe: under construction"); }
extern "C" int ocb_sort_pairs(int32_t, const void*, void*, uint32_t*, int64_t, void*) { OCB_FAIL(OCB_ERR_UNSUPPORTED, "sort: under construction"); }
