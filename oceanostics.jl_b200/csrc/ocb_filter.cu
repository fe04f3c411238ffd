#include "ocb_common.cuh"
extern "C" int ocb_filter_execute(int32_t, const ocb_field*, const int32_t*, const ocb_filter_pass*, int32_t, const ocb_field*, const int32_t*, const int32_t*, void*) { OCB_FAIL(OCB_ERR_UNSUPPORTED, "filters: under construction"); }
