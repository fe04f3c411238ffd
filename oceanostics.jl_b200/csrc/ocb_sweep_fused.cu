// placeholder: the budget kernel is developed next; until then no plan takes the fused path
#include "ocb_sweep_fused.cuh"
struct ocb_fused_plan { int nblocks; };
int ocb_fused_try_create(const ocb_grid*, const ocb_sweep_inputs*, const ocb_request*, int, const SweepParams<double>&, ocb_fused_plan** out) { *out = nullptr; return OCB_OK; }
int ocb_fused_try_create(const ocb_grid*, const ocb_sweep_inputs*, const ocb_request*, int, const SweepParams<float>&, ocb_fused_plan** out) { *out = nullptr; return OCB_OK; }
int ocb_fused_nblocks(const ocb_fused_plan* fp) { return fp->nblocks; }
int ocb_fused_execute(ocb_fused_plan*, double*, cudaStream_t) { return OCB_OK; }
void ocb_fused_destroy(ocb_fused_plan* fp) { delete fp; }
