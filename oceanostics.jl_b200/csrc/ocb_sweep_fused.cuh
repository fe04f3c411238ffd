// Interface between the plan object (ocb_sweep.cu) and the register-pipelined budget kernel
// (ocb_sweep_fused.cu): the fast path a plan takes when every request is one of the cell-centred budget
// terms over full Float64/Float32 array operands with a constant-coefficient closure.
#pragma once
#include "ocb_diag.cuh"

struct ocb_fused_plan;

// *out stays nullptr when the plan does not fit the envelope (the generic kernel then evaluates it)
int ocb_fused_try_create(const ocb_grid* grid, const ocb_sweep_inputs* in, const ocb_request* reqs, int n,
                         const SweepParams<double>& P, ocb_fused_plan** out);
int ocb_fused_try_create(const ocb_grid* grid, const ocb_sweep_inputs* in, const ocb_request* reqs, int n,
                         const SweepParams<float>& P, ocb_fused_plan** out);
bool ocb_fused_handles_diag(int diag);     // a diagnostic the budget kernel can evaluate (plans may be split on this)
int ocb_fused_nblocks(const ocb_fused_plan* fp);
int ocb_fused_execute(ocb_fused_plan* fp, double* partial, cudaStream_t st);
// the class-sum kernel in ring mode (whole periodic y extent): may run with its first / last tile rows gated on the
// halo flag words of a peer-memory ring (flags[0] low neighbour, flags[1] high neighbour, flags[2] set on timeout)
bool ocb_fused_ring_capable(const ocb_fused_plan* fp);
int ocb_fused_execute_gated(ocb_fused_plan* fp, double* partial, long long* flags, long long step, long long timeout_cycles, cudaStream_t st);
void ocb_fused_destroy(ocb_fused_plan* fp);
