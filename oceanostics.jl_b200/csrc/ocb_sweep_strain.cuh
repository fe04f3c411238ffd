// Interface between the plan object (ocb_sweep.cu) and the velocity-gradient kernel (ocb_sweep_strain.cu): the
// register-pipelined path for the closure-dependent dissipation rate and the strain / rotation moduli in either
// precision, on regular or stretched z, with constant and/or eddy-viscosity-field closures (BASELINE config 4).
#pragma once
#include "ocb_diag.cuh"

struct ocb_strain_plan;

bool ocb_strain_handles_diag(int diag);
// *out stays nullptr when the requests do not fit the envelope (the generic kernel then evaluates them)
int ocb_strain_try_create(const ocb_grid* grid, const ocb_sweep_inputs* in, const ocb_request* reqs, int n,
                          const SweepParams<double>& P, ocb_strain_plan** out);
int ocb_strain_try_create(const ocb_grid* grid, const ocb_sweep_inputs* in, const ocb_request* reqs, int n,
                          const SweepParams<float>& P, ocb_strain_plan** out);
int ocb_strain_nblocks(const ocb_strain_plan* sp);
int ocb_strain_execute(ocb_strain_plan* sp, double* partial, cudaStream_t st);
void ocb_strain_destroy(ocb_strain_plan* sp);
