/*
 * oceanostics_b200.h -- C ABI of liboceanostics_b200.so, the B200 (sm_100a) replacement for the
 * evaluation hot path of tomchor/Oceanostics.jl.
 *
 * The reference has no FFI of its own: every diagnostic is an Oceananigans
 * `KernelFunctionOperation` evaluated by `Field(op)` / `compute!` / `Integral` / `Average`
 * (reference: src/Oceanostics.jl:5 `CustomKFO`; the precedent for intercepting `compute!` is
 * src/SpatialFilters/SpatialFilters.jl:309,347-378, src/SpatialFilters/box_filter.jl:229-230,
 * src/SpatialFilters/gaussian_filter.jl:533-534, src/BackgroundPotentialEnergyEquation.jl:254-262,
 * 610-619).  The entry points below are what a Julia `ccall` layer binds from those `compute!`
 * overrides (INTEGRATION.md shows the stubs); the Python host package binds the same symbols
 * through ctypes.
 *
 * Conventions
 *   - plain C types only; every pointer marked "device" is a CUDA device pointer owned by the CALLER
 *     (CuArrays behind Oceananigans Fields).  The library never frees or reallocates caller memory and
 *     keeps no caller pointer beyond a plan's lifetime.
 *   - arrays are Oceananigans parent arrays: column-major, i fastest, interior index (1,1,1) at
 *     element offset halo[0]*stride[0] + halo[1]*stride[1] + halo[2]*stride[2].
 *   - every call is stream-ordered on the `stream` it is handed (a cudaStream_t passed as void*),
 *     never synchronises the device, returns 0 on success and a non-zero OCB_ERR_* otherwise;
 *     `ocb_last_error()` returns the message (thread-local).  No exceptions or abort cross the ABI.
 *   - there is NO CPU fallback: every compute entry point fails with OCB_ERR_CUDA when no CUDA
 *     device is usable.
 */
#ifndef OCEANOSTICS_B200_H
#define OCEANOSTICS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCB_ABI_VERSION 2

/* ---- status codes -------------------------------------------------------------------------- */
enum {
  OCB_OK = 0,
  OCB_ERR_INVALID = 1,      /* bad argument (the host layer raises ArgumentError, cf. src/Oceanostics.jl:114-120) */
  OCB_ERR_UNSUPPORTED = 2,  /* valid request this library does not evaluate (host materialises via Field(arg)) */
  OCB_ERR_CUDA = 3,         /* CUDA runtime error / no device */
  OCB_ERR_ALLOC = 4
};

/* ---- basic enums --------------------------------------------------------------------------- */
enum { OCB_F32 = 0, OCB_F64 = 1 };                       /* grid FT (eltype(grid)) */
enum { OCB_PERIODIC = 0, OCB_BOUNDED = 1 };              /* Flat axes are passed as Periodic, N = 1 */
enum { OCB_CENTER = 0, OCB_FACE = 1, OCB_NOTHING = 2 };  /* location; NOTHING = reduced axis */
enum { OCB_FIELD_ARRAY = 0, OCB_FIELD_ZERO = 1, OCB_FIELD_CONSTANT = 2 };

/* ---- grid: Oceananigans RectilinearGrid ----------------------------------------------------- */
typedef struct {
  int32_t dtype;          /* OCB_F32 / OCB_F64 */
  int32_t N[3];           /* Nx, Ny, Nz */
  int32_t H[3];           /* halo */
  int32_t topo[3];        /* OCB_PERIODIC / OCB_BOUNDED */
  double  L[3];           /* Lx, Ly, Lz */
  double  d[3];           /* uniform spacing per axis (ignored on a stretched axis) */
  double  z_bottom;       /* znode(1, Face) */
  /* stretched z: device arrays in the grid FT, entry m <-> index k = m + 1 - Hz,
     lengths Nz + 2Hz (+1 for zf/dzf). NULL on a regular z axis (only z may be stretched for the
     stencil sweep; filters take node/spacing arrays per pass). */
  const void* dzc;        /* Δzᵃᵃᶜ */
  const void* dzf;        /* Δzᵃᵃᶠ */
  const void* zc;         /* zᵃᵃᶜ */
  const void* zf;         /* zᵃᵃᶠ */
} ocb_grid;

/* ---- field operand -------------------------------------------------------------------------- */
typedef struct {
  void*   data;           /* device pointer to the PARENT array (ARRAY kind), else NULL */
  int32_t kind;           /* OCB_FIELD_ARRAY / ZERO (ZeroField) / CONSTANT (a Number) */
  int32_t loc[3];         /* OCB_CENTER / OCB_FACE / OCB_NOTHING (reduced: Average(u, dims=...)) */
  int32_t halo[3];        /* halo of the parent along each axis (0 on a reduced axis) */
  int32_t size[3];        /* parent extents (interior + 2*halo; 1 on a reduced axis) */
  int64_t stride[3];      /* element strides (0 on a reduced axis => broadcast) */
  double  value;          /* the Number for OCB_FIELD_CONSTANT */
} ocb_field;

/* ---- closures: ScalarDiffusivity / SmagorinskyLilly / AMD and tuples of them ----------------- */
/* Isotropic (ThreeDimensionalFormulation) closures only, as validate_dissipative_closure demands
 * (src/Oceanostics.jl:118-120).  A closure tuple sums its fluxes, i.e. its viscosities:
 *   nu@loc    = nu_const    + sum_f  I_loc(nu_fields[f])
 *   kappa@loc = kappa_const + sum_f  kappa_scale[f] * I_loc(kappa_fields[f])   (kappa = nu_e/Pr: scale = 1/Pr)
 * with I_loc the Oceananigans interpolation of a ccc field to ccc/ffc/fcf/cff (nu) or fcc/cfc/ccf (kappa). */
#define OCB_MAX_CLOSURE_FIELDS 2
typedef struct {
  double    nu_const;
  int32_t   n_nu_fields;
  ocb_field nu_fields[OCB_MAX_CLOSURE_FIELDS];
  double    kappa_const;
  int32_t   n_kappa_fields;
  ocb_field kappa_fields[OCB_MAX_CLOSURE_FIELDS];
  double    kappa_scale[OCB_MAX_CLOSURE_FIELDS];
} ocb_closure;

/* ---- diagnostics of the fused stencil sweep (K1) -------------------------------------------- */
/* Each id names the reference kernel function it replaces (file:line under /root/reference/src). */
enum {
  OCB_DIAG_KE = 0,        /* kinetic_energy_ccc                      KineticEnergyEquation.jl:37-39      ccc */
  OCB_DIAG_ADV,           /* uᵢ∂ⱼuⱼuᵢᶜᶜᶜ (Centered 2nd order)         KineticEnergyEquation.jl:147-152    ccc */
  OCB_DIAG_PR,            /* uᵢ∂ᵢpᶜᶜᶜ                                KineticEnergyEquation.jl:304-309    ccc */
  OCB_DIAG_WB,            /* uᵢbᵢᶜᶜᶜ                                 KineticEnergyEquation.jl:362-367    ccc */
  OCB_DIAG_EPS,           /* viscous_dissipation_rate_ccc            KineticEnergyEquation.jl:427-453    ccc */
  OCB_DIAG_EPS_ISO,       /* isotropic_viscous_dissipation_rate_ccc  KineticEnergyEquation.jl:497-510    ccc */
  OCB_DIAG_TKE,           /* turbulent_kinetic_energy_ccc            TurbulentKineticEnergyEquation.jl:28-30 */
  OCB_DIAG_SPX,           /* shear_production_rate_x_ccc             TurbulentKineticEnergyEquation.jl:101-117 */
  OCB_DIAG_SPY,           /* shear_production_rate_y_ccc             :163-179 */
  OCB_DIAG_SPZ,           /* shear_production_rate_z_ccc             :224-240 */
  OCB_DIAG_SP,            /* shear_production_rate_ccc               :285-289 */
  OCB_DIAG_CHI,           /* tracer_variance_dissipation_rate_ccc    TracerVarianceEquation.jl:146-161   ccc */
  OCB_DIAG_CDIFF,         /* c∇_dot_qᶜ                               TracerVarianceEquation.jl:93-98     ccc */
  OCB_DIAG_SMOD,          /* strain_rate_tensor_modulus_ccc          FlowDiagnostics.jl:431-441          ccc */
  OCB_DIAG_OMOD,          /* vorticity_tensor_modulus_ccc            FlowDiagnostics.jl:577-587          ccc */
  OCB_DIAG_Q,             /* Q_velocity_gradient_tensor_invariant_ccc FlowDiagnostics.jl:708-712         ccc */
  OCB_DIAG_S11, OCB_DIAG_S22, OCB_DIAG_S33,      /* StrainRateTensorKernel{I,J}  FlowDiagnostics.jl:484-491  ccc */
  OCB_DIAG_S12, OCB_DIAG_S13, OCB_DIAG_S23,      /*                                                ffc fcf cff */
  OCB_DIAG_O12, OCB_DIAG_O13, OCB_DIAG_O23,      /* VorticityTensorKernel{I,J}   FlowDiagnostics.jl:630-634  ffc fcf cff */
  OCB_DIAG_T11, OCB_DIAG_T22, OCB_DIAG_T33,      /* StressTensorKernel{I,I,false} :728-730          fcc cfc ccf */
  OCB_DIAG_T11C, OCB_DIAG_T22C, OCB_DIAG_T33C,   /* StressTensorKernel{I,I,true}  :725-727          ccc */
  OCB_DIAG_T12, OCB_DIAG_T13, OCB_DIAG_T23,      /* StressTensorKernel{I,J,false} :731-733          ffc fcf cff */
  OCB_DIAG_RI,            /* richardson_number_ccf                   FlowDiagnostics.jl:55-68            ccf */
  OCB_DIAG_RO,            /* rossby_number_fff                       FlowDiagnostics.jl:124-138          fff */
  OCB_DIAG_EPV,           /* ertel_potential_vorticity_fff           FlowDiagnostics.jl:216-233          fff */
  OCB_DIAG_TWPV,          /* potential_vorticity_in_thermal_wind_fff FlowDiagnostics.jl:200-214          fff */
  OCB_DIAG_DEPV,          /* directional_ertel_potential_vorticity_fff FlowDiagnostics.jl:358-380        fff */
  OCB_DIAG_VF11, OCB_DIAG_VF22, OCB_DIAG_VF33,   /* viscous_flux_ux / vy / wz (Oceananigans; filtered by FilteredKineticEnergyEquation.jl:377-387) ccc */
  OCB_DIAG_VF12, OCB_DIAG_VF13, OCB_DIAG_VF23,   /* viscous_flux_uy(=vx) / uz(=wx) / vz(=wy)        ffc fcf cff */
  OCB_DIAG_FEPS,          /* filtered_dissipation_rate_ccc           FilteredKineticEnergyEquation.jl:275-299  (aux[0..5] = filtered fluxes) */
  OCB_DIAG_SFEPS,         /* subfilter_ke_dissipation_rate: aux[6] - FEPS   SubFilterKineticEnergyEquation.jl:92,143-148 */
  OCB_DIAG_SFKE,          /* subfilter_kinetic_energy_ccc: aux[6] - KE(u,v,w) SubFilterKineticEnergyEquation.jl:29 */
  OCB_DIAG_PI,            /* cross_scale_ke_flux_ccc: -sum w_ij center(aux_ij - uᵢuⱼ) center(S_ij)  FilteredKineticEnergyEquation.jl:176-188 */
  OCB_DIAG_BPE,           /* minus_bz✶_ccc: -b * aux[6]              BackgroundPotentialEnergyEquation.jl:852 */
  OCB_DIAG_APE,           /* local_ape_ccc: aux[5] - b (z - aux[6])   AvailablePotentialEnergyEquation.jl:42  (next-row, §8f) */
  OCB_DIAG_UPSILON,       /* upsilon_ccc: aux[6] - z  (aux[6] = z✶)   AvailablePotentialEnergyEquation.jl:126 (next-row, §8f) */
  OCB_DIAG_APE_EPS,       /* ape_dissipation_rate_ccc: (1/V) sum center(-A δΥ q), Υ = aux[4], tracer b   AvailablePotentialEnergyEquation.jl:202-215 */
  OCB_DIAG_KE_STRESS,     /* uᵢ∂ⱼ_τᵢⱼᶜᶜᶜ: velocity x divergence of the viscous fluxes  KineticEnergyEquation.jl:191-202 (next-row, §8f) */
  OCB_DIAG_KE_FORCING,    /* uᵢFᵤᵢᶜᶜᶜ: velocity x forcing, forcings materialised at the velocity points in aux[0..2]  KineticEnergyEquation.jl:251-260 (next-row, §8f) */
  OCB_DIAG_C_TENDENCY,    /* c∂ₜcᶜᶜᶜ = 2c (-div_Uc[Centered 2nd order] - ∇·q + F), tracer in slot b, F in aux[0]  TracerVarianceEquation.jl:18-39 (next-row, §8f) */
  OCB_DIAG_KE_TENDENCY,   /* uᵢGᵢᶜᶜᶜ: velocity x NonhydrostaticModel momentum tendency (Centered 2nd-order advection, Coriolis f[3], buoyancy b x ghat,
                             hydrostatic pressure anomaly in slot p, viscous flux divergence, forcings in aux[0..2]); no background fields,
                             immersed boundaries or Stokes drift  KineticEnergyEquation.jl:74-95 (next-row, §8f) */
  OCB_DIAG_NU_SMAG,       /* the step before the path (§8f row 3): Smagorinsky-Lilly eddy viscosity νₑ = ς (C Δᶠ)² √(2 ΣᵢⱼΣᵢⱼ) at ccc, Δᶠ = ∛(ΔxΔyΔz),
                             ς² = max(0, 1 - Cb N²/(Pr ΣᵢⱼΣᵢⱼ)), N² = max(0, ℑz ∂z b); C, Cb, Pr in les[0..2] (Cb = 0: no stratification
                             correction, b unused) -- Oceananigans' calc_nonlinear_νᶜᶜᶜ for SmagorinskyLilly; ΣᵢⱼΣᵢⱼ as in FlowDiagnostics.jl:431-441 */
  /* ---- pass-through modules (§8f row 4): the Oceananigans tendency terms UMomentumEquation / VMomentumEquation /
   * WMomentumEquation / TracerEquation hand out one by one, at the velocity points (fcc / cfc / ccf) or the cell centre.  The
   * term is picked by OCB_REQ_TERM(code) in the request flags: OCB_TERM_TENDENCY the whole tendency (u_velocity_tendency /
   * tracer_tendency, the model without background fields, immersed boundaries, Stokes drift: as KE_TENDENCY / C_TENDENCY),
   * ADVECTION div_𝐯u / div_Uc (Centered 2nd order), BUOYANCY x_dot_g_b, CORIOLIS x_f_cross_U, PRESSURE ∂x of slot p
   * (hydrostatic_pressure_gradient_x), DIFFUSION ∂ⱼ_τ₁ⱼ / ∇_dot_qᶜ, FORCING the materialised functor (aux[0..2]; tracer: aux[0]).
   * UMomentumEquation.jl:99-504, VMomentumEquation.jl, WMomentumEquation.jl, TracerEquation.jl:70-327 */
  OCB_DIAG_U_TERM,        /* fcc */
  OCB_DIAG_V_TERM,        /* cfc */
  OCB_DIAG_W_TERM,        /* ccf */
  OCB_DIAG_C_TERM,        /* ccc: tracer in slot b */
  OCB_DIAG_QX, OCB_DIAG_QY, OCB_DIAG_QZ,   /* diffusive_flux_x / y / z of the tracer in slot b   TracerEquation.jl:221-294   fcc cfc ccf */
  OCB_DIAG_COUNT
};

/* request flags */
enum { OCB_TERM_TENDENCY = 0, OCB_TERM_ADVECTION = 1, OCB_TERM_BUOYANCY = 2, OCB_TERM_CORIOLIS = 3, OCB_TERM_PRESSURE = 4,
       OCB_TERM_DIFFUSION = 5, OCB_TERM_FORCING = 6 };
#define OCB_REQ_TERM_SHIFT 12
#define OCB_REQ_TERM(code) ((code) << OCB_REQ_TERM_SHIFT)     /* U_TERM / V_TERM / W_TERM / C_TERM: which term (bits 12..15) */
enum {
  OCB_REQ_PERTURBATION = 1,   /* velocities seen by this diagnostic are (u-U, v-V, w-W) and the tracer is c-C:
                                 perturbation_fields (src/Oceanostics.jl:157-173) / lazy `u-U` operands
                                 (src/TurbulentKineticEnergyEquation.jl:89-95, 327-330) evaluated on the fly */
  OCB_REQ_TENSOR_DIM1 = 1 << 4, OCB_REQ_TENSOR_DIM2 = 1 << 5, OCB_REQ_TENSOR_DIM3 = 1 << 6,  /* `dims` of PI (default all three) */
  OCB_REQ_COLLOCATED = 1 << 7, /* PI built with collocate_diagonals=true */
  OCB_REQ_CARRIED_HEIGHT = 1 << 8, /* APE on a VerticalSort column: the parcel's own height is the field aux[3], not the grid's z
                                      (4-argument local_ape_ccc, AvailablePotentialEnergyEquation.jl:43) */
  OCB_REQ_CENTERED4 = 1 << 10,     /* ADV / KE_TENDENCY / C_TENDENCY / U,V,W,C_TERM: Centered(order = 4) advection -- (-1, 7, 7, -1)/12 symmetric interpolation of
                                      the advecting and the advected quantity, order 2 next to the walls of a Bounded axis; default: Centered(order = 2) */
  OCB_REQ_HYDROSTATIC_SPLIT = 1 << 9 /* KE_TENDENCY / U,V,W_TERM tendency: the model carries a hydrostatic pressure anomaly pHY′ in slot p -- its
                                      horizontal gradient enters the u, v tendencies and the vertical momentum tendency has no buoyancy term
                                      (Oceananigans' w_velocity_tendency).  Without the flag the tendency does not read slot p at all. */
};

typedef struct {
  int32_t diag;             /* OCB_DIAG_* */
  int32_t flags;            /* OCB_REQ_* */
  void*   out;              /* device pointer to the destination PARENT array, or NULL (integral only) */
  int32_t out_halo[3];
  int64_t out_stride[3];
  int32_t window_lo[3];     /* 1-based inclusive index window of the destination (Field `indices=`; */
  int32_t window_hi[3];     /* SpatialFilters.jl:322-330).  lo = hi = 0  =>  the whole interior.      */
  int64_t out_offset[3];    /* index of the destination's first stored element along each axis (1 for a full field; the window start for a windowed one) */
  int32_t integral_slot;    /* >= 0: accumulate sum(value * V) into integrals[slot]; -1: none */
} ocb_request;

/* operands of one fused sweep: every request of a plan reads these */
#define OCB_N_AUX 7
typedef struct {
  ocb_field   u, v, w;      /* velocities at fcc / cfc / ccf (w may be ZERO: src/Oceanostics.jl:160-162) */
  ocb_field   b;            /* tracer slot: buoyancy b / tracer c / pressure-free scalar, ccc */
  ocb_field   p;            /* pressure, ccc */
  ocb_field   U, V, W, B;   /* mean flow / mean tracer: full, reduced, constant or zero */
  ocb_field   aux[OCB_N_AUX]; /* filtered fluxes etc. for FEPS / SFEPS / SFKE / PI / BPE */
  ocb_closure closure;
  double      f[3];         /* Coriolis fx, fy, fz (get_coriolis_frequency_components, src/Oceanostics.jl:178-192) */
  double      dir[3];       /* true vertical direction (Ri) or `direction` (DEPV) */
  double      f_dir;        /* DEPV: sum(f .* direction) */
  double      ghat[3];      /* -gravity_unit_vector (uᵢbᵢ); (0,0,1) by default */
  double      ro_bg[6];     /* Rossby background shears: dWdy, dVdz, dUdz, dWdx, dUdy, dVdx */
  double      les[4];       /* NU_SMAG: Smagorinsky coefficient C, buoyancy coefficient Cb (0 = none), turbulent Prandtl number Pr, reserved */
} ocb_sweep_inputs;

typedef struct ocb_plan ocb_plan;

/* ---- library ------------------------------------------------------------------------------- */
int         ocb_abi_version(void);
const char* ocb_last_error(void);
int         ocb_device_count(void);     /* number of usable CUDA devices (0 => every compute call fails) */

/* ---- K1: fused multi-diagnostic stencil sweep with fused volume integrals ------------------- */
/* Replaces launch!(arch, grid, params, _compute!, data, op) per KFO and sum!(…) per Integral
 * (Oceananigans compute!(::Field{…,<:KernelFunctionOperation}) / compute!(::ReducedComputedField)). */
int ocb_plan_create(const ocb_grid* grid, const ocb_sweep_inputs* inputs,
                    const ocb_request* requests, int32_t n_requests, ocb_plan** plan_out);
/* integrals: device pointer to double[n_slots] (overwritten; may be NULL when no request integrates) */
int ocb_plan_execute(ocb_plan* plan, double* integrals, void* stream);
int ocb_plan_destroy(ocb_plan* plan);
/* number of kernels one execute launches, the kernel variant chosen, and the algorithmic bytes it moves.
 * variant: 0 generic kernel, 1 budget kernel, 2 budget + generic, 3 velocity-gradient kernel,
 *          4 velocity-gradient + generic, 5 budget + velocity-gradient, 6 all three */
int ocb_plan_info(const ocb_plan* plan, int32_t* n_launches, int32_t* variant, double* algorithmic_bytes);

/* ---- halos: Oceananigans BoundaryConditions.fill_halo_regions! (default BCs) ------------------ */
/* Periodic: wrap all halo layers; Bounded Center-located/tangential: mirror the first layer (no-flux);
 * Bounded wall-normal Face: boundary faces set to 0 when `impenetrable` != 0.
 * `zgrad_bottom/top`: GradientBoundaryCondition values in z (NaN = none). */
int ocb_fill_halo(const ocb_grid* grid, const ocb_field* field, int32_t impenetrable,
                  double zgrad_bottom, double zgrad_top, void* stream);

/* ---- K2: separable spatial filters ---------------------------------------------------------- */
enum { OCB_FILTER_BOX = 0, OCB_FILTER_GAUSSIAN = 1, OCB_FILTER_GAUSSIAN_STRETCHED = 2 };
enum { OCB_POLICY_PERIODIC = 0, OCB_POLICY_SHRINK = 1, OCB_POLICY_EDGE = 2, OCB_POLICY_CONSTANT = 3 };
#define OCB_MAX_FILTER_WIDTH 64
typedef struct {
  int32_t dim;              /* 1, 2 or 3 */
  int32_t kind;             /* OCB_FILTER_* (BoxFilterKernel / GaussianFilterKernel / StretchedGaussianFilterKernel) */
  int32_t width;            /* half width; 2*width+1 taps */
  int32_t policy;           /* OCB_POLICY_* (SpatialFilters.jl:26-77) */
  int32_t extent;           /* operand's own length along dim: N, or N+1 for Face x Bounded (SizedBoundary) */
  double  left, right;      /* ConstantBoundary values */
  double  weights[2 * OCB_MAX_FILTER_WIDTH + 1];  /* uniform Gaussian weights (gaussian_weights, gaussian_filter.jl:312) */
  double  sigma;            /* stretched: σ in physical units */
  double  period;           /* stretched: domain length along dim (periodic image shift) */
  const void* nodes;        /* stretched: device array (grid FT) of node coordinates for indices 1..extent */
  const void* spacings;     /* stretched: device array (grid FT) of cell widths for indices 1..extent */
} ocb_filter_pass;

/* Staged evaluation of an n_pass (1..3) separable filter: _compute_staged_filter!
 * (SpatialFilters.jl:347-378).  src/dst are interior views described as fields (same location); the
 * destination may be a window.  Scratch for the intermediates is owned by the library (grown on
 * demand, stream-ordered). */
int ocb_filter_execute(int32_t dtype, const ocb_field* src, const int32_t src_extent[3],
                       const ocb_filter_pass* passes, int32_t n_pass,
                       const ocb_field* dst, const int32_t window_lo[3], const int32_t window_hi[3],
                       void* stream);

/* ---- K3: sort-based reference height z✶ (BackgroundPotentialEnergy) --------------------------- */
enum { OCB_BPE_THREE_DIMENSIONAL_SORT = 0, OCB_BPE_HEAVISIDE_INTEGRAL = 1 };
/* compute!(::SortedReferenceHeightField): rank_by_buoyancy! + slot_centers! + scatter +
 * fill_reference_potential! (BackgroundPotentialEnergyEquation.jl:310-466, 610-619).
 * b: ccc field; zstar, psi_delta: ccc destination fields (interior written; caller fills halos);
 * perm_out: optional device int64[N] (1-based ranks as `sortperm!` returns), b_sorted_out: optional FT[N]. */
int ocb_bpe_reference_height(const ocb_grid* grid, const ocb_field* b, int32_t method,
                             const ocb_field* zstar, const ocb_field* psi_delta,
                             int64_t* perm_out, void* b_sorted_out, void* stream);
/* VerticalSort (BackgroundPotentialEnergyEquation.jl:584-597): a model-grid ccc field read in sorted order,
 * dst[r] = src[perm[r]] with the 1-based `sortperm!` ranks ocb_bpe_reference_height returns (s.source_height[perm], z✶, Ψδ). */
int ocb_gather_sorted(int32_t dtype, const ocb_field* src, const int64_t* perm, void* dst, int64_t n, void* stream);
/* assign_reference_height!(::ProfileLookup) (BackgroundPotentialEnergyEquation.jl:550-582): match every cell to a sorted
 * profile (b_profile, z_profile: device FT[m], both non-decreasing -- validated by the host as lookup_profile :528-547
 * does) by binary search; slot faces are the midpoints of z_profile closed by the domain's bottom and top (:497-507). */
int ocb_bpe_profile_lookup(const ocb_grid* grid, const ocb_field* b, const void* b_profile, const void* z_profile,
                           int64_t m, const ocb_field* zstar, const ocb_field* psi_delta, void* stream);
/* the radix sort alone (keys FT[n] -> sorted keys + 0-based permutation), exposed for tests/benchmarks */
int ocb_sort_pairs(int32_t dtype, const void* keys_in, void* keys_out, uint32_t* perm_out,
                   int64_t n, void* stream);

/* ---- pointwise helper: Field(a ∘ b) with each operand interpolated to the destination location - */
enum { OCB_OP_COPY = 0, OCB_OP_ADD = 1, OCB_OP_SUB = 2, OCB_OP_MUL = 3 };
int ocb_pointwise(const ocb_grid* grid, int32_t op, double alpha, const ocb_field* a, const ocb_field* b,
                  const ocb_field* dst, void* stream);

/* ---- multi-GPU slab support (y-slabs, one rank per GPU) ---------------------------------------- */
/* pack / unpack `h` y-planes next to the low (side=0) or high (side=1) edge of the interior into a
 * contiguous buffer of size[0]*h*size[2] elements; the exchange itself is NCCL send/recv issued by the host
 * layer on the same stream. */
int ocb_halo_pack_y(int32_t dtype, const ocb_field* field, int32_t side, int32_t h, void* buffer, void* stream);
int ocb_halo_unpack_y(int32_t dtype, const ocb_field* field, int32_t side, int32_t h, const void* buffer, void* stream);
/* the same for both edges of up to 8 fields in one launch: buffers[2*f + side]; unpack = 0 packs the interior edge
 * planes, unpack = 1 writes the buffers into the halo planes (the overlapped slab step of the multi-GPU path) */
int ocb_halo_pack_y_many(int32_t dtype, const ocb_field* fields, int32_t n_fields, int32_t h, void* const* buffers,
                         int32_t unpack, void* stream);

/* ---- direct peer-memory halo exchange (one process per GPU of one box, CUDA IPC over NVLink) ----
 * ocb_ipc_export: the 64-byte cudaIpcMemHandle_t of the allocation holding `ptr` and ptr's byte offset inside it (the host
 * layer sends both to the neighbour ranks); ocb_ipc_import maps a received handle (peer access enabled lazily) and returns
 * the allocation base; ocb_ipc_close unmaps it.
 * ocb_halo_push_y_many: ONE kernel stores the `h` edge rows of every field into the neighbours' halo rows --
 * peer_arrays[2*f + side] is the neighbour's parent array of field f (side 0 = low-y neighbour, 1 = high-y neighbour),
 * peer_ny[2*f + side] its interior length along y -- then publishes `step` in the neighbours' flag words (lo_flag = the low
 * neighbour's flags[1], hi_flag = the high neighbour's flags[0]).
 * ocb_halo_wait: stream-ordered wait until both of this rank's flag words (flags[0] from the low neighbour, flags[1] from
 * the high one) reach `step`; after `timeout_seconds` it sets flags[2] = 1 and returns (the host layer checks it).
 * Ordering contract: a push for step k+1 must be enqueued after a collective every rank enters after its step-k reads of the
 * halo rows (e.g. the all-reduce of the integrals). */
int ocb_ipc_export(const void* ptr, void* handle64, int64_t* offset);
int ocb_ipc_import(const void* handle64, void** base);
int ocb_ipc_close(void* base);
int ocb_halo_push_y_many(int32_t dtype, const ocb_field* fields, int32_t n_fields, int32_t h, void* const* peer_arrays,
                         const int32_t* peer_ny, void* lo_flag, void* hi_flag, int64_t step, void* stream);
int ocb_halo_wait(void* flags, int64_t step, double timeout_seconds, void* stream);

/* ---- peer-memory ring: the slab step as ONE gated sweep + ONE reduction kernel (no NCCL on the data path) -----------
 * This is the survey's `ocb_comm_init` seam in peer-memory form: the host layer (Julia: MPI / NCCL
 * bootstrap of Oceananigans' Distributed; Python: torch.distributed) only carries the IPC handles at set-up.
 * Every rank owns one zero-initialised device "sync block" of OCB_RING_SYNC_WORDS 64-bit words and maps every other
 * rank's block (ocb_ipc_export / ocb_ipc_import):
 *   word 0, 1   halo flags: the step number of the last push from the low / high y neighbour (ocb_halo_push_y_many's
 *               lo_flag = low neighbour's word 1, hi_flag = high neighbour's word 0)
 *   word 2      set to 1 when a wait gave up (also: the integrals read NaN on every rank)
 *   word 8 + 16 p + r          "rank r's partial integrals of a step with parity p have landed" (the step number)
 *   word 64 + 16 (16 p + r) + s  rank r's partial integral of slot s (a double), parity p
 * ocb_plan_execute_ring runs the plan's sweep with the tiles that read a neighbour's halo row placed last in launch order
 * and made to wait (in the kernel) for that neighbour's push of `step`, so the sweep of the rows that need no neighbour
 * data overlaps the exchange; one more kernel then sums this rank's per-block partials, stores them into every rank's sync
 * block over NVLink, waits for all ranks' and adds them in rank order -- the same bits on every rank.  `step` must
 * increase by one per call and match the `step` of the pushes; the reduction is also the collective that orders the next
 * push after this step's halo reads (the ordering contract above).  Plans it accepts: integrals only, evaluated by the
 * class-sum budget kernel over the whole y extent of a y-periodic (slab) grid; anything else is OCB_ERR_UNSUPPORTED and
 * the host layer runs exchange + ocb_plan_execute + its own all-reduce instead. */
#define OCB_MAX_RANKS 16
#define OCB_RING_SYNC_WORDS 1024
typedef struct {
  int32_t rank, nranks;
  void*   sync[OCB_MAX_RANKS];    /* sync[r]: rank r's sync block as mapped in this process (sync[rank] = its own) */
  double  timeout_seconds;        /* a wait that lasts longer gives up (never a hang) */
} ocb_ring;
int ocb_plan_ring_capable(const ocb_plan* plan);   /* 1 if ocb_plan_execute_ring accepts this plan */
int ocb_plan_execute_ring(ocb_plan* plan, double* integrals, const ocb_ring* ring, int64_t step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OCEANOSTICS_B200_H */
