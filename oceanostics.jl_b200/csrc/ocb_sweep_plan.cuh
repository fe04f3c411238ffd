// Host-side translation of the C-ABI descriptors of one fused sweep into the device parameter block.
// Shared by the product library (ocb_sweep.cu) and by the CPU verification harness under
// tests/hostcheck/ (which runs the same evaluators on the host against the NumPy oracle).
#pragma once
#include "ocb_diag.cuh"

// location of each diagnostic's output (C = 0, F = 1), in OCB_DIAG_* order
static inline void ocb_diag_location(int diag, int loc[3]) {
  const int C = OCB_CENTER, F = OCB_FACE;
  int l[3] = {C, C, C};
  switch (diag) {
    case OCB_DIAG_S12: case OCB_DIAG_O12: case OCB_DIAG_T12: case OCB_DIAG_VF12: l[0] = F; l[1] = F; break;
    case OCB_DIAG_S13: case OCB_DIAG_O13: case OCB_DIAG_T13: case OCB_DIAG_VF13: l[0] = F; l[2] = F; break;
    case OCB_DIAG_S23: case OCB_DIAG_O23: case OCB_DIAG_T23: case OCB_DIAG_VF23: l[1] = F; l[2] = F; break;
    case OCB_DIAG_T11: case OCB_DIAG_U_TERM: case OCB_DIAG_QX: l[0] = F; break;
    case OCB_DIAG_T22: case OCB_DIAG_V_TERM: case OCB_DIAG_QY: l[1] = F; break;
    case OCB_DIAG_T33: case OCB_DIAG_W_TERM: case OCB_DIAG_QZ: l[2] = F; break;
    case OCB_DIAG_RI: l[2] = F; break;
    case OCB_DIAG_RO: case OCB_DIAG_EPV: case OCB_DIAG_TWPV: case OCB_DIAG_DEPV: l[0] = l[1] = l[2] = F; break;
    default: break;
  }
  loc[0] = l[0]; loc[1] = l[1]; loc[2] = l[2];
}

static inline const char* ocb_diag_name(int diag) {
#define OCB_NAME_CASE(D) case D: return #D;
  switch (diag) { OCB_FOR_EACH_DIAG(OCB_NAME_CASE) default: return "?"; }
#undef OCB_NAME_CASE
}

// which operands a diagnostic cannot do without (bit set => must be an ARRAY operand)
enum { OCB_NEED_B = 1, OCB_NEED_P = 2, OCB_NEED_AUX05 = 4, OCB_NEED_AUX6 = 8, OCB_NEED_AUX5 = 16, OCB_NEED_AUX4 = 32 };
static inline int ocb_diag_needs(int diag) {
  switch (diag) {
    case OCB_DIAG_PR: return OCB_NEED_P;
    case OCB_DIAG_WB: case OCB_DIAG_CHI: case OCB_DIAG_CDIFF: case OCB_DIAG_RI: case OCB_DIAG_EPV:
    case OCB_DIAG_TWPV: case OCB_DIAG_DEPV: case OCB_DIAG_C_TENDENCY: case OCB_DIAG_C_TERM: case OCB_DIAG_QX: case OCB_DIAG_QY:
    case OCB_DIAG_QZ: return OCB_NEED_B;
    case OCB_DIAG_FEPS: case OCB_DIAG_PI: return OCB_NEED_AUX05;
    case OCB_DIAG_SFEPS: return OCB_NEED_AUX05 | OCB_NEED_AUX6;
    case OCB_DIAG_SFKE: return OCB_NEED_AUX6;
    case OCB_DIAG_BPE: return OCB_NEED_B | OCB_NEED_AUX6;
    case OCB_DIAG_UPSILON: return OCB_NEED_AUX6;
    case OCB_DIAG_APE_EPS: return OCB_NEED_B | OCB_NEED_AUX4;
    case OCB_DIAG_APE: return OCB_NEED_B | OCB_NEED_AUX6 | OCB_NEED_AUX5;
    default: return 0;
  }
}

// how many halo cells a diagnostic reads beyond the interior along each axis: 2 for the terms that
// difference a flux which itself reaches one cell out (Centered second-order momentum advection, the viscous-flux divergence,
// the tendencies built from them) and for the tilted Richardson number, 1 for everything else
static inline int ocb_diag_reach(int diag, int flags = 0) {
  if (flags & OCB_REQ_CENTERED4) return 3;             // fourth-order fluxes reach one cell further than second-order ones
  switch (diag) {
    case OCB_DIAG_ADV: case OCB_DIAG_KE_STRESS: case OCB_DIAG_KE_TENDENCY: case OCB_DIAG_U_TERM: case OCB_DIAG_V_TERM:
    case OCB_DIAG_W_TERM: case OCB_DIAG_RI: return 2;
    default: return 1;
  }
}

template <class T>
struct ZTables {     // pointers pre-offset so that table[k] is valid for k = 1-Hz .. Nz+Hz+1
  const T *dzc, *dzf, *rdzc, *rdzf, *zc, *zf;
};

template <class T>
static int ocb_build_sweep_params(const ocb_grid* grid, const ocb_sweep_inputs* in, const ocb_request* reqs,
                                  int n, const ZTables<T>& zt, SweepParams<T>* out) {
  OCB_REQUIRE(grid && in && reqs && out, "null argument");
  OCB_REQUIRE(n >= 1 && n <= OCB_MAX_REQ, "a plan takes 1..%d requests; got %d", OCB_MAX_REQ, n);
  SweepParams<T>& P = *out;
  memset(&P, 0, sizeof(P));
  for (int d = 0; d < 3; ++d) {
    OCB_REQUIRE(grid->N[d] >= 1 && grid->H[d] >= 1, "grid size / halo must be >= 1 along axis %d", d + 1);
    OCB_REQUIRE(grid->topo[d] == OCB_PERIODIC || grid->topo[d] == OCB_BOUNDED, "bad topology along axis %d", d + 1);
  }
  OCB_REQUIRE(grid->d[0] > 0 && grid->d[1] > 0, "x and y spacings must be positive");
  DGrid<T>& g = P.g;
  g.Nx = grid->N[0]; g.Ny = grid->N[1]; g.Nz = grid->N[2];
  g.tx = grid->topo[0]; g.ty = grid->topo[1]; g.tz = grid->topo[2];
  g.dx = T(grid->d[0]); g.dy = T(grid->d[1]);
  g.rdx = T(1) / g.dx; g.rdy = T(1) / g.dy;
  g.dzc = zt.dzc; g.dzf = zt.dzf; g.rdzc = zt.rdzc; g.rdzf = zt.rdzf; g.zc = zt.zc; g.zf = zt.zf;

  SweepIn<T>& I = P.in;
  int rc;
#define OCB_MK(dst, src, name) if ((rc = make_fld<T>(src, &dst, name)) != OCB_OK) return rc;
  OCB_MK(I.u, in->u, "u") OCB_MK(I.v, in->v, "v") OCB_MK(I.w, in->w, "w") OCB_MK(I.b, in->b, "b") OCB_MK(I.p, in->p, "p")
  OCB_MK(I.U, in->U, "U") OCB_MK(I.V, in->V, "V") OCB_MK(I.W, in->W, "W") OCB_MK(I.B, in->B, "B")
  for (int a = 0; a < OCB_N_AUX; ++a) OCB_MK(I.aux[a], in->aux[a], "aux")
  const ocb_closure& c = in->closure;
  OCB_REQUIRE(c.n_nu_fields >= 0 && c.n_nu_fields <= OCB_MAX_CLOSURE_FIELDS && c.n_kappa_fields >= 0 &&
              c.n_kappa_fields <= OCB_MAX_CLOSURE_FIELDS, "closure: at most %d eddy fields", OCB_MAX_CLOSURE_FIELDS);
  I.nu_c = T(c.nu_const); I.n_nu = c.n_nu_fields;
  I.ka_c = T(c.kappa_const); I.n_ka = c.n_kappa_fields;
  for (int f = 0; f < c.n_nu_fields; ++f) {
    OCB_MK(I.nu_f[f], c.nu_fields[f], "closure.nu_fields")
    OCB_REQUIRE(I.nu_f[f].p, "closure.nu_fields[%d] must be an array operand", f);
  }
  for (int f = 0; f < c.n_kappa_fields; ++f) {
    OCB_MK(I.ka_f[f], c.kappa_fields[f], "closure.kappa_fields")
    OCB_REQUIRE(I.ka_f[f].p, "closure.kappa_fields[%d] must be an array operand", f);
    I.ka_s[f] = T(c.kappa_scale[f]);
  }
#undef OCB_MK
  for (int d = 0; d < 3; ++d) { I.f[d] = T(in->f[d]); I.dir[d] = T(in->dir[d]); I.ghat[d] = T(in->ghat[d]); }
  I.f_dir = T(in->f_dir);
  for (int d = 0; d < 6; ++d) I.ro_bg[d] = T(in->ro_bg[d]);
  for (int d = 0; d < 4; ++d) I.les[d] = T(in->les[d]);

  P.nreq = n;
  int nslots = 0;
  for (int d = 0; d < 3; ++d) { P.lo[d] = 1 << 30; P.hi[d] = -(1 << 30); }
  for (int r = 0; r < n; ++r) {
    const ocb_request& q = reqs[r];
    DReq& R = P.req[r];
    OCB_REQUIRE(q.diag >= 0 && q.diag < OCB_DIAG_COUNT, "request %d: unknown diagnostic id %d", r, q.diag);
    R.diag = q.diag; R.flags = q.flags; R.slot = q.integral_slot;
    if ((q.flags & OCB_REQ_PERTURBATION) && !ocb_diag_takes_perturbation(q.diag))
      OCB_FAIL(OCB_ERR_UNSUPPORTED, "request %d (%s) does not take lazy perturbation operands: materialise Field(u - U) first", r, ocb_diag_name(q.diag));
    ocb_diag_location(q.diag, R.loc);
    OCB_REQUIRE(q.out != nullptr || q.integral_slot >= 0, "request %d (%s) has neither an output field nor an integral slot",
                r, ocb_diag_name(q.diag));
    OCB_REQUIRE(q.integral_slot < OCB_MAX_REQ, "request %d: integral slot %d out of range", r, q.integral_slot);
    for (int r2 = 0; r2 < r; ++r2)
      OCB_REQUIRE(q.integral_slot < 0 || reqs[r2].integral_slot != q.integral_slot, "requests %d and %d share integral slot %d", r2, r, q.integral_slot);
    if (q.integral_slot + 1 > nslots) nslots = q.integral_slot + 1;
    const int needs = ocb_diag_needs(q.diag);
    if (needs & OCB_NEED_B) OCB_REQUIRE(I.b.p, "%s needs the tracer/buoyancy operand `b` as an array", ocb_diag_name(q.diag));
    if (needs & OCB_NEED_P) OCB_REQUIRE(I.p.p, "%s needs the pressure operand `p` as an array", ocb_diag_name(q.diag));
    if (needs & OCB_NEED_AUX6) OCB_REQUIRE(I.aux[6].p, "%s needs aux[6] as an array", ocb_diag_name(q.diag));
    if (needs & OCB_NEED_AUX5) OCB_REQUIRE(I.aux[5].p, "%s needs aux[5] as an array", ocb_diag_name(q.diag));
    if (needs & OCB_NEED_AUX4) OCB_REQUIRE(I.aux[4].p, "%s needs aux[4] as an array", ocb_diag_name(q.diag));
    if (q.diag == OCB_DIAG_NU_SMAG) {
      OCB_REQUIRE(in->les[0] > 0, "NU_SMAG needs the Smagorinsky coefficient C > 0 in les[0]");
      if (in->les[1] != 0) OCB_REQUIRE(I.b.p && in->les[2] > 0, "NU_SMAG with a buoyancy coefficient Cb needs the buoyancy operand `b` as an array and Pr > 0");
    }
    if (q.diag == OCB_DIAG_APE && (q.flags & OCB_REQ_CARRIED_HEIGHT)) OCB_REQUIRE(I.aux[3].p, "APE with a carried height needs aux[3] as an array");
    if (needs & OCB_NEED_AUX05) {
      const bool d1 = q.diag != OCB_DIAG_PI || (q.flags & OCB_REQ_TENSOR_DIM1), d2 = q.diag != OCB_DIAG_PI || (q.flags & OCB_REQ_TENSOR_DIM2),
                 d3 = q.diag != OCB_DIAG_PI || (q.flags & OCB_REQ_TENSOR_DIM3);
      const bool need[6] = {d1, d2, d3, d1 && d2, d1 && d3, d2 && d3};
      for (int a = 0; a < 6; ++a) if (need[a]) OCB_REQUIRE(I.aux[a].p, "%s needs aux[%d] as an array", ocb_diag_name(q.diag), a);
      if (q.diag == OCB_DIAG_PI) OCB_REQUIRE(d1 || d2 || d3, "cross-scale flux needs at least one tensor dimension flag");
    }
    const int topo[3] = {grid->topo[0], grid->topo[1], grid->topo[2]};
    for (int d = 0; d < 3; ++d) {
      const int ext = grid->N[d] + ((R.loc[d] == OCB_FACE && topo[d] == OCB_BOUNDED) ? 1 : 0);
      int lo = q.window_lo[d], hi = q.window_hi[d];
      if (lo == 0 && hi == 0) { lo = 1; hi = ext; }
      OCB_REQUIRE(lo >= 1 && hi <= ext && lo <= hi, "request %d (%s): window [%d, %d] outside 1..%d along axis %d", r,
                  ocb_diag_name(q.diag), lo, hi, ext, d + 1);
      R.lo[d] = lo; R.hi[d] = hi;
      if (lo < P.lo[d]) P.lo[d] = lo;
      if (hi > P.hi[d]) P.hi[d] = hi;
    }
    R.out = nullptr;
    if (q.out) {
      long long off = 0;
      for (int d = 0; d < 3; ++d) {
        R.os[d] = q.out_stride[d];
        const long long first = q.out_offset[d] == 0 ? 1 : q.out_offset[d];
        off += ((long long)q.out_halo[d] - first) * q.out_stride[d];
      }
      R.out = (void*)((T*)q.out + off);
    }
  }
  P.nslots = nslots;
  // every array operand must carry the halo the requested stencils reach into (a thin halo would be read out of bounds)
  int reach = 1;
  for (int r = 0; r < n; ++r) if (ocb_diag_reach(reqs[r].diag, reqs[r].flags) > reach) reach = ocb_diag_reach(reqs[r].diag, reqs[r].flags);
  const ocb_field* arr[] = {&in->u, &in->v, &in->w, &in->b, &in->p, &in->U, &in->V, &in->W, &in->B, &in->aux[0], &in->aux[1], &in->aux[2],
                            &in->aux[3], &in->aux[4], &in->aux[5], &in->aux[6], &in->closure.nu_fields[0], &in->closure.nu_fields[1],
                            &in->closure.kappa_fields[0], &in->closure.kappa_fields[1]};
  const char* names[] = {"u", "v", "w", "b", "p", "U", "V", "W", "B", "aux[0]", "aux[1]", "aux[2]", "aux[3]", "aux[4]", "aux[5]", "aux[6]",
                         "closure.nu_fields[0]", "closure.nu_fields[1]", "closure.kappa_fields[0]", "closure.kappa_fields[1]"};
  for (int a = 0; a < 20; ++a) {
    const ocb_field& f = *arr[a];
    if (f.kind != OCB_FIELD_ARRAY || !f.data) continue;
    if (a >= 16 && a < 18 && a - 16 >= c.n_nu_fields) continue;
    if (a >= 18 && a - 18 >= c.n_kappa_fields) continue;
    const int need = a < 5 ? reach : 1;                 // means, aux and closure fields are read one cell out at most
    for (int d = 0; d < 3; ++d) {
      if (f.loc[d] == OCB_NOTHING) continue;
      OCB_REQUIRE(f.halo[d] >= need, "operand `%s` has a halo of %d along axis %d, but the requested diagnostics read %d cell(s) beyond the interior",
                  names[a], f.halo[d], d + 1, need);
    }
  }
  return OCB_OK;
}
