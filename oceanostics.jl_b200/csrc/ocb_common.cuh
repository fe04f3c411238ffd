// Shared host/device plumbing of liboceanostics_b200 (sm_100a).
//
// Nothing here mirrors reference code: the reference (tomchor/Oceanostics.jl) has no native layer.  The
// structures below are the device-side image of the C-ABI descriptors in include/oceanostics_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include "../../include/oceanostics_b200.h"

#define OCB_DF __device__ __forceinline__
#define OCB_HD __host__ __device__ __forceinline__

// ---- error plumbing (thread-local message, integer status; nothing throws across the ABI) -------------
void ocb_set_error(const char* fmt, ...);
#define OCB_FAIL(code, ...) do { ocb_set_error(__VA_ARGS__); return (code); } while (0)
#define OCB_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
    ocb_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    return OCB_ERR_CUDA; } } while (0)
#define OCB_REQUIRE(cond, ...) do { if (!(cond)) OCB_FAIL(OCB_ERR_INVALID, __VA_ARGS__); } while (0)

int ocb_require_device();   // OCB_OK or OCB_ERR_CUDA ("no CUDA device": there is no CPU fallback)
int ocb_sm_count();
void ocb_keep_scratch_pool();   // make the stream-ordered scratch pool retain its memory between calls

// Every entry point runs on the device that OWNS the caller's arrays, whatever the calling thread's current device is
// (plan scratch, tables and launches would otherwise land on the current device next to another device's pointers and stream).
// The guard switches to the owner of `p` (or to an explicit device index) and restores the previous device on scope exit.
struct OcbDeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit OcbDeviceGuard(const void* p) {
    if (!p) return;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return; }
    if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged) return;
    enter(a.device);
  }
  explicit OcbDeviceGuard(int device) { if (device >= 0) enter(device); }
  void enter(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); return; }
    if (device != prev && cudaSetDevice(device) == cudaSuccess) switched = true;
  }
  ~OcbDeviceGuard() { if (switched) cudaSetDevice(prev); }
  OcbDeviceGuard(const OcbDeviceGuard&) = delete;
  OcbDeviceGuard& operator=(const OcbDeviceGuard&) = delete;
};

// ---- device view of one operand ------------------------------------------------------------------------
// p is pre-offset so that the 1-based (Oceananigans) index (i, j, k) lives at p[i*sx + j*sy + k*sz];
// reduced axes have stride 0; ZeroField / Number operands have p == nullptr and value c.
// (host+device: the same accessor is compiled into tests/hostcheck for CPU verification of the formulas)
#ifdef __CUDA_ARCH__
#define OCB_LD(ptr) __ldg(ptr)
#else
#define OCB_LD(ptr) (*(ptr))
#endif
#define OCB_X __host__ __device__ __forceinline__
template <class T>
struct HFld {
  const T* p;
  long long sx, sy, sz;
  T c;
  OCB_X T operator()(int i, int j, int k) const { return p ? OCB_LD(p + (i * sx + j * sy + k * sz)) : c; }
  OCB_X bool is_array() const { return p != nullptr; }
};

template <class T>
static inline int make_fld(const ocb_field& f, HFld<T>* out, const char* name, bool allow_null = true) {
  HFld<T> r;
  r.p = nullptr; r.sx = r.sy = r.sz = 0; r.c = T(0);
  if (f.kind == OCB_FIELD_ZERO || (f.kind == OCB_FIELD_ARRAY && f.data == nullptr)) {
    if (f.kind == OCB_FIELD_ARRAY && !allow_null) OCB_FAIL(OCB_ERR_INVALID, "operand `%s` has no data", name);
  } else if (f.kind == OCB_FIELD_CONSTANT) {
    r.c = T(f.value);
  } else if (f.kind == OCB_FIELD_ARRAY) {
    long long off = 0;
    long long s[3];
    for (int d = 0; d < 3; ++d) {
      s[d] = (f.loc[d] == OCB_NOTHING) ? 0 : f.stride[d];
      off += (long long)(f.halo[d] - 1) * s[d];
    }
    r.p = (const T*)f.data + off;
    r.sx = s[0]; r.sy = s[1]; r.sz = s[2];
  } else {
    OCB_FAIL(OCB_ERR_INVALID, "operand `%s`: unknown kind %d", name, f.kind);
  }
  *out = r;
  return OCB_OK;
}

// ---- device view of the grid ---------------------------------------------------------------------------
// x and y are regularly spaced (RectilinearGrid with stretched z only, as the reference's own stretched
// test grid: test/test_utils.jl:17-24).  z metrics are always read from library-owned tables indexed by
// the 1-based k (valid for k = 1-Hz .. Nz+Hz+1), filled with the constant on a regular axis.
template <class T>
struct DGrid {
  int Nx, Ny, Nz;
  int tx, ty, tz;              // topology
  T dx, dy, rdx, rdy;
  const T* dzc;  const T* dzf;   // Δzᵃᵃᶜ[k], Δzᵃᵃᶠ[k]
  const T* rdzc; const T* rdzf;  // their reciprocals
  const T* zc;   const T* zf;    // nodes
  int oi, oj, ok;                // absolute index of the origin the z tables / operand pointers are currently shifted to (0 when unshifted)
};

static inline size_t ocb_sizeof(int dtype) { return dtype == OCB_F64 ? 8 : 4; }

// deterministic two-stage reduction of per-block partial sums: partial[b * nslots + s] -> out[s]
void ocb_launch_finalize_partials(const double* partial, int nblocks, int nslots, double* out, cudaStream_t st);
// ... followed by the all-reduce over the ranks of a peer-memory ring (one kernel; ocb_api.cu)
int ocb_launch_finalize_ring(const double* partial, int nblocks, int nslots, double* out, const ocb_ring* ring, long long step,
                             long long timeout_cycles, cudaStream_t st);
long long ocb_cycles(double seconds);   // SM clock cycles of the current device in `seconds` (spin-wait bounds)
