// Shared plumbing for libemk: error reporting, launch checks, small device helpers.
#pragma once

#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cmath>

#include "../../include/emk.h"

namespace emk {

// thread-local message for emk_last_error()
char* last_error_buffer();
int fail(int code, const char* fmt, ...);

#define EMK_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) return ::emk::fail((int)e_, "%s failed: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

#define EMK_REQUIRE(cond, code, ...)                 \
  do {                                               \
    if (!(cond)) return ::emk::fail((code), __VA_ARGS__); \
  } while (0)

inline int launch_status(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, "launch of %s failed: %s", what, cudaGetErrorString(e));
  return EMK_OK;
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();
// stream-ordered scratch from a private per-device pool that keeps its memory (release with cudaFreeAsync)
int scratch_alloc(void** ptr, size_t bytes, cudaStream_t st);

// one-time per-device kernel configuration (cudaFuncSetAttribute is per device): returns true exactly once
// per (call site, device); call sites pass their own static flag array
constexpr int kMaxDevices = 64;
bool first_use_on_device(bool* flags);

// ---- sketch-map sigmoid, evaluated from the SQUARED distance ---------------------------------
// s(r) = 1 - (1 + c (r/sig)^a)^(-b/a),  c = 2^(a/b) - 1      (encodermap/misc/distances.py:86)
// Everything is a function of r^2: (r/sig)^a = (r^2/sig^2)^(a/2), so no sqrt is needed unless a
// is an odd integer.  Integer fast paths cover every parameter set the reference ships
// (a in {1,2,3,4,6,12}, b/a in {1/2,1,3/2,2,3,6}); anything else goes through powf.
struct SigSpec {
  float inv_sig2;   // 1/sig^2
  float c;          // 2^(a/b) - 1
  float half_a;     // a/2
  float e;          // b/a
  float dcoef;      // b*c/sig^2  (prefactor of s'(r)/r)
  int a_int;        // a if a is an integer in [1,17], else 0 (0 => generic powf path)
  int e2_int;       // 2*b/a if that is an integer in [1,15], else 0
};

SigSpec make_sig_spec(float sig, float a, float b);

// arguments of the force/torque backward kernel (backmap.cu), filled by the entry points in emk_api.cu
struct BwdParams {
  const float* lengths;   // may be null when no planar terms are needed
  int64_t lstride;
  const float* angles;    // (b, n-2) or null
  const float* xyz;       // (b, n, 3) final coordinates; null in planar mode (recomputed from lengths/angles)
  const float* grad_xyz;  // (b, n, 3)
  int64_t b;
  int n;
  int mid;                // first atom whose hinge/bond is "right of the anchor": n/2 (layer), 0 (planar / one-way)
  int dr0;                // first dihedral handled as right-side twist: n/2-1 (two-sided), 0 (one-way)
  int planar;             // 1: chain_in_plane backward (xyz recomputed, hinge normal = -(-1)^j e_z)
  float* grad_angles;     // (b, n-2) or null
  float* grad_dihedrals;  // (b, n-3) or null
  float* grad_lengths;    // (b, n-1) or null
};

#ifdef __CUDACC__
// 1/sqrt(x) as ONE MUFU.RSQ (2 ulp).  CUDA's rsqrtf() wraps the same instruction in a denormal fix-up (two compares,
// a select and three multiplies on every call); distances below 1e-19 are zero for every purpose here, so callers
// guard with x >= EMK_TINY instead.
#define EMK_TINY 1.17549435e-38f
__device__ __forceinline__ float rsqrt_fast(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// x^n for a kernel-uniform n in [0,15]: branch-free binary exponentiation (3 squarings, no MUFU).  Kept tiny on
// purpose: the pair-tile epilogue instantiates it 128 times and must stay inside the instruction cache.
__device__ __forceinline__ float ipow_uniform(float x, int n) {
  float r = (n & 1) ? x : 1.f;
  x *= x;
  r = (n & 2) ? r * x : r;
  x *= x;
  r = (n & 4) ? r * x : r;
  x *= x;
  r = (n & 8) ? r * x : r;
  return r;
}
// generic exponents go through one shared out-of-line powf (never taken by the shipped parameter sets)
static __device__ __noinline__ float pow_generic(float x, float y) { return powf(x, y); }

// value of the sigmoid at squared distance r2; if WITH_D also returns s'(r)/r in *dfac
// (finite for a >= 2; callers zero it when r2 == 0, the reference's zero-distance mask).
template <bool WITH_D>
__device__ __forceinline__ float sig_eval(float r2, const SigSpec& s, float* dfac) {
  const float u2 = r2 * s.inv_sig2;
  float p, pm1;  // p = u^a, pm1 = u^(a-2)
  if (s.a_int > 0) {
    if ((s.a_int & 1) == 0) {
      pm1 = ipow_uniform(u2, (s.a_int >> 1) - 1);
      p = pm1 * u2;
    } else {
      const float u = sqrtf(u2);
      if (s.a_int == 1) {
        p = u;
        pm1 = WITH_D ? (1.f / u) : 0.f;
      } else {
        pm1 = ipow_uniform(u, s.a_int - 2);
        p = pm1 * u2;
      }
    }
  } else {
    p = pow_generic(u2, s.half_a);
    pm1 = WITH_D ? (p / u2) : 0.f;
  }
  const float q = fmaf(s.c, p, 1.f);
  float qe;  // q^(-b/a)
  if (s.e2_int > 0) {
    const float t = (s.e2_int & 1) ? 1.f / sqrtf(q) : 1.f / q;
    qe = ipow_uniform(t, (s.e2_int & 1) ? s.e2_int : (s.e2_int >> 1));
  } else {
    qe = pow_generic(q, -s.e);
  }
  if (WITH_D) *dfac = s.dcoef * pm1 * qe / q;
  return 1.f - qe;
}
#endif

}  // namespace emk
