// Back-mapping kernels for sm_100a: internal coordinates -> Cartesian as an SE(3) prefix product
// along the atom chain, and the exact VJP from wrench (force / torque) recurrences.
//
// Forward: the chain is cut into per-lane contiguous chunks; each lane multiplies its chunk's transforms
// sequentially (pass 1), the chunk aggregates are combined with a shuffle scan, and each lane maps its
// chunk-local results through its prefix (pass 2).  The transform chain is float64 (inputs/outputs float32): in
// float32 the orientation error random-walks along the chain and exceeds the 1e-4 nm parity budget beyond ~100
// residues.
//
// Forward formulations (tools/proto_backmap.py checks all of them against the oracle):
//   * emk_backmap   : BackMapLayer.  NeRF placement from (L, theta, phi) with LOCAL transforms
//                     R <- R Rx(phi) Rz(pi - theta),  p <- p + L R e_x, built outward in both
//                     directions from the three middle atoms, which sit at their planar-chain positions.
//   * emk_dihedrals_to_cartesian : arbitrary start chain.  A_i = rotation by dihedral_i about the bond
//                     (t_{i+1}, t_{i+2}) of the START chain, C_i = C_{i-1} o A_i, out_k = C_{k-3}(start_k).
//   * emk_chain_in_plane : SE(2) prefix product (alternating-sign turns) + prefix sum of bond vectors.
// Backward: for every internal coordinate one end of the chain moves rigidly (twist about a bond, hinge
// about the normal of a bond angle, slide along a bond; left of the anchor the whole molecule also follows the
// planar chain), so each gradient is <axis, torque> or <direction, force> of that end's wrench, carried in
// float32 about a pivot that moves with the walk (see "backward, version 3").
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <mutex>

#include <cuda.h>

#include "emk_common.cuh"

namespace emk {

// ---- float64 sin/cos of a float32-exact argument, ~1e-14 absolute -----------------------------------
__device__ __forceinline__ void sincos_d(double x, double* s, double* c) {
  const double q = rint(x * 0.63661977236758134308);  // 2/pi
  double r = fma(-q, 1.5707963267948966, x);
  r = fma(-q, 6.123233995736766e-17, r);
  const double r2 = r * r;
  double sp = 1.6059043836821613e-10;                 // 1/13!
  sp = fma(sp, r2, -2.5052108385441720e-08);          // -1/11!
  sp = fma(sp, r2, 2.7557319223985893e-06);           // 1/9!
  sp = fma(sp, r2, -1.9841269841269841e-04);          // -1/7!
  sp = fma(sp, r2, 8.3333333333333332e-03);           // 1/5!
  sp = fma(sp, r2, -1.6666666666666666e-01);          // -1/3!
  const double sr = fma(sp * r2, r, r);
  double cp = -1.1470745597729725e-11;                // -1/14!
  cp = fma(cp, r2, 2.0876756987868100e-09);           // 1/12!
  cp = fma(cp, r2, -2.7557319223985888e-07);          // -1/10!
  cp = fma(cp, r2, 2.4801587301587302e-05);           // 1/8!
  cp = fma(cp, r2, -1.3888888888888889e-03);          // -1/6!
  cp = fma(cp, r2, 4.1666666666666664e-02);           // 1/4!
  cp = fma(cp, r2, -0.5);
  const double cr = fma(cp, r2, 1.0);
  const int iq = (int)q & 3;
  const double ss = (iq & 1) ? cr : sr;
  const double cc = (iq & 1) ? sr : cr;
  *s = (iq & 2) ? -ss : ss;
  *c = ((iq + 1) & 2) ? -cc : cc;
}

// ---- SE(3) element: x -> R x + p (column vectors), R row-major ---------------------------------------
struct Se3 {
  double r[9];
  double p[3];
};

__device__ __forceinline__ void se3_identity(Se3& a) {
#pragma unroll
  for (int i = 0; i < 9; i++) a.r[i] = (i % 4 == 0) ? 1.0 : 0.0;
  a.p[0] = a.p[1] = a.p[2] = 0.0;
}

// out = a o b  (apply b first)
__device__ __forceinline__ Se3 se3_mul(const Se3& a, const Se3& b) {
  Se3 o;
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) o.r[3 * i + j] = a.r[3 * i] * b.r[j] + a.r[3 * i + 1] * b.r[3 + j] + a.r[3 * i + 2] * b.r[6 + j];
    o.p[i] = a.r[3 * i] * b.p[0] + a.r[3 * i + 1] * b.p[1] + a.r[3 * i + 2] * b.p[2] + a.p[i];
  }
  return o;
}

__device__ __forceinline__ Se3 se3_shfl_up(const Se3& a, int delta, int width) {
  Se3 o;
#pragma unroll
  for (int i = 0; i < 9; i++) o.r[i] = __shfl_up_sync(0xffffffffu, a.r[i], delta, width);
#pragma unroll
  for (int i = 0; i < 3; i++) o.p[i] = __shfl_up_sync(0xffffffffu, a.p[i], delta, width);
  return o;
}

// same, but only rows 0 and 1 of the rotation travel: row 2 = row 0 x row 1 (9 FP64 operations instead of
// 6 32-bit shuffles -- shuffles share the L1 data pipe with shared memory, which bounds the forward kernel)
__device__ __forceinline__ Se3 se3_shfl_up_rot(const Se3& a, int delta) {
  Se3 o;
#pragma unroll
  for (int i = 0; i < 6; i++) o.r[i] = __shfl_up_sync(0xffffffffu, a.r[i], delta);
#pragma unroll
  for (int i = 0; i < 3; i++) o.p[i] = __shfl_up_sync(0xffffffffu, a.p[i], delta);
  o.r[6] = o.r[1] * o.r[5] - o.r[2] * o.r[4];
  o.r[7] = o.r[2] * o.r[3] - o.r[0] * o.r[5];
  o.r[8] = o.r[0] * o.r[4] - o.r[1] * o.r[3];
  return o;
}

// exclusive scan over `width` consecutive lanes (width = 16 or 32): lane q gets T_0 o ... o T_{q-1}
__device__ __forceinline__ Se3 se3_exclusive_scan(Se3 t, int lane_in_group, int width) {
  for (int d = 1; d < width; d <<= 1) {
    Se3 up = se3_shfl_up(t, d, width);
    if (lane_in_group >= d) t = se3_mul(up, t);
  }
  Se3 ex = se3_shfl_up(t, 1, width);
  if (lane_in_group == 0) se3_identity(ex);
  return ex;
}

// NeRF step: R <- R Rx(phi) Rz(g), p <- p + L R e_x     (cw,sw = cos/sin phi; cg,sg = cos/sin g)
__device__ __forceinline__ void nerf_step(Se3& f, double cw, double sw, double cg, double sg, double L) {
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double x = f.r[3 * i], y = f.r[3 * i + 1], z = f.r[3 * i + 2];
    const double y1 = y * cw + z * sw;   // column y of R Rx
    const double z1 = z * cw - y * sw;   // column z of R Rx
    const double x2 = x * cg + y1 * sg;  // column x of (R Rx) Rz
    const double y2 = y1 * cg - x * sg;
    f.r[3 * i] = x2;
    f.r[3 * i + 1] = y2;
    f.r[3 * i + 2] = z1;
    f.p[i] = fma(L, x2, f.p[i]);
  }
}

// ---- SE(2) element for the planar chain: unit complex rotation + displacement ----------------------
struct Se2 {
  double c, s, x, y;
};
__device__ __forceinline__ Se2 se2_mul(const Se2& a, const Se2& b) {  // a o b (b first, in a's frame)
  Se2 o;
  o.c = a.c * b.c - a.s * b.s;
  o.s = a.c * b.s + a.s * b.c;
  o.x = a.x + a.c * b.x - a.s * b.y;
  o.y = a.y + a.s * b.x + a.c * b.y;
  return o;
}
__device__ __forceinline__ Se2 se2_shfl_up(const Se2& a, int d) {
  Se2 o;
  o.c = __shfl_up_sync(0xffffffffu, a.c, d);
  o.s = __shfl_up_sync(0xffffffffu, a.s, d);
  o.x = __shfl_up_sync(0xffffffffu, a.x, d);
  o.y = __shfl_up_sync(0xffffffffu, a.y, d);
  return o;
}
__device__ __forceinline__ Se2 se2_shfl_down(const Se2& a, int d) {
  Se2 o;
  o.c = __shfl_down_sync(0xffffffffu, a.c, d);
  o.s = __shfl_down_sync(0xffffffffu, a.s, d);
  o.x = __shfl_down_sync(0xffffffffu, a.x, d);
  o.y = __shfl_down_sync(0xffffffffu, a.y, d);
  return o;
}
// inclusive scan over the 32 lanes; *excl receives the exclusive prefix
__device__ __forceinline__ Se2 se2_scan(Se2 t, int lane, Se2* excl) {
  for (int d = 1; d < 32; d <<= 1) {
    Se2 up = se2_shfl_up(t, d);
    if (lane >= d) t = se2_mul(up, t);
  }
  Se2 ex = se2_shfl_up(t, 1);
  if (lane == 0) ex = Se2{1.0, 0.0, 0.0, 0.0};
  *excl = ex;
  return t;
}

// planar bond k: position advances by L_k along the current direction, then the direction turns by
// -(-1)^k (pi - theta_k)  (encodermap_tf1/backmapping.py:105-110; for the last bond there is no turn)
__device__ __forceinline__ void planar_step(Se2& t, double L, bool has_turn, float theta, int k) {
  t.x = fma(L, t.c, t.x);
  t.y = fma(L, t.s, t.y);
  if (has_turn) {
    double st, ct;
    sincos_d((double)theta, &st, &ct);
    // turn angle w = -(-1)^k (pi - theta): cos w = -cos(theta), sin w = -(-1)^k sin(theta)
    const double cw = -ct, sw = (k & 1) ? st : -st;
    const double c2 = t.c * cw - t.s * sw, s2 = t.c * sw + t.s * cw;
    t.c = c2;
    t.s = s2;
  }
}

// ---- asynchronous, vectorised staging of one contiguous row into shared memory -------------------------
__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
// coalesced copy of `len` floats from shared to global; sm and dst are congruent modulo 16 bytes
__device__ __forceinline__ void store_row(float* __restrict__ dst, const float* sm, int len, int t, int nthreads) {
  const int head = min(len, (int)(((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15) >> 2));
  for (int e = t; e < head; e += nthreads) dst[e] = sm[e];
  const int nvec = (len - head) >> 2;
  const float4* s4 = reinterpret_cast<const float4*>(sm + head);
  float4* d4 = reinterpret_cast<float4*>(dst + head);
  for (int v = t; v < nvec; v += nthreads) d4[v] = s4[v];
  for (int e = head + (nvec << 2) + t; e < len; e += nthreads) dst[e] = sm[e];
}

__device__ __forceinline__ void stage_in(float* dst, const float* __restrict__ src, int count, int lane) {
  for (int i = lane; i < count; i += 32) dst[i] = src[i];
}

// ====================================================================================================
// sin/cos of a float32 argument in float64, by table: float32 Cody-Waite reduction to |r| <= pi/1024 (the first two
// steps are exact), float32 two-term corrections, a float64 table entry and 4 DFMA -- 2.3e-10 absolute error with 4
// FP64 instructions instead of ~30.  The 1024-entry table lives in global memory (backward kernels); the forward
// kernel copies every fourth entry into shared memory (sincos_tab256 below).
// ====================================================================================================
constexpr int SC_TABLE = 1024;               // table step 2 pi / 1024
__global__ void sincos_table_kernel(double2* tab) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < SC_TABLE) {
    double s, c;
    sincospi(2.0 * (double)i / (double)SC_TABLE, &s, &c);
    tab[i] = make_double2(s, c);
  }
}

__device__ __forceinline__ void sincos_tab(float x, const double2* __restrict__ tab, double* s, double* c) {
  if (fabsf(x) >= 48.f) {   // never produced by the models, allowed by the reference: full float64 reduction
    sincos_d((double)x, s, c);
    return;
  }
  // k = rint(x * 1024 / 2pi) by the magic-number trick; the low mantissa bits of (x*A + M) are k mod 2^22
  const float km = fmaf(x, 162.9746551513672f, 12582912.f);
  const float kf = km - 12582912.f;
  // 2pi/1024 = C1 + C2 + C3 with 11-bit C1, C2: the first two reductions are exact in float32
  float r = fmaf(-kf, 0.006134033203125f, x);
  r = fmaf(-kf, 1.889653503894806e-06f, r);
  r = fmaf(-kf, 2.949136768126692e-10f, r);
  const float r2 = r * r;
  const float sr = fmaf(r * r2, -0.16666667f, r);                  // sin r
  const float cm = r2 * fmaf(r2, 0.041666668f, -0.5f);             // cos r - 1
  const double2 t = __ldg(tab + (__float_as_int(km) & (SC_TABLE - 1)));
  *s = fma(t.x, (double)cm, fma(t.y, (double)sr, t.x));
  *c = fma(t.y, (double)cm, fma(-t.x, (double)sr, t.y));
}

// planar bond k with the table sin/cos (same step as planar_step)
__device__ __forceinline__ void planar_step_tab(Se2& t, double L, bool has_turn, float theta, int k, const double2* __restrict__ tab) {
  t.x = fma(L, t.c, t.x);
  t.y = fma(L, t.s, t.y);
  if (has_turn) {
    double st, ct;
    sincos_tab(theta, tab, &st, &ct);
    const double cw = -ct, sw = (k & 1) ? st : -st;
    const double c2 = t.c * cw - t.s * sw, s2 = t.c * sw + t.s * cw;
    t.c = c2;
    t.s = s2;
  }
}

// ====================================================================================================
// chain_in_plane forward
// ====================================================================================================
__global__ void chain_in_plane_kernel(const float* __restrict__ lengths, int64_t lstride, const float* __restrict__ angles,
                                      int64_t b, int n, float* __restrict__ xyz, const double2* __restrict__ tab) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const int64_t frame = (int64_t)blockIdx.x * wpc + warp;
  if (frame >= b) return;
  const int per_warp = (n - 1) + (n - 2) + 3 * n;
  float* sL = smem + (size_t)warp * per_warp;
  float* sA = sL + (n - 1);
  float* sO = sA + (n - 2);
  stage_in(sL, lengths + frame * lstride, n - 1, lane);
  stage_in(sA, angles + frame * (int64_t)(n - 2), n - 2, lane);
  __syncwarp();
  const int nb = n - 1;                 // bonds
  const int cs = (nb + 31) / 32;
  const int k0 = lane * cs, k1 = min(nb, k0 + cs);
  Se2 t{1.0, 0.0, 0.0, 0.0};
  for (int k = k0; k < k1; k++) planar_step_tab(t, (double)sL[k], k < n - 2, k < n - 2 ? sA[k] : 0.f, k, tab);
  Se2 ex;
  se2_scan(t, lane, &ex);
  Se2 run = ex;
  if (lane == 0) {
    sO[0] = 0.f;
    sO[1] = 0.f;
    sO[2] = 0.f;
  }
  for (int k = k0; k < k1; k++) {
    Se2 step{1.0, 0.0, 0.0, 0.0};
    planar_step_tab(step, (double)sL[k], k < n - 2, k < n - 2 ? sA[k] : 0.f, k, tab);
    run = se2_mul(run, step);
    sO[3 * (k + 1)] = (float)run.x;
    sO[3 * (k + 1) + 1] = (float)run.y;
    sO[3 * (k + 1) + 2] = 0.f;
  }
  __syncwarp();
  float* dst = xyz + frame * (int64_t)(3 * n);
  for (int i = lane; i < 3 * n; i += 32) dst[i] = sO[i];
}

// ====================================================================================================
// BackMapLayer forward, version 5: same mathematics as version 3, half the shared memory, ~2/3 of the
// instructions per NeRF step and a shared-memory sin/cos table, so that 24 instead of 12 warps are resident
// per SM (the kernel is bound by dependent FP64 latencies and by L1 data-pipe wavefronts, not by a math pipe).
//   * persistent CTAs of 6 independent 64-thread groups (one warp per side of the chain); a group loops over
//     frames and synchronises with its own named barrier, so a slow frame never holds up its neighbours.
//   * ONE shared-memory region of 3n floats per group.  The inputs of atom k -- (dihedral, angle, length) of the
//     step that places it -- are staged INTERLEAVED at T[3k .. 3k+2] (4-byte cp.async, stride-3 scatter, bank
//     conflict free), pass 1 overwrites each triple with the chunk-local position of the atom it placed, pass 2
//     transforms it in place, and the region leaves as one contiguous, vectorised row store.
//     slot of dihedral j: j (left of the anchor, j < s-1) or j+3;  angle j: j (j < s) or j+2;  length j: j
//     (j < s+1) or j+1 -- the middle angle theta_{s-1} and the two middle bond lengths land in the slots of
//     the three anchor atoms, which the anchor lane reads before it writes their planar positions.
//   * sin/cos: 256-entry float64 table in shared memory (v3 read a 1024-entry table through L1, which the
//     shared-memory carve-out had shrunk to ~30 KB: 51 % hit rate, 30 sectors per warp-wide lookup).
//   * per-lane trip counts (no dummy steps, no selects); arguments beyond the table path's range only raise a
//     flag, and a flagged warp re-stages its side and repeats pass 1 with the float64 polynomial sin/cos
//     (never taken by model outputs).
// ====================================================================================================
constexpr int FWD5_GROUPS = 6;
constexpr int FWD5_THREADS = 64 * FWD5_GROUPS;
constexpr int SC_SMALL = 256;                 // shared-memory table: step 2 pi / 256
constexpr float SC_FAST_LIMIT = 48.f;

__device__ __forceinline__ void group_barrier(int g) { asm volatile("bar.sync %0, 64;" ::"r"(g + 1) : "memory"); }

// both table sin/cos of one NeRF step at once (|x| < SC_FAST_LIMIT is the caller's business; tabS in shared memory): the float32 argument reduction and the two short polynomials run as
// packed FP32x2 instructions (FFMA2 / FMUL2 / FADD2), one lane of the pair per angle -- 10 issue slots instead of 20
__device__ __forceinline__ void sincos_tab256_x2(float xa, float xb, const double2* tabS, double* sa, double* ca, double* sb, double* cb) {
  const float2 x = make_float2(xa, xb);
  const float2 km = __ffma2_rn(x, make_float2(40.74366543152521f, 40.74366543152521f), make_float2(12582912.f, 12582912.f));
  const float2 kf = __fadd2_rn(km, make_float2(-12582912.f, -12582912.f));
  float2 r = __ffma2_rn(kf, make_float2(-0.0245361328125f, -0.0245361328125f), x);
  r = __ffma2_rn(kf, make_float2(-7.558614015579224e-06f, -7.558614015579224e-06f), r);
  r = __ffma2_rn(kf, make_float2(-1.1796547072506768e-09f, -1.1796547072506768e-09f), r);
  const float2 r2 = __fmul2_rn(r, r);
  const float2 sr = __ffma2_rn(__fmul2_rn(r, r2), make_float2(-0.16666667f, -0.16666667f), r);
  const float2 cm = __fmul2_rn(r2, __ffma2_rn(r2, make_float2(0.041666668f, 0.041666668f), make_float2(-0.5f, -0.5f)));
  const double2 ta = tabS[__float_as_int(km.x) & (SC_SMALL - 1)];
  const double2 tb = tabS[__float_as_int(km.y) & (SC_SMALL - 1)];
  *sa = fma(ta.x, (double)cm.x, fma(ta.y, (double)sr.x, ta.x));
  *ca = fma(ta.y, (double)cm.x, fma(-ta.x, (double)sr.x, ta.y));
  *sb = fma(tb.x, (double)cm.y, fma(tb.y, (double)sr.y, tb.x));
  *cb = fma(tb.y, (double)cm.y, fma(-tb.x, (double)sr.y, tb.y));
}

template <bool LEFT, bool SLOW>
__device__ __forceinline__ void fwd5_step(float* pT, int& kpar, const double2* tabS, Se3& f, Se2& pl, float& amax) {
  const float fd = pT[0], fa = pT[1], fl = pT[2];
  double sw, cw, sg, cg;
  if (SLOW) {
    sincos_d((double)fd, &sw, &cw);
    sincos_d((double)fa, &sg, &cg);
  } else {
    amax = fmaxf(amax, fmaxf(fabsf(fd), fabsf(fa)));
    sincos_tab256_x2(fd, fa, tabS, &sw, &cw, &sg, &cg);
  }
  const double L = (double)fl;
  nerf_step(f, cw, sw, -cg, sg, L);   // bend by pi - theta: cos = -cos(theta), sin = sin(theta)
  pT[0] = (float)f.p[0];
  pT[1] = (float)f.p[1];
  pT[2] = (float)f.p[2];
  if (LEFT) {
    // planar bond k: advance L along the current direction, then turn by -(-1)^k (pi - theta_k); this lane walks
    // k downwards, so the step is prepended: pl <- step_k o pl
    const double tc = -cg, ts = (kpar & 1) ? sg : -sg;
    const double nc = tc * pl.c - ts * pl.s, ns = tc * pl.s + ts * pl.c;
    const double nx = L + tc * pl.x - ts * pl.y, ny = ts * pl.x + tc * pl.y;
    pl.c = nc; pl.s = ns; pl.x = nx; pl.y = ny;
    kpar--;
  }
}

// pass 1 of one lane: `nvalid` steps walking the interleaved region from pT (LEFT: downwards).  Returns false if
// an argument was outside the table path's range (NaN compares false and propagates through the table path).
template <bool LEFT, bool SLOW>
__device__ __forceinline__ bool fwd5_pass1(float* pT, int nvalid, int kpar, const double2* tabS, Se3& f, Se2& pl) {
  constexpr int dk = LEFT ? -3 : 3;
  float amax = 0.f;
  int c = 0;
#pragma unroll 1
  for (; c + 1 < nvalid; c += 2) {
    fwd5_step<LEFT, SLOW>(pT, kpar, tabS, f, pl, amax);
    fwd5_step<LEFT, SLOW>(pT + dk, kpar, tabS, f, pl, amax);
    pT += 2 * dk;
  }
  if (c < nvalid) fwd5_step<LEFT, SLOW>(pT, kpar, tabS, f, pl, amax);
  return amax < SC_FAST_LIMIT;
}

// interleaving stage of row elements [j0, j1) of `src` into T[3 (j + shift) + COMP] by the 64 threads of a group:
// thread t copies j = j0 + t + 64 i.  Uniform trip count, 4 predicated copies with immediate offsets per iteration
// (about 3 instructions per element; the plain strided loop needed 7).
__device__ __forceinline__ void cp_async4_s(uint32_t dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
template <int COMP>
__device__ __forceinline__ void stage_interleaved(uint32_t sT, const float* __restrict__ src, int j0, int j1, int shift, int t) {
  uint32_t d = sT + 4u * (uint32_t)(3 * (j0 + t + shift) + COMP);
  const float* s = src + j0 + t;
  int left = j1 - j0 - t;                  // copy i of this thread is in range iff 64 i < left
  const int iters = (j1 - j0 + 63) >> 6;   // uniform over the group
  int i = 0;
#pragma unroll 1
  for (; i + 4 <= iters; i += 4) {
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (64 * u < left) cp_async4_s(d + 768u * u, s + 64 * u);
    d += 4 * 768;
    s += 256;
    left -= 256;
  }
#pragma unroll
  for (int u = 0; u < 3; u++)
    if (i + u < iters && 64 * u < left) cp_async4_s(d + 768u * u, s + 64 * u);
}

__global__ void __launch_bounds__(FWD5_THREADS, 2) backmap_fwd5_kernel(const float* __restrict__ lengths, int64_t lstride,
                                                                       const float* __restrict__ angles,
                                                                       const float* __restrict__ dihedrals, int64_t b, int n,
                                                                       float* __restrict__ xyz, const double2* __restrict__ tab) {
  extern __shared__ __align__(16) float smem[];
  __shared__ double anchor[FWD5_GROUPS][2][12];

  const int tid = threadIdx.x, lane = tid & 31;
  const int g = tid >> 6, t = tid & 63, side = (tid >> 5) & 1;
  double2* tabS = reinterpret_cast<double2*>(smem);
  const int ngroups = blockDim.x >> 6;   // 6 at <= 3000 atoms, fewer for longer chains (shared memory)
  for (int i = tid; i < SC_SMALL; i += blockDim.x) tabS[i] = __ldg(tab + i * (SC_TABLE / SC_SMALL));
  __syncthreads();

  const int out_floats = 3 * n;
  const int rO = (out_floats + 7) & ~3;
  float* region = smem + SC_SMALL * 4 + (size_t)g * rO;
  const int s = n / 2;
  const int steps = side == 0 ? (s - 1) : (n - s - 2);
  const int ch = ((max(s - 1, n - s - 2) + 31) / 32) | 1;   // odd chunk length: conflict-free strided smem access
  const int i0 = lane * ch;
  const int nvalid = max(0, min(ch, steps - i0));
  const int kfirst = side == 0 ? s - 2 - i0 : s + 2 + i0;

  for (int64_t frame = (int64_t)blockIdx.x * ngroups + g; frame < b; frame += (int64_t)gridDim.x * ngroups) {
    float* dst = xyz + frame * (int64_t)out_floats;
    // congruent with the global row modulo 16 bytes, so that the final copy is float4 on both sides
    float* T = region + (int)((reinterpret_cast<uintptr_t>(dst) & 15) >> 2);
    const float* rowD = dihedrals + frame * (int64_t)(n - 3);
    const float* rowA = angles + frame * (int64_t)(n - 2);
    const float* rowL = lengths + frame * lstride;
    const uint32_t sT = (uint32_t)__cvta_generic_to_shared(T);
    stage_interleaved<0>(sT, rowD, 0, s - 1, 0, t);
    stage_interleaved<0>(sT, rowD, s - 1, n - 3, 3, t);
    stage_interleaved<1>(sT, rowA, 0, s, 0, t);
    stage_interleaved<1>(sT, rowA, s, n - 2, 2, t);
    stage_interleaved<2>(sT, rowL, 0, s + 1, 0, t);
    stage_interleaved<2>(sT, rowL, s + 1, n - 1, 1, t);
    cp_async_wait_all();
    group_barrier(g);

    float* pT = T + 3 * kfirst;
    Se3 f;
    se3_identity(f);
    Se2 pl{1.0, 0.0, 0.0, 0.0};
    bool fast_ok;
    if (side == 0) fast_ok = fwd5_pass1<true, false>(pT, nvalid, kfirst, tabS, f, pl);
    else fast_ok = fwd5_pass1<false, false>(pT, nvalid, kfirst, tabS, f, pl);
    if (!__all_sync(0xffffffffu, fast_ok)) {
      // out-of-range argument somewhere on this side: restore the side's inputs and repeat with float64 sin/cos
      if (side == 0) {
        for (int j = lane; j < s - 1; j += 32) { T[3 * j] = rowD[j]; T[3 * j + 1] = rowA[j]; T[3 * j + 2] = rowL[j]; }
      } else {
        for (int k = s + 2 + lane; k < n; k += 32) { T[3 * k] = rowD[k - 3]; T[3 * k + 1] = rowA[k - 2]; T[3 * k + 2] = rowL[k - 1]; }
      }
      __syncwarp();
      se3_identity(f);
      pl = Se2{1.0, 0.0, 0.0, 0.0};
      if (side == 0) fwd5_pass1<true, true>(pT, nvalid, kfirst, tabS, f, pl);
      else fwd5_pass1<false, true>(pT, nvalid, kfirst, tabS, f, pl);
    }

    // ---- scan of the chunk aggregates inside the warp ---------------------------------------------------
    Se3 inc = f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      Se3 up = se3_shfl_up_rot(inc, d);
      if (lane >= d) inc = se3_mul(up, inc);
    }
    Se3 ex = se3_shfl_up_rot(inc, 1);
    if (lane == 0) se3_identity(ex);

    // ---- anchor: the left warp reduces the planar product (higher lanes hold lower bonds => go on the left)
    if (side == 0) {
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        Se2 o = se2_shfl_down(pl, d);
        if (lane + d < 32) pl = se2_mul(o, pl);
      }
      if (lane == 0) {
        // pl: direction of bond s-1 and position of atom s-1 in the plane
        const double dm_c = pl.c, dm_s = pl.s, am_x = pl.x, am_y = pl.y;
        const float th_mid = T[3 * (s - 1) + 1], l_m = T[3 * (s - 1) + 2], l_p = T[3 * s + 2];
        Se2 mid = pl;
        planar_step(mid, (double)l_m, true, th_mid, s - 1);
        const double a0_x = mid.x, a0_y = mid.y, dp_c = mid.c, dp_s = mid.s;
        const double ap_x = fma((double)l_p, dp_c, a0_x), ap_y = fma((double)l_p, dp_s, a0_y);
        const double zs = (dm_c * dp_s - dm_s * dp_c) >= 0.0 ? 1.0 : -1.0;
        T[3 * (s - 1)] = (float)am_x; T[3 * (s - 1) + 1] = (float)am_y; T[3 * (s - 1) + 2] = 0.f;
        T[3 * s] = (float)a0_x;       T[3 * s + 1] = (float)a0_y;       T[3 * s + 2] = 0.f;
        T[3 * (s + 1)] = (float)ap_x; T[3 * (s + 1) + 1] = (float)ap_y; T[3 * (s + 1) + 2] = 0.f;
        // anchor frames: x along the last anchored bond, z = +-e_z, y = z x x, origin at the last anchor atom
#pragma unroll
        for (int sd = 0; sd < 2; sd++) {
          const double xx = sd == 0 ? -dm_c : dp_c, xy = sd == 0 ? -dm_s : dp_s, zz = sd == 0 ? -zs : zs;
          double* a = anchor[g][sd];
          a[0] = xx; a[1] = -zz * xy; a[2] = 0.0;
          a[3] = xy; a[4] = zz * xx;  a[5] = 0.0;
          a[6] = 0.0; a[7] = 0.0;     a[8] = zz;
          a[9] = sd == 0 ? am_x : ap_x; a[10] = sd == 0 ? am_y : ap_y; a[11] = 0.0;
        }
      }
    }
    group_barrier(g);

    // ---- pass 2: out = (anchor o prefix)(local), float32 rotation (chunk-local coordinates are < 4 nm), translation
    //      as a float32 hi/lo pair so that the only full-magnitude rounding is the one of the float32 output
    if (nvalid > 0) {
      Se3 carry;
#pragma unroll
      for (int i = 0; i < 9; i++) carry.r[i] = anchor[g][side][i];
#pragma unroll
      for (int i = 0; i < 3; i++) carry.p[i] = anchor[g][side][9 + i];
      const Se3 pre = se3_mul(carry, ex);
      float rf[9], ph[3], plo[3];
#pragma unroll
      for (int i = 0; i < 9; i++) rf[i] = (float)pre.r[i];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        ph[i] = (float)pre.p[i];
        plo[i] = (float)(pre.p[i] - (double)ph[i]);
      }
      float* pO = pT;
      const int dk = side == 0 ? -3 : 3;
#pragma unroll 4
      for (int c = 0; c < nvalid; c++) {
        const float lx = pO[0], ly = pO[1], lz = pO[2];
        pO[0] = ph[0] + fmaf(rf[0], lx, fmaf(rf[1], ly, fmaf(rf[2], lz, plo[0])));
        pO[1] = ph[1] + fmaf(rf[3], lx, fmaf(rf[4], ly, fmaf(rf[5], lz, plo[1])));
        pO[2] = ph[2] + fmaf(rf[6], lx, fmaf(rf[7], ly, fmaf(rf[8], lz, plo[2])));
        pO += dk;
      }
    }
    group_barrier(g);
    store_row(dst, T, out_floats, t, 64);
    group_barrier(g);   // the next frame's staging overwrites T
  }
}

// ====================================================================================================
// BackMapLayer forward, version 6 (large batches): one LANE per (frame, side) -- the chain is walked sequentially in
// GLOBAL coordinates, 32 frames per warp, and the parallelism comes from the frames instead of from cutting one chain
// into 32 chunks.  What disappears against version 5: the SE(3) scan (21 % of the FP64 work, 26 % of the L1 wavefronts),
// pass 2 and its shared-memory round trip, the chunk-local / hi-lo bookkeeping, every barrier inside a frame:
// 6.2 k instead of 8.0 k warp instructions per 500-residue frame.
//   * work item = (tile of 32 consecutive frames, side); the two warps of a tile are neighbours in the CTA and split the
//     planar prologue: the anchor frame needs the SE(2) product of the bonds LEFT of the middle before either side can
//     start in global coordinates; each warp multiplies half of those bonds (angles only: 10 FP64 per bond), they swap
//     the two partial products through shared memory (one named barrier per tile) and compose them.
//   * inputs: rows of `dihedrals` / `angles` are streamed through per-warp shared-memory tiles of 8 steps x 32 frames,
//     dense [frame][8] with the two 16-byte halves of every other group of four rows swapped (the per-lane LDS.128 of a
//     half -- lane = frame -- is then conflict free), filled by cp.async (4 bytes per lane for the dihedrals, whose row pitch
//     n - 3 is odd; 8 bytes per lane for the angles, whose rows are 8-byte aligned when n % 4 == 0) with the L2::128B
//     prefetch hint, double buffered: the tile of the next 8 steps is in flight while the current one is consumed.  Neither
//     array can use the bulk-copy engine: its tensor maps need row pitches that are multiples of 16 bytes.
//     Bond lengths (shared by all frames) sit in shared memory as float64.
//   * outputs: a lane keeps the float32 positions of the 8 atoms of a group in registers, writes them as 6 STS.128 into its
//     96-byte row of the warp's group buffer, and ONE 2-d TMA tensor store per group and warp (box 24 floats x 32 frames of
//     xyz viewed as (b, 3 n); 16-byte aligned because n % 4 == 0; rows beyond the batch are clipped by the unit) sends the
//     buffer off: one instruction per 8 atoms x 32 frames and no LSU wavefronts.  The buffer is rewritten a full group later
//     (waiting for the store's shared-memory read any earlier cost 20 %).  Atoms outside whole 8-groups of a side
//     (<= 14 per frame) are stored directly.
//   * sin/cos: 128-entry float64 table replicated 8x (lane l reads replica l & 7: 4 wavefronts per warp-wide lookup), float32
//     Cody-Waite reduction and degree-3/4 corrections with literal constants; angles beyond +-384 rad are only noticed
//     (max |angle| per lane) and make the pair of warps repeat the tile with every step range-tested.
// Measured (profiles/r02_*): bound by issue slots with FP64 instructions counting twice -- (2 x 35 FP64 + 48 other) / 4
// schedulers = 29.6 cycles per warp-step is the measured compute-only rate (no copies in or out: 137 M frames/s = 0.63 of the
// HBM roofline at 500 residues); copies in add 21 %, stores 15 %.  The per-SM rate is the same for 12 .. 16 warps.
// Used when the batch fills whole waves of (frame tile, side) warps; version 5 stays the path for small batches,
// n % 4 != 0, per-frame bond lengths and very long chains.
// ====================================================================================================
// one persistent CTA per SM with 8 .. 20 warps (4 .. 10 tile pairs): the per-SM rate is the same from 12 to 20 warps (it is set
// by issue slots, FP64 instructions counting twice), so the count is chosen per launch to make the tile count come out as
// whole waves; emk_set_option("backmap_fwd6_warps", w) forces one (0 = automatic)
constexpr int F6_TILE = 8;                         // steps per input tile = atoms per output group
constexpr int F6_IN_TILE = 32 * F6_TILE;           // words per (array, buffer): [frame][8 elements], dense
constexpr int F6_OUT_TILE = 32 * 24;               // output group [frame][8 atoms x 3], dense: the box of the TMA tensor store
constexpr int F6_WARP_WORDS = 4 * F6_IN_TILE + F6_OUT_TILE;   // D/A x two buffers + output group = 7 KB
constexpr int SC6 = 128;                           // table entries (step 2 pi / 128), 8 replicas: 16 KB

// The float32 Cody-Waite reduction below is exact while k * C1 fits 24 bits: C1 = 2 pi / 128 rounded to 11 bits, |k| < 2^13,
// i.e. |x| < 402 rad.  Larger finite angles (legal for the reference, never produced by the models) take the float64 path;
// NaN / Inf pass through the table path as NaN on their own.
constexpr float SC6_LIMIT = 384.f;

// table sin/cos of one angle (|x| < SC6_LIMIT): tabB = byte address of THIS LANE's replica of the 128-entry float64 (sin, cos)
// table in shared memory (entry stride 8 x 16 bytes: lane l reads replica l & 7, so a warp-wide 16-byte lookup with random
// indices is 4 wavefronts instead of ~8 -- the shared-memory pipe is the busiest unit of this kernel next to the issue slots).
// |r| <= pi / 128: sin r = r - r^3 / 6 (+7e-11) and cos r - 1 = -r^2 / 2 + r^4 / 24 (+3e-13); the float32 evaluation adds
// 1.5e-9.  Scalar FFMA with literal constants: the packed FFMA2 form needs every constant in a register PAIR.
__device__ __forceinline__ void sincos_tab64(float x, uint32_t tabB, double* sn, double* cs) {
  const float km = fmaf(x, 20.371832715762604f, 12582912.f);
  const float kf = km - 12582912.f;
  float r = fmaf(kf, -0.0490722656250f, x);
  r = fmaf(kf, -1.5117228031158447e-05f, r);
  r = fmaf(kf, -2.3593094145013538e-09f, r);
  const float r2 = r * r;
  const float sr = fmaf(r * r2, -0.16666667f, r);                       // sin r
  const float cm = r2 * fmaf(r2, 0.041666668f, -0.5f);                  // cos r - 1
  double tx, ty;
  asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(tx), "=d"(ty) : "r"(tabB + (((uint32_t)__float_as_int(km) & (SC6 - 1)) << 7)));
  *sn = fma(tx, (double)cm, fma(ty, (double)sr, tx));
  *cs = fma(ty, (double)cm, fma(-tx, (double)sr, ty));
}
__device__ __forceinline__ void sincos_tab64_x2(float xa, float xb, uint32_t tabB, double* sa, double* ca, double* sb, double* cb) {
  sincos_tab64(xa, tabB, sa, ca);
  sincos_tab64(xb, tabB, sb, cb);
}
// out of line on purpose: the step body is instantiated ~20 times and must stay inside the instruction cache
static __device__ __noinline__ double2 sincos_slow(float x) {   // by value: pointer outputs would pin the caller's results to local memory
  double s, c;
  sincos_d((double)x, &s, &c);
  return make_double2(s, c);
}
__device__ __forceinline__ void sincos_f6_x2(float xa, float xb, uint32_t tabB, double* sa, double* ca, double* sb, double* cb) {
  if (!(fmaxf(fabsf(xa), fabsf(xb)) >= SC6_LIMIT)) {   // NaN takes the table path too (and comes out as NaN)
    sincos_tab64_x2(xa, xb, tabB, sa, ca, sb, cb);
  } else {   // beyond the exact range of the float32 reduction (never produced by the models), or NaN
    const double2 a = sincos_slow(xa), bb = sincos_slow(xb);
    *sa = a.x; *ca = a.y; *sb = bb.x; *cb = bb.y;
  }
}

// ---- float32 chain (first pass of version 6, round 2) ---------------------------------------------------------------
// The float64 chain is bound by issue slots with FP64 instructions counting twice (118 issue cycles per warp-step).  A
// float32 chain in coordinates RELATIVE TO THE SIDE'S ANCHOR ATOM needs ~50: rotation columns as FFMA2 pairs (rows 0 and
// 1 packed, row 2 scalar), correctly rounded float32 sin/cos from a hi/lo table, no conversions.  Its error is a random walk
// of the orientation: measured <= 4e-6 x (extent of the side) over 4 096 chains of 750 steps (random coils, helices,
// extended chains; tools/experiments/f32_chain_error.py), i.e. 2.6e-5 nm for a 16 nm side against the 1e-4 nm tolerance.
// Every lane therefore tracks the extent of its side (max |p - anchor|^2 at the group ends); if any lane of a tile pair
// exceeds the limit (emk_set_option("backmap_fwd6_f32_extent_nm"): 16 keeps the error below 7e-5 nm), the pair repeats the
// tile with the float64 chain -- the same second pass that catches angles beyond the table range.  The planar prologue
// (anchor frame) stays float64: its error would rotate the whole side.  OPT-IN: measured on B200 it does not pay (see
// g_fwd6_f32_extent below); kept as the tested answer to "would a float32 chain be faster".
struct Se3f {
  float2 x01, y01, z01, p01;   // rows 0 and 1 of the columns x, y, z of R, and of the position
  float x2, y2, z2, p2;        // row 2
};
__device__ __forceinline__ float2 bc2(float v) { return make_float2(v, v); }

// sin / cos of two angles, correctly rounded to float32 up to ~1e-9: entry {s_hi, c_hi, s_lo, c_lo} of the 128-entry table
// (8 replicas, lane l reads replica l & 7), sin(a + r) = s_hi + (s_lo + s_hi (cos r - 1) + c_hi sin r) and the like for cos;
// the bracket is below 0.025 in magnitude, so its float32 rounding is 1.5e-9.  |x| < SC6_LIMIT as for the float64 form.
__device__ __forceinline__ void sincos_tab32_x2(float xa, float xb, uint32_t tabF, float& sa, float& ca, float& sb, float& cb) {
  const float2 x = make_float2(xa, xb);
  const float2 km = __ffma2_rn(x, bc2(20.371832715762604f), bc2(12582912.f));
  const float2 kf = __fadd2_rn(km, bc2(-12582912.f));
  float2 r = __ffma2_rn(kf, bc2(-0.0490722656250f), x);
  r = __ffma2_rn(kf, bc2(-1.5117228031158447e-05f), r);
  r = __ffma2_rn(kf, bc2(-2.3593094145013538e-09f), r);
  const float2 r2 = __fmul2_rn(r, r);
  const float2 sr = __ffma2_rn(__fmul2_rn(r, r2), bc2(-0.16666667f), r);                 // sin r
  const float2 cm = __fmul2_rn(r2, __ffma2_rn(r2, bc2(0.041666668f), bc2(-0.5f)));       // cos r - 1
  float4 ta, tb;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(ta.x), "=f"(ta.y), "=f"(ta.z), "=f"(ta.w)
      : "r"(tabF + (((uint32_t)__float_as_int(km.x) & (SC6 - 1)) << 7)));
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(tb.x), "=f"(tb.y), "=f"(tb.z), "=f"(tb.w)
      : "r"(tabF + (((uint32_t)__float_as_int(km.y) & (SC6 - 1)) << 7)));
  const float2 ua = __ffma2_rn(make_float2(ta.x, ta.y), bc2(cm.x), make_float2(ta.z, ta.w));
  const float2 ub = __ffma2_rn(make_float2(tb.x, tb.y), bc2(cm.y), make_float2(tb.z, tb.w));
  sa = ta.x + fmaf(ta.y, sr.x, ua.x);
  ca = ta.y + fmaf(-ta.x, sr.x, ua.y);
  sb = tb.x + fmaf(tb.y, sr.y, ub.x);
  cb = tb.y + fmaf(-tb.x, sr.y, ub.y);
}

// NeRF step in float32: R <- R Rx(phi) Rz(g), p <- p + L R e_x   (cw, sw = cos / sin phi; cg, sg = cos / sin g)
__device__ __forceinline__ void nerf_step32(Se3f& f, float cw, float sw, float cg, float sg, float L) {
  const float2 y1 = __ffma2_rn(f.z01, bc2(sw), __fmul2_rn(f.y01, bc2(cw)));    // column y of R Rx
  const float2 z1 = __ffma2_rn(f.y01, bc2(-sw), __fmul2_rn(f.z01, bc2(cw)));   // column z of R Rx
  const float y1s = fmaf(f.z2, sw, f.y2 * cw), z1s = fmaf(-f.y2, sw, f.z2 * cw);
  const float2 xn = __ffma2_rn(y1, bc2(sg), __fmul2_rn(f.x01, bc2(cg)));       // column x of (R Rx) Rz
  const float2 yn = __ffma2_rn(f.x01, bc2(-sg), __fmul2_rn(y1, bc2(cg)));
  const float xns = fmaf(y1s, sg, f.x2 * cg), yns = fmaf(-f.x2, sg, y1s * cg);
  f.x01 = xn; f.y01 = yn; f.z01 = z1;
  f.x2 = xns; f.y2 = yns; f.z2 = z1s;
  f.p01 = __ffma2_rn(bc2(L), xn, f.p01);
  f.p2 = fmaf(L, xns, f.p2);
}

// stage elements [e0, e0 + 8) of 32 consecutive rows (frames frame0 ..) into a dense [frame][8] tile: lane -> (element
// q = lane & 7, frame lane >> 3 + 4 j), i.e. a warp instruction moves four 32-byte row segments into 128 consecutive bytes
// of shared memory.  Rows beyond the batch are skipped (their tile rows keep the zeros written at start-up); elements outside
// [0, len) are clamped to the row (read, never used).
__device__ __forceinline__ void f6_issue(uint32_t tile, const float* __restrict__ base, int64_t pitch, int64_t frame0, int64_t b,
                                         int e0, int len, int lane) {
  const int q = lane & 7, fr0 = lane >> 3;
  const int e = min(max(e0 + q, 0), len - 1);
  const float* src = base + (frame0 + fr0) * pitch + e;
  const int64_t hop = 4 * pitch;
  // rows fr with (fr >> 2) & 1 (= j & 1 here) keep their two 16-byte halves swapped, so that the per-lane LDS.128 of one
  // half (lane = frame) hits eight different 16-byte bank groups per quarter warp
  const uint32_t dst = tile + 4u * (uint32_t)lane;
  const uint32_t dst_swapped = tile + 4u * (uint32_t)(lane ^ 4);
  // a tile takes 32 bytes of each row; L2::128B makes L2 fetch the whole 128-byte line from DRAM, so that the next three
  // tiles of the row hit in L2 and DRAM sees full-line reads instead of four scattered 32-byte sectors
  if (frame0 + 32 <= b) {   // all 32 rows exist (every tile but the last of the batch): no per-copy predicate
#pragma unroll
    for (int j = 0; j < 8; j++) {
      asm volatile("cp.async.ca.shared.global.L2::128B [%0], [%1], 4;" ::"r"(((j & 1) ? dst_swapped : dst) + 128u * j), "l"(src) : "memory");
      src += hop;
    }
  } else {
    int64_t fr = frame0 + fr0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (fr < b) asm volatile("cp.async.ca.shared.global.L2::128B [%0], [%1], 4;" ::"r"(((j & 1) ? dst_swapped : dst) + 128u * j), "l"(src) : "memory");
      src += hop;
      fr += 4;
    }
  }
}
// same for an array whose rows are 8-byte aligned (even pitch, even e0): lane -> (element pair q = lane & 3, frame lane >> 2 + 8 j),
// four 8-byte-per-lane instructions instead of eight 4-byte ones
__device__ __forceinline__ void f6_issue8(uint32_t tile, const float* __restrict__ base, int64_t pitch, int64_t frame0, int64_t b,
                                          int e0, int len, int lane) {
  const int q = lane & 3, fr0 = lane >> 2;
  const int e = min(max(e0 + 2 * q, 0), len - 2);   // stays even: len is even
  const float* src = base + (frame0 + fr0) * pitch + e;
  const int64_t hop = 8 * pitch;
  // row fr = fr0 + 8 j: swapped halves when (fr >> 2) & 1 = (fr0 >> 2) & 1
  const uint32_t dst = tile + 4u * (uint32_t)((8 * fr0 + 2 * q) ^ (((fr0 >> 2) & 1) << 2));
  if (frame0 + 32 <= b) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      asm volatile("cp.async.ca.shared.global.L2::128B [%0], [%1], 8;" ::"r"(dst + 256u * j), "l"(src) : "memory");
      src += hop;
    }
  } else {
    int64_t fr = frame0 + fr0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (fr < b) asm volatile("cp.async.ca.shared.global.L2::128B [%0], [%1], 8;" ::"r"(dst + 256u * j), "l"(src) : "memory");
      src += hop;
      fr += 8;
    }
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// one 2-d TMA tensor store per 8-atom group and warp: box = 24 floats x 32 frames; rows beyond the batch are clipped by the unit
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const float* ssrc, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one NeRF step of a lane in global coordinates; returns the float32 position of the atom it placed
__device__ __forceinline__ void f6_step(float fd, float fa, double L, uint32_t tabB, Se3& f, float& ox, float& oy, float& oz) {
  double sw, cw, sg, cg;
  sincos_f6_x2(fd, fa, tabB, &sw, &cw, &sg, &cg);
  nerf_step(f, cw, sw, -cg, sg, L);   // bend by pi - theta: cos = -cos(theta), sin = sin(theta)
  ox = (float)f.p[0];
  oy = (float)f.p[1];
  oz = (float)f.p[2];
}

// the chain of one side: atoms k = s-2 .. 0 (SIDE 0: inputs d[k], a[k], L[k]) or s+2 .. n-1 (SIDE 1: d[k-3], a[k-2], L[k-1]).
// tD0 / tA0: this warp's dihedral / angle tiles (buffer 1 follows buffer 0 at + F6_IN_TILE words); outT: the output group.
// ROBUST = false: whole 8-atom groups run the table path unconditionally and only track max |angle| (returned): the caller
// repeats the tile with ROBUST = true -- every step tests its angles -- if any lane saw an angle beyond the table path's range.
// F32 = true: the float32 chain relative to the anchor (see above); *ext2 receives the squared extent of this lane's side
template <int SIDE, bool A8, bool ROBUST, bool F32>
__device__ __forceinline__ float f6_chain(Se3& f, const float* __restrict__ angles, const float* __restrict__ dihedrals, const double* L64,
                                          const float* L32, uint32_t tabB, uint32_t tabF, float* tD0, float* tA0, float* outT,
                                          const CUtensorMap* omap, float* __restrict__ dst, int64_t frame0, int64_t b, int n, bool valid,
                                          int lane, float* ext2) {
  const int s = n / 2, nd = n - 3, na = n - 2;
  constexpr int dir = SIDE ? 1 : -1;
  constexpr int sh_d = SIDE ? -3 : 0, sh_a = SIDE ? -2 : 0, sh_l = SIDE ? -1 : 0;
  const int k_first = SIDE ? s + 2 : s - 2, k_last = SIDE ? n - 1 : 0;     // inclusive
  float amax = 0.f;
  if (SIDE ? (k_first > k_last) : (k_first < k_last)) return amax;
  // float32 state: orientation as packed columns, position relative to the anchor atom, the anchor as a float32 hi/lo pair
  Se3f g32;
  float ahx = 0.f, ahy = 0.f, ahz = 0.f, alx = 0.f, aly = 0.f, alz = 0.f, e2 = 0.f;
  if (F32) {
    g32.x01 = make_float2((float)f.r[0], (float)f.r[3]); g32.x2 = (float)f.r[6];
    g32.y01 = make_float2((float)f.r[1], (float)f.r[4]); g32.y2 = (float)f.r[7];
    g32.z01 = make_float2((float)f.r[2], (float)f.r[5]); g32.z2 = (float)f.r[8];
    g32.p01 = make_float2(0.f, 0.f); g32.p2 = 0.f;
    ahx = (float)f.p[0]; ahy = (float)f.p[1]; ahz = (float)f.p[2];
    alx = (float)(f.p[0] - (double)ahx); aly = (float)(f.p[1] - (double)ahy); alz = (float)(f.p[2] - (double)ahz);
  }
  const int g_first = k_first >> 3, g_last = k_last >> 3;
  const uint32_t sD0 = (uint32_t)__cvta_generic_to_shared(tD0), sA0 = (uint32_t)__cvta_generic_to_shared(tA0);
  const int swz = (lane >> 2) & 1;                           // this row's halves are swapped in the tiles
  auto issue_a = [&](uint32_t tile, int e0) {
    if (A8) f6_issue8(tile, angles, na, frame0, b, e0, na, lane);
    else f6_issue(tile, angles, na, frame0, b, e0, na, lane);
  };
  int buf = 0;
  f6_issue(sD0, dihedrals, nd, frame0, b, 8 * g_first + sh_d, nd, lane);
  issue_a(sA0, 8 * g_first + sh_a);
  cp_async_commit();
  for (int g = g_first; g != g_last + dir; g += dir) {
    const int gn = g + dir;
    if (gn != g_last + dir) {
      f6_issue(sD0 + 4u * F6_IN_TILE * (buf ^ 1), dihedrals, nd, frame0, b, 8 * gn + sh_d, nd, lane);
      issue_a(sA0 + 4u * F6_IN_TILE * (buf ^ 1), 8 * gn + sh_a);
    }
    cp_async_commit();
    cp_async_wait1();
    __syncwarp();
    const float* td = tD0 + buf * F6_IN_TILE + lane * 8;     // this lane's (frame's) eight dihedrals / angles of the group
    const float* ta = tA0 + buf * F6_IN_TILE + lane * 8;
    // atoms of this group that belong to the side, in walking order
    const int ka = SIDE ? max(8 * g, k_first) : min(8 * g + 7, k_first);
    const int kb = SIDE ? min(8 * g + 7, k_last) : max(8 * g, k_last);    // inclusive
    const bool whole = SIDE ? (ka == 8 * g && kb == 8 * g + 7) : (ka == 8 * g + 7 && kb == 8 * g);
    if (whole) {
      // the eight positions of the group stay in registers until all are known: the buffer is written (and handed to the
      // tensor store) at the END of the group, i.e. a full group after the previous store was issued -- waiting for that
      // store's shared-memory read any earlier cost 20 % of the kernel
      float pend[24];
#pragma unroll
      for (int half = 0; half < 2; half++) {
        const int hm = SIDE ? half : 1 - half;   // memory half of the group: atoms 8 g + 4 hm .. + 3
        const float4 d4 = *reinterpret_cast<const float4*>(td + 4 * (hm ^ swz));
        const float4 a4 = *reinterpret_cast<const float4*>(ta + 4 * (hm ^ swz));
        const float dv[4] = {d4.x, d4.y, d4.z, d4.w}, av[4] = {a4.x, a4.y, a4.z, a4.w};
        if (F32) {
          const float* Lp = L32 + (8 * g + 4 * hm + sh_l);
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int pu = SIDE ? u : 3 - u;       // atom 4 hm + pu is placed by step u of this half
            float sw, cw, sg, cg;
            amax = fmaxf(amax, fmaxf(fabsf(dv[pu]), fabsf(av[pu])));
            sincos_tab32_x2(dv[pu], av[pu], tabF, sw, cw, sg, cg);
            nerf_step32(g32, cw, sw, -cg, sg, Lp[pu]);
            const float2 o01 = __fadd2_rn(make_float2(ahx, ahy), __fadd2_rn(make_float2(alx, aly), g32.p01));
            pend[12 * hm + 3 * pu] = o01.x;
            pend[12 * hm + 3 * pu + 1] = o01.y;
            pend[12 * hm + 3 * pu + 2] = ahz + (alz + g32.p2);
          }
        } else {
          const double* Lp = L64 + (8 * g + 4 * hm + sh_l);
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int pu = SIDE ? u : 3 - u;       // atom 4 hm + pu is placed by step u of this half
            double sw, cw, sg, cg;
            if (ROBUST) {
              sincos_f6_x2(dv[pu], av[pu], tabB, &sw, &cw, &sg, &cg);
            } else {
              amax = fmaxf(amax, fmaxf(fabsf(dv[pu]), fabsf(av[pu])));
              sincos_tab64_x2(dv[pu], av[pu], tabB, &sw, &cw, &sg, &cg);
            }
            nerf_step(f, cw, sw, -cg, sg, Lp[pu]);
            pend[12 * hm + 3 * pu] = (float)f.p[0];
            pend[12 * hm + 3 * pu + 1] = (float)f.p[1];
            pend[12 * hm + 3 * pu + 2] = (float)f.p[2];
          }
        }
      }
      if (F32) e2 = fmaxf(e2, fmaf(g32.p01.x, g32.p01.x, fmaf(g32.p01.y, g32.p01.y, g32.p2 * g32.p2)));
      if (lane == 0) bulk_wait_read0();          // the previous group has left the buffer (the storing lane tracks the bulk groups)
      __syncwarp();
      // a lane's row is 96 bytes: lanes l and l + 4 of a quarter warp would hit the same banks with the same 16-byte chunk
      // (half of the store wavefronts were conflict replays).  Lanes with bit 2 set write chunk q + 1 (mod 6) in the q-th
      // instruction: the eight lanes of a phase then cover eight different bank groups; the row ends up dense all the same.
      float4* row4 = reinterpret_cast<float4*>(outT + lane * 24);
      const bool rot = (lane >> 2) & 1;
#pragma unroll
      for (int q = 0; q < 6; q++) {
        const int c1 = (q + 1) % 6;
        const float4 v = make_float4(rot ? pend[4 * c1] : pend[4 * q], rot ? pend[4 * c1 + 1] : pend[4 * q + 1],
                                     rot ? pend[4 * c1 + 2] : pend[4 * q + 2], rot ? pend[4 * c1 + 3] : pend[4 * q + 3]);
        row4[rot ? c1 : q] = v;
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(omap, outT, 24 * g, (int)frame0);
        bulk_commit();
      }
    } else {
      for (int k = ka; k != kb + dir; k += dir) {
        const int slot = k & 7;
        float ox, oy, oz;
        if (F32) {
          const float fd = td[slot ^ (4 * swz)], fa = ta[slot ^ (4 * swz)];
          float sw, cw, sg, cg;
          amax = fmaxf(amax, fmaxf(fabsf(fd), fabsf(fa)));     // beyond the table range: the second pass redoes the tile
          sincos_tab32_x2(fd, fa, tabF, sw, cw, sg, cg);
          nerf_step32(g32, cw, sw, -cg, sg, L32[k + sh_l]);
          ox = ahx + (alx + g32.p01.x); oy = ahy + (aly + g32.p01.y); oz = ahz + (alz + g32.p2);
        } else {
          f6_step(td[slot ^ (4 * swz)], ta[slot ^ (4 * swz)], L64[k + sh_l], tabB, f, ox, oy, oz);
        }
        if (valid) { dst[3 * k] = ox; dst[3 * k + 1] = oy; dst[3 * k + 2] = oz; }
      }
      if (F32) e2 = fmaxf(e2, fmaf(g32.p01.x, g32.p01.x, fmaf(g32.p01.y, g32.p01.y, g32.p2 * g32.p2)));
    }
    __syncwarp();
    buf ^= 1;
  }
  cp_async_wait0();
  *ext2 = e2;
  return amax;
}

template <bool A8, int F6_WARPS, bool F32>
__global__ void __launch_bounds__(32 * F6_WARPS, 1) backmap_fwd6_kernel(const __grid_constant__ CUtensorMap omap, const float* __restrict__ lengths,
                                                                     const float* __restrict__ angles, const float* __restrict__ dihedrals,
                                                                     int64_t b, int n, float* __restrict__ xyz, const double2* __restrict__ tab,
                                                                     float ext2_limit) {
  extern __shared__ __align__(1024) float smem6[];
  double2* tabS = reinterpret_cast<double2*>(smem6);                       // [SC6][8]
  double* L64 = reinterpret_cast<double*>(smem6 + SC6 * 8 * 4);            // [n - 1], region rounded up to 128 bytes
  const int lwords = ((2 * (n - 1) + 31) / 32) * 32;
  // float32 first pass: hi/lo table [SC6][8] of float4 and the bond lengths as float32
  float4* tabF4 = reinterpret_cast<float4*>(smem6 + SC6 * 8 * 4 + lwords);
  float* L32 = smem6 + SC6 * 8 * 4 + lwords + (F32 ? SC6 * 8 * 4 : 0);
  const int l32words = F32 ? ((n - 1 + 31) / 32) * 32 : 0;
  float* warp_base = L32 + l32words;

  constexpr int F6_THREADS = 32 * F6_WARPS;
  const int tid = threadIdx.x, warp = tid >> 5;
  int lane = tid & 31;
  asm volatile("mov.u32 %0, %0;" : "+r"(lane));   // opaque: keeps ptxas from re-reading SR_TID.X inside the loops
  const int side = warp & 1, pair = warp >> 1;
  for (int i = tid; i < SC6 * 8; i += F6_THREADS) tabS[i] = __ldg(tab + (i >> 3) * (SC_TABLE / SC6));
  for (int i = tid; i < n - 1; i += F6_THREADS) L64[i] = (double)__ldg(lengths + i);
  if (F32) {
    for (int i = tid; i < SC6 * 8; i += F6_THREADS) {
      const double2 t = __ldg(tab + (i >> 3) * (SC_TABLE / SC6));
      const float sh = (float)t.x, ch = (float)t.y;
      tabF4[i] = make_float4(sh, ch, (float)(t.x - (double)sh), (float)(t.y - (double)ch));
    }
    for (int i = tid; i < n - 1; i += F6_THREADS) L32[i] = __ldg(lengths + i);
  }
  float* wmem = warp_base + (size_t)warp * F6_WARP_WORDS;
  for (int i = lane; i < F6_WARP_WORDS; i += 32) wmem[i] = 0.f;
  __syncthreads();

  const uint32_t tabB = (uint32_t)__cvta_generic_to_shared(tabS + (lane & 7));   // this lane's table replica
  const uint32_t tabF = (uint32_t)__cvta_generic_to_shared(tabF4 + (lane & 7));
  float* tD0 = wmem;                                                       // dihedral tiles: buffer 1 at + F6_IN_TILE
  float* tA0 = wmem + 2 * F6_IN_TILE;                                      // angle tiles
  float* outT = wmem + 4 * F6_IN_TILE;                                     // 128-byte aligned: 4 KB into a 7 KB region
  float* partner_out = warp_base + (size_t)(warp ^ 1) * F6_WARP_WORDS + 4 * F6_IN_TILE;
  const uint32_t sA0 = (uint32_t)__cvta_generic_to_shared(tA0);

  const int s = n / 2;
  const int na = n - 2;
  const int64_t n_tiles = (b + 31) >> 5;
  const int pairs_per_cta = F6_WARPS / 2;

  for (int64_t tile = (int64_t)blockIdx.x * pairs_per_cta + pair; tile < n_tiles; tile += (int64_t)gridDim.x * pairs_per_cta) {
    const int64_t frame0 = tile << 5;
    const int64_t frame = frame0 + lane;
    const bool valid = frame < b;
    float* dst = xyz + (valid ? frame : frame0) * (int64_t)(3 * n);
    // pass 0: table path without per-step range tests; pass 1 (only if some angle of the tile was beyond the table path's
    // range -- never for model outputs): the same with every step tested
    for (int pass = 0; pass < 2; pass++) {

    // ---- planar prologue: SE(2) product of bonds 0 .. s-2, split between the two warps of the pair ---------------------
    // right warp: bonds [0, h), left warp: bonds [h, s-1); h is a multiple of 8 (tile aligned)
    const int h = ((s - 1) / 2) & ~7;
    const int k_lo = side ? 0 : h, k_hi = side ? h : s - 1;
    Se2 pl{1.0, 0.0, 0.0, 0.0};
    {
      int buf = 0;
      auto issue_a = [&](uint32_t tile_addr, int e0) {
        if (A8) f6_issue8(tile_addr, angles, na, frame0, b, e0, na, lane);
        else f6_issue(tile_addr, angles, na, frame0, b, e0, na, lane);
      };
      if (k_lo < k_hi) issue_a(sA0, k_lo);
      cp_async_commit();
      for (int k0 = k_lo; k0 < k_hi; k0 += F6_TILE) {
        if (k0 + F6_TILE < k_hi) issue_a(sA0 + 4u * F6_IN_TILE * (buf ^ 1), k0 + F6_TILE);
        cp_async_commit();
        cp_async_wait1();
        __syncwarp();
        const float* t = tA0 + buf * F6_IN_TILE + lane * 8;
        const int cnt = min(F6_TILE, k_hi - k0);
        const int swz = (lane >> 2) & 1;
        const float4 t0 = *reinterpret_cast<const float4*>(t + 4 * swz), t1 = *reinterpret_cast<const float4*>(t + 4 * (swz ^ 1));
        const float tv[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
        for (int q = 0; q < F6_TILE; q += 2) {   // two bonds per iteration: their sin/cos are computed as one packed pair
          if (q + 1 < cnt) {
            const int k = k0 + q;                // even (k0 is a multiple of 8)
            double s0, c0, s1, c1;
            sincos_f6_x2(tv[q], tv[q + 1], tabB, &s0, &c0, &s1, &c1);
            const double La = L64[k], Lb = L64[k + 1];
            pl.x = fma(La, pl.c, pl.x);
            pl.y = fma(La, pl.s, pl.y);
            {   // turn by -(-1)^k (pi - theta_k): cos = -cos(theta), sin = -(-1)^k sin(theta); k even
              const double c2 = -(pl.c * c0) + pl.s * s0, s2 = -(pl.c * s0) - pl.s * c0;
              pl.c = c2; pl.s = s2;
            }
            pl.x = fma(Lb, pl.c, pl.x);
            pl.y = fma(Lb, pl.s, pl.y);
            {   // bond k + 1 (odd): sin = +sin(theta)
              const double c2 = -(pl.c * c1) - pl.s * s1, s2 = pl.c * s1 - pl.s * c1;
              pl.c = c2; pl.s = s2;
            }
          } else if (q < cnt) {
            planar_step(pl, L64[k0 + q], true, tv[q], k0 + q);
          }
        }
        __syncwarp();
        buf ^= 1;
      }
      cp_async_wait0();
    }
    // swap the partial products: mine goes into my output buffer, the partner reads it from there
    {
      double* x = reinterpret_cast<double*>(outT) + 4 * lane;
      x[0] = pl.c; x[1] = pl.s; x[2] = pl.x; x[3] = pl.y;
      asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory");
      const double* y = reinterpret_cast<const double*>(partner_out) + 4 * lane;
      const Se2 other{y[0], y[1], y[2], y[3]};
      pl = side ? se2_mul(pl, other) : se2_mul(other, pl);   // bonds [0, h) first, then [h, s-1)
      asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory");   // both have read: the buffers may be reused
    }

    // ---- anchor: the three middle atoms at their planar positions, anchor frame of this side ---------------------------
    Se3 f;
    {
      const double dm_c = pl.c, dm_s = pl.s, am_x = pl.x, am_y = pl.y;      // direction of bond s-1, position of atom s-1
      const float th_mid = valid ? __ldg(angles + frame * na + (s - 1)) : 2.f;
      Se2 mid = pl;
      planar_step(mid, L64[s - 1], true, th_mid, s - 1);
      const double a0_x = mid.x, a0_y = mid.y, dp_c = mid.c, dp_s = mid.s;
      const double ap_x = fma(L64[s], dp_c, a0_x), ap_y = fma(L64[s], dp_s, a0_y);
      const double zs = (dm_c * dp_s - dm_s * dp_c) >= 0.0 ? 1.0 : -1.0;
      const double xx = side == 0 ? -dm_c : dp_c, xy = side == 0 ? -dm_s : dp_s, zz = side == 0 ? -zs : zs;
      f.r[0] = xx; f.r[1] = -zz * xy; f.r[2] = 0.0;
      f.r[3] = xy; f.r[4] = zz * xx;  f.r[5] = 0.0;
      f.r[6] = 0.0; f.r[7] = 0.0;     f.r[8] = zz;
      f.p[0] = side == 0 ? am_x : ap_x; f.p[1] = side == 0 ? am_y : ap_y; f.p[2] = 0.0;
      if (valid) {
        if (side == 0) {
          float* o = dst + 3 * (s - 1);
          o[0] = (float)am_x; o[1] = (float)am_y; o[2] = 0.f;
          o[3] = (float)a0_x; o[4] = (float)a0_y; o[5] = 0.f;
        } else {
          float* o = dst + 3 * (s + 1);
          o[0] = (float)ap_x; o[1] = (float)ap_y; o[2] = 0.f;
        }
      }
    }

    // ---- the chain of this side (side is warp uniform: the instantiations keep every register index static) ----------
    // pass 0: float32 chain (F32) or float64 chain without per-step range tests; pass 1: float64 chain, every step tested
    float amax, ext2 = 0.f;
    if (pass == 0) {
      if (side) amax = f6_chain<1, A8, false, F32>(f, angles, dihedrals, L64, L32, tabB, tabF, tD0, tA0, outT, &omap, dst, frame0, b, n, valid, lane, &ext2);
      else amax = f6_chain<0, A8, false, F32>(f, angles, dihedrals, L64, L32, tabB, tabF, tD0, tA0, outT, &omap, dst, frame0, b, n, valid, lane, &ext2);
    } else {
      if (side) amax = f6_chain<1, A8, true, false>(f, angles, dihedrals, L64, L32, tabB, tabF, tD0, tA0, outT, &omap, dst, frame0, b, n, valid, lane, &ext2);
      else amax = f6_chain<0, A8, true, false>(f, angles, dihedrals, L64, L32, tabB, tabF, tD0, tA0, outT, &omap, dst, frame0, b, n, valid, lane, &ext2);
    }
    if (lane == 0) bulk_wait_read0();   // the output buffer doubles as the exchange buffer of the next tile / pass
    __syncwarp();
    // both warps of the pair must agree on repeating (they meet at the named barrier of the prologue): the flag travels
    // through the exchange buffers
    if (pass == 0) {
      // rows beyond the batch hold zeros (no extent); a NaN extent compares false: NaN frames stay NaN, they need no redo
      const bool bad = __any_sync(0xffffffffu, amax >= SC6_LIMIT || (F32 && valid && ext2 >= ext2_limit));
      if (lane == 0) reinterpret_cast<int*>(outT)[0] = bad ? 1 : 0;
      asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory");
      const bool redo = bad || reinterpret_cast<const int*>(partner_out)[0] != 0;
      asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory");
      if (!redo) break;
      if (lane == 0) bulk_wait0();   // the first pass's tensor stores must have landed before the second pass rewrites them
      __syncwarp();
    }
    }   // pass
  }
  if (lane == 0) bulk_wait0();
}

// ====================================================================================================
// dihedrals_to_cartesian on an arbitrary start chain (two-sided or one-way)
// ====================================================================================================
__device__ __forceinline__ void load3(const float* p, double* v) {
  v[0] = p[0];
  v[1] = p[1];
  v[2] = p[2];
}

__global__ void d2c_general_kernel(const float* __restrict__ dihedrals, const float* __restrict__ chain, int64_t cstride,
                                   int64_t b, int n, int one_way, float* __restrict__ xyz, const double2* __restrict__ tab) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const int64_t frame = (int64_t)blockIdx.x * wpc + warp;
  if (frame >= b) return;
  const int per_warp = (n - 3) + 3 * n;
  float* sD = smem + (size_t)warp * per_warp;
  float* sO = sD + (n - 3);
  stage_in(sD, dihedrals + frame * (int64_t)(n - 3), n - 3, lane);
  stage_in(sO, chain + frame * cstride, 3 * n, lane);
  __syncwarp();

  // side description: atom of side-local index m is a0 + dir*m; dihedral of step i is d0 + dir*i
  int side, q, width, steps, a0, dir, d0;
  const int s = n / 2;
  if (one_way) {
    side = 0; q = lane; width = 32; steps = n - 3; a0 = 0; dir = 1; d0 = 0;
  } else {
    side = lane >> 4; q = lane & 15; width = 16;
    if (side == 0) { steps = s - 1; a0 = s + 1; dir = -1; d0 = s - 2; }
    else           { steps = n - s - 2; a0 = s - 1; dir = 1; d0 = s - 1; }
  }
  int cs = (steps + width - 1) / width;
  cs |= 1;
  const int i0 = min(steps, q * cs), i1 = min(steps, i0 + cs);
  // originals of the two atoms before this lane's first moved atom (a neighbour lane overwrites them)
  double pa[3] = {0, 0, 0}, pb[3] = {0, 0, 0};
  if (i0 < i1) {
    load3(sO + 3 * (a0 + dir * (i0 + 1)), pa);
    load3(sO + 3 * (a0 + dir * (i0 + 2)), pb);
  }
  __syncwarp();
  Se3 f;
  se3_identity(f);
  for (int i = i0; i < i1; i++) {
    // A_i: rotate by +dihedral_i about the axis pa -> pb through pb (column-vector convention; the
    // reference negates the angle and multiplies row vectors on the right, misc/backmapping.py:1896-1909)
    double ux = pb[0] - pa[0], uy = pb[1] - pa[1], uz = pb[2] - pa[2];
    const double inv = rsqrt(ux * ux + uy * uy + uz * uz);
    ux *= inv; uy *= inv; uz *= inv;
    double sw, cw;
    sincos_tab(sD[d0 + dir * i], tab, &sw, &cw);
    const double oc = 1.0 - cw;
    Se3 a;
    a.r[0] = cw + oc * ux * ux;      a.r[1] = oc * ux * uy - sw * uz; a.r[2] = oc * ux * uz + sw * uy;
    a.r[3] = oc * uy * ux + sw * uz; a.r[4] = cw + oc * uy * uy;      a.r[5] = oc * uy * uz - sw * ux;
    a.r[6] = oc * uz * ux - sw * uy; a.r[7] = oc * uz * uy + sw * ux; a.r[8] = cw + oc * uz * uz;
#pragma unroll
    for (int r = 0; r < 3; r++) a.p[r] = pb[r] - (a.r[3 * r] * pb[0] + a.r[3 * r + 1] * pb[1] + a.r[3 * r + 2] * pb[2]);
    f = se3_mul(f, a);
    const int k = a0 + dir * (i + 3);
    double pc[3];
    load3(sO + 3 * k, pc);
#pragma unroll
    for (int r = 0; r < 3; r++) sO[3 * k + r] = (float)(f.r[3 * r] * pc[0] + f.r[3 * r + 1] * pc[1] + f.r[3 * r + 2] * pc[2] + f.p[r]);
#pragma unroll
    for (int r = 0; r < 3; r++) { pa[r] = pb[r]; pb[r] = pc[r]; }
  }
  Se3 pre = se3_exclusive_scan(f, q, width);
  for (int i = i0; i < i1; i++) {
    const int k = a0 + dir * (i + 3);
    const double lx = sO[3 * k], ly = sO[3 * k + 1], lz = sO[3 * k + 2];
    sO[3 * k] = (float)(pre.r[0] * lx + pre.r[1] * ly + pre.r[2] * lz + pre.p[0]);
    sO[3 * k + 1] = (float)(pre.r[3] * lx + pre.r[4] * ly + pre.r[5] * lz + pre.p[1]);
    sO[3 * k + 2] = (float)(pre.r[6] * lx + pre.r[7] * ly + pre.r[8] * lz + pre.p[2]);
  }
  __syncwarp();
  float* dst = xyz + frame * (int64_t)(3 * n);
  for (int i = lane; i < 3 * n; i += 32) dst[i] = sO[i];
}

constexpr int BWD_CA = 13;   // atoms per thread that size the thread count per frame; odd => conflict-free strided walks

// ====================================================================================================
// backward, version 3: float32 wrench recurrences about a MOVING pivot.
//
// Every internal coordinate moves one end of the chain rigidly (the end that does not hold the anchor), so its
// gradient is <axis, torque about the pivot> or <direction, force> of the wrench of that end.  Version 2 summed
// (g, x cross g) about the ORIGIN in float64 and subtracted pivot cross force afterwards (a cancellation of two
// ~80 nm lever arms that float32 cannot carry).  Here both ends are walked from the chain end towards the anchor
// and the torque is kept about the atom the walk has reached:
//     S_i = S_{i-1} + g(a_i),      M_i = M_{i-1} - (x(a_{i+1}) - x(a_i)) x S_i        (M_i is about x(a_{i+1}))
// with a_i = i on the left of the anchor and a_i = n-1-i on the right.  Lever arms are single bonds, nothing
// cancels, float32 is enough (errors random-walk to ~2e-6 relative over 750 steps, the tolerance is 1e-5), and
// the formulas of the two ends are mirror images: with e1 = x(a_{i+1}) - x(a_i), e2 = x(a_{i+2}) - x(a_{i+1})
//     dihedral term = -<unit(e2), M_i>,   angle term = <unit(e1 x e2), M_i>,   length term = -<unit(e1), S_i>
// on both sides (tests/test_oracle_kats.py::test_moving_pivot_backward_formulas states this algorithm sequentially in
// float64 and pins it to float64 autograd of the reference order; the GPU tests compare the kernel itself).
// Chunks of consecutive walk steps per thread; chunk aggregates are combined by a scan whose operator shifts the
// pivot:  (S_A, M_A about p_A) then (S_C, M_C about p_C)  ->  (S_A + S_C, M_C + M_A - (p_C - p_A) x S_A).
// Left of the anchor an angle / length also moves the anchor itself along the planar chain, i.e. the whole
// molecule: + (-1)^k <e_z, total torque about planar atom k+1>  and  + <planar bond direction, total force>; the
// planar chain is an SE(2) scan in float64 (left threads only).
// The term of the cut behind atom a_i overwrites g(a_i) in place; the three output rows are gathered from there.
// ====================================================================================================
struct WrenchF {
  float s[3];
  float m[3];
};

// A then C (C further along the walk): pivot moves from pa to pc
__device__ __forceinline__ void wrench_append(WrenchF& c, const WrenchF& a, const float* pa, const float* pc) {
  const float d0 = pc[0] - pa[0], d1 = pc[1] - pa[1], d2 = pc[2] - pa[2];
  c.m[0] += a.m[0] - (d1 * a.s[2] - d2 * a.s[1]);
  c.m[1] += a.m[1] - (d2 * a.s[0] - d0 * a.s[2]);
  c.m[2] += a.m[2] - (d0 * a.s[1] - d1 * a.s[0]);
  c.s[0] += a.s[0];
  c.s[1] += a.s[1];
  c.s[2] += a.s[2];
}

// 16-byte staging of one row by NT threads: uniform trip count, four predicated copies with immediate offsets per
// iteration.  `region` is 16-byte aligned with 4 floats of slack; the copy starts at the 16-byte boundary at or below
// `src` (those <= 3 leading floats belong to the previous row of the same tensor), so the body moves 16 bytes per request;
// returns the address of element 0.
__device__ __forceinline__ void cp_async16_s(uint32_t dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
template <int NT>
__device__ __forceinline__ float* stage_row16(float* region, const float* __restrict__ src, int len, int t) {
  const int sh = (int)((reinterpret_cast<uintptr_t>(src) & 15) >> 2);
  const float* asrc = src - sh;
  const int total = sh + len;
  const int nvec = total >> 2;
  uint32_t d = (uint32_t)__cvta_generic_to_shared(region) + 16u * (uint32_t)t;
  const float* s = asrc + 4 * t;
  int left = nvec - t;                          // copy i of this thread is in range iff NT * i < left
  const int iters = (nvec + NT - 1) / NT;
  int i = 0;
#pragma unroll 1
  for (; i + 4 <= iters; i += 4) {
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (NT * u < left) cp_async16_s(d + 16u * NT * u, s + 4 * NT * u);
    d += 64u * NT;
    s += 16 * NT;
    left -= 4 * NT;
  }
#pragma unroll
  for (int u = 0; u < 3; u++)
    if (i + u < iters && NT * u < left) cp_async16_s(d + 16u * NT * u, s + 4 * NT * u);
  const int e = (nvec << 2) + t;
  if (e < total) cp_async4(region + e, asrc + e);
  return region + sh;
}

// GEN = false: only dihedral gradients, no planar terms (the default ADC training step and the standalone
// dihedrals_to_cartesian op); GEN = true: everything (angle / length gradients, planar mode).
template <int T, bool GEN>
__global__ void __launch_bounds__((T < 128 ? 128 : T)) backmap_bwd3_kernel(const BwdParams p, const double2* __restrict__ tab) {
  constexpr int CTA = T < 128 ? 128 : T;
  constexpr int FPC = CTA / T;
  constexpr int WPF = T / 32;
  static_assert(T >= 32 && T % 32 == 0, "whole warps per frame");
  extern __shared__ __align__(16) float smem[];
  __shared__ float wagg[FPC][WPF][9];      // per-warp inclusive aggregate: S, M, end pivot
  __shared__ float side_tot[FPC][2][6];    // per-side total wrench about x_{dr0}
  __shared__ double psum[FPC][WPF][4];

  const int tid = threadIdx.x, lane = tid & 31;
  const int g = tid / T, t = tid % T, wf = t >> 5;
  const int64_t frame = (int64_t)blockIdx.x * FPC + g;
  const bool active = frame < p.b;
  const int64_t fr = active ? frame : 0;
  const int n = p.n;
  const bool planar = GEN && p.planar;
  const bool extras = GEN && !p.planar && (p.grad_angles || p.grad_lengths) && p.mid > 1;   // left-of-anchor planar terms
  const bool need_planar = planar || extras;

  // planar inputs: the whole rows in planar mode, only the bonds / angles left of the anchor for the left-side terms
  const int nL = planar ? n - 1 : p.dr0, nA = planar ? n - 2 : p.dr0;
  const int rX = (3 * n + 27) & ~3, rG = (3 * n + 7) & ~3, rL = (nL + 7) & ~3, rA = (nA + 7) & ~3;
  const int rP = extras ? 8 * (T / 2) : 0;   // one SE(2) element (4 doubles) per left thread: planar halves computed by the right threads
  float* base = smem + (size_t)g * (rX + rG + (need_planar ? rL + rA : 0) + rP);
  float* sX = base + 4;   // 4 floats of slack in front: the walk may read (never use) one atom before the chain
  float* sG = base + rX;
  float* sL = sG + rG;
  float* sA = sL + rL;
  double* pexch = reinterpret_cast<double*>(sA + rA);   // (rX + rG + rL + rA) is a multiple of 4 floats: 16-byte aligned
  if (active) {
    sG = stage_row16<T>(sG, p.grad_xyz + fr * (int64_t)(3 * n), 3 * n, t);
    if (!planar) sX = stage_row16<T>(sX, p.xyz + fr * (int64_t)(3 * n), 3 * n, t);
    if (need_planar) {
      sL = stage_row16<T>(sL, p.lengths + fr * p.lstride, nL, t);
      sA = stage_row16<T>(sA, p.angles + fr * (int64_t)(n - 2), nA, t);
    }
  }
  cp_async_wait_all();
  __syncthreads();

  // ---- sides and chunks ------------------------------------------------------------------------------------
  // left: cuts 0 .. dr0-1 (walk i = cut), right: cuts n-2 .. dr0 (walk i = n-2-cut); the walk visits a_0 .. a_{steps}
  const int TL = p.dr0 > 0 ? T / 2 : 0;
  const int side = t >= TL ? 1 : 0;
  const int q = side ? t - TL : t;                    // thread index within the side
  const int TS = side ? T - TL : TL;
  const int W = TS < 32 ? TS : 32;                    // shuffle segment (16 only for T = 32 with two sides)
  const int qs = q & (W - 1);                         // lane within the segment
  const int steps = side ? n - 1 - p.dr0 : p.dr0;
  const int C = ((steps + TS - 1) / TS) | 1;          // odd: the strided shared-memory walks are conflict-free
  const int i0 = min(steps, q * C), i1 = min(steps, i0 + C);
  const int st = side ? -3 : 3;
  const int a0 = side ? n - 1 - i0 : i0;              // atom of walk index i0

  // ---- planar chain ------------------------------------------------------------------------------------------
  Se2 pl_ex{1.0, 0.0, 0.0, 0.0}, pl_ex2{1.0, 0.0, 0.0, 0.0};   // left threads with extras: planar state before bond i0 / imid
  int imid = 0;
  if (planar) {
    // chain_in_plane backward: recompute the planar coordinates into sX (ascending chunks of BWD_CA atoms)
    const int k0 = min(n, t * BWD_CA), k1 = min(n, k0 + BWD_CA);
    Se2 part{1.0, 0.0, 0.0, 0.0};
    if (active)
      for (int k = k0; k < k1 && k < n - 1; k++) planar_step_tab(part, (double)sL[k], k < n - 2, k < n - 2 ? sA[k] : 0.f, k, tab);
    Se2 inc = part;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      Se2 up = se2_shfl_up(inc, d);
      if (lane >= d) inc = se2_mul(up, inc);
    }
    Se2 ex = se2_shfl_up(inc, 1);
    if (lane == 0) ex = Se2{1.0, 0.0, 0.0, 0.0};
    if (lane == 31) { psum[g][wf][0] = inc.c; psum[g][wf][1] = inc.s; psum[g][wf][2] = inc.x; psum[g][wf][3] = inc.y; }
    __syncthreads();
    Se2 run{1.0, 0.0, 0.0, 0.0};
    for (int w = 0; w < wf; w++) run = se2_mul(run, Se2{psum[g][w][0], psum[g][w][1], psum[g][w][2], psum[g][w][3]});
    run = se2_mul(run, ex);
    if (active) {
      if (t == 0) { sX[0] = 0.f; sX[1] = 0.f; sX[2] = 0.f; }
      for (int k = k0; k < k1 && k < n - 1; k++) {
        planar_step_tab(run, (double)sL[k], k < n - 2, k < n - 2 ? sA[k] : 0.f, k, tab);
        sX[3 * (k + 1)] = (float)run.x; sX[3 * (k + 1) + 1] = (float)run.y; sX[3 * (k + 1) + 2] = 0.f;
      }
    }
    __syncthreads();
  } else if (extras) {
    // Planar SE(2) chain of the bonds left of the anchor.  The chunk of left thread q is cut in two: the left thread
    // walks the first half, its idle partner on the right side (thread q + T/2) the second half, each leaving the planar
    // atoms in the frame of its own first bond in the slots of (L_k, theta_k); pass 2 then only needs float32 arithmetic on
    // these local coordinates (< 1 nm) and the float64 prefix of the half it is in.
    const int CLs = ((p.dr0 + TL - 1) / TL) | 1;
    const int lo = min(p.dr0, q * CLs), hi = min(p.dr0, lo + CLs);
    imid = lo + (hi - lo + 1) / 2;
    const int ka = side == 0 ? lo : imid, kb = side == 0 ? imid : hi;
    Se2 part{1.0, 0.0, 0.0, 0.0};
    if (active)
      for (int k = ka; k < kb; k++) {
        part.x = fma((double)sL[k], part.c, part.x);
        part.y = fma((double)sL[k], part.s, part.y);
        double sn, cs;
        sincos_tab(sA[k], tab, &sn, &cs);   // k <= dr0 - 1 = n/2 - 2 < n - 2: every left bond is followed by a turn
        const double cw = -cs, sw = (k & 1) ? sn : -sn;
        const double c2 = part.c * cw - part.s * sw, s2 = part.c * sw + part.s * cw;
        part.c = c2; part.s = s2;
        sL[k] = (float)part.x;
        sA[k] = (float)part.y;
      }
    if (side == 1) { pexch[4 * q] = part.c; pexch[4 * q + 1] = part.s; pexch[4 * q + 2] = part.x; pexch[4 * q + 3] = part.y; }
    __syncthreads();
    // scan over the left threads of (first half o second half); right threads carry the identity
    Se2 inc{1.0, 0.0, 0.0, 0.0};
    if (side == 0) inc = se2_mul(part, Se2{pexch[4 * q], pexch[4 * q + 1], pexch[4 * q + 2], pexch[4 * q + 3]});
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      Se2 up = se2_shfl_up(inc, d);
      if (lane >= d) inc = se2_mul(up, inc);
    }
    Se2 ex = se2_shfl_up(inc, 1);
    if (lane == 0) ex = Se2{1.0, 0.0, 0.0, 0.0};
    if (lane == 31) { psum[g][wf][0] = inc.c; psum[g][wf][1] = inc.s; psum[g][wf][2] = inc.x; psum[g][wf][3] = inc.y; }
    __syncthreads();
    Se2 prev{1.0, 0.0, 0.0, 0.0};
    for (int w = 0; w < wf; w++) prev = se2_mul(prev, Se2{psum[g][w][0], psum[g][w][1], psum[g][w][2], psum[g][w][3]});
    pl_ex = se2_mul(prev, ex);         // before the first half
    pl_ex2 = se2_mul(pl_ex, part);     // before the second half (left threads: part is the first half)
  }

  // ---- pass 1: wrench of this thread's chunk about x(a_{i1}) ------------------------------------------------------
  WrenchF w{{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
  float xs[3], xe[3];   // pivots before / after the chunk: x(a_{i0}), x(a_{i1})
  {
    const float* px = sX + 3 * a0;
    const float* pg = sG + 3 * a0;
    xs[0] = px[0]; xs[1] = px[1]; xs[2] = px[2];
    float xc0 = xs[0], xc1 = xs[1], xc2 = xs[2];
    if (active)
      for (int i = i0; i < i1; i++) {
        const float xn0 = px[st], xn1 = px[st + 1], xn2 = px[st + 2];
        w.s[0] += pg[0]; w.s[1] += pg[1]; w.s[2] += pg[2];
        const float e0 = xn0 - xc0, e1 = xn1 - xc1, e2 = xn2 - xc2;
        w.m[0] = fmaf(-e1, w.s[2], fmaf(e2, w.s[1], w.m[0]));
        w.m[1] = fmaf(-e2, w.s[0], fmaf(e0, w.s[2], w.m[1]));
        w.m[2] = fmaf(-e0, w.s[1], fmaf(e1, w.s[0], w.m[2]));
        xc0 = xn0; xc1 = xn1; xc2 = xn2;
        px += st; pg += st;
      }
    xe[0] = xc0; xe[1] = xc1; xe[2] = xc2;
  }

  // ---- inclusive scan over the side's threads (segments of W lanes, then warps through shared memory) -------------
  WrenchF inc = w;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    if (d < W) {
      WrenchF up;
      float pa[3];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        up.s[i] = __shfl_up_sync(0xffffffffu, inc.s[i], d);
        up.m[i] = __shfl_up_sync(0xffffffffu, inc.m[i], d);
        pa[i] = __shfl_up_sync(0xffffffffu, xe[i], d);
      }
      if (qs >= d) wrench_append(inc, up, pa, xe);
    }
  }
  WrenchF ex;   // exclusive prefix inside the segment, about xs
#pragma unroll
  for (int i = 0; i < 3; i++) {
    ex.s[i] = __shfl_up_sync(0xffffffffu, inc.s[i], 1);
    ex.m[i] = __shfl_up_sync(0xffffffffu, inc.m[i], 1);
    if (qs == 0) { ex.s[i] = 0.f; ex.m[i] = 0.f; }
  }
  if (lane == 31) {
#pragma unroll
    for (int i = 0; i < 3; i++) { wagg[g][wf][i] = inc.s[i]; wagg[g][wf][3 + i] = inc.m[i]; wagg[g][wf][6 + i] = xe[i]; }
  }
  __syncthreads();
  if (WPF > 2 || (WPF == 2 && TL == 0)) {
    // earlier warps of the same side (left warps are 0 .. TL/32-1, right warps follow)
    const int wfirst = side ? TL / 32 : 0;
    WrenchF acc{{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    float pacc[3] = {0.f, 0.f, 0.f};
    for (int wq = wfirst; wq < wf; wq++) {
      WrenchF c{{wagg[g][wq][0], wagg[g][wq][1], wagg[g][wq][2]}, {wagg[g][wq][3], wagg[g][wq][4], wagg[g][wq][5]}};
      const float pc[3] = {wagg[g][wq][6], wagg[g][wq][7], wagg[g][wq][8]};
      wrench_append(c, acc, pacc, pc);
      acc = c;
      pacc[0] = pc[0]; pacc[1] = pc[1]; pacc[2] = pc[2];
    }
    if (wf > wfirst) {
      wrench_append(ex, acc, pacc, xs);
      wrench_append(inc, acc, pacc, xe);
    }
  }
  if (extras) {
    if (q == TS - 1) {   // last thread of the side: total wrench of the side about x(a_steps) = x_{dr0}
#pragma unroll
      for (int i = 0; i < 3; i++) { side_tot[g][side][i] = inc.s[i]; side_tot[g][side][3 + i] = inc.m[i]; }
    }
    __syncthreads();
  }

  // ---- pass 2: walk the chunk again from the prefix, emit the terms of every cut ----------------------------------
  const bool wantD = p.grad_dihedrals != nullptr, wantA = GEN && p.grad_angles != nullptr, wantL = GEN && p.grad_lengths != nullptr;
  if (active && i0 < i1) {
    float ft[3] = {0.f, 0.f, 0.f}, mt[3] = {0.f, 0.f, 0.f}, xr[3] = {0.f, 0.f, 0.f};
    const bool lext = extras && side == 0;
    if (lext) {
#pragma unroll
      for (int i = 0; i < 3; i++) {
        xr[i] = sX[3 * p.dr0 + i];
        ft[i] = side_tot[g][0][i] + side_tot[g][1][i] + sG[3 * p.dr0 + i];   // g of atom dr0 belongs to neither end
        mt[i] = side_tot[g][0][3 + i] + side_tot[g][1][3 + i];
      }
    }
    float plc = (float)pl_ex.c, pls = (float)pl_ex.s, plx = (float)pl_ex.x, ply = (float)pl_ex.y;
    float llx = 0.f, lly = 0.f;   // local planar position of atom i (each half of the chunk starts at its own origin)
    float s0 = ex.s[0], s1 = ex.s[1], s2 = ex.s[2], m0 = ex.m[0], m1 = ex.m[1], m2 = ex.m[2];
    const float* px = sX + 3 * a0 + 2 * st;   // x(a_{i+2}); one atom beyond the chain end is slack / the neighbouring row
    float* pg = sG + 3 * a0;
    float xn0 = px[-st], xn1 = px[-st + 1], xn2 = px[-st + 2];
    float e10 = xn0 - xs[0], e11 = xn1 - xs[1], e12 = xn2 - xs[2];
#pragma unroll 2
    for (int i = i0; i < i1; i++) {
      const float xq0 = px[0], xq1 = px[1], xq2 = px[2];
      s0 += pg[0]; s1 += pg[1]; s2 += pg[2];
      m0 = fmaf(-e11, s2, fmaf(e12, s1, m0));
      m1 = fmaf(-e12, s0, fmaf(e10, s2, m1));
      m2 = fmaf(-e10, s1, fmaf(e11, s0, m2));
      const float e20 = xq0 - xn0, e21 = xq1 - xn1, e22 = xq2 - xn2;
      if (wantD) pg[0] = -(e20 * m0 + e21 * m1 + e22 * m2) * rsqrt_fast(fmaf(e20, e20, fmaf(e21, e21, e22 * e22)));
      if (GEN) {
        float rA = 0.f, rL = 0.f;
        if (wantA) {
          if (planar) {
            const int k = n - 2 - i;                      // cut index (planar mode has only the right side)
            rA = ((k - 1) & 1) ? -m2 : m2;                // hinge normal -(-1)^(k-1) e_z, term = -<normal, M>
          } else {
            const float c0 = e11 * e22 - e12 * e21, c1 = e12 * e20 - e10 * e22, c2 = e10 * e21 - e11 * e20;
            rA = (c0 * m0 + c1 * m1 + c2 * m2) * rsqrt_fast(fmaf(c0, c0, fmaf(c1, c1, c2 * c2)));
          }
        }
        if (wantL) rL = -(e10 * s0 + e11 * s1 + e12 * s2) * rsqrt_fast(fmaf(e10, e10, fmaf(e11, e11, e12 * e12)));
        if (lext) {
          if (i == imid) {   // second half of the chunk: its own prefix and origin
            plc = (float)pl_ex2.c; pls = (float)pl_ex2.s; plx = (float)pl_ex2.x; ply = (float)pl_ex2.y;
            llx = 0.f; lly = 0.f;
          }
          // planar atom k+1 (k = i) from its local coordinates; planar direction of bond k from two of them
          const float lx = sL[i], ly = sA[i];
          const float cn0 = fmaf(plc, lx, fmaf(-pls, ly, plx)), cn1 = fmaf(pls, lx, fmaf(plc, ly, ply));
          // z-torque of the whole molecule about planar atom k+1
          const float tz = mt[2] + ((xr[0] - cn0) * ft[1] - (xr[1] - cn1) * ft[0]);
          rA += (i & 1) ? -tz : tz;
          if (wantL) {
            const float dlx = lx - llx, dly = ly - lly;
            const float inv = rsqrt_fast(fmaf(dlx, dlx, dly * dly));
            const float q0 = dlx * inv, q1 = dly * inv;
            rL += (plc * q0 - pls * q1) * ft[0] + (pls * q0 + plc * q1) * ft[1];
          }
          llx = lx; lly = ly;
        }
        pg[1] = rA; pg[2] = rL;
      }
      xn0 = xq0; xn1 = xq1; xn2 = xq2;
      e10 = e20; e11 = e21; e12 = e22;
      px += st; pg += st;
    }
  }
  __syncthreads();

  // ---- gather: the term of cut k sits in slot k (left, k < dr0) or k+1 (right) ---------------------------------------
  if (active) {
    const int dr0 = p.dr0;
    if (wantD) {
      float* gD = p.grad_dihedrals + fr * (int64_t)(n - 3);
      for (int d = t; d < n - 3; d += T) gD[d] = sG[3 * (d < dr0 ? d : d + 3)];            // left: cut d, right: cut d+2
    }
    if (wantA) {
      float* gA = p.grad_angles + fr * (int64_t)(n - 2);
      for (int j = t; j < n - 2; j += T) gA[j] = sG[3 * (j < dr0 ? j : j + 2) + 1];        // left: cut j, right: cut j+1
    }
    if (wantL) {
      float* gL = p.grad_lengths + fr * (int64_t)(n - 1);
      for (int k = t; k < n - 1; k += T) gL[k] = sG[3 * (k < dr0 ? k : k + 1) + 2];        // cut k
    }
  }
}

// ====================================================================================================
// dihedrals_to_cartesian: gradient w.r.t. the START chain (the models never need it -- BackMapLayer carries the
// gradient of the fused composition -- but the reference's operator is differentiable in both arguments).
//
// Twisting about bonds keeps every bond length and bond angle and shifts every dihedral by its twist, so the build is
//     out_k = place(out_{k-3}, out_{k-2}, out_{k-1};  L_{k-1}(start), theta_{k-2}(start), delta_{k-3}(start) + twist_{k-3}),   out_{0,1,2} = start_{0,1,2}
// (per side, in the side's chain order).  Its reverse mode is one walk from the chain end to the anchor: the adjoint of
// out_k is pushed through `place` onto the three previous atoms and onto (L, theta, Delta), and those three scalars are
// pushed through the internal coordinates of the START window (start_{k-3..k}).  Everything `place` needs is read off the
// final coordinates.  Sequential along the chain: one thread per frame, float64, sliding windows in registers.
// tests/test_oracle_kats.py::test_start_chain_vjp_prototype states the same algorithm in numpy (1e-15 from autograd).
// ====================================================================================================
struct D3 {
  double x, y, z;
};
__device__ __forceinline__ D3 d3(double x, double y, double z) { return D3{x, y, z}; }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return D3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return D3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ D3 operator*(double s, D3 a) { return D3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ double dot3(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 cross3(D3 a, D3 b) { return D3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ D3 ldf3(const float* p) { return D3{(double)p[0], (double)p[1], (double)p[2]}; }

struct FrameD {
  D3 u, n, m, ab;
  double lbc, ln;
};
// u = unit(c - b), n = unit((b - a) x u), m = n x u
__device__ __forceinline__ FrameD frame_of(D3 a, D3 b, D3 c) {
  FrameD f;
  const D3 bc = c - b;
  f.ab = b - a;
  f.lbc = sqrt(dot3(bc, bc));
  f.u = (1.0 / f.lbc) * bc;
  const D3 nraw = cross3(f.ab, f.u);
  f.ln = sqrt(dot3(nraw, nraw));
  f.n = (1.0 / f.ln) * nraw;
  f.m = cross3(f.n, f.u);
  return f;
}
// adjoint of frame_of: (ubar, nbar, mbar) -> (abar, bbar, cbar)
__device__ __forceinline__ void frame_vjp(const FrameD& f, D3 ub, D3 nb, D3 mb, D3* ab_, D3* bb_, D3* cb_) {
  nb = nb + cross3(f.u, mb);
  ub = ub + cross3(mb, f.n);
  const D3 nrawb = (1.0 / f.ln) * (nb - dot3(f.n, nb) * f.n);
  const D3 abb = cross3(f.u, nrawb);
  ub = ub + cross3(nrawb, f.ab);
  const D3 bcb = (1.0 / f.lbc) * (ub - dot3(f.u, ub) * f.u);
  *ab_ = -1.0 * abb;
  *bb_ = abb - bcb;
  *cb_ = bcb;
}

// one side.  Chain position p is atom first + dir * p; `len` positions.  gscale0 = 0 drops the upstream gradient of the
// first three positions (the right build's anchor outputs are discarded by the reference's concat); add_first3 adds
// to grad_chain there instead of storing (the two sides share the three middle atoms).
__device__ void chain_bwd_side(const float* __restrict__ start, const float* __restrict__ fin, const float* __restrict__ g,
                               float* __restrict__ out, int first, int dir, int len, double gscale0, bool add_first3) {
  auto at = [&](int pidx) { return 3 * (first + dir * pidx); };
  D3 P0 = d3(0, 0, 0), P1 = P0, P2 = P0;   // propagated adjoints of final points k, k-1, k-2
  D3 S0 = P0, S1 = P0, S2 = P0;             // start adjoints of points k, k-1, k-2 (from the windows already visited)
  for (int k = len - 1; k >= 3; k--) {
    const D3 fa = ldf3(fin + at(k - 3)), fb = ldf3(fin + at(k - 2)), fc = ldf3(fin + at(k - 1)), fd = ldf3(fin + at(k));
    const D3 db = ldf3(g + at(k)) + P0;
    // ---- adjoint of place() at the final points
    const FrameD f = frame_of(fa, fb, fc);
    const D3 w = fd - fc;
    const double L = sqrt(dot3(w, w));
    const double ct = -dot3(w, f.u) / L, st = sqrt(fmax(0.0, 1.0 - ct * ct));
    const double xx = dot3(w, f.m), yy = dot3(w, f.n);
    const double cd = xx / (L * st), sd = yy / (L * st);
    const double Lb = dot3(db, w) / L;
    const double tb = L * dot3(db, st * f.u + ct * (cd * f.m + sd * f.n));
    const double Db = L * st * dot3(db, cd * f.n - sd * f.m);
    D3 ab_, bb_, cb_;
    frame_vjp(f, (-L * ct) * db, (L * st * sd) * db, (L * st * cd) * db, &ab_, &bb_, &cb_);
    P0 = P1 + cb_ + db;
    P1 = P2 + bb_;
    P2 = ab_;
    // ---- adjoint of (L, theta, Delta) of the START window
    const D3 sa = ldf3(start + at(k - 3)), sb = ldf3(start + at(k - 2)), sc = ldf3(start + at(k - 1)), sdp = ldf3(start + at(k));
    const FrameD fs = frame_of(sa, sb, sc);
    const D3 ws = sdp - sc;
    const double Ls = sqrt(dot3(ws, ws));
    const double wu = dot3(ws, fs.u);
    const double q = -wu / Ls, sts = sqrt(fmax(1e-300, 1.0 - q * q));
    const double xs_ = dot3(ws, fs.m), ys_ = dot3(ws, fs.n);
    const double qb = -tb / sts;
    const double r2 = xs_ * xs_ + ys_ * ys_;
    const double xb = -ys_ / r2 * Db, yb = xs_ / r2 * Db;
    const D3 wb = (Lb / Ls) * ws + qb * ((-1.0 / Ls) * fs.u + (wu / (Ls * Ls * Ls)) * ws) + xb * fs.m + yb * fs.n;
    D3 a2, b2, c2;
    frame_vjp(fs, (-qb / Ls) * ws, yb * ws, xb * ws, &a2, &b2, &c2);
    const D3 outk = S0 + wb;
    float* o = out + at(k);
    o[0] = (float)outk.x; o[1] = (float)outk.y; o[2] = (float)outk.z;
    S0 = S1 + (c2 - wb);
    S1 = S2 + b2;
    S2 = a2;
  }
  // the three anchor positions of the side (positions 2, 1, 0 hold P0/S0, P1/S1, P2/S2 when len > 3; shorter chains
  // never entered the loop and all six are zero)
  const D3 Pk[3] = {P0, P1, P2}, Sk[3] = {S0, S1, S2};
  for (int j = 0; j < 3 && 2 - j < len; j++) {
    const int k = 2 - j;
    const D3 v = Sk[j] + Pk[j] + gscale0 * ldf3(g + at(k));
    float* o = out + at(k);
    if (add_first3) { o[0] += (float)v.x; o[1] += (float)v.y; o[2] += (float)v.z; }
    else { o[0] = (float)v.x; o[1] = (float)v.y; o[2] = (float)v.z; }
  }
}

__global__ void __launch_bounds__(128) d2c_chain_bwd_kernel(const float* __restrict__ chain, int64_t cstride, const float* __restrict__ xyz,
                                                           const float* __restrict__ grad_xyz, int64_t b, int n, int one_way,
                                                           float* __restrict__ grad_chain) {
  const int64_t frame = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (frame >= b) return;
  const float* st = chain + frame * cstride;
  const float* fin = xyz + frame * (int64_t)(3 * n);
  const float* g = grad_xyz + frame * (int64_t)(3 * n);
  float* out = grad_chain + frame * (int64_t)(3 * n);
  if (one_way) {
    chain_bwd_side(st, fin, g, out, 0, 1, n, 1.0, false);
  } else {
    const int s = n / 2;
    chain_bwd_side(st, fin, g, out, s + 1, -1, s + 2, 1.0, false);          // atoms s+1 .. 0
    chain_bwd_side(st, fin, g, out, s - 1, 1, n - s + 1, 0.0, true);        // atoms s-1 .. n-1; its first three outputs are dropped
  }
}

int d2c_chain_bwd_device(const float* chain, int64_t cstride, const float* xyz, const float* grad_xyz, int64_t b, int64_t n, int one_way,
                         float* grad_chain, cudaStream_t st) {
  EMK_REQUIRE(b == 0 || (chain && xyz && grad_xyz && grad_chain), EMK_E_NULL, "emk_dihedrals_to_cartesian_chain_bwd: NULL pointer argument");
  EMK_REQUIRE(n >= 4 && n < (1 << 20), EMK_E_SHAPE, "emk_dihedrals_to_cartesian_chain_bwd: need 4 <= n_atoms < 2^20");
  EMK_REQUIRE(b >= 0 && (cstride == 0 || cstride == 3 * n), EMK_E_ARG, "emk_dihedrals_to_cartesian_chain_bwd: chain_batch_stride must be 0 or 3*n_atoms");
  if (b == 0) return EMK_OK;
  d2c_chain_bwd_kernel<<<(unsigned)((b + 127) / 128), 128, 0, st>>>(chain, cstride, xyz, grad_xyz, b, (int)n, one_way, grad_chain);
  return launch_status("d2c_chain_bwd_kernel");
}

// ====================================================================================================
// host launchers
// ====================================================================================================
static int pick_warps(size_t floats_per_warp, size_t shared_floats, int* warps, size_t* smem_bytes) {
  const size_t budget = 200 * 1024;
  size_t per = floats_per_warp * sizeof(float);
  size_t sh = shared_floats * sizeof(float);
  if (per + sh > 227 * 1024 - 1024) return fail(EMK_E_UNSUPPORTED, "chain too long for one warp's shared-memory staging (%zu bytes)", per + sh);
  int w = (int)((budget - sh) / per);
  if (w < 1) w = 1;
  if (w > 8) w = 8;
  *warps = w;
  *smem_bytes = (size_t)w * per + sh;
  return EMK_OK;
}

template <typename K>
static int set_smem(K kern, size_t bytes) {
  EMK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
  (void)bytes;
  return EMK_OK;
}

static int get_sincos_table(const double2** out) {
  static double2* tables[kMaxDevices] = {nullptr};
  static std::mutex mu;
  int dev = 0;
  EMK_CUDA(cudaGetDevice(&dev));
  EMK_REQUIRE(dev >= 0 && dev < kMaxDevices, EMK_E_UNSUPPORTED, "device ordinal %d out of range", dev);
  std::lock_guard<std::mutex> lock(mu);
  if (!tables[dev]) {
    double2* t = nullptr;
    EMK_CUDA(cudaMalloc(&t, SC_TABLE * sizeof(double2)));
    sincos_table_kernel<<<SC_TABLE / 256, 256>>>(t);
    int rc = launch_status("sincos_table_kernel");
    if (rc) return rc;
    EMK_CUDA(cudaDeviceSynchronize());
    tables[dev] = t;
  }
  *out = tables[dev];
  return EMK_OK;
}

// version 6 (lane per frame and side): shared bond lengths, n % 4 == 0 (16-byte aligned 8-atom output groups), 16-byte aligned
// output, bond lengths + per-warp tiles within the shared memory of one SM, and a batch that fills the machine
static std::atomic<int64_t> g_fwd6_warps{0};
int64_t fwd6_warps() { return g_fwd6_warps.load(); }
void set_fwd6_warps(int64_t v) { g_fwd6_warps.store(v); }
// extent limit (nm) of the float32 first pass; 0 (default) = float64 chain only (emk_set_option("backmap_fwd6_f32_extent_nm", v)).
// OFF by default, by measurement (tools/experiments/fwd6_extent_sweep.py, 65 536 x 1 500): the float32 pass takes 0.634 ms
// against 0.661 ms for the float64 chain when no tile falls back -- the kernel is bound by the LATENCY of the dependent
// chain of a step at 16 warps per SM, not by issue slots, so halving the issue slots buys 4 % (and more resident warps do
// not help either: 20 warps are 30 % slower in both precisions, 24 float32 warps at 80 registers 2.5x slower:
// profiles/r02_backmap_fwd_f32_experiment.txt) -- and with the 16 nm limit
// that the 1e-4 nm tolerance needs, the 0.06 % of frames beyond it make 2 % of the tiles run twice, which costs 33 %
// through the makespan of the persistent pairs (0.843 ms).
static std::atomic<int64_t> g_fwd6_f32_extent{0};
int64_t fwd6_f32_extent() { return g_fwd6_f32_extent.load(); }
void set_fwd6_f32_extent(int64_t v) { g_fwd6_f32_extent.store(v < 0 ? 0 : v); }
static size_t fwd6_smem_bytes(int64_t n, int warps, bool f32) {
  const int64_t extra = f32 ? SC6 * 8 * 4 + ((n - 1 + 31) / 32) * 32 : 0;   // hi/lo table + float32 bond lengths
  return (size_t)(SC6 * 8 * 4 + ((2 * (n - 1) + 31) / 32) * 32 + extra + warps * F6_WARP_WORDS) * sizeof(float);
}
int encode_f32_map_2d(void* map_out, const float* base, uint64_t cols, uint64_t rows, uint64_t row_pitch_bytes, uint32_t box_cols,
                      uint32_t box_rows, int swizzle_128b);   // pair_tile.cu
// batch size from which the lane-per-frame kernel takes over (emk_set_option("backmap_fwd6_min_batch", v); 0 = whenever
// eligible, negative = never).  Default from EMK_BACKMAP_FWD6_MIN_BATCH or the measured cross-over.
static std::atomic<int64_t> g_fwd6_min_batch{[] {
  const char* e = getenv("EMK_BACKMAP_FWD6_MIN_BATCH");
  return e ? (int64_t)atoll(e) : (int64_t)4096;
}()};
int64_t fwd6_min_batch() { return g_fwd6_min_batch.load(); }
void set_fwd6_min_batch(int64_t v) { g_fwd6_min_batch.store(v); }

template <bool A8, int W, bool F32>
static int fwd6_launch_w(const CUtensorMap& omap, const float* lengths, const float* angles, const float* dihedrals, int64_t b, int64_t n,
                         float* xyz, const double2* tab, float ext2_limit, cudaStream_t st) {
  auto kern = backmap_fwd6_kernel<A8, W, F32>;
  static bool cfg[kMaxDevices] = {false};
  if (first_use_on_device(cfg)) EMK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const int64_t tiles = (b + 31) / 32;
  const int64_t blocks = std::min<int64_t>((tiles + W / 2 - 1) / (W / 2), (int64_t)sm_count());
  kern<<<(unsigned)blocks, 32 * W, fwd6_smem_bytes(n, W, F32), st>>>(omap, lengths, angles, dihedrals, b, (int)n, xyz, tab, ext2_limit);
  return launch_status("backmap_fwd6_kernel");
}
static int backmap_fwd6_launch(const float* lengths, const float* angles, const float* dihedrals, int64_t b, int64_t n, float* xyz,
                               const double2* tab, int warps, cudaStream_t st) {
  // output viewed as (b, 3 n) float32 rows: a warp stores 8 atoms (24 floats) of its 32 frames with one tensor copy
  alignas(64) CUtensorMap omap;
  int rc = encode_f32_map_2d(&omap, xyz, (uint64_t)(3 * n), (uint64_t)b, (uint64_t)(12 * n), 24, 32, 0);
  if (rc) return rc;
  // n % 4 == 0 makes the rows of `angles` (n - 2 floats) 8-byte aligned whenever the base is: 8-byte cp.async for that array
  const bool a8 = (reinterpret_cast<uintptr_t>(angles) & 7) == 0;
  const bool f32 = fwd6_f32_extent() > 0;
  const float lim = (float)fwd6_f32_extent();
  const float ext2_limit = lim * lim;
#define EMK_F6(W)                                                                                                              \
  do {                                                                                                                         \
    if (f32) return a8 ? fwd6_launch_w<true, W, true>(omap, lengths, angles, dihedrals, b, n, xyz, tab, ext2_limit, st)        \
                       : fwd6_launch_w<false, W, true>(omap, lengths, angles, dihedrals, b, n, xyz, tab, ext2_limit, st);      \
    return a8 ? fwd6_launch_w<true, W, false>(omap, lengths, angles, dihedrals, b, n, xyz, tab, ext2_limit, st)                \
              : fwd6_launch_w<false, W, false>(omap, lengths, angles, dihedrals, b, n, xyz, tab, ext2_limit, st);              \
  } while (0)
  if (warps <= 8) EMK_F6(8);
  if (warps <= 12) EMK_F6(12);
  if (warps <= 14) EMK_F6(14);
  if (warps <= 16) EMK_F6(16);
  if (warps <= 18) EMK_F6(18);
  EMK_F6(20);
#undef EMK_F6
}

int backmap_fwd_device(const float* lengths, int64_t lstride, const float* angles, const float* dihedrals, int64_t b,
                       int64_t n, float* xyz, cudaStream_t st) {
  EMK_REQUIRE(b == 0 || (lengths && angles && dihedrals && xyz), EMK_E_NULL, "emk_backmap: NULL pointer argument");
  EMK_REQUIRE(n >= 4 && n < (1 << 20), EMK_E_SHAPE, "emk_backmap: need 4 <= n_atoms < 2^20, got %lld", (long long)n);
  EMK_REQUIRE(b >= 0 && (lstride == 0 || lstride == n - 1), EMK_E_ARG, "emk_backmap: lengths_batch_stride must be 0 or n_atoms-1");
  if (b == 0) return EMK_OK;
  const double2* tab;
  int rc = get_sincos_table(&tab);
  if (rc) return rc;
  if (lstride == 0 && n % 4 == 0 && n >= 16 && (reinterpret_cast<uintptr_t>(xyz) & 15) == 0 && fwd6_min_batch() >= 0 && b >= fwd6_min_batch()) {
    // warps per CTA: the candidate whose waves are fullest (time ~ rounds x pairs: the per-SM rate does not depend on the count)
    const int64_t tiles = (b + 31) / 32, sms = sm_count();
    static const int cand[] = {20, 18, 16, 14, 12, 8};
    int best = 0;
    double best_eff = 0.0;
    for (int w : cand) {
      if (fwd6_smem_bytes(n, w, fwd6_f32_extent() > 0) > 227 * 1024) continue;   // long chains: the bond lengths take the room of warps
      if (fwd6_warps() != 0 && w != fwd6_warps()) continue;
      const int64_t cap = sms * (w / 2);
      const double eff = (double)tiles / (double)(((tiles + cap - 1) / cap) * cap) * (w == 8 ? 0.85 : 1.0);   // 8 warps hide less latency
      if (eff > best_eff + 1e-9) { best_eff = eff; best = w; }
    }
    // the chunk-scan kernel reaches ~0.8 of this kernel's full-wave rate at any batch size
    if (best != 0 && (best_eff >= 0.8 || fwd6_min_batch() == 0))
      return backmap_fwd6_launch(lengths, angles, dihedrals, b, n, xyz, tab, best, st);
  }
  const size_t per_group = ((3 * (size_t)n + 7) & ~(size_t)3) * sizeof(float);
  const size_t budget = 224 * 1024 - SC_SMALL * sizeof(double2);
  const int groups = (int)std::min<size_t>(FWD5_GROUPS, budget / per_group);   // frames in flight per CTA
  EMK_REQUIRE(groups >= 1, EMK_E_UNSUPPORTED, "emk_backmap: chain of %lld atoms needs %zu bytes of staging shared memory", (long long)n, per_group);
  const size_t smem = SC_SMALL * sizeof(double2) + (size_t)groups * per_group;
  static bool cfg[kMaxDevices] = {false};
  if (first_use_on_device(cfg)) EMK_CUDA(cudaFuncSetAttribute(backmap_fwd5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  // persistent CTAs: as many as can be resident (2 per SM at 500 residues), each group strides over the frames
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(2, (227 * 1024) / (smem + 2048)));
  const int64_t want = (b + groups - 1) / groups;
  const int64_t blocks = std::min<int64_t>(want, (int64_t)sm_count() * per_sm);
  backmap_fwd5_kernel<<<(unsigned)blocks, 64 * groups, smem, st>>>(lengths, lstride, angles, dihedrals, b, (int)n, xyz, tab);
  return launch_status("backmap_fwd5_kernel");
}

int chain_in_plane_device(const float* lengths, int64_t lstride, const float* angles, int64_t b, int64_t n, float* xyz, cudaStream_t st) {
  EMK_REQUIRE(b == 0 || (lengths && angles && xyz), EMK_E_NULL, "emk_chain_in_plane: NULL pointer argument");
  EMK_REQUIRE(n >= 3 && n < (1 << 20), EMK_E_SHAPE, "emk_chain_in_plane: need 3 <= n_atoms < 2^20, got %lld", (long long)n);
  EMK_REQUIRE(b >= 0 && (lstride == 0 || lstride == n - 1), EMK_E_ARG, "emk_chain_in_plane: lengths_batch_stride must be 0 or n_atoms-1");
  if (b == 0) return EMK_OK;
  int warps;
  size_t smem;
  int rc = pick_warps((size_t)(n - 1) + (n - 2) + 3 * n, 0, &warps, &smem);
  if (rc) return rc;
  static bool cfg[kMaxDevices] = {false};
  if (first_use_on_device(cfg)) { rc = set_smem(chain_in_plane_kernel, smem); if (rc) return rc; }
  const int64_t blocks = (b + warps - 1) / warps;
  const double2* tab;
  rc = get_sincos_table(&tab);
  if (rc) return rc;
  chain_in_plane_kernel<<<(unsigned)blocks, warps * 32, smem, st>>>(lengths, lstride, angles, b, (int)n, xyz, tab);
  return launch_status("chain_in_plane_kernel");
}

int d2c_general_device(const float* dihedrals, const float* chain, int64_t cstride, int64_t b, int64_t n, int one_way,
                       float* xyz, cudaStream_t st) {
  EMK_REQUIRE(b == 0 || (dihedrals && chain && xyz), EMK_E_NULL, "emk_dihedrals_to_cartesian: NULL pointer argument");
  EMK_REQUIRE(n >= 4 && n < (1 << 20), EMK_E_SHAPE, "emk_dihedrals_to_cartesian: need 4 <= n_atoms < 2^20, got %lld", (long long)n);
  EMK_REQUIRE(b >= 0 && (cstride == 0 || cstride == 3 * n), EMK_E_ARG, "emk_dihedrals_to_cartesian: chain_batch_stride must be 0 or 3*n_atoms");
  if (b == 0) return EMK_OK;
  int warps;
  size_t smem;
  int rc = pick_warps((size_t)(n - 3) + 3 * n, 0, &warps, &smem);
  if (rc) return rc;
  static bool cfg[kMaxDevices] = {false};
  if (first_use_on_device(cfg)) { rc = set_smem(d2c_general_kernel, smem); if (rc) return rc; }
  const int64_t blocks = (b + warps - 1) / warps;
  const double2* tab;
  rc = get_sincos_table(&tab);
  if (rc) return rc;
  d2c_general_kernel<<<(unsigned)blocks, warps * 32, smem, st>>>(dihedrals, chain, cstride, b, (int)n, one_way, xyz, tab);
  return launch_status("d2c_general_kernel");
}

template <int T, bool GEN>
static int launch_bwd3(const BwdParams& p, const double2* tab, bool need_planar, cudaStream_t st) {
  constexpr int CTA = T < 128 ? 128 : T;
  constexpr int FPC = CTA / T;
  const size_t n = (size_t)p.n;
  const size_t nL = p.planar ? n - 1 : (size_t)p.dr0, nA = p.planar ? n - 2 : (size_t)p.dr0;
  const size_t per_frame = ((3 * n + 27) & ~(size_t)3) + ((3 * n + 7) & ~(size_t)3) +
                           (need_planar ? ((nL + 7) & ~(size_t)3) + ((nA + 7) & ~(size_t)3) : 0) +
                           (need_planar && !p.planar ? 8 * (size_t)(T / 2) : 0);
  const size_t smem = FPC * per_frame * sizeof(float);
  EMK_REQUIRE(smem <= 200 * 1024, EMK_E_UNSUPPORTED, "back-mapping backward: chain of %d atoms needs %zu bytes of staging shared memory", p.n, smem);
  auto kern = backmap_bwd3_kernel<T, GEN>;
  static bool cfg[kMaxDevices] = {false};
  if (first_use_on_device(cfg)) EMK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int64_t blocks = (p.b + FPC - 1) / FPC;
  kern<<<(unsigned)blocks, CTA, smem, st>>>(p, tab);
  return launch_status("backmap_bwd3_kernel");
}

int backmap_bwd_device(const BwdParams& p, cudaStream_t st) {
  if (p.b == 0) return EMK_OK;
  EMK_REQUIRE(p.mid == 0 || p.mid - 1 == p.dr0, EMK_E_ARG, "back-mapping backward: inconsistent split (mid=%d, dr0=%d)", p.mid, p.dr0);
  const double2* tab;
  int rc = get_sincos_table(&tab);
  if (rc) return rc;
  const bool need_planar = p.planar || ((p.grad_angles || p.grad_lengths) && p.mid > 1);
  const int n = p.n;
  const bool gen = p.planar || p.grad_angles || p.grad_lengths;
#define EMK_BWD3(TT) return gen ? launch_bwd3<TT, true>(p, tab, need_planar, st) : launch_bwd3<TT, false>(p, tab, need_planar, st)
  if (n <= 32 * BWD_CA) EMK_BWD3(32);
  if (n <= 64 * BWD_CA) EMK_BWD3(64);
  if (n <= 128 * BWD_CA) EMK_BWD3(128);
  if (n <= 256 * BWD_CA) EMK_BWD3(256);
  if (n <= 512 * BWD_CA) EMK_BWD3(512);
  EMK_REQUIRE(n <= 1024 * BWD_CA, EMK_E_UNSUPPORTED, "back-mapping backward: chains above %d atoms are not supported", 1024 * BWD_CA);
  EMK_BWD3(1024);
#undef EMK_BWD3
}

}  // namespace emk
