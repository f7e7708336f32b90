// Bandwidth-bound helpers of the hot path: periodic_distance, sigmoid, PeriodicInput, rotation_matrix,
// column mean, and the batched small-d pairwise_dist (PairwiseDistances layer) forward / backward.
#include <algorithm>

#include "emk_common.cuh"

namespace emk {

static inline int grid_for(int64_t work, int threads = 256) {
  int64_t blocks = (work + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---- pairwise_dist_periodic backward -------------------------------------------------------------------
// Reference autodiff of distances.py:164-175 (TF conventions): Dist_ij = sqrt(sum_k V^2) + 1e-12,
// V = min(d, P - d) (+1e-12 where exactly 0), d = |x_jk - x_ik|; abs' = sign, minimum routes to its first operand on
// ties.  Both orders of a pair touch row i:
//     grad_x[i,k] = - sum_j (G_ij + G_ji) * V_ijk / S_ij * m_ijk * sign(x_jk - x_ik),   S = Dist - 1e-12, m = +1 if d <= P - d else -1.
// One CTA owns PDB_ROWS rows i and 128 columns k; the pair weights w_ij = (G_ij + G_ji) / S_ij of a block of 128 rows j
// are staged in shared memory, x_j is read coalesced.  O(N^2 D); the models never need it (the high-d side is input
// data), it exists so that the operator is differentiable like the reference's.
constexpr int PDB_ROWS = 8;
__global__ void __launch_bounds__(128) pairwise_periodic_bwd_kernel(const float* __restrict__ x, int64_t n, int64_t d, float P,
                                                                    const float* __restrict__ dist, const float* __restrict__ go,
                                                                    float* __restrict__ gx) {
  __shared__ float w[PDB_ROWS][128];
  const int64_t i0 = (int64_t)blockIdx.y * PDB_ROWS;
  const int64_t k = (int64_t)blockIdx.x * 128 + threadIdx.x;
  const bool kv = k < d;
  float xi[PDB_ROWS], acc[PDB_ROWS];
#pragma unroll
  for (int r = 0; r < PDB_ROWS; r++) {
    xi[r] = (kv && i0 + r < n) ? x[(i0 + r) * d + k] : 0.f;
    acc[r] = 0.f;
  }
  for (int64_t j0 = 0; j0 < n; j0 += 128) {
    __syncthreads();
    {
      const int64_t j = j0 + threadIdx.x;
#pragma unroll
      for (int r = 0; r < PDB_ROWS; r++) {
        const int64_t i = i0 + r;
        float v = 0.f;
        if (i < n && j < n) {
          const float sij = fmaxf(dist[i * n + j] - 1e-12f, 1e-30f);
          v = (go[i * n + j] + go[j * n + i]) / sij;
        }
        w[r][threadIdx.x] = v;
      }
    }
    __syncthreads();
    if (kv) {
      const int jn = (int)min((int64_t)128, n - j0);
      for (int jj = 0; jj < jn; jj++) {
        const float xj = x[(j0 + jj) * d + k];
#pragma unroll
        for (int r = 0; r < PDB_ROWS; r++) {
          const float df = xj - xi[r];
          const float ad = fabsf(df), pd = P - ad;
          float v = fminf(ad, pd);
          v = v == 0.f ? 1e-12f : v;
          // m * sign(x_j - x_i): +-1, 0 where the difference is exactly 0
          const float sg = df == 0.f ? 0.f : ((ad <= pd) == (df > 0.f) ? 1.f : -1.f);
          acc[r] = fmaf(-w[r][jj] * v, sg, acc[r]);
        }
      }
    }
  }
  if (kv) {
#pragma unroll
    for (int r = 0; r < PDB_ROWS; r++)
      if (i0 + r < n) gx[(i0 + r) * d + k] = acc[r];
  }
}
int pairwise_periodic_bwd_device(const float* x, int64_t n, int64_t d, double P, const float* dist, const float* go, float* gx,
                                 cudaStream_t st) {
  EMK_REQUIRE(n == 0 || (x && dist && go && gx), EMK_E_NULL, "emk_pairwise_dist_periodic_bwd: NULL pointer argument");
  EMK_REQUIRE(n >= 0 && d >= 0 && P > 0, EMK_E_ARG, "emk_pairwise_dist_periodic_bwd: bad arguments");
  if (n * d == 0) return EMK_OK;
  EMK_REQUIRE((n + PDB_ROWS - 1) / PDB_ROWS <= 65535, EMK_E_UNSUPPORTED, "emk_pairwise_dist_periodic_bwd: more than %d rows", 65535 * PDB_ROWS);
  dim3 grid((unsigned)((d + 127) / 128), (unsigned)((n + PDB_ROWS - 1) / PDB_ROWS));
  pairwise_periodic_bwd_kernel<<<grid, 128, 0, st>>>(x, n, d, std::isinf(P) ? INFINITY : (float)P, dist, go, gx);
  return launch_status("pairwise_periodic_bwd_kernel");
}

// ---- FP32 pipe probe ---------------------------------------------------------------------------------
// Measured denominator of the pair-tile kernel's roofline: register-only FFMA chains (16 independent accumulators
// per thread, 8 CTAs of 256 threads per SM), i.e. what the FP32 pipe of this very device issues per second when
// nothing else is in the way.  Used by bench.py; not part of the hot path.
__global__ void __launch_bounds__(256) fp32_probe_kernel(float* __restrict__ sink, int iters, float a, float b) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = (float)(threadIdx.x + i);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int i = 0; i < 16; i++) acc[i] = fmaf(acc[i], a, b);
  }
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 16; i++) sum += acc[i];
  if (sum == 12345.678f) sink[0] = sum;   // never true: keeps the chains alive
}
int fp32_probe_device(double* lane_instr_per_s) {
  EMK_REQUIRE(lane_instr_per_s, EMK_E_NULL, "emk_probe_fp32: NULL output");
  float* sink = nullptr;
  EMK_CUDA(cudaMalloc(&sink, sizeof(float)));
  cudaEvent_t e0, e1;
  EMK_CUDA(cudaEventCreate(&e0));
  EMK_CUDA(cudaEventCreate(&e1));
  const int iters = 1 << 14, blocks = sm_count() * 8;
  fp32_probe_kernel<<<blocks, 256>>>(sink, 256, 0.999f, 0.001f);   // warm-up
  double best = 0.0;
  int rc = launch_status("fp32_probe_kernel");
  for (int rep = 0; rep < 3 && rc == EMK_OK; rep++) {
    cudaEventRecord(e0);
    fp32_probe_kernel<<<blocks, 256>>>(sink, iters, 0.999f, 0.001f);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { rc = fail(EMK_E_ARG, "emk_probe_fp32: kernel failed"); break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double rate = (double)blocks * 256.0 * (double)iters * 64.0 / (ms * 1e-3);
    if (rate > best) best = rate;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  *lane_instr_per_s = best;
  return rc;
}

// ---- periodic_distance -----------------------------------------------------------------------------
__global__ void periodic_distance_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t count, float P,
                                         float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = fabsf(b[i] - a[i]);
    out[i] = fminf(d, P - d);
  }
}
// TF autodiff conventions: abs' = sign, minimum routes to the first operand on ties
__global__ void periodic_distance_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t count, float P,
                                             const float* __restrict__ go, float* __restrict__ ga, float* __restrict__ gb) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const float diff = b[i] - a[i];
    const float d = fabsf(diff);
    const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
    const float branch = (d <= P - d) ? 1.f : -1.f;
    const float g = go[i] * sgn * branch;
    if (gb) gb[i] = g;
    if (ga) ga[i] = -g;
  }
}

// ---- sigmoid ---------------------------------------------------------------------------------------
__global__ void sigmoid_kernel(const float* __restrict__ r, int64_t count, SigSpec s, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = r[i];
    out[i] = sig_eval<false>(x * x, s, nullptr);
  }
}
__global__ void sigmoid_bwd_kernel(const float* __restrict__ r, int64_t count, SigSpec s, const float* __restrict__ go,
                                   float* __restrict__ gr) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = r[i];
    float w;
    sig_eval<true>(x * x, s, &w);
    // s'(r) = r * (s'(r)/r); at r == 0 the limit is 0 for a > 1 and b*c/sig for a == 1
    float ds = x == 0.f ? ((s.a_int == 1) ? s.dcoef * sqrtf(1.f / s.inv_sig2) : 0.f) : w * x;
    gr[i] = go[i] * ds;
  }
}

// ---- PeriodicInput -----------------------------------------------------------------------------------
// (rows, d) -> (rows, 2d) = [sin x | cos x].  CTAs stride over rows, threads over the columns of a row: no index
// division, and 16-byte loads / stores when d is a multiple of 4 and the bases are aligned (VEC).
template <bool VEC>
__global__ void __launch_bounds__(256) periodic_input_kernel(const float* __restrict__ x, int64_t rows, int d, float scale, int rescale,
                                                             float* __restrict__ out) {
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* xr = x + r * d;
    float* so = out + r * 2 * d;
    float* co = so + d;
    if (VEC) {
      for (int k = threadIdx.x * 4; k < d; k += blockDim.x * 4) {
        float4 v = *reinterpret_cast<const float4*>(xr + k);
        if (rescale) { v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale; }
        float4 s4, c4;
        sincosf(v.x, &s4.x, &c4.x);
        sincosf(v.y, &s4.y, &c4.y);
        sincosf(v.z, &s4.z, &c4.z);
        sincosf(v.w, &s4.w, &c4.w);
        *reinterpret_cast<float4*>(so + k) = s4;
        *reinterpret_cast<float4*>(co + k) = c4;
      }
    } else {
      for (int k = threadIdx.x; k < d; k += blockDim.x) {
        float v = xr[k];
        if (rescale) v *= scale;
        float sn, cs;
        sincosf(v, &sn, &cs);
        so[k] = sn;
        co[k] = cs;
      }
    }
  }
}
template <bool VEC>
__global__ void __launch_bounds__(256) periodic_input_bwd_kernel(const float* __restrict__ x, int64_t rows, int d, float scale, int rescale,
                                                                 const float* __restrict__ go, float* __restrict__ gx) {
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* xr = x + r * d;
    const float* gs = go + r * 2 * d;
    const float* gc = gs + d;
    float* gr = gx + r * d;
    if (VEC) {
      for (int k = threadIdx.x * 4; k < d; k += blockDim.x * 4) {
        float4 v = *reinterpret_cast<const float4*>(xr + k);
        const float4 a = *reinterpret_cast<const float4*>(gs + k), c = *reinterpret_cast<const float4*>(gc + k);
        if (rescale) { v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale; }
        float sn, cs;
        float4 o;
        sincosf(v.x, &sn, &cs); o.x = a.x * cs - c.x * sn;
        sincosf(v.y, &sn, &cs); o.y = a.y * cs - c.y * sn;
        sincosf(v.z, &sn, &cs); o.z = a.z * cs - c.z * sn;
        sincosf(v.w, &sn, &cs); o.w = a.w * cs - c.w * sn;
        if (rescale) { o.x *= scale; o.y *= scale; o.z *= scale; o.w *= scale; }
        *reinterpret_cast<float4*>(gr + k) = o;
      }
    } else {
      for (int k = threadIdx.x; k < d; k += blockDim.x) {
        float v = xr[k];
        if (rescale) v *= scale;
        float sn, cs;
        sincosf(v, &sn, &cs);
        float g = gs[k] * cs - gc[k] * sn;
        if (rescale) g *= scale;
        gr[k] = g;
      }
    }
  }
}

// ---- rotation_matrix -----------------------------------------------------------------------------------
__global__ void rotation_matrix_kernel(const float* __restrict__ axis, const float* __restrict__ angle, int64_t b,
                                       float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < b; i += (int64_t)gridDim.x * blockDim.x) {
    const float ux = axis[3 * i], uy = axis[3 * i + 1], uz = axis[3 * i + 2];
    float s, c;
    sincosf(angle[i], &s, &c);
    const float oc = 1.f - c;
    float* o = out + 9 * i;
    o[0] = c + oc * ux * ux;      o[1] = oc * ux * uy - s * uz; o[2] = oc * ux * uz + s * uy;
    o[3] = oc * uy * ux + s * uz; o[4] = c + oc * uy * uy;      o[5] = oc * uy * uz - s * ux;
    o[6] = oc * uz * ux - s * uy; o[7] = oc * uz * uy + s * ux; o[8] = c + oc * uz * uz;
  }
}

// ---- column mean: (rows, cols) -> (cols) ------------------------------------------------------------------
constexpr int CM_SPLIT = 64;
__global__ void column_partial_kernel(const float* __restrict__ x, int64_t rows, int64_t cols, double* __restrict__ part) {
  // block: 32 columns x 8 row lanes; grid.x over column tiles, grid.y over CM_SPLIT row slabs
  __shared__ double sm[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int64_t col = (int64_t)blockIdx.x * 32 + cx;
  const int64_t slab = (rows + gridDim.y - 1) / gridDim.y;
  const int64_t r0 = (int64_t)blockIdx.y * slab, r1 = min(rows, r0 + slab);
  double acc = 0.0;
  if (col < cols) {
    // eight independent loads in flight per thread (the one-load-per-iteration loop reached 0.5 of the HBM rate);
    // groups of four are summed in float32 first (relative 1e-7 on bond lengths, averaged away over the column)
    int64_t r = r0 + ry;
    const float* px = x + r * cols + col;
    const int64_t step = 8 * cols;
    for (; r + 56 < r1; r += 64, px += 8 * step) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = __ldg(px + u * step);
      acc += (double)((v[0] + v[1]) + (v[2] + v[3])) + (double)((v[4] + v[5]) + (v[6] + v[7]));
    }
    for (; r < r1; r += 8, px += step) acc += (double)__ldg(px);
  }
  sm[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && col < cols) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) t += sm[i][cx];
    part[(int64_t)blockIdx.y * cols + col] = t;
  }
}
__global__ void column_final_kernel(const double* __restrict__ part, int split, int64_t rows, int64_t cols, float* __restrict__ out) {
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += (int64_t)gridDim.x * blockDim.x) {
    double t = 0.0;
    for (int s = 0; s < split; s++) t += part[(int64_t)s * cols + c];
    out[c] = (float)(t / (double)rows);
  }
}

// ---- batched pairwise_dist for small d (PairwiseDistances layer) ------------------------------------------
// strict-upper-triangle index p -> (i, j), row-major: row i starts at i*(2n-i-1)/2.  32-bit arithmetic with a
// float estimate + fix-up (exact for n < 4096); the 64-bit / double form covers larger n.
__device__ __forceinline__ void triu_decode(int64_t p, int64_t n, int64_t* i_out, int64_t* j_out) {
  if (n < 4096) {
    const int pi = (int)p, ni = (int)n;
    const float nn = (float)(2 * ni - 1);
    int i = (int)((nn - sqrtf(fmaf(nn, nn, -8.f * (float)pi))) * 0.5f);
    i = max(0, min(i, ni - 2));
    while (i > 0 && i * (2 * ni - i - 1) / 2 > pi) --i;
    while ((i + 1) * (2 * ni - i - 2) / 2 <= pi) ++i;
    *i_out = i;
    *j_out = pi - i * (2 * ni - i - 1) / 2 + i + 1;
    return;
  }
  const double nn = (double)(2 * n - 1);
  int64_t i = (int64_t)((nn - sqrt(nn * nn - 8.0 * (double)p)) * 0.5);
  if (i < 0) i = 0;
  while (i > 0 && i * (2 * n - i - 1) / 2 > p) --i;
  while ((i + 1) * (2 * n - i - 2) / 2 <= p) ++i;
  *i_out = i;
  *j_out = p - i * (2 * n - i - 1) / 2 + i + 1;
}

// grid: (ceil(per / 256), b): one frame per blockIdx.y, coalesced output writes, inputs from L1
// (general fallback: non-flat output, wide d, very long chains)
__global__ void pairwise_small_kernel(const float* __restrict__ x, int64_t b, int64_t n, int64_t d, int64_t bstride,
                                      int64_t rstride, int squared, int flat, float* __restrict__ out) {
  const int64_t per = flat ? n * (n - 1) / 2 : n * n;
  for (int64_t bi = blockIdx.y; bi < b; bi += gridDim.y) {
    const float* xb = x + bi * bstride;
    float* ob = out + bi * per;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < per; p += (int64_t)gridDim.x * blockDim.x) {
      int64_t i, j;
      if (flat) triu_decode(p, n, &i, &j);
      else { i = p / n; j = p - i * n; }
      const float* xi = xb + i * rstride;
      const float* xj = xb + j * rstride;
      float s = 0.f;
      for (int64_t k = 0; k < d; k++) {
        const float t = xi[k] - xj[k];
        s = fmaf(t, t, s);
      }
      ob[p] = squared ? s : sqrtf(s);
    }
  }
}

// PairwiseDistances layer, forward: flat upper triangle of 3-d points, up to PWT_MAX_N selected atoms (the Calpha
// selections of the ADC models).  One thread per output element with coalesced 4-byte stores; the pair (i, j) of
// flat index p comes from a table in shared memory that the CTA builds once and reuses for every frame it
// processes (no per-element decode, no divergence).  The table holds BYTE OFFSETS: i * 16 into a float4 copy of the
// coordinates (x_i is one mostly-broadcast LDS.128) and j * 4 into a component-major copy (x_j is three conflict-free
// LDS.32 on consecutive lanes), so an output costs ~20 instructions and ~5 shared-memory wavefronts per warp.
// G frames are staged per barrier pair.
constexpr int PWF_THREADS = 256;
constexpr int PWT_MAX_N = 181;   // table of at most 64 KB / 4

__global__ void __launch_bounds__(PWF_THREADS) pairwise_flat3_tab_kernel(const float* __restrict__ x, int64_t b, int n, int64_t bstride,
                                                                         int64_t rstride, int squared, int G, float* __restrict__ out) {
  extern __shared__ __align__(16) float sx[];   // [G][n] float4 (x, y, z, 0), then [G][3][n] component-major, then per x uint32
  const int per = n * (n - 1) / 2;
  float4* s4 = reinterpret_cast<float4*>(sx);
  float* ssoa = sx + (size_t)G * 4 * n;
  unsigned* tab = reinterpret_cast<unsigned*>(ssoa + (size_t)G * 3 * n);
  const int tid = threadIdx.x;
  {
    const int nn = 2 * n - 1;
    for (int p = tid; p < per; p += PWF_THREADS) {
      int i = (int)(((float)nn - sqrtf((float)(nn * nn - 8 * p))) * 0.5f);
      i = max(0, min(i, n - 2));
      if (i * (nn - i) / 2 > p) --i;
      if ((i + 1) * (nn - i - 1) / 2 <= p) ++i;
      const int j = p - i * (nn - i) / 2 + i + 1;
      tab[p] = (unsigned)(i * 16) | ((unsigned)(j * 4) << 16);
    }
  }
  for (int64_t g0 = (int64_t)blockIdx.x * G; g0 < b; g0 += (int64_t)gridDim.x * G) {
    const int gc = (int)min((int64_t)G, b - g0);
    __syncthreads();
    for (int f = 0; f < gc; f++) {
      const float* xb = x + (g0 + f) * bstride;
      for (int a = tid; a < n; a += PWF_THREADS) {
        const float vx = xb[a * rstride], vy = xb[a * rstride + 1], vz = xb[a * rstride + 2];
        s4[f * n + a] = make_float4(vx, vy, vz, 0.f);
        ssoa[(f * 3 + 0) * n + a] = vx;
        ssoa[(f * 3 + 1) * n + a] = vy;
        ssoa[(f * 3 + 2) * n + a] = vz;
      }
    }
    __syncthreads();
    for (int f = 0; f < gc; f++) {
      const char* a4 = reinterpret_cast<const char*>(s4 + f * n);
      const char* cx = reinterpret_cast<const char*>(ssoa + (f * 3 + 0) * n);
      const char* cy = reinterpret_cast<const char*>(ssoa + (f * 3 + 1) * n);
      const char* cz = reinterpret_cast<const char*>(ssoa + (f * 3 + 2) * n);
      float* ob = out + (g0 + f) * per;
#pragma unroll 4
      for (int p = tid; p < per; p += PWF_THREADS) {
        const unsigned e = tab[p];
        const unsigned oi = e & 0xffffu, oj = e >> 16;
        const float4 pi = *reinterpret_cast<const float4*>(a4 + oi);
        const float dx = pi.x - *reinterpret_cast<const float*>(cx + oj);
        const float dy = pi.y - *reinterpret_cast<const float*>(cy + oj);
        const float dz = pi.z - *reinterpret_cast<const float*>(cz + oj);
        const float s2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        ob[p] = squared ? s2 : (s2 >= EMK_TINY ? s2 * rsqrt_fast(s2) : 0.f);   // sqrt as s2 * rsqrt(s2): MUFU + FMUL, 2 ulp
      }
    }
  }
}

// PairwiseDistances layer, forward: flat upper triangle of 3-d points.  One CTA per frame: the selected atoms are
// staged in shared memory (SoA), each warp walks whole rows of the triangle, lanes over j, so that the output
// writes are contiguous runs and no per-element index decode is needed.
constexpr int PW_THREADS = 128;
__global__ void __launch_bounds__(PW_THREADS) pairwise_rows3_kernel(const float* __restrict__ x, int64_t b, int n, int64_t bstride,
                                                                    int64_t rstride, int squared, float* __restrict__ out) {
  extern __shared__ float sx[];   // [3][n]
  const int64_t per = (int64_t)n * (n - 1) / 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t bi = blockIdx.x; bi < b; bi += gridDim.x) {
    const float* xb = x + bi * bstride;
    __syncthreads();
    for (int idx = threadIdx.x; idx < 3 * n; idx += PW_THREADS) {
      const int a = idx / 3, c = idx - 3 * a;
      sx[c * n + a] = xb[a * rstride + c];
    }
    __syncthreads();
    float* ob = out + bi * per;
    // rows are dealt to the warps in (i, n-2-i) pairs so that every warp gets the same number of elements
    for (int r = warp; r < (n - 1 + 1) / 2; r += PW_THREADS / 32) {
      for (int half = 0; half < 2; half++) {
        const int i = half == 0 ? r : n - 2 - r;
        if (half == 1 && i == r) break;
        const float xi = sx[i], yi = sx[n + i], zi = sx[2 * n + i];
        float* orow = ob + (int64_t)i * (2 * n - i - 1) / 2 - i - 1;   // + j
        // two columns per lane and iteration; sqrt as s2 * rsqrt(s2) (MUFU + FMUL, 2 ulp) with the zero guarded
        int j = i + 1 + lane;
        for (; j + 32 < n; j += 64) {
          const float dx0 = xi - sx[j], dy0 = yi - sx[n + j], dz0 = zi - sx[2 * n + j];
          const float dx1 = xi - sx[j + 32], dy1 = yi - sx[n + j + 32], dz1 = zi - sx[2 * n + j + 32];
          const float a0 = fmaf(dx0, dx0, fmaf(dy0, dy0, dz0 * dz0));
          const float a1 = fmaf(dx1, dx1, fmaf(dy1, dy1, dz1 * dz1));
          orow[j] = squared ? a0 : (a0 >= EMK_TINY ? a0 * rsqrt_fast(a0) : 0.f);
          orow[j + 32] = squared ? a1 : (a1 >= EMK_TINY ? a1 * rsqrt_fast(a1) : 0.f);
        }
        if (j < n) {
          const float dx = xi - sx[j], dy = yi - sx[n + j], dz = zi - sx[2 * n + j];
          const float s2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
          orow[j] = squared ? s2 : (s2 >= EMK_TINY ? s2 * rsqrt_fast(s2) : 0.f);
        }
      }
    }
  }
}

// PairwiseDistances layer, backward: one CTA per frame, one thread per atom i looping over ALL j (uniform trip
// count, unrolled by 4); positions (and the upstream gradient when it fits) are staged in shared memory.
// grad_x[i] = sum_j coef_ij (x_i - x_j), coef = g/dist (2g when squared), 0 at zero distance.
// The flat index of the pair {i, j} is walked incrementally: it starts at i - 1 (pair (0, i)), advances by n - j - 2
// while j < i (down column i of the triangle) and by 1 afterwards (along row i); at j == i the walk passes through
// rowstart(i) - 1, an unrelated element whose weight is multiplied by x_i - x_i = 0.
template <bool G_IN_SMEM, bool SQUARED>
__global__ void __launch_bounds__(PW_THREADS) pairwise_flat3_bwd_kernel(const float* __restrict__ x, int64_t b, int n, int64_t bstride,
                                                                        int64_t rstride, const float* __restrict__ go,
                                                                        float* __restrict__ gx) {
  extern __shared__ float sm[];   // [3][n] positions, then (optionally) the frame's upstream gradient
  float* sx = sm;
  float* sg = sm + ((3 * n + 3) & ~3);
  const int64_t per = (int64_t)n * (n - 1) / 2;
  for (int64_t bi = blockIdx.x; bi < b; bi += gridDim.x) {
    const float* xb = x + bi * bstride;
    const float* g = go + bi * per;
    __syncthreads();
    for (int idx = threadIdx.x; idx < 3 * n; idx += PW_THREADS) {
      const int a = idx / 3, c = idx - 3 * a;
      sx[c * n + a] = xb[a * rstride + c];
    }
    if (G_IN_SMEM)
      for (int64_t idx = threadIdx.x; idx < per; idx += PW_THREADS) sg[idx] = g[idx];
    __syncthreads();
    const float* gg = G_IN_SMEM ? sg : g;
    for (int i = threadIdx.x; i < n; i += PW_THREADS) {
      const float xi = sx[i], yi = sx[n + i], zi = sx[2 * n + i];
      float ax = 0.f, ay = 0.f, az = 0.f;
      int off = i - 1;
#pragma unroll 4
      for (int j = 0; j < n; j++) {
        const float dx = xi - sx[j], dy = yi - sx[n + j], dz = zi - sx[2 * n + j];
        const float s2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        const float gij = gg[max(off, 0)];
        off += j < i ? n - j - 2 : 1;
        const float coef = SQUARED ? 2.f * gij : (s2 >= EMK_TINY ? gij * rsqrt_fast(s2) : 0.f);
        ax = fmaf(coef, dx, ax); ay = fmaf(coef, dy, ay); az = fmaf(coef, dz, az);
      }
      float* o = gx + bi * bstride + (int64_t)i * rstride;
      o[0] = ax; o[1] = ay; o[2] = az;
    }
  }
}

// PairwiseDistances layer, backward, up to 32 * C selected atoms: every pair is visited ONCE.  A warp walks whole rows
// i of the triangle with its lanes on FIXED columns j = 32 c + lane, so that (a) the upstream gradient of a row is
// read straight from global memory in contiguous runs (no staging), (b) x_j and the column sums sum_i -coef_ij d_ij
// live in registers, (c) the row sum sum_j coef_ij d_ij is one shuffle reduction per row.  The four warps' column
// sums and the row sums meet in shared memory.  (The thread-per-atom version visited every pair twice and read the
// row part of g with a stride of ~n floats between lanes: 0.09 of the HBM roofline at 300 atoms.)
template <int C>
__global__ void __launch_bounds__(PW_THREADS) pairwise_flat3_bwd_cols_kernel(const float* __restrict__ x, int64_t b, int n, int64_t bstride,
                                                                             int64_t rstride, int squared, const float* __restrict__ go,
                                                                             float* __restrict__ gx) {
  extern __shared__ float sm[];
  float* xs = sm;                 // [3][n]
  float* racc = sm + 3 * n;       // [3][n] row sums
  float* cacc = sm + 6 * n;       // [4][3][n] column sums per warp
  const int per = n * (n - 1) / 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t bi = blockIdx.x; bi < b; bi += gridDim.x) {
    const float* xb = x + bi * bstride;
    const float* g = go + bi * per;
    __syncthreads();
    for (int idx = threadIdx.x; idx < 3 * n; idx += PW_THREADS) {
      const int a = idx / 3, c = idx - 3 * a;
      xs[c * n + a] = xb[a * rstride + c];
    }
    if (threadIdx.x < 3) racc[threadIdx.x * n + n - 1] = 0.f;   // the last atom has no row
    __syncthreads();
    float xj[C][3], ca[C][3];
#pragma unroll
    for (int c = 0; c < C; c++) {
      const int j = 32 * c + lane;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        xj[c][k] = j < n ? xs[k * n + j] : 0.f;
        ca[c][k] = 0.f;
      }
    }
    // the upstream gradient of the NEXT row of this warp is in flight while the current row is processed (one row is
    // a dependent chain of global latency + shuffle reduction otherwise)
    float gcur[C], gnxt[C];
    auto load_row = [&](int i, float* dstv) {
      const float* grow = g + i * (2 * n - i - 1) / 2 - i - 1;   // + j
#pragma unroll
      for (int c = 0; c < C; c++) {
        const int j = 32 * c + lane;
        dstv[c] = (i < n - 1 && j > i && j < n) ? __ldg(grow + j) : 0.f;
      }
    };
    load_row(warp, gcur);
    for (int i = warp; i < n - 1; i += PW_THREADS / 32) {
      load_row(i + PW_THREADS / 32, gnxt);
      const float xi = xs[i], yi = xs[n + i], zi = xs[2 * n + i];
      float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
      for (int c = 0; c < C; c++) {
        const int j = 32 * c + lane;
        if (j > i && j < n) {
          const float gij = gcur[c];
          const float dx = xi - xj[c][0], dy = yi - xj[c][1], dz = zi - xj[c][2];
          const float s2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
          const float coef = squared ? 2.f * gij : (s2 >= EMK_TINY ? gij * rsqrt_fast(s2) : 0.f);
          r0 = fmaf(coef, dx, r0); r1 = fmaf(coef, dy, r1); r2 = fmaf(coef, dz, r2);
          ca[c][0] = fmaf(-coef, dx, ca[c][0]); ca[c][1] = fmaf(-coef, dy, ca[c][1]); ca[c][2] = fmaf(-coef, dz, ca[c][2]);
        }
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        r0 += __shfl_xor_sync(0xffffffffu, r0, d);
        r1 += __shfl_xor_sync(0xffffffffu, r1, d);
        r2 += __shfl_xor_sync(0xffffffffu, r2, d);
      }
      if (lane == 0) { racc[i] = r0; racc[n + i] = r1; racc[2 * n + i] = r2; }
#pragma unroll
      for (int c = 0; c < C; c++) gcur[c] = gnxt[c];
    }
#pragma unroll
    for (int c = 0; c < C; c++) {
      const int j = 32 * c + lane;
      if (j < n) {
#pragma unroll
        for (int k = 0; k < 3; k++) cacc[(warp * 3 + k) * n + j] = ca[c][k];
      }
    }
    __syncthreads();
    for (int a = threadIdx.x; a < n; a += PW_THREADS) {
      float* o = gx + bi * bstride + (int64_t)a * rstride;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        float v = racc[k * n + a];
#pragma unroll
        for (int w = 0; w < PW_THREADS / 32; w++) v += cacc[(w * 3 + k) * n + a];
        o[k] = v;
      }
    }
  }
}

// one thread per (frame, atom i): grad_x[i] = sum_j coef_ij (x_i - x_j); coef = g/dist (or 2g when squared).
// Threads of a warp own consecutive atoms of one frame: x_j is a broadcast read, g[pair(j,i)] for j < i is
// contiguous across the warp; the running flat offsets avoid any per-pair index arithmetic.
__global__ void pairwise_small_bwd_kernel(const float* __restrict__ x, int64_t b, int64_t n, int64_t d, int64_t bstride,
                                          int64_t rstride, int squared, int flat, const float* __restrict__ go,
                                          float* __restrict__ gx) {
  const int64_t per = flat ? n * (n - 1) / 2 : n * n;
  for (int64_t bi = blockIdx.y; bi < b; bi += gridDim.y) {
    const float* xb = x + bi * bstride;
    const float* g = go + bi * per;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const float* xi = xb + i * rstride;
      for (int64_t k0 = 0; k0 < d; k0 += 8) {   // up to 8 components per sweep keeps the accumulators in registers
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        float xik[8];
        const int kc = (int)min((int64_t)8, d - k0);
#pragma unroll
        for (int k = 0; k < 8; k++) xik[k] = k < kc ? xi[k0 + k] : 0.f;
        // flat: pair (j,i) with j < i sits at j*(2n-j-1)/2 + (i-j-1): starts at i-1 and advances by n-j-2 per j
        int64_t off_lo = i - 1;
        const int64_t off_hi = i * (2 * n - i - 1) / 2 - i - 1;   // + j for j > i
        for (int64_t j = 0; j < n; j++) {
          float gij;
          if (flat) {
            if (j < i) { gij = g[off_lo]; off_lo += n - j - 2; }
            else if (j > i) gij = g[off_hi + j];
            else continue;
          } else {
            if (j == i) continue;
            gij = g[i * n + j] + g[j * n + i];
          }
          const float* xj = xb + j * rstride;
          float s = 0.f;
          if (d <= 8) {
#pragma unroll
            for (int k = 0; k < 8; k++)
              if (k < kc) { const float t = xik[k] - xj[k]; s = fmaf(t, t, s); }
          } else {
            for (int64_t k = 0; k < d; k++) { const float t = xi[k] - xj[k]; s = fmaf(t, t, s); }
          }
          const float coef = squared ? 2.f * gij : (s >= EMK_TINY ? gij * rsqrt_fast(s) : 0.f);
#pragma unroll
          for (int k = 0; k < 8; k++)
            if (k < kc) acc[k] = fmaf(coef, xik[k] - xj[k0 + k], acc[k]);
        }
#pragma unroll
        for (int k = 0; k < 8; k++)
          if (k < kc) gx[bi * bstride + i * rstride + k0 + k] = acc[k];
      }
    }
  }
}

// ---- PairwiseDistances layer, up to 128 selected atoms (the C-alpha selections of the ADC models): one WARP per frame ----
// Lane l owns the fixed columns j = 32 c + l (C = ceil(n / 32) chunks), positions in registers; the warp walks the rows i of
// the strict upper triangle, x_i is one broadcast LDS.128, and chunks that lie entirely on or below the diagonal are
// skipped (warp-uniform).  Two chunks are processed per instruction with the packed FADD2 / FMUL2 / FFMA2 forms.  The
// outputs of a row are contiguous in the flat order, so every store (forward) or load of the upstream gradient (backward)
// is one contiguous run per chunk, addressed from ONE per-lane pointer with the chunk as an immediate offset.  No index
// table, no per-element decode, no barrier: ~12 instructions per two chunk visits.
constexpr int PWW_THREADS = 128;
constexpr int PWW_MAX_N = 128;       // backward (gradient staged in shared memory)
constexpr int PWW_FWD_MAX_N = 320;   // forward (C <= 10 column chunks in registers)

__device__ __forceinline__ float sqrt_fast(float x) {   // one MUFU.SQRT (2 ulp), exact 0 at 0; no denormal fix-up
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

template <int C, bool SQUARED>
__global__ void __launch_bounds__(PWW_THREADS) pairwise_flat3_warp_kernel(const float* __restrict__ x, int64_t b, int n, int64_t bstride,
                                                                          int64_t rstride, float* __restrict__ out, int64_t out_pitch) {
  __shared__ float4 sx4[PWW_THREADS / 32][C * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = n * (n - 1) / 2;
  float4* sx = sx4[warp];
  for (int64_t f = (int64_t)blockIdx.x * (PWW_THREADS / 32) + warp; f < b; f += (int64_t)gridDim.x * (PWW_THREADS / 32)) {
    const float* xb = x + f * bstride;
    float nx[C], ny[C], nz[C];       // -x_j of this lane's columns
    int jj[C];                       // the column, or -1 beyond the last atom (never "right of the diagonal")
    __syncwarp();
#pragma unroll
    for (int c = 0; c < C; c++) {
      const int j = 32 * c + lane;
      float vx = 0.f, vy = 0.f, vz = 0.f;
      if (j < n) {
        const float* a = xb + (int64_t)j * rstride;
        vx = a[0]; vy = a[1]; vz = a[2];
        sx[j] = make_float4(vx, vy, vz, 0.f);
      }
      nx[c] = -vx; ny[c] = -vy; nz[c] = -vz;
      jj[c] = j < n ? j : -1;
    }
    __syncwarp();
    // out_pitch >= per floats per frame; the tail [per, out_pitch) is zero-filled (a 16-byte row pitch lets the pair-tile
    // kernel read the matrix through TMA as it stands: emk_cartesian_distance_cost)
    for (int64_t k = per + lane; k < out_pitch; k += 32) out[f * out_pitch + k] = 0.f;
    float* orow = out + f * out_pitch - 1 + lane;   // orow[32 c] is the element (i, 32 c + lane) of the current row
#pragma unroll 1
    for (int i = 0; i < n - 1; i++) {
      const float4 xi = sx[i];
#pragma unroll
      for (int c = 0; c + 1 < C; c += 2) {
        if (32 * c + 63 <= i) continue;    // both chunks lie on or below the diagonal
        const float2 dx = __fadd2_rn(make_float2(xi.x, xi.x), make_float2(nx[c], nx[c + 1]));
        const float2 dy = __fadd2_rn(make_float2(xi.y, xi.y), make_float2(ny[c], ny[c + 1]));
        const float2 dz = __fadd2_rn(make_float2(xi.z, xi.z), make_float2(nz[c], nz[c + 1]));
        const float2 s2 = __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, __fmul2_rn(dz, dz)));
        const float v0 = SQUARED ? s2.x : sqrt_fast(s2.x), v1 = SQUARED ? s2.y : sqrt_fast(s2.y);
        if (jj[c] > i) orow[32 * c] = v0;
        if (jj[c + 1] > i) orow[32 * c + 32] = v1;
      }
      if (C & 1) {
        constexpr int c = C - 1;
        const float dx = xi.x + nx[c], dy = xi.y + ny[c], dz = xi.z + nz[c];
        const float s2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        const float v = SQUARED ? s2 : sqrt_fast(s2);
        if (jj[c] > i) orow[32 * c] = v;
      }
      orow += n - i - 2;
    }
  }
}

// PairwiseDistances backward, <= 128 selected atoms: grad_x[a] = sum_b coef_ab (x_a - x_b), coef = g / dist (2 g when squared),
// 0 at zero distance.  One CTA (four warps) per frame, one THREAD per atom i that walks all columns j -- every pair is visited
// from both ends, so there is nothing to reduce across lanes -- with the two things that made round 1's thread-per-atom kernel
// slow removed: x_j is one broadcast LDS.128 (float4 copy) instead of three scalar loads, and two columns are processed per
// iteration with the packed FADD2 / FFMA2 forms (18 instead of 30 instructions per visit).  Measured against a warp-per-frame
// form that visits every pair once (column sums in registers, row sums through a shared-memory transpose; removed again):
// 65 536 x 100 atoms 783 us against 1 064 us, 1 024 x 100 15.6 against 30.4 us (profiles/r02_pairwise_bwd_forms.txt) -- the
// single visit saves FP work but pays it back in row bookkeeping, and its 7 KB of transposes per warp cap the SM at 24 warps.  The frame's upstream gradient is staged once with 8-byte cp.async; its flat offsets are walked
// incrementally (start at pair (0, i); advance by n - j - 2 down column i while j < i, by 1 along row i afterwards; at
// j == i the walk passes an unrelated element whose weight is masked by the zero distance).  21 KB of shared memory per
// CTA: ten CTAs = 40 warps per SM.
constexpr int PWB2_THREADS = 128;

template <bool SQUARED>
__global__ void __launch_bounds__(PWB2_THREADS) pairwise_flat3_bwd2_kernel(const float* __restrict__ x, int64_t b, int n, int64_t bstride,
                                                                           int64_t rstride, const float* __restrict__ go,
                                                                           float* __restrict__ gx) {
  extern __shared__ __align__(16) float smb[];
  float4* sx = reinterpret_cast<float4*>(smb);          // [PWW_MAX_N]
  float* sg = smb + 4 * PWW_MAX_N;                      // [n (n-1) / 2], 8-byte aligned
  const int t = threadIdx.x;
  const int per = n * (n - 1) / 2;
  for (int64_t f = blockIdx.x; f < b; f += gridDim.x) {
    const float* g = go + f * per;
    __syncthreads();                                    // the previous frame has been consumed
    if (t < n) {
      const float* a = x + f * bstride + (int64_t)t * rstride;
      sx[t] = make_float4(a[0], a[1], a[2], 0.f);
    }
    if ((reinterpret_cast<uintptr_t>(g) & 7) == 0) {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sg);
      for (int e = 2 * t; e + 1 < per; e += 2 * PWB2_THREADS)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 4u * e), "l"(g + e) : "memory");
      if ((per & 1) && t == 0) sg[per - 1] = g[per - 1];
      asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    } else {
      for (int e = t; e < per; e += PWB2_THREADS) sg[e] = g[e];
    }
    __syncthreads();
    if (t < n) {
      const int i = t;
      const float4 xi = sx[i];
      float2 ax = make_float2(0.f, 0.f), ay = ax, az = ax;
      int off = i - 1;
      auto coef_of = [](float gv, float s2) { return SQUARED ? 2.f * gv : (s2 >= EMK_TINY ? gv * rsqrt_fast(s2) : 0.f); };
      int j = 0;
#pragma unroll 2
      for (; j + 1 < n; j += 2) {
        const int offa = off;
        const int offb = offa + (j < i ? n - j - 2 : 1);
        off = offb + (j + 1 < i ? n - j - 3 : 1);
        const float ga = sg[max(offa, 0)], gb = sg[min(max(offb, 0), per - 1)];
        const float4 xa = sx[j], xb = sx[j + 1];
        const float2 dx = __fadd2_rn(make_float2(xi.x, xi.x), make_float2(-xa.x, -xb.x));
        const float2 dy = __fadd2_rn(make_float2(xi.y, xi.y), make_float2(-xa.y, -xb.y));
        const float2 dz = __fadd2_rn(make_float2(xi.z, xi.z), make_float2(-xa.z, -xb.z));
        const float2 s2 = __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, __fmul2_rn(dz, dz)));
        const float2 cf = make_float2(coef_of(ga, s2.x), coef_of(gb, s2.y));
        ax = __ffma2_rn(cf, dx, ax); ay = __ffma2_rn(cf, dy, ay); az = __ffma2_rn(cf, dz, az);
      }
      float rx = ax.x + ax.y, ry = ay.x + ay.y, rz = az.x + az.y;
      if (j < n) {                                      // odd n: the last column
        const float gv = sg[min(max(off, 0), per - 1)];
        const float4 xa = sx[j];
        const float dx = xi.x - xa.x, dy = xi.y - xa.y, dz = xi.z - xa.z;
        const float cf = coef_of(gv, fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
        rx = fmaf(cf, dx, rx); ry = fmaf(cf, dy, ry); rz = fmaf(cf, dz, rz);
      }
      float* o = gx + f * bstride + (int64_t)i * rstride;
      o[0] = rx; o[1] = ry; o[2] = rz;
    }
  }
}

// ---- host launchers -------------------------------------------------------------------------------------------
int periodic_distance_device(const float* a, const float* b, int64_t count, double P, float* out, cudaStream_t st) {
  EMK_REQUIRE(count == 0 || (a && b && out), EMK_E_NULL, "emk_periodic_distance: NULL pointer argument");
  EMK_REQUIRE(count >= 0, EMK_E_SHAPE, "emk_periodic_distance: negative count");
  if (count == 0) return EMK_OK;
  periodic_distance_kernel<<<grid_for(count), 256, 0, st>>>(a, b, count, std::isinf(P) ? INFINITY : (float)P, out);
  return launch_status("periodic_distance_kernel");
}
int periodic_distance_bwd_device(const float* a, const float* b, int64_t count, double P, const float* go, float* ga, float* gb,
                                 cudaStream_t st) {
  EMK_REQUIRE(count == 0 || (a && b && go && (ga || gb)), EMK_E_NULL, "emk_periodic_distance_bwd: NULL pointer argument");
  if (count == 0) return EMK_OK;
  periodic_distance_bwd_kernel<<<grid_for(count), 256, 0, st>>>(a, b, count, std::isinf(P) ? INFINITY : (float)P, go, ga, gb);
  return launch_status("periodic_distance_bwd_kernel");
}
int sigmoid_device(const float* r, int64_t count, float sig, float a, float b, float* out, cudaStream_t st) {
  EMK_REQUIRE(count == 0 || (r && out), EMK_E_NULL, "emk_sigmoid: NULL pointer argument");
  EMK_REQUIRE(sig > 0 && a > 0 && b > 0, EMK_E_ARG, "emk_sigmoid: parameters must be > 0");
  if (count == 0) return EMK_OK;
  sigmoid_kernel<<<grid_for(count), 256, 0, st>>>(r, count, make_sig_spec(sig, a, b), out);
  return launch_status("sigmoid_kernel");
}
int sigmoid_bwd_device(const float* r, int64_t count, float sig, float a, float b, const float* go, float* gr, cudaStream_t st) {
  EMK_REQUIRE(count == 0 || (r && go && gr), EMK_E_NULL, "emk_sigmoid_bwd: NULL pointer argument");
  EMK_REQUIRE(sig > 0 && a > 0 && b > 0, EMK_E_ARG, "emk_sigmoid_bwd: parameters must be > 0");
  if (count == 0) return EMK_OK;
  sigmoid_bwd_kernel<<<grid_for(count), 256, 0, st>>>(r, count, make_sig_spec(sig, a, b), go, gr);
  return launch_status("sigmoid_bwd_kernel");
}
static inline void periodic_scale(double P, float* scale, int* rescale) {
  // reference: `if periodicity != 2*pi: x = x / periodicity * 2 * pi` (layers.py:207-208)
  *rescale = (P != 2.0 * M_PI) ? 1 : 0;
  *scale = (float)(2.0 * M_PI / P);
}
int periodic_input_device(const float* x, int64_t rows, int64_t d, double P, float* out, cudaStream_t st) {
  EMK_REQUIRE(rows * d == 0 || (x && out), EMK_E_NULL, "emk_periodic_input: NULL pointer argument");
  EMK_REQUIRE(rows >= 0 && d >= 0 && P > 0, EMK_E_ARG, "emk_periodic_input: bad arguments");
  if (rows * d == 0) return EMK_OK;
  float sc; int rs;
  periodic_scale(P, &sc, &rs);
  EMK_REQUIRE(d < (1LL << 30), EMK_E_UNSUPPORTED, "emk_periodic_input: more than 2^30 columns");
  const bool vec = d % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  const int threads = (int)std::min<int64_t>(256, std::max<int64_t>(32, ((vec ? d / 4 : d) + 31) / 32 * 32));
  const unsigned grid = (unsigned)std::min<int64_t>(rows, (int64_t)sm_count() * 32);
  if (vec) periodic_input_kernel<true><<<grid, threads, 0, st>>>(x, rows, (int)d, sc, rs, out);
  else periodic_input_kernel<false><<<grid, threads, 0, st>>>(x, rows, (int)d, sc, rs, out);
  return launch_status("periodic_input_kernel");
}
int periodic_input_bwd_device(const float* x, int64_t rows, int64_t d, double P, const float* go, float* gx, cudaStream_t st) {
  EMK_REQUIRE(rows * d == 0 || (x && go && gx), EMK_E_NULL, "emk_periodic_input_bwd: NULL pointer argument");
  EMK_REQUIRE(rows >= 0 && d >= 0 && P > 0, EMK_E_ARG, "emk_periodic_input_bwd: bad arguments");
  if (rows * d == 0) return EMK_OK;
  float sc; int rs;
  periodic_scale(P, &sc, &rs);
  EMK_REQUIRE(d < (1LL << 30), EMK_E_UNSUPPORTED, "emk_periodic_input_bwd: more than 2^30 columns");
  const bool vec = d % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(go) | reinterpret_cast<uintptr_t>(gx)) & 15) == 0;
  const int threads = (int)std::min<int64_t>(256, std::max<int64_t>(32, ((vec ? d / 4 : d) + 31) / 32 * 32));
  const unsigned grid = (unsigned)std::min<int64_t>(rows, (int64_t)sm_count() * 32);
  if (vec) periodic_input_bwd_kernel<true><<<grid, threads, 0, st>>>(x, rows, (int)d, sc, rs, go, gx);
  else periodic_input_bwd_kernel<false><<<grid, threads, 0, st>>>(x, rows, (int)d, sc, rs, go, gx);
  return launch_status("periodic_input_bwd_kernel");
}
int rotation_matrix_device(const float* axis, const float* angle, int64_t b, float* out, cudaStream_t st) {
  EMK_REQUIRE(b <= 0 || (axis && angle && out), EMK_E_NULL, "emk_rotation_matrix: NULL pointer argument");
  if (b <= 0) return b == 0 ? EMK_OK : fail(EMK_E_SHAPE, "emk_rotation_matrix: negative batch");
  rotation_matrix_kernel<<<grid_for(b), 256, 0, st>>>(axis, angle, b, out);
  return launch_status("rotation_matrix_kernel");
}
int column_mean_device(const float* x, int64_t rows, int64_t cols, float* out, cudaStream_t st) {
  EMK_REQUIRE(x && out, EMK_E_NULL, "emk_column_mean: NULL pointer argument");
  EMK_REQUIRE(rows >= 1 && cols >= 1, EMK_E_SHAPE, "emk_column_mean: need rows, cols >= 1");
  const int split = (int)min((int64_t)CM_SPLIT, (rows + 255) / 256);
  double* part = nullptr;
  {
    int rc0 = scratch_alloc(reinterpret_cast<void**>(&part), (size_t)split * cols * sizeof(double), st);
    if (rc0) return rc0;
  }
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)split);
  column_partial_kernel<<<grid, 256, 0, st>>>(x, rows, cols, part);
  int rc = launch_status("column_partial_kernel");
  if (rc == EMK_OK) {
    column_final_kernel<<<grid_for(cols), 256, 0, st>>>(part, split, rows, cols, out);
    rc = launch_status("column_final_kernel");
  }
  cudaFreeAsync(part, st);
  return rc;
}
// flat upper triangle of (b, n <= 128, 3) points, one warp per frame, `out_pitch` floats per output row (>= n (n-1) / 2)
int pairwise_flat3_warp_device(const float* x, int64_t b, int64_t n, int64_t bstride, int64_t rstride, int squared, float* out,
                               int64_t out_pitch, cudaStream_t st) {
  EMK_REQUIRE(n >= 2 && n <= PWW_FWD_MAX_N && out_pitch >= n * (n - 1) / 2, EMK_E_ARG, "pairwise (warp per frame): bad geometry");
  if (b == 0) return EMK_OK;
  const unsigned grid = (unsigned)std::min<int64_t>((b + PWW_THREADS / 32 - 1) / (PWW_THREADS / 32), (int64_t)sm_count() * 16);
#define EMK_PWW(CC)                                                                                                                 \
  do {                                                                                                                              \
    if (squared) pairwise_flat3_warp_kernel<CC, true><<<grid, PWW_THREADS, 0, st>>>(x, b, (int)n, bstride, rstride, out, out_pitch); \
    else pairwise_flat3_warp_kernel<CC, false><<<grid, PWW_THREADS, 0, st>>>(x, b, (int)n, bstride, rstride, out, out_pitch);        \
  } while (0)
  if (n <= 32) EMK_PWW(1);
  else if (n <= 64) EMK_PWW(2);
  else if (n <= 96) EMK_PWW(3);
  else if (n <= 128) EMK_PWW(4);
  else if (n <= 192) EMK_PWW(6);
  else if (n <= 256) EMK_PWW(8);
  else EMK_PWW(10);
#undef EMK_PWW
  return launch_status("pairwise_flat3_warp_kernel");
}
int pairwise_warp_max_atoms() { return PWW_FWD_MAX_N; }

int pairwise_small_device(const float* x, int64_t b, int64_t n, int64_t d, int64_t bstride, int64_t rstride, int squared,
                          int flat, float* out, cudaStream_t st) {
  const int64_t per = flat ? n * (n - 1) / 2 : n * n;
  if (b * per == 0) return EMK_OK;
  if (flat && d == 3 && n >= 2 && n <= PWW_FWD_MAX_N) return pairwise_flat3_warp_device(x, b, n, bstride, rstride, squared, out, per, st);
  if (flat && d == 3 && n >= 2 && n <= 8192) {
    if (n <= PWT_MAX_N) {
      const int64_t per3 = n * (n - 1) / 2;
      const int G = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(8, 32768 / per3 + 1), b));
      const size_t smem = (size_t)G * 7 * (size_t)n * sizeof(float) + (size_t)per3 * 4;
      static bool cfg[kMaxDevices] = {false};
      if (first_use_on_device(cfg)) EMK_CUDA(cudaFuncSetAttribute(pairwise_flat3_tab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
      const unsigned grid = (unsigned)std::min<int64_t>((b + G - 1) / G, (int64_t)sm_count() * 8);
      pairwise_flat3_tab_kernel<<<grid, PWF_THREADS, smem, st>>>(x, b, (int)n, bstride, rstride, squared, G, out);
      return launch_status("pairwise_flat3_tab_kernel");
    }
    const size_t smem = 3 * (size_t)n * sizeof(float);   // up to 96 KB at n = 8192: above the 48 KB default limit
    static bool cfg_rows[kMaxDevices] = {false};
    if (first_use_on_device(cfg_rows)) EMK_CUDA(cudaFuncSetAttribute(pairwise_rows3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    const unsigned grid = (unsigned)std::min<int64_t>(b, (int64_t)sm_count() * 16);
    pairwise_rows3_kernel<<<grid, PW_THREADS, smem, st>>>(x, b, (int)n, bstride, rstride, squared, out);
    return launch_status("pairwise_rows3_kernel");
  }
  const int64_t bx = std::min<int64_t>((per + 255) / 256, 1024);
  dim3 grid((unsigned)bx, (unsigned)std::min<int64_t>(b, 65535));
  pairwise_small_kernel<<<grid, 256, 0, st>>>(x, b, n, d, bstride, rstride, squared, flat, out);
  return launch_status("pairwise_small_kernel");
}
int pairwise_small_bwd_device(const float* x, int64_t b, int64_t n, int64_t d, int64_t bstride, int64_t rstride, int squared,
                              int flat, const float* go, float* gx, cudaStream_t st) {
  if (b * n == 0) return EMK_OK;
  // 129 .. 320 atoms: pair-once kernel (2.5x the thread-per-atom kernel at 300 atoms); up to 128 atoms the whole
  // upstream gradient of a frame fits in shared memory next to the coordinates and the thread-per-atom kernel is as fast
  if (flat && d == 3 && n >= 2 && n <= PWW_MAX_N) {
    // CTA per frame, thread per atom, packed two-column walk
    const size_t smem = (4 * (size_t)PWW_MAX_N + (size_t)((n * (n - 1) / 2 + 1) & ~(int64_t)1)) * sizeof(float);
    const unsigned grid = (unsigned)std::min<int64_t>(b, (int64_t)sm_count() * 10);
    if (squared) pairwise_flat3_bwd2_kernel<true><<<grid, PWB2_THREADS, smem, st>>>(x, b, (int)n, bstride, rstride, go, gx);
    else pairwise_flat3_bwd2_kernel<false><<<grid, PWB2_THREADS, smem, st>>>(x, b, (int)n, bstride, rstride, go, gx);
    return launch_status("pairwise_flat3_bwd2_kernel");
  }
  if (flat && d == 3 && n > 128 && n <= 320) {
    const size_t smem = 18 * (size_t)n * sizeof(float);
    const unsigned grid = (unsigned)std::min<int64_t>(b, (int64_t)sm_count() * 16);
#define EMK_PWB(CC) pairwise_flat3_bwd_cols_kernel<CC><<<grid, PW_THREADS, smem, st>>>(x, b, (int)n, bstride, rstride, squared, go, gx)
    if (n <= 192) EMK_PWB(6);
    else if (n <= 256) EMK_PWB(8);
    else EMK_PWB(10);
#undef EMK_PWB
    return launch_status("pairwise_flat3_bwd_cols_kernel");
  }
  if (flat && d == 3 && n >= 2 && n <= 8192) {
    const size_t sx = ((3 * (size_t)n + 3) & ~(size_t)3) * sizeof(float);
    const size_t sg = (size_t)n * (n - 1) / 2 * sizeof(float);
    const unsigned grid = (unsigned)std::min<int64_t>(b, (int64_t)sm_count() * 16);
#define EMK_PWB1(GS, SQ, SMEM)                                                                                              \
  do {                                                                                                                      \
    static bool cfg[kMaxDevices] = {false};                                                                                 \
    if (first_use_on_device(cfg))                                                                                           \
      EMK_CUDA(cudaFuncSetAttribute(pairwise_flat3_bwd_kernel<GS, SQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); \
    pairwise_flat3_bwd_kernel<GS, SQ><<<grid, PW_THREADS, SMEM, st>>>(x, b, (int)n, bstride, rstride, go, gx);               \
  } while (0)
    if (sx + sg <= 96 * 1024) {
      if (squared) EMK_PWB1(true, true, sx + sg);
      else EMK_PWB1(true, false, sx + sg);
    } else {
      if (squared) EMK_PWB1(false, true, sx);
      else EMK_PWB1(false, false, sx);
    }
#undef EMK_PWB1
    return launch_status("pairwise_flat3_bwd_kernel");
  }
  const int threads = n >= 128 ? 128 : (int)((n + 31) / 32 * 32);
  dim3 grid((unsigned)((n + threads - 1) / threads), (unsigned)std::min<int64_t>(b, 65535));
  pairwise_small_bwd_kernel<<<grid, threads, 0, st>>>(x, b, n, d, bstride, rstride, squared, flat, go, gx);
  return launch_status("pairwise_small_bwd_kernel");
}

}  // namespace emk
