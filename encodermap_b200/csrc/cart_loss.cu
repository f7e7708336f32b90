// Fused Cartesian branch of the ADC model:
//   PairwiseDistances("output")  ->  cartesian_loss  (+ ADCClashMetric)
// i.e. encodermap/models/layers.py:1252-1267 (atom selection + flat upper-triangle pairwise distances),
// encodermap/loss_functions/loss_functions.py:1020-1065 (mean |d_in - d_out|, mean square, or mean per-frame norm) and
// encodermap/callbacks/metrics.py:512-520 (pairs closer than clash_distance), forward AND the gradient w.r.t. the
// back-mapped coordinates from one launch.  The (frames, n_pairs) matrix of output distances, its gradient, and -- when the
// input coordinates are passed instead of their distance matrix -- the (frames, n_pairs) input matrix never exist in HBM:
// per frame the kernel reads 12 bytes per atom and writes 12 bytes per atom instead of moving ~5 x 4 bytes per PAIR
// (100 C-alpha atoms: 3.6 KB instead of 99 KB per frame; the step's five launches become one).
//
// One WARP per frame.  The selected atoms are staged in shared memory (SoA); lane l owns the fixed columns j = l + 32 c
// (positions and column sums of the gradient in registers), the warp walks the rows i of the upper triangle: x_i is a
// broadcast read, a stored target row is read coalesced (the pairs (i, j > i) are contiguous in the flat order), the row sum
// of the gradient is one butterfly reduction per row.  Every pair is visited once.
#include "emk_common.cuh"

namespace emk {

constexpr int CL_THREADS = 128;   // four frames per CTA

enum CartVariant : int { kMeanAbs = 0, kMeanSquare = 1, kMeanNorm = 2 };

struct CartLossParams {
  const float* xyz;        // (b, n_atoms, 3) back-mapped coordinates
  const float* target;     // (b, n_pairs) input pair distances, or (b, n_atoms, 3) input coordinates
  int64_t b;
  int n_atoms;
  int first, ns, step;     // selected atoms: first + a * step, a < ns
  int target_is_xyz;
  int variant;
  float clash2;            // squared clash distance (< 0: no clash count)
  double* loss_sum;        // += sum over frames of: sum_p |diff| (abs), sum_p diff^2 (square), sqrt(sum_p diff^2) (norm)
  float* grad_xyz;         // (b, n_atoms, 3) or null: d(loss_sum)/d(xyz), zero for unselected atoms
  long long* clashes;      // (b) or null
};

// WPF = warps per frame: 1 (large batches: four frames per CTA) or 4 (training batches of ~1000 frames, where one warp per
// frame leaves most of the machine idle: the rows of the triangle are dealt round-robin to the CTA's four warps, column sums
// and the frame's partial sums meet in shared memory).  VARIANT and the kind of target are template parameters so that the
// inner loop carries no selects for the variants it is not.
template <int C, int VARIANT, bool TXYZ, int WPF>
__global__ void __launch_bounds__(CL_THREADS) cart_pair_loss_kernel(const CartLossParams p) {
  extern __shared__ float sm[];
  constexpr int NW = CL_THREADS / 32;
  constexpr int FPB = NW / WPF;                    // frames per CTA
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = warp / WPF, sub = warp % WPF;   // frame slot within the CTA, this warp's share of the rows
  const int64_t frame = (int64_t)blockIdx.x * FPB + slot;
  const int ns = p.ns;
  const int per_frame = (9 + (WPF > 1 ? 3 * WPF : 0)) * ns;   // out coords [3][ns], in coords [3][ns], row sums [3][ns], column sums [WPF][3][ns]
  float* so = sm + (size_t)slot * per_frame;
  float* si = so + 3 * ns;
  float* sr = si + 3 * ns;
  float* sc = sr + 3 * ns;
  __shared__ double wsum[NW];
  __shared__ float fsum[NW];
  __shared__ int csum[NW];
  const bool live = frame < p.b;                   // uniform per frame slot; with WPF == NW uniform per CTA
  const bool want_grad = p.grad_xyz != nullptr;
  double frame_loss = 0.0;
  float co[C][3], ci[C][3], ca[C][3];
  if (live) {
    const float* xo = p.xyz + frame * (int64_t)(3 * p.n_atoms);
    const float* xi = TXYZ ? p.target + frame * (int64_t)(3 * p.n_atoms) : nullptr;
    for (int a = sub * 32 + lane; a < ns; a += 32 * WPF) {
      const int at = 3 * (p.first + a * p.step);
      so[a] = xo[at]; so[ns + a] = xo[at + 1]; so[2 * ns + a] = xo[at + 2];
      if (TXYZ) { si[a] = xi[at]; si[ns + a] = xi[at + 1]; si[2 * ns + a] = xi[at + 2]; }
      if (a == ns - 1) { sr[a] = 0.f; sr[ns + a] = 0.f; sr[2 * ns + a] = 0.f; }   // the last atom has no row
    }
  }
  if (WPF > 1) __syncthreads(); else __syncwarp();
  if (live) {
#pragma unroll
    for (int c = 0; c < C; c++) {
      const int j = 32 * c + lane;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        co[c][k] = j < ns ? so[k * ns + j] : 0.f;
        ci[c][k] = (TXYZ && j < ns) ? si[k * ns + j] : 0.f;
        ca[c][k] = 0.f;
      }
    }
  }
  const int64_t n_pairs = (int64_t)ns * (ns - 1) / 2;
  const float* trow_base = TXYZ ? nullptr : p.target + frame * n_pairs;
  // mean_norm needs the frame's norm before any gradient term: a first sweep sums the squares
  constexpr int SWEEPS = VARIANT == kMeanNorm ? 2 : 1;
  float inv_norm = 1.f;
#pragma unroll 1
  for (int sweep = 0; sweep < SWEEPS; sweep++) {
    const bool grad_sweep = want_grad && sweep == SWEEPS - 1;
    const bool loss_sweep = sweep == 0;
    float lsum = 0.f;
    int clash = 0;
    if (live && (grad_sweep || loss_sweep)) {
#pragma unroll 1
      for (int i = sub; i < ns - 1; i += WPF) {
        const float xo0 = so[i], xo1 = so[ns + i], xo2 = so[2 * ns + i];
        float xi0 = 0.f, xi1 = 0.f, xi2 = 0.f;
        if (TXYZ) { xi0 = si[i]; xi1 = si[ns + i]; xi2 = si[2 * ns + i]; }
        const float* trow = TXYZ ? nullptr : trow_base + (i * (2 * ns - i - 1) / 2 - i - 1);   // + j
        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
        for (int c = 0; c < C; c++) {
          if (32 * c + 31 <= i) continue;           // the whole chunk lies on or below the diagonal (warp-uniform)
          const int j = 32 * c + lane;
          if (j > i && j < ns) {
            const float dx = xo0 - co[c][0], dy = xo1 - co[c][1], dz = xo2 - co[c][2];
            const float s2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
            const float rinv = s2 >= EMK_TINY ? rsqrt_fast(s2) : 0.f;
            // __fmul_rn: a product the compiler may not contract into `fma(s2, rinv, -din)` below -- identical structures must
            // give a difference of exactly 0 as they do in the reference, not the rounding error of one of the two products
            const float dout = __fmul_rn(s2, rinv);             // pairwise_dist: 0 at zero distance (distances.py:244-253)
            float din;
            if (!TXYZ) {
              din = __ldg(trow + j);
            } else {
              const float ex = xi0 - ci[c][0], ey = xi1 - ci[c][1], ez = xi2 - ci[c][2];
              const float t2 = fmaf(ex, ex, fmaf(ey, ey, ez * ez));
              din = t2 >= EMK_TINY ? __fmul_rn(t2, rsqrt_fast(t2)) : 0.f;
            }
            const float diff = dout - din;                      // the reference forms y_true - y_pred = -(diff): same |.| and square
            if (loss_sweep) {
              lsum += VARIANT == kMeanAbs ? fabsf(diff) : diff * diff;
              if (s2 < p.clash2) clash++;
            }
            if (grad_sweep) {
              // d/d(dout): sign(diff) (abs; 0 at 0 as tf.abs), 2 diff (square), diff / norm (norm); d(dout)/dx_i = (x_i - x_j)/dout,
              // zero at zero distance (the reference's mask)
              const float w = VARIANT == kMeanAbs ? (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f))
                                                  : (VARIANT == kMeanSquare ? 2.f * diff : diff * inv_norm);
              const float coef = w * rinv;
              r0 = fmaf(coef, dx, r0); r1 = fmaf(coef, dy, r1); r2 = fmaf(coef, dz, r2);
              ca[c][0] = fmaf(-coef, dx, ca[c][0]); ca[c][1] = fmaf(-coef, dy, ca[c][1]); ca[c][2] = fmaf(-coef, dz, ca[c][2]);
            }
          }
        }
        if (grad_sweep) {
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) {
            r0 += __shfl_xor_sync(0xffffffffu, r0, d);
            r1 += __shfl_xor_sync(0xffffffffu, r1, d);
            r2 += __shfl_xor_sync(0xffffffffu, r2, d);
          }
          if (lane == 0) { sr[i] = r0; sr[ns + i] = r1; sr[2 * ns + i] = r2; }
        }
      }
    }
    if (loss_sweep) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        lsum += __shfl_xor_sync(0xffffffffu, lsum, d);
        clash += __shfl_xor_sync(0xffffffffu, clash, d);
      }
      if (WPF > 1) {                                 // the frame's sums over its WPF warps
        if (lane == 0) { fsum[warp] = lsum; csum[warp] = clash; }
        __syncthreads();
        lsum = 0.f; clash = 0;
#pragma unroll
        for (int w = 0; w < WPF; w++) { lsum += fsum[slot * WPF + w]; clash += csum[slot * WPF + w]; }
      }
      if (VARIANT == kMeanNorm) {
        const float nrm = sqrtf(lsum);
        frame_loss = (double)nrm;
        inv_norm = nrm > 0.f ? 1.f / nrm : 0.f;                 // tf.norm's gradient at 0 is NaN in TensorFlow; 0 here
      } else {
        frame_loss = (double)lsum;
      }
      if (live && sub == 0 && lane == 0 && p.clashes) p.clashes[frame] = clash;
    }
  }
  if (want_grad) {
    float* g = p.grad_xyz + frame * (int64_t)(3 * p.n_atoms);
    if (WPF > 1) {
      if (live) {
#pragma unroll
        for (int c = 0; c < C; c++) {
          const int j = 32 * c + lane;
          if (j < ns) {
#pragma unroll
            for (int k = 0; k < 3; k++) sc[(sub * 3 + k) * ns + j] = ca[c][k];
          }
        }
      }
      __syncthreads();
      if (live) {
        // unselected atoms get zero: the whole row is written once, selected atoms with their sums
        const int first = p.first, step = p.step;
        for (int a = sub * 32 + lane; a < p.n_atoms; a += 32 * WPF) {
          const int rel = a - first;
          const int js = (rel >= 0 && rel % step == 0) ? rel / step : -1;
          float v0 = 0.f, v1 = 0.f, v2 = 0.f;
          if (js >= 0 && js < ns) {
            v0 = sr[js]; v1 = sr[ns + js]; v2 = sr[2 * ns + js];
#pragma unroll
            for (int w = 0; w < WPF; w++) { v0 += sc[(w * 3 + 0) * ns + js]; v1 += sc[(w * 3 + 1) * ns + js]; v2 += sc[(w * 3 + 2) * ns + js]; }
          }
          g[3 * a] = v0; g[3 * a + 1] = v1; g[3 * a + 2] = v2;
        }
      }
    } else if (live) {
      // unselected atoms get zero; the whole row is cleared first (coalesced), then the selected atoms are filled in
      for (int e = lane; e < 3 * p.n_atoms; e += 32) g[e] = 0.f;
      __syncwarp();
#pragma unroll
      for (int c = 0; c < C; c++) {
        const int j = 32 * c + lane;
        if (j < ns) {
          float* o = g + 3 * (p.first + j * p.step);
          o[0] = ca[c][0] + sr[j]; o[1] = ca[c][1] + sr[ns + j]; o[2] = ca[c][2] + sr[2 * ns + j];
        }
      }
    }
  }
  if (lane == 0) wsum[warp] = (sub == 0) ? frame_loss : 0.0;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < NW; w++) t += wsum[w];
    if (t != 0.0) atomicAdd(p.loss_sum, t);
  }
}

template <int C, int WPF>
static int launch_cart(const CartLossParams& p, cudaStream_t st) {
  constexpr int FPB = (CL_THREADS / 32) / WPF;
  const size_t smem = (size_t)FPB * (9 + (WPF > 1 ? 3 * WPF : 0)) * p.ns * sizeof(float);
  const unsigned grid = (unsigned)((p.b + FPB - 1) / FPB);
#define EMK_CL(V, T)                                                                                                              \
  do {                                                                                                                            \
    static bool cfg[kMaxDevices] = {false};                                                                                       \
    if (first_use_on_device(cfg))                                                                                                 \
      EMK_CUDA(cudaFuncSetAttribute(cart_pair_loss_kernel<C, V, T, WPF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); \
    cart_pair_loss_kernel<C, V, T, WPF><<<grid, CL_THREADS, smem, st>>>(p);                                                       \
  } while (0)
  if (p.target_is_xyz) {
    if (p.variant == kMeanAbs) EMK_CL(kMeanAbs, true);
    else if (p.variant == kMeanSquare) EMK_CL(kMeanSquare, true);
    else EMK_CL(kMeanNorm, true);
  } else {
    if (p.variant == kMeanAbs) EMK_CL(kMeanAbs, false);
    else if (p.variant == kMeanSquare) EMK_CL(kMeanSquare, false);
    else EMK_CL(kMeanNorm, false);
  }
#undef EMK_CL
  return launch_status("cart_pair_loss_kernel");
}

int cart_pair_loss_device(const CartLossParams& p, cudaStream_t st) {
  EMK_REQUIRE(p.loss_sum && (p.b == 0 || (p.xyz && p.target)), EMK_E_NULL, "emk_cartesian_pair_loss: NULL pointer argument");
  EMK_REQUIRE(p.b >= 0 && p.n_atoms >= 1 && p.ns >= 0 && p.step >= 1 && p.first >= 0 &&
                  (p.ns == 0 || p.first + (int64_t)(p.ns - 1) * p.step < p.n_atoms),
              EMK_E_SHAPE, "emk_cartesian_pair_loss: atom selection (first %d, count %d, step %d) outside %d atoms", p.first, p.ns, p.step,
              p.n_atoms);
  EMK_REQUIRE(p.variant >= 0 && p.variant <= 2, EMK_E_ARG, "emk_cartesian_pair_loss: variant must be 0 (mean_abs), 1 (mean_square) or 2 (mean_norm)");
  EMK_REQUIRE(p.ns <= 320, EMK_E_UNSUPPORTED, "emk_cartesian_pair_loss: more than 320 selected atoms per frame (got %d)", p.ns);
  if (p.b == 0) return EMK_OK;
  if (p.ns < 2) {   // no pairs: zero loss, zero gradient
    if (p.grad_xyz) EMK_CUDA(cudaMemsetAsync(p.grad_xyz, 0, (size_t)p.b * p.n_atoms * 3 * sizeof(float), st));
    if (p.clashes) EMK_CUDA(cudaMemsetAsync(p.clashes, 0, (size_t)p.b * sizeof(long long), st));
    return EMK_OK;
  }
  // one warp per frame when that alone fills the machine (eight warps per SM sub-partition would be 148 x 32 frames);
  // four warps per frame below
  const bool wide = p.b < (int64_t)sm_count() * 16;
#define EMK_CLC(CC) return wide ? launch_cart<CC, 4>(p, st) : launch_cart<CC, 1>(p, st)
  if (p.ns <= 32) EMK_CLC(1);
  else if (p.ns <= 64) EMK_CLC(2);
  else if (p.ns <= 128) EMK_CLC(4);
  else if (p.ns <= 192) EMK_CLC(6);
  else EMK_CLC(10);
#undef EMK_CLC
}

}  // namespace emk
