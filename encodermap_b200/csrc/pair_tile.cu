// Pair-tile kernels: the all-pairs sketch-map sigmoid cost (fused forward + backward) and the
// (periodic) distance matrices, for sm_100a.
//
// Work decomposition.  The N x N pair space is cut into tiles of 128 rows x 64 columns; only tiles
// that touch the upper triangle are evaluated (two "diagonal" tiles per 128-row block are computed in
// full and weighted 1, every other tile stands for itself and its mirror image and is weighted 2).
// One 256-thread CTA owns one tile; two CTAs are resident per SM so that one CTA's epilogue
// (sigmoids, reductions, atomics) overlaps the other's FP32 main loop.
//
// Main loop.  High-d rows are streamed through shared memory in 32-float (128-byte) k-chunks by TMA
// (cp.async.bulk.tensor.2d, SWIZZLE_128B, mbarrier complete_tx), 4 stages deep.  Each thread owns an
// 8 x 4 register micro-tile and reads operands with conflict-free LDS.128 (the TMA swizzle is undone
// in the address).  Per (pair, dim) the periodic minimum-image distance costs
//     d = a - b            FADD2  (packed, 2 dims per issue slot)
//     t = P - |d|          FADD   (|.| and - are free source modifiers)
//     m = min(|d|, t)      FMNMX  (ALU pipe)
//     acc += m * m         FFMA2  (packed)
// i.e. 3 issue slots instead of 4 -- measured at 95% of the 4-instruction issue roofline
// (profiles/r01_pipe_probe.txt).  The accumulator is a float2 over even/odd k.
//
// Epilogue.  Both sigmoids are evaluated from SQUARED distances (emk_common.cuh), the squared
// difference is accumulated in double, and dL/dz is reduced by warp shuffles (row side) and shared
// memory atomics (column side) down to one red.global.add.f32 per (row, component, tile).
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>
#include <cuda.h>

#include "emk_common.cuh"

namespace emk {

namespace cg = cooperative_groups;

constexpr int KC = 32;             // floats per k-chunk (128 B = one swizzle row)
constexpr int STAGES = 4;
constexpr int NTHREADS = 256;
constexpr int MAX_LATENT = 8;

// Thread (ty, tx) of the 16 x 16 grid.  A warp covers WLY consecutive ty and WLX = 32 / WLY consecutive tx, so that a warp-wide
// LDS.128 of the row operand touches WLY and one of the column operand WLX distinct 16-byte pieces: with 4 x 8 both are at
// most 128 bytes = one shared-memory wavefront each (the 2 x 16 arrangement of round 1 read 256 bytes = two wavefronts per
// column-operand load; the Euclidean main loop is co-limited by shared-memory wavefronts).  The eight warps tile the grid 4 x 2.
constexpr int WLY = 4, WLX = 32 / WLY;
__device__ __forceinline__ int thread_ty(int tid) { return WLY * ((tid >> 5) / (16 / WLX)) + ((tid & 31) / WLX); }
__device__ __forceinline__ int thread_tx(int tid) { return WLX * ((tid >> 5) % (16 / WLX)) + ((tid & 31) % WLX); }

// Tile geometry.  Threads form a 16 x 16 grid; thread (ty, tx) owns rows ty + 16 i (i < MI) and columns tx + 16 j (j < MJ).
//   Big   (128 x 64, 8 x 4 micro-tile): the shape the public tile numbering (emk_pair_tile_count / _range / _decode) is in, best
//         operand reuse (12 LDS.128 per 32 pair-float4 products), two CTAs per SM -- large evaluations and tile ranges.
//   Small (64 x 32, 4 x 2 micro-tile): four times as many tiles and a quarter of the shared memory per CTA (3 CTAs per SM), for
//         whole evaluations of up to a few thousand rows, where the big shape leaves SMs idle (1 024 rows: 72 big tiles on 296
//         CTA slots; split over clusters each CTA still ran alone on its SM).
template <int TM_, int TN_>
struct Geom {
  static constexpr int TM = TM_, TN = TN_;
  static constexpr int MI = TM_ / 16, MJ = TN_ / 16;
  static constexpr int BOX_ROWS = TN_;                       // rows per TMA box: one box for the columns, TM / TN for the rows
  static constexpr int STAGE_FLOATS = (TM_ + TN_) * KC;
  static constexpr int STAGE_BYTES = STAGE_FLOATS * 4;       // 24 KB (big), 12 KB (small)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 128 + (MAX_LATENT * (TM_ + 2 * TN_)) * 4;
  static constexpr int MIN_CTAS = TM_ >= 128 ? 2 : 3;
};
using GeomBig = Geom<EMK_TILE_ROWS, EMK_TILE_COLS>;   // 128 x 64
using GeomSmall = Geom<64, 32>;
constexpr int TM = GeomBig::TM;   // the public tile shape (tile ranges, small_cost_kernel)
constexpr int TN = GeomBig::TN;

enum class Epi : int { kCost = 0, kDistMatrix = 1 };

struct PairParams {
  // common
  int64_t n;          // rows
  int n_chunks;       // ceil(d / KC)
  int tiles_per_row;  // Tc = ceil(n / TN)
  int tile_rows;      // Tr = ceil(n / TM)
  int64_t tile_begin;
  float period;       // +inf => Euclidean
  // cost epilogue
  const float* low;   // (n, l)
  int l;
  SigSpec sh, sl;
  double* loss;
  float* grad;        // (n, l) or nullptr
  double loss_scale;  // 1 / n^2
  float grad_scale;   // 4 / n^2
  // distance-matrix epilogue
  float* out;         // (n, n)
  int squared;
  int d_total;        // true feature count (for the all-zero eps of the periodic form)
  int periodic_eps;   // 1: reproduce pairwise_dist_periodic's +1e-12 conventions
};

// ---- PTX helpers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}

// tile id -> (tile row I over 128-row blocks, tile column J over 64-column blocks), J in [2I, Tc).
// Tiles are numbered band by band: a band is BAND consecutive tile rows (1024 rows of `high`, 4 MB at
// D = 1024); inside a band the order is column-major, so CTAs that run together share one column block and
// the band's row blocks stay in L2 -- every column block is fetched from HBM once per band instead of
// once per tile row.
constexpr int BAND = 8;
__host__ __device__ inline int64_t band_prefix(int64_t gb, int64_t tc) {   // tiles before full band gb
  return BAND * gb * (tc - BAND + 1) - (int64_t)BAND * BAND * gb * (gb - 1);
}
__host__ __device__ inline void tile_decode(int64_t t, int64_t tc, int64_t tr, int64_t* I_out, int64_t* J_out) {
  // band index from the quadratic prefix, then fix up
  const double a = (double)BAND * BAND, bq = (double)BAND * (tc - BAND + 1) + a;
  double disc = bq * bq - 4.0 * a * (double)t;
  if (disc < 0) disc = 0;
  int64_t gb = (int64_t)((bq - sqrt(disc)) / (2.0 * a));
  const int64_t nb = (tr + BAND - 1) / BAND;
  if (gb < 0) gb = 0;
  if (gb > nb - 1) gb = nb - 1;
  while (gb > 0 && band_prefix(gb, tc) > t) --gb;
  while (gb + 1 < nb && band_prefix(gb + 1, tc) <= t) ++gb;
  const int64_t I0 = gb * BAND;
  const int64_t R = (tr - I0) < BAND ? (tr - I0) : BAND;
  int64_t u = t - band_prefix(gb, tc);
  if (u < R * (R - 1)) {
    // ragged head of the band: column pair p holds p+1 valid rows per column
    int64_t pq = (int64_t)((sqrt(4.0 * (double)u + 1.0) - 1.0) * 0.5);
    while (pq > 0 && pq * (pq + 1) > u) --pq;
    while ((pq + 1) * (pq + 2) <= u) ++pq;
    const int64_t rem = u - pq * (pq + 1);
    *J_out = 2 * I0 + 2 * pq + (rem >= pq + 1 ? 1 : 0);
    *I_out = I0 + rem % (pq + 1);
  } else {
    u -= R * (R - 1);
    *J_out = 2 * (I0 + R - 1) + u / R;
    *I_out = I0 + u % R;
  }
}

// ---- cost epilogue of NI row groups (rows ty + 16 (i0 + i)) x 4 column groups of this thread ------------------
// NI = 8: the whole micro-tile (no cluster); NI = 8 / S: this CTA's share after the reduce-scatter over a cluster.
// stage latent components [c0, c0 + lc) of the tile's rows / columns (component-major, zero for out-of-range rows)
template <class G = GeomBig>
__device__ __forceinline__ void stage_latent(const PairParams& p, float* zA, float* zB, float* colsum, const int64_t row0,
                                             const int64_t col0, const int c0, const int lc, const int tid) {
  constexpr int TM = G::TM, TN = G::TN;
  for (int idx = tid; idx < lc * TM; idx += NTHREADS) {
    const int c = idx / TM, r = idx - c * TM;
    zA[c * TM + r] = (row0 + r < p.n) ? p.low[(row0 + r) * p.l + c0 + c] : 0.f;
  }
  for (int idx = tid; idx < lc * TN; idx += NTHREADS) {
    const int c = idx / TN, r = idx - c * TN;
    zB[c * TN + r] = (col0 + r < p.n) ? p.low[(col0 + r) * p.l + c0 + c] : 0.f;
    colsum[c * TN + r] = 0.f;
  }
}

template <int NI, class G = GeomBig>
__device__ __forceinline__ void cost_epilogue(const float (&d2h)[NI][G::MJ], const int i0, const PairParams& p, float* zA,
                                              float* zB, float* colsum, double* red_d, const int64_t row0,
                                              const int64_t col0, const bool diag, const int ty, const int tx, const int tid,
                                              const int lane, const int warp) {
  constexpr int TM = G::TM, TN = G::TN, MJ = G::MJ;
  // low-d squared distances, summed in the SAME order as the main loop sums the high-d ones (even
  // components in one fused chain, odd components in the other, then one add): identical inputs and
  // sigmoids on both sides then cancel exactly, as they do in the reference (tests/test_losses.py:897-904).
  // Latent widths above MAX_LATENT (the reference accepts any n_neurons[-1], parameters.py:612) are walked in chunks of
  // MAX_LATENT components that are re-staged into the same shared-memory rows; the usual 2..8-wide latent never loops.
  float dl2[NI][MJ];
#pragma unroll
  for (int i = 0; i < NI; i++)
#pragma unroll
    for (int j = 0; j < MJ; j++) dl2[i][j] = 0.f;
  int staged = 0;   // first component of the chunk that sits in zA / zB
#pragma unroll 1
  for (int c0 = 0; c0 < p.l; c0 += MAX_LATENT) {
    const int lc = min(MAX_LATENT, p.l - c0);
    if (c0 != staged) {
      __syncthreads();
      stage_latent<G>(p, zA, zB, colsum, row0, col0, c0, lc, tid);
      staged = c0;
      __syncthreads();
    }
#pragma unroll 1
    for (int par = 0; par < 2; par++) {
      float part[NI][MJ];
#pragma unroll
      for (int i = 0; i < NI; i++)
#pragma unroll
        for (int j = 0; j < MJ; j++) part[i][j] = 0.f;
#pragma unroll 1
      for (int c = par; c < lc; c += 2) {
        float za[NI], zb[MJ];
#pragma unroll
        for (int i = 0; i < NI; i++) za[i] = zA[c * TM + ty + 16 * (i0 + i)];
#pragma unroll
        for (int j = 0; j < MJ; j++) zb[j] = zB[c * TN + tx + 16 * j];
#pragma unroll
        for (int i = 0; i < NI; i++)
#pragma unroll
          for (int j = 0; j < MJ; j++) {
            const float t = za[i] - zb[j];
            part[i][j] = fmaf(t, t, part[i][j]);
          }
      }
#pragma unroll
      for (int i = 0; i < NI; i++)
#pragma unroll
        for (int j = 0; j < MJ; j++) dl2[i][j] += part[i][j];
    }
  }

  float lsum = 0.f;
  // after this loop dl2 holds the gradient coefficient (s_l - s_h) * s_l'(d_l) / d_l
#pragma unroll
  for (int i = 0; i < NI; i++) {
    const bool rv = row0 + ty + 16 * (i0 + i) < p.n;
#pragma unroll
    for (int j = 0; j < MJ; j++) {
      const bool valid = rv && (col0 + tx + 16 * j < p.n);
      const float sh = sig_eval<false>(d2h[i][j], p.sh, nullptr);
      float w;
      const float sl = sig_eval<true>(dl2[i][j], p.sl, &w);
      float diff = sh - sl;
      if (!valid) diff = 0.f;
      if (dl2[i][j] == 0.f || !valid) w = 0.f;
      lsum = fmaf(diff, diff, lsum);
      dl2[i][j] = -diff * w;
    }
  }

  // loss: per-thread float (<= 32 terms) -> double across the CTA
  double ld = (double)lsum;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ld += __shfl_xor_sync(0xffffffffu, ld, o);
  if (lane == 0) red_d[warp] = ld;
  __syncthreads();
  if (tid == 0) {
    double t = 0;
#pragma unroll
    for (int w = 0; w < NTHREADS / 32; w++) t += red_d[w];
    atomicAdd(p.loss, t * (diag ? 1.0 : 2.0) * p.loss_scale);
  }

  if (p.grad == nullptr) return;
  const float gs = p.grad_scale;
  // gradient: chunks in descending order, so that the chunk staged last by the distance loop is used first
#pragma unroll 1
  for (int c0 = ((p.l - 1) / MAX_LATENT) * MAX_LATENT; c0 >= 0; c0 -= MAX_LATENT) {
    const int lc = min(MAX_LATENT, p.l - c0);
    if (c0 != staged) {
      __syncthreads();   // the previous chunk's column sums have been flushed
      stage_latent<G>(p, zA, zB, colsum, row0, col0, c0, lc, tid);
      staged = c0;
      __syncthreads();
    }
#pragma unroll 1
    for (int c = 0; c < lc; c++) {
      float za[NI], zb[MJ], rs[NI], cs[MJ];
#pragma unroll
      for (int i = 0; i < NI; i++) {
        za[i] = zA[c * TM + ty + 16 * (i0 + i)];
        rs[i] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < MJ; j++) {
        zb[j] = zB[c * TN + tx + 16 * j];
        cs[j] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < NI; i++)
#pragma unroll
        for (int j = 0; j < MJ; j++) {
          const float t = dl2[i][j] * (za[i] - zb[j]);
          rs[i] += t;
          cs[j] -= t;
        }
      // row side: the WLX lanes of a warp that share ty
#pragma unroll
      for (int i = 0; i < NI; i++) {
#pragma unroll
        for (int o = WLX / 2; o > 0; o >>= 1) rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], o);   // lanes that share ty
        const int64_t r = row0 + ty + 16 * (i0 + i);
        if ((lane & (WLX - 1)) == 0 && r < p.n) atomicAdd(&p.grad[r * p.l + c0 + c], rs[i] * gs);   // one per warp and row
      }
      // column side (mirror image of the tile); diagonal tiles already visit both orders
      if (!diag) {
#pragma unroll
        for (int j = 0; j < MJ; j++) {
#pragma unroll
          for (int o = WLX; o < 32; o <<= 1) cs[j] += __shfl_xor_sync(0xffffffffu, cs[j], o);   // lanes that share tx
          if (lane < WLX) atomicAdd(&colsum[c * TN + tx + 16 * j], cs[j]);
        }
      }
    }
    if (!diag) {
      __syncthreads();
      for (int idx = tid; idx < lc * TN; idx += NTHREADS) {
        const int c = idx / TN, r = idx - c * TN;
        if (col0 + r < p.n) atomicAdd(&p.grad[(col0 + r) * p.l + c0 + c], colsum[idx] * gs);
      }
    }
  }
}

// this CTA's share of the cluster's partial sums: row groups [i0, i0 + NI) summed over all ranks through DSMEM
template <int NI, class G>
__device__ __forceinline__ void cluster_reduce_scatter(cooperative_groups::cluster_group& cluster, float* part, const int S, const int i0,
                                                       const int tid, float (&out)[NI][G::MJ]) {
  constexpr int MJ = G::MJ;
#pragma unroll
  for (int i = 0; i < NI; i++)
#pragma unroll
    for (int j = 0; j < MJ; j++) out[i][j] = 0.f;
  for (int r = 0; r < S; r++) {
    const float* peer = cluster.map_shared_rank(part, r);
#pragma unroll
    for (int i = 0; i < NI; i++)
#pragma unroll
      for (int j = 0; j < MJ; j++) out[i][j] += peer[((i0 + i) * MJ + j) * NTHREADS + tid];
  }
}

// CLUSTERED = false: one CTA per tile (large problems; none of the cluster code is compiled in -- the extra epilogue
// variants cost the big kernel 3 % when they shared one instantiation)
template <bool PERIODIC, Epi EPI, bool CLUSTERED, class G>
__global__ void __launch_bounds__(NTHREADS, G::MIN_CTAS) pair_tile_kernel(const __grid_constant__ CUtensorMap tmap, const PairParams p) {
  constexpr int TM = G::TM, TN = G::TN, MI = G::MI, MJ = G::MJ, BOX_ROWS = G::BOX_ROWS, STAGE_FLOATS = G::STAGE_FLOATS,
                STAGE_BYTES = G::STAGE_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  float* stage_base = reinterpret_cast<float*>(smem_raw);
  uint8_t* tail = smem_raw + STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);             // STAGES barriers
  double* red_d = reinterpret_cast<double*>(tail + 64);               // 8 doubles
  float* zA = reinterpret_cast<float*>(tail + 128);                   // [l][TM]
  float* zB = zA + MAX_LATENT * TM;                                   // [l][TN]
  float* colsum = zB + MAX_LATENT * TN;                               // [l][TN]

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int ty = thread_ty(tid);  // 0..15 -> rows ty + 16 i
  const int tx = thread_tx(tid);  // 0..15 -> cols tx + 16 j

  // Small problems (few tiles) are spread over the SMs by splitting the feature axis over a thread-block
  // cluster: the S CTAs of a cluster own the same tile and S interleaved shares of the k-chunks; partial squared
  // distances are summed into rank 0 through distributed shared memory, rank 0 runs the epilogue.
  cg::cluster_group cluster = cg::this_cluster();
  const int S = CLUSTERED ? (int)cluster.num_blocks() : 1;
  const int crank = CLUSTERED ? (int)cluster.block_rank() : 0;
  int64_t I, J;
  tile_decode(p.tile_begin + blockIdx.x / S, p.tiles_per_row, p.tile_rows, &I, &J);
  const int64_t row0 = I * TM;
  const int64_t col0 = J * TN;
  const bool diag = (J >> 1) == I;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; s++) mbar_init(&full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // this CTA's share of the k-chunks: global chunk index = kbase + local index
  const int kbase = (int)((int64_t)p.n_chunks * crank / S);
  const int nk = (int)((int64_t)p.n_chunks * (crank + 1) / S) - kbase;
  auto issue = [&](int kc) {
    const int s = kc % STAGES;
    float* dst = stage_base + s * STAGE_FLOATS;
    const int kx = (kbase + kc) * KC;
    mbar_expect_tx(&full_bar[s], STAGE_BYTES);
#pragma unroll
    for (int q = 0; q < TM / BOX_ROWS; q++) tma_load_2d(dst + q * BOX_ROWS * KC, &tmap, kx, (int)row0 + q * BOX_ROWS, &full_bar[s]);
    tma_load_2d(dst + TM * KC, &tmap, kx, (int)col0, &full_bar[s]);
  };
  if (tid == 0) {
    for (int kc = 0; kc < STAGES - 1 && kc < nk; kc++) issue(kc);
  }

  if (EPI == Epi::kCost) stage_latent<G>(p, zA, zB, colsum, row0, col0, 0, min(MAX_LATENT, p.l), tid);   // first chunk of the latent

  float2 acc[MI][MJ];
#pragma unroll
  for (int i = 0; i < MI; i++)
#pragma unroll
    for (int j = 0; j < MJ; j++) acc[i][j] = make_float2(0.f, 0.f);

  const float P = p.period;
  const int swa = ty & 7, swb = tx & 7;  // TMA SWIZZLE_128B: 16-byte chunk index ^= (row & 7)

  for (int kc = 0; kc < nk; kc++) {
    // every thread is done with chunk kc-1 => its stage may be refilled with chunk kc+STAGES-1
    __syncthreads();
    if (tid == 0 && kc + STAGES - 1 < nk) issue(kc + STAGES - 1);
    const int s = kc % STAGES;
    mbar_wait(&full_bar[s], (kc / STAGES) & 1);
    const float* As = stage_base + s * STAGE_FLOATS;
    const float* Bs = As + TM * KC;
#pragma unroll 1
    for (int k4 = 0; k4 < KC / 4; k4++) {
      float4 av[MI];
#pragma unroll
      for (int i = 0; i < MI; i++) av[i] = *reinterpret_cast<const float4*>(As + (ty + 16 * i) * KC + ((k4 ^ swa) << 2));
#pragma unroll
      for (int j = 0; j < MJ; j++) {
        const float4 bv = *reinterpret_cast<const float4*>(Bs + (tx + 16 * j) * KC + ((k4 ^ swb) << 2));
        const float2 nb0 = make_float2(-bv.x, -bv.y), nb1 = make_float2(-bv.z, -bv.w);
#pragma unroll
        for (int i = 0; i < MI; i++) {
          float2 d0 = __fadd2_rn(make_float2(av[i].x, av[i].y), nb0);
          float2 d1 = __fadd2_rn(make_float2(av[i].z, av[i].w), nb1);
          if (PERIODIC) {
            d0.x = fminf(fabsf(d0.x), P - fabsf(d0.x));
            d0.y = fminf(fabsf(d0.y), P - fabsf(d0.y));
            d1.x = fminf(fabsf(d1.x), P - fabsf(d1.x));
            d1.y = fminf(fabsf(d1.y), P - fabsf(d1.y));
          }
          acc[i][j] = __ffma2_rn(d0, d0, acc[i][j]);
          acc[i][j] = __ffma2_rn(d1, d1, acc[i][j]);
        }
      }
    }
  }

  if (CLUSTERED && S > 1) {
    // partial sums -> own shared memory ([value][thread], conflict-free)
    __syncthreads();   // all TMA data of this CTA has been consumed: the stage buffers are free
    float* part = stage_base;
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
      for (int j = 0; j < MJ; j++) part[(i * MJ + j) * NTHREADS + tid] = acc[i][j].x + acc[i][j].y;
    cluster.sync();
    if (EPI == Epi::kDistMatrix) {
      // distance matrices: rank 0 gathers over DSMEM and writes the tile
      if (crank == 0) {
        for (int r = 1; r < S; r++) {
          const float* peer = cluster.map_shared_rank(part, r);
#pragma unroll
          for (int i = 0; i < MI; i++)
#pragma unroll
            for (int j = 0; j < MJ; j++) acc[i][j].x += peer[(i * MJ + j) * NTHREADS + tid];
        }
      }
      cluster.sync();    // peers stay resident until rank 0 has read their partials
      if (crank != 0) return;
    }
    // cost: reduce-scatter below -- every rank sums and finishes 8/S of the 8 row groups, so the epilogue (sigmoids,
    // gradient reductions, atomics) is spread over the cluster instead of idling S-1 CTAs behind rank 0
  }

  // ------------------------------------------------------------------------------------------
  if (EPI == Epi::kDistMatrix) {
#pragma unroll
    for (int i = 0; i < MI; i++) {
      const int64_t r = row0 + ty + 16 * i;
#pragma unroll
      for (int j = 0; j < MJ; j++) {
        const int64_t c = col0 + tx + 16 * j;
        if (r < p.n && c < p.n) {
          float v = acc[i][j].x + acc[i][j].y;
          if (p.periodic_eps) {
            // pairwise_dist_periodic: zero components become 1e-12 before squaring, result + 1e-12
            // (encodermap/misc/distances.py:169-175); only the all-zero case is above float32 resolution
            v = (v == 0.f) ? sqrtf((float)p.d_total) * 1e-12f + 1e-12f : sqrtf(v) + 1e-12f;
          } else if (!p.squared) {
            v = sqrtf(v);
          }
          p.out[r * p.n + c] = v;
          if (!diag) p.out[c * p.n + r] = v;
        }
      }
    }
    return;
  }

  // ---- cost epilogue --------------------------------------------------------------------------
  if (!CLUSTERED || S == 1) {
    float d2h[MI][MJ];
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
      for (int j = 0; j < MJ; j++) d2h[i][j] = acc[i][j].x + acc[i][j].y;
    cost_epilogue<MI, G>(d2h, 0, p, zA, zB, colsum, red_d, row0, col0, diag, ty, tx, tid, lane, warp);
  } else if (S == 2) {
    constexpr int NI = MI / 2;
    float d2h[NI][MJ];
    cluster_reduce_scatter<NI, G>(cluster, stage_base, S, NI * crank, tid, d2h);
    cluster.sync();   // every rank has read what it needs: shared memory may be released
    cost_epilogue<NI, G>(d2h, NI * crank, p, zA, zB, colsum, red_d, row0, col0, diag, ty, tx, tid, lane, warp);
  } else if (S == 4) {
    constexpr int NI = MI / 4;
    float d2h[NI][MJ];
    cluster_reduce_scatter<NI, G>(cluster, stage_base, S, NI * crank, tid, d2h);
    cluster.sync();
    cost_epilogue<NI, G>(d2h, NI * crank, p, zA, zB, colsum, red_d, row0, col0, diag, ty, tx, tid, lane, warp);
  } else if (MI >= 8) {   // S == 8: one row group per rank (big tiles only; pick_cluster never asks a small tile for it)
    constexpr int NI = MI >= 8 ? MI / 8 : 1;
    float d2h[NI][MJ];
    cluster_reduce_scatter<NI, G>(cluster, stage_base, S, NI * crank, tid, d2h);
    cluster.sync();
    cost_epilogue<NI, G>(d2h, NI * crank, p, zA, zB, colsum, red_d, row0, col0, diag, ty, tx, tid, lane, warp);
  }
}

// ---- narrow inputs (D <= SMALL_D: the cube example's 3-d points, 2-d projections) ---------------------------------------
// No TMA, no padded copy, no 32-wide k-chunk that is mostly zeros: the D columns of the tile's 128 + 64 rows are staged
// component-major in shared memory with plain loads and every thread forms its squared distances in registers.  The tile
// numbering, the micro-tile layout (rows ty + 16 i, columns tx + 16 j), the even/odd summation order and the epilogue are
// those of pair_tile_kernel, so tile ranges (multi-GPU) and results carry over.  A launch with few tiles (a training batch
// of 256 rows is 6 tiles) is spread over 8 / NI CTAs per tile, each finishing NI of the eight 16-row groups: there is no
// feature axis worth splitting, so no cluster and no reduction is needed.  The epilogue (two sigmoids per pair, MUFU) is
// the whole cost.
constexpr int SMALL_D = 8;

template <bool PERIODIC, int NI>
__global__ void __launch_bounds__(NTHREADS) small_cost_kernel(const PairParams p, const float* __restrict__ high, const int d) {
  __shared__ float hA[SMALL_D * TM];
  __shared__ float hB[SMALL_D * TN];
  __shared__ float zA[MAX_LATENT * TM];
  __shared__ float zB[MAX_LATENT * TN];
  __shared__ float colsum[MAX_LATENT * TN];
  __shared__ double red_d[NTHREADS / 32];
  constexpr int S = 8 / NI;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, ty = thread_ty(tid), tx = thread_tx(tid);
  const int i0 = (int)(blockIdx.x % S) * NI;
  int64_t I, J;
  tile_decode(p.tile_begin + blockIdx.x / S, p.tiles_per_row, p.tile_rows, &I, &J);
  const int64_t row0 = I * TM, col0 = J * TN;
  const bool diag = (J >> 1) == I;
  for (int idx = tid; idx < d * TM; idx += NTHREADS) {
    const int r = idx / d, k = idx - r * d;      // consecutive threads read consecutive floats of the row block
    hA[k * TM + r] = (row0 + r < p.n) ? high[(row0 + r) * d + k] : 0.f;
  }
  for (int idx = tid; idx < d * TN; idx += NTHREADS) {
    const int r = idx / d, k = idx - r * d;
    hB[k * TN + r] = (col0 + r < p.n) ? high[(col0 + r) * d + k] : 0.f;
  }
  stage_latent(p, zA, zB, colsum, row0, col0, 0, min(MAX_LATENT, p.l), tid);
  __syncthreads();
  const float P = p.period;
  float even[NI][4], odd[NI][4];
#pragma unroll
  for (int i = 0; i < NI; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) even[i][j] = odd[i][j] = 0.f;
#pragma unroll 1
  for (int k = 0; k < d; k++) {
    float a[NI], b[4];
#pragma unroll
    for (int i = 0; i < NI; i++) a[i] = hA[k * TM + ty + 16 * (i0 + i)];
#pragma unroll
    for (int j = 0; j < 4; j++) b[j] = hB[k * TN + tx + 16 * j];
#pragma unroll
    for (int i = 0; i < NI; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        float t = a[i] - b[j];
        if (PERIODIC) t = fminf(fabsf(t), P - fabsf(t));
        if (k & 1) odd[i][j] = fmaf(t, t, odd[i][j]);
        else even[i][j] = fmaf(t, t, even[i][j]);
      }
  }
  float d2h[NI][4];
#pragma unroll
  for (int i = 0; i < NI; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) d2h[i][j] = even[i][j] + odd[i][j];
  cost_epilogue<NI>(d2h, i0, p, zA, zB, colsum, red_d, row0, col0, diag, ty, tx, tid, lane, warp);
}

// emk_set_option("cost_small_d_max", v): widest input that takes the register path (default SMALL_D; 0 sends everything
// through the TMA kernel -- the parity tests compare the two)
static int g_small_d_max = SMALL_D;
void set_cost_small_d_max(int64_t v) { g_small_d_max = (int)std::max<int64_t>(0, std::min<int64_t>(SMALL_D, v)); }
int64_t cost_small_d_max() { return g_small_d_max; }

template <bool PERIODIC>
static int launch_small_cost(const PairParams& p, const float* high, int d, int64_t n_tiles, cudaStream_t st) {
  // CTAs per tile: enough to give every SM a few CTAs while a tile's fixed work (staging, one loss atomic, the column
  // flush) is not repeated more often than needed
  const int64_t slots = 4 * (int64_t)sm_count();
  if (n_tiles * 8 <= slots) small_cost_kernel<PERIODIC, 1><<<(unsigned)(n_tiles * 8), NTHREADS, 0, st>>>(p, high, d);
  else if (n_tiles * 4 <= slots) small_cost_kernel<PERIODIC, 2><<<(unsigned)(n_tiles * 4), NTHREADS, 0, st>>>(p, high, d);
  else if (n_tiles * 2 <= slots) small_cost_kernel<PERIODIC, 4><<<(unsigned)(n_tiles * 2), NTHREADS, 0, st>>>(p, high, d);
  else small_cost_kernel<PERIODIC, 8><<<(unsigned)n_tiles, NTHREADS, 0, st>>>(p, high, d);
  return launch_status("small_cost_kernel");
}

// ---- host side -----------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode_fn(EncodeTiledFn* fn) {
  static EncodeTiledFn cached = nullptr;
  if (!cached) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    EMK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
    EMK_REQUIRE(sym != nullptr && q == cudaDriverEntryPointSuccess, EMK_E_UNSUPPORTED,
                "driver does not export cuTensorMapEncodeTiled");
    cached = reinterpret_cast<EncodeTiledFn>(sym);
  }
  *fn = cached;
  return EMK_OK;
}

// generic 2-d float32 tiled map (rows x cols, row pitch in bytes a multiple of 16) -- also used by backmap.cu for its
// output rows.  `map_out` is a CUtensorMap (128 bytes, 64-byte aligned).
int encode_f32_map_2d(void* map_out, const float* base, uint64_t cols, uint64_t rows, uint64_t row_pitch_bytes, uint32_t box_cols,
                      uint32_t box_rows, int swizzle_128b) {
  EncodeTiledFn enc;
  int rc = get_encode_fn(&enc);
  if (rc) return rc;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(static_cast<CUtensorMap*>(map_out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_128b ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EMK_REQUIRE(r == CUDA_SUCCESS, EMK_E_ARG, "cuTensorMapEncodeTiled failed with CUresult %d (cols=%llu rows=%llu pitch=%llu)", (int)r,
              (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)row_pitch_bytes);
  return EMK_OK;
}

// (n, d_pad) row-major float32, 16-byte aligned base, d_pad % 4 == 0
static int make_tensor_map(CUtensorMap* map, const float* base, int64_t n, int64_t d_pad, int box_rows) {
  EncodeTiledFn enc;
  int rc = get_encode_fn(&enc);
  if (rc) return rc;
  cuuint64_t dims[2] = {(cuuint64_t)d_pad, (cuuint64_t)n};
  cuuint64_t strides[1] = {(cuuint64_t)d_pad * sizeof(float)};
  cuuint32_t box[2] = {KC, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EMK_REQUIRE(r == CUDA_SUCCESS, EMK_E_ARG, "cuTensorMapEncodeTiled failed with CUresult %d (n=%lld d=%lld)", (int)r,
              (long long)n, (long long)d_pad);
  return EMK_OK;
}

__global__ void pad_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n, int64_t d, int64_t d_pad) {
  const int64_t total = n * d_pad;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / d_pad, k = idx - r * d_pad;
    dst[idx] = k < d ? src[r * d + k] : 0.f;
  }
}

// TMA needs a 16-byte aligned base and a row pitch that is a multiple of 16 bytes; anything else is
// re-laid out once into a stream-ordered scratch buffer (O(N D), negligible against O(N^2 D)).
struct HighView {
  const float* ptr;
  int64_t d_pad;
  float* scratch;
};

static int prepare_high(const float* high, int64_t n, int64_t d, cudaStream_t st, HighView* v) {
  v->scratch = nullptr;
  if ((d % 4) == 0 && (reinterpret_cast<uintptr_t>(high) % 16) == 0) {
    v->ptr = high;
    v->d_pad = d;
    return EMK_OK;
  }
  const int64_t d_pad = (d + 3) / 4 * 4;
  {
    int rc0 = scratch_alloc(reinterpret_cast<void**>(&v->scratch), (size_t)(n * d_pad) * sizeof(float), st);
    if (rc0) return rc0;
  }
  const int64_t total = n * d_pad;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  pad_rows_kernel<<<blocks, 256, 0, st>>>(high, v->scratch, n, d, d_pad);
  int rc = launch_status("pad_rows_kernel");
  if (rc) return rc;
  v->ptr = v->scratch;
  v->d_pad = d_pad;
  return EMK_OK;
}

// Cluster size (K-split) for a launch of `n_tiles` tiles with `n_chunks` k-chunks each.  Every rank of a cluster runs
// 1/S of the main loop AND 1/S of the epilogue (reduce-scatter over DSMEM), so the cost of splitting is the per-CTA
// fixed work (barrier set-up, latent staging, pipeline fill, two cluster syncs): measured ~0 % at S = 2, 9 % at 4,
// 27 % at 8 on 4096 x 1024.  Measured on B200 (tools/bench_small_cost.py, EMK_CLUSTER sweep):
//   * up to 37 tiles (N <= 640): S = 8 fills at most the 296 CTA slots            256 x 1024: 185 -> 38 us
//   * up to 127 tiles: the smallest S that gives every SM a CTA (clusters of 4 / 8 no longer fit in one wave: GPCs
//     hold whole clusters only)                                                   1024 x 1024: 186 -> 115 us, 1024 x 4950: 450 -> 247 us
//   * 297 .. 2048 tiles: S = 2 halves the work unit of the ragged last wave       4096 x 1024: 1205 -> 1168 us
//   * otherwise one CTA per tile                                                  8192 x 1024: 4265 us (S = 2: 4358)
static int pick_cluster(int64_t n_tiles, int n_chunks) {
  int s = 1;
  static const int forced = [] {   // experiments only: EMK_CLUSTER forces a cluster size (1, 2, 4, 8); read once per process
    const char* e = getenv("EMK_CLUSTER");
    const int v = e ? atoi(e) : 0;
    return (v >= 1 && v <= 8 && (v & (v - 1)) == 0) ? v : 0;
  }();
  if (forced) {
    s = forced;
  } else if (n_tiles * 8 <= 2 * (int64_t)sm_count()) {
    s = 8;
  } else if (n_tiles < 128) {
    while (s < 8 && n_tiles * s < 128) s *= 2;
  } else if (n_tiles > 2 * (int64_t)sm_count() && n_tiles <= 2048) {
    s = 2;
  }
  while (s > 1 && 2 * s > n_chunks) s /= 2;   // every rank needs at least two k-chunks
  return s;
}
// small tiles (64 x 32, three CTAs per SM, at most MI = 4 ranks per cluster): split the feature axis until the launch has more
// than half as many CTAs as the machine has slots
static int pick_cluster_small(int64_t n_tiles, int n_chunks) {
  const int64_t slots = (int64_t)GeomSmall::MIN_CTAS * sm_count();
  int s = 1;
  while (s < GeomSmall::MI && n_tiles * s * 2 <= slots) s *= 2;
  while (s > 1 && 2 * s > n_chunks) s /= 2;
  return s;
}

template <bool PERIODIC, Epi EPI, bool CLUSTERED, class G>
static int launch_pair_c(const CUtensorMap& map, const PairParams& p, int64_t n_tiles, int cluster, cudaStream_t st) {
  auto kern = pair_tile_kernel<PERIODIC, EPI, CLUSTERED, G>;
  constexpr int SMEM_BYTES = G::SMEM_BYTES;
  static bool configured[kMaxDevices] = {false};
  if (first_use_on_device(configured)) {
    EMK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    EMK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(n_tiles * cluster));
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  EMK_CUDA(cudaLaunchKernelEx(&cfg, kern, map, p));
  return launch_status("pair_tile_kernel");
}
template <bool PERIODIC, Epi EPI>
static int launch_pair(const CUtensorMap& map, const PairParams& p, int64_t n_tiles, cudaStream_t st) {
  const int cluster = pick_cluster(n_tiles, p.n_chunks);
  return cluster > 1 ? launch_pair_c<PERIODIC, EPI, true, GeomBig>(map, p, n_tiles, cluster, st)
                     : launch_pair_c<PERIODIC, EPI, false, GeomBig>(map, p, n_tiles, 1, st);
}
template <bool PERIODIC>
static int launch_pair_small(const CUtensorMap& map, const PairParams& p, int64_t n_tiles, cudaStream_t st) {
  const int cluster = pick_cluster_small(n_tiles, p.n_chunks);
  return cluster > 1 ? launch_pair_c<PERIODIC, Epi::kCost, true, GeomSmall>(map, p, n_tiles, cluster, st)
                     : launch_pair_c<PERIODIC, Epi::kCost, false, GeomSmall>(map, p, n_tiles, 1, st);
}
// emk_set_option("cost_small_tile_max_rows", v): whole evaluations of up to v rows (and up to SMALL_TILE_MAX_COLS columns) use
// the 64 x 32 tile shape (0 = never).  Measured on B200, device time, small / big tiles (tools/bench_small_cost.py):
//   256 x 1024: 23 / 37 us    512 x 1024: 43 / 53    1024 x 1024: 97 / 113    2048 x 1024: 336 / 329    3072 x 1024: 655 / 710
//   4096 x 1024: 1166 / 1168    8192 x 1024: 4462 / 4276    1024 x 4950: 270 / 234 (a long feature axis favours operand reuse)
static int64_t g_small_tile_max_rows = 1024;
constexpr int64_t SMALL_TILE_MAX_COLS = 2048;
void set_cost_small_tile_max_rows(int64_t v) { g_small_tile_max_rows = v < 0 ? 0 : v; }
int64_t cost_small_tile_max_rows() { return g_small_tile_max_rows; }

void tile_decode_host(int64_t t, int64_t n_rows, int64_t* I, int64_t* J) {
  tile_decode(t, (n_rows + TN - 1) / TN, (n_rows + TM - 1) / TM, I, J);
}

int64_t pair_tile_count(int64_t n) {
  if (n <= 0) return 0;
  const int64_t tr = (n + TM - 1) / TM, tc = (n + TN - 1) / TN;
  return tr * tc - tr * (tr - 1);
}

int sigmoid_cost_device(const float* high, int64_t n, int64_t d, const float* low, int64_t l, double periodicity,
                        const float sig[6], int64_t tile_begin, int64_t tile_end, double* loss, float* grad_low,
                        uint32_t flags, cudaStream_t st) {
  EMK_REQUIRE(sig && loss && (n == 0 || (high && low)), EMK_E_NULL, "emk_sigmoid_cost: NULL pointer argument");
  EMK_REQUIRE((flags & EMK_COST_NO_GRAD) || grad_low || n == 0, EMK_E_NULL, "emk_sigmoid_cost: grad_low is NULL without EMK_COST_NO_GRAD");
  EMK_REQUIRE(n >= 0 && d >= 1, EMK_E_SHAPE, "emk_sigmoid_cost: bad shape n=%lld d=%lld", (long long)n, (long long)d);
  EMK_REQUIRE(l >= 1 && l < (1 << 20), EMK_E_SHAPE, "emk_sigmoid_cost: latent width %lld outside [1,2^20)", (long long)l);
  EMK_REQUIRE(n < (int64_t)1 << 30, EMK_E_UNSUPPORTED, "emk_sigmoid_cost: n=%lld too large", (long long)n);
  EMK_REQUIRE(periodicity > 0 || std::isinf(periodicity), EMK_E_ARG, "emk_sigmoid_cost: periodicity must be > 0");
  for (int k = 0; k < 6; k++) EMK_REQUIRE(sig[k] > 0.f && std::isfinite(sig[k]), EMK_E_ARG, "emk_sigmoid_cost: sig[%d]=%g must be finite and > 0", k, sig[k]);
  const int64_t total = pair_tile_count(n);
  EMK_REQUIRE(tile_begin >= 0 && tile_begin <= tile_end && tile_end <= total, EMK_E_ARG,
              "emk_sigmoid_cost: tile range [%lld,%lld) outside [0,%lld]", (long long)tile_begin, (long long)tile_end, (long long)total);
  const bool want_grad = !(flags & EMK_COST_NO_GRAD);
  if (flags & EMK_COST_ZERO_OUTPUTS) {
    EMK_CUDA(cudaMemsetAsync(loss, 0, sizeof(double), st));
    if (want_grad && n > 0) EMK_CUDA(cudaMemsetAsync(grad_low, 0, (size_t)(n * l) * sizeof(float), st));
  }
  if (tile_end == tile_begin) return EMK_OK;

  PairParams p{};
  p.n = n;
  p.n_chunks = (int)((d + KC - 1) / KC);
  p.tiles_per_row = (int)((n + TN - 1) / TN);
  p.tile_rows = (int)((n + TM - 1) / TM);
  p.tile_begin = tile_begin;
  p.period = std::isinf(periodicity) ? INFINITY : (float)periodicity;
  p.low = low;
  p.l = (int)l;
  p.sh = make_sig_spec(sig[0], sig[1], sig[2]);
  p.sl = make_sig_spec(sig[3], sig[4], sig[5]);
  p.loss = loss;
  p.grad = want_grad ? grad_low : nullptr;
  p.loss_scale = 1.0 / ((double)n * (double)n);
  p.grad_scale = (float)(4.0 / ((double)n * (double)n));
  if (d <= g_small_d_max) {   // narrow inputs: register path, no TMA, no padded copy
    return std::isinf(periodicity) ? launch_small_cost<false>(p, high, (int)d, tile_end - tile_begin, st)
                                   : launch_small_cost<true>(p, high, (int)d, tile_end - tile_begin, st);
  }

  HighView hv;
  int rc = prepare_high(high, n, d, st, &hv);
  if (rc) return rc;
  CUtensorMap map;
  // a whole evaluation of a training-size batch: the small tile shape (its own tile numbering, from 0)
  const bool small_tiles = tile_begin == 0 && tile_end == total && n <= g_small_tile_max_rows &&
                           (d <= SMALL_TILE_MAX_COLS || g_small_tile_max_rows >= (1 << 20));   // 2^20: forced (tests)
  rc = make_tensor_map(&map, hv.ptr, n, hv.d_pad, small_tiles ? GeomSmall::BOX_ROWS : GeomBig::BOX_ROWS);
  if (rc == EMK_OK) {
    if (small_tiles) {
      p.tiles_per_row = (int)((n + GeomSmall::TN - 1) / GeomSmall::TN);
      p.tile_rows = (int)((n + GeomSmall::TM - 1) / GeomSmall::TM);
      const int64_t tiles = (int64_t)p.tile_rows * p.tiles_per_row - (int64_t)p.tile_rows * (p.tile_rows - 1);
      rc = std::isinf(periodicity) ? launch_pair_small<false>(map, p, tiles, st) : launch_pair_small<true>(map, p, tiles, st);
    } else if (std::isinf(periodicity)) {
      rc = launch_pair<false, Epi::kCost>(map, p, tile_end - tile_begin, st);
    } else {
      rc = launch_pair<true, Epi::kCost>(map, p, tile_end - tile_begin, st);
    }
  }
  if (hv.scratch) cudaFreeAsync(hv.scratch, st);
  return rc;
}

// (n,n) distance matrix through the same main loop
int dist_matrix_device(const float* x, int64_t n, int64_t d, double periodicity, bool periodic_form, int squared, float* out,
                       cudaStream_t st) {
  EMK_REQUIRE(n == 0 || (x && out), EMK_E_NULL, "distance matrix: NULL pointer argument");
  EMK_REQUIRE(n >= 0 && d >= 1, EMK_E_SHAPE, "distance matrix: bad shape n=%lld d=%lld", (long long)n, (long long)d);
  EMK_REQUIRE(n < (int64_t)1 << 30, EMK_E_UNSUPPORTED, "distance matrix: n=%lld too large", (long long)n);
  if (n == 0) return EMK_OK;
  HighView hv;
  int rc = prepare_high(x, n, d, st, &hv);
  if (rc) return rc;
  CUtensorMap map;
  rc = make_tensor_map(&map, hv.ptr, n, hv.d_pad, GeomBig::BOX_ROWS);
  if (rc == EMK_OK) {
    PairParams p{};
    p.n = n;
    p.n_chunks = (int)((d + KC - 1) / KC);
    p.tiles_per_row = (int)((n + TN - 1) / TN);
    p.tile_rows = (int)((n + TM - 1) / TM);
    p.tile_begin = 0;
    p.period = std::isinf(periodicity) ? INFINITY : (float)periodicity;
    p.out = out;
    p.squared = squared;
    p.d_total = (int)d;
    p.periodic_eps = periodic_form ? 1 : 0;
    const int64_t tiles = pair_tile_count(n);
    if (std::isinf(periodicity))
      rc = launch_pair<false, Epi::kDistMatrix>(map, p, tiles, st);
    else
      rc = launch_pair<true, Epi::kDistMatrix>(map, p, tiles, st);
  }
  if (hv.scratch) cudaFreeAsync(hv.scratch, st);
  return rc;
}

}  // namespace emk
