// Generation-side back-mapping (the part of it that is tensor arithmetic): amide hydrogens and
// carbonyl oxygens guessed from the backbone, and the merge of backbone + guessed atoms into one coordinate array --
// encodermap/misc/backmapping.py:1920-1990 (guess_sp2_atom, guess_amide_H, guess_amide_O, merge_cartesians; TF1 twins
// encodermap/encodermap_tf1/backmapping.py:256-318).  The reference builds them from ~10 TensorFlow ops per guessed atom
// inside Python loops over the atoms; here ONE launch writes the merged (frames, n + n_H + n_O, 3) array straight from the
// backbone, one thread per output atom, following a small "plan" (one 32-bit word per output atom: what it is and where it
// comes from) that the host builds from the index lists with the reference's own loop.  HBM-bound: 12 n bytes read and
// 12 (n + n_H + n_O) written per frame.
//
// guess_sp2_atom for centre atom i (reference order of operations):
//   prev = x[i-1] - x[i];  next = x[i+1] - x[i]   (x[i-2] - x[i] when i is the last atom)
//   u = prev x next / |prev x next|
//   bond = prev . R(u, angle)     R = cos I + sin [u]x + (1 - cos) u u^T, row vector on the left:  cos prev + sin (prev x u) + (1 - cos)(prev . u) u
//   out = x[i] + bond * bond_length / |bond|
// Negative indices wrap as Python / TensorFlow slicing does (i - 1 at i = 0 is the last atom).
#include <vector>

#include "emk_common.cuh"

namespace emk {

enum PlanKind : uint32_t { kCopy0 = 0, kGuessA = 1, kGuessB = 2, kCopy1 = 3, kCopy2 = 4 };
constexpr int PLAN_SHIFT = 28;
constexpr uint32_t PLAN_MASK = (1u << PLAN_SHIFT) - 1;

struct Sp2Params { float c, s, len; };

struct GenParams {
  const float* src0;   // (b, n0, 3) backbone
  const float* src1;   // (b, n1, 3) or null (merge of given hydrogens)
  const float* src2;   // (b, n2, 3) or null
  int n0, n1, n2;
  const uint32_t* plan;
  int n_out;
  int64_t b;
  Sp2Params a, bb;
  float* out;          // (b, n_out, 3)
};

__device__ __forceinline__ void sp2_atom(const float* __restrict__ xf, int n, int i, const Sp2Params& q, float* o) {
  const int ip = i - 1 < 0 ? i - 1 + n : i - 1;
  int in = i + 1;
  if (in >= n) in = i - 2 < 0 ? i - 2 + n : i - 2;
  const float cx = xf[3 * i], cy = xf[3 * i + 1], cz = xf[3 * i + 2];
  const float px = xf[3 * ip] - cx, py = xf[3 * ip + 1] - cy, pz = xf[3 * ip + 2] - cz;
  const float nx = xf[3 * in] - cx, ny = xf[3 * in + 1] - cy, nz = xf[3 * in + 2] - cz;
  float ux = py * nz - pz * ny, uy = pz * nx - px * nz, uz = px * ny - py * nx;
  const float inv = rsqrt_fast(ux * ux + uy * uy + uz * uz);      // one MUFU (2 ulp): IEEE sqrt + divide are ~30 instructions
  ux *= inv; uy *= inv; uz *= inv;
  const float wx = py * uz - pz * uy, wy = pz * ux - px * uz, wz = px * uy - py * ux;   // prev x u
  const float pu = (1.f - q.c) * (px * ux + py * uy + pz * uz);
  float bx = q.c * px + q.s * wx + pu * ux, by = q.c * py + q.s * wy + pu * uy, bz = q.c * pz + q.s * wz + pu * uz;
  const float sc = q.len * rsqrt_fast(bx * bx + by * by + bz * bz);
  o[0] = cx + bx * sc; o[1] = cy + by * sc; o[2] = cz + bz * sc;
}

// grid: x over chunks of 256 output atoms, y over frames (grid-stride): no index division; the plan entry of a thread is the
// same for every frame it visits
__global__ void __launch_bounds__(256) generate_atoms_kernel(const GenParams p) {
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k >= p.n_out) return;
  const uint32_t e = __ldg(p.plan + k);
  for (int64_t f = blockIdx.y; f < p.b; f += gridDim.y) {
    const int64_t idx = f * p.n_out + k;
    const uint32_t kind = e >> PLAN_SHIFT;
    const int i = (int)(e & PLAN_MASK);
    float o[3];
    if (kind == kCopy0 || kind == kCopy1 || kind == kCopy2) {
      const float* s = kind == kCopy0 ? p.src0 + (f * p.n0 + i) * 3 : (kind == kCopy1 ? p.src1 + (f * p.n1 + i) * 3 : p.src2 + (f * p.n2 + i) * 3);
      o[0] = s[0]; o[1] = s[1]; o[2] = s[2];
    } else {
      sp2_atom(p.src0 + f * p.n0 * 3, p.n0, i, kind == kGuessA ? p.a : p.bb, o);
    }
    float* d = p.out + idx * 3;
    d[0] = o[0]; d[1] = o[1]; d[2] = o[2];
  }
}

static int run_plan(GenParams p, const std::vector<uint32_t>& plan, cudaStream_t st) {
  p.n_out = (int)plan.size();
  if (p.b == 0 || p.n_out == 0) return EMK_OK;
  uint32_t* dplan = nullptr;
  int rc = scratch_alloc(reinterpret_cast<void**>(&dplan), plan.size() * sizeof(uint32_t), st);
  if (rc) return rc;
  // pageable source: the runtime stages it before the call returns, so `plan` may go out of scope afterwards
  cudaError_t e = cudaMemcpyAsync(dplan, plan.data(), plan.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) {
    cudaFreeAsync(dplan, st);
    return fail((int)e, "generation plan upload failed: %s", cudaGetErrorString(e));
  }
  p.plan = dplan;
  const unsigned gx = (unsigned)((p.n_out + 255) / 256);
  const unsigned gy = (unsigned)std::max<int64_t>(1, std::min<int64_t>(p.b, std::min<int64_t>(65535, (int64_t)sm_count() * 16 / gx)));
  generate_atoms_kernel<<<dim3(gx, gy), 256, 0, st>>>(p);
  rc = launch_status("generate_atoms_kernel");
  cudaFreeAsync(dplan, st);
  return rc;
}

static Sp2Params sp2_params(double angle, double len) { return Sp2Params{(float)std::cos(angle), (float)std::sin(angle), (float)len}; }

static int check_indices(const char* what, const int64_t* idx, int64_t count, int64_t n) {
  EMK_REQUIRE(count == 0 || idx, EMK_E_NULL, "%s: NULL index list", what);
  for (int64_t k = 0; k < count; k++)
    EMK_REQUIRE(idx[k] >= 0 && idx[k] < n, EMK_E_ARG, "%s: index %lld outside [0,%lld)", what, (long long)idx[k], (long long)n);
  return EMK_OK;
}

// guess_sp2_atom(cartesians, indices, angle_to_previous, bond_length) -> (b, n_idx, 3)
int guess_sp2_device(const float* xyz, int64_t b, int64_t n, const int64_t* indices, int64_t n_idx, double angle, double bond_length,
                     float* out, cudaStream_t st) {
  EMK_REQUIRE((xyz && out) || b == 0 || (xyz && n_idx == 0), EMK_E_NULL, "emk_guess_sp2_atoms: NULL pointer argument");
  EMK_REQUIRE(b >= 0 && n >= 3 && n < (int64_t)PLAN_MASK, EMK_E_SHAPE, "emk_guess_sp2_atoms: need (b, n >= 3, 3) coordinates");
  int rc = check_indices("emk_guess_sp2_atoms", indices, n_idx, n);
  if (rc) return rc;
  std::vector<uint32_t> plan((size_t)n_idx);
  for (int64_t k = 0; k < n_idx; k++) plan[(size_t)k] = ((uint32_t)kGuessA << PLAN_SHIFT) | (uint32_t)indices[k];
  GenParams p{};
  p.src0 = xyz; p.n0 = (int)n; p.b = b; p.a = sp2_params(angle, bond_length); p.bb = p.a; p.out = out;
  return run_plan(p, plan, st);
}

// The merge loop of the reference (misc/backmapping.py:1970-1990): atom 0, then for i = 1 .. n-1: atom i, followed by the next
// hydrogen if i is in h_after, ELSE by the next oxygen if i is in o_after.  `h_after` / `o_after` are what the reference
// tests membership against (N_indices[1:] and O_indices in the TF2 form, the positions of "N" / "C" in the TF1 form).
static int merge_plan(int64_t n, const int64_t* h_after, int64_t nh_after, const int64_t* o_after, int64_t no_after, bool guess,
                      std::vector<uint32_t>* plan, int64_t* nh_used, int64_t* no_used) {
  std::vector<uint8_t> in_h((size_t)n, 0), in_o((size_t)n, 0);
  for (int64_t k = 0; k < nh_after; k++) in_h[(size_t)h_after[k]] = 1;
  for (int64_t k = 0; k < no_after; k++) in_o[(size_t)o_after[k]] = 1;
  plan->clear();
  plan->push_back(((uint32_t)kCopy0 << PLAN_SHIFT) | 0u);
  int64_t h = 0, o = 0;
  for (int64_t i = 1; i < n; i++) {
    plan->push_back(((uint32_t)kCopy0 << PLAN_SHIFT) | (uint32_t)i);
    if (in_h[(size_t)i]) {
      plan->push_back(guess ? (((uint32_t)kGuessA << PLAN_SHIFT) | (uint32_t)i) : (((uint32_t)kCopy1 << PLAN_SHIFT) | (uint32_t)h));
      ++h;
    } else if (in_o[(size_t)i]) {
      plan->push_back(guess ? (((uint32_t)kGuessB << PLAN_SHIFT) | (uint32_t)i) : (((uint32_t)kCopy2 << PLAN_SHIFT) | (uint32_t)o));
      ++o;
    }
  }
  *nh_used = h;
  *no_used = o;
  return EMK_OK;
}

int merge_cartesians_device(const float* central, int64_t b, int64_t n, const int64_t* h_after, int64_t nh_after, const int64_t* o_after,
                            int64_t no_after, const float* h_xyz, int64_t n_h, const float* o_xyz, int64_t n_o, float* out,
                            cudaStream_t st) {
  EMK_REQUIRE(b == 0 || (central && out && (n_h == 0 || h_xyz) && (n_o == 0 || o_xyz)), EMK_E_NULL, "emk_merge_cartesians: NULL pointer argument");
  EMK_REQUIRE(b >= 0 && n >= 1 && n < (int64_t)PLAN_MASK && n_h >= 0 && n_o >= 0, EMK_E_SHAPE, "emk_merge_cartesians: bad shape");
  int rc = check_indices("emk_merge_cartesians (h_after)", h_after, nh_after, n);
  if (rc) return rc;
  rc = check_indices("emk_merge_cartesians (o_after)", o_after, no_after, n);
  if (rc) return rc;
  std::vector<uint32_t> plan;
  int64_t h, o;
  merge_plan(n, h_after, nh_after, o_after, no_after, false, &plan, &h, &o);
  // the reference's closing assert: every supplied atom is used exactly once (h > n_h would already have raised in the loop)
  EMK_REQUIRE(h == n_h && o == n_o, EMK_E_SHAPE,
              "emk_merge_cartesians: the index lists place %lld hydrogens and %lld oxygens, %lld and %lld were supplied", (long long)h,
              (long long)o, (long long)n_h, (long long)n_o);
  GenParams p{};
  p.src0 = central; p.src1 = h_xyz; p.src2 = o_xyz; p.n0 = (int)n; p.n1 = (int)n_h; p.n2 = (int)n_o; p.b = b; p.out = out;
  return run_plan(p, plan, st);
}

// guess_amide_H + guess_amide_O + merge_cartesians in one launch: the H / O arrays never exist
int backbone_amide_device(const float* central, int64_t b, int64_t n, const int64_t* h_after, int64_t nh_after, const int64_t* o_after,
                          int64_t no_after, double h_angle, double h_len, double o_angle, double o_len, float* out, int64_t n_out,
                          cudaStream_t st) {
  EMK_REQUIRE(b == 0 || (central && out), EMK_E_NULL, "emk_backbone_amide_atoms: NULL pointer argument");
  EMK_REQUIRE(b >= 0 && n >= 3 && n < (int64_t)PLAN_MASK, EMK_E_SHAPE, "emk_backbone_amide_atoms: need (b, n >= 3, 3) coordinates");
  int rc = check_indices("emk_backbone_amide_atoms (h_after)", h_after, nh_after, n);
  if (rc) return rc;
  rc = check_indices("emk_backbone_amide_atoms (o_after)", o_after, no_after, n);
  if (rc) return rc;
  std::vector<uint32_t> plan;
  int64_t h, o;
  merge_plan(n, h_after, nh_after, o_after, no_after, true, &plan, &h, &o);
  EMK_REQUIRE((int64_t)plan.size() == n_out, EMK_E_SHAPE, "emk_backbone_amide_atoms: output holds %lld atoms, the index lists make %lld",
              (long long)n_out, (long long)plan.size());
  GenParams p{};
  p.src0 = central; p.n0 = (int)n; p.b = b; p.a = sp2_params(h_angle, h_len); p.bb = sp2_params(o_angle, o_len); p.out = out;
  return run_plan(p, plan, st);
}

// ---- topology-aware back-mapping: set the dihedrals of an all-atom structure one after the other ------------------------------
// The numeric core of mdtraj_backmapping (encodermap/misc/backmapping.py:1661-1690 for the backbone dihedrals, :1722-1745 for
// the side-chain dihedrals; the untested numba twin parallel_rotation_application :384-405): for every frame, for every
// dihedral j IN ORDER: measure the current dihedral of its four atoms (misc/rotate.py:547-581), rotate the atoms on the far
// side of its central bond by (target - current) about the bond axis through the bond's first atom (transformations'
// rotation_matrix, restated in the reference as _rotmat_jit :356-381).  Which atoms are "far" comes from the bond graph
// (networkx in the reference, misc/rotate.py:409-511): index lists handed over by the host.  The reference runs this as a
// Python loop over frames x dihedrals (with a progress bar); here one CTA owns a frame, keeps its coordinates in shared memory
// as float64 (the reference stores float32 and rounds after every rotation), warp 0 builds the rotation of the current
// dihedral, all warps apply it to the far side.  O(frames x sum of far-side sizes); the index lists are read from L2.
struct SetDihedralParams {
  const float* start;        // (1 | frames, n_atoms, 3)
  int64_t start_stride;      // 0 or 3 * n_atoms
  const int* quads;          // (D, 4) device
  const int* bonds;          // (D, 2) device
  const int* far_offsets;    // (D + 1) device
  const int* far_atoms;      // concatenated far sides, device
  const float* targets;      // (frames, D)
  int64_t frames;
  int n_atoms, n_dihedrals;
  float* out;                // (frames, n_atoms, 3)
};

__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}

__global__ void __launch_bounds__(256) set_dihedrals_kernel(const SetDihedralParams p) {
  extern __shared__ double sxyz[];          // [n_atoms][3], then the 12 numbers of the current transform
  double* xf = sxyz;
  double* tr = sxyz + 3 * (size_t)p.n_atoms;   // R (9, row-major) and the pivot (3)
  const int tid = threadIdx.x, nth = blockDim.x;
  for (int64_t f = blockIdx.x; f < p.frames; f += gridDim.x) {
    __syncthreads();
    const float* src = p.start + f * p.start_stride;
    for (int e = tid; e < 3 * p.n_atoms; e += nth) xf[e] = (double)src[e];
    __syncthreads();
    for (int j = 0; j < p.n_dihedrals; j++) {
      if (tid < 32) {
        // every lane of warp 0 computes the same numbers (no divergence, no shuffles); lane 0 publishes them
        const int a = __ldg(p.quads + 4 * j), b = __ldg(p.quads + 4 * j + 1), c = __ldg(p.quads + 4 * j + 2), d = __ldg(p.quads + 4 * j + 3);
        double b1[3], b2[3], b3[3], c1[3], c2[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
          b1[k] = xf[3 * b + k] - xf[3 * a + k];
          b2[k] = xf[3 * c + k] - xf[3 * b + k];
          b3[k] = xf[3 * d + k] - xf[3 * c + k];
        }
        cross3(b2, b3, c1);
        cross3(b1, b2, c2);
        double p1 = b1[0] * c1[0] + b1[1] * c1[1] + b1[2] * c1[2];
        p1 *= sqrt(b2[0] * b2[0] + b2[1] * b2[1] + b2[2] * b2[2]);
        const double p2 = c1[0] * c2[0] + c1[1] * c2[1] + c1[2] * c2[2];
        const double current = atan2(p1, p2);
        const double angle = (double)__ldg(p.targets + f * p.n_dihedrals + j) - current;
        const int u = __ldg(p.bonds + 2 * j), v = __ldg(p.bonds + 2 * j + 1);
        double dir[3] = {xf[3 * v] - xf[3 * u], xf[3 * v + 1] - xf[3 * u + 1], xf[3 * v + 2] - xf[3 * u + 2]};
        const double inv = 1.0 / sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
        dir[0] *= inv; dir[1] *= inv; dir[2] *= inv;
        double sa, ca;
        sincos(angle, &sa, &ca);
        const double oc = 1.0 - ca;
        if (tid == 0) {
          tr[0] = ca + dir[0] * dir[0] * oc;          tr[1] = dir[0] * dir[1] * oc - dir[2] * sa; tr[2] = dir[0] * dir[2] * oc + dir[1] * sa;
          tr[3] = dir[1] * dir[0] * oc + dir[2] * sa; tr[4] = ca + dir[1] * dir[1] * oc;          tr[5] = dir[1] * dir[2] * oc - dir[0] * sa;
          tr[6] = dir[2] * dir[0] * oc - dir[1] * sa; tr[7] = dir[2] * dir[1] * oc + dir[0] * sa; tr[8] = ca + dir[2] * dir[2] * oc;
          tr[9] = xf[3 * u]; tr[10] = xf[3 * u + 1]; tr[11] = xf[3 * u + 2];
        }
      }
      __syncthreads();
      const double r0 = tr[0], r1 = tr[1], r2 = tr[2], r3 = tr[3], r4 = tr[4], r5 = tr[5], r6 = tr[6], r7 = tr[7], r8 = tr[8];
      const double px = tr[9], py = tr[10], pz = tr[11];
      const int e0 = __ldg(p.far_offsets + j), e1 = __ldg(p.far_offsets + j + 1);
      for (int e = e0 + tid; e < e1; e += nth) {
        const int at = __ldg(p.far_atoms + e);
        const double x = xf[3 * at] - px, y = xf[3 * at + 1] - py, z = xf[3 * at + 2] - pz;
        xf[3 * at] = px + r0 * x + r1 * y + r2 * z;
        xf[3 * at + 1] = py + r3 * x + r4 * y + r5 * z;
        xf[3 * at + 2] = pz + r6 * x + r7 * y + r8 * z;
      }
      __syncthreads();
    }
    float* dst = p.out + f * (int64_t)(3 * p.n_atoms);
    for (int e = tid; e < 3 * p.n_atoms; e += nth) dst[e] = (float)xf[e];
  }
}

int set_dihedrals_device(const float* start, int64_t start_frames, int64_t n_atoms, const int32_t* quads, const int32_t* bonds,
                         const int32_t* far_offsets, const int32_t* far_atoms, int64_t n_dihedrals, const float* targets, int64_t frames,
                         float* out, cudaStream_t st) {
  EMK_REQUIRE(frames == 0 || (start && out && (n_dihedrals == 0 || (quads && bonds && far_offsets && targets))), EMK_E_NULL,
              "emk_set_dihedrals: NULL pointer argument");
  EMK_REQUIRE(frames >= 0 && n_atoms >= 1 && n_dihedrals >= 0 && (start_frames == 1 || start_frames == frames), EMK_E_SHAPE,
              "emk_set_dihedrals: need start (1 | frames, n_atoms, 3), targets (frames, n_dihedrals)");
  const size_t smem = (3 * (size_t)n_atoms + 12) * sizeof(double);
  EMK_REQUIRE(smem <= 227 * 1024, EMK_E_UNSUPPORTED, "emk_set_dihedrals: %lld atoms need %zu bytes of shared memory per frame (limit 227 KB)",
              (long long)n_atoms, smem);
  if (frames == 0) return EMK_OK;
  // validate the index lists on the host (they ARE host arrays): a bad index would otherwise corrupt shared memory
  int64_t total_far = 0;
  for (int64_t j = 0; j < n_dihedrals; j++) {
    for (int k = 0; k < 4; k++)
      EMK_REQUIRE(quads[4 * j + k] >= 0 && quads[4 * j + k] < n_atoms, EMK_E_ARG, "emk_set_dihedrals: dihedral %lld names atom %d of %lld",
                  (long long)j, quads[4 * j + k], (long long)n_atoms);
    for (int k = 0; k < 2; k++)
      EMK_REQUIRE(bonds[2 * j + k] >= 0 && bonds[2 * j + k] < n_atoms, EMK_E_ARG, "emk_set_dihedrals: bond %lld names atom %d of %lld",
                  (long long)j, bonds[2 * j + k], (long long)n_atoms);
    EMK_REQUIRE(far_offsets[j] <= far_offsets[j + 1] && far_offsets[j] >= 0, EMK_E_ARG, "emk_set_dihedrals: far_offsets must be non-decreasing");
  }
  if (n_dihedrals > 0) {
    EMK_REQUIRE(far_offsets[0] == 0, EMK_E_ARG, "emk_set_dihedrals: far_offsets[0] must be 0");
    total_far = far_offsets[n_dihedrals];
    EMK_REQUIRE(total_far == 0 || far_atoms, EMK_E_NULL, "emk_set_dihedrals: NULL far_atoms");
    for (int64_t e = 0; e < total_far; e++)
      EMK_REQUIRE(far_atoms[e] >= 0 && far_atoms[e] < n_atoms, EMK_E_ARG, "emk_set_dihedrals: far side names atom %d of %lld", far_atoms[e],
                  (long long)n_atoms);
  }
  // one device buffer for all four index arrays
  const size_t n_ints = (size_t)(6 * n_dihedrals + n_dihedrals + 1 + total_far);
  int* dbuf = nullptr;
  int rc = scratch_alloc(reinterpret_cast<void**>(&dbuf), std::max<size_t>(n_ints, 1) * sizeof(int), st);
  if (rc) return rc;
  std::vector<int> host(n_ints);
  int* hq = host.data();
  int* hb = hq + 4 * n_dihedrals;
  int* ho = hb + 2 * n_dihedrals;
  int* hf = ho + n_dihedrals + 1;
  for (int64_t k = 0; k < 4 * n_dihedrals; k++) hq[k] = quads[k];
  for (int64_t k = 0; k < 2 * n_dihedrals; k++) hb[k] = bonds[k];
  for (int64_t k = 0; k <= n_dihedrals; k++) ho[k] = n_dihedrals > 0 ? far_offsets[k] : 0;
  for (int64_t k = 0; k < total_far; k++) hf[k] = far_atoms[k];
  cudaError_t e = cudaMemcpyAsync(dbuf, host.data(), n_ints * sizeof(int), cudaMemcpyHostToDevice, st);   // pageable: staged before return
  if (e != cudaSuccess) {
    cudaFreeAsync(dbuf, st);
    return fail((int)e, "emk_set_dihedrals: index upload failed: %s", cudaGetErrorString(e));
  }
  SetDihedralParams p{};
  p.start = start; p.start_stride = start_frames == 1 ? 0 : 3 * n_atoms;
  p.quads = dbuf; p.bonds = dbuf + 4 * n_dihedrals; p.far_offsets = dbuf + 6 * n_dihedrals; p.far_atoms = dbuf + 7 * n_dihedrals + 1;
  p.targets = targets; p.frames = frames; p.n_atoms = (int)n_atoms; p.n_dihedrals = (int)n_dihedrals; p.out = out;
  static bool cfg[kMaxDevices] = {false};
  if (first_use_on_device(cfg)) EMK_CUDA(cudaFuncSetAttribute(set_dihedrals_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (227 * 1024) / (smem + 1024)));
  const unsigned grid = (unsigned)std::min<int64_t>(frames, (int64_t)sm_count() * per_sm);
  set_dihedrals_kernel<<<grid, 256, smem, st>>>(p);
  rc = launch_status("set_dihedrals_kernel");
  cudaFreeAsync(dbuf, st);
  return rc;
}

// number of atoms the merge produces (host only)
int64_t merged_atom_count(int64_t n, const int64_t* h_after, int64_t nh_after, const int64_t* o_after, int64_t no_after) {
  if (n < 1) return 0;
  for (int64_t k = 0; k < nh_after; k++) if (h_after[k] < 0 || h_after[k] >= n) return -1;
  for (int64_t k = 0; k < no_after; k++) if (o_after[k] < 0 || o_after[k] >= n) return -1;
  std::vector<uint32_t> plan;
  int64_t h, o;
  merge_plan(n, h_after, nh_after, o_after, no_after, true, &plan, &h, &o);
  return (int64_t)plan.size();
}

}  // namespace emk
