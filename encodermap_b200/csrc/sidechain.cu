// Back-mapping WITH side chains: BackMapLayerWithSidechains, forward and exact backward, plus the atom gather
// PairwiseDistances uses on its output.
//
// Reference: encodermap/models/layers.py:218-843 (numpy twin encodermap/misc/backmapping.py:424-1002).  The layer lays all atoms
// out in the z = 0 plane -- the 3 n backbone atoms on the x axis at the running sum of the bond lengths, every side chain
// straight up in y from its CA -- and then sets every bond angle and every dihedral ONE AFTER THE OTHER: measure the current
// value on the atoms as they stand, rotate the atoms "behind" the pivot / bond by the difference.  Which atoms are behind comes
// from boolean masks (n_ops x n_atoms) built in the constructor.  In TensorFlow every one of those ~6 n + 2 S steps is a dozen
// kernels that gather, rotate and re-stack the whole (batch, n_atoms, 4) tensor.
//
// Here: the masks of this construction are at most two index ranges per step (a suffix of the backbone and a suffix of the
// side-chain atoms, or a suffix of one side chain), so the plan is 12 ints per step, built once on the host
// (emk_sidechain_plan_create) and kept on the device.  One CTA owns a frame: coordinates live in shared memory as float64 for
// the whole sequence (a bond angle is measured on a still straight triplet at every step but the first of a side chain, where
// the reference's acos turns the rounding of the cosine into sqrt(2 eps): 3.5e-4 rad in float32, 1.5e-8 in float64), warp 0
// measures and builds the rotation -- from sines and cosines only, no inverse trigonometric function -- and all warps apply it
// to the moving ranges.  The backward kernel re-runs the forward keeping sin / cos of every rotation angle, then walks the steps
// in reverse: every rotation is undone in place (the pivot and the axis atoms do not move, so the
// inverse is known from the state after the step), which restores the coordinates each measurement saw without storing them,
// while the gradient is pulled back through the rotation, its angle, its axis, its pivot and the measured value.
#include <algorithm>
#include <vector>

#include "emk_common.cuh"

namespace emk {

constexpr int kOpInts = 12;   // kind, a, b, c, d, column, lo0, hi0, lo1, hi1, 0, 0   (three int4 loads)
enum OpKind { kCentralAngle = 0, kSideAngle = 1, kCentralDihedral = 2, kSideDihedral = 3 };
// a measured bond angle with 1 - cos^2 below this is a constant for the backward pass (acos is not differentiable at a straight
// triplet; TensorFlow gives 0, a huge number or NaN there depending on how the float32 cosine rounded)
constexpr double kStraightEps = 1e-12;

struct SidechainPlan {
  int n_res = 0, n_bb = 0, n_side = 0, n_atoms = 0, n_ops = 0;
  int cols[6] = {0, 0, 0, 0, 0, 0};   // central distances / angles / dihedrals, side distances / angles / dihedrals
  std::vector<int> ops;               // n_ops x kOpInts
  std::vector<int> side;              // n_side x 4: CA atom, first atom of its chain, one past its last atom, 0
  std::vector<int> res;               // residues with a side chain x 4: N atom, first side-chain atom, dihedral count, first dihedral column
  int n_ca = 0, n_cd = 0;             // backbone bond-angle / dihedral steps
  int device = -1;
  int* d_mem = nullptr;               // ops, then side
};

// ---- host: the index construction of the layer, as ranges --------------------------------------------------------------------
// layers.py:249-474.  Mask row r of the backbone bonds keeps backbone atoms <= r and the first T[r] side-chain atoms; T is built
// exactly as the reference builds its "right side" rows (:254-359): one row of zeros, three rows per residue holding the number
// of side-chain atoms placed so far (a residue without side chain repeats the previous residue's rows; the first and the last
// residue contribute nothing when they have none), one row of everything.  That only adds up to the 3 n - 1 bonds when exactly
// one of the two end residues has no side chain: any other description fails in the reference (np.hstack) and is refused here.
static int build_plan(int64_t n_res, const int32_t* counts, SidechainPlan* pl) {
  EMK_REQUIRE(n_res >= 1 && n_res < (1 << 20) && counts, EMK_E_ARG, "emk_sidechain_plan_create: need the side-chain dihedral count of >= 1 residues");
  int64_t n_side = 0, n_sdih = 0;
  for (int64_t r = 0; r < n_res; r++) {
    EMK_REQUIRE(counts[r] >= 0 && counts[r] < 64, EMK_E_ARG, "emk_sidechain_plan_create: residue %lld has %d side-chain dihedrals", (long long)(r + 1), counts[r]);
    if (counts[r] > 0) n_side += counts[r] + 1;
    n_sdih += counts[r];
  }
  EMK_REQUIRE(n_side > 0, EMK_E_UNSUPPORTED, "emk_sidechain_plan_create: no residue has a side chain (the reference layer cannot be built either: layers.py:477)");
  const int n_bb = (int)(3 * n_res);
  std::vector<int> kept;   // T
  kept.push_back(0);
  int filled = 0, last = -1;
  for (int64_t r = 0; r < n_res; r++) {
    if (counts[r] == 0) {
      if (r == 0 || r == n_res - 1) continue;
      EMK_REQUIRE(last >= 0, EMK_E_UNSUPPORTED,
                  "emk_sidechain_plan_create: residue %lld has no side chain and no residue before it has one (NameError in the reference, layers.py:287)",
                  (long long)(r + 1));
    } else {
      filled += counts[r] + 1;
      last = filled;
    }
    for (int k = 0; k < 3; k++) kept.push_back(last);
  }
  kept.push_back((int)n_side);
  EMK_REQUIRE((int)kept.size() == n_bb - 1, EMK_E_UNSUPPORTED,
              "emk_sidechain_plan_create: exactly one of the first / last residue must be without side chain (the reference's index "
              "construction yields %d rows for %d backbone bonds, layers.py:370-372)", (int)kept.size(), n_bb - 1);
  pl->n_res = (int)n_res; pl->n_bb = n_bb; pl->n_side = (int)n_side; pl->n_atoms = n_bb + (int)n_side;
  pl->cols[0] = n_bb - 1; pl->cols[1] = n_bb - 2; pl->cols[2] = n_bb - 3; pl->cols[3] = (int)n_side; pl->cols[4] = (int)n_side; pl->cols[5] = (int)n_sdih;
  auto& ops = pl->ops;
  auto push = [&](int kind, int a, int b, int c, int d, int col, int lo0, int hi0, int lo1, int hi1) {
    const int rec[kOpInts] = {kind, a, b, c, d, col, lo0, hi0, lo1, hi1, 0, 0};
    ops.insert(ops.end(), rec, rec + kOpInts);
  };
  const int n_atoms = pl->n_atoms;
  // backbone bond angles: mask row i + 1, about +z through atom i + 1 (:654-717)
  for (int i = 0; i + 2 < n_bb; i++) push(kCentralAngle, i, i + 1, i + 2, -1, i, i + 2, n_bb, n_bb + kept[i + 1], n_atoms);
  // side-chain bond angles, about -z (:720-783); the chain of a residue is N, CA, then its own atoms
  pl->side.assign((size_t)n_side * 4, 0);
  {
    int first = n_bb, j = 0;
    for (int64_t r = 0; r < n_res; r++) {
      const int c = counts[r];
      if (c == 0) continue;
      const int end = first + c + 1;
      auto chain = [&](int q) { return q == 0 ? (int)(3 * r) : q == 1 ? (int)(3 * r + 1) : first + q - 2; };
      for (int q = 0; q <= c; q++, j++) {
        push(kSideAngle, chain(q), chain(q + 1), chain(q + 2), -1, j, first + q, end, 0, 0);
        pl->side[4 * (size_t)j] = (int)(3 * r + 1); pl->side[4 * (size_t)j + 1] = first; pl->side[4 * (size_t)j + 2] = end;
      }
      first = end;
    }
  }
  // backbone dihedrals: mask row i + 1, about the bond (i + 1, i + 2) (:786-841)
  for (int i = 0; i + 3 < n_bb; i++) push(kCentralDihedral, i, i + 1, i + 2, i + 3, i, i + 2, n_bb, n_bb + kept[i + 1], n_atoms);
  // side-chain dihedrals: the rows of the side-chain mask without the last one of every residue (:423-428)
  {
    int first = n_bb, col = 0;
    for (int64_t r = 0; r < n_res; r++) {
      const int c = counts[r];
      if (c == 0) continue;
      const int end = first + c + 1;
      auto chain = [&](int q) { return q == 0 ? (int)(3 * r) : q == 1 ? (int)(3 * r + 1) : first + q - 2; };
      const int rec[4] = {(int)(3 * r), first, c, col};
      pl->res.insert(pl->res.end(), rec, rec + 4);
      for (int q = 0; q < c; q++, col++) push(kSideDihedral, chain(q), chain(q + 1), chain(q + 2), chain(q + 3), col, first + q, end, 0, 0);
      first = end;
    }
  }
  pl->n_ca = n_bb - 2; pl->n_cd = n_bb - 3;
  pl->n_ops = (int)(ops.size() / kOpInts);
  return EMK_OK;
}

// ---- device --------------------------------------------------------------------------------------------------------------------
struct ScParams {
  const int4* ops;
  const int4* side;
  const int4* res;
  int n_res_side, n_ca, n_cd;
  const float* in[6];
  int cols[6];
  int64_t frames;
  int n_bb, n_side, n_atoms, n_ops;
  float* out;
  const float* gout;
  float* gin[6];
  double* save;           // forward: per frame (sin, cos) of every rotation, the angle flags and the float64 coordinates, or nullptr
  const double* saved;    // backward: the same block, or nullptr (the forward pass is then repeated inside the kernel)
  int64_t save_stride;    // doubles per frame
};

// slots of the published transform
enum { TR_R = 0, TR_P = 9, TR_U = 12, TR_S = 15, TR_C = 16, TR_N = 18 };

__device__ __forceinline__ void cross3d(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3d(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

__device__ __forceinline__ int input_of(int kind) { return kind == kCentralAngle ? 1 : kind == kSideAngle ? 4 : kind == kCentralDihedral ? 2 : 5; }

// No inverse trigonometric function is evaluated anywhere: a step needs sin and cos of (target - measured), and
//   cos(t - m) = cos t cos m + sin t sin m,   sin(t - m) = sin t cos m - cos t sin m
// with sin / cos of the TARGETS tabulated per frame by all threads before the sequence starts (they do not depend on the
// coordinates) and sin / cos of the MEASURED value read off the vectors: for a bond angle cos m = ba.bc / (|ba||bc|),
// sin m = |ba x bc| / (|ba||bc|) >= 0; for a dihedral (cos m, sin m) = (p2, p1) / hypot(p1, p2)  (layers.py:674-683, 800-809).
// The cross product also gives sin m of a (nearly) straight triplet without the cancellation of acos(-1 + eps).
struct Measured {
  double s, c;      // sin, cos of (target - measured)
  double u[3];      // unit rotation axis
  double sm;        // bond angles: sin of the measured angle
};

__device__ __forceinline__ Measured sc_measure(const double* xf, int kind, int a, int b, int c, int d, double sin_t, double cos_t) {
  Measured m;
  if (kind <= kSideAngle) {
    double ba[3], bc[3], cr[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { ba[k] = xf[3 * a + k] - xf[3 * b + k]; bc[k] = xf[3 * c + k] - xf[3 * b + k]; }
    const double inv = rsqrt(dot3d(ba, ba) * dot3d(bc, bc));
    const double cm = fmin(fmax(dot3d(ba, bc) * inv, -1.0), 1.0);
    cross3d(ba, bc, cr);
    m.sm = fmin(sqrt(dot3d(cr, cr)) * inv, 1.0);
    m.c = cos_t * cm + sin_t * m.sm;
    m.s = sin_t * cm - cos_t * m.sm;
    m.u[0] = 0.0; m.u[1] = 0.0; m.u[2] = kind == kCentralAngle ? 1.0 : -1.0;               // self.up / self.down
  } else {
    double b1[3], b2[3], b3[3], c1[3], c2[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { b1[k] = xf[3 * b + k] - xf[3 * a + k]; b2[k] = xf[3 * c + k] - xf[3 * b + k]; b3[k] = xf[3 * d + k] - xf[3 * c + k]; }
    const double l2 = dot3d(b2, b2), il = rsqrt(l2);
    cross3d(b2, b3, c1);
    cross3d(b1, b2, c2);
    const double p1 = dot3d(b1, c1) * (l2 * il), p2 = dot3d(c1, c2);
    const double h2 = p1 * p1 + p2 * p2;
    double cm = 1.0, sm = 0.0;                                                             // atan2(0, 0) = 0
    if (h2 > 0.0) { const double ih = rsqrt(h2); cm = p2 * ih; sm = p1 * ih; }
    m.sm = sm;
    m.c = cos_t * cm + sin_t * sm;
    m.s = sin_t * cm - cos_t * sm;
    m.u[0] = b2[0] * il; m.u[1] = b2[1] * il; m.u[2] = b2[2] * il;
  }
  return m;
}

// rotation by (s, c) about the unit axis u into tr[TR_R..]
__device__ __forceinline__ void publish_rotation(double* tr, const double* u, double s, double c) {
  const double oc = 1.0 - c;
  tr[0] = c + u[0] * u[0] * oc;        tr[1] = u[0] * u[1] * oc - u[2] * s; tr[2] = u[0] * u[2] * oc + u[1] * s;
  tr[3] = u[1] * u[0] * oc + u[2] * s; tr[4] = c + u[1] * u[1] * oc;        tr[5] = u[1] * u[2] * oc - u[0] * s;
  tr[6] = u[2] * u[0] * oc - u[1] * s; tr[7] = u[2] * u[1] * oc + u[0] * s; tr[8] = c + u[2] * u[2] * oc;
  tr[TR_U] = u[0]; tr[TR_U + 1] = u[1]; tr[TR_U + 2] = u[2]; tr[TR_S] = s; tr[TR_C] = c;
}

// initial planar layout (layers.py:555-648) and the per-frame table of sin / cos of the targets; flags[k]: 0 the target of a
// bond angle lies in [0, pi] (the sign of target - measured is the sign of its sine), 1 above, 2 below
__device__ __forceinline__ void sc_layout(const ScParams& p, int64_t f, double* xf, double* tg, unsigned char* flags) {
  const int tid = threadIdx.x, nth = blockDim.x;
  const float* cd = p.in[0] + f * p.cols[0];
  const float* sd = p.in[3] + f * p.cols[3];
  for (int k = tid; k < p.n_bb; k += nth) {
    xf[3 * k] = k == 0 ? 0.0 : (double)__ldg(cd + k - 1);
    xf[3 * k + 1] = 0.0; xf[3 * k + 2] = 0.0;
  }
  for (int k = tid; k < p.n_ops; k += nth) {
    const int4 o0 = __ldg(p.ops + 3 * k), o1 = __ldg(p.ops + 3 * k + 1);
    const int src = input_of(o0.x);
    const double t = (double)__ldg(p.in[src] + f * p.cols[src] + o1.y);
    sincos(t, tg + 2 * k, tg + 2 * k + 1);
    if (flags) flags[k] = t > 3.141592653589793 ? 1 : t < 0.0 ? 2 : 0;
  }
  __syncthreads();
  if (tid < 32) {       // running sum of the bond lengths: warp scan over 32 atoms at a time
    double carry = 0.0;
    for (int base = 0; base < p.n_bb; base += 32) {
      const int k = base + tid;
      double v = k < p.n_bb ? xf[3 * k] : 0.0;
#pragma unroll
      for (int m = 1; m < 32; m <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, v, m);
        if (tid >= m) v += up;
      }
      v += carry;
      if (k < p.n_bb) xf[3 * k] = v;
      carry = __shfl_sync(0xffffffffu, v, 31);
    }
  }
  __syncthreads();
  for (int m = tid; m < p.n_side; m += nth) {
    const int4 s = __ldg(p.side + m);
    double y = 0.0;
    for (int q = s.y; q <= p.n_bb + m; q++) y += (double)__ldg(sd + q - p.n_bb);
    const int at = p.n_bb + m;
    xf[3 * at] = xf[3 * s.x]; xf[3 * at + 1] = y; xf[3 * at + 2] = 0.0;
  }
  __syncthreads();
}

// x -> p + R (x - p) for one atom
__device__ __forceinline__ void sc_rotate_atom(double* xf, const double* tr, int at) {
  const double x = xf[3 * at] - tr[TR_P], y = xf[3 * at + 1] - tr[TR_P + 1], z = xf[3 * at + 2] - tr[TR_P + 2];
  xf[3 * at] = tr[TR_P] + tr[0] * x + tr[1] * y + tr[2] * z;
  xf[3 * at + 1] = tr[TR_P + 1] + tr[3] * x + tr[4] * y + tr[5] * z;
  xf[3 * at + 2] = tr[TR_P + 2] + tr[6] * x + tr[7] * y + tr[8] * z;
}

// The steps of the BACKBONE are strictly sequential (each measures atoms the previous one moved).  The steps of a SIDE CHAIN only
// measure N, CA and the chain's own atoms and only move the chain's own atoms, and no backbone step runs between them: the side
// chains of different residues are independent, so each gets one thread that walks its chain alone -- no barrier, no hand-over.
// Sequential depth per frame: 6 n - 5 + (longest side chain) instead of 6 n - 5 + 2 (side-chain atoms).

// backbone steps [k0, k1) by the whole CTA.  KEEP: leave (sin(target - measured), cos) of every step in tg for the backward pass
template <bool KEEP>
__device__ __forceinline__ void sc_backbone_steps(const ScParams& p, int k0, int k1, double* xf, double* tg, double* tr) {
  const int tid = threadIdx.x, nth = blockDim.x;
  int4 o0 = __ldg(p.ops + 3 * k0), o1 = __ldg(p.ops + 3 * k0 + 1), o2 = __ldg(p.ops + 3 * k0 + 2);
  for (int k = k0; k < k1; k++) {
    int4 n0 = o0, n1 = o1, n2 = o2;
    if (k + 1 < k1) { n0 = __ldg(p.ops + 3 * k + 3); n1 = __ldg(p.ops + 3 * k + 4); n2 = __ldg(p.ops + 3 * k + 5); }
    if (tid == 0) {
      // one thread measures and publishes the rotation (a serial chain of ~300 float64 instructions: nothing to share out)
      const Measured m = sc_measure(xf, o0.x, o0.y, o0.z, o0.w, o1.x, tg[2 * k], tg[2 * k + 1]);
      publish_rotation(tr, m.u, o0.x <= kSideAngle ? fabs(m.s) : m.s, m.c);              // bond angles rotate by |target - measured|
      tr[TR_P] = xf[3 * o0.z]; tr[TR_P + 1] = xf[3 * o0.z + 1]; tr[TR_P + 2] = xf[3 * o0.z + 2];
      if (KEEP) { tg[2 * k] = m.s; tg[2 * k + 1] = m.c; }
    }
    __syncthreads();
    const double r0 = tr[0], r1 = tr[1], r2 = tr[2], r3 = tr[3], r4 = tr[4], r5 = tr[5], r6 = tr[6], r7 = tr[7], r8 = tr[8];
    const double px = tr[TR_P], py = tr[TR_P + 1], pz = tr[TR_P + 2];
    const int c0 = o1.w - o1.z, total = c0 + (o2.y - o2.x);
    for (int e = tid; e < total; e += nth) {
      const int at = e < c0 ? o1.z + e : o2.x + (e - c0);
      const double x = xf[3 * at] - px, y = xf[3 * at + 1] - py, z = xf[3 * at + 2] - pz;
      xf[3 * at] = px + r0 * x + r1 * y + r2 * z;
      xf[3 * at + 1] = py + r3 * x + r4 * y + r5 * z;
      xf[3 * at + 2] = pz + r6 * x + r7 * y + r8 * z;
    }
    __syncthreads();
    o0 = n0; o1 = n1; o2 = n2;
  }
}

// atom q of the chain N, CA, CB, CG, ... of a residue
__device__ __forceinline__ int sc_chain(const int4& r, int q) { return q < 2 ? r.x + q : r.y + q - 2; }

// side-chain steps of one kind (bond angles: DIH = false, k_base = first side-angle step; dihedrals: DIH = true), one thread per
// residue with a side chain
template <bool DIH, bool KEEP>
__device__ __forceinline__ void sc_side_steps(const ScParams& p, int k_base, double* xf, double* tg) {
  for (int ri = threadIdx.x; ri < p.n_res_side; ri += blockDim.x) {
    const int4 r = __ldg(p.res + ri);                       // N atom, first side-chain atom, dihedral count, first dihedral column
    const int end = r.y + r.z + 1;
    const int k0 = k_base + (DIH ? r.w : r.y - p.n_bb);
    const int steps = DIH ? r.z : r.z + 1;
    for (int q = 0; q < steps; q++) {
      const int k = k0 + q;
      const int b = sc_chain(r, q + 1);
      const Measured m = sc_measure(xf, DIH ? kSideDihedral : kSideAngle, sc_chain(r, q), b, sc_chain(r, q + 2), DIH ? sc_chain(r, q + 3) : -1,
                                    tg[2 * k], tg[2 * k + 1]);
      double tr[TR_N];
      publish_rotation(tr, m.u, DIH ? m.s : fabs(m.s), m.c);
      tr[TR_P] = xf[3 * b]; tr[TR_P + 1] = xf[3 * b + 1]; tr[TR_P + 2] = xf[3 * b + 2];
      if (KEEP) { tg[2 * k] = m.s; tg[2 * k + 1] = m.c; }
      for (int at = r.y + q; at < end; at++) sc_rotate_atom(xf, tr, at);
    }
  }
  __syncthreads();
}

template <bool KEEP>
__device__ __forceinline__ void sc_forward(const ScParams& p, double* xf, double* tg, double* tr) {
  sc_backbone_steps<KEEP>(p, 0, p.n_ca, xf, tg, tr);                                   // layers.py:654-717
  sc_side_steps<false, KEEP>(p, p.n_ca, xf, tg);                                       // :720-783
  const int k_cd = p.n_ca + p.n_side;
  sc_backbone_steps<KEEP>(p, k_cd, k_cd + p.n_cd, xf, tg, tr);                         // :786-841, backbone rows
  sc_side_steps<true, KEEP>(p, k_cd + p.n_cd, xf, tg);                                 // :786-841, side-chain rows
}

// doubles per frame of the state the backward pass can take over from the forward pass
__host__ __device__ inline int64_t sc_saved_doubles(int n_atoms, int n_ops) { return 2 * (int64_t)n_ops + (n_ops + 7) / 8 + 3 * (int64_t)n_atoms; }

template <bool SAVE>
__global__ void __launch_bounds__(128, 7) sidechain_fwd_kernel(const ScParams p) {
  extern __shared__ double sc_smem[];
  double* xf = sc_smem;
  double* tg = xf + 3 * (size_t)p.n_atoms;
  double* tr = tg + 2 * (size_t)p.n_ops;
  unsigned char* flags = SAVE ? reinterpret_cast<unsigned char*>(tr + TR_N) : nullptr;
  for (int64_t f = blockIdx.x; f < p.frames; f += gridDim.x) {
    __syncthreads();
    sc_layout(p, f, xf, tg, flags);
    sc_forward<SAVE>(p, xf, tg, tr);
    float* dst = p.out + f * (int64_t)(3 * p.n_atoms);
    for (int e = threadIdx.x; e < 3 * p.n_atoms; e += blockDim.x) dst[e] = (float)xf[e];
    if (SAVE) {
      double* sv = p.save + f * p.save_stride;
      for (int e = threadIdx.x; e < 2 * p.n_ops; e += blockDim.x) sv[e] = tg[e];
      unsigned char* sf = reinterpret_cast<unsigned char*>(sv + 2 * (size_t)p.n_ops);
      for (int e = threadIdx.x; e < p.n_ops; e += blockDim.x) sf[e] = flags[e];
      double* sx = sv + 2 * (size_t)p.n_ops + (p.n_ops + 7) / 8;
      for (int e = threadIdx.x; e < 3 * p.n_atoms; e += blockDim.x) sx[e] = xf[e];
    }
  }
}

// ---- backward --------------------------------------------------------------------------------------------------------------------
// For the step X' = where(static, X, p + R(theta, u)(X - p)) with theta = target - m(X) (dihedrals) or |target - m(X)| (bond
// angles), u = e_z / -e_z or the normalised bond X_c - X_b, p = X_b, and incoming gradient G' = dL/dX':
//   moving atoms:  G_j = R^T G'_j;   g_p += G'_j - R^T G'_j;   g_theta += G'_j . (u x (X'_j - p));
//                  g_u += sin(theta) (v_j x G'_j) + (1 - cos(theta)) ((u . v_j) G'_j + (u . G'_j) v_j),  v_j = X_j - p
//   then G_b += g_p;  through the normalisation g_d = (g_u - u (u . g_u)) / |d| to atoms c (+) and b (-);
//   dL/dtarget = +-g_theta and -+g_theta dm/dX to the three / four measured atoms.
constexpr int kRed = 7;
constexpr int kScThreads = 128;     // threads per frame (>= 64: thread 32 prepares the next rotation of the backward pass)

// the rotation of step k seen from the state AFTER it (pivot and axis atoms are fixed points of the rotation)
__device__ __forceinline__ void sc_publish_inverse(const double* xf, const double* tg, double* tr, int k, int kind, int b, int c) {
  double u[3];
  if (kind <= kSideAngle) {
    u[0] = 0.0; u[1] = 0.0; u[2] = kind == kCentralAngle ? 1.0 : -1.0;
  } else {
#pragma unroll
    for (int q = 0; q < 3; q++) u[q] = xf[3 * c + q] - xf[3 * b + q];
    const double inv = rsqrt(dot3d(u, u));
    u[0] *= inv; u[1] *= inv; u[2] *= inv;
  }
  publish_rotation(tr, u, kind <= kSideAngle ? fabs(tg[2 * k]) : tg[2 * k], tg[2 * k + 1]);
  tr[TR_P] = xf[3 * b]; tr[TR_P + 1] = xf[3 * b + 1]; tr[TR_P + 2] = xf[3 * b + 2];
}

// the published transform of a step, in registers
struct Rot {
  double r[9], p[3], u[3], s, oc;
};
__device__ __forceinline__ Rot sc_load_rot(const double* tr) {
  Rot t;
#pragma unroll
  for (int q = 0; q < 9; q++) t.r[q] = tr[q];
#pragma unroll
  for (int q = 0; q < 3; q++) { t.p[q] = tr[TR_P + q]; t.u[q] = tr[TR_U + q]; }
  t.s = tr[TR_S]; t.oc = 1.0 - tr[TR_C];
  return t;
}

// one moving atom of a step, backward: restores its position before the step, pulls its gradient back, adds its share of
// (g_theta, g_p, g_u) to acc
__device__ __forceinline__ void sc_atom_bwd(double* xf, double* gf, const Rot& t, int at, bool dihedral, double* acc) {
  const double* r = t.r;
  const double w[3] = {xf[3 * at] - t.p[0], xf[3 * at + 1] - t.p[1], xf[3 * at + 2] - t.p[2]};
  const double v[3] = {r[0] * w[0] + r[3] * w[1] + r[6] * w[2], r[1] * w[0] + r[4] * w[1] + r[7] * w[2], r[2] * w[0] + r[5] * w[1] + r[8] * w[2]};
  xf[3 * at] = t.p[0] + v[0]; xf[3 * at + 1] = t.p[1] + v[1]; xf[3 * at + 2] = t.p[2] + v[2];
  const double g[3] = {gf[3 * at], gf[3 * at + 1], gf[3 * at + 2]};
  const double rg[3] = {r[0] * g[0] + r[3] * g[1] + r[6] * g[2], r[1] * g[0] + r[4] * g[1] + r[7] * g[2], r[2] * g[0] + r[5] * g[1] + r[8] * g[2]};
  double uxw[3];
  cross3d(t.u, w, uxw);
  acc[0] += dot3d(g, uxw);
  acc[1] += g[0] - rg[0]; acc[2] += g[1] - rg[1]; acc[3] += g[2] - rg[2];
  if (dihedral) {
    double vxg[3];
    cross3d(v, g, vxg);
    const double uv = dot3d(t.u, v), ug = dot3d(t.u, g);
#pragma unroll
    for (int q = 0; q < 3; q++) acc[4 + q] += t.s * vxg[q] + t.oc * (uv * g[q] + ug * v[q]);
  }
  gf[3 * at] = rg[0]; gf[3 * at + 1] = rg[1]; gf[3 * at + 2] = rg[2];
}

// the same for a bond-angle step of the backbone, whose axis is +-e_z: the rotation is planar (x, y mix; z, and the z component
// of the gradient, pass through), and (u x w) = u_z (-w_y, w_x, 0)
__device__ __forceinline__ void sc_atom_bwd_planar(double* xf, double* gf, const Rot& t, int at, double* acc) {
  const double c = t.r[0], sn = t.r[3];                 // R = [[c, -sn, 0], [sn, c, 0], [0, 0, 1]]
  const double wx = xf[3 * at] - t.p[0], wy = xf[3 * at + 1] - t.p[1];
  xf[3 * at] = t.p[0] + (c * wx + sn * wy);
  xf[3 * at + 1] = t.p[1] + (c * wy - sn * wx);
  const double gx = gf[3 * at], gy = gf[3 * at + 1];
  const double rgx = c * gx + sn * gy, rgy = c * gy - sn * gx;
  acc[0] += t.u[2] * (gy * wx - gx * wy);
  acc[1] += gx - rgx; acc[2] += gy - rgy;
  gf[3 * at] = rgx; gf[3 * at + 1] = rgy;
}

// what the sums of a step do to the pivot, the axis atoms and the measured atoms (xf holds the state BEFORE the step); returns
// dL/dtarget.  sd = sin(target - measured) kept by the forward pass, flag: see sc_layout
__device__ __forceinline__ double sc_finish_step(const double* xf, double* gf, int kind, int a, int b, int c, int d, const double* sum, double sd,
                                                 int flag) {
  // contributions are collected per atom in registers and added to shared memory once at the end (a, b, c, d are four different
  // atoms): read-modify-write chains through shared memory on the same address would serialise
  double ga[3] = {0.0, 0.0, 0.0}, gb[3] = {sum[1], sum[2], sum[3]}, gc[3] = {0.0, 0.0, 0.0}, gd[3] = {0.0, 0.0, 0.0};
  const double g_theta = sum[0];
  double g_target;
  double xa[3], xb[3], xc[3];
#pragma unroll
  for (int q = 0; q < 3; q++) { xa[q] = xf[3 * a + q]; xb[q] = xf[3 * b + q]; xc[q] = xf[3 * c + q]; }
  if (kind <= kSideAngle) {
    double ba[3], bc[3], cr[3];
#pragma unroll
    for (int q = 0; q < 3; q++) { ba[q] = xa[q] - xb[q]; bc[q] = xc[q] - xb[q]; }
    const double na2 = dot3d(ba, ba), nc2 = dot3d(bc, bc), inv = rsqrt(na2 * nc2);
    const double t = dot3d(ba, bc) * inv;
    cross3d(ba, bc, cr);
    const double sm2 = dot3d(cr, cr) * inv * inv;                                   // sin^2 of the measured angle
    const double sgn = flag == 1 ? 1.0 : flag == 2 ? -1.0 : sd > 0.0 ? 1.0 : sd < 0.0 ? -1.0 : 0.0;     // d|x|/dx
    g_target = sgn * g_theta;
    if (t >= -1.0 && t <= 1.0 && sm2 >= kStraightEps) {
      const double gt = g_target * rsqrt(sm2);           // dL/dcos = (-sgn g_theta) (-1 / sin(measured))
      const double ta = t / na2, tc = t / nc2;
#pragma unroll
      for (int q = 0; q < 3; q++) {
        ga[q] = gt * (bc[q] * inv - ta * ba[q]);
        gc[q] = gt * (ba[q] * inv - tc * bc[q]);
        gb[q] -= ga[q] + gc[q];
      }
    }
  } else {
    g_target = g_theta;
    // the axis: u = dvec / |dvec|
    double dv[3], xd[3];
    const double gu[3] = {sum[4], sum[5], sum[6]};
#pragma unroll
    for (int q = 0; q < 3; q++) { dv[q] = xc[q] - xb[q]; xd[q] = xf[3 * d + q]; }
    const double l2 = dot3d(dv, dv), il = rsqrt(l2), len = l2 * il;
    const double ugu = dot3d(dv, gu) * il * il;          // (u . g_u) / |d|
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const double gdir = gu[q] * il - dv[q] * ugu * il;      // (g_u - u (u . g_u)) / |d|
      gc[q] = gdir; gb[q] -= gdir;
    }
    // the measured dihedral m = atan2(p1, p2) of layers.py:800-808, dL/dm = -g_theta.  With A = b1 x b2, B = b2 x b3:
    //   dm/da = -|b2| A / A^2,   dm/dd = |b2| B / B^2,
    //   dm/db = -(1 + b1.b2 / b2^2) dm/da + (b3.b2 / b2^2) dm/dd,   dm/dc = (b1.b2 / b2^2) dm/da - (1 + b3.b2 / b2^2) dm/dd
    double b1[3], b3[3], A[3], B[3];
#pragma unroll
    for (int q = 0; q < 3; q++) { b1[q] = xb[q] - xa[q]; b3[q] = xd[q] - xc[q]; }
    cross3d(b1, dv, A);
    cross3d(dv, b3, B);
    const double a2 = dot3d(A, A), b2n = dot3d(B, B);
    if (a2 > 0.0 && b2n > 0.0) {
      const double ka = g_theta * len / a2, kd = -g_theta * len / b2n;      // -g_theta dm/da = ka A,  -g_theta dm/dd = kd B
      const double fa = dot3d(b1, dv) * il * il, fb = dot3d(b3, dv) * il * il;
#pragma unroll
      for (int q = 0; q < 3; q++) {
        const double pa = ka * A[q], pd = kd * B[q];
        ga[q] = pa;
        gd[q] = pd;
        gb[q] += fb * pd - (1.0 + fa) * pa;
        gc[q] += fa * pa - (1.0 + fb) * pd;
      }
    }
  }
  // all loads first, then all stores (a, b, c, d are different atoms; written as += the compiler must assume they may alias and
  // serialises twelve load - add - store round trips)
  const bool four = kind > kSideAngle;
  double oa[3], ob[3], oc[3], od[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int q = 0; q < 3; q++) { oa[q] = gf[3 * a + q]; ob[q] = gf[3 * b + q]; oc[q] = gf[3 * c + q]; }
  if (four) {
#pragma unroll
    for (int q = 0; q < 3; q++) od[q] = gf[3 * d + q];
  }
#pragma unroll
  for (int q = 0; q < 3; q++) { gf[3 * a + q] = oa[q] + ga[q]; gf[3 * b + q] = ob[q] + gb[q]; gf[3 * c + q] = oc[q] + gc[q]; }
  if (four) {
#pragma unroll
    for (int q = 0; q < 3; q++) gf[3 * d + q] = od[q] + gd[q];
  }
  return g_target;
}

// backbone steps [k0, k1) in reverse, by the whole CTA; all bond angles (DIH = false) or all dihedrals
template <bool DIH>
__device__ __forceinline__ void sc_backbone_steps_bwd(const ScParams& p, int64_t f, int k0, int k1, double* xf, double* gf, const double* tg,
                                                      const unsigned char* flags, double* tr, double* red) {
  const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5;
  if (k1 <= k0) return;
  constexpr int kind = DIH ? kCentralDihedral : kCentralAngle;
  constexpr bool dihedral = DIH;
  constexpr int src = DIH ? 2 : 1;                               // input_of(kind): where dL/dtarget of these steps goes
  float* const gdst = p.gin[src] ? p.gin[src] + f * p.cols[src] : nullptr;
  int4 o0 = __ldg(p.ops + 3 * (k1 - 1)), o1 = __ldg(p.ops + 3 * (k1 - 1) + 1), o2 = __ldg(p.ops + 3 * (k1 - 1) + 2);
  if (tid == 0) sc_publish_inverse(xf, tg, tr, k1 - 1, kind, o0.z, o0.w);
  __syncthreads();
  for (int k = k1 - 1; k >= k0; k--) {
    int4 n0 = o0, n1 = o1, n2 = o2;
    if (k > k0) { n0 = __ldg(p.ops + 3 * k - 3); n1 = __ldg(p.ops + 3 * k - 2); n2 = __ldg(p.ops + 3 * k - 1); }
    const int c0 = o1.w - o1.z, total = c0 + (o2.y - o2.x);
    {
      // every warp leaves its partial sums (zeros when none of its lanes has a moving atom of this step): thread 0 then adds a
      // FIXED number of slots with independent loads -- a loop over "the warps that had work" is a chain of dependent
      // LDS + DADD round trips (measured: 1 600 of the 3 800 cycles of a step, profiles/r02_sidechain_backmap.txt)
      double acc[kRed] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      if (warp * 32 < total) {
        const Rot rot = sc_load_rot(tr);
        for (int e = tid; e < total; e += nth) {
          const int at = e < c0 ? o1.z + e : o2.x + (e - c0);
          if (DIH) sc_atom_bwd(xf, gf, rot, at, true, acc);
          else sc_atom_bwd_planar(xf, gf, rot, at, acc);
        }
#pragma unroll
        for (int q = 0; q < kRed; q++) {
          if (q >= 3 && !dihedral) break;
#pragma unroll
          for (int m = 16; m >= 1; m >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], m);
        }
      }
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < kRed; q++) red[warp * kRed + q] = acc[q];
      }
    }
    __syncthreads();
    if (tid == 0) {
      double part[(kScThreads / 32) * kRed];
#pragma unroll
      for (int q = 0; q < (kScThreads / 32) * kRed; q++) part[q] = red[q];
      double sum[kRed];
#pragma unroll
      for (int q = 0; q < kRed; q++) {
        double v = part[q];
#pragma unroll
        for (int w = 1; w < kScThreads / 32; w++) v += part[w * kRed + q];
        sum[q] = v;
      }
      const double g_target = sc_finish_step(xf, gf, kind, o0.y, o0.z, o0.w, o1.x, sum, tg[2 * k], flags[k]);
      if (gdst) gdst[o1.y] = (float)g_target;
    } else if (tid == 32 && k > k0) {
      // meanwhile another warp prepares the rotation of the next (earlier) step: it only needs the restored coordinates
      sc_publish_inverse(xf, tg, tr, k - 1, kind, n0.z, n0.w);
    }
    __syncthreads();
    o0 = n0; o1 = n1; o2 = n2;
  }
}

// side-chain steps of one kind in reverse, one thread per residue with a side chain
template <bool DIH>
__device__ __forceinline__ void sc_side_steps_bwd(const ScParams& p, int64_t f, int k_base, double* xf, double* gf, const double* tg,
                                                  const unsigned char* flags) {
  const int kind = DIH ? kSideDihedral : kSideAngle;
  const int src = input_of(kind);
  float* gdst = p.gin[src];
  for (int ri = threadIdx.x; ri < p.n_res_side; ri += blockDim.x) {
    const int4 r = __ldg(p.res + ri);
    const int end = r.y + r.z + 1;
    const int k0 = k_base + (DIH ? r.w : r.y - p.n_bb);
    const int steps = DIH ? r.z : r.z + 1;
    for (int q = steps - 1; q >= 0; q--) {
      const int k = k0 + q;
      const int a = sc_chain(r, q), b = sc_chain(r, q + 1), c = sc_chain(r, q + 2), d = DIH ? sc_chain(r, q + 3) : -1;
      double tr[TR_N];
      sc_publish_inverse(xf, tg, tr, k, kind, b, c);
      double acc[kRed] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      const Rot rot = sc_load_rot(tr);
      for (int at = r.y + q; at < end; at++) sc_atom_bwd(xf, gf, rot, at, DIH, acc);
      const double g_target = sc_finish_step(xf, gf, kind, a, b, c, d, acc, tg[2 * k], flags[k]);
      if (gdst) gdst[f * p.cols[src] + k - k_base] = (float)g_target;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kScThreads) sidechain_bwd_kernel(const ScParams p) {
  extern __shared__ double sc_smem[];
  double* xf = sc_smem;
  double* gf = xf + 3 * (size_t)p.n_atoms;
  double* tg = gf + 3 * (size_t)p.n_atoms;
  double* tr = tg + 2 * (size_t)p.n_ops;
  double* red = tr + TR_N;                 // (blockDim / 32) x kRed
  unsigned char* flags = reinterpret_cast<unsigned char*>(red + (kScThreads / 32) * kRed);
  const int tid = threadIdx.x, nth = blockDim.x;
  const int k_cd = p.n_ca + p.n_side;
  for (int64_t f = blockIdx.x; f < p.frames; f += gridDim.x) {
    __syncthreads();
    if (p.saved) {
      const double* sv = p.saved + f * p.save_stride;
      for (int e = tid; e < 2 * p.n_ops; e += nth) tg[e] = sv[e];
      const unsigned char* sf = reinterpret_cast<const unsigned char*>(sv + 2 * (size_t)p.n_ops);
      for (int e = tid; e < p.n_ops; e += nth) flags[e] = sf[e];
      const double* sx = sv + 2 * (size_t)p.n_ops + (p.n_ops + 7) / 8;
      for (int e = tid; e < 3 * p.n_atoms; e += nth) xf[e] = sx[e];
    } else {
      sc_layout(p, f, xf, tg, flags);
      sc_forward<true>(p, xf, tg, tr);
    }
    const float* gsrc = p.gout + f * (int64_t)(3 * p.n_atoms);
    for (int e = tid; e < 3 * p.n_atoms; e += nth) gf[e] = (double)__ldg(gsrc + e);
    __syncthreads();
    sc_side_steps_bwd<true>(p, f, k_cd + p.n_cd, xf, gf, tg, flags);
    sc_backbone_steps_bwd<true>(p, f, k_cd, k_cd + p.n_cd, xf, gf, tg, flags, tr, red);
    sc_side_steps_bwd<false>(p, f, p.n_ca, xf, gf, tg, flags);
    sc_backbone_steps_bwd<false>(p, f, 0, p.n_ca, xf, gf, tg, flags, tr, red);
    // the planar layout: x of backbone atom k = sum of the bonds before it, x of a side-chain atom = x of its CA, y = running sum
    // of its own chain's bonds (layers.py:593-628)
    float* g_sd = p.gin[3];
    for (int m = tid; m < p.n_side; m += nth) {
      const int4 s = __ldg(p.side + m);
      if (g_sd) {
        double y = 0.0;
        for (int q = p.n_bb + m; q < s.z; q++) y += gf[3 * q + 1];
        g_sd[f * p.cols[3] + m] = (float)y;
      }
    }
    __syncthreads();
    float* g_cd = p.gin[0];
    if (g_cd) {
      // every side-chain atom starts at the x of its CA: one thread per residue folds its chain into the CA
      for (int ri = tid; ri < p.n_res_side; ri += nth) {
        const int4 r = __ldg(p.res + ri);
        double v = 0.0;
        for (int q = r.y; q <= r.y + r.z; q++) v += gf[3 * q];
        gf[3 * (r.x + 1)] += v;
      }
      __syncthreads();
      if (tid < 32) {     // d x_k / d bond_i = 1 for k > i: suffix sums, warp scan over 32 atoms at a time from the far end
        double carry = 0.0;
        for (int base = p.n_bb - 1; base >= 1; base -= 32) {
          const int k = base - tid;
          double v = k >= 1 ? gf[3 * k] : 0.0;
#pragma unroll
          for (int m = 1; m < 32; m <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, v, m);
            if (tid >= m) v += up;
          }
          v += carry;
          if (k >= 1) g_cd[f * p.cols[0] + k - 1] = (float)v;
          carry = __shfl_sync(0xffffffffu, v, 31);
        }
      }
    }
  }
}

// ---- gather / scatter of atoms (PairwiseDistances with reconstruct_sidechains, layers.py:1260-1265) ---------------------------
__global__ void gather_atoms_kernel(const float* __restrict__ x, int64_t b, int n, const int* __restrict__ index, int m, float* __restrict__ out) {
  const int64_t total = b * (int64_t)m * 3;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t fr = e / (3 * m);
    const int r = (int)(e - fr * 3 * m);
    const int at = __ldg(index + r / 3);
    out[e] = (at >= 0 && at < n) ? __ldg(x + (fr * n + at) * 3 + r % 3) : __int_as_float(0x7fc00000);
  }
}
__global__ void scatter_atoms_kernel(const float* __restrict__ g, int64_t b, int n, const int* __restrict__ index, int m, float* __restrict__ gx) {
  const int64_t total = b * (int64_t)m * 3;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t fr = e / (3 * m);
    const int r = (int)(e - fr * 3 * m);
    const int at = __ldg(index + r / 3);
    if (at >= 0 && at < n) atomicAdd(gx + (fr * n + at) * 3 + r % 3, __ldg(g + e));
  }
}

int gather_atoms_device(const float* x, int64_t b, int64_t n, const int32_t* index_dev, int64_t m, float* out, cudaStream_t st) {
  EMK_REQUIRE(b >= 0 && n >= 1 && m >= 0 && n < (1 << 28) && m < (1 << 28), EMK_E_SHAPE, "emk_gather_atoms: bad shape");
  if (b == 0 || m == 0) return EMK_OK;
  EMK_REQUIRE(x && index_dev && out, EMK_E_NULL, "emk_gather_atoms: NULL pointer argument");
  const int64_t total = b * m * 3;
  const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
  gather_atoms_kernel<<<grid, 256, 0, st>>>(x, b, (int)n, index_dev, (int)m, out);
  return launch_status("gather_atoms_kernel");
}
int gather_atoms_bwd_device(const float* grad_out, int64_t b, int64_t n, const int32_t* index_dev, int64_t m, float* grad_x, cudaStream_t st) {
  EMK_REQUIRE(b >= 0 && n >= 1 && m >= 0 && n < (1 << 28) && m < (1 << 28), EMK_E_SHAPE, "emk_gather_atoms_bwd: bad shape");
  if (b == 0) return EMK_OK;
  EMK_REQUIRE(grad_x && (m == 0 || (grad_out && index_dev)), EMK_E_NULL, "emk_gather_atoms_bwd: NULL pointer argument");
  EMK_CUDA(cudaMemsetAsync(grad_x, 0, (size_t)(b * n * 3) * sizeof(float), st));
  if (m == 0) return EMK_OK;
  const int64_t total = b * m * 3;
  const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
  scatter_atoms_kernel<<<grid, 256, 0, st>>>(grad_out, b, (int)n, index_dev, (int)m, grad_x);
  return launch_status("scatter_atoms_kernel");
}

// ---- host entry points ------------------------------------------------------------------------------------------------------------
int sidechain_plan_create(int64_t n_res, const int32_t* counts, SidechainPlan** out) {
  EMK_REQUIRE(out, EMK_E_NULL, "emk_sidechain_plan_create: NULL plan pointer");
  *out = nullptr;
  SidechainPlan* pl = new SidechainPlan();
  int rc = build_plan(n_res, counts, pl);
  if (rc) { delete pl; return rc; }
  // the device copy is optional: without a GPU the plan can still be inspected (emk_sidechain_plan_info / _ops)
  int dev = -1;
  if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0) {
    const size_t n_ints = pl->ops.size() + pl->side.size() + pl->res.size();
    int* mem = nullptr;
    cudaError_t e = cudaMalloc(&mem, n_ints * sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpy(mem, pl->ops.data(), pl->ops.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(mem + pl->ops.size(), pl->side.data(), pl->side.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
      e = cudaMemcpy(mem + pl->ops.size() + pl->side.size(), pl->res.data(), pl->res.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      if (mem) cudaFree(mem);
      cudaGetLastError();
      mem = nullptr;
      dev = -1;
    }
    pl->d_mem = mem;
    pl->device = dev;
  } else {
    cudaGetLastError();
  }
  *out = pl;
  return EMK_OK;
}
void sidechain_plan_destroy(SidechainPlan* pl) {
  if (!pl) return;
  if (pl->d_mem) {
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != pl->device) cudaSetDevice(pl->device);
    cudaFree(pl->d_mem);
    if (cur >= 0 && cur != pl->device) cudaSetDevice(cur);
  }
  delete pl;
}
int sidechain_plan_info(const SidechainPlan* pl, int64_t* info) {
  EMK_REQUIRE(pl && info, EMK_E_NULL, "emk_sidechain_plan_info: NULL pointer argument");
  info[0] = pl->n_atoms; info[1] = pl->n_side; info[2] = pl->n_ops; info[3] = pl->n_res;
  for (int k = 0; k < 6; k++) info[4 + k] = pl->cols[k];
  return EMK_OK;
}
int sidechain_plan_ops(const SidechainPlan* pl, int32_t* ops) {
  EMK_REQUIRE(pl && ops, EMK_E_NULL, "emk_sidechain_plan_ops: NULL pointer argument");
  std::copy(pl->ops.begin(), pl->ops.end(), ops);
  return EMK_OK;
}
const int* sidechain_plan_cols(const SidechainPlan* pl) { return pl->cols; }
int sidechain_plan_atoms(const SidechainPlan* pl) { return pl->n_atoms; }

static int fill_params(const char* who, const SidechainPlan* pl, const float* const* in, int64_t frames, ScParams* p) {
  EMK_REQUIRE(pl, EMK_E_NULL, "%s: NULL plan", who);
  EMK_REQUIRE(frames >= 0, EMK_E_SHAPE, "%s: negative frame count", who);
  int dev = -1;
  EMK_CUDA(cudaGetDevice(&dev));
  EMK_REQUIRE(pl->d_mem && pl->device == dev, EMK_E_DEVICE, "%s: the plan was created on device %d, the current device is %d", who, pl->device, dev);
  for (int k = 0; k < 6; k++) {
    EMK_REQUIRE(frames == 0 || pl->cols[k] == 0 || in[k], EMK_E_NULL, "%s: NULL input %d", who, k);
    p->in[k] = in[k];
    p->cols[k] = pl->cols[k];
  }
  p->ops = reinterpret_cast<const int4*>(pl->d_mem);
  p->side = reinterpret_cast<const int4*>(pl->d_mem + pl->ops.size());
  p->res = reinterpret_cast<const int4*>(pl->d_mem + pl->ops.size() + pl->side.size());
  p->n_res_side = (int)(pl->res.size() / 4); p->n_ca = pl->n_ca; p->n_cd = pl->n_cd;
  p->frames = frames; p->n_bb = pl->n_bb; p->n_side = pl->n_side; p->n_atoms = pl->n_atoms; p->n_ops = pl->n_ops;
  return EMK_OK;
}

static unsigned frames_grid(int64_t frames, size_t smem) {
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(12, (227 * 1024) / (smem + 1024)));
  return (unsigned)std::min<int64_t>(frames, (int64_t)sm_count() * per_sm);
}

int64_t sidechain_saved_doubles(const SidechainPlan* pl) { return pl ? sc_saved_doubles(pl->n_atoms, pl->n_ops) : -1; }

int sidechain_backmap_device(const SidechainPlan* pl, const float* const* in, int64_t frames, float* out, double* save, cudaStream_t st) {
  ScParams p{};
  int rc = fill_params("emk_sidechain_backmap", pl, in, frames, &p);
  if (rc) return rc;
  if (frames == 0) return EMK_OK;
  EMK_REQUIRE(out, EMK_E_NULL, "emk_sidechain_backmap: NULL output");
  const size_t smem = (3 * (size_t)pl->n_atoms + 2 * (size_t)pl->n_ops + TR_N) * sizeof(double) + (save ? (((size_t)pl->n_ops + 15) & ~(size_t)15) : 0);
  EMK_REQUIRE(smem <= 227 * 1024, EMK_E_UNSUPPORTED, "emk_sidechain_backmap: %d atoms need %zu bytes of shared memory per frame (limit 227 KB)", pl->n_atoms, smem);
  p.out = out;
  p.save = save;
  p.save_stride = sc_saved_doubles(pl->n_atoms, pl->n_ops);
  static bool cfg[kMaxDevices] = {false};
  if (first_use_on_device(cfg)) {
    EMK_CUDA(cudaFuncSetAttribute(sidechain_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    EMK_CUDA(cudaFuncSetAttribute(sidechain_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  if (save) sidechain_fwd_kernel<true><<<frames_grid(frames, smem), 128, smem, st>>>(p);
  else sidechain_fwd_kernel<false><<<frames_grid(frames, smem), 128, smem, st>>>(p);
  return launch_status("sidechain_fwd_kernel");
}

int sidechain_backmap_bwd_device(const SidechainPlan* pl, const float* const* in, int64_t frames, const float* grad_out, const double* saved,
                                 float* const* grad_in, cudaStream_t st) {
  ScParams p{};
  int rc = fill_params("emk_sidechain_backmap_bwd", pl, in, frames, &p);
  if (rc) return rc;
  if (frames == 0) return EMK_OK;
  EMK_REQUIRE(grad_out, EMK_E_NULL, "emk_sidechain_backmap_bwd: NULL grad_out");
  const size_t smem = (6 * (size_t)pl->n_atoms + 2 * (size_t)pl->n_ops + TR_N + (kScThreads / 32) * kRed) * sizeof(double) + (((size_t)pl->n_ops + 15) & ~(size_t)15);
  EMK_REQUIRE(smem <= 227 * 1024, EMK_E_UNSUPPORTED, "emk_sidechain_backmap_bwd: %d atoms / %d steps need %zu bytes of shared memory per frame (limit 227 KB)",
              pl->n_atoms, pl->n_ops, smem);
  p.gout = grad_out;
  p.saved = saved;
  p.save_stride = sc_saved_doubles(pl->n_atoms, pl->n_ops);
  for (int k = 0; k < 6; k++) p.gin[k] = grad_in[k];
  static bool cfg[kMaxDevices] = {false};
  if (first_use_on_device(cfg)) EMK_CUDA(cudaFuncSetAttribute(sidechain_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  sidechain_bwd_kernel<<<frames_grid(frames, smem), kScThreads, smem, st>>>(p);
  return launch_status("sidechain_bwd_kernel");
}

// PairwiseDistances.__init__ with reconstruct_sidechains (layers.py:1188-1208): the sliced backbone plus one index per residue with
// a side chain; the running index starts at 3 n + 1 and advances by the residue's number of side-chain DIHEDRALS (as the reference
// does).  out may be NULL to ask for the count.
int64_t sidechain_pairwise_indices(int64_t n_res, const int32_t* counts, int64_t first, int64_t count, int64_t step, int64_t* out) {
  int64_t k = 0;
  for (int64_t i = 0; i < count; i++, k++)
    if (out) out[k] = first + i * step;
  int64_t atom = 3 * n_res + 1;
  for (int64_t r = 0; r < n_res; r++) {
    if (counts[r] == 0) continue;
    atom += counts[r];
    if (out) out[k] = atom;
    k++;
  }
  return k;
}

}  // namespace emk
