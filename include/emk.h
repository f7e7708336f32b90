/* libemk -- C ABI of the B200 (sm_100a) kernels for EncoderMap's training hot path.
 *
 * The reference (AG-Peter/encodermap) has no FFI for this path: the boundary is a set of
 * plain Python callables taking/returning tensors.  Every entry point
 * below names the reference callable (file:line under /root/reference) whose arithmetic it
 * replaces; the Python shims in encodermap_b200/ keep those callables' signatures and
 * call in here through ctypes.
 *
 * Conventions
 *   - return 0 on success; negative EMK_E_* for argument errors (validated before any
 *     launch); positive values are cudaError_t.  emk_last_error() gives a thread-local
 *     message for the last non-zero return.  Nothing throws or aborts.
 *   - every tensor is float32, C-contiguous, on the CURRENT CUDA device unless stated;
 *     caller allocates all inputs and outputs; libemk borrows pointers for the call only.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy
 *     default stream); no entry point synchronises the device except the *_host ones.
 *   - `emk_dl_*` twins take DLPack tensors (`DLManagedTensor*` from the capsule a framework
 *     exports; the capsule is NOT consumed) and validate dtype/device/contiguity/shape.
 *   - there is no CPU fallback: on a machine without a CUDA device compute calls fail with
 *     the CUDA error; the host-only helpers (index construction) always work.
 */
#ifndef EMK_H_
#define EMK_H_

#include <stdint.h>
#include "emk_dlpack.h"

#ifdef __cplusplus
extern "C" {
#endif

#define EMK_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define EMK_API __attribute__((visibility("default")))
#else
#define EMK_API
#endif

/* error codes */
#define EMK_OK 0
#define EMK_E_NULL (-1)        /* required pointer is NULL */
#define EMK_E_DTYPE (-2)       /* DLPack tensor is not float32 / float64 as required */
#define EMK_E_DEVICE (-3)      /* DLPack tensor is not on a CUDA device */
#define EMK_E_SHAPE (-4)       /* rank / extent mismatch */
#define EMK_E_CONTIG (-5)      /* tensor is not C-contiguous */
#define EMK_E_ARG (-6)         /* scalar argument out of range */
#define EMK_E_UNSUPPORTED (-7) /* valid request this build does not implement */

/* flags for emk_sigmoid_cost */
#define EMK_COST_ZERO_OUTPUTS 1u /* memset loss and grad_low on the stream before accumulating */
#define EMK_COST_NO_GRAD 2u      /* forward only: grad_low may be NULL */

EMK_API int emk_version(void);
EMK_API const char* emk_last_error(void);
/* "sm_100a;<nvcc version>;<build date>" */
EMK_API const char* emk_build_info(void);
/* Measurement aid (bench.py): FP32 FMA lane-instructions per second this device sustains on register-only FFMA
 * chains -- the measured denominator of the pair-tile kernel's FP32-issue roofline.  Synchronises the device. */
EMK_API int emk_probe_fp32(double* lane_instr_per_s);
/* Process-wide tuning knobs (kernel selection thresholds; results do not depend on them beyond float32 rounding):
 *   "backmap_fwd6_min_batch"  batch size from which emk_backmap uses the lane-per-frame kernel (default 4096; 0 = whenever
 *                             the shape is eligible, negative = never)
 *   "backmap_fwd6_warps"      warps per CTA of that kernel: 0 (default: chosen per launch so that the frame tiles fill whole
 *                             waves), or one of 8, 12, 14, 16, 18, 20
 *   "backmap_fwd6_f32_extent_nm"  0 (default): float64 chain, error ~2e-6 nm.  v > 0 (opt-in, <= 64): the lane-per-frame kernel runs
 *                             its first pass with a float32 chain relative to the anchor atom (error <= 4e-6 x extent) and repeats
 *                             a frame tile in float64 when a side extends further than v nm from its anchor; measured slower
 *                             than the default on B200 at v = 16
 *   "cost_small_tile_max_rows"  emk_sigmoid_cost calls that cover the whole tile list of up to this many rows use 64 x 32 pair tiles
 *                             instead of 128 x 64 (default 1024, and only up to 2048 columns; 0 = never; >= 2^20 = always).  Tile
 *                             ranges always refer to the 128 x 64 numbering.
 *   "cost_small_d_max"        widest high-d input (columns) for which emk_sigmoid_cost uses the register kernel instead of
 *                             the TMA pair-tile kernel (default 8; 0..8) */
EMK_API int emk_set_option(const char* name, int64_t value);
EMK_API int emk_get_option(const char* name, int64_t* value);

/* ------------------------------------------------------------------------------------------
 * Host-only index construction (integer work: bit-exact contracts; no GPU needed)
 * ---------------------------------------------------------------------------------------- */

/* Pair order of pairwise_dist(flat=True): strict upper triangle, row-major
 * (encodermap/misc/distances.py:235-242).  count = n(n-1)/2; i_out/j_out hold `count` entries. */
EMK_API int64_t emk_triu_pair_count(int64_t n);
EMK_API int emk_triu_pair_indices(int64_t n, int32_t* i_out, int32_t* j_out);

/* Atom / dihedral index lists of the two-sided back-mapping
 * (split_and_reverse_cartesians / split_and_reverse_dihedrals, encodermap/misc/backmapping.py:179-256;
 * loop counts left_split = n/2-1, right_split = (n-3)/2, encodermap/models/models.py:661-671).
 * counts[4] = {n_left_atoms, n_right_atoms, n_left_dihedrals, n_right_dihedrals}. */
EMK_API int emk_backmap_split_counts(int64_t n_atoms, int64_t counts[4]);
EMK_API int emk_backmap_split_indices(int64_t n_atoms, int32_t* left_atoms, int32_t* right_atoms,
                              int32_t* left_dihedrals, int32_t* right_dihedrals);

/* Work decomposition of the all-pairs cost: the N x N pair space is cut into tiles of
 * EMK_TILE_ROWS x EMK_TILE_COLS; only tiles that touch the upper triangle are evaluated.
 * Tiles are numbered band by band (8 tile rows per band, column-major inside a band, so that
 * concurrently running tiles share their column block in L2); a multi-GPU run gives rank r the
 * contiguous range emk_pair_tile_range(n, r, world). */
#define EMK_TILE_ROWS 128
#define EMK_TILE_COLS 64
EMK_API int64_t emk_pair_tile_count(int64_t n_rows);
EMK_API int emk_pair_tile_decode(int64_t n_rows, int64_t tile, int64_t* tile_row, int64_t* tile_col);
EMK_API int emk_pair_tile_range(int64_t n_rows, int rank, int world, int64_t* begin, int64_t* end);

/* ------------------------------------------------------------------------------------------
 * All-pairs sketch-map sigmoid cost, fused forward + backward
 *   replaces sigmoid_loss_func (encodermap/loss_functions/loss_functions.py:335-369) and through
 *   it distance_loss_func (:266-296) and cartesian_distance_loss_func (:917-942); fuses
 *   pairwise_dist_periodic / pairwise_dist (encodermap/misc/distances.py:144-255) and
 *   sigmoid (:66-88).  No N x N matrix is written.
 *
 *   high (n,d)   high-dimensional rows;   low (n,l) latent rows, any l >= 1 (walked in chunks of 8 components)
 *   periodicity  +inf => Euclidean high-d distances
 *   sig[6]       sig_h, a_h, b_h, sig_l, a_l, b_l
 *   tile_begin/tile_end   slice of the tile list to evaluate (0, emk_pair_tile_count(n) for all)
 *   loss         device double[1]:  += sum over the slice of (s_h - s_l)^2 / n^2
 *   grad_low     device (n,l):      += d(loss)/d(low) contributions of the slice
 * ---------------------------------------------------------------------------------------- */
EMK_API int emk_sigmoid_cost(const float* high, int64_t n, int64_t d, const float* low, int64_t l,
                     double periodicity, const float sig[6], int64_t tile_begin, int64_t tile_end,
                     double* loss, float* grad_low, uint32_t flags, void* stream);
EMK_API int emk_dl_sigmoid_cost(const DLManagedTensor* high, const DLManagedTensor* low, double periodicity,
                        const float sig[6], int64_t tile_begin, int64_t tile_end,
                        DLManagedTensor* loss, DLManagedTensor* grad_low, uint32_t flags, void* stream);
/* Same through HOST buffers: copies in, evaluates all tiles, copies out, synchronises.
 * loss_host: double[1]; grad_low_host: (n,l) or NULL. */
EMK_API int emk_sigmoid_cost_host(const float* high_host, int64_t n, int64_t d, const float* low_host, int64_t l,
                          double periodicity, const float sig[6], double* loss_host, float* grad_low_host);

/* ------------------------------------------------------------------------------------------
 * Distance matrices (API-compatible standalone ops; the fused cost never calls them)
 * ---------------------------------------------------------------------------------------- */

/* pairwise_dist_periodic(positions (n,d), periodicity) -> (n,n)   encodermap/misc/distances.py:144-176 */
EMK_API int emk_pairwise_dist_periodic(const float* x, int64_t n, int64_t d, double periodicity, float* out, void* stream);
EMK_API int emk_dl_pairwise_dist_periodic(const DLManagedTensor* x, double periodicity, DLManagedTensor* out, void* stream);
/* VJP of pairwise_dist_periodic w.r.t. positions (TensorFlow's autodiff of distances.py:164-175: abs' = sign, minimum routes
 * to its first operand on ties); dist is the forward output, grad_out the upstream (n,n) gradient. */
EMK_API int emk_pairwise_dist_periodic_bwd(const float* x, int64_t n, int64_t d, double periodicity, const float* dist,
                                           const float* grad_out, float* grad_x, void* stream);
EMK_API int emk_dl_pairwise_dist_periodic_bwd(const DLManagedTensor* x, double periodicity, const DLManagedTensor* dist,
                                              const DLManagedTensor* grad_out, DLManagedTensor* grad_x, void* stream);

/* pairwise_dist(positions (b,n,d), squared, flat)   encodermap/misc/distances.py:179-255
 *   x is addressed as x[bi*batch_stride + i*row_stride + k] (element strides) so that the atom
 *   selection inputs[:, start:stop:step] of PairwiseDistances.call (encodermap/models/layers.py:1252-1267)
 *   needs no copy.  out: (b,n,n) or, flat, (b, n(n-1)/2) in emk_triu_pair_indices order. */
EMK_API int emk_pairwise_dist(const float* x, int64_t b, int64_t n, int64_t d, int64_t batch_stride, int64_t row_stride,
                      int squared, int flat, float* out, void* stream);
/* grad_x (same addressing as x, must be zero-initialised by the caller when strided) +=
 * d(out)/d(x)^T grad_out;  zero distances get zero gradient (the reference's mask). */
EMK_API int emk_pairwise_dist_bwd(const float* x, int64_t b, int64_t n, int64_t d, int64_t batch_stride, int64_t row_stride,
                          int squared, int flat, const float* grad_out, float* grad_x, void* stream);
EMK_API int emk_dl_pairwise_dist(const DLManagedTensor* x, int64_t start, int64_t stop, int64_t step, int squared, int flat,
                         DLManagedTensor* out, void* stream);
EMK_API int emk_dl_pairwise_dist_bwd(const DLManagedTensor* x, int64_t start, int64_t stop, int64_t step, int squared, int flat,
                             const DLManagedTensor* grad_out, DLManagedTensor* grad_x, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused Cartesian branch of the ADC model: PairwiseDistances("output") -> cartesian_loss (+ ADCClashMetric),
 * forward and gradient w.r.t. the coordinates in ONE launch; the (b, n_pairs) distance matrices never exist in HBM.
 *   replaces, as a unit, PairwiseDistances.call (encodermap/models/layers.py:1252-1267) on the back-mapped coordinates,
 *   cartesian_loss_func (encodermap/loss_functions/loss_functions.py:1020-1065) and the clash count of ADCClashMetric.update_state
 *   (encodermap/callbacks/metrics.py:512-520) -- the caller, ADCFunctionalModel.get_loss (encodermap/models/models.py:2385-2387),
 *   passes the two coordinate tensors instead of the two distance matrices.
 *   xyz (b,n_atoms,3) back-mapped coordinates; selected atoms first + a * step, a < count (inputs[:, start:stop:step])
 *   target       target_is_xyz = 1: input coordinates (b,n_atoms,3) (their pair distances are formed on the fly)
 *                target_is_xyz = 0: input pair distances (b, count (count-1)/2) in emk_triu_pair_indices order
 *   variant      0 mean_abs, 1 mean_square, 2 mean_norm (p.cartesian_cost_variant)
 *   loss_sum     device double[1]: += sum over frames and pairs of |d_in - d_out| (0), (d_in - d_out)^2 (1), or sum over frames of
 *                the 2-norm over pairs (2); the caller divides by the element count and applies cost_reference / scale
 *   grad_xyz     (b,n_atoms,3) or NULL: d(loss_sum)/d(xyz), zero for unselected atoms (written, not accumulated)
 *   clashes      (b) int64 or NULL: pairs closer than clash_distance per frame (clash_distance <= 0: not counted)
 * ---------------------------------------------------------------------------------------- */
EMK_API int emk_cartesian_pair_loss(const float* xyz, int64_t b, int64_t n_atoms, int64_t first, int64_t count, int64_t step,
                                    const float* target, int target_is_xyz, int variant, float clash_distance, double* loss_sum,
                                    float* grad_xyz, int64_t* clashes, void* stream);
EMK_API int emk_dl_cartesian_pair_loss(const DLManagedTensor* xyz, int64_t start, int64_t stop, int64_t step, const DLManagedTensor* target,
                                       int variant, float clash_distance, DLManagedTensor* loss_sum, DLManagedTensor* grad_xyz,
                                       DLManagedTensor* clashes, void* stream);

/* ------------------------------------------------------------------------------------------
 * cartesian_distance_loss straight from the coordinates (no stored pair matrix)
 *   reference: encodermap/models/models.py:837-839 (inp_pair = PairwiseDistances(p, "input")(inp_cartesians)) followed by
 *   :2419-2422 (cartesian_distance_loss(model, p)(inp_pair, latent)), loss_functions.py:873-944.
 * Same result as emk_pairwise_dist(flat) on the selected atoms + emk_sigmoid_cost(periodicity = +inf); the (frames, n_pairs)
 * matrix lives in stream-ordered scratch between the two kernels (16-byte row pitch for TMA, zero-filled tail) and is released
 * before the call returns.  Arguments as emk_sigmoid_cost with the frames as the rows of the pair problem; tile ranges refer to
 * emk_pair_tile_count(b).  The selection needs at least two atoms.
 * ---------------------------------------------------------------------------------------- */
EMK_API int emk_cartesian_distance_cost(const float* xyz, int64_t b, int64_t n_atoms, int64_t first, int64_t count, int64_t step,
                                        const float* low, int64_t l, const float sig[6], int64_t tile_begin, int64_t tile_end, double* loss,
                                        float* grad_low, uint32_t flags, void* stream);
EMK_API int emk_dl_cartesian_distance_cost(const DLManagedTensor* xyz, int64_t start, int64_t stop, int64_t step, const DLManagedTensor* low,
                                           const float sig[6], int64_t tile_begin, int64_t tile_end, DLManagedTensor* loss,
                                           DLManagedTensor* grad_low, uint32_t flags, void* stream);

/* ------------------------------------------------------------------------------------------
 * Elementwise ops
 * ---------------------------------------------------------------------------------------- */

/* periodic_distance(a, b, periodicity) on equal-shape operands   encodermap/misc/distances.py:113-141 */
EMK_API int emk_periodic_distance(const float* a, const float* b, int64_t count, double periodicity, float* out, void* stream);
EMK_API int emk_periodic_distance_bwd(const float* a, const float* b, int64_t count, double periodicity, const float* grad_out,
                              float* grad_a, float* grad_b, void* stream);
EMK_API int emk_dl_periodic_distance(const DLManagedTensor* a, const DLManagedTensor* b, double periodicity, DLManagedTensor* out, void* stream);
EMK_API int emk_dl_periodic_distance_bwd(const DLManagedTensor* a, const DLManagedTensor* b, double periodicity,
                                 const DLManagedTensor* grad_out, DLManagedTensor* grad_a, DLManagedTensor* grad_b, void* stream);

/* sigmoid(sig,a,b)(r) elementwise and its derivative   encodermap/misc/distances.py:66-88 */
EMK_API int emk_sigmoid(const float* r, int64_t count, float sig, float a, float b, float* out, void* stream);
EMK_API int emk_sigmoid_bwd(const float* r, int64_t count, float sig, float a, float b, const float* grad_out, float* grad_r, void* stream);
EMK_API int emk_dl_sigmoid(const DLManagedTensor* r, float sig, float a, float b, DLManagedTensor* out, void* stream);
EMK_API int emk_dl_sigmoid_bwd(const DLManagedTensor* r, float sig, float a, float b, const DLManagedTensor* grad_out, DLManagedTensor* grad_r, void* stream);

/* PeriodicInput.call: (rows,d) -> (rows,2d) = [sin(x 2pi/P), cos(x 2pi/P)]   encodermap/models/layers.py:204-215 */
EMK_API int emk_periodic_input(const float* x, int64_t rows, int64_t d, double periodicity, float* out, void* stream);
EMK_API int emk_periodic_input_bwd(const float* x, int64_t rows, int64_t d, double periodicity, const float* grad_out, float* grad_x, void* stream);
EMK_API int emk_dl_periodic_input(const DLManagedTensor* x, double periodicity, DLManagedTensor* out, void* stream);
EMK_API int emk_dl_periodic_input_bwd(const DLManagedTensor* x, double periodicity, const DLManagedTensor* grad_out, DLManagedTensor* grad_x, void* stream);

/* rotation_matrix(axis_unit_vec (b,3), angle (b)) -> (b,3,3)   encodermap/misc/backmapping.py:1950-1968 */
EMK_API int emk_rotation_matrix(const float* axis, const float* angle, int64_t b, float* out, void* stream);
EMK_API int emk_dl_rotation_matrix(const DLManagedTensor* axis, const DLManagedTensor* angle, DLManagedTensor* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Generation-side back-mapping: amide H / carbonyl O guessed from the backbone, and the merge of backbone + guessed atoms
 *   encodermap/misc/backmapping.py:1920-1947 (guess_sp2_atom, guess_amide_H: angle 123 deg, length 1.10; guess_amide_O: 121 deg,
 *   1.24), :1970-1990 (merge_cartesians); TF1 twins encodermap/encodermap_tf1/backmapping.py:256-318.
 * Index lists are HOST arrays (int64).  Forward only (the reference uses these on the generate / summary path).  The index
 * lists are uploaded with a pageable copy, so these calls cannot be captured into a CUDA graph.
 *   emk_guess_sp2_atoms       out (b, n_idx, 3): one atom per entry of `indices` (centre atom i; neighbours i-1 and i+1, or
 *                             i-2 when i is the last atom; negative positions wrap as Python indexing does)
 *   emk_merge_cartesians      out (b, n + n_h + n_o, 3): atom 0, then for i = 1..n-1 atom i followed by the next hydrogen if i
 *                             is in h_after, else by the next oxygen if i is in o_after (h_after = N_indices[1:], o_after =
 *                             O_indices in the reference's call); EMK_E_SHAPE unless every supplied atom is placed exactly once
 *   emk_backbone_amide_atoms  guess_amide_H + guess_amide_O + merge_cartesians in one launch (the H / O arrays never exist);
 *                             n_out must equal emk_merged_atom_count(...)
 * ---------------------------------------------------------------------------------------- */
EMK_API int emk_guess_sp2_atoms(const float* xyz, int64_t b, int64_t n_atoms, const int64_t* indices, int64_t n_idx, double angle_to_previous,
                                double bond_length, float* out, void* stream);
EMK_API int emk_merge_cartesians(const float* central, int64_t b, int64_t n_atoms, const int64_t* h_after, int64_t n_h_after,
                                 const int64_t* o_after, int64_t n_o_after, const float* h_xyz, int64_t n_h, const float* o_xyz, int64_t n_o,
                                 float* out, void* stream);
EMK_API int emk_backbone_amide_atoms(const float* central, int64_t b, int64_t n_atoms, const int64_t* h_after, int64_t n_h_after,
                                     const int64_t* o_after, int64_t n_o_after, double h_angle, double h_length, double o_angle,
                                     double o_length, float* out, int64_t n_out, void* stream);
EMK_API int64_t emk_merged_atom_count(int64_t n_atoms, const int64_t* h_after, int64_t n_h_after, const int64_t* o_after, int64_t n_o_after);

/* ------------------------------------------------------------------------------------------
 * Topology-aware back-mapping: set the dihedrals of an all-atom structure, one after the other
 *   the numeric core of mdtraj_backmapping -- encodermap/misc/backmapping.py:1661-1690 (backbone dihedrals), :1722-1745
 *   (side-chain dihedrals; pass them after the backbone ones), dihedral formula misc/rotate.py:547-581, rotation matrix as in
 *   _rotmat_jit misc/backmapping.py:356-381.  For every frame and every dihedral j in order: rotate the atoms of far side j by
 *   (targets[frame, j] - current dihedral of quads[j]) about the axis bonds[j][0] -> bonds[j][1] through atom bonds[j][0].
 *   start        device (1 | frames, n_atoms, 3): the structure every frame starts from (one shared, or one per frame)
 *   quads, bonds HOST int32 (D, 4), (D, 2);  far_offsets HOST int32 (D + 1), far_atoms HOST int32: far side j is
 *                far_atoms[far_offsets[j] .. far_offsets[j + 1])  (misc/rotate.py:409-511 builds them from the bond graph)
 *   targets      device (frames, D) radians;   out device (frames, n_atoms, 3)
 * Coordinates are held in float64 between the rotations.  At most 9 600 atoms (shared memory).  Forward only; not capturable
 * (pageable index upload).
 * ---------------------------------------------------------------------------------------- */
EMK_API int emk_set_dihedrals(const float* start, int64_t start_frames, int64_t n_atoms, const int32_t* quads, const int32_t* bonds,
                              const int32_t* far_offsets, const int32_t* far_atoms, int64_t n_dihedrals, const float* targets,
                              int64_t frames, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Back-mapping WITH side chains: BackMapLayerWithSidechains, forward and exact backward
 *   encodermap/models/layers.py:218-843 (constructor :234-500, call :533-843; numpy twin _full_backmapping_np,
 *   encodermap/misc/backmapping.py:424-1002).  Atom order of the output: the 3 n backbone atoms (N, CA, C per residue), then per
 *   residue with a side chain its (dihedrals + 1) side-chain atoms.  All atoms start in the z = 0 plane (backbone on the x axis,
 *   side chains straight up from their CA); then every bond angle (backbone about +z, side chains about -z, by |target -
 *   measured|) and every dihedral (by target - measured, about its central bond) is set one after the other on the atoms behind
 *   the pivot, in the reference's order: backbone angles, side-chain angles, backbone dihedrals, side-chain dihedrals.
 *
 *   emk_sidechain_plan_create   HOST: n_side_dihedrals[r] = feature_description[-1][r + 1].  Builds the 12-int step table
 *                               (the reference's boolean masks, as index ranges) and uploads it to the CURRENT device (the plan
 *                               is bound to it; without a device the plan can still be inspected).  Refuses, with
 *                               EMK_E_UNSUPPORTED, the descriptions the reference's constructor cannot build either: no side
 *                               chain at all, not exactly one of the first / last residue without side chain, a residue without
 *                               side chain before the first one with.
 *   emk_sidechain_plan_info     info[10] = n_atoms, n_side_atoms, n_steps, n_residues, then the column counts of the six inputs
 *   emk_sidechain_plan_ops      ops (n_steps, 12) int32: kind (0 backbone angle, 1 side angle, 2 backbone dihedral, 3 side
 *                               dihedral), atoms a b c d (d = -1 for angles), input column, moving ranges [lo0, hi0) [lo1, hi1)
 *   emk_sidechain_backmap       inputs (frames, columns) float32 device, xyz (frames, n_atoms, 3).  One CTA per frame,
 *                               coordinates in shared memory as float64 for the whole sequence; capturable.
 *                               saved_state (optional, (frames, emk_sidechain_saved_size(plan)) float64): sin / cos of every
 *                               rotation and the float64 coordinates, for the backward pass
 *   emk_sidechain_backmap_bwd   exact VJP (any gradient pointer may be NULL): takes the forward state over from saved_state, or
 *                               re-runs the forward when that is NULL, then undoes the rotations in reverse order.  A bond angle measured on a straight triplet (1 - cos^2 < 1e-12) is a constant
 *                               (acos is not differentiable there; the reference's float32 autodiff gives 0, a huge number or
 *                               NaN depending on rounding).  At most ~3 500 atoms (shared memory).
 *   emk_sidechain_pairwise_indices  HOST: the atoms PairwiseDistances selects when side chains are reconstructed
 *                               (layers.py:1188-1208): backbone[start:stop:step] (INT64_MIN = None) plus one index per residue
 *                               with a side chain; returns the count (indices may be NULL), -1 on bad arguments
 *   emk_gather_atoms(_bwd)      out[b, k] = xyz[b, index[k]] (tf.gather, layers.py:1260-1265) and its VJP (scatter-add into a
 *                               zeroed grad_xyz); index is a DEVICE int32 array, an index outside [0, n_atoms) yields NaN
 * ---------------------------------------------------------------------------------------- */
typedef struct emk_sidechain_plan emk_sidechain_plan;
EMK_API int emk_sidechain_plan_create(int64_t n_residues, const int32_t* n_side_dihedrals, emk_sidechain_plan** plan);
EMK_API void emk_sidechain_plan_destroy(emk_sidechain_plan* plan);
EMK_API int emk_sidechain_plan_info(const emk_sidechain_plan* plan, int64_t* info);
EMK_API int emk_sidechain_plan_ops(const emk_sidechain_plan* plan, int32_t* ops);
EMK_API int emk_sidechain_backmap(const emk_sidechain_plan* plan, const float* central_distances, const float* central_angles,
                                  const float* central_dihedrals, const float* side_distances, const float* side_angles,
                                  const float* side_dihedrals, int64_t frames, float* xyz, double* saved_state, void* stream);
EMK_API int emk_sidechain_backmap_bwd(const emk_sidechain_plan* plan, const float* central_distances, const float* central_angles,
                                      const float* central_dihedrals, const float* side_distances, const float* side_angles,
                                      const float* side_dihedrals, int64_t frames, const float* grad_xyz, const double* saved_state,
                                      float* grad_central_distances, float* grad_central_angles, float* grad_central_dihedrals,
                                      float* grad_side_distances, float* grad_side_angles, float* grad_side_dihedrals, void* stream);
/* inputs / grad_inputs: arrays of six tensors in the order above (grad_inputs entries may be NULL) */
EMK_API int emk_dl_sidechain_backmap(const emk_sidechain_plan* plan, const DLManagedTensor* const* inputs, DLManagedTensor* xyz,
                                     DLManagedTensor* saved_state, void* stream);
EMK_API int emk_dl_sidechain_backmap_bwd(const emk_sidechain_plan* plan, const DLManagedTensor* const* inputs, const DLManagedTensor* grad_xyz,
                                         const DLManagedTensor* saved_state, DLManagedTensor* const* grad_inputs, void* stream);
EMK_API int64_t emk_sidechain_saved_size(const emk_sidechain_plan* plan);
EMK_API int64_t emk_sidechain_pairwise_indices(int64_t n_residues, const int32_t* n_side_dihedrals, int64_t start, int64_t stop,
                                               int64_t step, int64_t* indices);
EMK_API int emk_gather_atoms(const float* xyz, int64_t b, int64_t n_atoms, const int32_t* index_dev, int64_t m, float* out, void* stream);
EMK_API int emk_gather_atoms_bwd(const float* grad_out, int64_t b, int64_t n_atoms, const int32_t* index_dev, int64_t m, float* grad_xyz,
                                 void* stream);
EMK_API int emk_dl_gather_atoms(const DLManagedTensor* xyz, const DLManagedTensor* index, DLManagedTensor* out, void* stream);
EMK_API int emk_dl_gather_atoms_bwd(const DLManagedTensor* grad_out, const DLManagedTensor* index, DLManagedTensor* grad_xyz, void* stream);

/* mean over rows: (rows,cols) -> (cols)   the `tf.reduce_mean(distances, 0)` of BackMapLayer.call, layers.py:970 */
EMK_API int emk_column_mean(const float* x, int64_t rows, int64_t cols, float* out, void* stream);
EMK_API int emk_dl_column_mean(const DLManagedTensor* x, DLManagedTensor* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Back-mapping (internal coordinates -> Cartesian), SE(3) scan along the chain
 * ---------------------------------------------------------------------------------------- */

/* BackMapLayer.call fused   encodermap/models/layers.py:957-986  (= chain_in_plane ->
 * dihedrals + pi -> dihedrals_to_cartesian_tf_layers, encodermap/misc/backmapping.py:259-309)
 *   lengths: (n-1) shared by all frames when lengths_batch_stride == 0, else (b, n-1)
 *   angles (b,n-2), dihedrals (b,n-3) as the LAYER receives them (the +pi is applied inside)
 *   xyz (b,n,3) */
EMK_API int emk_backmap(const float* lengths, int64_t lengths_batch_stride, const float* angles, const float* dihedrals,
                int64_t b, int64_t n_atoms, float* xyz, void* stream);
/* exact VJP of emk_backmap from the final coordinates:
 *   grad_angles (b,n-2), grad_dihedrals (b,n-3), grad_lengths (b,n-1) per frame -- any may be NULL */
EMK_API int emk_backmap_bwd(const float* lengths, int64_t lengths_batch_stride, const float* angles, const float* xyz,
                    const float* grad_xyz, int64_t b, int64_t n_atoms, float* grad_angles, float* grad_dihedrals,
                    float* grad_lengths, void* stream);
EMK_API int emk_dl_backmap(const DLManagedTensor* lengths, const DLManagedTensor* angles, const DLManagedTensor* dihedrals,
                   DLManagedTensor* xyz, void* stream);
EMK_API int emk_dl_backmap_bwd(const DLManagedTensor* lengths, const DLManagedTensor* angles, const DLManagedTensor* xyz,
                       const DLManagedTensor* grad_xyz, DLManagedTensor* grad_angles, DLManagedTensor* grad_dihedrals,
                       DLManagedTensor* grad_lengths, void* stream);
EMK_API int emk_backmap_host(const float* lengths_host, const float* angles_host, const float* dihedrals_host, int64_t b,
                     int64_t n_atoms, float* xyz_host);

/* chain_in_plane(lengths, angles) -> (b,n,3), z = 0   encodermap/encodermap_tf1/backmapping.py:97-119 */
EMK_API int emk_chain_in_plane(const float* lengths, int64_t lengths_batch_stride, const float* angles, int64_t b,
                       int64_t n_atoms, float* xyz, void* stream);
EMK_API int emk_chain_in_plane_bwd(const float* lengths, int64_t lengths_batch_stride, const float* angles, const float* grad_xyz,
                           int64_t b, int64_t n_atoms, float* grad_angles, float* grad_lengths, void* stream);
EMK_API int emk_dl_chain_in_plane(const DLManagedTensor* lengths, const DLManagedTensor* angles, DLManagedTensor* xyz, void* stream);
EMK_API int emk_dl_chain_in_plane_bwd(const DLManagedTensor* lengths, const DLManagedTensor* angles, const DLManagedTensor* grad_xyz,
                              DLManagedTensor* grad_angles, DLManagedTensor* grad_lengths, void* stream);

/* dihedrals_to_cartesian_tf(dihedrals, cartesian) on an arbitrary start chain
 *   two-sided:  encodermap/encodermap_tf1/backmapping.py:164-195, encodermap/misc/backmapping.py:259-309
 *   one_way=1:  dihedral_to_cartesian_tf_one_way(_layers), encodermap/misc/backmapping.py:1873-1912
 *   dihedrals (b, n-3) exactly as the reference function receives them (no +pi here);
 *   chain: (n,3) shared when chain_batch_stride == 0, else (b,n,3) */
EMK_API int emk_dihedrals_to_cartesian(const float* dihedrals, const float* chain, int64_t chain_batch_stride, int64_t b,
                               int64_t n_atoms, int one_way, float* xyz, void* stream);
/* gradient w.r.t. the dihedrals, from the FINAL coordinates (twist of the downstream body about each bond) */
EMK_API int emk_dihedrals_to_cartesian_bwd(const float* xyz, const float* grad_xyz, int64_t b, int64_t n_atoms, int one_way,
                                   float* grad_dihedrals, void* stream);
EMK_API int emk_dl_dihedrals_to_cartesian(const DLManagedTensor* dihedrals, const DLManagedTensor* chain, int one_way,
                                  DLManagedTensor* xyz, void* stream);
EMK_API int emk_dl_dihedrals_to_cartesian_bwd(const DLManagedTensor* xyz, const DLManagedTensor* grad_xyz, int one_way,
                                      DLManagedTensor* grad_dihedrals, void* stream);
/* VJP of dihedrals_to_cartesian w.r.t. the START chain (grad_chain is (b, n_atoms, 3), one row per frame even when the
 * chain is shared: the caller sums over frames).  xyz is the forward output.  Reverse mode of the placement recursion,
 * one thread per frame in float64 (encodermap_tf1/backmapping.py:164-214 is differentiable in both arguments). */
EMK_API int emk_dihedrals_to_cartesian_chain_bwd(const float* chain, int64_t chain_batch_stride, const float* xyz, const float* grad_xyz,
                                                 int64_t b, int64_t n_atoms, int one_way, float* grad_chain, void* stream);
EMK_API int emk_dl_dihedrals_to_cartesian_chain_bwd(const DLManagedTensor* chain, const DLManagedTensor* xyz,
                                                    const DLManagedTensor* grad_xyz, int one_way, DLManagedTensor* grad_chain, void* stream);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU exchange (one process per GPU).  libemk keeps its own NCCL communicator, bound at run time
 * with dlopen("libnccl.so.2") -- the reference has no counterpart: its multi-device story is tf.distribute inside Keras
 * (encodermap/models/models.py:3367-3401 runs under whatever strategy the caller set up).
 *   emk_comm_unique_id : rank 0 fills 128 bytes (ncclUniqueId) and ships them to the other ranks by any means
 *   emk_comm_init      : every rank, with its CUDA device current; collective
 *   emk_comm_allreduce : full-set cost: sum of the float64 loss scalar and of grad_count float32 gradient values over all
 *                        ranks, in place, as ONE fused NCCL launch on `stream` (either pointer may be NULL)
 *   emk_comm_allgather2: data-parallel batch cost, exchange in: recv = concatenation over ranks of send, for two tensors
 *                        (high-d rows and latent rows) in one fused launch; counts are float32 elements per rank
 *   emk_comm_reduce_cost_scatter : exchange out: all-reduce of the loss + reduce-scatter of the (n,l) gradient (this rank keeps
 *                        count_per_rank values starting at rank * count_per_rank) in one fused launch
 * ---------------------------------------------------------------------------------------- */
EMK_API int emk_comm_unique_id(void* id_out_128_bytes);
EMK_API int emk_comm_init(int rank, int world, const void* nccl_unique_id);
EMK_API int emk_comm_info(int* rank, int* world);
EMK_API int emk_comm_allreduce(double* loss, float* grad, int64_t grad_count, void* stream);
EMK_API int emk_comm_allgather2(const float* a_send, float* a_recv, int64_t a_count, const float* b_send, float* b_recv,
                                int64_t b_count, void* stream);
EMK_API int emk_comm_reduce_cost_scatter(double* loss, const float* grad_full, float* grad_mine, int64_t count_per_rank, void* stream);
EMK_API int emk_comm_destroy(void);

#ifdef __cplusplus
}
#endif
#endif /* EMK_H_ */
