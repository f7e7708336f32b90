// extern "C" surface of libemk (include/emk.h): argument validation, DLPack views, host-only index
// construction, host-buffer convenience entry points.  Kernels live in pair_tile.cu, backmap.cu,
// elementwise.cu.
#include <cstring>
#include <mutex>
#include <vector>

#include "emk_common.cuh"

namespace emk {

// ---- error plumbing ----------------------------------------------------------------------------------
char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}
int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
    return n;
  return 148;
}

// Stream-ordered scratch from a PRIVATE memory pool per device whose release threshold is unlimited: the default
// pool trims back to zero at every synchronisation, so a scratch buffer requested once per call (padded copy of a
// D % 4 != 0 input, column-mean partials) cost a fresh device allocation each time -- 50 .. 500 us of jitter on calls
// that run for 100 us.  Blocks return to the pool with cudaFreeAsync as before.
int scratch_alloc(void** ptr, size_t bytes, cudaStream_t st) {
  static cudaMemPool_t pools[kMaxDevices] = {nullptr};
  static std::mutex mu;
  int dev = 0;
  EMK_CUDA(cudaGetDevice(&dev));
  EMK_REQUIRE(dev >= 0 && dev < kMaxDevices, EMK_E_UNSUPPORTED, "device ordinal %d out of range", dev);
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!pools[dev]) {
      cudaMemPoolProps props{};
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = dev;
      cudaMemPool_t pool;
      EMK_CUDA(cudaMemPoolCreate(&pool, &props));
      uint64_t keep = UINT64_MAX;
      EMK_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
      pools[dev] = pool;
    }
  }
  EMK_CUDA(cudaMallocFromPoolAsync(ptr, bytes, pools[dev], st));
  return EMK_OK;
}

bool first_use_on_device(bool* flags) {
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return true;
  std::lock_guard<std::mutex> lock(mu);
  if (flags[dev]) return false;
  flags[dev] = true;
  return true;
}

SigSpec make_sig_spec(float sig, float a, float b) {
  SigSpec s;
  s.inv_sig2 = (float)(1.0 / ((double)sig * (double)sig));
  s.c = (float)(std::pow(2.0, (double)a / (double)b) - 1.0);
  s.half_a = 0.5f * a;
  s.e = b / a;
  s.dcoef = (float)((double)b * (std::pow(2.0, (double)a / (double)b) - 1.0) / ((double)sig * (double)sig));
  s.a_int = (a == std::floor(a) && a >= 1.f && a <= 17.f) ? (int)a : 0;
  const float e2 = 2.f * b / a;
  s.e2_int = (e2 == std::floor(e2) && e2 >= 1.f && e2 <= 15.f) ? (int)e2 : 0;
  return s;
}

// device entry points implemented in the kernel files
int64_t pair_tile_count(int64_t n);
void tile_decode_host(int64_t t, int64_t n_rows, int64_t* I, int64_t* J);
int sigmoid_cost_device(const float*, int64_t, int64_t, const float*, int64_t, double, const float*, int64_t, int64_t, double*,
                        float*, uint32_t, cudaStream_t);
int dist_matrix_device(const float*, int64_t, int64_t, double, bool, int, float*, cudaStream_t);
int pairwise_flat3_warp_device(const float*, int64_t, int64_t, int64_t, int64_t, int, float*, int64_t, cudaStream_t);
int pairwise_warp_max_atoms();
void set_cost_small_d_max(int64_t v);
void set_cost_small_tile_max_rows(int64_t v);
int64_t cost_small_tile_max_rows();
int64_t cost_small_d_max();
int periodic_distance_device(const float*, const float*, int64_t, double, float*, cudaStream_t);
int periodic_distance_bwd_device(const float*, const float*, int64_t, double, const float*, float*, float*, cudaStream_t);
int sigmoid_device(const float*, int64_t, float, float, float, float*, cudaStream_t);
int sigmoid_bwd_device(const float*, int64_t, float, float, float, const float*, float*, cudaStream_t);
int periodic_input_device(const float*, int64_t, int64_t, double, float*, cudaStream_t);
int periodic_input_bwd_device(const float*, int64_t, int64_t, double, const float*, float*, cudaStream_t);
int rotation_matrix_device(const float*, const float*, int64_t, float*, cudaStream_t);
int guess_sp2_device(const float*, int64_t, int64_t, const int64_t*, int64_t, double, double, float*, cudaStream_t);
int merge_cartesians_device(const float*, int64_t, int64_t, const int64_t*, int64_t, const int64_t*, int64_t, const float*, int64_t,
                            const float*, int64_t, float*, cudaStream_t);
int backbone_amide_device(const float*, int64_t, int64_t, const int64_t*, int64_t, const int64_t*, int64_t, double, double, double, double,
                          float*, int64_t, cudaStream_t);
int64_t merged_atom_count(int64_t, const int64_t*, int64_t, const int64_t*, int64_t);
int set_dihedrals_device(const float*, int64_t, int64_t, const int32_t*, const int32_t*, const int32_t*, const int32_t*, int64_t, const float*,
                         int64_t, float*, cudaStream_t);
int column_mean_device(const float*, int64_t, int64_t, float*, cudaStream_t);
struct SidechainPlan;
int sidechain_plan_create(int64_t, const int32_t*, SidechainPlan**);
void sidechain_plan_destroy(SidechainPlan*);
int sidechain_plan_info(const SidechainPlan*, int64_t*);
int sidechain_plan_ops(const SidechainPlan*, int32_t*);
const int* sidechain_plan_cols(const SidechainPlan*);
int sidechain_plan_atoms(const SidechainPlan*);
int sidechain_backmap_device(const SidechainPlan*, const float* const*, int64_t, float*, double*, cudaStream_t);
int sidechain_backmap_bwd_device(const SidechainPlan*, const float* const*, int64_t, const float*, const double*, float* const*, cudaStream_t);
int64_t sidechain_saved_doubles(const SidechainPlan*);
int64_t sidechain_pairwise_indices(int64_t, const int32_t*, int64_t, int64_t, int64_t, int64_t*);
int gather_atoms_device(const float*, int64_t, int64_t, const int32_t*, int64_t, float*, cudaStream_t);
int gather_atoms_bwd_device(const float*, int64_t, int64_t, const int32_t*, int64_t, float*, cudaStream_t);
int pairwise_small_device(const float*, int64_t, int64_t, int64_t, int64_t, int64_t, int, int, float*, cudaStream_t);
int pairwise_small_bwd_device(const float*, int64_t, int64_t, int64_t, int64_t, int64_t, int, int, const float*, float*, cudaStream_t);
int backmap_fwd_device(const float*, int64_t, const float*, const float*, int64_t, int64_t, float*, cudaStream_t);
int fp32_probe_device(double*);
struct CartLossParams {
  const float* xyz;
  const float* target;
  int64_t b;
  int n_atoms;
  int first, ns, step;
  int target_is_xyz;
  int variant;
  float clash2;
  double* loss_sum;
  float* grad_xyz;
  long long* clashes;
};
int cart_pair_loss_device(const CartLossParams&, cudaStream_t);
int64_t fwd6_min_batch();
void set_fwd6_min_batch(int64_t);
int64_t fwd6_f32_extent();
void set_fwd6_f32_extent(int64_t);
int64_t fwd6_warps();
void set_fwd6_warps(int64_t);
int d2c_chain_bwd_device(const float*, int64_t, const float*, const float*, int64_t, int64_t, int, float*, cudaStream_t);
int pairwise_periodic_bwd_device(const float*, int64_t, int64_t, double, const float*, const float*, float*, cudaStream_t);
int chain_in_plane_device(const float*, int64_t, const float*, int64_t, int64_t, float*, cudaStream_t);
int d2c_general_device(const float*, const float*, int64_t, int64_t, int64_t, int, float*, cudaStream_t);
int backmap_bwd_device(const BwdParams&, cudaStream_t);

// ---- DLPack views ---------------------------------------------------------------------------------------
struct View {
  void* data;
  int ndim;
  int64_t shape[4];
  int64_t numel;
};

static int view_of(const DLManagedTensor* t, const char* name, int code, int bits, int min_rank, int max_rank, View* v) {
  EMK_REQUIRE(t != nullptr, EMK_E_NULL, "%s: NULL DLManagedTensor", name);
  const DLTensor& d = t->dl_tensor;
  EMK_REQUIRE(d.device.device_type == kDLCUDA || d.device.device_type == kDLCUDAManaged, EMK_E_DEVICE,
              "%s: tensor is on device_type %d, need a CUDA tensor (there is no CPU fallback)", name, d.device.device_type);
  EMK_REQUIRE(d.dtype.code == code && d.dtype.bits == bits && d.dtype.lanes == 1, EMK_E_DTYPE,
              "%s: dtype (code %d, %d bits) is not %s%d", name, d.dtype.code, d.dtype.bits, code == kDLFloat ? "float" : "int", bits);
  EMK_REQUIRE(d.ndim >= min_rank && d.ndim <= max_rank && d.ndim <= 4, EMK_E_SHAPE, "%s: rank %d outside [%d,%d]", name, d.ndim,
              min_rank, max_rank);
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess)
    EMK_REQUIRE(dev == d.device.device_id, EMK_E_DEVICE, "%s: tensor lives on cuda:%d but the current device is cuda:%d", name,
                d.device.device_id, dev);
  v->ndim = d.ndim;
  v->numel = 1;
  for (int i = 0; i < d.ndim; i++) {
    v->shape[i] = d.shape[i];
    v->numel *= d.shape[i];
  }
  if (d.strides != nullptr && v->numel > 0) {
    int64_t expect = 1;
    for (int i = d.ndim - 1; i >= 0; i--) {
      EMK_REQUIRE(d.shape[i] == 1 || d.strides[i] == expect, EMK_E_CONTIG, "%s: tensor is not C-contiguous (stride[%d]=%lld, expected %lld)",
                  name, i, (long long)d.strides[i], (long long)expect);
      expect *= d.shape[i];
    }
  }
  v->data = static_cast<char*>(d.data) + d.byte_offset;
  return EMK_OK;
}
#define VIEW(var, tensor, name, minr, maxr)                                   \
  View var;                                                                   \
  do {                                                                        \
    int rc_ = view_of(tensor, name, kDLFloat, 32, minr, maxr, &var);          \
    if (rc_) return rc_;                                                      \
  } while (0)
#define F(v) static_cast<float*>((v).data)

// python-style slice resolution for inputs[:, start:stop:step]; INT64_MIN means "None"
static int resolve_slice(int64_t n, int64_t start, int64_t stop, int64_t step, int64_t* first, int64_t* count, int64_t* stride) {
  const int64_t none = INT64_MIN;
  if (step == none) step = 1;
  EMK_REQUIRE(step > 0, EMK_E_UNSUPPORTED, "atom selection: only positive steps are supported (got %lld)", (long long)step);
  if (start == none) start = 0;
  if (stop == none) stop = n;
  if (start < 0) start += n;
  if (stop < 0) stop += n;
  if (start < 0) start = 0;
  if (start > n) start = n;
  if (stop < 0) stop = 0;
  if (stop > n) stop = n;
  *first = start;
  *count = stop > start ? (stop - start + step - 1) / step : 0;
  *stride = step;
  return EMK_OK;
}

static void split_geometry(int64_t n, int64_t* split, int64_t counts[4]) {
  const int64_t s = n / 2;
  *split = s;
  counts[0] = s + 2;          // left atoms  s+1 .. 0
  counts[1] = n - s + 1;      // right atoms s-1 .. n-1
  const int64_t nd = n - 3, m = nd / 2;
  counts[2] = (nd % 2 == 0) ? m : m + 1;
  counts[3] = nd - counts[2];
}

}  // namespace emk

using namespace emk;

extern "C" {

int emk_version(void) { return EMK_VERSION; }
const char* emk_last_error(void) { return last_error_buffer(); }
const char* emk_build_info(void) {
  static char info[128];
  snprintf(info, sizeof(info), "sm_100a;nvcc %d.%d;%s", __CUDACC_VER_MAJOR__, __CUDACC_VER_MINOR__, __DATE__);
  return info;
}

int emk_probe_fp32(double* lane_instr_per_s) { return fp32_probe_device(lane_instr_per_s); }

int emk_set_option(const char* name, int64_t value) {
  EMK_REQUIRE(name, EMK_E_NULL, "emk_set_option: NULL name");
  if (strcmp(name, "backmap_fwd6_min_batch") == 0) {
    set_fwd6_min_batch(value);
    return EMK_OK;
  }
  if (strcmp(name, "backmap_fwd6_warps") == 0) {
    EMK_REQUIRE(value == 0 || value == 8 || value == 12 || value == 14 || value == 16 || value == 18 || value == 20, EMK_E_ARG,
                "emk_set_option: backmap_fwd6_warps must be 0 (automatic), 8, 12, 14, 16, 18 or 20");
    set_fwd6_warps(value);
    return EMK_OK;
  }
  if (strcmp(name, "cost_small_d_max") == 0) {
    set_cost_small_d_max(value);
    return EMK_OK;
  }
  if (strcmp(name, "cost_small_tile_max_rows") == 0) {
    set_cost_small_tile_max_rows(value);
    return EMK_OK;
  }

  if (strcmp(name, "backmap_fwd6_f32_extent_nm") == 0) {
    EMK_REQUIRE(value >= 0 && value <= 64, EMK_E_ARG, "emk_set_option: backmap_fwd6_f32_extent_nm must be in [0, 64]");
    set_fwd6_f32_extent(value);
    return EMK_OK;
  }
  return fail(EMK_E_ARG, "emk_set_option: unknown option '%s'", name);
}
int emk_get_option(const char* name, int64_t* value) {
  EMK_REQUIRE(name && value, EMK_E_NULL, "emk_get_option: NULL argument");
  if (strcmp(name, "backmap_fwd6_min_batch") == 0) {
    *value = fwd6_min_batch();
    return EMK_OK;
  }
  if (strcmp(name, "backmap_fwd6_warps") == 0) {
    *value = fwd6_warps();
    return EMK_OK;
  }
  if (strcmp(name, "cost_small_d_max") == 0) {
    *value = cost_small_d_max();
    return EMK_OK;
  }
  if (strcmp(name, "cost_small_tile_max_rows") == 0) {
    *value = cost_small_tile_max_rows();
    return EMK_OK;
  }

  if (strcmp(name, "backmap_fwd6_f32_extent_nm") == 0) {
    *value = fwd6_f32_extent();
    return EMK_OK;
  }
  return fail(EMK_E_ARG, "emk_get_option: unknown option '%s'", name);
}

// ---- host-only index construction --------------------------------------------------------------------------
int64_t emk_triu_pair_count(int64_t n) { return n < 2 ? 0 : n * (n - 1) / 2; }

int emk_triu_pair_indices(int64_t n, int32_t* i_out, int32_t* j_out) {
  EMK_REQUIRE(n >= 0 && n < (1 << 30), EMK_E_ARG, "emk_triu_pair_indices: bad n=%lld", (long long)n);
  EMK_REQUIRE(n < 2 || (i_out && j_out), EMK_E_NULL, "emk_triu_pair_indices: NULL output");
  int64_t p = 0;
  for (int64_t i = 0; i < n; i++)
    for (int64_t j = i + 1; j < n; j++, p++) {
      i_out[p] = (int32_t)i;
      j_out[p] = (int32_t)j;
    }
  return EMK_OK;
}

int emk_backmap_split_counts(int64_t n_atoms, int64_t counts[4]) {
  EMK_REQUIRE(counts, EMK_E_NULL, "emk_backmap_split_counts: NULL output");
  EMK_REQUIRE(n_atoms >= 3, EMK_E_SHAPE, "emk_backmap_split_counts: need n_atoms >= 3, got %lld", (long long)n_atoms);
  int64_t s;
  split_geometry(n_atoms, &s, counts);
  return EMK_OK;
}

int emk_backmap_split_indices(int64_t n_atoms, int32_t* left_atoms, int32_t* right_atoms, int32_t* left_dihedrals,
                              int32_t* right_dihedrals) {
  EMK_REQUIRE(n_atoms >= 3, EMK_E_SHAPE, "emk_backmap_split_indices: need n_atoms >= 3, got %lld", (long long)n_atoms);
  EMK_REQUIRE(left_atoms && right_atoms && left_dihedrals && right_dihedrals, EMK_E_NULL, "emk_backmap_split_indices: NULL output");
  int64_t s, c[4];
  split_geometry(n_atoms, &s, c);
  for (int64_t m = 0; m < c[0]; m++) left_atoms[m] = (int32_t)(s + 1 - m);
  for (int64_t m = 0; m < c[1]; m++) right_atoms[m] = (int32_t)(s - 1 + m);
  for (int64_t i = 0; i < c[2]; i++) left_dihedrals[i] = (int32_t)(c[2] - 1 - i);
  for (int64_t i = 0; i < c[3]; i++) right_dihedrals[i] = (int32_t)(c[2] + i);
  return EMK_OK;
}

int64_t emk_pair_tile_count(int64_t n_rows) { return pair_tile_count(n_rows); }

int emk_pair_tile_decode(int64_t n_rows, int64_t tile, int64_t* tile_row, int64_t* tile_col) {
  EMK_REQUIRE(tile_row && tile_col, EMK_E_NULL, "emk_pair_tile_decode: NULL output");
  EMK_REQUIRE(tile >= 0 && tile < pair_tile_count(n_rows), EMK_E_ARG, "emk_pair_tile_decode: tile %lld outside [0,%lld)", (long long)tile,
              (long long)pair_tile_count(n_rows));
  tile_decode_host(tile, n_rows, tile_row, tile_col);
  return EMK_OK;
}

int emk_pair_tile_range(int64_t n_rows, int rank, int world, int64_t* begin, int64_t* end) {
  EMK_REQUIRE(begin && end, EMK_E_NULL, "emk_pair_tile_range: NULL output");
  EMK_REQUIRE(world >= 1 && rank >= 0 && rank < world, EMK_E_ARG, "emk_pair_tile_range: bad rank %d / world %d", rank, world);
  const int64_t total = pair_tile_count(n_rows);
  const int64_t base = total / world, rem = total % world;
  *begin = rank * base + (rank < rem ? rank : rem);
  *end = *begin + base + (rank < rem ? 1 : 0);
  return EMK_OK;
}

// ---- sigmoid cost ------------------------------------------------------------------------------------------------
int emk_sigmoid_cost(const float* high, int64_t n, int64_t d, const float* low, int64_t l, double periodicity, const float sig[6],
                     int64_t tile_begin, int64_t tile_end, double* loss, float* grad_low, uint32_t flags, void* stream) {
  return sigmoid_cost_device(high, n, d, low, l, periodicity, sig, tile_begin, tile_end, loss, grad_low, flags, as_stream(stream));
}

int emk_dl_sigmoid_cost(const DLManagedTensor* high, const DLManagedTensor* low, double periodicity, const float sig[6],
                        int64_t tile_begin, int64_t tile_end, DLManagedTensor* loss, DLManagedTensor* grad_low, uint32_t flags,
                        void* stream) {
  VIEW(h, high, "high", 2, 2);
  VIEW(z, low, "low", 2, 2);
  EMK_REQUIRE(h.shape[0] == z.shape[0], EMK_E_SHAPE, "emk_dl_sigmoid_cost: high has %lld rows, low has %lld", (long long)h.shape[0],
              (long long)z.shape[0]);
  View lv;
  int rc = view_of(loss, "loss", kDLFloat, 64, 0, 1, &lv);
  if (rc) return rc;
  EMK_REQUIRE(lv.numel == 1, EMK_E_SHAPE, "emk_dl_sigmoid_cost: loss must hold exactly one float64");
  float* g = nullptr;
  if (!(flags & EMK_COST_NO_GRAD)) {
    VIEW(gv, grad_low, "grad_low", 2, 2);
    EMK_REQUIRE(gv.shape[0] == z.shape[0] && gv.shape[1] == z.shape[1], EMK_E_SHAPE, "emk_dl_sigmoid_cost: grad_low shape differs from low");
    g = F(gv);
  }
  return sigmoid_cost_device(F(h), h.shape[0], h.shape[1], F(z), z.shape[1], periodicity, sig, tile_begin, tile_end,
                             static_cast<double*>(lv.data), g, flags, as_stream(stream));
}

int emk_sigmoid_cost_host(const float* high_host, int64_t n, int64_t d, const float* low_host, int64_t l, double periodicity,
                          const float sig[6], double* loss_host, float* grad_low_host) {
  EMK_REQUIRE(high_host && low_host && sig && loss_host, EMK_E_NULL, "emk_sigmoid_cost_host: NULL pointer argument");
  EMK_REQUIRE(n >= 1 && d >= 1 && l >= 1, EMK_E_SHAPE, "emk_sigmoid_cost_host: bad shape");
  cudaStream_t st;
  EMK_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  float *dh = nullptr, *dz = nullptr, *dg = nullptr;
  double* dl = nullptr;
  int rc = EMK_OK;
  auto cleanup = [&]() {
    if (dh) cudaFreeAsync(dh, st);
    if (dz) cudaFreeAsync(dz, st);
    if (dg) cudaFreeAsync(dg, st);
    if (dl) cudaFreeAsync(dl, st);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
  };
#define HOSTCK(call)                                                                          \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      cleanup();                                                                              \
      return fail((int)e_, "%s failed: %s", #call, cudaGetErrorString(e_));                   \
    }                                                                                         \
  } while (0)
  HOSTCK(cudaMallocAsync(&dh, (size_t)n * d * sizeof(float), st));
  HOSTCK(cudaMallocAsync(&dz, (size_t)n * l * sizeof(float), st));
  HOSTCK(cudaMallocAsync(&dg, (size_t)n * l * sizeof(float), st));
  HOSTCK(cudaMallocAsync(&dl, sizeof(double), st));
  HOSTCK(cudaMemcpyAsync(dh, high_host, (size_t)n * d * sizeof(float), cudaMemcpyHostToDevice, st));
  HOSTCK(cudaMemcpyAsync(dz, low_host, (size_t)n * l * sizeof(float), cudaMemcpyHostToDevice, st));
  rc = sigmoid_cost_device(dh, n, d, dz, l, periodicity, sig, 0, pair_tile_count(n), dl, dg,
                           EMK_COST_ZERO_OUTPUTS | (grad_low_host ? 0u : EMK_COST_NO_GRAD), st);
  if (rc == EMK_OK) {
    HOSTCK(cudaMemcpyAsync(loss_host, dl, sizeof(double), cudaMemcpyDeviceToHost, st));
    if (grad_low_host) HOSTCK(cudaMemcpyAsync(grad_low_host, dg, (size_t)n * l * sizeof(float), cudaMemcpyDeviceToHost, st));
    HOSTCK(cudaStreamSynchronize(st));
  }
  cleanup();
  return rc;
}

// ---- cartesian_distance_loss straight from coordinates ------------------------------------------------------
// The high-d side of cartesian_distance_loss is the matrix of flat pair distances of the selected atoms, one row per frame
// (models/models.py:837-839: inp_pair = PairwiseDistances(...)(inp_cartesians); :2419-2422: loss(inp_pair, latent)).  Here the
// caller hands over the COORDINATES: the matrix is formed into stream-ordered scratch (16-byte row pitch, zero-filled tail, so
// the pair-tile kernel reads it through TMA without the re-padding pass that D % 4 != 0 inputs otherwise need -- 4 950 and
// 44 850 both are), consumed once and released; it never becomes a tensor of the host framework (no activation kept for the
// backward pass, the largest one of the ADC step) and at training sizes (1 024 x 4 950: 20 MB) it lives in L2 between the two
// kernels.  Forming the pair distances inside the pair-tile kernel instead was considered and rejected: a 128 x 64 tile would
// recompute 192 frames' distances (13x redundantly over the 72 tiles of a 1 024-frame batch), ~25 % more issue slots on a
// kernel that is bound by them, and 3x the L2 -> SM traffic of loading the finished rows.
int emk_cartesian_distance_cost(const float* xyz, int64_t b, int64_t n_atoms, int64_t first, int64_t count, int64_t step, const float* low,
                                int64_t l, const float sig[6], int64_t tile_begin, int64_t tile_end, double* loss, float* grad_low,
                                uint32_t flags, void* stream) {
  EMK_REQUIRE(sig && loss && (b == 0 || (xyz && low)), EMK_E_NULL, "emk_cartesian_distance_cost: NULL pointer argument");
  EMK_REQUIRE(b >= 0 && n_atoms >= 1 && count >= 2 && step >= 1 && first >= 0 && first + (count - 1) * step < n_atoms, EMK_E_SHAPE,
              "emk_cartesian_distance_cost: atom selection (first %lld, count %lld, step %lld) needs >= 2 of %lld atoms", (long long)first,
              (long long)count, (long long)step, (long long)n_atoms);
  cudaStream_t st = as_stream(stream);
  const int64_t per = count * (count - 1) / 2;
  const bool pitched = count <= pairwise_warp_max_atoms();
  const int64_t pitch = pitched ? (per + 3) / 4 * 4 : per;
  if (b == 0) return sigmoid_cost_device(xyz, 0, pitch, low, l, INFINITY, sig, tile_begin, tile_end, loss, grad_low, flags, st);
  float* pairs = nullptr;
  int rc = scratch_alloc(reinterpret_cast<void**>(&pairs), (size_t)(b * pitch) * sizeof(float), st);
  if (rc) return rc;
  rc = pitched ? pairwise_flat3_warp_device(xyz + 3 * first, b, count, 3 * n_atoms, 3 * step, 0, pairs, pitch, st)
               : pairwise_small_device(xyz + 3 * first, b, count, 3, 3 * n_atoms, 3 * step, 0, 1, pairs, st);
  if (rc == EMK_OK) rc = sigmoid_cost_device(pairs, b, pitch, low, l, INFINITY, sig, tile_begin, tile_end, loss, grad_low, flags, st);
  cudaFreeAsync(pairs, st);
  return rc;
}

int emk_dl_cartesian_distance_cost(const DLManagedTensor* xyz, int64_t start, int64_t stop, int64_t step, const DLManagedTensor* low,
                                   const float sig[6], int64_t tile_begin, int64_t tile_end, DLManagedTensor* loss, DLManagedTensor* grad_low,
                                   uint32_t flags, void* stream) {
  VIEW(xv, xyz, "cartesians", 3, 3);
  VIEW(z, low, "low", 2, 2);
  EMK_REQUIRE(xv.shape[2] == 3, EMK_E_SHAPE, "emk_dl_cartesian_distance_cost: cartesians must be (b, n_atoms, 3)");
  EMK_REQUIRE(xv.shape[0] == z.shape[0], EMK_E_SHAPE, "emk_dl_cartesian_distance_cost: %lld frames, low has %lld rows", (long long)xv.shape[0],
              (long long)z.shape[0]);
  int64_t first, cnt, st;
  int rc = resolve_slice(xv.shape[1], start, stop, step, &first, &cnt, &st);
  if (rc) return rc;
  View lv;
  rc = view_of(loss, "loss", kDLFloat, 64, 0, 1, &lv);
  if (rc) return rc;
  EMK_REQUIRE(lv.numel == 1, EMK_E_SHAPE, "emk_dl_cartesian_distance_cost: loss must hold exactly one float64");
  float* g = nullptr;
  if (!(flags & EMK_COST_NO_GRAD)) {
    VIEW(gv, grad_low, "grad_low", 2, 2);
    EMK_REQUIRE(gv.shape[0] == z.shape[0] && gv.shape[1] == z.shape[1], EMK_E_SHAPE, "emk_dl_cartesian_distance_cost: grad_low shape differs from low");
    g = F(gv);
  }
  return emk_cartesian_distance_cost(F(xv), xv.shape[0], xv.shape[1], first, cnt, st, F(z), z.shape[1], sig, tile_begin, tile_end,
                                     static_cast<double*>(lv.data), g, flags, stream);
}

// ---- distance matrices -----------------------------------------------------------------------------------------------
int emk_pairwise_dist_periodic(const float* x, int64_t n, int64_t d, double periodicity, float* out, void* stream) {
  return dist_matrix_device(x, n, d, periodicity, true, 0, out, as_stream(stream));
}
int emk_dl_pairwise_dist_periodic(const DLManagedTensor* x, double periodicity, DLManagedTensor* out, void* stream) {
  VIEW(xv, x, "positions", 2, 2);
  VIEW(ov, out, "out", 2, 2);
  EMK_REQUIRE(ov.shape[0] == xv.shape[0] && ov.shape[1] == xv.shape[0], EMK_E_SHAPE, "emk_dl_pairwise_dist_periodic: out must be (n,n)");
  return dist_matrix_device(F(xv), xv.shape[0], xv.shape[1], periodicity, true, 0, F(ov), as_stream(stream));
}

int emk_pairwise_dist_periodic_bwd(const float* x, int64_t n, int64_t d, double periodicity, const float* dist, const float* grad_out,
                                   float* grad_x, void* stream) {
  return pairwise_periodic_bwd_device(x, n, d, periodicity, dist, grad_out, grad_x, as_stream(stream));
}
int emk_dl_pairwise_dist_periodic_bwd(const DLManagedTensor* x, double periodicity, const DLManagedTensor* dist,
                                      const DLManagedTensor* grad_out, DLManagedTensor* grad_x, void* stream) {
  VIEW(xv, x, "positions", 2, 2);
  VIEW(dv, dist, "dist", 2, 2);
  VIEW(gv, grad_out, "grad_out", 2, 2);
  VIEW(ov, grad_x, "grad_x", 2, 2);
  const int64_t n = xv.shape[0];
  EMK_REQUIRE(dv.shape[0] == n && dv.shape[1] == n && gv.shape[0] == n && gv.shape[1] == n, EMK_E_SHAPE,
              "emk_dl_pairwise_dist_periodic_bwd: dist and grad_out must be (n,n)");
  EMK_REQUIRE(ov.shape[0] == n && ov.shape[1] == xv.shape[1], EMK_E_SHAPE, "emk_dl_pairwise_dist_periodic_bwd: grad_x must be (n,d)");
  return pairwise_periodic_bwd_device(F(xv), n, xv.shape[1], periodicity, F(dv), F(gv), F(ov), as_stream(stream));
}

int emk_pairwise_dist(const float* x, int64_t b, int64_t n, int64_t d, int64_t batch_stride, int64_t row_stride, int squared,
                      int flat, float* out, void* stream) {
  EMK_REQUIRE((x && out) || b == 0 || n < 2, EMK_E_NULL, "emk_pairwise_dist: NULL pointer argument");
  EMK_REQUIRE(b >= 0 && n >= 0 && d >= 1, EMK_E_SHAPE, "emk_pairwise_dist: bad shape b=%lld n=%lld d=%lld", (long long)b, (long long)n, (long long)d);
  // one big contiguous rank-2 problem with a wide feature axis goes through the TMA pair-tile kernel
  if (b == 1 && !flat && row_stride == d && d >= 16 && n >= 256)
    return dist_matrix_device(x, n, d, INFINITY, false, squared, out, as_stream(stream));
  return pairwise_small_device(x, b, n, d, batch_stride, row_stride, squared, flat, out, as_stream(stream));
}
int emk_pairwise_dist_bwd(const float* x, int64_t b, int64_t n, int64_t d, int64_t batch_stride, int64_t row_stride, int squared,
                          int flat, const float* grad_out, float* grad_x, void* stream) {
  EMK_REQUIRE((x && grad_out && grad_x) || b == 0 || n == 0 || (x && grad_x && n < 2), EMK_E_NULL, "emk_pairwise_dist_bwd: NULL pointer argument");
  EMK_REQUIRE(b >= 0 && n >= 0 && d >= 1, EMK_E_SHAPE, "emk_pairwise_dist_bwd: bad shape");
  return pairwise_small_bwd_device(x, b, n, d, batch_stride, row_stride, squared, flat, grad_out, grad_x, as_stream(stream));
}

static int pw_geometry(const View& xv, int64_t start, int64_t stop, int64_t step, int64_t* b, int64_t* n, int64_t* d, int64_t* bs,
                       int64_t* rs, int64_t* first) {
  if (xv.ndim == 2) {
    *b = 1;
    *d = xv.shape[1];
    *bs = xv.numel;
    int64_t cnt, st;
    int rc = resolve_slice(xv.shape[0], start, stop, step, first, &cnt, &st);
    if (rc) return rc;
    *n = cnt;
    *rs = st * xv.shape[1];
    *first *= xv.shape[1];
  } else {
    *b = xv.shape[0];
    *d = xv.shape[2];
    *bs = xv.shape[1] * xv.shape[2];
    int64_t cnt, st;
    int rc = resolve_slice(xv.shape[1], start, stop, step, first, &cnt, &st);
    if (rc) return rc;
    *n = cnt;
    *rs = st * xv.shape[2];
    *first *= xv.shape[2];
  }
  return EMK_OK;
}

int emk_dl_pairwise_dist(const DLManagedTensor* x, int64_t start, int64_t stop, int64_t step, int squared, int flat,
                         DLManagedTensor* out, void* stream) {
  VIEW(xv, x, "positions", 2, 3);
  int64_t b, n, d, bs, rs, first;
  int rc = pw_geometry(xv, start, stop, step, &b, &n, &d, &bs, &rs, &first);
  if (rc) return rc;
  VIEW(ov, out, "out", 2, 3);
  const int64_t want = b * (flat ? n * (n - 1) / 2 : n * n);
  EMK_REQUIRE(ov.numel == want, EMK_E_SHAPE, "emk_dl_pairwise_dist: out has %lld elements, expected %lld", (long long)ov.numel, (long long)want);
  return emk_pairwise_dist(F(xv) + first, b, n, d, bs, rs, squared, flat, F(ov), stream);
}
int emk_dl_pairwise_dist_bwd(const DLManagedTensor* x, int64_t start, int64_t stop, int64_t step, int squared, int flat,
                             const DLManagedTensor* grad_out, DLManagedTensor* grad_x, void* stream) {
  VIEW(xv, x, "positions", 2, 3);
  int64_t b, n, d, bs, rs, first;
  int rc = pw_geometry(xv, start, stop, step, &b, &n, &d, &bs, &rs, &first);
  if (rc) return rc;
  VIEW(gv, grad_out, "grad_out", 2, 3);
  VIEW(gx, grad_x, "grad_x", 2, 3);
  const int64_t want = b * (flat ? n * (n - 1) / 2 : n * n);
  EMK_REQUIRE(gv.numel == want, EMK_E_SHAPE, "emk_dl_pairwise_dist_bwd: grad_out has %lld elements, expected %lld", (long long)gv.numel, (long long)want);
  EMK_REQUIRE(gx.numel == xv.numel, EMK_E_SHAPE, "emk_dl_pairwise_dist_bwd: grad_x shape differs from positions");
  return emk_pairwise_dist_bwd(F(xv) + first, b, n, d, bs, rs, squared, flat, F(gv), F(gx) + first, stream);
}

// ---- fused Cartesian branch: pairwise distances of selected atoms + cartesian loss (+ clash count), forward and gradient ---------
int emk_cartesian_pair_loss(const float* xyz, int64_t b, int64_t n_atoms, int64_t first, int64_t count, int64_t step, const float* target,
                            int target_is_xyz, int variant, float clash_distance, double* loss_sum, float* grad_xyz, int64_t* clashes,
                            void* stream) {
  EMK_REQUIRE(n_atoms >= 1 && n_atoms < (1 << 24) && count >= 0 && count < (1 << 24) && step >= 1 && step < (1 << 24) && first >= 0 &&
                  first < (1 << 24),
              EMK_E_SHAPE, "emk_cartesian_pair_loss: bad selection / atom count");
  CartLossParams p{xyz, target, b, (int)n_atoms, (int)first, (int)count, (int)step, target_is_xyz, variant,
                   clash_distance > 0.f ? clash_distance * clash_distance : -1.f, loss_sum, grad_xyz, reinterpret_cast<long long*>(clashes)};
  return cart_pair_loss_device(p, as_stream(stream));
}
int emk_dl_cartesian_pair_loss(const DLManagedTensor* xyz, int64_t start, int64_t stop, int64_t step, const DLManagedTensor* target, int variant,
                               float clash_distance, DLManagedTensor* loss_sum, DLManagedTensor* grad_xyz, DLManagedTensor* clashes,
                               void* stream) {
  VIEW(xv, xyz, "xyz", 3, 3);
  EMK_REQUIRE(xv.shape[2] == 3, EMK_E_SHAPE, "emk_dl_cartesian_pair_loss: xyz must be (b, n_atoms, 3)");
  const int64_t b = xv.shape[0], n = xv.shape[1];
  int64_t first, cnt, st;
  int rc = resolve_slice(n, start, stop, step, &first, &cnt, &st);
  if (rc) return rc;
  VIEW(tv, target, "target", 2, 3);
  int is_xyz;
  if (tv.ndim == 3) {
    EMK_REQUIRE(tv.shape[0] == b && tv.shape[1] == n && tv.shape[2] == 3, EMK_E_SHAPE, "emk_dl_cartesian_pair_loss: input coordinates must be (b, n_atoms, 3)");
    is_xyz = 1;
  } else {
    EMK_REQUIRE(tv.shape[0] == b && tv.shape[1] == cnt * (cnt - 1) / 2, EMK_E_SHAPE,
                "emk_dl_cartesian_pair_loss: input pair distances must be (b, %lld) for %lld selected atoms, got (%lld, %lld)",
                (long long)(cnt * (cnt - 1) / 2), (long long)cnt, (long long)tv.shape[0], (long long)tv.shape[1]);
    is_xyz = 0;
  }
  View lv;
  rc = view_of(loss_sum, "loss_sum", kDLFloat, 64, 0, 1, &lv);
  if (rc) return rc;
  EMK_REQUIRE(lv.numel == 1, EMK_E_SHAPE, "emk_dl_cartesian_pair_loss: loss_sum must hold exactly one float64");
  float* g = nullptr;
  if (grad_xyz) {
    VIEW(gv, grad_xyz, "grad_xyz", 3, 3);
    EMK_REQUIRE(gv.numel == xv.numel, EMK_E_SHAPE, "emk_dl_cartesian_pair_loss: grad_xyz shape differs from xyz");
    g = F(gv);
  }
  int64_t* cl = nullptr;
  if (clashes) {
    View cv;
    rc = view_of(clashes, "clashes", kDLInt, 64, 1, 1, &cv);
    if (rc) return rc;
    EMK_REQUIRE(cv.numel == b, EMK_E_SHAPE, "emk_dl_cartesian_pair_loss: clashes must be (b) int64");
    cl = static_cast<int64_t*>(cv.data);
  }
  return emk_cartesian_pair_loss(F(xv), b, n, first, cnt, st, F(tv), is_xyz, variant, clash_distance, static_cast<double*>(lv.data), g, cl, stream);
}

// ---- elementwise ---------------------------------------------------------------------------------------------------------
int emk_periodic_distance(const float* a, const float* b, int64_t count, double periodicity, float* out, void* stream) {
  return periodic_distance_device(a, b, count, periodicity, out, as_stream(stream));
}
int emk_periodic_distance_bwd(const float* a, const float* b, int64_t count, double periodicity, const float* grad_out, float* grad_a,
                              float* grad_b, void* stream) {
  return periodic_distance_bwd_device(a, b, count, periodicity, grad_out, grad_a, grad_b, as_stream(stream));
}
int emk_dl_periodic_distance(const DLManagedTensor* a, const DLManagedTensor* b, double periodicity, DLManagedTensor* out, void* stream) {
  VIEW(av, a, "a", 0, 4);
  VIEW(bv, b, "b", 0, 4);
  VIEW(ov, out, "out", 0, 4);
  EMK_REQUIRE(av.numel == bv.numel && av.numel == ov.numel, EMK_E_SHAPE, "emk_dl_periodic_distance: operands must have equal element counts (broadcast first)");
  return periodic_distance_device(F(av), F(bv), av.numel, periodicity, F(ov), as_stream(stream));
}
int emk_dl_periodic_distance_bwd(const DLManagedTensor* a, const DLManagedTensor* b, double periodicity, const DLManagedTensor* grad_out,
                                 DLManagedTensor* grad_a, DLManagedTensor* grad_b, void* stream) {
  VIEW(av, a, "a", 0, 4);
  VIEW(bv, b, "b", 0, 4);
  VIEW(gv, grad_out, "grad_out", 0, 4);
  EMK_REQUIRE(av.numel == bv.numel && av.numel == gv.numel, EMK_E_SHAPE, "emk_dl_periodic_distance_bwd: element counts differ");
  float *ga = nullptr, *gb = nullptr;
  if (grad_a) {
    VIEW(t, grad_a, "grad_a", 0, 4);
    EMK_REQUIRE(t.numel == av.numel, EMK_E_SHAPE, "emk_dl_periodic_distance_bwd: grad_a element count differs");
    ga = F(t);
  }
  if (grad_b) {
    VIEW(t, grad_b, "grad_b", 0, 4);
    EMK_REQUIRE(t.numel == av.numel, EMK_E_SHAPE, "emk_dl_periodic_distance_bwd: grad_b element count differs");
    gb = F(t);
  }
  return periodic_distance_bwd_device(F(av), F(bv), av.numel, periodicity, F(gv), ga, gb, as_stream(stream));
}

int emk_sigmoid(const float* r, int64_t count, float sig, float a, float b, float* out, void* stream) {
  return sigmoid_device(r, count, sig, a, b, out, as_stream(stream));
}
int emk_sigmoid_bwd(const float* r, int64_t count, float sig, float a, float b, const float* grad_out, float* grad_r, void* stream) {
  return sigmoid_bwd_device(r, count, sig, a, b, grad_out, grad_r, as_stream(stream));
}
int emk_dl_sigmoid(const DLManagedTensor* r, float sig, float a, float b, DLManagedTensor* out, void* stream) {
  VIEW(rv, r, "r", 0, 4);
  VIEW(ov, out, "out", 0, 4);
  EMK_REQUIRE(rv.numel == ov.numel, EMK_E_SHAPE, "emk_dl_sigmoid: element counts differ");
  return sigmoid_device(F(rv), rv.numel, sig, a, b, F(ov), as_stream(stream));
}
int emk_dl_sigmoid_bwd(const DLManagedTensor* r, float sig, float a, float b, const DLManagedTensor* grad_out, DLManagedTensor* grad_r,
                       void* stream) {
  VIEW(rv, r, "r", 0, 4);
  VIEW(gv, grad_out, "grad_out", 0, 4);
  VIEW(ov, grad_r, "grad_r", 0, 4);
  EMK_REQUIRE(rv.numel == ov.numel && rv.numel == gv.numel, EMK_E_SHAPE, "emk_dl_sigmoid_bwd: element counts differ");
  return sigmoid_bwd_device(F(rv), rv.numel, sig, a, b, F(gv), F(ov), as_stream(stream));
}

int emk_periodic_input(const float* x, int64_t rows, int64_t d, double periodicity, float* out, void* stream) {
  return periodic_input_device(x, rows, d, periodicity, out, as_stream(stream));
}
int emk_periodic_input_bwd(const float* x, int64_t rows, int64_t d, double periodicity, const float* grad_out, float* grad_x, void* stream) {
  return periodic_input_bwd_device(x, rows, d, periodicity, grad_out, grad_x, as_stream(stream));
}
int emk_dl_periodic_input(const DLManagedTensor* x, double periodicity, DLManagedTensor* out, void* stream) {
  VIEW(xv, x, "x", 2, 2);
  VIEW(ov, out, "out", 2, 2);
  EMK_REQUIRE(ov.shape[0] == xv.shape[0] && ov.shape[1] == 2 * xv.shape[1], EMK_E_SHAPE, "emk_dl_periodic_input: out must be (rows, 2d)");
  return periodic_input_device(F(xv), xv.shape[0], xv.shape[1], periodicity, F(ov), as_stream(stream));
}
int emk_dl_periodic_input_bwd(const DLManagedTensor* x, double periodicity, const DLManagedTensor* grad_out, DLManagedTensor* grad_x,
                              void* stream) {
  VIEW(xv, x, "x", 2, 2);
  VIEW(gv, grad_out, "grad_out", 2, 2);
  VIEW(ov, grad_x, "grad_x", 2, 2);
  EMK_REQUIRE(gv.shape[0] == xv.shape[0] && gv.shape[1] == 2 * xv.shape[1] && ov.numel == xv.numel, EMK_E_SHAPE,
              "emk_dl_periodic_input_bwd: shape mismatch");
  return periodic_input_bwd_device(F(xv), xv.shape[0], xv.shape[1], periodicity, F(gv), F(ov), as_stream(stream));
}

int emk_rotation_matrix(const float* axis, const float* angle, int64_t b, float* out, void* stream) {
  return rotation_matrix_device(axis, angle, b, out, as_stream(stream));
}
int emk_dl_rotation_matrix(const DLManagedTensor* axis, const DLManagedTensor* angle, DLManagedTensor* out, void* stream) {
  VIEW(av, axis, "axis_unit_vec", 2, 2);
  VIEW(gv, angle, "angle", 1, 1);
  VIEW(ov, out, "out", 3, 3);
  EMK_REQUIRE(av.shape[1] == 3 && gv.shape[0] == av.shape[0] && ov.shape[0] == av.shape[0] && ov.shape[1] == 3 && ov.shape[2] == 3,
              EMK_E_SHAPE, "emk_dl_rotation_matrix: need axis (b,3), angle (b), out (b,3,3)");
  return rotation_matrix_device(F(av), F(gv), av.shape[0], F(ov), as_stream(stream));
}

int emk_column_mean(const float* x, int64_t rows, int64_t cols, float* out, void* stream) {
  return column_mean_device(x, rows, cols, out, as_stream(stream));
}
int emk_dl_column_mean(const DLManagedTensor* x, DLManagedTensor* out, void* stream) {
  VIEW(xv, x, "x", 2, 2);
  VIEW(ov, out, "out", 1, 2);
  EMK_REQUIRE(ov.numel == xv.shape[1], EMK_E_SHAPE, "emk_dl_column_mean: out must hold one value per column");
  return column_mean_device(F(xv), xv.shape[0], xv.shape[1], F(ov), as_stream(stream));
}

// ---- back-mapping -------------------------------------------------------------------------------------------------------------
int emk_backmap(const float* lengths, int64_t lengths_batch_stride, const float* angles, const float* dihedrals, int64_t b,
                int64_t n_atoms, float* xyz, void* stream) {
  return backmap_fwd_device(lengths, lengths_batch_stride, angles, dihedrals, b, n_atoms, xyz, as_stream(stream));
}

int emk_backmap_bwd(const float* lengths, int64_t lengths_batch_stride, const float* angles, const float* xyz, const float* grad_xyz,
                    int64_t b, int64_t n_atoms, float* grad_angles, float* grad_dihedrals, float* grad_lengths, void* stream) {
  EMK_REQUIRE(b == 0 || (lengths && angles && xyz && grad_xyz), EMK_E_NULL, "emk_backmap_bwd: NULL pointer argument");
  EMK_REQUIRE(n_atoms >= 4 && n_atoms < (1 << 20), EMK_E_SHAPE, "emk_backmap_bwd: need 4 <= n_atoms < 2^20");
  EMK_REQUIRE(lengths_batch_stride == 0 || lengths_batch_stride == n_atoms - 1, EMK_E_ARG, "emk_backmap_bwd: lengths_batch_stride must be 0 or n_atoms-1");
  BwdParams p{lengths, lengths_batch_stride, angles, xyz, grad_xyz, b, (int)n_atoms, (int)(n_atoms / 2), (int)(n_atoms / 2 - 1), 0,
              grad_angles, grad_dihedrals, grad_lengths};
  return backmap_bwd_device(p, as_stream(stream));
}

static int lengths_view(const DLManagedTensor* lengths, int64_t b, int64_t n, View* lv, int64_t* stride) {
  int rc = view_of(lengths, "lengths", kDLFloat, 32, 1, 2, lv);
  if (rc) return rc;
  if (lv->numel == n - 1) *stride = 0;
  else if (lv->numel == b * (n - 1)) *stride = n - 1;
  else return fail(EMK_E_SHAPE, "lengths must hold n_atoms-1 = %lld values, or one such row per frame (got %lld)", (long long)(n - 1), (long long)lv->numel);
  return EMK_OK;
}

int emk_dl_backmap(const DLManagedTensor* lengths, const DLManagedTensor* angles, const DLManagedTensor* dihedrals, DLManagedTensor* xyz,
                   void* stream) {
  VIEW(av, angles, "angles", 2, 2);
  VIEW(dv, dihedrals, "dihedrals", 2, 2);
  VIEW(ov, xyz, "xyz", 3, 3);
  const int64_t b = av.shape[0], n = av.shape[1] + 2;
  EMK_REQUIRE(dv.shape[0] == b && dv.shape[1] == n - 3, EMK_E_SHAPE, "emk_dl_backmap: dihedrals must be (b, n_atoms-3) = (%lld,%lld)", (long long)b, (long long)(n - 3));
  EMK_REQUIRE(ov.shape[0] == b && ov.shape[1] == n && ov.shape[2] == 3, EMK_E_SHAPE, "emk_dl_backmap: xyz must be (b, n_atoms, 3)");
  View lv;
  int64_t ls;
  int rc = lengths_view(lengths, b, n, &lv, &ls);
  if (rc) return rc;
  return backmap_fwd_device(F(lv), ls, F(av), F(dv), b, n, F(ov), as_stream(stream));
}

int emk_dl_backmap_bwd(const DLManagedTensor* lengths, const DLManagedTensor* angles, const DLManagedTensor* xyz,
                       const DLManagedTensor* grad_xyz, DLManagedTensor* grad_angles, DLManagedTensor* grad_dihedrals,
                       DLManagedTensor* grad_lengths, void* stream) {
  VIEW(av, angles, "angles", 2, 2);
  VIEW(xv, xyz, "xyz", 3, 3);
  VIEW(gv, grad_xyz, "grad_xyz", 3, 3);
  const int64_t b = av.shape[0], n = av.shape[1] + 2;
  EMK_REQUIRE(xv.shape[0] == b && xv.shape[1] == n && xv.shape[2] == 3 && gv.numel == xv.numel, EMK_E_SHAPE, "emk_dl_backmap_bwd: xyz / grad_xyz must be (b, n_atoms, 3)");
  View lv;
  int64_t ls;
  int rc = lengths_view(lengths, b, n, &lv, &ls);
  if (rc) return rc;
  float *ga = nullptr, *gd = nullptr, *gl = nullptr;
  if (grad_angles) { VIEW(t, grad_angles, "grad_angles", 2, 2); EMK_REQUIRE(t.numel == av.numel, EMK_E_SHAPE, "grad_angles shape"); ga = F(t); }
  if (grad_dihedrals) { VIEW(t, grad_dihedrals, "grad_dihedrals", 2, 2); EMK_REQUIRE(t.numel == b * (n - 3), EMK_E_SHAPE, "grad_dihedrals shape"); gd = F(t); }
  if (grad_lengths) { VIEW(t, grad_lengths, "grad_lengths", 2, 2); EMK_REQUIRE(t.numel == b * (n - 1), EMK_E_SHAPE, "grad_lengths must be (b, n_atoms-1)"); gl = F(t); }
  return emk_backmap_bwd(F(lv), ls, F(av), F(xv), F(gv), b, n, ga, gd, gl, stream);
}

int emk_backmap_host(const float* lengths_host, const float* angles_host, const float* dihedrals_host, int64_t b, int64_t n_atoms,
                     float* xyz_host) {
  EMK_REQUIRE(lengths_host && angles_host && dihedrals_host && xyz_host, EMK_E_NULL, "emk_backmap_host: NULL pointer argument");
  EMK_REQUIRE(b >= 1 && n_atoms >= 4, EMK_E_SHAPE, "emk_backmap_host: bad shape");
  const int64_t n = n_atoms;
  float *dl = nullptr, *da = nullptr, *dd = nullptr, *dx = nullptr;
  cudaStream_t st;
  EMK_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  auto cleanup = [&]() {
    if (dl) cudaFreeAsync(dl, st);
    if (da) cudaFreeAsync(da, st);
    if (dd) cudaFreeAsync(dd, st);
    if (dx) cudaFreeAsync(dx, st);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
  };
  HOSTCK(cudaMallocAsync(&dl, (size_t)(n - 1) * sizeof(float), st));
  HOSTCK(cudaMallocAsync(&da, (size_t)b * (n - 2) * sizeof(float), st));
  HOSTCK(cudaMallocAsync(&dd, (size_t)b * (n - 3) * sizeof(float), st));
  HOSTCK(cudaMallocAsync(&dx, (size_t)b * n * 3 * sizeof(float), st));
  HOSTCK(cudaMemcpyAsync(dl, lengths_host, (size_t)(n - 1) * sizeof(float), cudaMemcpyHostToDevice, st));
  HOSTCK(cudaMemcpyAsync(da, angles_host, (size_t)b * (n - 2) * sizeof(float), cudaMemcpyHostToDevice, st));
  HOSTCK(cudaMemcpyAsync(dd, dihedrals_host, (size_t)b * (n - 3) * sizeof(float), cudaMemcpyHostToDevice, st));
  int rc = backmap_fwd_device(dl, 0, da, dd, b, n, dx, st);
  if (rc == EMK_OK) {
    HOSTCK(cudaMemcpyAsync(xyz_host, dx, (size_t)b * n * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    HOSTCK(cudaStreamSynchronize(st));
  }
  cleanup();
  return rc;
}

int emk_chain_in_plane(const float* lengths, int64_t lengths_batch_stride, const float* angles, int64_t b, int64_t n_atoms, float* xyz,
                       void* stream) {
  return chain_in_plane_device(lengths, lengths_batch_stride, angles, b, n_atoms, xyz, as_stream(stream));
}
int emk_chain_in_plane_bwd(const float* lengths, int64_t lengths_batch_stride, const float* angles, const float* grad_xyz, int64_t b,
                           int64_t n_atoms, float* grad_angles, float* grad_lengths, void* stream) {
  EMK_REQUIRE(b == 0 || (lengths && angles && grad_xyz), EMK_E_NULL, "emk_chain_in_plane_bwd: NULL pointer argument");
  EMK_REQUIRE(n_atoms >= 3 && n_atoms < (1 << 20), EMK_E_SHAPE, "emk_chain_in_plane_bwd: need 3 <= n_atoms < 2^20");
  EMK_REQUIRE(lengths_batch_stride == 0 || lengths_batch_stride == n_atoms - 1, EMK_E_ARG, "emk_chain_in_plane_bwd: lengths_batch_stride must be 0 or n_atoms-1");
  BwdParams p{lengths, lengths_batch_stride, angles, nullptr, grad_xyz, b, (int)n_atoms, 0, 0, 1, grad_angles, nullptr, grad_lengths};
  return backmap_bwd_device(p, as_stream(stream));
}
int emk_dl_chain_in_plane(const DLManagedTensor* lengths, const DLManagedTensor* angles, DLManagedTensor* xyz, void* stream) {
  VIEW(av, angles, "angles", 2, 2);
  VIEW(ov, xyz, "xyz", 3, 3);
  const int64_t b = av.shape[0], n = av.shape[1] + 2;
  EMK_REQUIRE(ov.shape[0] == b && ov.shape[1] == n && ov.shape[2] == 3, EMK_E_SHAPE, "emk_dl_chain_in_plane: xyz must be (b, n_atoms, 3)");
  View lv;
  int64_t ls;
  int rc = lengths_view(lengths, b, n, &lv, &ls);
  if (rc) return rc;
  return chain_in_plane_device(F(lv), ls, F(av), b, n, F(ov), as_stream(stream));
}
int emk_dl_chain_in_plane_bwd(const DLManagedTensor* lengths, const DLManagedTensor* angles, const DLManagedTensor* grad_xyz,
                              DLManagedTensor* grad_angles, DLManagedTensor* grad_lengths, void* stream) {
  VIEW(av, angles, "angles", 2, 2);
  VIEW(gv, grad_xyz, "grad_xyz", 3, 3);
  const int64_t b = av.shape[0], n = av.shape[1] + 2;
  EMK_REQUIRE(gv.shape[0] == b && gv.shape[1] == n && gv.shape[2] == 3, EMK_E_SHAPE, "emk_dl_chain_in_plane_bwd: grad_xyz must be (b, n_atoms, 3)");
  View lv;
  int64_t ls;
  int rc = lengths_view(lengths, b, n, &lv, &ls);
  if (rc) return rc;
  float *ga = nullptr, *gl = nullptr;
  if (grad_angles) { VIEW(t, grad_angles, "grad_angles", 2, 2); EMK_REQUIRE(t.numel == av.numel, EMK_E_SHAPE, "grad_angles shape"); ga = F(t); }
  if (grad_lengths) { VIEW(t, grad_lengths, "grad_lengths", 2, 2); EMK_REQUIRE(t.numel == b * (n - 1), EMK_E_SHAPE, "grad_lengths must be (b, n_atoms-1)"); gl = F(t); }
  return emk_chain_in_plane_bwd(F(lv), ls, F(av), F(gv), b, n, ga, gl, stream);
}

int emk_dihedrals_to_cartesian(const float* dihedrals, const float* chain, int64_t chain_batch_stride, int64_t b, int64_t n_atoms,
                               int one_way, float* xyz, void* stream) {
  return d2c_general_device(dihedrals, chain, chain_batch_stride, b, n_atoms, one_way, xyz, as_stream(stream));
}
int emk_dihedrals_to_cartesian_bwd(const float* xyz, const float* grad_xyz, int64_t b, int64_t n_atoms, int one_way,
                                   float* grad_dihedrals, void* stream) {
  EMK_REQUIRE(b == 0 || (xyz && grad_xyz && grad_dihedrals), EMK_E_NULL, "emk_dihedrals_to_cartesian_bwd: NULL pointer argument");
  EMK_REQUIRE(n_atoms >= 4 && n_atoms < (1 << 20), EMK_E_SHAPE, "emk_dihedrals_to_cartesian_bwd: need 4 <= n_atoms < 2^20");
  BwdParams p{nullptr, 0, nullptr, xyz, grad_xyz, b, (int)n_atoms, 0, one_way ? 0 : (int)(n_atoms / 2 - 1), 0, nullptr, grad_dihedrals, nullptr};
  return backmap_bwd_device(p, as_stream(stream));
}
int emk_dl_dihedrals_to_cartesian(const DLManagedTensor* dihedrals, const DLManagedTensor* chain, int one_way, DLManagedTensor* xyz,
                                  void* stream) {
  VIEW(dv, dihedrals, "dihedrals", 2, 2);
  VIEW(cv, chain, "cartesian", 2, 3);
  VIEW(ov, xyz, "xyz", 3, 3);
  const int64_t b = dv.shape[0], n = dv.shape[1] + 3;
  int64_t cs;
  if (cv.ndim == 2) {
    EMK_REQUIRE(cv.shape[0] == n && cv.shape[1] == 3, EMK_E_SHAPE, "emk_dl_dihedrals_to_cartesian: cartesian must be (n_atoms,3) = (%lld,3)", (long long)n);
    cs = 0;
  } else {
    EMK_REQUIRE(cv.shape[0] == b && cv.shape[1] == n && cv.shape[2] == 3, EMK_E_SHAPE, "emk_dl_dihedrals_to_cartesian: cartesian must be (b, n_atoms, 3)");
    cs = 3 * n;
  }
  EMK_REQUIRE(ov.shape[0] == b && ov.shape[1] == n && ov.shape[2] == 3, EMK_E_SHAPE, "emk_dl_dihedrals_to_cartesian: xyz must be (b, n_atoms, 3)");
  return d2c_general_device(F(dv), F(cv), cs, b, n, one_way, F(ov), as_stream(stream));
}
int emk_dl_dihedrals_to_cartesian_bwd(const DLManagedTensor* xyz, const DLManagedTensor* grad_xyz, int one_way,
                                      DLManagedTensor* grad_dihedrals, void* stream) {
  VIEW(xv, xyz, "xyz", 3, 3);
  VIEW(gv, grad_xyz, "grad_xyz", 3, 3);
  VIEW(ov, grad_dihedrals, "grad_dihedrals", 2, 2);
  const int64_t b = xv.shape[0], n = xv.shape[1];
  EMK_REQUIRE(xv.shape[2] == 3 && gv.numel == xv.numel && ov.shape[0] == b && ov.shape[1] == n - 3, EMK_E_SHAPE, "emk_dl_dihedrals_to_cartesian_bwd: shape mismatch");
  return emk_dihedrals_to_cartesian_bwd(F(xv), F(gv), b, n, one_way, F(ov), stream);
}

int emk_dihedrals_to_cartesian_chain_bwd(const float* chain, int64_t chain_batch_stride, const float* xyz, const float* grad_xyz, int64_t b,
                                         int64_t n_atoms, int one_way, float* grad_chain, void* stream) {
  return d2c_chain_bwd_device(chain, chain_batch_stride, xyz, grad_xyz, b, n_atoms, one_way, grad_chain, as_stream(stream));
}
int emk_dl_dihedrals_to_cartesian_chain_bwd(const DLManagedTensor* chain, const DLManagedTensor* xyz, const DLManagedTensor* grad_xyz,
                                            int one_way, DLManagedTensor* grad_chain, void* stream) {
  VIEW(cv, chain, "cartesian", 2, 3);
  VIEW(xv, xyz, "xyz", 3, 3);
  VIEW(gv, grad_xyz, "grad_xyz", 3, 3);
  VIEW(ov, grad_chain, "grad_chain", 3, 3);
  const int64_t b = xv.shape[0], n = xv.shape[1];
  EMK_REQUIRE(xv.shape[2] == 3 && gv.numel == xv.numel && ov.numel == xv.numel, EMK_E_SHAPE,
              "emk_dl_dihedrals_to_cartesian_chain_bwd: xyz, grad_xyz and grad_chain must be (b, n_atoms, 3)");
  int64_t cs;
  if (cv.ndim == 2) {
    EMK_REQUIRE(cv.shape[0] == n && cv.shape[1] == 3, EMK_E_SHAPE, "emk_dl_dihedrals_to_cartesian_chain_bwd: cartesian must be (n_atoms,3)");
    cs = 0;
  } else {
    EMK_REQUIRE(cv.shape[0] == b && cv.shape[1] == n && cv.shape[2] == 3, EMK_E_SHAPE, "emk_dl_dihedrals_to_cartesian_chain_bwd: cartesian must be (b,n_atoms,3)");
    cs = 3 * n;
  }
  return d2c_chain_bwd_device(F(cv), cs, F(xv), F(gv), b, n, one_way, F(ov), as_stream(stream));
}

int emk_guess_sp2_atoms(const float* xyz, int64_t b, int64_t n_atoms, const int64_t* indices, int64_t n_idx, double angle_to_previous,
                        double bond_length, float* out, void* stream) {
  return guess_sp2_device(xyz, b, n_atoms, indices, n_idx, angle_to_previous, bond_length, out, as_stream(stream));
}
int emk_merge_cartesians(const float* central, int64_t b, int64_t n_atoms, const int64_t* h_after, int64_t n_h_after, const int64_t* o_after,
                         int64_t n_o_after, const float* h_xyz, int64_t n_h, const float* o_xyz, int64_t n_o, float* out, void* stream) {
  return merge_cartesians_device(central, b, n_atoms, h_after, n_h_after, o_after, n_o_after, h_xyz, n_h, o_xyz, n_o, out, as_stream(stream));
}
int emk_backbone_amide_atoms(const float* central, int64_t b, int64_t n_atoms, const int64_t* h_after, int64_t n_h_after, const int64_t* o_after,
                             int64_t n_o_after, double h_angle, double h_length, double o_angle, double o_length, float* out, int64_t n_out,
                             void* stream) {
  return backbone_amide_device(central, b, n_atoms, h_after, n_h_after, o_after, n_o_after, h_angle, h_length, o_angle, o_length, out, n_out,
                               as_stream(stream));
}
int64_t emk_merged_atom_count(int64_t n_atoms, const int64_t* h_after, int64_t n_h_after, const int64_t* o_after, int64_t n_o_after) {
  return merged_atom_count(n_atoms, h_after, n_h_after, o_after, n_o_after);
}

int emk_set_dihedrals(const float* start, int64_t start_frames, int64_t n_atoms, const int32_t* quads, const int32_t* bonds,
                      const int32_t* far_offsets, const int32_t* far_atoms, int64_t n_dihedrals, const float* targets, int64_t frames,
                      float* out, void* stream) {
  return set_dihedrals_device(start, start_frames, n_atoms, quads, bonds, far_offsets, far_atoms, n_dihedrals, targets, frames, out,
                              as_stream(stream));
}

// ---- back-mapping with side chains ---------------------------------------------------------------------------------
static inline const emk::SidechainPlan* as_plan(const emk_sidechain_plan* p) { return reinterpret_cast<const emk::SidechainPlan*>(p); }

int emk_sidechain_plan_create(int64_t n_residues, const int32_t* n_side_dihedrals, emk_sidechain_plan** plan) {
  emk::SidechainPlan* pl = nullptr;
  int rc = sidechain_plan_create(n_residues, n_side_dihedrals, &pl);
  if (plan) *plan = reinterpret_cast<emk_sidechain_plan*>(pl);
  else if (pl) sidechain_plan_destroy(pl);
  return rc;
}
void emk_sidechain_plan_destroy(emk_sidechain_plan* plan) { sidechain_plan_destroy(reinterpret_cast<emk::SidechainPlan*>(plan)); }
int emk_sidechain_plan_info(const emk_sidechain_plan* plan, int64_t* info) { return sidechain_plan_info(as_plan(plan), info); }
int emk_sidechain_plan_ops(const emk_sidechain_plan* plan, int32_t* ops) { return sidechain_plan_ops(as_plan(plan), ops); }

int emk_sidechain_backmap(const emk_sidechain_plan* plan, const float* central_distances, const float* central_angles,
                          const float* central_dihedrals, const float* side_distances, const float* side_angles,
                          const float* side_dihedrals, int64_t frames, float* xyz, double* saved_state, void* stream) {
  const float* in[6] = {central_distances, central_angles, central_dihedrals, side_distances, side_angles, side_dihedrals};
  return sidechain_backmap_device(as_plan(plan), in, frames, xyz, saved_state, as_stream(stream));
}
int emk_sidechain_backmap_bwd(const emk_sidechain_plan* plan, const float* central_distances, const float* central_angles,
                              const float* central_dihedrals, const float* side_distances, const float* side_angles,
                              const float* side_dihedrals, int64_t frames, const float* grad_xyz, const double* saved_state,
                              float* grad_central_distances,
                              float* grad_central_angles, float* grad_central_dihedrals, float* grad_side_distances,
                              float* grad_side_angles, float* grad_side_dihedrals, void* stream) {
  const float* in[6] = {central_distances, central_angles, central_dihedrals, side_distances, side_angles, side_dihedrals};
  float* gin[6] = {grad_central_distances, grad_central_angles, grad_central_dihedrals, grad_side_distances, grad_side_angles, grad_side_dihedrals};
  return sidechain_backmap_bwd_device(as_plan(plan), in, frames, grad_xyz, saved_state, gin, as_stream(stream));
}

static int sidechain_views(const char* who, const emk_sidechain_plan* plan, const DLManagedTensor* const* t, bool optional, const float** ptr,
                           int64_t* frames) {
  EMK_REQUIRE(plan, EMK_E_NULL, "%s: NULL plan", who);
  static const char* names[6] = {"central_distances", "central_angles", "central_dihedrals", "side_distances", "side_angles", "side_dihedrals"};
  const int* cols = sidechain_plan_cols(as_plan(plan));
  for (int k = 0; k < 6; k++) {
    ptr[k] = nullptr;
    if (optional && !t[k]) continue;
    View v;
    int rc = view_of(t[k], names[k], kDLFloat, 32, 2, 2, &v);
    if (rc) return rc;
    EMK_REQUIRE(v.shape[1] == cols[k], EMK_E_SHAPE, "%s: %s has %lld columns, the plan needs %d", who, names[k], (long long)v.shape[1], cols[k]);
    if (*frames < 0) *frames = v.shape[0];
    EMK_REQUIRE(v.shape[0] == *frames, EMK_E_SHAPE, "%s: %s has %lld frames, expected %lld", who, names[k], (long long)v.shape[0], (long long)*frames);
    ptr[k] = F(v);
  }
  return EMK_OK;
}

static int saved_view(const char* who, const emk_sidechain_plan* plan, const DLManagedTensor* t, int64_t frames, double** ptr) {
  *ptr = nullptr;
  if (!t) return EMK_OK;
  View v;
  int rc = view_of(t, "saved_state", kDLFloat, 64, 2, 2, &v);
  if (rc) return rc;
  EMK_REQUIRE(v.shape[0] == frames && v.shape[1] == sidechain_saved_doubles(as_plan(plan)), EMK_E_SHAPE,
              "%s: saved_state must be (%lld, %lld) float64 (emk_sidechain_saved_size)", who, (long long)frames,
              (long long)sidechain_saved_doubles(as_plan(plan)));
  *ptr = static_cast<double*>(v.data);
  return EMK_OK;
}
int64_t emk_sidechain_saved_size(const emk_sidechain_plan* plan) { return sidechain_saved_doubles(as_plan(plan)); }

int emk_dl_sidechain_backmap(const emk_sidechain_plan* plan, const DLManagedTensor* const* inputs, DLManagedTensor* xyz,
                             DLManagedTensor* saved_state, void* stream) {
  EMK_REQUIRE(inputs, EMK_E_NULL, "emk_dl_sidechain_backmap: NULL inputs");
  const float* in[6];
  int64_t frames = -1;
  int rc = sidechain_views("emk_dl_sidechain_backmap", plan, inputs, false, in, &frames);
  if (rc) return rc;
  VIEW(ov, xyz, "xyz", 3, 3);
  EMK_REQUIRE(ov.shape[0] == frames && ov.shape[1] == sidechain_plan_atoms(as_plan(plan)) && ov.shape[2] == 3, EMK_E_SHAPE,
              "emk_dl_sidechain_backmap: xyz must be (%lld, %d, 3)", (long long)frames, sidechain_plan_atoms(as_plan(plan)));
  double* sv;
  rc = saved_view("emk_dl_sidechain_backmap", plan, saved_state, frames, &sv);
  if (rc) return rc;
  return sidechain_backmap_device(as_plan(plan), in, frames, F(ov), sv, as_stream(stream));
}
int emk_dl_sidechain_backmap_bwd(const emk_sidechain_plan* plan, const DLManagedTensor* const* inputs, const DLManagedTensor* grad_xyz,
                                 const DLManagedTensor* saved_state, DLManagedTensor* const* grad_inputs, void* stream) {
  EMK_REQUIRE(inputs && grad_inputs, EMK_E_NULL, "emk_dl_sidechain_backmap_bwd: NULL inputs");
  const float* in[6];
  const float* gin_c[6];
  int64_t frames = -1;
  int rc = sidechain_views("emk_dl_sidechain_backmap_bwd", plan, inputs, false, in, &frames);
  if (rc) return rc;
  rc = sidechain_views("emk_dl_sidechain_backmap_bwd (gradients)", plan, grad_inputs, true, gin_c, &frames);
  if (rc) return rc;
  VIEW(gv, grad_xyz, "grad_xyz", 3, 3);
  EMK_REQUIRE(gv.shape[0] == frames && gv.shape[1] == sidechain_plan_atoms(as_plan(plan)) && gv.shape[2] == 3, EMK_E_SHAPE,
              "emk_dl_sidechain_backmap_bwd: grad_xyz must be (%lld, %d, 3)", (long long)frames, sidechain_plan_atoms(as_plan(plan)));
  float* gin[6];
  for (int k = 0; k < 6; k++) gin[k] = const_cast<float*>(gin_c[k]);
  double* sv;
  rc = saved_view("emk_dl_sidechain_backmap_bwd", plan, saved_state, frames, &sv);
  if (rc) return rc;
  return sidechain_backmap_bwd_device(as_plan(plan), in, frames, F(gv), sv, gin, as_stream(stream));
}

int64_t emk_sidechain_pairwise_indices(int64_t n_residues, const int32_t* n_side_dihedrals, int64_t start, int64_t stop, int64_t step,
                                       int64_t* indices) {
  if (n_residues < 1 || !n_side_dihedrals) return -1;
  int64_t first, cnt, st;
  if (resolve_slice(3 * n_residues, start, stop, step, &first, &cnt, &st)) return -1;
  return sidechain_pairwise_indices(n_residues, n_side_dihedrals, first, cnt, st, indices);
}

int emk_gather_atoms(const float* xyz, int64_t b, int64_t n_atoms, const int32_t* index_dev, int64_t m, float* out, void* stream) {
  return gather_atoms_device(xyz, b, n_atoms, index_dev, m, out, as_stream(stream));
}
int emk_gather_atoms_bwd(const float* grad_out, int64_t b, int64_t n_atoms, const int32_t* index_dev, int64_t m, float* grad_xyz, void* stream) {
  return gather_atoms_bwd_device(grad_out, b, n_atoms, index_dev, m, grad_xyz, as_stream(stream));
}
static int index_view(const DLManagedTensor* t, View* v) { return view_of(t, "index", kDLInt, 32, 1, 1, v); }
int emk_dl_gather_atoms(const DLManagedTensor* xyz, const DLManagedTensor* index, DLManagedTensor* out, void* stream) {
  VIEW(xv, xyz, "xyz", 3, 3);
  View iv;
  int rc = index_view(index, &iv);
  if (rc) return rc;
  VIEW(ov, out, "out", 3, 3);
  EMK_REQUIRE(xv.shape[2] == 3 && ov.shape[0] == xv.shape[0] && ov.shape[1] == iv.shape[0] && ov.shape[2] == 3, EMK_E_SHAPE,
              "emk_dl_gather_atoms: need xyz (b, n, 3), index (m), out (b, m, 3)");
  return gather_atoms_device(F(xv), xv.shape[0], xv.shape[1], static_cast<const int32_t*>(iv.data), iv.shape[0], F(ov), as_stream(stream));
}
int emk_dl_gather_atoms_bwd(const DLManagedTensor* grad_out, const DLManagedTensor* index, DLManagedTensor* grad_xyz, void* stream) {
  VIEW(gv, grad_out, "grad_out", 3, 3);
  View iv;
  int rc = index_view(index, &iv);
  if (rc) return rc;
  VIEW(xv, grad_xyz, "grad_xyz", 3, 3);
  EMK_REQUIRE(xv.shape[2] == 3 && gv.shape[0] == xv.shape[0] && gv.shape[1] == iv.shape[0] && gv.shape[2] == 3, EMK_E_SHAPE,
              "emk_dl_gather_atoms_bwd: need grad_out (b, m, 3), index (m), grad_xyz (b, n, 3)");
  return gather_atoms_bwd_device(F(gv), xv.shape[0], xv.shape[1], static_cast<const int32_t*>(iv.data), iv.shape[0], F(xv), as_stream(stream));
}

}  // extern "C"
