// K10: node-feature assembly for the GNN input.
// Reference: /root/reference/hovernet_tile_inference.ipynb:2903 (cell 21: per-column z-score with
// mean / std(ddof=0), both NaN-skipping; a column whose sigma is 0 or NaN becomes all 0.0) and
// :2950 (cell 23: pd.get_dummies(type, prefix="type") one-hot columns in ascending type order, features =
// one-hot columns followed by the *_z columns).
//
// Three launches, no atomics on floats, so the result does not depend on scheduling:
//   stats x2 pandas' own two-pass mean / variance (see feature_stats_kernel), all columns per launch;
//   assemble one thread per output element of x [N, n_onehot + n_feat] (row-major float32, coalesced stores).
#include <algorithm>
#include "pg_common.cuh"

namespace {

constexpr int TPB = 256;

// One reduction pass over every column (blockIdx.y = column): PASS 0 accumulates (count, sum) of the non-NaN
// values, PASS 1 the squared deviations from the mean PASS 0 left in stats_out - pandas' own two-pass nanvar.
// Fixed order throughout: a thread's strided share, a shuffle tree, the warps of the CTA in index order, and the
// CTA partials in index order by the last CTA of the column (ticket); nothing depends on scheduling.
template <int PASS>
__global__ void __launch_bounds__(TPB)
feature_stats_kernel(const double* __restrict__ feat, int n, double2* __restrict__ partial, unsigned int* tickets,
                     double* __restrict__ stats_out) {
  __shared__ double s_a[TPB / 32], s_b[TPB / 32];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c = blockIdx.y;
  const double* col = feat + (int64_t)c * n;
  const double mean = PASS == 1 ? stats_out[2 * c] : 0.0;
  double a = 0.0, b = 0.0;  // PASS 0: count, sum; PASS 1: count, sum of squared deviations
  for (int i = blockIdx.x * TPB + threadIdx.x; i < n; i += gridDim.x * TPB) {
    const double v = col[i];
    if (v == v) {  // pandas skips NaN in both mean and std
      a += 1.0;
      if (PASS == 0) b += v;
      else { const double d = v - mean; b += d * d; }
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, d); b += __shfl_xor_sync(0xffffffffu, b, d); }
  if (lane == 0) { s_a[warp] = a; s_b[warp] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int w = 0; w < TPB / 32; ++w) { ta += s_a[w]; tb += s_b[w]; }
    partial[(int64_t)c * gridDim.x + blockIdx.x] = make_double2(ta, tb);
    __threadfence();
    s_last = atomicAdd(&tickets[c], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence();
  const volatile double2* p = partial + (int64_t)c * gridDim.x;
  double ta = 0.0, tb = 0.0;
  for (unsigned int g = 0; g < gridDim.x; ++g) { ta += p[g].x; tb += p[g].y; }
  const double nan_ = __longlong_as_double(0x7ff8000000000000ll);
  if (PASS == 0) stats_out[2 * c] = ta > 0.0 ? tb / ta : nan_;
  else stats_out[2 * c + 1] = ta > 0.0 ? sqrt(tb / ta) : nan_;  // ddof = 0
  tickets[c] = 0u;  // re-armed for the next launch
}

__global__ void __launch_bounds__(TPB)
feature_assemble_kernel(const double* __restrict__ feat, const int32_t* __restrict__ type,
                        const int32_t* __restrict__ onehot_values, int n, int n_feat, int n_onehot,
                        const double* __restrict__ stats, float* __restrict__ x) {
  const int width = n_onehot + n_feat;
  const int64_t e = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (e >= (int64_t)n * width) return;
  const int i = (int)(e / width), c = (int)(e - (int64_t)i * width);
  float v;
  if (c < n_onehot) {
    v = type[i] == onehot_values[c] ? 1.0f : 0.0f;
  } else {
    const int f = c - n_onehot;
    const double mu = stats[2 * f], sigma = stats[2 * f + 1];
    // cell 21: `if sigma == 0 or np.isnan(sigma): col_z = 0.0` (the whole column, NaN rows included)
    v = (sigma == 0.0 || sigma != sigma) ? 0.0f : (float)((feat[(int64_t)f * n + i] - mu) / sigma);
  }
  x[e] = v;
}

// the same for narrow matrices (width <= TILE_WIDTH_MAX): a CTA owns TPB consecutive rows; thread t computes row t
// column by column (coalesced reads of type / feat, no integer division) into a shared-memory tile whose
// TPB x width floats are then written out as one contiguous, fully coalesced range of x
constexpr int TILE_WIDTH_MAX = 40;
__global__ void __launch_bounds__(TPB)
feature_assemble_tile_kernel(const double* __restrict__ feat, const int32_t* __restrict__ type,
                             const int32_t* __restrict__ onehot_values, int n, int n_feat, int n_onehot,
                             const double* __restrict__ stats, float* __restrict__ x) {
  extern __shared__ float s_tile[];  // [TPB][width], row stride = width (odd widths are conflict-free as they are)
  const int width = n_onehot + n_feat;
  const int row0 = blockIdx.x * TPB, i = row0 + threadIdx.x;
  if (i < n) {
    float* mine = s_tile + threadIdx.x * width;
    const int ty = n_onehot ? type[i] : 0;
    for (int c = 0; c < n_onehot; ++c) mine[c] = ty == onehot_values[c] ? 1.0f : 0.0f;
    for (int f = 0; f < n_feat; ++f) {
      const double mu = stats[2 * f], sigma = stats[2 * f + 1];
      mine[n_onehot + f] = (sigma == 0.0 || sigma != sigma) ? 0.0f : (float)((feat[(int64_t)f * n + i] - mu) / sigma);
    }
  }
  __syncthreads();
  const int rows = min(TPB, n - row0);
  float* dst = x + (int64_t)row0 * width;
  for (int e = threadIdx.x; e < rows * width; e += TPB) dst[e] = s_tile[e];
}

}  // namespace

extern "C" int pg_node_features(pg_handle* h, int32_t n, int32_t n_feat, const double* feat, const int32_t* type,
                                const int32_t* onehot_values, int32_t n_onehot, float* x, double* stats,
                                pg_stream stream) {
  if (!h) return PG_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  PG_CUDA(h, cudaSetDevice(h->device));
  h->last_stream = s;
  PG_REQUIRE(h, n >= 0 && n_feat >= 0 && n_onehot >= 0, "pg_node_features: negative size");
  PG_REQUIRE(h, n_feat == 0 || (feat && stats), "pg_node_features: feat / stats is NULL");
  PG_REQUIRE(h, n_onehot == 0 || (type && onehot_values), "pg_node_features: type / onehot_values is NULL");
  PG_REQUIRE(h, (int64_t)n * (n_feat + n_onehot) == 0 || x, "pg_node_features: x is NULL");
  if (n_feat > 0) {
    PG_REQUIRE(h, n_feat <= PG_FEAT_MAX_COLS, "pg_node_features: at most %d feature columns", PG_FEAT_MAX_COLS);
    const int gx = n > 0 ? std::max(1, std::min(pg_div_up(n, TPB * 4), h->sm_count * 8 / n_feat + 1)) : 1;
    int rc = pg_reserve(h, h->cell_of, (size_t)n_feat * gx * sizeof(double2) + 64);
    if (rc) return rc;
    unsigned int* tickets = (unsigned int*)((char*)h->misc.p + PG_MISC_FEAT_TICKET);
    const dim3 grid(gx, n_feat);
    PG_LAUNCH(h, s, "feature_stats_kernel<0>", feature_stats_kernel<0><<<grid, TPB, 0, s>>>(feat, n, (double2*)h->cell_of.p, tickets, stats));
    PG_LAUNCH(h, s, "feature_stats_kernel<1>", feature_stats_kernel<1><<<grid, TPB, 0, s>>>(feat, n, (double2*)h->cell_of.p, tickets, stats));
    PG_LAUNCH_CHECK(h);
  }
  const int64_t elems = (int64_t)n * (n_feat + n_onehot);
  if (elems > 0) {
    PG_REQUIRE(h, elems / TPB < 0x7fffffff, "pg_node_features: feature matrix too large");
    const int width = n_feat + n_onehot;
    if (width <= TILE_WIDTH_MAX)
      PG_LAUNCH(h, s, "feature_assemble_tile_kernel", feature_assemble_tile_kernel<<<pg_div_up(n, TPB), TPB, (size_t)TPB * width * sizeof(float), s>>>(feat, type, onehot_values, n, n_feat, n_onehot, stats, x));
    else
    PG_LAUNCH(h, s, "feature_assemble_kernel", feature_assemble_kernel<<<pg_div_up(elems, TPB), TPB, 0, s>>>(feat, type, onehot_values, n, n_feat, n_onehot, stats, x));
    PG_LAUNCH_CHECK(h);
  }
  return PG_OK;
}
