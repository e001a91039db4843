// K10: node-feature assembly for the GNN input.
// Reference: /root/reference/hovernet_tile_inference.ipynb:2903 (cell 21: per-column z-score with
// mean / std(ddof=0), both NaN-skipping; a column whose sigma is 0 or NaN becomes all 0.0) and
// :2950 (cell 23: pd.get_dummies(type, prefix="type") one-hot columns in ascending type order, features =
// one-hot columns followed by the *_z columns).
//
// Two launches, no atomics on floats, so the result does not depend on scheduling:
//   stats    a fixed grid; every thread folds its strided share of a column into (count, mean, M2) and the
//            partials are merged pairwise (Chan et al.) lane -> warp -> CTA in a fixed order; the CTA's partial
//            goes to the workspace and the LAST CTA to finish (a ticket) merges the G partials of every column
//            in index order and leaves mean / sigma where the next launch and the caller find them.
//   assemble one thread per output element of x [N, n_onehot + n_feat] (row-major float32, coalesced stores).
#include <algorithm>
#include "pg_common.cuh"

namespace {

constexpr int TPB = 256;

struct moments {
  double n, mean, m2;
};

__device__ __forceinline__ moments merge(const moments& a, const moments& b) {
  if (b.n == 0.0) return a;
  if (a.n == 0.0) return b;
  moments r;
  r.n = a.n + b.n;
  const double d = b.mean - a.mean;
  r.mean = a.mean + d * (b.n / r.n);
  r.m2 = a.m2 + b.m2 + d * d * (a.n * b.n / r.n);
  return r;
}

__device__ __forceinline__ moments shfl_xor(const moments& v, int lane_mask) {
  moments r;
  r.n = __shfl_xor_sync(0xffffffffu, v.n, lane_mask);
  r.mean = __shfl_xor_sync(0xffffffffu, v.mean, lane_mask);
  r.m2 = __shfl_xor_sync(0xffffffffu, v.m2, lane_mask);
  return r;
}

// feat is column-major [n_feat][n]; partial is [n_feat][gridDim.x]; stats_out is [n_feat][2] = {mean, sigma}
__global__ void __launch_bounds__(TPB)
feature_stats_kernel(const double* __restrict__ feat, int n, int n_feat, moments* __restrict__ partial,
                     unsigned int* ticket, double* __restrict__ stats_out) {
  __shared__ moments s_part[TPB / 32];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = 0; c < n_feat; ++c) {
    const double* col = feat + (int64_t)c * n;
    moments m{0.0, 0.0, 0.0};
    for (int i = blockIdx.x * TPB + threadIdx.x; i < n; i += gridDim.x * TPB) {
      const double v = col[i];
      if (v == v) {  // pandas skips NaN in both mean and std
        m.n += 1.0;
        const double d = v - m.mean;
        m.mean += d / m.n;
        m.m2 += d * (v - m.mean);
      }
    }
    // lane l merges with lane l ^ d: written so that both sides compute the same value (lower lane first)
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const moments o = shfl_xor(m, d);
      m = (lane & d) ? merge(o, m) : merge(m, o);
    }
    if (lane == 0) s_part[warp] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      moments t = s_part[0];
      for (int w = 1; w < TPB / 32; ++w) t = merge(t, s_part[w]);
      partial[(int64_t)c * gridDim.x + blockIdx.x] = t;
    }
    __syncthreads();
  }
  __threadfence();
  if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int c = threadIdx.x; c < n_feat; c += TPB) {
    const volatile moments* p = partial + (int64_t)c * gridDim.x;
    moments t{0.0, 0.0, 0.0};
    for (unsigned int g = 0; g < gridDim.x; ++g) {
      moments q;
      q.n = p[g].n; q.mean = p[g].mean; q.m2 = p[g].m2;
      t = merge(t, q);
    }
    const double nan_ = __longlong_as_double(0x7ff8000000000000ll);
    stats_out[2 * c] = t.n > 0.0 ? t.mean : nan_;
    stats_out[2 * c + 1] = t.n > 0.0 ? sqrt(t.m2 / t.n) : nan_;  // ddof = 0
  }
  if (threadIdx.x == 0) *ticket = 0u;  // re-armed for the next call
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
    const int grid = n > 0 ? std::min(pg_div_up(n, TPB), h->sm_count * 4) : 1;
    int rc = pg_reserve(h, h->cell_of, (size_t)n_feat * grid * sizeof(moments) + 64);
    if (rc) return rc;
    unsigned int* ticket = (unsigned int*)((char*)h->misc.p + PG_MISC_FEAT_TICKET);
    PG_LAUNCH(h, s, "feature_stats_kernel", feature_stats_kernel<<<grid, TPB, 0, s>>>(feat, n, n_feat, (moments*)h->cell_of.p, ticket, stats));
    PG_LAUNCH_CHECK(h);
  }
  const int64_t elems = (int64_t)n * (n_feat + n_onehot);
  if (elems > 0) {
    PG_REQUIRE(h, elems / TPB < 0x7fffffff, "pg_node_features: feature matrix too large");
    PG_LAUNCH(h, s, "feature_assemble_kernel", feature_assemble_kernel<<<pg_div_up(elems, TPB), TPB, 0, s>>>(feat, type, onehot_values, n, n_feat, n_onehot, stats, x));
    PG_LAUNCH_CHECK(h);
  }
  return PG_OK;
}
