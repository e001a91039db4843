// K6 radius graph (two-pass count / scan / fill CSR) with K8 (neighbour-type histogram + degree
// statistics) fused into the count pass.
// Reference: /root/reference/hovernet_tile_inference.ipynb:2964-2975 (cKDTree.query_ball_tree, i<j edge
// loop), :3041-3042 (np.linalg.norm distances, float32 edge_attr); composition / degree per SURVEY A.5.
// Acceptance test is d2 <= r*r in float64 with d2 = fl(fl(dx*dx)+fl(dy*dy)) - what scipy evaluates.
#include <cmath>
#include "pg_query.cuh"

namespace {

constexpr int TPB = 128;
constexpr int FILL_CAP = 32;
constexpr int HIST_SMEM_MAX = 2048;

struct block_stats {
  int mn, mx;
  long long sum, sumsq;
};

// block-wide reduction of the degree statistics, one set of atomics per CTA
__device__ __forceinline__ void reduce_degree_stats(int deg, bool valid, pg_degree_stats* stats) {
  __shared__ int s_mn[TPB / 32], s_mx[TPB / 32];
  __shared__ long long s_sum[TPB / 32], s_sq[TPB / 32];
  __shared__ int s_cnt[TPB / 32];
  int mn = valid ? deg : 0x7fffffff, mx = valid ? deg : -1, cnt = valid ? 1 : 0;
  long long sum = valid ? deg : 0, sq = valid ? (long long)deg * deg : 0;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    sum += __shfl_xor_sync(0xffffffffu, sum, d);
    sq += __shfl_xor_sync(0xffffffffu, sq, d);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; s_sum[warp] = sum; s_sq[warp] = sq; s_cnt[warp] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < TPB / 32; ++w) {
      mn = min(mn, s_mn[w]); mx = max(mx, s_mx[w]); sum += s_sum[w]; sq += s_sq[w]; cnt += s_cnt[w];
    }
    if (cnt > 0) {
      atomicMin(&stats->min_degree, mn);
      atomicMax(&stats->max_degree, mx);
      atomicAdd((unsigned long long*)&stats->sum_degree, (unsigned long long)sum);
      atomicAdd((unsigned long long*)&stats->sumsq_degree, (unsigned long long)sq);
      atomicAdd((unsigned long long*)&stats->n_nodes, (unsigned long long)cnt);
    }
  }
}

__global__ void init_stats_kernel(pg_degree_stats* stats) {
  stats->min_degree = 0x7fffffff;
  stats->max_degree = -1;
  stats->sum_degree = 0;
  stats->sumsq_degree = 0;
  stats->n_nodes = 0;
}
// an empty graph reports min = max = 0
__global__ void finish_stats_kernel(pg_degree_stats* stats) {
  if (stats->n_nodes == 0) { stats->min_degree = 0; stats->max_degree = 0; }
}

// Count pass. One thread per point in cell order (a warp = spatial neighbours, so the candidate
// runs it walks are shared through L1). Writes, at the point's own row: the CSR row count (all
// neighbours, or only gid_j > gid_i when `upper`), and fused over all neighbours: degree,
// per-type neighbour counts, degree statistics and histogram.
template <int TMAX>
__global__ void __launch_bounds__(TPB)
radius_count_kernel(pg_grid_view g, double r2, int R, int upper, int32_t* __restrict__ row_count,
                    int32_t* __restrict__ degree, int32_t* __restrict__ nbr_count, int n_types,
                    pg_degree_stats* stats, int32_t* hist, int hist_len) {
  extern __shared__ int s_hist[];
  const bool use_smem_hist = hist != nullptr && hist_len <= HIST_SMEM_MAX;
  if (use_smem_hist) {
    for (int i = threadIdx.x; i < hist_len; i += TPB) s_hist[i] = 0;
    __syncthreads();
  }
  const int p = blockIdx.x * TPB + threadIdx.x;
  bool valid = p < g.n;
  int4 me = make_int4(0, 0, 0, 0);
  if (valid) {
    me = g.s_meta[p];
    valid = me.x < g.n_query;
  }
  int deg = 0;
  if (valid) {
    const double2 q = g.s_xy[p];
    const int cx = pg_cell_coord(q.x, g.x0, g.inv_cell, g.nx);
    const int cy = pg_cell_coord(q.y, g.y0, g.inv_cell, g.ny);
    int tc[TMAX];
#pragma unroll
    for (int t = 0; t < TMAX; ++t) tc[t] = 0;
    int up = 0;
    pg_visit_block(g, cx, cy, R, [&](int b, int e) {
      for (int j = b; j < e; ++j) {
        const double2 c = g.s_xy[j];
        const double d2 = pg_dist2(q.x, q.y, c.x, c.y);
        if (d2 <= r2 && j != p) {
          const int4 m = g.s_meta[j];
          ++deg;
          up += (m.y > me.y);
#pragma unroll
          for (int t = 0; t < TMAX; ++t) tc[t] += (m.z == t + 1);
        }
      }
    });
    row_count[me.x] = upper ? up : deg;
    if (degree) degree[me.x] = deg;
    if (nbr_count) {
#pragma unroll
      for (int t = 0; t < TMAX; ++t)
        if (t < n_types) nbr_count[(int64_t)me.x * n_types + t] = tc[t];
    }
    if (hist) {
      const int bin = min(deg, hist_len - 1);
      if (use_smem_hist) atomicAdd(&s_hist[bin], 1);
      else atomicAdd(&hist[bin], 1);
    }
  }
  if (stats) reduce_degree_stats(deg, valid, stats);
  if (use_smem_hist) {
    __syncthreads();
    for (int i = threadIdx.x; i < hist_len; i += TPB) {
      const int c = s_hist[i];
      if (c) atomicAdd(&hist[i], c);
    }
  }
}

// Fill pass. Same walk; a row is emitted in ascending column (gid) order in chunks of FILL_CAP:
// each pass keeps the FILL_CAP smallest accepted ids above the last one written, so ordinary rows
// take one walk and heavy rows take ceil(count / FILL_CAP) walks without extra memory.
__global__ void __launch_bounds__(TPB)
radius_fill_kernel(pg_grid_view g, double r2, int R, int upper, const int32_t* __restrict__ row_ptr,
                   int32_t* __restrict__ col, float* __restrict__ dist32, double* __restrict__ dist64,
                   long long* __restrict__ edges, long long* __restrict__ edge_index,
                   float* __restrict__ edge_attr, long long n_edges, long long capacity, int32_t* overflow) {
  const int p = blockIdx.x * TPB + threadIdx.x;
  if (p >= g.n) return;
  const int4 me = g.s_meta[p];
  if (me.x >= g.n_query) return;
  const int base = row_ptr[me.x];
  const int cnt = row_ptr[me.x + 1] - base;
  if (cnt <= 0) return;
  if ((long long)base + cnt > capacity) { atomicExch(overflow, 1); return; }
  const double2 q = g.s_xy[p];
  const int cx = pg_cell_coord(q.x, g.x0, g.inv_cell, g.nx);
  const int cy = pg_cell_coord(q.y, g.y0, g.inv_cell, g.ny);
  pg_sorted_chunk<FILL_CAP> buf;
  int emitted = 0;
  int last = upper ? me.y : -1;
  while (emitted < cnt) {
    buf.reset(last);
    pg_visit_block(g, cx, cy, R, [&](int b, int e) {
      for (int j = b; j < e; ++j) {
        const double2 c = g.s_xy[j];
        const double d2 = pg_dist2(q.x, q.y, c.x, c.y);
        if (d2 <= r2 && j != p) buf.push(g.s_meta[j].y, d2);
      }
    });
    if (buf.m == 0) break;  // cannot happen when row_ptr came from the matching count pass
    for (int t = 0; t < buf.m; ++t) {
      const long long o = (long long)base + emitted + t;
      const double d = sqrt(buf.val[t]);
      col[o] = buf.key[t];
      if (dist32) dist32[o] = (float)d;
      if (dist64) dist64[o] = d;
      if (edges) { edges[2 * o] = me.y; edges[2 * o + 1] = buf.key[t]; }
      if (edge_index) {  // [2, 2E] = hstack(edges.T, edges[:, ::-1].T), ipynb:3021 (see SURVEY B-3)
        edge_index[o] = me.y; edge_index[n_edges + o] = buf.key[t];
        edge_index[2 * n_edges + o] = buf.key[t]; edge_index[3 * n_edges + o] = me.y;
      }
      if (edge_attr) { edge_attr[o] = (float)d; edge_attr[n_edges + o] = (float)d; }  // ipynb:3041-3042
    }
    emitted += buf.m;
    last = buf.key[buf.m - 1];
  }
}

__global__ void copy_total_kernel(const int32_t* src, int32_t* dst) { *dst = *src; }

}  // namespace

extern "C" {

int pg_radius_count(pg_handle* h, double r, int32_t flags, int32_t* row_ptr, int32_t* degree,
                    int32_t* nbr_count, int32_t n_types, pg_degree_stats* stats, int32_t* hist,
                    int32_t hist_len, pg_stream stream) {
  if (!h) return PG_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  PG_CUDA(h, cudaSetDevice(h->device));
  h->last_stream = s;
  if (!h->grid.built) return pg_set_error(h, PG_ERR_STATE, "pg_radius_count: call pg_grid_build first");
  PG_REQUIRE(h, r >= 0 && std::isfinite(r), "pg_radius_count: r must be finite and >= 0");
  PG_REQUIRE(h, row_ptr != nullptr, "pg_radius_count: row_ptr is NULL");
  PG_REQUIRE(h, flags == PG_RADIUS_SYMMETRIC || flags == PG_RADIUS_UPPER, "pg_radius_count: bad flags %d", flags);
  PG_REQUIRE(h, !nbr_count || (n_types >= 1 && n_types <= PG_MAX_TYPES), "pg_radius_count: n_types must be in 1..%d", PG_MAX_TYPES);
  PG_REQUIRE(h, !hist || hist_len >= 1, "pg_radius_count: hist_len must be >= 1");
  const pg_grid& gr = h->grid;
  const int nq = gr.n_query;
  int rc;
  if ((rc = pg_reserve(h, h->row_count, ((size_t)nq + 4) * sizeof(int32_t)))) return rc;
  // ring radius in cells: every point within r lies at most R cells away along each axis
  // (relative slack far above the rounding of the cell coordinate)
  int R = (int)std::ceil(r * gr.inv_cell * (1.0 + 1e-9) + 1e-9);
  if (R < 1) R = 1;
  h->radius_r = r;
  h->radius_flags = flags;
  if (stats) { PG_LAUNCH(h, s, "init_stats_kernel", init_stats_kernel<<<1, 1, 0, s>>>(stats)); }
  if (hist) PG_CUDA(h, cudaMemsetAsync(hist, 0, (size_t)hist_len * sizeof(int32_t), s));
  // ghost rows never write their count: clear so the scan sees zeros (n_query rows only are scanned)
  if (gr.n > 0 && nq > 0) {
    pg_grid_view v = pg_make_view(h);
    const double r2 = r * r;
    const int blocks = pg_div_up(gr.n, TPB);
    const size_t smem = (hist && hist_len <= HIST_SMEM_MAX) ? (size_t)hist_len * sizeof(int) : 0;
    const int upper = flags == PG_RADIUS_UPPER;
    if (!nbr_count || n_types <= 8)
      PG_LAUNCH(h, s, "radius_count_kernel<8>", radius_count_kernel<8><<<blocks, TPB, smem, s>>>(v, r2, R, upper, (int32_t*)h->row_count.p, degree, nbr_count,
                                                      n_types, stats, hist, hist_len));
    else
      PG_LAUNCH(h, s, "radius_count_kernel<16>", radius_count_kernel<16><<<blocks, TPB, smem, s>>>(v, r2, R, upper, (int32_t*)h->row_count.p, degree, nbr_count,
                                                       n_types, stats, hist, hist_len));
    PG_LAUNCH_CHECK(h);
  }
  if (stats) { PG_LAUNCH(h, s, "finish_stats_kernel", finish_stats_kernel<<<1, 1, 0, s>>>(stats)); }
  if ((rc = pg_scan_i32(h, (const int32_t*)h->row_count.p, row_ptr, nq, s))) return rc;
  PG_LAUNCH(h, s, "copy_total_kernel", copy_total_kernel<<<1, 1, 0, s>>>(row_ptr + nq, (int32_t*)((char*)h->misc.p + PG_MISC_TOTALS)));
  PG_LAUNCH_CHECK(h);
  return PG_OK;
}

int pg_radius_total(pg_handle* h, int64_t* total) {
  if (!h || !total) return PG_ERR_INVALID;
  PG_CUDA(h, cudaSetDevice(h->device));
  PG_CUDA(h, cudaMemcpyAsync(&h->pinned[0], (char*)h->misc.p + PG_MISC_TOTALS, sizeof(int32_t),
                             cudaMemcpyDeviceToHost, h->last_stream));
  PG_CUDA(h, cudaStreamSynchronize(h->last_stream));
  *total = h->pinned[0];
  return PG_OK;
}

int pg_radius_fill(pg_handle* h, const int32_t* row_ptr, int32_t* col, float* dist32, double* dist64,
                   int64_t* edges_i64, int64_t* edge_index, float* edge_attr, int64_t n_edges,
                   int64_t capacity, pg_stream stream) {
  if (!h) return PG_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  PG_CUDA(h, cudaSetDevice(h->device));
  h->last_stream = s;
  if (!h->grid.built) return pg_set_error(h, PG_ERR_STATE, "pg_radius_fill: call pg_grid_build / pg_radius_count first");
  PG_REQUIRE(h, row_ptr != nullptr, "pg_radius_fill: row_ptr is NULL");
  PG_REQUIRE(h, capacity >= 0, "pg_radius_fill: capacity < 0");
  PG_REQUIRE(h, capacity == 0 || col != nullptr, "pg_radius_fill: col is NULL");
  PG_REQUIRE(h, !(edge_index || edge_attr) || (h->radius_flags == PG_RADIUS_UPPER && n_edges >= 0 && n_edges <= capacity),
             "pg_radius_fill: edge_index / edge_attr need the UPPER count pass and 0 <= n_edges <= capacity");
  const pg_grid& gr = h->grid;
  if (gr.n == 0 || gr.n_query == 0 || capacity == 0) {
    // an empty buffer is only fine for an empty result: let the kernel flag rows it cannot place
    if (gr.n == 0 || gr.n_query == 0) return PG_OK;
  }
  int R = (int)std::ceil(h->radius_r * gr.inv_cell * (1.0 + 1e-9) + 1e-9);
  if (R < 1) R = 1;
  pg_grid_view v = pg_make_view(h);
  PG_LAUNCH(h, s, "radius_fill_kernel", radius_fill_kernel<<<pg_div_up(gr.n, TPB), TPB, 0, s>>>(
      v, h->radius_r * h->radius_r, R, h->radius_flags == PG_RADIUS_UPPER, row_ptr, col, dist32, dist64,
      (long long*)edges_i64, (long long*)edge_index, edge_attr, (long long)n_edges, (long long)capacity,
      (int32_t*)((char*)h->misc.p + PG_MISC_OVERFLOW)));
  PG_LAUNCH_CHECK(h);
  return PG_OK;
}

}  // extern "C"
