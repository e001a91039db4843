// K7 undirected union of directed kNN lists, `edges` (i<j) extraction from a symmetric CSR, and the
// stand-alone K8 (neighbour-type composition + degree statistics over any CSR).
// Reference: /root/reference/hovernet_tile_inference.ipynb:1865-1894 (nx.Graph union, weight = min),
// :2969-2975 (i<j edge list); composition / degree per SURVEY A.5 (README.md:127,136).
#include "pg_query.cuh"

namespace {

constexpr int TPB = 256;
constexpr int SORT_CAP = 48;
constexpr uint8_t RECIP_INVALID = 255, RECIP_NONE = 254;

// pass A: one thread per directed edge i->j. Records the slot of i inside j's list (or NONE) and
// counts, per node, the reverse-only edges it will receive.
// Lists hold ids in "column space" (global ids when the rows are a strip + halo of a larger slide):
// row_id[i] is the id of row i, id_map[id] the row of an id (-1 / >= n: no row here). NULL = identity.
__global__ void __launch_bounds__(TPB)
sym_mark_kernel(const int32_t* __restrict__ knn_idx, int n, int k, const int32_t* __restrict__ row_id,
                const int32_t* __restrict__ id_map, int n_ids, uint8_t* __restrict__ recip,
                int32_t* __restrict__ extra) {
  const int64_t e = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (e >= (int64_t)n * k) return;
  const int i = (int)(e / k);
  const int my_id = row_id ? row_id[i] : i;
  const int jid = knn_idx[e];
  if (jid < 0 || jid == my_id || (id_map && jid >= n_ids) || (!id_map && jid >= n)) { recip[e] = RECIP_INVALID; return; }
  const int j = id_map ? id_map[jid] : jid;
  if (j < 0 || j >= n) { recip[e] = RECIP_NONE; return; }  // neighbour has no row here: own entry only
  const int32_t* row = knn_idx + (int64_t)j * k;
  int found = RECIP_NONE;
  for (int s = 0; s < k; ++s)
    if (row[s] == my_id) { found = s; break; }
  recip[e] = (uint8_t)found;
  if (found == RECIP_NONE) atomicAdd(&extra[j], 1);
}

// per node: own valid entries (start value of its append cursor) and the undirected degree
__global__ void __launch_bounds__(TPB)
sym_degree_kernel(const uint8_t* __restrict__ recip, const int32_t* __restrict__ extra, int n, int k,
                  int32_t* __restrict__ cursor, int32_t* __restrict__ row_count) {
  const int i = blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  int own = 0;
  for (int s = 0; s < k; ++s) own += recip[(int64_t)i * k + s] != RECIP_INVALID;
  cursor[i] = own;
  row_count[i] = own + extra[i];
}

template <class DT>
__global__ void __launch_bounds__(TPB)
sym_scatter_kernel(const int32_t* __restrict__ knn_idx, const DT* __restrict__ dist, int n, int k,
                   const int32_t* __restrict__ row_id, const int32_t* __restrict__ id_map,
                   const uint8_t* __restrict__ recip, const int32_t* __restrict__ row_ptr,
                   int32_t* __restrict__ cursor, int32_t* __restrict__ tmp_col, double* __restrict__ tmp_w) {
  const int64_t e = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (e >= (int64_t)n * k) return;
  const uint8_t rc = recip[e];
  if (rc == RECIP_INVALID) return;
  const int i = (int)(e / k), slot = (int)(e - (int64_t)i * k);
  const int jid = knn_idx[e];
  const int j = id_map ? id_map[jid] : jid;
  double w = (double)dist[e];
  if (rc != RECIP_NONE) w = fmin(w, (double)dist[(int64_t)j * k + rc]);  // weight = min over directions
  int own_rank = 0;
  for (int s = 0; s < slot; ++s) own_rank += recip[(int64_t)i * k + s] != RECIP_INVALID;
  const int64_t o = (int64_t)row_ptr[i] + own_rank;
  tmp_col[o] = jid;
  tmp_w[o] = w;
  if (rc == RECIP_NONE && j >= 0 && j < n) {
    const int64_t r = (int64_t)row_ptr[j] + atomicAdd(&cursor[j], 1);
    tmp_col[r] = row_id ? row_id[i] : i;
    tmp_w[r] = w;
  }
}

// rows come out of the scatter in arrival order; emit them ascending by column
__global__ void __launch_bounds__(TPB)
sym_sort_rows_kernel(const int32_t* __restrict__ row_ptr, int n, const int32_t* __restrict__ tmp_col,
                     const double* __restrict__ tmp_w, int32_t* __restrict__ col, double* __restrict__ w64,
                     float* __restrict__ w32) {
  const int i = blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  const int64_t base = row_ptr[i];
  const int cnt = row_ptr[i + 1] - row_ptr[i];
  pg_sorted_chunk<SORT_CAP> buf;
  int emitted = 0, last = -1;
  while (emitted < cnt) {
    buf.reset(last);
    for (int t = 0; t < cnt; ++t) buf.push(tmp_col[base + t], tmp_w[base + t]);
    if (buf.m == 0) break;
    for (int t = 0; t < buf.m; ++t) {
      col[base + emitted + t] = buf.key[t];
      if (w64) w64[base + emitted + t] = buf.val[t];
      if (w32) w32[base + emitted + t] = (float)buf.val[t];
    }
    emitted += buf.m;
    last = buf.key[buf.m - 1];
  }
}

// rows are ascending by column: entries above the row's own id start at the first col > id
__global__ void __launch_bounds__(TPB)
upper_count_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                   const int32_t* __restrict__ row_id, int n, int32_t* __restrict__ up_count) {
  const int i = blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  const int id = row_id ? row_id[i] : i;
  int lo = row_ptr[i], hi = row_ptr[i + 1];
  const int end = hi;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (col[mid] > id) hi = mid; else lo = mid + 1;
  }
  up_count[i] = end - lo;
}

__global__ void __launch_bounds__(TPB)
upper_fill_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                  const double* __restrict__ w64, const float* __restrict__ w32,
                  const int32_t* __restrict__ row_id, const int32_t* __restrict__ up_ptr, int n,
                  long long* __restrict__ edges, double* __restrict__ ew64, float* __restrict__ ew32) {
  const int i = blockIdx.x * TPB + threadIdx.x;
  if (i >= n) return;
  const int id = row_id ? row_id[i] : i;
  const int cnt = up_ptr[i + 1] - up_ptr[i];
  const int64_t src = (int64_t)row_ptr[i + 1] - cnt, dst = up_ptr[i];
  for (int t = 0; t < cnt; ++t) {
    if (edges) { edges[2 * (dst + t)] = id; edges[2 * (dst + t) + 1] = col[src + t]; }
    if (ew64) ew64[dst + t] = w64 ? w64[src + t] : (double)w32[src + t];
    if (ew32) ew32[dst + t] = w32 ? w32[src + t] : (float)w64[src + t];
  }
}

// ---- K8 over a CSR ------------------------------------------------------------------------
constexpr int HIST_SMEM_MAX = 2048;

template <int TMAX>
__global__ void __launch_bounds__(TPB)
compose_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
               const int32_t* __restrict__ type, int n, int n_types, int32_t* __restrict__ nbr_count,
               int32_t* __restrict__ degree, pg_degree_stats* stats, int32_t* hist, int hist_len) {
  extern __shared__ int s_hist[];
  __shared__ int s_mn[TPB / 32], s_mx[TPB / 32], s_cnt[TPB / 32];
  __shared__ long long s_sum[TPB / 32], s_sq[TPB / 32];
  const bool use_smem_hist = hist != nullptr && hist_len <= HIST_SMEM_MAX;
  if (use_smem_hist) {
    for (int i = threadIdx.x; i < hist_len; i += TPB) s_hist[i] = 0;
    __syncthreads();
  }
  const int i = blockIdx.x * TPB + threadIdx.x;
  const bool valid = i < n;
  int deg = 0;
  if (valid) {
    const int b = row_ptr[i], e = row_ptr[i + 1];
    deg = e - b;
    if (nbr_count) {
      int tc[TMAX];
#pragma unroll
      for (int t = 0; t < TMAX; ++t) tc[t] = 0;
      for (int j = b; j < e; ++j) {
        const int ty = __ldg(&type[col[j]]);
#pragma unroll
        for (int t = 0; t < TMAX; ++t) tc[t] += (ty == t + 1);
      }
#pragma unroll
      for (int t = 0; t < TMAX; ++t)
        if (t < n_types) nbr_count[(int64_t)i * n_types + t] = tc[t];
    }
    if (degree) degree[i] = deg;
    if (hist) {
      const int bin = min(deg, hist_len - 1);
      if (use_smem_hist) atomicAdd(&s_hist[bin], 1); else atomicAdd(&hist[bin], 1);
    }
  }
  if (stats) {
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
  if (use_smem_hist) {
    __syncthreads();
    for (int b = threadIdx.x; b < hist_len; b += TPB) {
      const int c = s_hist[b];
      if (c) atomicAdd(&hist[b], c);
    }
  }
}

__global__ void init_stats_kernel(pg_degree_stats* stats) {
  stats->min_degree = 0x7fffffff; stats->max_degree = -1;
  stats->sum_degree = 0; stats->sumsq_degree = 0; stats->n_nodes = 0;
}
__global__ void finish_stats_kernel(pg_degree_stats* stats) {
  if (stats->n_nodes == 0) { stats->min_degree = 0; stats->max_degree = 0; }
}

// ---- K9 halo pack / unpack ---------------------------------------------------------------------
__global__ void __launch_bounds__(TPB)
halo_pack_kernel(const double2* __restrict__ xy, const int32_t* __restrict__ type, const int32_t* __restrict__ gid,
                 int n, double lo_edge, double hi_edge, pg_halo_rec* __restrict__ out, int capacity,
                 int32_t* count, int32_t* overflow) {
  const int i = blockIdx.x * TPB + threadIdx.x;
  bool take = false;
  double2 p = make_double2(0, 0);
  if (i < n) { p = xy[i]; take = p.x < lo_edge || p.x >= hi_edge; }
  // warp-aggregated append: one atomic per warp
  const unsigned m = __ballot_sync(0xffffffffu, take);
  if (m == 0) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == __ffs(m) - 1) base = atomicAdd(count, __popc(m));
  base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
  if (take) {
    const int o = base + __popc(m & ((1u << lane) - 1));
    if (o < capacity) {
      pg_halo_rec r;
      r.x = p.x; r.y = p.y; r.gid = gid ? gid[i] : i; r.type = type ? type[i] : 0;
      out[o] = r;
    } else {
      atomicExch(overflow, 1);
    }
  }
}

__global__ void __launch_bounds__(TPB)
halo_unpack_kernel(const pg_halo_rec* __restrict__ recs, int n_recs, int skip_begin, int skip_end, double x_lo,
                   double x_hi, double2* __restrict__ xy, int32_t* __restrict__ type, int32_t* __restrict__ gid,
                   int n_base, int capacity, int32_t* count, int32_t* overflow) {
  const int i = blockIdx.x * TPB + threadIdx.x;
  bool take = false;
  pg_halo_rec r;
  r.x = r.y = 0; r.gid = r.type = 0;
  if (i < n_recs && !(i >= skip_begin && i < skip_end)) {
    r = recs[i];
    take = r.x >= x_lo && r.x < x_hi;
  }
  const unsigned m = __ballot_sync(0xffffffffu, take);
  if (m == 0) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == __ffs(m) - 1) base = atomicAdd(count, __popc(m));
  base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
  if (take) {
    const int o = n_base + base + __popc(m & ((1u << lane) - 1));
    if (o < capacity) {
      xy[o] = make_double2(r.x, r.y);
      if (type) type[o] = r.type;
      if (gid) gid[o] = r.gid;
    } else {
      atomicExch(overflow, 1);
    }
  }
}

}  // namespace

extern "C" {

int pg_knn_symmetrize_count(pg_handle* h, int32_t n, int32_t k, const int32_t* knn_idx, const int32_t* row_id,
                            const int32_t* id_map, int32_t n_ids, int32_t* und_row_ptr, pg_stream stream) {
  if (!h) return PG_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  PG_CUDA(h, cudaSetDevice(h->device));
  h->last_stream = s;
  PG_REQUIRE(h, n >= 0 && k >= 1 && k <= PG_MAX_K, "pg_knn_symmetrize_count: need n >= 0 and 1 <= k <= %d", PG_MAX_K);
  PG_REQUIRE(h, und_row_ptr != nullptr && (n == 0 || knn_idx != nullptr), "pg_knn_symmetrize_count: NULL argument");
  PG_REQUIRE(h, (int64_t)n * k < (int64_t)1 << 31, "pg_knn_symmetrize_count: n*k must be < 2^31");
  PG_REQUIRE(h, (row_id == nullptr) == (id_map == nullptr) && (!id_map || n_ids > 0),
             "pg_knn_symmetrize_count: row_id and id_map go together (both NULL = identity)");
  int rc;
  if ((rc = pg_reserve(h, h->sym_extra, ((size_t)n + 4) * sizeof(int32_t)))) return rc;
  if ((rc = pg_reserve(h, h->sym_cursor, ((size_t)n + 4) * sizeof(int32_t)))) return rc;
  if ((rc = pg_reserve(h, h->sym_recip, (size_t)n * k + 16))) return rc;
  if ((rc = pg_reserve(h, h->row_count, ((size_t)n + 4) * sizeof(int32_t)))) return rc;
  if (n > 0) {
    PG_CUDA(h, cudaMemsetAsync(h->sym_extra.p, 0, (size_t)n * sizeof(int32_t), s));
    PG_LAUNCH(h, s, "sym_mark_kernel", sym_mark_kernel<<<pg_div_up((int64_t)n * k, TPB), TPB, 0, s>>>(knn_idx, n, k, row_id, id_map, n_ids, (uint8_t*)h->sym_recip.p,
                                                                  (int32_t*)h->sym_extra.p));
    PG_LAUNCH(h, s, "sym_degree_kernel", sym_degree_kernel<<<pg_div_up(n, TPB), TPB, 0, s>>>((const uint8_t*)h->sym_recip.p, (const int32_t*)h->sym_extra.p,
                                                       n, k, (int32_t*)h->sym_cursor.p, (int32_t*)h->row_count.p));
    PG_LAUNCH_CHECK(h);
  }
  return pg_scan_i32(h, (const int32_t*)h->row_count.p, und_row_ptr, n, s, (int32_t*)((char*)h->misc.p + PG_MISC_TOTALS) + 1);
}

int pg_knn_symmetrize_total(pg_handle* h, int64_t* total) {
  if (!h || !total) return PG_ERR_INVALID;
  PG_CUDA(h, cudaSetDevice(h->device));
  PG_CUDA(h, cudaMemcpyAsync(&h->pinned[1], (int32_t*)((char*)h->misc.p + PG_MISC_TOTALS) + 1, sizeof(int32_t),
                             cudaMemcpyDeviceToHost, h->last_stream));
  PG_CUDA(h, cudaStreamSynchronize(h->last_stream));
  *total = h->pinned[1];
  return PG_OK;
}

int pg_knn_symmetrize_fill(pg_handle* h, int32_t n, int32_t k, const int32_t* knn_idx, const double* dist64,
                           const float* dist32, const int32_t* row_id, const int32_t* id_map,
                           const int32_t* und_row_ptr, int32_t* und_col, double* und_w64, float* und_w32,
                           pg_stream stream) {
  if (!h) return PG_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  PG_CUDA(h, cudaSetDevice(h->device));
  h->last_stream = s;
  PG_REQUIRE(h, n >= 0 && k >= 1 && k <= PG_MAX_K, "pg_knn_symmetrize_fill: bad n / k");
  PG_REQUIRE(h, dist64 || dist32, "pg_knn_symmetrize_fill: one of dist64 / dist32 is required");
  PG_REQUIRE(h, und_row_ptr && (n == 0 || (knn_idx && und_col)), "pg_knn_symmetrize_fill: NULL argument");
  if (n == 0) return PG_OK;
  // the total of the matching count pass sizes the staging rows
  int64_t total = 0;
  int rc = pg_knn_symmetrize_total(h, &total);
  if (rc) return rc;
  // staging (arrival-order rows) lives in the grid's scratch that is free at this point
  pg_buf& tcol = h->cell_of;
  pg_buf& tw = h->rank;
  if ((rc = pg_reserve(h, tcol, ((size_t)total + 4) * sizeof(int32_t)))) return rc;
  if ((rc = pg_reserve(h, tw, ((size_t)total + 4) * sizeof(double)))) return rc;
  const int blocks_e = pg_div_up((int64_t)n * k, TPB);
  if (dist64)
    PG_LAUNCH(h, s, "sym_scatter_kernel<double>", sym_scatter_kernel<double><<<blocks_e, TPB, 0, s>>>(knn_idx, dist64, n, k, row_id, id_map, (const uint8_t*)h->sym_recip.p, und_row_ptr,
                                                       (int32_t*)h->sym_cursor.p, (int32_t*)tcol.p, (double*)tw.p));
  else
    PG_LAUNCH(h, s, "sym_scatter_kernel<float>", sym_scatter_kernel<float><<<blocks_e, TPB, 0, s>>>(knn_idx, dist32, n, k, row_id, id_map, (const uint8_t*)h->sym_recip.p, und_row_ptr,
                                                      (int32_t*)h->sym_cursor.p, (int32_t*)tcol.p, (double*)tw.p));
  PG_LAUNCH(h, s, "sym_sort_rows_kernel", sym_sort_rows_kernel<<<pg_div_up(n, TPB), TPB, 0, s>>>(und_row_ptr, n, (const int32_t*)tcol.p, (const double*)tw.p,
                                                        und_col, und_w64, und_w32));
  PG_LAUNCH_CHECK(h);
  return PG_OK;
}

int pg_csr_upper_count(pg_handle* h, int32_t n, const int32_t* row_ptr, const int32_t* col,
                       const int32_t* row_id, int32_t* up_ptr, pg_stream stream) {
  if (!h) return PG_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  PG_CUDA(h, cudaSetDevice(h->device));
  h->last_stream = s;
  PG_REQUIRE(h, n >= 0 && row_ptr && up_ptr, "pg_csr_upper_count: bad argument");
  int rc;
  if ((rc = pg_reserve(h, h->row_count, ((size_t)n + 4) * sizeof(int32_t)))) return rc;
  if (n > 0) {
    PG_LAUNCH(h, s, "upper_count_kernel", upper_count_kernel<<<pg_div_up(n, TPB), TPB, 0, s>>>(row_ptr, col, row_id, n, (int32_t*)h->row_count.p));
    PG_LAUNCH_CHECK(h);
  }
  return pg_scan_i32(h, (const int32_t*)h->row_count.p, up_ptr, n, s, (int32_t*)((char*)h->misc.p + PG_MISC_TOTALS) + 2);
}

int pg_csr_upper_total(pg_handle* h, int64_t* total) {
  if (!h || !total) return PG_ERR_INVALID;
  PG_CUDA(h, cudaSetDevice(h->device));
  PG_CUDA(h, cudaMemcpyAsync(&h->pinned[2], (int32_t*)((char*)h->misc.p + PG_MISC_TOTALS) + 2, sizeof(int32_t),
                             cudaMemcpyDeviceToHost, h->last_stream));
  PG_CUDA(h, cudaStreamSynchronize(h->last_stream));
  *total = h->pinned[2];
  return PG_OK;
}

int pg_csr_upper_fill(pg_handle* h, int32_t n, const int32_t* row_ptr, const int32_t* col, const double* w64,
                      const float* w32, const int32_t* row_id, const int32_t* up_ptr, int64_t* edges_i64,
                      double* ew64, float* ew32, pg_stream stream) {
  if (!h) return PG_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  PG_CUDA(h, cudaSetDevice(h->device));
  h->last_stream = s;
  PG_REQUIRE(h, n >= 0 && row_ptr && up_ptr, "pg_csr_upper_fill: bad argument");
  PG_REQUIRE(h, !(ew64 || ew32) || (w64 || w32), "pg_csr_upper_fill: weights requested but none given");
  if (n == 0) return PG_OK;
  PG_LAUNCH(h, s, "upper_fill_kernel", upper_fill_kernel<<<pg_div_up(n, TPB), TPB, 0, s>>>(row_ptr, col, w64, w32, row_id, up_ptr, n, (long long*)edges_i64,
                                                     ew64, ew32));
  PG_LAUNCH_CHECK(h);
  return PG_OK;
}

int pg_compose_degree(pg_handle* h, int32_t n, const int32_t* row_ptr, const int32_t* col, const int32_t* type,
                      int32_t n_types, int32_t* nbr_count, int32_t* degree, pg_degree_stats* stats, int32_t* hist,
                      int32_t hist_len, pg_stream stream) {
  if (!h) return PG_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  PG_CUDA(h, cudaSetDevice(h->device));
  h->last_stream = s;
  PG_REQUIRE(h, n >= 0 && row_ptr, "pg_compose_degree: bad argument");
  PG_REQUIRE(h, !nbr_count || (type && col && n_types >= 1 && n_types <= PG_MAX_TYPES),
             "pg_compose_degree: nbr_count needs type, col and 1 <= n_types <= %d", PG_MAX_TYPES);
  PG_REQUIRE(h, !hist || hist_len >= 1, "pg_compose_degree: hist_len must be >= 1");
  if (stats) PG_LAUNCH(h, s, "init_stats_kernel", init_stats_kernel<<<1, 1, 0, s>>>(stats));
  if (hist) PG_CUDA(h, cudaMemsetAsync(hist, 0, (size_t)hist_len * sizeof(int32_t), s));
  if (n > 0) {
    const size_t smem = (hist && hist_len <= HIST_SMEM_MAX) ? (size_t)hist_len * sizeof(int) : 0;
    if (!nbr_count || n_types <= 8)
      PG_LAUNCH(h, s, "compose_kernel<8>", compose_kernel<8><<<pg_div_up(n, TPB), TPB, smem, s>>>(row_ptr, col, type, n, n_types, nbr_count, degree, stats, hist, hist_len));
    else
      PG_LAUNCH(h, s, "compose_kernel<16>", compose_kernel<16><<<pg_div_up(n, TPB), TPB, smem, s>>>(row_ptr, col, type, n, n_types, nbr_count, degree, stats, hist, hist_len));
    PG_LAUNCH_CHECK(h);
  }
  if (stats) PG_LAUNCH(h, s, "finish_stats_kernel", finish_stats_kernel<<<1, 1, 0, s>>>(stats));
  PG_LAUNCH_CHECK(h);
  return PG_OK;
}

int pg_halo_pack(pg_handle* h, int32_t n, const double* xy, const int32_t* type, const int32_t* gid,
                 double lo_edge, double hi_edge, pg_halo_rec* out, int32_t capacity, int32_t* count_out,
                 pg_stream stream) {
  if (!h) return PG_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  PG_CUDA(h, cudaSetDevice(h->device));
  h->last_stream = s;
  PG_REQUIRE(h, n >= 0 && capacity >= 0 && count_out && (capacity == 0 || out), "pg_halo_pack: bad argument");
  PG_CUDA(h, cudaMemsetAsync(count_out, 0, sizeof(int32_t), s));
  if (n > 0) {
    PG_LAUNCH(h, s, "halo_pack_kernel", halo_pack_kernel<<<pg_div_up(n, TPB), TPB, 0, s>>>((const double2*)xy, type, gid, n, lo_edge, hi_edge, out, capacity,
                                                      count_out, (int32_t*)((char*)h->misc.p + PG_MISC_OVERFLOW)));
    PG_LAUNCH_CHECK(h);
  }
  return PG_OK;
}

int pg_halo_unpack(pg_handle* h, const pg_halo_rec* recs, int32_t n_recs, int32_t skip_begin, int32_t skip_end,
                   double x_lo, double x_hi, double* xy, int32_t* type, int32_t* gid, int32_t n_base,
                   int32_t capacity, int32_t* count_out, pg_stream stream) {
  if (!h) return PG_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  PG_CUDA(h, cudaSetDevice(h->device));
  h->last_stream = s;
  PG_REQUIRE(h, n_recs >= 0 && n_base >= 0 && capacity >= n_base && count_out && xy, "pg_halo_unpack: bad argument");
  PG_CUDA(h, cudaMemsetAsync(count_out, 0, sizeof(int32_t), s));
  if (n_recs > 0) {
    PG_LAUNCH(h, s, "halo_unpack_kernel", halo_unpack_kernel<<<pg_div_up(n_recs, TPB), TPB, 0, s>>>(recs, n_recs, skip_begin, skip_end, x_lo, x_hi, (double2*)xy,
                                                             type, gid, n_base, capacity, count_out,
                                                             (int32_t*)((char*)h->misc.p + PG_MISC_OVERFLOW)));
    PG_LAUNCH_CHECK(h);
  }
  return PG_OK;
}

}  // extern "C"
