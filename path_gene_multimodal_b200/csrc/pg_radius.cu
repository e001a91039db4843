// K6 radius graph (two-pass count / scan / fill CSR) with K8 (neighbour-type histogram + degree
// statistics) fused into the count pass.
// Reference: /root/reference/hovernet_tile_inference.ipynb:2964-2975 (cKDTree.query_ball_tree, i<j edge
// loop), :3021 (edge_index), :3041-3042 (np.linalg.norm distances, float32 edge_attr); composition /
// degree per SURVEY A.5. Acceptance test is d2 <= r*r in float64 with d2 = fl(fl(dx*dx)+fl(dy*dy)).
//
// Work split: one CTA per run of W consecutive cells of a grid row. The three cell rows around the
// run (W+2 cells each, three contiguous runs of the cell-ordered point array) are staged in shared
// memory with coalesced loads; then each thread takes one query point of the run and walks its own
// 3x3 block out of shared memory in ONE merged, branch-free loop (accept flag folded into the counters,
// per-type counts in 12-bit fields of one 64-bit word). A run whose neighbourhood does not fit the
// staging buffer, a ring radius > 1 or more than 5 types fall back to walking global memory.
#include <cmath>
#include "pg_query.cuh"

namespace {

constexpr int TPB = 128;
constexpr int FILL_CAP = 32;
constexpr int HIST_SMEM_MAX = 512;
constexpr int STAGE_CAP = 1024;  // candidates staged per CTA (24 B each)
constexpr int MAX_W = 256;       // query cells per CTA
constexpr int TYPE_BITS = 12;    // packed per-type counters; STAGE_CAP < 2^12 so a field cannot overflow
constexpr int PACKED_TYPES = 5;

// per-thread degree statistics -> one set of atomics per CTA
__device__ __forceinline__ void reduce_degree_stats(int mn, int mx, long long sum, long long sq, int cnt,
                                                    pg_degree_stats* stats) {
  __shared__ int s_mn[TPB / 32], s_mx[TPB / 32], s_cnt[TPB / 32];
  __shared__ long long s_sum[TPB / 32], s_sq[TPB / 32];
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

// stats block + histogram cleared by one tiny launch (replaces memset + init)
__global__ void prep_stats_kernel(pg_degree_stats* stats, int32_t* hist, int hist_len, int empty) {
  if (stats && threadIdx.x == 0) {
    stats->min_degree = empty ? 0 : 0x7fffffff;
    stats->max_degree = empty ? 0 : -1;
    stats->sum_degree = 0; stats->sumsq_degree = 0; stats->n_nodes = 0;
  }
  if (hist)
    for (int i = threadIdx.x; i < hist_len; i += blockDim.x) hist[i] = 0;
}

// Geometry of one CTA's run of cells and its staged neighbourhood.
struct run_geom {
  int y, c0, c1, xa, ncell;  // query cells [c0, c1) of row y; candidate cells [xa, xa + ncell)
  int q0, q1;                // query points (cell-order positions)
  int base[3], first[3];     // shared-memory base and first global position of each staged row
  int total;
};

struct stage_smem {
  double2 xy[STAGE_CAP];
  int2 ia[STAGE_CAP];            // {id, aux}
  int cs[3][MAX_W + 4];          // cell_start of the three rows over the candidate cells (+1)
};

// Loads the cell_start slices, derives the geometry and (when it fits) stages the candidates.
// aux_mode 0: aux = bit shift of the packed type counter; 1: aux unused.
__device__ __forceinline__ bool stage_run(const pg_grid_view& g, int W, int nbx, int n_types, int aux_mode,
                                          stage_smem& sm, run_geom& rg) {
  const int tid = threadIdx.x;
  rg.y = blockIdx.x / nbx;
  rg.c0 = (blockIdx.x - rg.y * nbx) * W;
  rg.c1 = min(rg.c0 + W, g.nx);
  rg.xa = max(rg.c0 - 1, 0);
  const int xb = min(rg.c1, g.nx - 1);
  rg.ncell = xb - rg.xa + 1;
  const int stride = rg.ncell + 1;
  for (int i = tid; i < 3 * stride; i += TPB) {
    const int r = i / stride, c = i - r * stride;
    const int yy = rg.y - 1 + r;
    sm.cs[r][c] = (yy >= 0 && yy < g.ny) ? g.cell_start[yy * g.nx + rg.xa + c] : 0;
  }
  __syncthreads();
  rg.q0 = sm.cs[1][rg.c0 - rg.xa];
  rg.q1 = sm.cs[1][rg.c1 - rg.xa];
  int acc = 0;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    rg.first[r] = sm.cs[r][0];
    rg.base[r] = acc;
    acc += sm.cs[r][rg.ncell] - sm.cs[r][0];
  }
  rg.total = acc;
  if (rg.q1 == rg.q0 || acc > STAGE_CAP) return false;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int cnt = sm.cs[r][rg.ncell] - rg.first[r];
    for (int t = tid; t < cnt; t += TPB) {
      const int j = rg.first[r] + t;
      const int2 m = g.s_meta[j];
      sm.xy[rg.base[r] + t] = g.s_xy[j];
      int aux = 0;
      if (aux_mode == 0) aux = (m.y >= 1 && m.y <= n_types) ? (m.y - 1) * TYPE_BITS : 60;
      sm.ia[rg.base[r] + t] = make_int2(pg_id_of(g, j, m.x), aux);
    }
  }
  __syncthreads();
  return true;
}

// shared-memory sub-ranges of the 3x3 block of cell cx inside the staged rows
__device__ __forceinline__ void block_ranges(const stage_smem& sm, const run_geom& rg, int cx, int nx,
                                             int& s0, int& n0, int& s1, int& n1, int& s2, int& n2) {
  const int i0 = max(cx - 1, rg.xa) - rg.xa, i1 = min(cx + 1, nx - 1) - rg.xa + 1;
  const int a0 = sm.cs[0][i0], a1 = sm.cs[1][i0], a2 = sm.cs[2][i0];
  n0 = sm.cs[0][i1] - a0; n1 = sm.cs[1][i1] - a1; n2 = sm.cs[2][i1] - a2;
  s0 = rg.base[0] + a0 - rg.first[0];
  s1 = rg.base[1] + a1 - rg.first[1];
  s2 = rg.base[2] + a2 - rg.first[2];
}

// Count pass: CSR row count (all neighbours, or only id_j > id_i when `upper`) and, fused over all
// neighbours: degree, per-type neighbour counts, degree statistics and histogram.
__global__ void __launch_bounds__(TPB)
radius_count_kernel(pg_grid_view g, double r2, int R, int W, int nbx, int upper, int32_t* __restrict__ row_count,
                    int32_t* __restrict__ degree, int32_t* __restrict__ nbr_count, int n_types,
                    pg_degree_stats* stats, int32_t* hist, int hist_len) {
  __shared__ stage_smem sm;
  __shared__ int s_hist[HIST_SMEM_MAX];
  const bool use_smem_hist = hist != nullptr && hist_len <= HIST_SMEM_MAX;
  if (use_smem_hist)
    for (int i = threadIdx.x; i < hist_len; i += TPB) s_hist[i] = 0;
  run_geom rg;
  const bool fast = R == 1 && n_types <= PACKED_TYPES;
  const bool staged = stage_run(g, W, nbx, n_types, 0, sm, rg) && fast;   // block-uniform
  if (rg.q1 == rg.q0) return;                                             // block-uniform: empty run

  int st_mn = 0x7fffffff, st_mx = -1, st_cnt = 0;
  long long st_sum = 0, st_sq = 0;
  for (int q = rg.q0 + threadIdx.x; q < rg.q1; q += TPB) {
    const int2 me = g.s_meta[q];
    if (me.x >= g.n_query) continue;  // halo point: no row
    const int my_id = pg_id_of(g, q, me.x);
    int deg = 0, up = 0;
    if (staged) {
      const int slot = rg.base[1] + q - rg.first[1];
      const double2 p = sm.xy[slot];
      const int cx = pg_cell_coord(p.x, g.x0, g.inv_cell, g.nx);
      int s0, n0, s1, n1, s2, n2;
      block_ranges(sm, rg, cx, g.nx, s0, n0, s1, n1, s2, n2);
      const int n01 = n0 + n1, tot = n01 + n2;
      const int off1 = s1 - n0, off2 = s2 - n01;
      unsigned long long pk = 0;
      for (int t = 0; t < tot; ++t) {
        const int idx = t + (t < n0 ? s0 : (t < n01 ? off1 : off2));
        const double2 c = sm.xy[idx];
        const int2 ia = sm.ia[idx];
        const double d2 = pg_dist2(p.x, p.y, c.x, c.y);
        const int acc = (d2 <= r2) & (idx != slot);
        deg += acc;
        up += acc & (ia.x > my_id);
        pk += (unsigned long long)acc << ia.y;
      }
      if (nbr_count) {
#pragma unroll
        for (int t = 0; t < PACKED_TYPES; ++t)
          if (t < n_types) nbr_count[(int64_t)me.x * n_types + t] = (int)((pk >> (TYPE_BITS * t)) & ((1u << TYPE_BITS) - 1));
      }
    } else {
      const double2 p = g.s_xy[q];
      const int cx = pg_cell_coord(p.x, g.x0, g.inv_cell, g.nx);
      int tc[PG_MAX_TYPES];
#pragma unroll
      for (int t = 0; t < PG_MAX_TYPES; ++t) tc[t] = 0;
      pg_visit_block(g, cx, rg.y, R, [&](int b, int e) {
        for (int j = b; j < e; ++j) {
          const double2 c = g.s_xy[j];
          const double d2 = pg_dist2(p.x, p.y, c.x, c.y);
          if (d2 <= r2 && j != q) {
            const int2 m = g.s_meta[j];
            ++deg;
            up += (pg_id_of(g, j, m.x) > my_id);
#pragma unroll
            for (int t = 0; t < PG_MAX_TYPES; ++t) tc[t] += (m.y == t + 1);
          }
        }
      });
      if (nbr_count) {
#pragma unroll
        for (int t = 0; t < PG_MAX_TYPES; ++t)
          if (t < n_types) nbr_count[(int64_t)me.x * n_types + t] = tc[t];
      }
    }
    row_count[me.x] = upper ? up : deg;
    if (degree) degree[me.x] = deg;
    if (hist) {
      const int bin = min(deg, hist_len - 1);
      if (use_smem_hist) atomicAdd(&s_hist[bin], 1); else atomicAdd(&hist[bin], 1);
    }
    st_mn = min(st_mn, deg); st_mx = max(st_mx, deg); st_sum += deg; st_sq += (long long)deg * deg; ++st_cnt;
  }
  if (stats) reduce_degree_stats(st_mn, st_mx, st_sum, st_sq, st_cnt, stats);
  if (use_smem_hist) {
    __syncthreads();
    for (int i = threadIdx.x; i < hist_len; i += TPB) {
      const int c = s_hist[i];
      if (c) atomicAdd(&hist[i], c);
    }
  }
}

// Fill pass. Same staging; a row is emitted in ascending column (id) order in chunks of FILL_CAP:
// each pass keeps the FILL_CAP smallest accepted ids above the last one written, so ordinary rows
// take one walk and heavy rows take ceil(count / FILL_CAP) walks without extra memory.
__global__ void __launch_bounds__(TPB)
radius_fill_kernel(pg_grid_view g, double r2, int R, int W, int nbx, int upper, const int32_t* __restrict__ row_ptr,
                   int32_t* __restrict__ col, float* __restrict__ dist32, double* __restrict__ dist64,
                   long long* __restrict__ edges, long long* __restrict__ edge_index,
                   float* __restrict__ edge_attr, long long n_edges, long long capacity, int32_t* overflow) {
  __shared__ stage_smem sm;
  run_geom rg;
  const bool staged = stage_run(g, W, nbx, 0, 1, sm, rg) && R == 1;
  if (rg.q1 == rg.q0) return;
  for (int q = rg.q0 + threadIdx.x; q < rg.q1; q += TPB) {
    const int2 me = g.s_meta[q];
    if (me.x >= g.n_query) continue;
    const int base = row_ptr[me.x];
    const int cnt = row_ptr[me.x + 1] - base;
    if (cnt <= 0) continue;
    if ((long long)base + cnt > capacity) { atomicExch(overflow, 1); continue; }
    const int my_id = pg_id_of(g, q, me.x);
    const double2 p = g.s_xy[q];
    const int cx = pg_cell_coord(p.x, g.x0, g.inv_cell, g.nx);
    int s0 = 0, n0 = 0, s1 = 0, n1 = 0, s2 = 0, n2 = 0, slot = 0;
    if (staged) {
      slot = rg.base[1] + q - rg.first[1];
      block_ranges(sm, rg, cx, g.nx, s0, n0, s1, n1, s2, n2);
    }
    pg_sorted_chunk<FILL_CAP> buf;
    int emitted = 0;
    int last = upper ? my_id : -1;
    while (emitted < cnt) {
      buf.reset(last);
      if (staged) {
        const int n01 = n0 + n1, tot = n01 + n2;
        const int off1 = s1 - n0, off2 = s2 - n01;
        for (int t = 0; t < tot; ++t) {
          const int idx = t + (t < n0 ? s0 : (t < n01 ? off1 : off2));
          const double2 c = sm.xy[idx];
          const double d2 = pg_dist2(p.x, p.y, c.x, c.y);
          if (d2 <= r2 && idx != slot) buf.push(sm.ia[idx].x, d2);
        }
      } else {
        pg_visit_block(g, cx, rg.y, R, [&](int b, int e) {
          for (int j = b; j < e; ++j) {
            const double2 c = g.s_xy[j];
            const double d2 = pg_dist2(p.x, p.y, c.x, c.y);
            if (d2 <= r2 && j != q) buf.push(pg_id_of(g, j, g.s_meta[j].x), d2);
          }
        });
      }
      if (buf.m == 0) break;  // cannot happen when row_ptr came from the matching count pass
      for (int t = 0; t < buf.m; ++t) {
        const long long o = (long long)base + emitted + t;
        const double d = sqrt(buf.val[t]);
        col[o] = buf.key[t];
        if (dist32) dist32[o] = (float)d;
        if (dist64) dist64[o] = d;
        if (edges) { edges[2 * o] = my_id; edges[2 * o + 1] = buf.key[t]; }
        if (edge_index) {  // [2, 2E] = hstack(edges.T, edges[:, ::-1].T), ipynb:3021 (see SURVEY B-3)
          edge_index[o] = my_id; edge_index[n_edges + o] = buf.key[t];
          edge_index[2 * n_edges + o] = buf.key[t]; edge_index[3 * n_edges + o] = my_id;
        }
        if (edge_attr) { edge_attr[o] = (float)d; edge_attr[n_edges + o] = (float)d; }  // ipynb:3041-3042
      }
      emitted += buf.m;
      last = buf.key[buf.m - 1];
    }
  }
}

// cells per CTA: about one query point per thread, bounded by the staging buffer
static inline int pick_run_width(const pg_grid& gr) {
  const double cells = (double)gr.nx * (double)gr.ny;
  const double lambda = gr.n > 0 ? (double)gr.n / cells : 1.0;   // points per cell
  int w = (int)std::floor(TPB / std::max(lambda, 1e-3));
  const int by_stage = (int)std::floor(STAGE_CAP / (3.0 * 2.5 * std::max(lambda, 1e-3))) - 2;  // 2.5x headroom
  w = std::min(w, by_stage);
  w = std::max(1, std::min(w, MAX_W));
  return std::min(w, std::max(gr.nx, 1));
}

static inline int ring_radius(double r, const pg_grid& gr) {
  // every point within r lies at most R cells away along each axis (slack far above the rounding
  // of the cell coordinate)
  int R = (int)std::ceil(r * gr.inv_cell * (1.0 + 1e-9) + 1e-9);
  return R < 1 ? 1 : R;
}

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
  h->radius_r = r;
  h->radius_flags = flags;
  if (stats || hist)
    PG_LAUNCH(h, s, "prep_stats_kernel", prep_stats_kernel<<<1, 256, 0, s>>>(stats, hist, hist ? hist_len : 0, nq == 0));
  if (gr.n > 0 && nq > 0) {
    pg_grid_view v = pg_make_view(h);
    const int W = pick_run_width(gr);
    const int nbx = pg_div_up(gr.nx, W);
    PG_LAUNCH(h, s, "radius_count_kernel", radius_count_kernel<<<nbx * gr.ny, TPB, 0, s>>>(
        v, r * r, ring_radius(r, gr), W, nbx, flags == PG_RADIUS_UPPER, (int32_t*)h->row_count.p, degree, nbr_count,
        nbr_count ? n_types : 1, stats, hist, hist_len));
    PG_LAUNCH_CHECK(h);
  }
  return pg_scan_i32(h, (const int32_t*)h->row_count.p, row_ptr, nq, s, (int32_t*)((char*)h->misc.p + PG_MISC_TOTALS));
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
  if (gr.n == 0 || gr.n_query == 0) return PG_OK;
  pg_grid_view v = pg_make_view(h);
  const int W = pick_run_width(gr);
  const int nbx = pg_div_up(gr.nx, W);
  PG_LAUNCH(h, s, "radius_fill_kernel", radius_fill_kernel<<<nbx * gr.ny, TPB, 0, s>>>(
      v, h->radius_r * h->radius_r, ring_radius(h->radius_r, gr), W, nbx, h->radius_flags == PG_RADIUS_UPPER, row_ptr, col,
      dist32, dist64, (long long*)edges_i64, (long long*)edge_index, edge_attr, (long long)n_edges, (long long)capacity,
      (int32_t*)((char*)h->misc.p + PG_MISC_OVERFLOW)));
  PG_LAUNCH_CHECK(h);
  return PG_OK;
}

}  // extern "C"
