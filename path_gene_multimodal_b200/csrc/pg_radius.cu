// K6 radius graph (two-pass count / scan / fill CSR) with K8 (neighbour-type histogram + degree
// statistics) fused into the count pass.
// Reference: /root/reference/hovernet_tile_inference.ipynb:2964-2975 (cKDTree.query_ball_tree, i<j edge
// loop), :3021 (edge_index), :3041-3042 (np.linalg.norm distances, float32 edge_attr); composition /
// degree per SURVEY A.5. Acceptance test is d2 <= r*r in float64 with d2 = fl(fl(dx*dx)+fl(dy*dy)).
//
// Work split: one thread per point, in cell order, so the 32 lanes of a warp sit in the same or adjacent
// cells and their candidate records (one 32-byte sector each, read with one 256-bit load) hit in L1.
// With the usual cell ~ r the 3x3 block of a point is three contiguous runs of the cell-ordered array;
// they are walked as ONE merged loop (a branch-free index map) so a warp iterates max-over-lanes of the
// block population once instead of three times. The count pass keeps the per-type neighbour counts in
// 12-bit fields of one 64-bit register (the field offset is precomputed in the record); degree statistics
// are reduced per warp, then per CTA (no barrier), then added to accumulators in the handle which the row_ptr
// scan that follows publishes and resets - no init / finish launches. The fill pass collects a row's accepted (id, position) pairs
// in shared memory (slot-major, conflict-free), sorts the handful of entries by id and writes the row.
#include <cmath>
#include "pg_query.cuh"

namespace {

constexpr int TPB_COUNT = 256;
constexpr int TPB_FILL = 128;
constexpr int FILL_CAP = 16;         // row entries kept per thread in shared memory (2 x 4 B x CAP x TPB_FILL = 16 KB)
constexpr int FIELD_MAX = (1 << PG_TYPE_BITS) - 1;

// Calls f(position, record) for every candidate of the (2R+1)^2 block around (cx, cy) and flush() at least
// once every FIELD_MAX candidates (and once at the end). Records are fetched four at a time so that four
// 256-bit loads are in flight per thread; the slots past the end of a run are pointed at `self`, the
// caller's own position, which every f rejects anyway (a point is not its own neighbour).
template <bool MERGED, class F, class FL>
__device__ __forceinline__ void walk_block(const pg_grid_view& g, int R, int cx, int cy, int self, F&& f, FL&& flush) {
  if (MERGED) {  // R == 1: three runs, one loop
    const int xa = max(cx - 1, 0), xe = min(cx + 1, g.nx - 1) + 1;
    const int32_t* c1 = g.cell_start + (int64_t)cy * g.nx;
    const int b1 = c1[xa], e1 = c1[xe];
    int b0 = 0, e0 = 0, b2 = 0, e2 = 0;
    if (cy > 0) { b0 = c1[xa - g.nx]; e0 = c1[xe - g.nx]; }
    if (cy + 1 < g.ny) { b2 = c1[xa + g.nx]; e2 = c1[xe + g.nx]; }
    const int n0 = e0 - b0, n01 = n0 + (e1 - b1), tot = n01 + (e2 - b2);
    const int off1 = b1 - n0, off2 = b2 - n01;
    auto pos = [&](int t) { return t < tot ? t + (t < n0 ? b0 : (t < n01 ? off1 : off2)) : self; };
    for (int t0 = 0; t0 < tot; t0 += FIELD_MAX - 3) {  // FIELD_MAX - 3 is a multiple of 4
      const int t1 = min(tot, t0 + FIELD_MAX - 3);
      for (int t = t0; t < t1; t += 4) {
        const int j0 = pos(t), j1 = pos(t + 1), j2 = pos(t + 2), j3 = pos(t + 3);
        const pg_rec r0 = pg_ld_rec(g.rec + j0), r1 = pg_ld_rec(g.rec + j1);
        const pg_rec r2 = pg_ld_rec(g.rec + j2), r3 = pg_ld_rec(g.rec + j3);
        f(j0, r0); f(j1, r1); f(j2, r2); f(j3, r3);
      }
      flush();
    }
  } else {
    pg_visit_block(g, cx, cy, R, [&](int b, int e) {
      for (int j0 = b; j0 < e; j0 += FIELD_MAX - 3) {
        const int j1 = min(e, j0 + FIELD_MAX - 3);
        for (int j = j0; j < j1; j += 4) {
          const int p0 = j, p1 = j + 1 < j1 ? j + 1 : self, p2 = j + 2 < j1 ? j + 2 : self, p3 = j + 3 < j1 ? j + 3 : self;
          const pg_rec r0 = pg_ld_rec(g.rec + p0), r1 = pg_ld_rec(g.rec + p1);
          const pg_rec r2 = pg_ld_rec(g.rec + p2), r3 = pg_ld_rec(g.rec + p3);
          f(p0, r0); f(p1, r1); f(p2, r2); f(p3, r3);
        }
        flush();
      }
    });
  }
}

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// Count pass: CSR row count (all neighbours, or only id_j > id_i when `upper`) and, fused over all
// neighbours: degree, per-type neighbour counts, degree statistics and histogram.
// hist_mode 0: no histogram; 1: shared-memory bins -> accumulators -> published by the scan that follows;
// 2: hist_len > PG_ACC_HIST_MAX, bins added straight into the caller's (pre-zeroed) array.
template <bool MERGED, bool WIDE_TYPES>
__global__ void __launch_bounds__(TPB_COUNT)
radius_count_kernel(pg_grid_view g, double r2, int R, int upper, int32_t* __restrict__ row_count,
                    int32_t* __restrict__ degree, int32_t* __restrict__ nbr_count, int n_types,
                    pg_stats_acc* acc, int32_t* acc_hist, bool stats, int32_t* hist, int hist_len, int hist_mode) {
  __shared__ int s_hist[PG_ACC_HIST_MAX];
  __shared__ int s_mn, s_mx, s_cnt, s_arrived;
  __shared__ unsigned long long s_sum, s_sq;
  const int tid = threadIdx.x;
  if (hist_mode == 1)
    for (int i = tid; i < hist_len; i += TPB_COUNT) s_hist[i] = 0;
  if (tid == 0) { s_mn = 0x7fffffff; s_mx = -1; s_cnt = 0; s_sum = 0; s_sq = 0; s_arrived = 0; }
  __syncthreads();

  const int q = blockIdx.x * TPB_COUNT + tid;
  const pg_rec me = pg_ld_rec(g.rec + min(q, g.n - 1));
  const bool active = q < g.n && me.row < g.n_query;  // halo points own no row
  int deg = 0, up = 0;
  if (active) {
    const int cx = pg_cell_coord(me.x, g.x0, g.inv_cell, g.nx);
    const int cy = pg_cell_coord(me.y, g.y0, g.inv_cell, g.ny);
    unsigned long long pk = 0;
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
    walk_block<MERGED>(g, R, cx, cy, q,
      [&](int j, const pg_rec& c) {
        const double d2 = pg_dist2(me.x, me.y, c.x, c.y);
        const int a = (d2 <= r2) & (j != q);
        deg += a;
        up += a & (c.id > me.id);
        pk += (unsigned long long)a << c.tshift;
      },
      [&]() {
        c0 += (int)(pk & FIELD_MAX); c1 += (int)((pk >> PG_TYPE_BITS) & FIELD_MAX);
        c2 += (int)((pk >> (2 * PG_TYPE_BITS)) & FIELD_MAX); c3 += (int)((pk >> (3 * PG_TYPE_BITS)) & FIELD_MAX);
        c4 += (int)((pk >> (4 * PG_TYPE_BITS)) & FIELD_MAX);
        pk = 0;
      });
    row_count[me.row] = upper ? up : deg;
    if (degree) degree[me.row] = deg;
    if (nbr_count) {
      int32_t* o = nbr_count + (int64_t)me.row * n_types;
      if (!WIDE_TYPES) {
        o[0] = c0;
        if (n_types > 1) o[1] = c1;
        if (n_types > 2) o[2] = c2;
        if (n_types > 3) o[3] = c3;
        if (n_types > 4) o[4] = c4;
      } else {  // more than PG_PACKED_TYPES types: a second walk with counters in local memory (rare)
        int tc[PG_MAX_TYPES];
#pragma unroll
        for (int t = 0; t < PG_MAX_TYPES; ++t) tc[t] = 0;
        walk_block<MERGED>(g, R, cx, cy, q,
          [&](int j, const pg_rec& c) {
            const double d2 = pg_dist2(me.x, me.y, c.x, c.y);
            if (d2 <= r2 && j != q && c.type >= 1 && c.type <= n_types) tc[c.type - 1] += 1;
          },
          [&]() {});
        for (int t = 0; t < n_types; ++t) o[t] = tc[t];
      }
    }
    if (hist_mode == 1) atomicAdd(&s_hist[min(deg, hist_len - 1)], 1);
    else if (hist_mode == 2) atomicAdd(&hist[min(deg, hist_len - 1)], 1);
  }
  if (!stats && hist_mode != 1) return;  // grid-uniform

  // ---- degree statistics: warp -> CTA (shared-memory atomics, no barrier: the last warp to arrive
  // forwards the CTA's totals) -> accumulators in the handle. The scan that follows publishes them.
  const int wmn = __reduce_min_sync(0xffffffffu, active ? deg : 0x7fffffff);
  const int wmx = __reduce_max_sync(0xffffffffu, active ? deg : -1);
  const int wcnt = __reduce_add_sync(0xffffffffu, active ? 1 : 0);
  const long long wsum = warp_sum_ll(active ? (long long)deg : 0ll);
  const long long wsq = warp_sum_ll(active ? (long long)deg * deg : 0ll);
  int last_warp = 0;
  if ((tid & 31) == 0) {
    if (wcnt > 0) {
      atomicMin(&s_mn, wmn); atomicMax(&s_mx, wmx); atomicAdd(&s_cnt, wcnt);
      atomicAdd(&s_sum, (unsigned long long)wsum); atomicAdd(&s_sq, (unsigned long long)wsq);
    }
    __threadfence_block();
    last_warp = atomicAdd(&s_arrived, 1) == TPB_COUNT / 32 - 1;
  }
  if (!__shfl_sync(0xffffffffu, last_warp, 0)) return;
  __threadfence_block();
  if ((tid & 31) == 0 && *(volatile int*)&s_cnt > 0) {
    atomicMin(&acc->min_degree, *(volatile int*)&s_mn); atomicMax(&acc->max_degree, *(volatile int*)&s_mx);
    atomicAdd(&acc->sum_degree, *(volatile unsigned long long*)&s_sum);
    atomicAdd(&acc->sumsq_degree, *(volatile unsigned long long*)&s_sq);
    atomicAdd(&acc->n_nodes, (unsigned long long)*(volatile int*)&s_cnt);
  }
  if (hist_mode == 1)
    for (int i = tid & 31; i < hist_len; i += 32) {
      const int c = *(volatile int*)&s_hist[i];
      if (c) atomicAdd(&acc_hist[i], c);
    }
}

// nothing to query: the statistics of an empty graph
__global__ void empty_stats_kernel(pg_degree_stats* stats, int32_t* hist, int hist_len) {
  if (stats && threadIdx.x == 0) {
    stats->min_degree = 0; stats->max_degree = 0; stats->sum_degree = 0; stats->sumsq_degree = 0; stats->n_nodes = 0;
  }
  if (hist)
    for (int i = threadIdx.x; i < hist_len; i += blockDim.x) hist[i] = 0;
}

struct fill_out {
  int32_t* __restrict__ col;
  float* __restrict__ dist32;
  double* __restrict__ dist64;
  long long* __restrict__ edges;
  long long* __restrict__ edge_index;
  float* __restrict__ edge_attr;
  long long n_edges;
};

__device__ __forceinline__ void emit_entry(const fill_out& o, long long pos, int my_id, int key, double d2) {
  const double d = sqrt(d2);
  o.col[pos] = key;
  if (o.dist32) o.dist32[pos] = (float)d;
  if (o.dist64) o.dist64[pos] = d;
  if (o.edges) *reinterpret_cast<longlong2*>(o.edges + 2 * pos) = make_longlong2(my_id, key);
  if (o.edge_index) {  // [2, 2E] = hstack(edges.T, edges[:, ::-1].T), ipynb:3021 (see SURVEY B-3)
    o.edge_index[pos] = my_id; o.edge_index[o.n_edges + pos] = key;
    o.edge_index[2 * o.n_edges + pos] = key; o.edge_index[3 * o.n_edges + pos] = my_id;
  }
  if (o.edge_attr) { o.edge_attr[pos] = (float)d; o.edge_attr[o.n_edges + pos] = (float)d; }  // ipynb:3041-3042
}

// Fill pass. A row of up to FILL_CAP entries (known from row_ptr before the walk) takes one walk:
// accepted (id, position) pairs are appended to the thread's column of two shared-memory arrays, sorted
// by id there, and written out with the distance recomputed from the (L1-resident) record. Longer rows
// are emitted in chunks of FILL_CAP: each walk keeps the FILL_CAP smallest ids above the last one written.
template <bool MERGED>
__global__ void __launch_bounds__(TPB_FILL)
radius_fill_kernel(pg_grid_view g, double r2, int R, int upper, const int32_t* __restrict__ row_ptr, fill_out o,
                   long long capacity, int32_t* overflow) {
  __shared__ int s_key[FILL_CAP][TPB_FILL];
  __shared__ int s_pos[FILL_CAP][TPB_FILL];
  const int tid = threadIdx.x;
  const int q = blockIdx.x * TPB_FILL + tid;
  if (q >= g.n) return;
  const pg_rec me = pg_ld_rec(g.rec + q);
  if (me.row >= g.n_query) return;
  const int base = row_ptr[me.row];
  const int cnt = row_ptr[me.row + 1] - base;
  if (cnt <= 0) return;
  if ((long long)base + cnt > capacity) { atomicExch(overflow, 1); return; }
  const int cx = pg_cell_coord(me.x, g.x0, g.inv_cell, g.nx);
  const int cy = pg_cell_coord(me.y, g.y0, g.inv_cell, g.ny);

  if (cnt <= FILL_CAP) {
    int m = 0;
    walk_block<MERGED>(g, R, cx, cy, q,
      [&](int j, const pg_rec& c) {
        const double d2 = pg_dist2(me.x, me.y, c.x, c.y);
        if (d2 <= r2 && j != q && (!upper || c.id > me.id) && m < FILL_CAP) {
          s_key[m][tid] = c.id; s_pos[m][tid] = j; ++m;
        }
      },
      [&]() {});
    for (int a = 1; a < m; ++a) {  // insertion sort of a handful of entries, own column only
      const int k = s_key[a][tid], p = s_pos[a][tid];
      int s = a;
      while (s > 0 && s_key[s - 1][tid] > k) { s_key[s][tid] = s_key[s - 1][tid]; s_pos[s][tid] = s_pos[s - 1][tid]; --s; }
      s_key[s][tid] = k; s_pos[s][tid] = p;
    }
    for (int t = 0; t < m; ++t) {
      const double2 c = pg_ld_xy(g.rec + s_pos[t][tid]);
      emit_entry(o, (long long)base + t, me.id, s_key[t][tid], pg_dist2(me.x, me.y, c.x, c.y));
    }
    return;
  }

  int emitted = 0;
  int last = upper ? me.id : -1;
  while (emitted < cnt) {
    int m = 0;
    walk_block<MERGED>(g, R, cx, cy, q,
      [&](int j, const pg_rec& c) {
        const double d2 = pg_dist2(me.x, me.y, c.x, c.y);
        if (!(d2 <= r2) || j == q || c.id <= last) return;
        if (m == FILL_CAP) {
          if (c.id >= s_key[FILL_CAP - 1][tid]) return;
          m = FILL_CAP - 1;
        }
        int s = m;
        while (s > 0 && s_key[s - 1][tid] > c.id) { s_key[s][tid] = s_key[s - 1][tid]; s_pos[s][tid] = s_pos[s - 1][tid]; --s; }
        s_key[s][tid] = c.id; s_pos[s][tid] = j;
        ++m;
      },
      [&]() {});
    if (m == 0) break;  // cannot happen when row_ptr came from the matching count pass
    for (int t = 0; t < m; ++t) {
      const double2 c = pg_ld_xy(g.rec + s_pos[t][tid]);
      emit_entry(o, (long long)base + emitted + t, me.id, s_key[t][tid], pg_dist2(me.x, me.y, c.x, c.y));
    }
    emitted += m;
    last = s_key[m - 1][tid];
  }
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
  bool launched = false;
  int hist_mode_used = 0;
  if (gr.n > 0 && nq > 0) {
    launched = true;
    pg_grid_view v = pg_make_view(h);
    const int R = ring_radius(r, gr);
    const int upper = flags == PG_RADIUS_UPPER;
    const int nt = nbr_count ? n_types : 1;
    int hist_mode = 0;
    if (hist) {
      hist_mode = hist_len <= PG_ACC_HIST_MAX ? 1 : 2;
      if (hist_mode == 2) PG_CUDA(h, cudaMemsetAsync(hist, 0, (size_t)hist_len * sizeof(int32_t), s));
    }
    hist_mode_used = hist_mode;
    pg_stats_acc* acc = (pg_stats_acc*)((char*)h->misc.p + PG_MISC_ACC);
    int32_t* acc_hist = (int32_t*)((char*)h->misc.p + PG_MISC_ACC_HIST);
    const int blocks = pg_div_up(gr.n, TPB_COUNT);
    const bool wide = nbr_count && n_types > PG_PACKED_TYPES;
#define PG_COUNT_LAUNCH(M, W)                                                                                   \
  PG_LAUNCH(h, s, "radius_count_kernel", radius_count_kernel<M, W><<<blocks, TPB_COUNT, 0, s>>>(                \
      v, r * r, R, upper, (int32_t*)h->row_count.p, degree, nbr_count, nt, acc, acc_hist, stats != nullptr, hist, hist_len, hist_mode))
    if (R == 1 && !wide) PG_COUNT_LAUNCH(true, false);
    else if (R == 1) PG_COUNT_LAUNCH(true, true);
    else if (!wide) PG_COUNT_LAUNCH(false, false);
    else PG_COUNT_LAUNCH(false, true);
#undef PG_COUNT_LAUNCH
    PG_LAUNCH_CHECK(h);
  } else if (stats || hist) {
    PG_LAUNCH(h, s, "empty_stats_kernel", empty_stats_kernel<<<1, 256, 0, s>>>(stats, hist, hist ? hist_len : 0));
    PG_LAUNCH_CHECK(h);
  }
  pg_scan_publish pub;
  if (launched && (stats || hist_mode_used == 1)) {
    pub.acc = (pg_stats_acc*)((char*)h->misc.p + PG_MISC_ACC);
    pub.acc_hist = (int32_t*)((char*)h->misc.p + PG_MISC_ACC_HIST);
    pub.stats = stats;
    pub.hist = hist_mode_used == 1 ? hist : nullptr;
    pub.hist_len = hist_len;
  }
  return pg_scan_i32(h, (const int32_t*)h->row_count.p, row_ptr, nq, s, (int32_t*)((char*)h->misc.p + PG_MISC_TOTALS), false,
                     pub.acc ? &pub : nullptr);
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
  PG_REQUIRE(h, ((uintptr_t)edges_i64 & 15) == 0, "pg_radius_fill: edges must be 16-byte aligned");
  const pg_grid& gr = h->grid;
  if (gr.n == 0 || gr.n_query == 0) return PG_OK;
  pg_grid_view v = pg_make_view(h);
  const int R = ring_radius(h->radius_r, gr);
  const int upper = h->radius_flags == PG_RADIUS_UPPER;
  fill_out o{col, dist32, dist64, (long long*)edges_i64, (long long*)edge_index, edge_attr, (long long)n_edges};
  int32_t* ovf = (int32_t*)((char*)h->misc.p + PG_MISC_OVERFLOW);
  const int blocks = pg_div_up(gr.n, TPB_FILL);
  const double r2 = h->radius_r * h->radius_r;
  if (R == 1)
    PG_LAUNCH(h, s, "radius_fill_kernel", radius_fill_kernel<true><<<blocks, TPB_FILL, 0, s>>>(v, r2, R, upper, row_ptr, o, (long long)capacity, ovf));
  else
    PG_LAUNCH(h, s, "radius_fill_kernel", radius_fill_kernel<false><<<blocks, TPB_FILL, 0, s>>>(v, r2, R, upper, row_ptr, o, (long long)capacity, ovf));
  PG_LAUNCH_CHECK(h);
  return PG_OK;
}

}  // extern "C"
