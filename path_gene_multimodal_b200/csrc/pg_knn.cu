// K5: k-nearest-neighbour query on the uniform grid.
// Reference: /root/reference/hovernet_tile_inference.ipynb:1815-1850 - KNN.from_array(coords, k)
// (libpysal -> scipy cKDTree.query(k+1) minus self) and the per-neighbour sqrt(dx*dx+dy*dy).
// Order is the canonical (d^2, id) of north_star; self is removed by index, so duplicates at
// distance 0 are ordinary neighbours.
#include <cmath>
#include "pg_query.cuh"

namespace {

constexpr int TPB = 128;

template <int KMAX>
struct topk {
  double d2[KMAX];
  int id[KMAX];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int s = 0; s < KMAX; ++s) { d2[s] = __longlong_as_double(0x7ff0000000000000ll); id[s] = 0x7fffffff; }
  }
  // insert (cd2, cid), known to sort before slot k-1; fully unrolled so the lists stay in registers
  __device__ __forceinline__ void insert(double cd2, int cid, int k) {
    bool placed = false;
#pragma unroll
    for (int s = KMAX - 1; s > 0; --s) {
      if (s < k && !placed) {
        const bool before_prev = cd2 < d2[s - 1] || (cd2 == d2[s - 1] && cid < id[s - 1]);
        if (before_prev) { d2[s] = d2[s - 1]; id[s] = id[s - 1]; }
        else { d2[s] = cd2; id[s] = cid; placed = true; }
      }
    }
    if (!placed) { d2[0] = cd2; id[0] = cid; }
  }
  __device__ __forceinline__ double worst_d2(int k) const {
    double w = d2[0];
#pragma unroll
    for (int s = 1; s < KMAX; ++s) if (s == k - 1) w = d2[s];
    return w;
  }
  __device__ __forceinline__ int worst_id(int k) const {
    int w = id[0];
#pragma unroll
    for (int s = 1; s < KMAX; ++s) if (s == k - 1) w = id[s];
    return w;
  }
};

// One thread per point in cell order. Ring expansion: after the (2R+1)^2 block has been searched
// the result is final once the k-th best distance is no larger than the distance from the query
// to the nearest face of the block that still has cells behind it.
template <int KMAX>
__global__ void __launch_bounds__(TPB)
knn_kernel(pg_grid_view g, int k, int32_t* __restrict__ knn_idx, double* __restrict__ dist64,
           float* __restrict__ dist32, double x_lo, double x_hi, int32_t* halo_ok) {
  const int p = blockIdx.x * TPB + threadIdx.x;
  if (p >= g.n) return;
  const pg_rec me = pg_ld_rec(g.rec + p);
  if (me.row >= g.n_query) return;
  const double2 q = make_double2(me.x, me.y);
  const int cx = pg_cell_coord(q.x, g.x0, g.inv_cell, g.nx);
  const int cy = pg_cell_coord(q.y, g.y0, g.inv_cell, g.ny);
  topk<KMAX> top;
  top.init();
  double wd2 = top.d2[0];
  int wid = top.id[0];
  auto scan = [&](int b, int e) {
    for (int j = b; j < e; ++j) {
      const pg_rec c = pg_ld_rec(g.rec + j);
      const double d2 = pg_dist2(q.x, q.y, c.x, c.y);
      if (d2 <= wd2 && j != p) {
        const int cid = c.id;
        if (d2 < wd2 || cid < wid) {
          top.insert(d2, cid, k);
          wd2 = top.worst_d2(k);
          wid = top.worst_id(k);
        }
      }
    }
  };
  const double margin = 1e-6 * g.cell;
  int R = 1;
  pg_visit_block(g, cx, cy, 1, scan);
  while (true) {
    const bool covers = cx - R <= 0 && cx + R >= g.nx - 1 && cy - R <= 0 && cy + R >= g.ny - 1;
    if (covers) break;
    // distance to the faces of the searched block that have unsearched cells behind them
    double bound = __longlong_as_double(0x7ff0000000000000ll);
    if (cx - R > 0) bound = fmin(bound, q.x - (g.x0 + (double)(cx - R) * g.cell));
    if (cx + R < g.nx - 1) bound = fmin(bound, (g.x0 + (double)(cx + R + 1) * g.cell) - q.x);
    if (cy - R > 0) bound = fmin(bound, q.y - (g.y0 + (double)(cy - R) * g.cell));
    if (cy + R < g.ny - 1) bound = fmin(bound, (g.y0 + (double)(cy + R + 1) * g.cell) - q.y);
    bound -= margin;
    if (bound > 0.0 && wd2 <= bound * bound) break;
    ++R;
    pg_visit_ring(g, cx, cy, R, scan);
  }
  const int64_t o = (int64_t)me.row * k;
#pragma unroll
  for (int s = 0; s < KMAX; ++s) {
    if (s < k) {
      knn_idx[o + s] = top.id[s];
      const double d = sqrt(top.d2[s]);
      if (dist64) dist64[o + s] = d;
      if (dist32) dist32[o + s] = (float)d;
    }
  }
  if (halo_ok) {
    // every point with x in [x_lo, x_hi) is present; the answer is complete iff the k-th
    // neighbour is strictly closer than both faces (slack keeps the test conservative)
    const double dk = sqrt(wd2) * (1.0 + 1e-9);
    if (!(dk < q.x - x_lo && dk < x_hi - q.x)) atomicExch(halo_ok, 0);
  }
}

__global__ void set_flag_kernel(int32_t* f, int v) { *f = v; }

}  // namespace

extern "C" int pg_knn(pg_handle* h, int32_t k, int32_t* knn_idx, double* dist64, float* dist32,
                      double x_lo, double x_hi, int32_t* halo_ok, pg_stream stream) {
  if (!h) return PG_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  PG_CUDA(h, cudaSetDevice(h->device));
  h->last_stream = s;
  if (!h->grid.built) return pg_set_error(h, PG_ERR_STATE, "pg_knn: call pg_grid_build first");
  const pg_grid& gr = h->grid;
  PG_REQUIRE(h, k >= 1 && k <= PG_MAX_K, "pg_knn: k must be in 1..%d (got %d)", PG_MAX_K, k);
  PG_REQUIRE(h, k < gr.n, "pg_knn: k (%d) must be smaller than the number of points (%d)", k, gr.n);
  PG_REQUIRE(h, knn_idx != nullptr, "pg_knn: knn_idx is NULL");
  if (halo_ok) { PG_LAUNCH(h, s, "set_flag_kernel", set_flag_kernel<<<1, 1, 0, s>>>(halo_ok, 1)); }
  if (gr.n_query == 0) return PG_OK;
  pg_grid_view v = pg_make_view(h);
  const int blocks = pg_div_up(gr.n, TPB);
  if (k <= 4) PG_LAUNCH(h, s, "knn_kernel<4>", knn_kernel<4><<<blocks, TPB, 0, s>>>(v, k, knn_idx, dist64, dist32, x_lo, x_hi, halo_ok));
  else if (k <= 8) PG_LAUNCH(h, s, "knn_kernel<8>", knn_kernel<8><<<blocks, TPB, 0, s>>>(v, k, knn_idx, dist64, dist32, x_lo, x_hi, halo_ok));
  else if (k <= 16) PG_LAUNCH(h, s, "knn_kernel<16>", knn_kernel<16><<<blocks, TPB, 0, s>>>(v, k, knn_idx, dist64, dist32, x_lo, x_hi, halo_ok));
  else if (k <= 32) PG_LAUNCH(h, s, "knn_kernel<32>", knn_kernel<32><<<blocks, TPB, 0, s>>>(v, k, knn_idx, dist64, dist32, x_lo, x_hi, halo_ok));
  else PG_LAUNCH(h, s, "knn_kernel<64>", knn_kernel<64><<<blocks, TPB, 0, s>>>(v, k, knn_idx, dist64, dist32, x_lo, x_hi, halo_ok));
  PG_LAUNCH_CHECK(h);
  return PG_OK;
}
