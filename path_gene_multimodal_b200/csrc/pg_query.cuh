// Shared device-side pieces of the grid queries (K5 kNN, K6 radius).
#pragma once
#include "pg_common.cuh"

struct pg_grid_view {
  const int32_t* __restrict__ cell_start;
  const double2* __restrict__ s_xy;
  const int2* __restrict__ s_meta;  // {local idx, type}
  const int32_t* __restrict__ s_gid;  // global ids in cell order, or NULL when ids are the local indices
  int32_t n, n_query, nx, ny;
  double x0, y0, cell, inv_cell;
};

static inline pg_grid_view pg_make_view(const pg_handle* h) {
  pg_grid_view v;
  v.cell_start = (const int32_t*)h->cell_start.p + 3;  // see pg_grid.cu: B[c] = first point of cell c
  v.s_xy = (const double2*)h->s_xy.p;
  v.s_meta = (const int2*)h->s_meta.p;
  v.s_gid = h->grid.has_gid ? (const int32_t*)h->s_gid.p : nullptr;
  v.n = h->grid.n; v.n_query = h->grid.n_query; v.nx = h->grid.nx; v.ny = h->grid.ny;
  v.x0 = h->grid.x0; v.y0 = h->grid.y0; v.cell = h->grid.cell; v.inv_cell = h->grid.inv_cell;
  return v;
}

#ifdef __CUDACC__
// id used for ordering / output columns of the point at cell-order position j
__device__ __forceinline__ int pg_id_of(const pg_grid_view& g, int j, int local_idx) {
  return g.s_gid ? g.s_gid[j] : local_idx;
}

// Visit the (2R+1)^2 block of cells around (cx, cy) as 2R+1 contiguous runs of the cell-ordered
// point array (cells are row-major, so one grid row of the block is one run).
template <class F>
__device__ __forceinline__ void pg_visit_block(const pg_grid_view& g, int cx, int cy, int R, F&& f) {
  const int xa = max(cx - R, 0), xb = min(cx + R, g.nx - 1);
  const int ya = max(cy - R, 0), yb = min(cy + R, g.ny - 1);
  for (int y = ya; y <= yb; ++y) {
    const int row = y * g.nx;
    f(g.cell_start[row + xa], g.cell_start[row + xb + 1]);
  }
}

// Visit only the ring at Chebyshev distance exactly R (R >= 1) around (cx, cy).
template <class F>
__device__ __forceinline__ void pg_visit_ring(const pg_grid_view& g, int cx, int cy, int R, F&& f) {
  const int xa = max(cx - R, 0), xb = min(cx + R, g.nx - 1);
  if (cy - R >= 0) { const int row = (cy - R) * g.nx; f(g.cell_start[row + xa], g.cell_start[row + xb + 1]); }
  if (cy + R < g.ny) { const int row = (cy + R) * g.nx; f(g.cell_start[row + xa], g.cell_start[row + xb + 1]); }
  const int ya = max(cy - R + 1, 0), yb = min(cy + R - 1, g.ny - 1);
  const bool left = cx - R >= 0, right = cx + R < g.nx;
  for (int y = ya; y <= yb; ++y) {
    const int row = y * g.nx;
    if (left) f(g.cell_start[row + cx - R], g.cell_start[row + cx - R + 1]);
    if (right) f(g.cell_start[row + cx + R], g.cell_start[row + cx + R + 1]);
  }
}

// Bounded sorted buffer used to emit a row in ascending key order in chunks of CAP entries:
// pass m keeps the CAP smallest keys greater than the last key already emitted.
template <int CAP>
struct pg_sorted_chunk {
  int key[CAP];
  double val[CAP];
  int m;
  int last;
  __device__ __forceinline__ void reset(int last_key) { m = 0; last = last_key; }
  __device__ __forceinline__ void push(int k, double v) {
    if (k <= last) return;
    if (m == CAP) {
      if (k >= key[CAP - 1]) return;
      m = CAP - 1;
    }
    int s = m;
    while (s > 0 && key[s - 1] > k) { key[s] = key[s - 1]; val[s] = val[s - 1]; --s; }
    key[s] = k; val[s] = v;
    ++m;
  }
};
#endif
