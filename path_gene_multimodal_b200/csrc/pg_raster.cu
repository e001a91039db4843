// K12: raster region properties of a HoverNeXt instance map (first half).
// Reference: /root/reference/aggregated_hovernet_run.py:172-181 (regionprops(inst_map): per-instance bbox) and
// /root/reference/hovernet_tile_inference.ipynb:2415-2429 (cell 18: regionprops_table area, perimeter,
// eccentricity, major / minor axis length, orientation). skimage's definitions:
//   area = pixel count; bbox = [min_row, min_col, max_row + 1, max_col + 1]; centroid = mean (row, col);
//   inertia tensor T = [[mu02, -mu11], [-mu11, mu20]] / mu00 (axis 0 = rows), eigenvalues l1 >= l2 >= 0;
//   eccentricity = sqrt(1 - l2 / l1) (0 when l1 == 0); major / minor = 4 sqrt(l1 / l2);
//   orientation = 0.5 atan2(-2 b, c - a) with (a, b, c) = (T00, T01, T11), +-pi/4 when a == c;
//   perimeter (4-neighbourhood): border = pixels with a 4-neighbour outside the region; every border pixel gets
//   the code sum of [[10,2,10],[2,1,2],[10,2,10]] over the border pixels around it and contributes 1 for codes
//   5,7,15,17,25,27, sqrt(2) for 21,33 and (1+sqrt(2))/2 for 13,23.
// The reference does this with one full-mask pass PER INSTANCE (inst_map == inst_id, :183); here the map is read
// twice in total. All accumulators are integers (pixel coordinates are integers: raw moments are exact in int64,
// perimeter is three counters), so the result does not depend on the order of the atomics.
#include <algorithm>
#include "pg_common.cuh"

namespace {

constexpr int TPB = 256;

struct raster_acc {  // one per label
  unsigned long long sr, sc, srr, scc, src;
  int area, min_r, min_c, max_r, max_c;
  int p1, p2, p3;  // perimeter code classes: weight 1, sqrt(2), (1 + sqrt(2)) / 2
};

__global__ void __launch_bounds__(TPB)
raster_init_kernel(raster_acc* acc, int n_labels) {
  const int i = blockIdx.x * TPB + threadIdx.x;
  if (i >= n_labels) return;
  raster_acc a;
  a.sr = a.sc = a.srr = a.scc = a.src = 0ull;
  a.area = 0; a.min_r = a.min_c = 0x7fffffff; a.max_r = a.max_c = -1;
  a.p1 = a.p2 = a.p3 = 0;
  acc[i] = a;
}

__device__ __forceinline__ int label_at(const int32_t* __restrict__ m, int h, int w, int r, int c) {
  return (r >= 0 && r < h && c >= 0 && c < w) ? m[(int64_t)r * w + c] : 0;
}

// pass 1: moments, bbox, border map. One thread per pixel; the lanes of a warp that carry the same label (a run
// of a nucleus along the row) are folded together before they touch the accumulators.
__global__ void __launch_bounds__(TPB)
raster_moments_kernel(const int32_t* __restrict__ m, int h, int w, int n_labels, raster_acc* __restrict__ acc,
                      uint8_t* __restrict__ border) {
  const int64_t px = (int64_t)blockIdx.x * TPB + threadIdx.x;
  const bool in = px < (int64_t)h * w;
  const int r = in ? (int)(px / w) : 0, c = in ? (int)(px - (int64_t)r * w) : 0;
  int lab = in ? m[px] : 0;
  if (lab < 0 || lab > n_labels) lab = 0;
  bool is_border = false;
  if (lab > 0)
    is_border = label_at(m, h, w, r - 1, c) != lab || label_at(m, h, w, r + 1, c) != lab ||
                label_at(m, h, w, r, c - 1) != lab || label_at(m, h, w, r, c + 1) != lab;
  if (in) border[px] = is_border ? 1 : 0;
  const unsigned peers = __match_any_sync(0xffffffffu, lab);
  if (lab == 0) return;
  const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
  // group sums: values are < 2^31 per lane for maps up to 46340 pixels a side; 64-bit sums via two 32-bit halves
  const unsigned cnt = __popc(peers);
  const unsigned s_r = __reduce_add_sync(peers, (unsigned)r), s_c = __reduce_add_sync(peers, (unsigned)c);
  const unsigned long long rr = (unsigned long long)r * r, cc = (unsigned long long)c * c, rc = (unsigned long long)r * c;
  const unsigned long long s_rr = (unsigned long long)__reduce_add_sync(peers, (unsigned)(rr & 0xffffu)) +
                                  ((unsigned long long)__reduce_add_sync(peers, (unsigned)(rr >> 16)) << 16);
  const unsigned long long s_cc = (unsigned long long)__reduce_add_sync(peers, (unsigned)(cc & 0xffffu)) +
                                  ((unsigned long long)__reduce_add_sync(peers, (unsigned)(cc >> 16)) << 16);
  const unsigned long long s_rc = (unsigned long long)__reduce_add_sync(peers, (unsigned)(rc & 0xffffu)) +
                                  ((unsigned long long)__reduce_add_sync(peers, (unsigned)(rc >> 16)) << 16);
  const int mn_r = __reduce_min_sync(peers, r), mx_r = __reduce_max_sync(peers, r);
  const int mn_c = __reduce_min_sync(peers, c), mx_c = __reduce_max_sync(peers, c);
  if (lane != leader) return;
  raster_acc* a = acc + (lab - 1);
  atomicAdd(&a->area, (int)cnt);
  atomicAdd(&a->sr, (unsigned long long)s_r); atomicAdd(&a->sc, (unsigned long long)s_c);
  atomicAdd(&a->srr, s_rr); atomicAdd(&a->scc, s_cc); atomicAdd(&a->src, s_rc);
  atomicMin(&a->min_r, mn_r); atomicMax(&a->max_r, mx_r);
  atomicMin(&a->min_c, mn_c); atomicMax(&a->max_c, mx_c);
}

// pass 2: perimeter codes of the border pixels
__global__ void __launch_bounds__(TPB)
raster_perimeter_kernel(const int32_t* __restrict__ m, const uint8_t* __restrict__ border, int h, int w,
                        int n_labels, raster_acc* __restrict__ acc) {
  const int64_t px = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (px >= (int64_t)h * w || !border[px]) return;
  const int lab = m[px];
  const int r = (int)(px / w), c = (int)(px - (int64_t)r * w);
  int code = 1;
#pragma unroll
  for (int dr = -1; dr <= 1; ++dr)
#pragma unroll
    for (int dc = -1; dc <= 1; ++dc) {
      if (dr == 0 && dc == 0) continue;
      const int rr = r + dr, cc = c + dc;
      if (rr < 0 || rr >= h || cc < 0 || cc >= w) continue;
      const int64_t q = (int64_t)rr * w + cc;
      if (border[q] && m[q] == lab) code += (dr != 0 && dc != 0) ? 10 : 2;
    }
  raster_acc* a = acc + (lab - 1);
  if (code == 5 || code == 7 || code == 15 || code == 17 || code == 25 || code == 27) atomicAdd(&a->p1, 1);
  else if (code == 21 || code == 33) atomicAdd(&a->p2, 1);
  else if (code == 13 || code == 23) atomicAdd(&a->p3, 1);
}

__global__ void __launch_bounds__(TPB)
raster_finish_kernel(const raster_acc* __restrict__ acc, int n_labels, pg_raster_out o) {
  const int i = blockIdx.x * TPB + threadIdx.x;
  if (i >= n_labels) return;
  const raster_acc a = acc[i];
  const double nan_ = __longlong_as_double(0x7ff8000000000000ll);
  if (o.area) o.area[i] = a.area;
  const bool some = a.area > 0;
  if (o.bbox) {
    o.bbox[4 * i + 0] = some ? a.min_r : 0; o.bbox[4 * i + 1] = some ? a.min_c : 0;
    o.bbox[4 * i + 2] = some ? a.max_r + 1 : 0; o.bbox[4 * i + 3] = some ? a.max_c + 1 : 0;
  }
  double cr = nan_, cc = nan_, ecc = nan_, major = nan_, minor = nan_, orient = nan_, per = nan_;
  if (some) {
    const double n = (double)a.area;
    cr = (double)a.sr / n; cc = (double)a.sc / n;
    // central moments from the exact integer raw moments: mu20 = sum r^2 - (sum r)^2 / n, ...
    const double mu20 = (double)a.srr - (double)a.sr * (double)a.sr / n;
    const double mu02 = (double)a.scc - (double)a.sc * (double)a.sc / n;
    const double mu11 = (double)a.src - (double)a.sr * (double)a.sc / n;
    const double ta = mu02 / n, tb = -mu11 / n, tc = mu20 / n;  // inertia tensor [[ta, tb], [tb, tc]]
    const double mid = 0.5 * (ta + tc), hd = 0.5 * (ta - tc);
    const double rad = sqrt(hd * hd + tb * tb);
    const double l1 = fmax(mid + rad, 0.0), l2 = fmax(mid - rad, 0.0);
    ecc = l1 == 0.0 ? 0.0 : sqrt(1.0 - l2 / l1);
    major = 4.0 * sqrt(l1);
    minor = 4.0 * sqrt(l2);
    const double pi = 3.14159265358979323846;
    orient = (ta - tc == 0.0) ? (tb < 0.0 ? pi / 4.0 : -pi / 4.0) : 0.5 * atan2(-2.0 * tb, tc - ta);
    per = (double)a.p1 + (double)a.p2 * 1.4142135623730951 + (double)a.p3 * ((1.0 + 1.4142135623730951) / 2.0);
  }
  if (o.centroid) { o.centroid[2 * i] = cr; o.centroid[2 * i + 1] = cc; }
  if (o.eccentricity) o.eccentricity[i] = ecc;
  if (o.major_axis) o.major_axis[i] = major;
  if (o.minor_axis) o.minor_axis[i] = minor;
  if (o.orientation) o.orientation[i] = orient;
  if (o.perimeter) o.perimeter[i] = per;
}

}  // namespace

extern "C" int pg_raster_props(pg_handle* h, int32_t height, int32_t width, const int32_t* inst_map, int32_t n_labels,
                               const pg_raster_out* out, pg_stream stream) {
  if (!h) return PG_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  PG_CUDA(h, cudaSetDevice(h->device));
  h->last_stream = s;
  PG_REQUIRE(h, height >= 0 && width >= 0 && n_labels >= 0 && out, "pg_raster_props: bad argument");
  PG_REQUIRE(h, height <= 46340 && width <= 46340, "pg_raster_props: maps up to 46340 pixels a side");
  if (n_labels == 0) return PG_OK;
  const int64_t pixels = (int64_t)height * width;
  PG_REQUIRE(h, pixels == 0 || inst_map, "pg_raster_props: inst_map is NULL");
  int rc;
  if ((rc = pg_reserve(h, h->rank, (size_t)n_labels * sizeof(raster_acc)))) return rc;
  if ((rc = pg_reserve(h, h->cell_of, (size_t)pixels + 16))) return rc;
  raster_acc* acc = (raster_acc*)h->rank.p;
  uint8_t* border = (uint8_t*)h->cell_of.p;
  PG_LAUNCH(h, s, "raster_init_kernel", raster_init_kernel<<<pg_div_up(n_labels, TPB), TPB, 0, s>>>(acc, n_labels));
  if (pixels > 0) {
    const int blocks = pg_div_up(pixels, TPB);
    PG_LAUNCH(h, s, "raster_moments_kernel", raster_moments_kernel<<<blocks, TPB, 0, s>>>(inst_map, height, width, n_labels, acc, border));
    PG_LAUNCH(h, s, "raster_perimeter_kernel", raster_perimeter_kernel<<<blocks, TPB, 0, s>>>(inst_map, border, height, width, n_labels, acc));
  }
  PG_LAUNCH(h, s, "raster_finish_kernel", raster_finish_kernel<<<pg_div_up(n_labels, TPB), TPB, 0, s>>>(acc, n_labels, *out));
  PG_LAUNCH_CHECK(h);
  return PG_OK;
}
