// K1: fused tile->WSI shift + polygon morphology over CSR-packed rings.
// Reference: /root/reference/aggregated_hovernet_run.py:302-334 (centroid / bbox / polygon shift),
// /root/reference/polygon_morphology.py:240-248 and create_and_overlay_polygon_from_prediction.py:298-299
// (shapely area / length / centroid / bounds), hovernet_tile_inference.ipynb:2415-2456 (eccentricity,
// axes, compactness). Formulas: SURVEY A.4 (Green's-theorem area moments in a frame at vertex 0).
//
// Work split: a warp owns 32 consecutive polygons. The vertex stream is walked by sub-warps of 8
// lanes (group g takes polygon 8g+t in step t), each lane accumulating the edge terms of every 8th
// edge in float64; a 3-step butterfly folds the 8 partials, and lane 8g+t keeps the totals of "its"
// polygon so the closed-form epilogue and all per-polygon loads / stores run one polygon per lane,
// fully coalesced.
#include "pg_common.cuh"

namespace {

constexpr int TPB = 256;

template <typename T> struct vec2_of;
template <> struct vec2_of<float> { using type = float2; };
template <> struct vec2_of<double> { using type = double2; };

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double bfly_add(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
__device__ __forceinline__ float bfly_addf(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
template <typename T> __device__ __forceinline__ T bfly_min(T v) {
  v = min(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = min(v, __shfl_xor_sync(0xffffffffu, v, 2));
  v = min(v, __shfl_xor_sync(0xffffffffu, v, 4));
  return v;
}
template <typename T> __device__ __forceinline__ T bfly_max(T v) {
  v = max(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = max(v, __shfl_xor_sync(0xffffffffu, v, 2));
  v = max(v, __shfl_xor_sync(0xffffffffu, v, 4));
  return v;
}

template <typename T, bool EXTRA>
__global__ void __launch_bounds__(TPB)
map_morph_kernel(int n, const int32_t* __restrict__ poly_off, const typename vec2_of<T>::type* __restrict__ poly,
                 const int32_t* __restrict__ nuc_tile, const int32_t* __restrict__ tile_x,
                 const int32_t* __restrict__ tile_y, const double2* __restrict__ centroid,
                 const int4* __restrict__ bbox, typename vec2_of<T>::type* __restrict__ wsi_poly,
                 double2* __restrict__ wsi_centroid, int4* __restrict__ wsi_bbox, pg_morph_out out) {
  using V2 = typename vec2_of<T>::type;
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * TPB + threadIdx.x) >> 5;
  const int my_poly = warp_global * 32 + lane;
  if (warp_global * 32 >= n) return;  // whole warp out of range

  // ---- one polygon per lane: offsets, tile shift, centroid / bbox shift (aggregated_hovernet_run.py:302-319)
  int off0 = 0, nv = 0, itx = 0, ity = 0;
  if (my_poly < n) {
    off0 = poly_off[my_poly];
    nv = poly_off[my_poly + 1] - off0;
    if (nuc_tile) { const int t = nuc_tile[my_poly]; itx = tile_x[t]; ity = tile_y[t]; }
    if (centroid) { const double2 c = centroid[my_poly]; wsi_centroid[my_poly] = make_double2(c.x + (double)itx, c.y + (double)ity); }
    if (bbox) { const int4 b = bbox[my_poly]; wsi_bbox[my_poly] = make_int4(b.x + itx, b.y + ity, b.z + itx, b.w + ity); }
  }

  // totals of my polygon (filled in when my group processes it)
  double S = 0, Sx = 0, Sy = 0, Ixx = 0, Iyy = 0, Ixy = 0;
  float P = 0;
  double fx = 0, fy = 0;  // frame origin = first vertex
  T bx0 = 0, by0 = 0, bx1 = 0, by1 = 0;

  const int grp = lane >> 3, sub = lane & 7;
#pragma unroll 1
  for (int t = 0; t < 8; ++t) {
    const int src = (grp << 3) + t;
    const int o = __shfl_sync(0xffffffffu, off0, src);
    const int cnt = __shfl_sync(0xffffffffu, nv, src);
    const int sx = __shfl_sync(0xffffffffu, itx, src);
    const int sy = __shfl_sync(0xffffffffu, ity, src);
    double a_s = 0, a_sx = 0, a_sy = 0, a_ixx = 0, a_iyy = 0, a_ixy = 0;
    float a_p = 0;
    V2 v0; v0.x = 0; v0.y = 0;
    T mnx = 0, mny = 0, mxx = 0, mxy = 0;
    if (cnt > 0) {
      v0 = poly[o];
      if (EXTRA) { mnx = mxx = v0.x; mny = mxy = v0.y; }
      const T tsx = (T)sx, tsy = (T)sy;
      for (int e = sub; e < cnt; e += 8) {
        const V2 va = poly[o + e];
        const V2 vb = poly[o + ((e + 1 == cnt) ? 0 : e + 1)];
        if (wsi_poly) { V2 w; w.x = va.x + tsx; w.y = va.y + tsy; wsi_poly[o + e] = w; }  // :322-334
        const double xa = (double)va.x - (double)v0.x, ya = (double)va.y - (double)v0.y;
        const double xb = (double)vb.x - (double)v0.x, yb = (double)vb.y - (double)v0.y;
        const double a = xa * yb - xb * ya;
        const double sxx = xa + xb, syy = ya + yb;
        a_s += a;
        a_sx += sxx * a;
        a_sy += syy * a;
        a_ixx += (syy * syy - ya * yb) * a;   // ya^2 + ya*yb + yb^2
        a_iyy += (sxx * sxx - xa * xb) * a;   // xa^2 + xa*xb + xb^2
        a_ixy += (sxx * syy + xa * ya + xb * yb) * a;  // xa*yb + 2xa*ya + 2xb*yb + xb*ya
        const float dx = (float)(xb - xa), dy = (float)(yb - ya);
        a_p += sqrtf(dx * dx + dy * dy);
        if (EXTRA) { mnx = min(mnx, va.x); mxx = max(mxx, va.x); mny = min(mny, va.y); mxy = max(mxy, va.y); }
      }
    }
    a_s = bfly_add(a_s); a_sx = bfly_add(a_sx); a_sy = bfly_add(a_sy);
    a_ixx = bfly_add(a_ixx); a_iyy = bfly_add(a_iyy); a_ixy = bfly_add(a_ixy);
    a_p = bfly_addf(a_p);
    if (EXTRA) { mnx = bfly_min(mnx); mny = bfly_min(mny); mxx = bfly_max(mxx); mxy = bfly_max(mxy); }
    if (sub == t) {
      S = a_s; Sx = a_sx; Sy = a_sy; Ixx = a_ixx; Iyy = a_iyy; Ixy = a_ixy; P = a_p;
      fx = (double)v0.x; fy = (double)v0.y;
      if (EXTRA) { bx0 = mnx; by0 = mny; bx1 = mxx; by1 = mxy; }
    }
  }

  // ---- epilogue: one polygon per lane
  if (my_poly >= n) return;
  const float nanf_ = __int_as_float(0x7fc00000);
  const double nand_ = __longlong_as_double(0x7ff8000000000000ll);
  const bool ok = nv >= 3;
  const double A = 0.5 * S;
  const double area = fabs(A);
  const double per = (double)P;
  if (out.area) out.area[my_poly] = ok ? (float)area : nanf_;
  if (out.perimeter) out.perimeter[my_poly] = ok ? P : nanf_;
  if (out.circularity) {
    const double pm = fmax(per, 1.0);
    out.circularity[my_poly] = ok ? (float)(4.0 * 3.14159265358979323846 * area / (pm * pm)) : nanf_;
  }
  const bool good = ok && A != 0.0;
  double ecc = nand_, major = nand_, minor = nand_, cxw = nand_, cyw = nand_;
  if (good) {
    const double inv = 1.0 / A;
    const double cx = Sx * inv * (1.0 / 6.0), cy = Sy * inv * (1.0 / 6.0);
    const double mu20 = Iyy * inv * (1.0 / 12.0) - cx * cx;
    const double mu02 = Ixx * inv * (1.0 / 12.0) - cy * cy;
    const double mu11 = Ixy * inv * (1.0 / 24.0) - cx * cy;
    const double m = 0.5 * (mu20 + mu02);
    const double hd = 0.5 * (mu20 - mu02);
    const double c = sqrt(hd * hd + mu11 * mu11);
    const double l1 = m + c;
    const double l2 = fmax(m - c, 0.0);
    ecc = l1 > 0.0 ? sqrt(1.0 - l2 / l1) : 0.0;
    major = 4.0 * sqrt(fmax(l1, 0.0));
    minor = 4.0 * sqrt(l2);
    cxw = cx + fx + (double)itx;
    cyw = cy + fy + (double)ity;
  }
  if (out.eccentricity) out.eccentricity[my_poly] = (float)ecc;
  if (EXTRA) {
    if (out.major_axis) out.major_axis[my_poly] = (float)major;
    if (out.minor_axis) out.minor_axis[my_poly] = (float)minor;
    if (out.centroid_x) out.centroid_x[my_poly] = cxw;
    if (out.centroid_y) out.centroid_y[my_poly] = cyw;
    if (out.poly_bbox) {
      double* b = out.poly_bbox + (int64_t)my_poly * 4;
      b[0] = ok ? (double)bx0 + (double)itx : nand_; b[1] = ok ? (double)by0 + (double)ity : nand_;
      b[2] = ok ? (double)bx1 + (double)itx : nand_; b[3] = ok ? (double)by1 + (double)ity : nand_;
    }
  }
}

template <typename T>
int launch_map_morph(pg_handle* h, int32_t n, const int32_t* poly_off, const T* poly_xy, const int32_t* nuc_tile,
                     const int32_t* tile_x, const int32_t* tile_y, const double* centroid, const int32_t* bbox,
                     T* wsi_poly_xy, double* wsi_centroid, int32_t* wsi_bbox, const pg_morph_out* out, pg_stream stream) {
  if (!h) return PG_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  PG_CUDA(h, cudaSetDevice(h->device));
  h->last_stream = s;
  using V2 = typename vec2_of<T>::type;
  PG_REQUIRE(h, n >= 0, "pg_map_morph: n < 0");
  PG_REQUIRE(h, n == 0 || poly_off, "pg_map_morph: poly_off is NULL");
  PG_REQUIRE(h, !nuc_tile || (tile_x && tile_y), "pg_map_morph: nuc_tile given without tile_x / tile_y");
  PG_REQUIRE(h, !centroid || wsi_centroid, "pg_map_morph: centroid given without wsi_centroid");
  PG_REQUIRE(h, !bbox || wsi_bbox, "pg_map_morph: bbox given without wsi_bbox");
  PG_REQUIRE(h, (((uintptr_t)poly_xy | (uintptr_t)wsi_poly_xy) & (sizeof(V2) - 1)) == 0 &&
                    (((uintptr_t)centroid | (uintptr_t)wsi_centroid | (uintptr_t)bbox | (uintptr_t)wsi_bbox) & 15) == 0,
             "pg_map_morph: vertex / centroid / bbox arrays must be aligned to their vector width");
  if (n == 0) return PG_OK;
  pg_morph_out o{};
  if (out) o = *out;
  const bool extra = o.major_axis || o.minor_axis || o.centroid_x || o.centroid_y || o.poly_bbox;
  const int blocks = pg_div_up(n, TPB);  // 32 polygons per warp, 8 warps per CTA
  if (extra)
    PG_LAUNCH(h, s, "map_morph_kernel<T, true>", map_morph_kernel<T, true><<<blocks, TPB, 0, s>>>(n, poly_off, (const V2*)poly_xy, nuc_tile, tile_x, tile_y,
                                                     (const double2*)centroid, (const int4*)bbox, (V2*)wsi_poly_xy,
                                                     (double2*)wsi_centroid, (int4*)wsi_bbox, o));
  else
    PG_LAUNCH(h, s, "map_morph_kernel<T, false>", map_morph_kernel<T, false><<<blocks, TPB, 0, s>>>(n, poly_off, (const V2*)poly_xy, nuc_tile, tile_x, tile_y,
                                                      (const double2*)centroid, (const int4*)bbox, (V2*)wsi_poly_xy,
                                                      (double2*)wsi_centroid, (int4*)wsi_bbox, o));
  PG_LAUNCH_CHECK(h);
  return PG_OK;
}

}  // namespace

extern "C" {

int pg_map_morph_f32(pg_handle* h, int32_t n, const int32_t* poly_off, const float* poly_xy, const int32_t* nuc_tile,
                     const int32_t* tile_x, const int32_t* tile_y, const double* centroid, const int32_t* bbox,
                     float* wsi_poly_xy, double* wsi_centroid, int32_t* wsi_bbox, const pg_morph_out* out,
                     pg_stream stream) {
  return launch_map_morph<float>(h, n, poly_off, poly_xy, nuc_tile, tile_x, tile_y, centroid, bbox, wsi_poly_xy,
                                 wsi_centroid, wsi_bbox, out, stream);
}

int pg_map_morph_f64(pg_handle* h, int32_t n, const int32_t* poly_off, const double* poly_xy, const int32_t* nuc_tile,
                     const int32_t* tile_x, const int32_t* tile_y, const double* centroid, const int32_t* bbox,
                     double* wsi_poly_xy, double* wsi_centroid, int32_t* wsi_bbox, const pg_morph_out* out,
                     pg_stream stream) {
  return launch_map_morph<double>(h, n, poly_off, poly_xy, nuc_tile, tile_x, tile_y, centroid, bbox, wsi_poly_xy,
                                  wsi_centroid, wsi_bbox, out, stream);
}

}  // extern "C"
