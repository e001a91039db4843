// K3: single-pass exclusive scan (int32) used for cell_start and every CSR row_ptr.
// Decoupled look-back organised for short chains: tiles of 4096 items are handed out by an atomic
// ticket (so a tile only ever waits on tiles that are already running); every tile publishes its
// aggregate at once, then sums the aggregates of the earlier tiles of its GROUP (256 tiles) directly,
// one descriptor per thread - a single L2 round trip - and adds the inclusive prefix published by the
// last tile of the previous group. 1M items = 245 tiles = one group: no serial chain at all.
// Descriptor words carry an epoch tag, so nothing has to be cleared between scans.
#include "pg_common.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr int SCAN_GROUP = SCAN_THREADS;  // tiles per group: one predecessor per thread

__device__ __forceinline__ uint64_t pack_desc(uint32_t epoch, int32_t value) {
  return ((uint64_t)epoch << 32) | (uint32_t)value;
}
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t* p, uint64_t v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// agg[tile]: aggregate of one tile; gpre[group]: inclusive prefix up to the end of a group.
// CLEAR: the input is zeroed as it is read (the cell histogram is handed back clean for the next build,
// which saves a memset node per build).
template <bool CLEAR>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_kernel(int32_t* in, int32_t* __restrict__ out, int32_t n, uint64_t* agg, uint64_t* gpre,
            unsigned int* ticket, uint32_t epoch, int num_tiles, int32_t* total_copy) {
  __shared__ int s_tile;
  __shared__ int s_warp_sum[SCAN_THREADS / PG_WARP];
  __shared__ int s_red[SCAN_THREADS / PG_WARP];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    int t = (int)atomicAdd(ticket, 1u);
    if (t == num_tiles - 1) atomicExch(ticket, 0u);  // every ticket is out: re-arm for the next scan
    s_tile = t;
  }
  __syncthreads();
  const int tile = s_tile;
  const int64_t base = (int64_t)tile * SCAN_TILE + (int64_t)tid * SCAN_ITEMS;

  int v[SCAN_ITEMS];
  if (base + SCAN_ITEMS <= n) {
    int4* p = reinterpret_cast<int4*>(in + base);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS / 4; ++i) {
      int4 q = p[i];
      v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
    }
    if (CLEAR) {
#pragma unroll
      for (int i = 0; i < SCAN_ITEMS / 4; ++i) p[i] = make_int4(0, 0, 0, 0);
    }
  } else {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
      v[i] = (base + i < n) ? in[base + i] : 0;
      if (CLEAR && base + i < n) in[base + i] = 0;
    }
  }
  int tsum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) { int x = v[i]; v[i] = tsum; tsum += x; }
  int incl = tsum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += y;
  }
  if (lane == 31) s_warp_sum[warp] = incl;
  __syncthreads();
  int warp_off = 0, tile_sum = 0;
#pragma unroll
  for (int w = 0; w < SCAN_THREADS / PG_WARP; ++w) {
    int s = s_warp_sum[w];
    if (w < warp) warp_off += s;
    tile_sum += s;
  }
  const int thread_off = warp_off + incl - tsum;
  if (tid == 0) st_relaxed_u64(&agg[tile], pack_desc(epoch, tile_sum));

  // prefix of this tile = prefix of the previous group + aggregates of the earlier tiles of my group
  const int group = tile / SCAN_GROUP, first = group * SCAN_GROUP;
  int contrib = 0;
  const int pred = first + tid;
  if (pred < tile) {
    uint64_t w;
    do { w = ld_relaxed_u64(&agg[pred]); } while ((uint32_t)(w >> 32) != epoch);
    contrib = (int32_t)(uint32_t)w;
  }
  if (tid == SCAN_THREADS - 1 && group > 0) {   // this lane never has a predecessor (pred >= first + 255 >= tile)
    uint64_t w;
    do { w = ld_relaxed_u64(&gpre[group - 1]); } while ((uint32_t)(w >> 32) != epoch);
    contrib = (int32_t)(uint32_t)w;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
  if (lane == 0) s_red[warp] = contrib;
  __syncthreads();
  int prefix = 0;
#pragma unroll
  for (int w = 0; w < SCAN_THREADS / PG_WARP; ++w) prefix += s_red[w];
  if (tid == 0) {
    if (tile == first + SCAN_GROUP - 1) st_relaxed_u64(&gpre[group], pack_desc(epoch, prefix + tile_sum));
    if (tile == num_tiles - 1) {
      out[n] = prefix + tile_sum;
      if (total_copy) *total_copy = prefix + tile_sum;
    }
  }
  const int off = prefix + thread_off;
  if (base + SCAN_ITEMS <= n) {
    int4* p = reinterpret_cast<int4*>(out + base);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS / 4; ++i)
      p[i] = make_int4(v[4 * i] + off, v[4 * i + 1] + off, v[4 * i + 2] + off, v[4 * i + 3] + off);
  } else {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
      if (base + i < n) out[base + i] = v[i] + off;
  }
}

__global__ void scan_empty_kernel(int32_t* out, int32_t* total_copy) {
  out[0] = 0;
  if (total_copy) *total_copy = 0;
}

}  // namespace

int pg_scan_i32(pg_handle* h, const int32_t* in, int32_t* out, int32_t n, cudaStream_t s, int32_t* total_copy,
                bool clear_in) {
  if (n < 0) return pg_set_error(h, PG_ERR_INVALID, "scan: n < 0");
  if (n == 0) {
    PG_LAUNCH(h, s, "scan_empty_kernel", scan_empty_kernel<<<1, 1, 0, s>>>(out, total_copy));
    PG_LAUNCH_CHECK(h);
    return PG_OK;
  }
  if (((uintptr_t)in & 15) || ((uintptr_t)out & 15))
    return pg_set_error(h, PG_ERR_INVALID, "scan: in/out must be 16-byte aligned");
  const int num_tiles = pg_div_up(n, SCAN_TILE);
  const int num_groups = pg_div_up(num_tiles, SCAN_GROUP);
  const size_t need = 256 + (size_t)(num_tiles + num_groups + 2) * sizeof(uint64_t);
  if (need > h->scan_state.cap) {
    int rc = pg_reserve(h, h->scan_state, need);
    if (rc) return rc;
    PG_CUDA(h, cudaMemsetAsync(h->scan_state.p, 0, h->scan_state.cap, s));
  }
  h->scan_epoch += 1;
  if (h->scan_epoch == 0xffffffffu) {  // epoch 0 means "never written": start over on a clean slate
    PG_CUDA(h, cudaMemsetAsync(h->scan_state.p, 0, h->scan_state.cap, s));
    h->scan_epoch = 1;
  }
  unsigned int* ticket = (unsigned int*)h->scan_state.p;
  uint64_t* agg = (uint64_t*)((char*)h->scan_state.p + 256);
  uint64_t* gpre = agg + num_tiles;
  if (clear_in)
    PG_LAUNCH(h, s, "scan_kernel", scan_kernel<true><<<num_tiles, SCAN_THREADS, 0, s>>>(const_cast<int32_t*>(in), out, n, agg, gpre, ticket, h->scan_epoch, num_tiles, total_copy));
  else
    PG_LAUNCH(h, s, "scan_kernel", scan_kernel<false><<<num_tiles, SCAN_THREADS, 0, s>>>(const_cast<int32_t*>(in), out, n, agg, gpre, ticket, h->scan_epoch, num_tiles, total_copy));
  PG_LAUNCH_CHECK(h);
  return PG_OK;
}

extern "C" int pg_exclusive_scan_i32(pg_handle* h, const int32_t* in, int32_t* out, int32_t n, pg_stream stream) {
  if (!h) return PG_ERR_INVALID;
  PG_CUDA(h, cudaSetDevice(h->device));
  h->last_stream = (cudaStream_t)stream;
  return pg_scan_i32(h, in, out, n, (cudaStream_t)stream, nullptr, false);
}
