// pointnet2.cu — sampling / grouping ops of PointNet++ (FPS, ball query, group, gather + gradients,
// three_nn / three_interpolate) for sm_100a.
//
//  * FPS keeps the whole cloud state on chip: coordinates and the running min-distance of every point
//    live in registers (P points per thread), the cloud is mirrored in shared memory only to broadcast
//    the last selected point.  One round = P fused distance updates + a 2-instruction warp arg-max
//    (REDUX on the packed key) + ONE block barrier (double-buffered partials), instead of the
//    reference's 10 barriers and global-memory temp round trip.  The selection order of the reference
//    (max distance; ties -> the point whose reference thread id k mod BS is smallest in BIT-REVERSED
//    order — that is what its shared-memory tournament with "lower slot keeps ties" does — then the
//    smallest k; points with |p|^2 <= 1e-3 never touched) is encoded in a 64-bit sort key so any thread
//    layout reproduces it bit for bit.
//  * ball query runs a warp per centroid: 64 candidates per step on the packed fp32 pipe, ballots give
//    the in-order write positions, the scan stops as soon as nsample hits are found.
//  * group_points stages the gathered rows in shared memory (random 4-byte global gathers would cost a
//    wavefront per sector) and writes float4 along nsample — the op is output-write bound.
//  * scatter gradients (group / gather / three_interpolate) gather through a deterministic CSR (csr.cuh).
#include <cstdlib>

#include "common.cuh"
#include "csr.cuh"

namespace geoa3 {

// ============================================================================ FPS
constexpr int FPS_MAX_WARPS = 32;

// Distance used by the sampling kernels.  MODE 0: the pointnet2 squared distance (y,x,z contraction).  MODE 1/2: the
// Euclidean NORM of Lib/utility.py:183 (`torch.norm(diff, dim=1)`): sqrt of the sum of squares, accumulated as an
// fma chain in x,y,z order (1) or as separately rounded squares (2) — the arg-max (and its ties) is then taken on
// the same rounded values the reference compares.
template <int MODE>
__device__ __forceinline__ float fps_dist(float px, float py, float pz, float qx, float qy, float qz) {
  if constexpr (MODE == 0) return dist2_pn2(px, py, pz, qx, qy, qz);
  const float dx = __fsub_rn(px, qx), dy = __fsub_rn(py, qy), dz = __fsub_rn(pz, qz);
  if constexpr (MODE == 1) return __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))));
  return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
}

// reference block size: opt_n_threads(n) = clamp(2^floor(log2 n), 1, 512) (cuda_utils.h:13-19)
static int ref_fps_block(int n) {
  int p = 1;
  while (p * 2 <= n && p < 512) p *= 2;
  return p;
}

template <int P, int MODE>
__global__ void __launch_bounds__(1024)
fps_kernel(const float* __restrict__ xyz, int n, int m, int ref_bs, int ref_bits, int32_t* __restrict__ idxs,
           const int32_t* __restrict__ start) {
  extern __shared__ __align__(16) float s_xyz[];  // [3][n] SoA mirror of the cloud
  __shared__ unsigned long long s_part[2][FPS_MAX_WARPS];

  const int cloud = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  const int lane = tid & 31, w = tid >> 5, nw = T >> 5;
  const float* p = xyz + (size_t)cloud * n * 3;
  int32_t* out = idxs + (size_t)cloud * m;

  // start != nullptr: plain FPS of Lib/utility.py:175-187 (given first pick, no frozen points, ties -> lowest index)
  const bool plain = start != nullptr;
  float px[P], py[P], pz[P], temp[P];
  unsigned tk[P];  // low word of the sort key; 0 => point is frozen (skipped) or out of range
#pragma unroll
  for (int t = 0; t < P; ++t) {
    const int k = tid + t * T;
    px[t] = py[t] = pz[t] = 0.f;
    temp[t] = plain ? 3.0e38f : 1e10f;
    tk[t] = 0u;
    if (k < n) {
      px[t] = p[k * 3]; py[t] = p[k * 3 + 1]; pz[t] = p[k * 3 + 2];
      s_xyz[k] = px[t]; s_xyz[n + k] = py[t]; s_xyz[2 * n + k] = pz[t];
      const float mag = __fmaf_rn(pz[t], pz[t], __fmaf_rn(px[t], px[t], __fmul_rn(py[t], py[t])));  // y,x,z (SASS)
      // reference tournament (sampling_gpu.cu:115-168): at stride s slot t keeps ties against slot t+s, so
      // among equal distances the winner has the smallest bit-reversed thread id; inside one reference
      // thread (k = tid, tid+BS, ...) the strict '>' keeps the smallest k.
      const unsigned rtid = ref_bits ? (__brev((unsigned)(k % ref_bs)) >> (32 - ref_bits)) : 0u;
      if (plain || !((double)mag <= 1e-3)) tk[t] = ~((rtid << 20) | (unsigned)k);
    }
  }
  int old = plain ? min(max(start[cloud], 0), n - 1) : 0;
  if (tid == 0) out[0] = old;
  __syncthreads();

  for (int j = 1; j < m; ++j) {
    const float x1 = s_xyz[old], y1 = s_xyz[n + old], z1 = s_xyz[2 * n + old];
    // branch-free update: frozen / out-of-range points carry tk == 0 and contribute the minimal key (0,0)
    float hmax = 0.f;
#pragma unroll
    for (int t = 0; t < P; ++t) {
      const float d = fps_dist<MODE>(px[t], py[t], pz[t], x1, y1, z1);
      temp[t] = fminf(d, temp[t]);
      hmax = fmaxf(hmax, tk[t] != 0u ? temp[t] : 0.f);  // distances are >= +0: float order == bit order
    }
    unsigned bl = 0u;
#pragma unroll
    for (int t = 0; t < P; ++t) bl = max(bl, temp[t] == hmax ? tk[t] : 0u);
    const unsigned bh = __float_as_uint(hmax);
    // warp arg-max of the 64-bit key (bh, bl) in two REDUX steps
    const unsigned mh = __reduce_max_sync(0xffffffffu, bh);
    const unsigned ml = __reduce_max_sync(0xffffffffu, bh == mh ? bl : 0u);
    if (lane == 0) s_part[j & 1][w] = ((unsigned long long)mh << 32) | ml;
    __syncthreads();
    // cross-warp arg-max, lane-parallel (nw <= 32 partials), again two REDUX steps; every warp redundantly
    const unsigned long long pv = lane < nw ? s_part[j & 1][lane] : 0ull;
    const unsigned ph = (unsigned)(pv >> 32), pl = (unsigned)pv;
    const unsigned gh = __reduce_max_sync(0xffffffffu, ph);
    const unsigned lo = __reduce_max_sync(0xffffffffu, ph == gh ? pl : 0u);
    old = lo != 0u ? (int)((~lo) & 0xFFFFFu) : 0;
    if (tid == 0) out[j] = old;
  }
}

// Single-warp variant for n <= 1024: a whole cloud lives in the registers of ONE warp (P points per lane, packed
// two per register pair), so a round needs no block barrier and no shared-memory exchange at all.  A round is
// issue-bound on that one warp (IPC ~0.55), so the loop is cut to ~5 instructions per point: packed distance
// (FADD2/FMUL2/FFMA2, pointnet2 contraction order), one FMNMX per point for the running minimum, FMNMX3 trees for
// per-group maxima, and the tie-break keys are only inspected inside the group(s) that attain the warp-wide
// maximum (found with one REDUX).  Frozen / padding points keep temp = 0 and key 0, so they need no select.
template <int P>
__global__ void __launch_bounds__(32)
fps_warp_kernel(const float* __restrict__ xyz, int n, int m, int ref_bs, int ref_bits, int32_t* __restrict__ idxs) {
  static_assert(P % 8 == 0, "groups of 8 points");
  extern __shared__ __align__(16) float s_xyz[];  // [3][n] SoA mirror (broadcast of the last pick)
  const int cloud = blockIdx.x, lane = threadIdx.x;
  const float* p = xyz + (size_t)cloud * n * 3;
  int32_t* out = idxs + (size_t)cloud * m;
  constexpr int H = P / 2, G = P / 8;
  float2 px[H], py[H], pz[H], temp[H];
  unsigned tk[P];
  {
    float lx[P], ly[P], lz[P];
#pragma unroll
    for (int t = 0; t < P; ++t) {  // all loads first (independent), then the per-point setup
      const int k = lane + t * 32;
      const bool ok = k < n;
      lx[t] = ok ? p[k * 3] : 0.f; ly[t] = ok ? p[k * 3 + 1] : 0.f; lz[t] = ok ? p[k * 3 + 2] : 0.f;
    }
#pragma unroll
    for (int t = 0; t < P; ++t) {
      const int k = lane + t * 32;
      tk[t] = 0u;
      float t0 = 0.f;
      if (k < n) {
        s_xyz[k] = lx[t]; s_xyz[n + k] = ly[t]; s_xyz[2 * n + k] = lz[t];
        const float mag = __fmaf_rn(lz[t], lz[t], __fmaf_rn(lx[t], lx[t], __fmul_rn(ly[t], ly[t])));
        const unsigned rtid = ref_bits ? (__brev((unsigned)(k % ref_bs)) >> (32 - ref_bits)) : 0u;
        if (!((double)mag <= 1e-3)) { tk[t] = ~((rtid << 20) | (unsigned)k); t0 = 1e10f; }
      }
      if (t & 1) { px[t >> 1].y = lx[t]; py[t >> 1].y = ly[t]; pz[t >> 1].y = lz[t]; temp[t >> 1].y = t0; }
      else { px[t >> 1].x = lx[t]; py[t >> 1].x = ly[t]; pz[t >> 1].x = lz[t]; temp[t >> 1].x = t0; }
    }
  }
  if (lane == 0) out[0] = 0;
  __syncwarp();
  int old = 0;
  for (int j = 1; j < m; ++j) {
    const float x1 = -s_xyz[old], y1 = -s_xyz[n + old], z1 = -s_xyz[2 * n + old];
    const float2 nx = make_float2(x1, x1), ny = make_float2(y1, y1), nz = make_float2(z1, z1);
    float hg[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float mx = 0.f;
#pragma unroll
      for (int h = g * 4; h < g * 4 + 4; ++h) {
        const float2 d = dist2x2_pn2(px[h], py[h], pz[h], nx, ny, nz);
        temp[h].x = fminf(d.x, temp[h].x);  // frozen / padding points sit at 0 forever
        temp[h].y = fminf(d.y, temp[h].y);
        mx = fmaxf(mx, fmaxf(temp[h].x, temp[h].y));
      }
      hg[g] = mx;
    }
    float hmax = hg[0];
#pragma unroll
    for (int g = 1; g < G; ++g) hmax = fmaxf(hmax, hg[g]);
    const float gmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(hmax)));  // >= +0: bit order
    unsigned bl = 0u;
    if (hmax == gmax) {  // only the lane(s) holding the maximum look at their tie-break keys
#pragma unroll
      for (int g = 0; g < G; ++g)
        if (hg[g] == gmax) {
#pragma unroll
          for (int h = g * 4; h < g * 4 + 4; ++h) {
            bl = max(bl, temp[h].x == gmax ? tk[2 * h] : 0u);
            bl = max(bl, temp[h].y == gmax ? tk[2 * h + 1] : 0u);
          }
        }
    }
    const unsigned lo = __reduce_max_sync(0xffffffffu, bl);
    old = lo != 0u ? (int)((~lo) & 0xFFFFFu) : 0;
    if (lane == 0) out[j] = old;
  }
}

// Four-warp variant for n <= 2048 (P = 2..16 points per lane).  ubench/fps_probe.cu shows that a round of the
// single-warp kernel is a ~600-cycle LATENCY chain (distance/min/max 290, tie-break search 200, REDUX + store 80,
// LDS 37) that barely shrinks with fewer points per lane, i.e. one warp cannot hide its own dependent chains.
// Here a cloud is spread over the 4 schedulers of an SM (P/4 as many points per lane => short chains), the
// per-lane tie-break key is computed unconditionally next to the first REDUX instead of after it (no divergent
// search), and the 4 warp partials are merged through a double-buffered shared slot with ONE barrier per round.
constexpr int FPSQ_WARPS = 4;

template <int P, int MODE>
__global__ void __launch_bounds__(FPSQ_WARPS * 32)
fps_quad_kernel(const float* __restrict__ xyz, int n, int m, int ref_bs, int ref_bits, int32_t* __restrict__ idxs,
                const int32_t* __restrict__ start) {
  static_assert(P % 2 == 0, "points are packed in pairs");
  extern __shared__ __align__(16) float s_xyz[];  // [3][n] SoA mirror (broadcast of the last pick), then int s_out[m]
  __shared__ __align__(16) unsigned long long s_part[2][FPSQ_WARPS];
  // picks are collected in shared memory and written out once: a global store per round sits in front of the
  // next round's block barrier (which orders global accesses too) and costs ~100 cycles per round (fps_probe)
  int* s_out = reinterpret_cast<int*>(s_xyz + 3 * n);
  constexpr int T = FPSQ_WARPS * 32, H = P / 2;
  const int cloud = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float* p = xyz + (size_t)cloud * n * 3;
  int32_t* out = idxs + (size_t)cloud * m;
  const bool plain = start != nullptr;  // Lib/utility.py:175-187 semantics (see fps_kernel)
  float2 px[H], py[H], pz[H], temp[H];
  unsigned tk[P];
#pragma unroll
  for (int t = 0; t < P; ++t) {
    const int k = tid + t * T;
    float x = 0.f, y = 0.f, z = 0.f, t0 = 0.f;
    tk[t] = 0u;
    if (k < n) {
      x = p[k * 3]; y = p[k * 3 + 1]; z = p[k * 3 + 2];
      s_xyz[k] = x; s_xyz[n + k] = y; s_xyz[2 * n + k] = z;
      const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
      const unsigned rtid = ref_bits ? (__brev((unsigned)(k % ref_bs)) >> (32 - ref_bits)) : 0u;
      if (plain || !((double)mag <= 1e-3)) {  // else frozen: temp 0, key 0
        tk[t] = ~((rtid << 20) | (unsigned)k);
        t0 = plain ? 3.0e38f : 1e10f;
      }
    }
    if (t & 1) { px[t >> 1].y = x; py[t >> 1].y = y; pz[t >> 1].y = z; temp[t >> 1].y = t0; }
    else { px[t >> 1].x = x; py[t >> 1].x = y; pz[t >> 1].x = z; temp[t >> 1].x = t0; }
  }
  int old = plain ? min(max(start[cloud], 0), n - 1) : 0;
  if (tid == 0) s_out[0] = old;
  __syncthreads();
  for (int j = 1; j < m; ++j) {
    const float x1 = -s_xyz[old], y1 = -s_xyz[n + old], z1 = -s_xyz[2 * n + old];
    const float2 nx = make_float2(x1, x1), ny = make_float2(y1, y1), nz = make_float2(z1, z1);
    float hm[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float2 d;
      if constexpr (MODE == 0) {
        d = dist2x2_pn2(px[h], py[h], pz[h], nx, ny, nz);
      } else {
        d.x = fps_dist<MODE>(px[h].x, py[h].x, pz[h].x, -x1, -y1, -z1);
        d.y = fps_dist<MODE>(px[h].y, py[h].y, pz[h].y, -x1, -y1, -z1);
      }
      temp[h].x = fminf(d.x, temp[h].x);  // frozen / padding points sit at 0 forever
      temp[h].y = fminf(d.y, temp[h].y);
      hm[h] = fmaxf(temp[h].x, temp[h].y);
    }
#pragma unroll
    for (int st = 1; st < H; st <<= 1)
#pragma unroll
      for (int h = 0; h + st < H; h += 2 * st) hm[h] = fmaxf(hm[h], hm[h + st]);
    const float hmax = hm[0];
    const unsigned gw = __reduce_max_sync(0xffffffffu, __float_as_uint(hmax));  // distances >= +0: bit order
    // this lane's best tie-break key among its points at hmax — independent of the REDUX above
    unsigned kk[H];
#pragma unroll
    for (int h = 0; h < H; ++h)
      kk[h] = max(temp[h].x == hmax ? tk[2 * h] : 0u, temp[h].y == hmax ? tk[2 * h + 1] : 0u);
#pragma unroll
    for (int st = 1; st < H; st <<= 1)
#pragma unroll
      for (int h = 0; h + st < H; h += 2 * st) kk[h] = max(kk[h], kk[h + st]);
    const unsigned lw = __reduce_max_sync(0xffffffffu, __float_as_uint(hmax) == gw ? kk[0] : 0u);
    if (lane == 0) s_part[j & 1][w] = ((unsigned long long)gw << 32) | lw;
    __syncthreads();
    const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(&s_part[j & 1][0]);
    const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(&s_part[j & 1][2]);
    const unsigned long long ab = a.x > a.y ? a.x : a.y, cd = b.x > b.y ? b.x : b.y;
    const unsigned lo = (unsigned)(ab > cd ? ab : cd);
    old = lo != 0u ? (int)((~lo) & 0xFFFFFu) : 0;
    if (tid == 0) s_out[j] = old;
  }
  __syncthreads();
  for (int i = tid; i < m; i += T) out[i] = s_out[i];
}

// ============================================================================ ball query
constexpr int BQ_WARPS = 8;
constexpr int BQ_PER_WARP = 8;  // centroids handled sequentially by one warp
constexpr int BQ_MAX_NS = 256;

__global__ void __launch_bounds__(BQ_WARPS * 32)
ball_query_kernel(const float* __restrict__ new_xyz, const float* __restrict__ xyz, int n, int m, float radius,
                  int nsample, int32_t* __restrict__ idx) {
  extern __shared__ __align__(16) float s_pts[];  // [3][n64] SoA, padded to a multiple of 64 with +inf
  __shared__ int s_row[BQ_WARPS][BQ_MAX_NS];
  const int cloud = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int n64 = (n + 63) & ~63;
  const float* p = xyz + (size_t)cloud * n * 3;
  float* sx = s_pts; float* sy = s_pts + n64; float* sz = s_pts + 2 * n64;
  for (int i = tid; i < n64; i += BQ_WARPS * 32) {
    const bool ok = i < n;
    sx[i] = ok ? p[i * 3] : __int_as_float(0x7f800000);  // padding is never inside a ball
    sy[i] = ok ? p[i * 3 + 1] : 0.f;
    sz[i] = ok ? p[i * 3 + 2] : 0.f;
  }
  __syncthreads();
  const float r2 = __fmul_rn(radius, radius);
  const unsigned lt = (1u << lane) - 1u;
  const int c_beg = (blockIdx.x * BQ_WARPS + w) * BQ_PER_WARP;
  for (int c = c_beg; c < min(m, c_beg + BQ_PER_WARP); ++c) {
    int32_t* o = idx + ((size_t)cloud * m + c) * nsample;
    // rows of more than BQ_MAX_NS samples (uniform_loss on dense clouds) are collected in the output row itself
    int* row = nsample <= BQ_MAX_NS ? s_row[w] : o;
    const float* q = new_xyz + ((size_t)cloud * m + c) * 3;
    const float qx = -q[0], qy = -q[1], qz = -q[2];
    const float2 nqx = make_float2(qx, qx), nqy = make_float2(qy, qy), nqz = make_float2(qz, qz);
    int cnt = 0;
    for (int j0 = 0; j0 < n64 && cnt < nsample; j0 += 64) {
      const int ja = j0 + lane, jb = j0 + 32 + lane;
      const float2 d = dist2x2_pn2(make_float2(sx[ja], sx[jb]), make_float2(sy[ja], sy[jb]),
                                   make_float2(sz[ja], sz[jb]), nqx, nqy, nqz);
      const bool ha = d.x < r2, hb = d.y < r2;
      const unsigned ba = __ballot_sync(0xffffffffu, ha), bb = __ballot_sync(0xffffffffu, hb);
      const int pa = cnt + __popc(ba & lt);
      if (ha && pa < nsample) row[pa] = ja;
      cnt += __popc(ba);
      const int pb = cnt + __popc(bb & lt);
      if (hb && pb < nsample) row[pb] = jb;
      cnt += __popc(bb);
    }
    __syncwarp();
    cnt = min(cnt, nsample);
    const int first = cnt > 0 ? row[0] : 0;  // first-hit fill; a centroid with no hit keeps zeros
    for (int l = lane; l < nsample; l += 32) o[l] = l < cnt ? row[l] : first;
    __syncwarp();
  }
}

// ---- thread-per-centroid variant (n <= 65535).  The warp-per-centroid kernel above pays a ballot + prefix per 64
// candidates; balls are small (typically ~1 % of the cloud), so almost every candidate is a miss and the cheap
// thing is the knn inner loop: one thread = one centroid, candidates broadcast from shared memory as float4,
// packed distance, pass bit = sign of (d - r^2) shifted into a 32-candidate mask.  Hits (rare) go to a per-thread
// shared row in scan order; rows are written out coalesced with the first-hit padding.
constexpr int BQT_THREADS = 128;
constexpr int BQT_CHUNK = 1024;

__global__ void __launch_bounds__(BQT_THREADS)
ball_query_thread_kernel(const float* __restrict__ new_xyz, const float* __restrict__ xyz, int n, int m, float radius,
                         int nsample, int32_t* __restrict__ idx) {
  __shared__ __align__(16) float sx[BQT_CHUNK];
  __shared__ __align__(16) float sy[BQT_CHUNK];
  __shared__ __align__(16) float sz[BQT_CHUNK];
  extern __shared__ uint16_t s_hits[];  // [BQT_THREADS][pitch], pitch = nsample + 2 (skews the banks)
  const int cloud = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
  const int c = blockIdx.x * BQT_THREADS + tid;
  const int pitch = nsample + 2;
  const float* q = new_xyz + ((size_t)cloud * m + min(c, m - 1)) * 3;
  const float qx = -q[0], qy = -q[1], qz = -q[2];
  const float2 nqx = make_float2(qx, qx), nqy = make_float2(qy, qy), nqz = make_float2(qz, qz);
  const float r2 = __fmul_rn(radius, radius);
  const float2 nr2 = make_float2(-r2, -r2);
  const float* p = xyz + (size_t)cloud * n * 3;
  uint16_t* mine = s_hits + (size_t)tid * pitch;
  int cnt = c < m ? 0 : nsample;  // threads past the last centroid count as full
  for (int c0 = 0; c0 < n; c0 += BQT_CHUNK) {
    const int cn = min(BQT_CHUNK, n - c0), cn32 = (cn + 31) & ~31;
    __syncthreads();
    for (int i = tid; i < cn32; i += BQT_THREADS) {
      const bool ok = i < cn;
      sx[i] = ok ? p[(size_t)(c0 + i) * 3] : __int_as_float(0x7f800000);  // padding is never inside a ball
      sy[i] = ok ? p[(size_t)(c0 + i) * 3 + 1] : 0.f;
      sz[i] = ok ? p[(size_t)(c0 + i) * 3 + 2] : 0.f;
    }
    __syncthreads();
    for (int j = 0; j < cn32; j += 32) {
      if (__all_sync(0xffffffffu, cnt >= nsample)) break;
      unsigned mask = 0u;
#pragma unroll 2
      for (int u = 0; u < 32; u += 8) {
        const float4 cxa = *reinterpret_cast<const float4*>(sx + j + u), cxb = *reinterpret_cast<const float4*>(sx + j + u + 4);
        const float4 cya = *reinterpret_cast<const float4*>(sy + j + u), cyb = *reinterpret_cast<const float4*>(sy + j + u + 4);
        const float4 cza = *reinterpret_cast<const float4*>(sz + j + u), czb = *reinterpret_cast<const float4*>(sz + j + u + 4);
        // d < r2  <=>  d - r2 < 0 exactly (a float difference is zero only for equal operands; no flush-to-zero)
        const float2 s01 = __fadd2_rn(dist2x2_pn2(make_float2(cxa.x, cxa.y), make_float2(cya.x, cya.y), make_float2(cza.x, cza.y), nqx, nqy, nqz), nr2);
        const float2 s23 = __fadd2_rn(dist2x2_pn2(make_float2(cxa.z, cxa.w), make_float2(cya.z, cya.w), make_float2(cza.z, cza.w), nqx, nqy, nqz), nr2);
        const float2 s45 = __fadd2_rn(dist2x2_pn2(make_float2(cxb.x, cxb.y), make_float2(cyb.x, cyb.y), make_float2(czb.x, czb.y), nqx, nqy, nqz), nr2);
        const float2 s67 = __fadd2_rn(dist2x2_pn2(make_float2(cxb.z, cxb.w), make_float2(cyb.z, cyb.w), make_float2(czb.z, czb.w), nqx, nqy, nqz), nr2);
        mask = __funnelshift_l(__float_as_uint(s01.x), mask, 1);
        mask = __funnelshift_l(__float_as_uint(s01.y), mask, 1);
        mask = __funnelshift_l(__float_as_uint(s23.x), mask, 1);
        mask = __funnelshift_l(__float_as_uint(s23.y), mask, 1);
        mask = __funnelshift_l(__float_as_uint(s45.x), mask, 1);
        mask = __funnelshift_l(__float_as_uint(s45.y), mask, 1);
        mask = __funnelshift_l(__float_as_uint(s67.x), mask, 1);
        mask = __funnelshift_l(__float_as_uint(s67.y), mask, 1);
      }
      while (mask && cnt < nsample) {  // bit 31 is candidate j: ascending index = descending bit
        const int t = __clz(mask);
        mask &= ~(0x80000000u >> t);
        mine[cnt++] = (uint16_t)(c0 + j + t);
      }
    }
  }
  __syncwarp();
  // rows of this warp's 32 centroids, one coalesced row at a time; first-hit fill, an empty ball keeps zeros
  const int wbase = tid & ~31;
  for (int r = 0; r < 32; ++r) {
    const int cr = blockIdx.x * BQT_THREADS + wbase + r;
    if (cr >= m) break;
    const int cnt_r = __shfl_sync(0xffffffffu, cnt, r);
    const uint16_t* row = s_hits + (size_t)(wbase + r) * pitch;
    const int first = cnt_r > 0 ? (int)row[0] : 0;
    int32_t* o = idx + ((size_t)cloud * m + cr) * nsample;
    for (int l = lane; l < nsample; l += 32) o[l] = l < cnt_r ? (int)row[l] : first;
  }
}

// ============================================================================ group_points (forward)
constexpr int GP_THREADS = 256;
constexpr int GP_CC = 16;        // channels staged per CTA
constexpr int GP_ETILE = 8192;   // (j,k) positions per CTA

__global__ void __launch_bounds__(GP_THREADS)
group_points_kernel(const float* __restrict__ points, const int32_t* __restrict__ idx, int c, int n, int E,
                    float* __restrict__ out, int gcc) {
  extern __shared__ __align__(16) float s_rows[];  // [cc][n]; gcc = channels per CTA (GP_CC unless n is very large)
  const int cloud = blockIdx.z, c0 = blockIdx.y * gcc, cc = min(gcc, c - c0);
  const int tid = threadIdx.x;
  const float* src = points + ((size_t)cloud * c + c0) * n;
  // The cc feature rows of this CTA are ONE contiguous block of cc*n floats: when it is 16-byte aligned and sized,
  // a single TMA bulk copy (cp.async.bulk, completion on an mbarrier) stages it — no per-thread load/store loop.
  const unsigned bytes = (unsigned)(cc * n) * 4u;
  __shared__ __align__(8) unsigned long long gp_bar;
  const bool bulk = (bytes & 15u) == 0u && (reinterpret_cast<uintptr_t>(src) & 15u) == 0u;
  if (bulk) {
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&gp_bar);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the init must be visible to the async proxy
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"((unsigned)__cvta_generic_to_shared(s_rows)), "l"(src), "r"(bytes), "r"(bar) : "memory");
    }
    __syncthreads();  // everybody sees the initialised barrier before polling it
    unsigned done;
    do {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(bar) : "memory");
    } while (!done);
  } else {
    for (int i = tid; i < cc * n; i += GP_THREADS) s_rows[i] = src[i];
    __syncthreads();
  }
  const int32_t* ix = idx + (size_t)cloud * E;
  float* dst = out + ((size_t)cloud * c + c0) * E;
  const int e_beg = blockIdx.x * GP_ETILE, e_end = min(E, e_beg + GP_ETILE);
  if ((E & 3) == 0) {
    for (int e = e_beg + tid * 4; e < e_end; e += GP_THREADS * 4) {
      const int4 i4 = *reinterpret_cast<const int4*>(ix + e);
      for (int l = 0; l < cc; ++l) {
        const float* r = s_rows + l * n;
        const float4 v = make_float4(r[i4.x], r[i4.y], r[i4.z], r[i4.w]);
        __stcs(reinterpret_cast<float4*>(dst + (size_t)l * E + e), v);  // streaming: output is write-once
      }
    }
  } else {
    for (int e = e_beg + tid; e < e_end; e += GP_THREADS) {
      const int i1 = ix[e];
      for (int l = 0; l < cc; ++l) dst[(size_t)l * E + e] = s_rows[l * n + i1];
    }
  }
}

// ============================================================================ CSR in global workspace
// workspace layout per cloud: offs[n+1] ints, then ent[E] ints
constexpr int CSRG_THREADS = 512;

__global__ void __launch_bounds__(CSRG_THREADS)
csr_global_kernel(const int32_t* __restrict__ idx, int E, int n, int W, int* __restrict__ ws,
                  const int* __restrict__ fast_flag) {
  extern __shared__ __align__(16) int s_whist[];
  __shared__ int scan_scratch[CSRG_THREADS / 32 + 1];
  if (fast_flag && *fast_flag == 0) return;  // the forward-order kernel handles this call (see fwd_order_ok_kernel)
  const int cloud = blockIdx.x;
  int* offs = ws + (size_t)cloud * (n + 1 + E);
  int* ent = offs + n + 1;
  build_csr<CSRG_THREADS, int, int>(idx + (size_t)cloud * E, E, n, 0u, offs, s_whist, W, ent, scan_scratch);
}

// grad_points[b][c][p] = sum over segment(p) of grad_out[b][c][e] (ascending e), optional weights
constexpr int GG_THREADS = 256;

__global__ void __launch_bounds__(GG_THREADS)
csr_gather_grad_kernel(const float* __restrict__ grad_out, const float* __restrict__ weight, const int* __restrict__ ws,
                       int c, int n, int E, int e_div, int tile, float* __restrict__ grad_points,
                       const int* __restrict__ fast_flag) {
  if (fast_flag && *fast_flag == 0) return;
  // grad_out rows have E/e_div entries per channel (e_div = 3 for three_interpolate, where the three
  // (idx,weight) slots of one output share one grad_out value); weight (nullable) is [b][E].
  extern __shared__ __align__(16) float s_go[];  // one tile of the row
  const int cloud = blockIdx.y, ch = blockIdx.x, tid = threadIdx.x;
  const int Eg = E / e_div;
  const int* offs = ws + (size_t)cloud * (n + 1 + E);
  const int* ent = offs + n + 1;
  const float* go = grad_out + ((size_t)cloud * c + ch) * Eg;
  const float* wt = weight ? weight + (size_t)cloud * E : nullptr;
  float* gp = grad_points + ((size_t)cloud * c + ch) * n;
  if (tile >= Eg) {  // whole row fits
    for (int i = tid; i < Eg; i += GG_THREADS) s_go[i] = go[i];
    __syncthreads();
    for (int p = tid; p < n; p += GG_THREADS) {
      float acc = 0.f;
      const int e1 = offs[p + 1];
      for (int t = offs[p]; t < e1; ++t) {
        const int e = ent[t];
        acc += wt ? s_go[e / e_div] * wt[e] : s_go[e];
      }
      gp[p] = acc;
    }
  } else {  // tiled over the source range; segments are ascending so each thread keeps a cursor
    constexpr int MAXP = 8;
    float acc[MAXP];
    int cur[MAXP];
#pragma unroll
    for (int q = 0; q < MAXP; ++q) {
      const int p = tid + q * GG_THREADS;
      acc[q] = 0.f;
      cur[q] = p < n ? offs[p] : 0;
    }
    for (int t0 = 0; t0 < Eg; t0 += tile) {
      const int tn = min(tile, Eg - t0);
      __syncthreads();
      for (int i = tid; i < tn; i += GG_THREADS) s_go[i] = go[t0 + i];
      __syncthreads();
#pragma unroll
      for (int q = 0; q < MAXP; ++q) {
        const int p = tid + q * GG_THREADS;
        if (p < n) {
          const int e1 = offs[p + 1];
          int t = cur[q];
          while (t < e1) {
            const int e = ent[t];
            const int g = e / e_div;
            if (g >= t0 + tn) break;
            acc[q] += wt ? s_go[g - t0] * wt[e] : s_go[g - t0];
            ++t;
          }
          cur[q] = t;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < MAXP; ++q) {
      const int p = tid + q * GG_THREADS;
      if (p < n) gp[p] = acc[q];
    }
  }
}

// Cooperative variant for many channels (the feature-grouping gradient, C >= 8): the CSR of one cloud is staged
// ONCE per CTA in shared memory (entries as uint16) and reused for GG2_CC channels; each channel's grad_out row
// is staged with 16-byte async copies; eight lanes share a target (lane-strided partial sums + a fixed shuffle
// tree), which bounds the damage of the long segments ball-query padding creates (the first index of a
// sparse ball is repeated up to nsample times).  Everything the inner loop touches is in shared memory.
// (Position-balanced segmented-reduction, double-buffered and bank-swizzled variants were measured slower on
// B200 at the SSG shapes — see DESIGN.md §7; this op is latency-bound per CTA, not bandwidth-bound, today.)
constexpr int GG2_THREADS = 256;
constexpr int GG2_CC = 8;

__global__ void __launch_bounds__(GG2_THREADS)
csr_gather_grad_coop_kernel(const float* __restrict__ grad_out, const int* __restrict__ ws, int c, int n, int E,
                            float* __restrict__ grad_points, const int* __restrict__ fast_flag) {
  extern __shared__ __align__(16) unsigned char gg2_smem[];
  if (fast_flag && *fast_flag == 0) return;
  float* s_go = reinterpret_cast<float*>(gg2_smem);                            // E floats
  int* s_offs = reinterpret_cast<int*>(gg2_smem + (size_t)E * 4);              // n+1 ints
  uint16_t* s_ent = reinterpret_cast<uint16_t*>(s_offs + ((n + 1 + 3) & ~3));  // E uint16
  const int cloud = blockIdx.y, c0 = blockIdx.x * GG2_CC, cc = min(GG2_CC, c - c0), tid = threadIdx.x;
  const int* offs = ws + (size_t)cloud * (n + 1 + E);
  const int* ent = offs + n + 1;
  for (int i = tid; i <= n; i += GG2_THREADS) s_offs[i] = offs[i];
  for (int i = tid; i < E; i += GG2_THREADS) s_ent[i] = (uint16_t)ent[i];
  const int sub = tid & 7, grp = tid >> 3;  // 8 lanes per target, 32 targets in flight per CTA
  for (int l = 0; l < cc; ++l) {
    const float* go = grad_out + ((size_t)cloud * c + c0 + l) * E;
    __syncthreads();  // previous channel fully consumed (and the CSR staged, first time round)
    for (int i = tid * 4; i < E; i += GG2_THREADS * 4) {
      const unsigned dst = (unsigned)__cvta_generic_to_shared(s_go + i);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(go + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    float* gp = grad_points + ((size_t)cloud * c + c0 + l) * n;
    for (int p = grp; p < n; p += GG2_THREADS / 8) {
      const int e1 = s_offs[p + 1];
      float acc = 0.f;
      for (int t = s_offs[p] + sub; t < e1; t += 8) acc += s_go[s_ent[t]];
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (sub == 0) gp[p] = acc;
    }
  }
}

// ============================================================================ group_points_grad, forward order
// Ball-query index rows have a known shape: a strictly ascending prefix of distinct hits followed by copies of
// the row's first index.  For such rows the scatter can run in FORWARD order with perfectly coalesced grad_out
// reads and still be deterministic: warp w owns a contiguous block of rows and a private accumulator array in
// shared memory; inside a 32-wide row slice all targets are distinct except the copies of the first index,
// which are summed by a fixed shuffle tree and added once.  Per target the order is (warp block, row, slice)
// + a fixed cross-warp reduction => bitwise reproducible, no atomics, no CSR, no sorting.
// fwd_order_ok_kernel verifies the shape on the device (violations counter); if any row deviates this kernel
// returns immediately and the generic CSR path (launched right after, predicated the other way) does the work.
__global__ void fwd_order_ok_kernel(const int32_t* __restrict__ idx, int rows, int ns, int n, int* __restrict__ violations) {
  // one pass, one load per entry: a row is fine iff r[0] is in range, the entries before the first repeat of r[0]
  // are strictly ascending (and < n), and everything from that repeat on equals r[0]
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int32_t* r = idx + (size_t)row * ns;
  const int p0 = r[0];
  bool bad = p0 < 0 || p0 >= n;
  bool in_pad = false;   // warp-uniform: the padding suffix has started in an earlier chunk
  int carry = p0;        // last entry of the previous chunk
  for (int k0 = 0; k0 < ns; k0 += 32) {
    const int k = k0 + lane;
    const bool live = k < ns;
    const int v = live ? r[k] : p0;
    int prev = __shfl_up_sync(0xffffffffu, v, 1);
    if (lane == 0) prev = carry;
    const unsigned pad = __ballot_sync(0xffffffffu, live && k >= 1 && v == p0);
    // first padding position of this chunk (or 32); entries at or after it must all be p0
    const int fp = in_pad ? 0 : (pad ? __ffs(pad) - 1 : 32);
    if (live && k >= 1) bad |= lane >= fp ? v != p0 : (!(v > prev) || v >= n);
    in_pad = in_pad || pad != 0u;
    carry = __shfl_sync(0xffffffffu, v, 31);
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicAdd(violations, 1);
}

constexpr int GF_WARPS = 8;
constexpr int GF_THREADS = GF_WARPS * 32;

template <int CC>
__global__ void __launch_bounds__(GF_THREADS)
group_grad_fwd_order_kernel(const float* __restrict__ grad_out, const int32_t* __restrict__ idx, int c, int n, int m,
                            int ns, float* __restrict__ grad_points, const int* __restrict__ violations) {
  extern __shared__ __align__(16) float gf_acc[];  // [GF_WARPS][CC][n]
  if (*violations != 0) return;
  const int cloud = blockIdx.y, c0 = blockIdx.x * CC, cc = min(CC, c - c0);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int i = tid; i < GF_WARPS * CC * n; i += GF_THREADS) gf_acc[i] = 0.f;
  __syncthreads();
  float* acc = gf_acc + (size_t)w * CC * n;
  const int E = m * ns;
  const int32_t* ix = idx + (size_t)cloud * E;
  const float* go = grad_out + ((size_t)cloud * c + c0) * E;
  const int rows_per_warp = (m + GF_WARPS - 1) / GF_WARPS;
  const int j_beg = min(m, w * rows_per_warp), j_end = min(m, j_beg + rows_per_warp);
  const int spr = (ns + 31) >> 5;                 // 32-wide slices per row
  const int n_slices = (j_end - j_beg) * spr;
  constexpr int D = 4;                            // slices fetched ahead: D*CC coalesced 128-byte loads in flight
  const float* gol[CC];
#pragma unroll
  for (int l = 0; l < CC; ++l) gol[l] = go + (size_t)min(l, cc - 1) * E;
  int row_e = j_beg * ns, k0 = 0;                 // running position of the next slice to fetch (no divisions)
  for (int s0 = 0; s0 < n_slices; s0 += D) {
    int pd[D], p0d[D];
    float vd[D][CC];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const bool live = s0 + d < n_slices;
      const int k = k0 + lane;
      const bool valid = live && k < ns;
      const int e = row_e + k;
      pd[d] = valid ? ix[e] : -1;
      p0d[d] = live ? ix[row_e] : -2;
#pragma unroll
      for (int l = 0; l < CC; ++l) vd[d][l] = valid ? __ldcs(gol[l] + e) : 0.f;
      k0 += 32;
      if (k0 >= ns) { k0 = 0; row_e += ns; }
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const int p = pd[d];
      const bool valid = p >= 0;
      const bool is0 = p == p0d[d];            // a copy of the row's first index (or that entry itself)
      const unsigned m0 = __ballot_sync(0xffffffffu, is0);
      const bool leader = is0 && (m0 & ((1u << lane) - 1u)) == 0u;   // lowest lane holding it
      const bool rmw = valid && (!is0 || leader);
      // control flow stays warp-uniform: the copies are ALWAYS summed by the fixed shuffle tree (a slice
      // without copies just sums one value); a uniform-looking branch here makes the compiler wrap every
      // shuffle in convergence barriers, which tripled the instruction count
#pragma unroll
      for (int l = 0; l < CC; ++l) {
        const float v = vd[d][l];
        const float t = warp_sum(is0 ? v : 0.f);
        if (rmw) acc[l * n + p] += leader ? t : v;   // channels >= cc mirror channel cc-1 and are never written out
      }
      __syncwarp();  // the next slice / row may touch the same targets from other lanes
    }
  }
  __syncthreads();
  float* gp = grad_points + ((size_t)cloud * c + c0) * n;
  for (int i = tid; i < cc * n; i += GF_THREADS) {
    float sacc = 0.f;
#pragma unroll
    for (int ww = 0; ww < GF_WARPS; ++ww) sacc += gf_acc[(size_t)ww * CC * n + i];
    gp[i] = sacc;
  }
}

// ---- row-structured variant: nsample = 32*SPR.  The generic kernel above spends its issue slots and LSU
// wavefronts (ncu: LSU data pipe 83 %) on one 5-step shuffle tree + leader vote per SLICE and on bank-conflicted
// 32-bit RMWs.  Here (a) the padding copies are summed once per ROW: position 0 of the row holds the first index
// itself, so lane 0 of slice 0 is always the leader and no vote is needed; (b) the CC channel sums share ONE
// butterfly (after the first exchange half the warp reduces channel 0 and the other half channel 1: 6 shuffles
// for 2 channels instead of 10, 9 for 4 instead of 20); (c) the CC accumulators of a target are adjacent, so the
// RMW is one LDS.64/128 + one STS.64/128; (d) the slices tile the index array linearly, so every load is a
// running pointer + immediate.  All targets of a row except the padding are distinct => one __syncwarp per row.
// sums z[0..CC) over the warp; lane 0 receives every total (other lanes: unspecified).  Fixed order.
template <int CC>
__device__ __forceinline__ void gf_reduce(float (&z)[CC], int lane) {
  constexpr unsigned F = 0xffffffffu;
  if constexpr (CC == 1) {
    z[0] = warp_sum(z[0]);
  } else if constexpr (CC == 2) {
    const bool hi = lane & 16;
    float keep = hi ? z[1] : z[0];
    keep += __shfl_xor_sync(F, hi ? z[0] : z[1], 16);
#pragma unroll
    for (int o = 8; o; o >>= 1) keep += __shfl_xor_sync(F, keep, o);
    z[0] = keep;                                  // valid on lanes 0..15
    z[1] = __shfl_xor_sync(F, keep, 16);          // lanes 0..15 read channel 1 from the upper half
  } else {
    const bool h16 = lane & 16, h8 = lane & 8;
    float a = h16 ? z[1] : z[0], b = h16 ? z[3] : z[2];
    a += __shfl_xor_sync(F, h16 ? z[0] : z[1], 16);   // a: ch0 (low half) / ch1 (high half)
    b += __shfl_xor_sync(F, h16 ? z[2] : z[3], 16);   // b: ch2 / ch3
    float k = h8 ? b : a;
    k += __shfl_xor_sync(F, h8 ? a : b, 8);           // lanes 0-7 ch0, 8-15 ch2, 16-23 ch1, 24-31 ch3
#pragma unroll
    for (int o = 4; o; o >>= 1) k += __shfl_xor_sync(F, k, o);
    z[0] = k;
    z[2] = __shfl_sync(F, k, 8);
    z[1] = __shfl_sync(F, k, 16);
    z[3] = __shfl_sync(F, k, 24);
  }
}

template <int CC> __device__ __forceinline__ void lds_vec(float (&a)[CC], uint32_t addr) {
  if constexpr (CC == 1) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a[0]) : "r"(addr) : "memory");
  else if constexpr (CC == 2)
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(a[0]), "=f"(a[1]) : "r"(addr) : "memory");
  else
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]) : "r"(addr) : "memory");
}
template <int CC> __device__ __forceinline__ void sts_vec(uint32_t addr, const float (&a)[CC]) {
  if constexpr (CC == 1) asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(a[0]) : "memory");
  else if constexpr (CC == 2)
    asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(addr), "f"(a[0]), "f"(a[1]) : "memory");
  else
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a[0]), "f"(a[1]), "f"(a[2]), "f"(a[3])
                 : "memory");
}

template <int CC, int SPR>
__global__ void __launch_bounds__(GF_THREADS)
group_grad_rows_kernel(const float* __restrict__ grad_out, const int32_t* __restrict__ idx, int c, int n, int m,
                       float* __restrict__ grad_points, const int* __restrict__ violations) {
  extern __shared__ __align__(16) float gf_acc[];  // [GF_WARPS][n][CC]
  if (*violations != 0) return;
  constexpr int NS = 32 * SPR;
  constexpr int R = SPR >= 4 ? 1 : 4 / SPR;       // rows in flight per iteration (4 slices of loads)
  const int cloud = blockIdx.y, c0 = blockIdx.x * CC, cc = min(CC, c - c0);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int i = tid; i < GF_WARPS * CC * n; i += GF_THREADS) gf_acc[i] = 0.f;
  __syncthreads();
  // one 32-bit shared address per warp; the RMWs below are explicit ld/st.shared so that the compiler does not
  // rebuild a generic pointer per access
  const uint32_t acc_s = (uint32_t)__cvta_generic_to_shared(gf_acc) + (uint32_t)w * (uint32_t)n * (4u * CC);
  const size_t E = (size_t)m * NS;
  const int rows_per_warp = (m + GF_WARPS - 1) / GF_WARPS;
  const int j_beg = min(m, w * rows_per_warp), j_end = min(m, j_beg + rows_per_warp);
  const int32_t* ixp = idx + (size_t)cloud * E + (size_t)j_beg * NS + lane;
  const float* gop[CC];
#pragma unroll
  for (int l = 0; l < CC; ++l)
    gop[l] = grad_out + ((size_t)cloud * c + c0 + min(l, cc - 1)) * E + (size_t)j_beg * NS + lane;
#pragma unroll
  for (int l = 0; l < CC; ++l) asm volatile("" : "+l"(gop[l]));  // opaque: keeps the bases out of the loop body
  for (int j = j_beg; j < j_end; j += R) {
    int p[R][SPR];
    float v[R][SPR][CC];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const bool live = j + r < j_end;
#pragma unroll
      for (int sl = 0; sl < SPR; ++sl) {
        p[r][sl] = live ? ixp[(r * SPR + sl) * 32] : -1;
#pragma unroll
        for (int l = 0; l < CC; ++l) v[r][sl][l] = live ? __ldcs(gop[l] + (r * SPR + sl) * 32) : 0.f;
      }
    }
    ixp += R * NS;
#pragma unroll
    for (int l = 0; l < CC; ++l) gop[l] += R * NS;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int p0 = __shfl_sync(0xffffffffu, p[r][0], 0);
      bool is0[SPR];
      float z[CC];
#pragma unroll
      for (int sl = 0; sl < SPR; ++sl) {
        is0[sl] = p[r][sl] == p0;
#pragma unroll
        for (int l = 0; l < CC; ++l) {
          const float t = is0[sl] ? v[r][sl][l] : 0.f;
          z[l] = sl ? z[l] + t : t;
        }
      }
      gf_reduce<CC>(z, lane);
#pragma unroll
      for (int sl = 0; sl < SPR; ++sl) {
        const bool lead = sl == 0 && lane == 0;      // holds the row's first index itself
        if (p[r][sl] >= 0 && (!is0[sl] || lead)) {
          float a[CC];
          const uint32_t addr = acc_s + (uint32_t)p[r][sl] * (4u * CC);
          lds_vec<CC>(a, addr);
#pragma unroll
          for (int l = 0; l < CC; ++l) a[l] += lead ? z[l] : v[r][sl][l];
          sts_vec<CC>(addr, a);
        }
      }
      __syncwarp();  // the next row may touch the same targets from other lanes
    }
  }
  __syncthreads();
  float* gp = grad_points + ((size_t)cloud * c + c0) * n;
  for (int i = tid; i < cc * n; i += GF_THREADS) {
    const int l = i / n, t = i - l * n;
    float sacc = 0.f;
#pragma unroll
    for (int ww = 0; ww < GF_WARPS; ++ww) sacc += gf_acc[((size_t)ww * n + t) * CC + l];
    gp[i] = sacc;
  }
}

// ============================================================================ gather_points (+grad)
__global__ void gather_points_kernel(const float* __restrict__ points, const int32_t* __restrict__ idx, int c, int n,
                                     int m, float* __restrict__ out) {
  const int cloud = blockIdx.z, ch = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m) out[((size_t)cloud * c + ch) * m + j] = points[((size_t)cloud * c + ch) * n + idx[(size_t)cloud * m + j]];
}

// ============================================================================ three_nn / three_interpolate
constexpr int TN_THREADS = 128;
constexpr int TN_CHUNK = 2048;

__global__ void __launch_bounds__(TN_THREADS)
three_nn_kernel(const float* __restrict__ unknown, const float* __restrict__ known, int n, int m,
                float* __restrict__ dist2o, int32_t* __restrict__ idxo) {
  __shared__ float sx[TN_CHUNK], sy[TN_CHUNK], sz[TN_CHUNK];
  const int cloud = blockIdx.y, tid = threadIdx.x;
  const int j = blockIdx.x * TN_THREADS + tid;
  const float* u = unknown + ((size_t)cloud * n + min(j, n - 1)) * 3;
  const float ux = u[0], uy = u[1], uz = u[2];
  const float* kn = known + (size_t)cloud * m * 3;
  // reference keeps best1..3 as doubles initialised to 1e40 (interpolate_gpu.cu:27): a float candidate
  // always compares below that, so +inf sentinels with a strict '<' reproduce it, except that a row with
  // fewer than 3 known points keeps index 0 and the (float)1e40 = +inf distance — same here.
  float b1 = __int_as_float(0x7f800000), b2 = b1, b3 = b1;
  int i1 = 0, i2 = 0, i3 = 0;
  for (int c0 = 0; c0 < m; c0 += TN_CHUNK) {
    const int cn = min(TN_CHUNK, m - c0);
    __syncthreads();
    for (int t = tid; t < cn; t += TN_THREADS) {
      sx[t] = kn[(c0 + t) * 3]; sy[t] = kn[(c0 + t) * 3 + 1]; sz[t] = kn[(c0 + t) * 3 + 2];
    }
    __syncthreads();
    for (int t = 0; t < cn; ++t) {
      const float d = dist2_pn2(ux, uy, uz, sx[t], sy[t], sz[t]);
      const int k = c0 + t;
      if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
      else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
      else if (d < b3) { b3 = d; i3 = k; }
    }
  }
  if (j < n) {
    float* dn = dist2o + ((size_t)cloud * n + j) * 3;
    int32_t* io = idxo + ((size_t)cloud * n + j) * 3;
    dn[0] = b1; dn[1] = b2; dn[2] = b3;
    io[0] = i1; io[1] = i2; io[2] = i3;
  }
}

__global__ void three_interpolate_kernel(const float* __restrict__ points, const int32_t* __restrict__ idx,
                                         const float* __restrict__ weight, int c, int m, int n,
                                         float* __restrict__ out) {
  const int cloud = blockIdx.z, ch = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const size_t o = ((size_t)cloud * n + j) * 3;
  const float* p = points + ((size_t)cloud * c + ch) * m;
  // interpolate_gpu.cu:98-99 `p1*w1 + p2*w2 + p3*w3` is contracted by nvcc into
  // t = p2*w2; t = fma(p1,w1,t); t = fma(p3,w3,t)  (read off the reference SASS) — reproduced exactly.
  out[((size_t)cloud * c + ch) * n + j] =
      __fmaf_rn(p[idx[o + 2]], weight[o + 2], __fmaf_rn(p[idx[o]], weight[o], __fmul_rn(p[idx[o + 1]], weight[o + 1])));
}

static size_t csr_ws_bytes(int b, int n, int E) { return 16 + (size_t)b * ((size_t)n + 1 + (size_t)E) * sizeof(int); }

static int pick_W(int n, int threads) {
  int W = threads / 32;
  while (W > 1 && (size_t)W * (n + 1) * 4 > 96 * 1024) W >>= 1;
  return W;
}

static int run_csr_grad(const float* grad_out, const float* weight, const int32_t* idx, int b, int c, int n, int E,
                        int e_div, float* grad_points, void* workspace, size_t workspace_bytes, cudaStream_t s,
                        const int* fast_flag = nullptr) {
  if (workspace_bytes < csr_ws_bytes(b, n, E) || !workspace) return GEOA3_EWORKSPACE;
  int* csr = (int*)workspace + 4;  // the first 16 bytes hold the dispatch flag
  const int W = pick_W(n, CSRG_THREADS);
  const size_t sm1 = (size_t)W * ((n + 1) & ~1) * 4;
  if (sm1 > 200 * 1024) return GEOA3_EUNSUPPORTED;
  static PerDeviceOnce attr_done;
  if (attr_done.needed()) {
    cudaError_t e = cudaFuncSetAttribute(csr_global_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(csr_gather_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_done.done();
  }
  csr_global_kernel<<<b, CSRG_THREADS, sm1, s>>>(idx, E, n, W, csr, fast_flag);
  int err = GEOA3_LAUNCH_RESULT();
  if (err) return err;
  const size_t coop_smem = (size_t)E * 4 + (size_t)((n + 1 + 3) & ~3) * 4 + (size_t)E * 2;
  if (!weight && e_div == 1 && c >= GG2_CC && (E & 3) == 0 && E <= 65536 && coop_smem <= 110 * 1024) {
    static PerDeviceOnce coop_attr;
    if (coop_attr.needed()) {
      cudaError_t e = cudaFuncSetAttribute(csr_gather_grad_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           110 * 1024);
      if (e != cudaSuccess) return (int)e;
      coop_attr.done();
    }
    csr_gather_grad_coop_kernel<<<dim3(ceil_div(c, GG2_CC), b), GG2_THREADS, coop_smem, s>>>(
        grad_out, csr, c, n, E, grad_points, fast_flag);
    return GEOA3_LAUNCH_RESULT();
  }
  const int Eg = E / e_div;
  int tile = Eg;
  if ((size_t)Eg * 4 > 96 * 1024) {
    tile = 24 * 1024;  // 96 KB tiles, two CTAs per SM
    if (n > 8 * GG_THREADS) return GEOA3_EUNSUPPORTED;
  }
  csr_gather_grad_kernel<<<dim3(c, b), GG_THREADS, (size_t)tile * 4, s>>>(grad_out, weight, csr, c, n, E, e_div, tile,
                                                                         grad_points, fast_flag);
  return GEOA3_LAUNCH_RESULT();
}

template <int CC>
static int launch_fwd_order(const float* grad_out, const int32_t* idx, int b, int c, int n, int m, int ns,
                            float* grad_points, const int* flag, cudaStream_t s) {
  const size_t smem = (size_t)GF_WARPS * CC * n * 4;
  cudaError_t e = cudaFuncSetAttribute(group_grad_fwd_order_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       200 * 1024);
  if (e != cudaSuccess) return (int)e;
  group_grad_fwd_order_kernel<CC><<<dim3(ceil_div(c, CC), b), GF_THREADS, smem, s>>>(grad_out, idx, c, n, m, ns,
                                                                                    grad_points, flag);
  return GEOA3_LAUNCH_RESULT();
}

template <int CC, int SPR>
static int launch_rows_t(const float* grad_out, const int32_t* idx, int b, int c, int n, int m, float* grad_points,
                         const int* flag, cudaStream_t s) {
  const size_t smem = (size_t)GF_WARPS * CC * n * 4;
  cudaError_t e = cudaFuncSetAttribute(group_grad_rows_kernel<CC, SPR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       200 * 1024);
  if (e != cudaSuccess) return (int)e;
  group_grad_rows_kernel<CC, SPR><<<dim3(ceil_div(c, CC), b), GF_THREADS, smem, s>>>(grad_out, idx, c, n, m,
                                                                                    grad_points, flag);
  return GEOA3_LAUNCH_RESULT();
}

static int launch_rows(int CC, int SPR, const float* grad_out, const int32_t* idx, int b, int c, int n, int m,
                       float* grad_points, const int* flag, cudaStream_t s) {
#define GF_ROWS_CASE(cc_, spr_) \
  if (CC == cc_ && SPR == spr_) return launch_rows_t<cc_, spr_>(grad_out, idx, b, c, n, m, grad_points, flag, s)
  GF_ROWS_CASE(1, 1); GF_ROWS_CASE(1, 2); GF_ROWS_CASE(1, 4);
  GF_ROWS_CASE(2, 1); GF_ROWS_CASE(2, 2); GF_ROWS_CASE(2, 4);
  GF_ROWS_CASE(4, 1); GF_ROWS_CASE(4, 2); GF_ROWS_CASE(4, 4);
#undef GF_ROWS_CASE
  return GEOA3_EUNSUPPORTED;
}

// group_points_grad: forward-order kernel when every index row has the ball-query shape (checked on the device),
// generic CSR path otherwise; both are launched, each exits immediately when it is not its turn.
static int run_group_grad(const float* grad_out, const int32_t* idx, int b, int c, int n, int m, int ns,
                          float* grad_points, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  const int E = m * ns;
  if (workspace_bytes < csr_ws_bytes(b, n, E) || !workspace) return GEOA3_EWORKSPACE;
  int* flag = (int*)workspace;
  // channels per CTA.  Generic kernel: 2 measured best (4 KB of private accumulators per warp at n = 512 keeps ~50
  // warps resident per SM).  Row-structured kernel (nsample 32/64/128): 4 when >= 3 CTAs/SM still fit (fewer
  // shuffles and RMW instructions per channel), 1 for the xyz-only case where 4 would leave the grid too small.
  const bool rows_kernel = ns == 32 || ns == 64 || ns == 128;
  int CC = 0;
  if ((size_t)GF_WARPS * 2 * n * 4 <= 100 * 1024) CC = 2;
  else if ((size_t)GF_WARPS * n * 4 <= 200 * 1024) CC = 1;
  if (!CC) return run_csr_grad(grad_out, nullptr, idx, b, c, n, E, 1, grad_points, workspace, workspace_bytes, s);
  if (rows_kernel) {
    if (c < 8) CC = 1;
    else if ((size_t)GF_WARPS * 4 * n * 4 <= 72 * 1024) CC = 4;
  }
  if (c == 1) CC = 1;
  if (const char* ov = getenv("GEOA3_GF_CC")) {  // tuning knob (tools/time_kernels.py), not part of the API
    const int v = atoi(ov);
    if ((v == 1 || v == 2 || v == 4 || v == 8) && (size_t)GF_WARPS * v * n * 4 <= 200 * 1024) CC = v;
  }
  cudaError_t e = cudaMemsetAsync(flag, 0, 16, s);
  if (e != cudaSuccess) return (int)e;
  fwd_order_ok_kernel<<<ceil_div(b * m, 8), 256, 0, s>>>(idx, b * m, ns, n, flag);
  int err = GEOA3_LAUNCH_RESULT();
  if (err) return err;
  if (rows_kernel) {
    if (CC > 4) CC = 4;
    err = launch_rows(CC, ns / 32, grad_out, idx, b, c, n, m, grad_points, flag, s);
    if (err) return err;
    return run_csr_grad(grad_out, nullptr, idx, b, c, n, E, 1, grad_points, workspace, workspace_bytes, s, flag);
  }
  switch (CC) {
    case 8: err = launch_fwd_order<8>(grad_out, idx, b, c, n, m, ns, grad_points, flag, s); break;
    case 4: err = launch_fwd_order<4>(grad_out, idx, b, c, n, m, ns, grad_points, flag, s); break;
    case 2: err = launch_fwd_order<2>(grad_out, idx, b, c, n, m, ns, grad_points, flag, s); break;
    default: err = launch_fwd_order<1>(grad_out, idx, b, c, n, m, ns, grad_points, flag, s); break;
  }
  if (err) return err;
  return run_csr_grad(grad_out, nullptr, idx, b, c, n, E, 1, grad_points, workspace, workspace_bytes, s, flag);
}

}  // namespace geoa3

using namespace geoa3;

// start == nullptr: pointnet2 semantics (first pick 0, frozen points, reference tie order);
// start != nullptr: plain FPS from the given first picks (Lib/utility.py:175-187)
template <int MODE>
static int launch_fps(const float* xyz, int b, int n, int m, const int32_t* start, int32_t* idx, size_t smem, int bs,
                      int bits, cudaStream_t s) {
  const bool one_warp = !start && getenv("GEOA3_FPS_WARP") != nullptr;  // A/B knob for tools/time_kernels.py, not part of the API
  if (one_warp && n <= 512) {
    fps_warp_kernel<16><<<b, 32, smem, s>>>(xyz, n, m, bs, bits, idx);
  } else if (one_warp && n <= 1024) {
    fps_warp_kernel<32><<<b, 32, smem, s>>>(xyz, n, m, bs, bits, idx);
  } else if (n <= 256) {
    fps_quad_kernel<2, MODE><<<b, FPSQ_WARPS * 32, smem + (size_t)m * 4, s>>>(xyz, n, m, bs, bits, idx, start);
  } else if (n <= 512) {
    fps_quad_kernel<4, MODE><<<b, FPSQ_WARPS * 32, smem + (size_t)m * 4, s>>>(xyz, n, m, bs, bits, idx, start);
  } else if (n <= 1024) {
    fps_quad_kernel<8, MODE><<<b, FPSQ_WARPS * 32, smem + (size_t)m * 4, s>>>(xyz, n, m, bs, bits, idx, start);
  } else if (n <= 2048) {
    fps_quad_kernel<16, MODE><<<b, FPSQ_WARPS * 32, smem + (size_t)m * 4, s>>>(xyz, n, m, bs, bits, idx, start);
  } else if (n <= 4096) {
    const int T = min(1024, max(32, ((n + 3) / 4 + 31) & ~31));
    fps_kernel<4, MODE><<<b, T, smem, s>>>(xyz, n, m, bs, bits, idx, start);
  } else if (n <= 8192) {
    fps_kernel<8, MODE><<<b, 1024, smem, s>>>(xyz, n, m, bs, bits, idx, start);
  } else if (n <= 16384) {
    fps_kernel<16, MODE><<<b, 1024, smem, s>>>(xyz, n, m, bs, bits, idx, start);
  } else {
    return GEOA3_EUNSUPPORTED;
  }
  return GEOA3_LAUNCH_RESULT();
}

static int run_fps(const float* xyz, int b, int n, int m, const int32_t* start, int32_t* idx, geoa3_stream_t stream) {
  GEOA3_CHECK_ARG(xyz && idx && b > 0 && n > 0 && m > 0);
  if (n >= (1 << 20) || (size_t)n * 12 > 200 * 1024) return GEOA3_EUNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = (size_t)n * 12;
  const int bs = ref_fps_block(n);
  int bits = 0;
  while ((1 << bits) < bs) ++bits;
  if (start) bits = 0;  // ties -> lowest index
  static PerDeviceOnce attr_done;
  if (attr_done.needed()) {
    cudaError_t e = cudaSuccess;
#define GEOA3_FPS_ATTR(P_, M_) \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(fps_kernel<P_, M_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)
    GEOA3_FPS_ATTR(4, 0); GEOA3_FPS_ATTR(8, 0); GEOA3_FPS_ATTR(16, 0);
    GEOA3_FPS_ATTR(4, 1); GEOA3_FPS_ATTR(8, 1); GEOA3_FPS_ATTR(16, 1);
    GEOA3_FPS_ATTR(4, 2); GEOA3_FPS_ATTR(8, 2); GEOA3_FPS_ATTR(16, 2);
#undef GEOA3_FPS_ATTR
    if (e != cudaSuccess) return (int)e;
    attr_done.done();
  }
  // norm arithmetic of the plain variant: separately rounded squares — what torch's CUDA norm reduction computes
  // (tools/time_kernels.py --fps-plain: picks identical to the torch loop; GEOA3_FPS_NORM=1 selects the fma chain,
  // which diverges from it at near-ties)
  const char* nm = getenv("GEOA3_FPS_NORM");
  if (!start) return launch_fps<0>(xyz, b, n, m, start, idx, smem, bs, bits, s);
  if (nm && atoi(nm) == 1) return launch_fps<1>(xyz, b, n, m, start, idx, smem, bs, bits, s);
  return launch_fps<2>(xyz, b, n, m, start, idx, smem, bs, bits, s);
}

extern "C" int geoa3_furthest_point_sampling(const float* xyz, int b, int n, int m, int32_t* idx,
                                             geoa3_stream_t stream) {
  return run_fps(xyz, b, n, m, nullptr, idx, stream);
}

extern "C" int geoa3_farthest_points_sample(const float* xyz, int b, int n, int m, const int32_t* start, int32_t* idx,
                                            geoa3_stream_t stream) {
  GEOA3_CHECK_ARG(start);
  return run_fps(xyz, b, n, m, start, idx, stream);
}

extern "C" int geoa3_ball_query(const float* new_xyz, const float* xyz, int b, int n, int m, float radius,
                                int nsample, int32_t* idx, geoa3_stream_t stream) {
  GEOA3_CHECK_ARG(new_xyz && xyz && idx && b > 0 && n > 0 && m > 0 && nsample > 0);
  if (b > 65535) return GEOA3_EUNSUPPORTED;
  const size_t smem = (size_t)((n + 63) & ~63) * 12;
  if (smem > 200 * 1024) return GEOA3_EUNSUPPORTED;
  static PerDeviceOnce attr_done;
  if (attr_done.needed()) {
    cudaError_t e = cudaFuncSetAttribute(ball_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_done.done();
  }
  if (n <= 65535 && nsample <= BQ_MAX_NS && !getenv("GEOA3_BQ_WARP")) {  // (env: A/B knob for tools/time_kernels.py, not part of the API)
    const size_t hs = (size_t)BQT_THREADS * (nsample + 2) * sizeof(uint16_t);
    static PerDeviceOnce attr_t;
    if (attr_t.needed()) {
      cudaError_t e = cudaFuncSetAttribute(ball_query_thread_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           BQT_THREADS * (BQ_MAX_NS + 2) * (int)sizeof(uint16_t));
      if (e != cudaSuccess) return (int)e;
      attr_t.done();
    }
    ball_query_thread_kernel<<<dim3(ceil_div(m, BQT_THREADS), b), BQT_THREADS, hs, (cudaStream_t)stream>>>(
        new_xyz, xyz, n, m, radius, nsample, idx);
    return GEOA3_LAUNCH_RESULT();
  }
  dim3 grid(ceil_div(m, BQ_WARPS * BQ_PER_WARP), b);
  ball_query_kernel<<<grid, BQ_WARPS * 32, smem, (cudaStream_t)stream>>>(new_xyz, xyz, n, m, radius, nsample, idx);
  return GEOA3_LAUNCH_RESULT();
}

extern "C" int geoa3_group_points(const float* points, const int32_t* idx, int b, int c, int n, int npoints,
                                  int nsample, float* out, geoa3_stream_t stream) {
  GEOA3_CHECK_ARG(points && idx && out && b > 0 && c > 0 && n > 0 && npoints > 0 && nsample > 0);
  if (b > 65535 || (size_t)n * 4 * 1 > 200 * 1024) return GEOA3_EUNSUPPORTED;
  const int E = npoints * nsample;
  // channels per CTA limited by shared memory (GP_CC rows of n floats; fewer rows for very long ones)
  int gcc = c < GP_CC ? c : GP_CC;
  while (gcc > 1 && (size_t)gcc * n * 4 > 200 * 1024) gcc >>= 1;
  const size_t smem = (size_t)gcc * n * 4;
  if (smem > 200 * 1024) return GEOA3_EUNSUPPORTED;
  static PerDeviceOnce attr_done;
  if (attr_done.needed()) {
    cudaError_t e = cudaFuncSetAttribute(group_points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_done.done();
  }
  dim3 grid(ceil_div(E, GP_ETILE), ceil_div(c, gcc), b);
  group_points_kernel<<<grid, GP_THREADS, smem, (cudaStream_t)stream>>>(points, idx, c, n, E, out, gcc);
  return GEOA3_LAUNCH_RESULT();
}

extern "C" size_t geoa3_group_points_grad_workspace_bytes(int b, int n, int npoints, int nsample) {
  return csr_ws_bytes(b, n, npoints * nsample);
}

extern "C" int geoa3_group_points_grad(const float* grad_out, const int32_t* idx, int b, int c, int n, int npoints,
                                       int nsample, float* grad_points, void* workspace, size_t workspace_bytes,
                                       geoa3_stream_t stream) {
  GEOA3_CHECK_ARG(grad_out && idx && grad_points && b > 0 && c > 0 && n > 0 && npoints > 0 && nsample > 0);
  if (b > 65535 || c > 65535) return GEOA3_EUNSUPPORTED;
  return run_group_grad(grad_out, idx, b, c, n, npoints, nsample, grad_points, workspace, workspace_bytes,
                        (cudaStream_t)stream);
}

extern "C" int geoa3_gather_points(const float* points, const int32_t* idx, int b, int c, int n, int m, float* out,
                                   geoa3_stream_t stream) {
  GEOA3_CHECK_ARG(points && idx && out && b > 0 && c > 0 && n > 0 && m > 0);
  if (b > 65535 || c > 65535) return GEOA3_EUNSUPPORTED;
  gather_points_kernel<<<dim3(ceil_div(m, 256), c, b), 256, 0, (cudaStream_t)stream>>>(points, idx, c, n, m, out);
  return GEOA3_LAUNCH_RESULT();
}

extern "C" int geoa3_gather_points_grad(const float* grad_out, const int32_t* idx, int b, int c, int n, int m,
                                        float* grad_points, void* workspace, size_t workspace_bytes,
                                        geoa3_stream_t stream) {
  GEOA3_CHECK_ARG(grad_out && idx && grad_points && b > 0 && c > 0 && n > 0 && m > 0);
  if (b > 65535 || c > 65535) return GEOA3_EUNSUPPORTED;
  return run_csr_grad(grad_out, nullptr, idx, b, c, n, m, 1, grad_points, workspace, workspace_bytes,
                      (cudaStream_t)stream);
}

extern "C" int geoa3_three_nn(const float* unknown, const float* known, int b, int n, int m, float* dist2, int32_t* idx,
                              geoa3_stream_t stream) {
  GEOA3_CHECK_ARG(unknown && known && dist2 && idx && b > 0 && n > 0 && m > 0);
  if (b > 65535) return GEOA3_EUNSUPPORTED;
  three_nn_kernel<<<dim3(ceil_div(n, TN_THREADS), b), TN_THREADS, 0, (cudaStream_t)stream>>>(unknown, known, n, m, dist2,
                                                                                            idx);
  return GEOA3_LAUNCH_RESULT();
}

extern "C" int geoa3_three_interpolate(const float* points, const int32_t* idx, const float* weight, int b, int c,
                                       int m, int n, float* out, geoa3_stream_t stream) {
  GEOA3_CHECK_ARG(points && idx && weight && out && b > 0 && c > 0 && m > 0 && n > 0);
  if (b > 65535 || c > 65535) return GEOA3_EUNSUPPORTED;
  three_interpolate_kernel<<<dim3(ceil_div(n, 256), c, b), 256, 0, (cudaStream_t)stream>>>(points, idx, weight, c, m, n,
                                                                                          out);
  return GEOA3_LAUNCH_RESULT();
}

extern "C" int geoa3_three_interpolate_grad(const float* grad_out, const int32_t* idx, const float* weight, int b,
                                            int c, int n, int m, float* grad_points, void* workspace,
                                            size_t workspace_bytes, geoa3_stream_t stream) {
  GEOA3_CHECK_ARG(grad_out && idx && weight && grad_points && b > 0 && c > 0 && n > 0 && m > 0);
  if (b > 65535 || c > 65535) return GEOA3_EUNSUPPORTED;
  // targets are the m known points, sources the 3n (point, slot) pairs
  return run_csr_grad(grad_out, weight, idx, b, c, m, 3 * n, 3, grad_points, workspace, workspace_bytes,
                      (cudaStream_t)stream);
}

extern "C" int geoa3_version(void) { return 1000; }

extern "C" const char* geoa3_error_string(int code) {
  switch (code) {
    case GEOA3_OK: return "ok";
    case GEOA3_EINVAL: return "geoa3: invalid argument (null pointer or non-positive size)";
    case GEOA3_EUNSUPPORTED: return "geoa3: size outside the supported range of the sm_100a kernels";
    case GEOA3_EWORKSPACE: return "geoa3: workspace missing or too small";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "geoa3: unknown error";
  }
}
