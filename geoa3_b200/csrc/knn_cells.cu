// knn_cells.cu — the neighbour SETS of the curvature loss through a per-cloud CELL GRID (self-query, K = k+1 <= 33,
// clouds of up to 65 535 points).  Same members as geoa3_knn / geoa3_knn_set — the K lexicographically smallest
// (pinned distance, ORIGINAL index) pairs minus the `drop` smallest — found by looking only at the cells a query's
// search ball touches instead of streaming the whole cloud past every query.
//
//  geoa3_cell_sort   one CTA per cloud: bounding box, G^3 uniform cells, stable counting sort (cell-major, ascending
//                    original index inside a cell: csr.cuh) -> the cloud as float4 (x, y, z, original index) in cell
//                    order, the cell start table and the inverse permutation.  Cell index = (cz*G + cy)*G + cx, so one
//                    (cz, cy) ROW of cells is one contiguous range of positions.
//  geoa3_knn_cells   one query per thread, a warp's 32 queries are neighbours in cell order.  tau = largest pinned
//                    distance to the hinted candidates (previous step's neighbours): an upper bound of the K-th distance
//                    whenever the hints are K-1 distinct points other than the query (verified at the end: fewer than K
//                    survivors => the warp searches again from tau = +inf).  For every cell row within sqrt(tau) of the
//                    query (conservative test on the row's y/z slab) the x interval that the ball can reach inside
//                    that row is mapped to a position range; every candidate in it is evaluated with the PINNED fma
//                    chain and kept when d <= tau (list in shared memory, arrival = ascending position).  When the
//                    list overflows, and once at the end, the largest (distance, original index) keys are removed
//                    until K remain.  Members are written in list order = ascending position in the cell arrangement,
//                    a function of the cloud alone (never of the hint).
//
// Exactness.  cell(x) = trunc(clamp((x - lo) * inv_h, 0, G-1)) is monotone in x and is the SAME function on both
// sides (sorting and querying), so |c.x - q.x| <= rx implies cell(q.x - rx) <= cell(c.x) <= cell(q.x + rx) whatever
// the rounding; radii are inflated (relative 1e-5 + `slack` = 8e-6 * max|coordinate|, ~64 ulp of the largest
// coordinate) over the rounding of the pinned chain (<= 5 ulp relative), of the cell boundaries and of the sums.
// The grid only decides which candidates are LOOKED AT; what is kept is decided by the pinned arithmetic.
#include <climits>

#include "cells.cuh"
#include "csr.cuh"

#ifndef KC_EXTRA
#define KC_EXTRA 11
#endif
#ifndef KC_THREADS
#define KC_THREADS 128
#endif

namespace geoa3 {

constexpr int KC_SORT_THREADS = 1024;

__host__ __device__ inline size_t kc_sort_smem(int n, int nc) {
  return (size_t)n * 4 + (size_t)(nc + 1) * 4 + (size_t)nc * 4 + (size_t)((n + 1) & ~1) * 2 + 32 * 4 + 32 * 8 * 4 + KC_GP * 4;
}

__global__ void __launch_bounds__(KC_SORT_THREADS)
cell_sort_kernel(const float* __restrict__ pc, int n, int G, unsigned char* __restrict__ blobs) {
  extern __shared__ __align__(16) unsigned char kc_smem[];
  const int nc = G * G * G;
  int* keys = reinterpret_cast<int*>(kc_smem);      // [n]
  int* offs = keys + n;                             // [nc + 1]
  int* cnt = offs + nc + 1;                         // [nc]
  int* scan = cnt + nc;                             // [32]
  float* red = reinterpret_cast<float*>(scan + 32); // [32][8]
  float* gp = red + 32 * 8;                         // [KC_GP]
  uint16_t* ent = reinterpret_cast<uint16_t*>(gp + KC_GP);  // [n]
  const int cloud = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float* p = pc + (size_t)cloud * 3 * n;
  unsigned char* blob = blobs + (size_t)cloud * kc_blob_bytes(n, nc);

  float lx = 3e38f, ly = 3e38f, lz = 3e38f, hx = -3e38f, hy = -3e38f, hz = -3e38f;
  for (int i = tid; i < n; i += KC_SORT_THREADS) {
    const float x = p[i], y = p[n + i], z = p[2 * n + i];
    lx = fminf(lx, x); hx = fmaxf(hx, x);
    ly = fminf(ly, y); hy = fmaxf(hy, y);
    lz = fminf(lz, z); hz = fmaxf(hz, z);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, s)); hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, s));
    ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, s)); hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, s));
    lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, s)); hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, s));
  }
  if (lane == 0) {
    float* o = red + w * 8;
    o[0] = lx; o[1] = ly; o[2] = lz; o[3] = hx; o[4] = hy; o[5] = hz;
  }
  __syncthreads();
  if (w == 0) {
    const float* o = red + lane * 8;
    lx = o[0]; ly = o[1]; lz = o[2]; hx = o[3]; hy = o[4]; hz = o[5];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, s)); hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, s));
      ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, s)); hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, s));
      lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, s)); hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, s));
    }
    if (lane == 0) {
      const float lo[3] = {lx, ly, lz}, hi[3] = {hx, hy, hz};
      float ma = 0.f;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float ext = hi[a] - lo[a];
        const bool ok = ext > 0.f && ext < 3e38f;  // empty / single-valued / non-finite axis: one cell layer
        gp[a] = ok ? lo[a] : (lo[a] < 3e38f ? lo[a] : 0.f);
        gp[3 + a] = ok ? (float)G / ext : 0.f;
        gp[6 + a] = ok ? ext / (float)G : 0.f;
        ma = fmaxf(ma, fmaxf(fabsf(lo[a]), fabsf(hi[a])));
      }
      gp[9] = ma < 3e38f ? 8e-6f * ma + 1e-30f : 0.f;
      gp[10] = (float)(G - 1);
      gp[11] = 0.f;
    }
  }
  __syncthreads();
  if (tid < KC_HDR / 4) reinterpret_cast<float*>(blob)[tid] = tid < KC_GP ? gp[tid] : 0.f;
  const float gm1 = gp[10];
  for (int i = tid; i < n; i += KC_SORT_THREADS) {
    const int cx = kc_cell(p[i], gp[0], gp[3], gm1), cy = kc_cell(p[n + i], gp[1], gp[4], gm1),
              cz = kc_cell(p[2 * n + i], gp[2], gp[5], gm1);
    keys[i] = (cz * G + cy) * G + cx;
  }
  __syncthreads();
  build_csr_sorted<KC_SORT_THREADS, uint16_t>(keys, n, nc, 0u, offs, cnt, ent, scan);
  float4* o4 = reinterpret_cast<float4*>(blob + KC_HDR);
  uint16_t* ip = reinterpret_cast<uint16_t*>(blob + kc_ip_off(n, nc));
  for (int t = tid; t < n; t += KC_SORT_THREADS) {
    const int o = ent[t];
    o4[t] = make_float4(p[o], p[n + o], p[2 * n + o], __int_as_float(o));
    ip[o] = (uint16_t)t;
  }
  uint16_t* cs = reinterpret_cast<uint16_t*>(blob + kc_cs_off(n));
  for (int c = tid; c < ((nc + 1 + 7) & ~7); c += KC_SORT_THREADS) cs[c] = (uint16_t)offs[min(c, nc)];
  for (int t = n + tid; t < ((n + 7) & ~7); t += KC_SORT_THREADS) ip[t] = 0;
}

template <int K>
struct KcCfg {
  static constexpr int R = K + KC_EXTRA;  // list slots per query
  static constexpr int T = KC_THREADS;    // one query per thread
};

// list entry = (pinned distance bits, ORIGINAL index); dead entries carry a negative distance
__host__ __device__ inline size_t kc_list_bytes(int r, int t) { return (size_t)r * t * 8; }

// Marks the `rm` largest (distance, original index) keys of a query's list dead (d = -2).  Per thread, loop form:
// it runs under divergence.  l_: the thread's column (stride COLS).
template <int COLS>
__device__ __forceinline__ void kc_mark(uint2* __restrict__ l_, int cnt, int rm) {
  for (; rm > 0; --rm) {
    float md = -1.f;
    int bs = 0;
    unsigned bi = 0u;
    for (int s = 0; s < cnt; ++s) {
      const uint2 e = l_[s * COLS];
      const float d = __uint_as_float(e.x);
      if (d >= md) {
        if (d > md || e.y > bi) { md = d; bs = s; bi = e.y; }
      }
    }
    l_[bs * COLS].x = __float_as_uint(-2.f);
  }
}

// List overflow (careful path only): keeps the kk smallest keys in arrival order, returns the new length and lowers
// *tau to the largest kept distance.
template <int COLS>
__device__ __forceinline__ int kc_cut(uint2* __restrict__ l_, int cnt, int kk, float* tau) {
  if (cnt <= kk) return cnt;
  kc_mark<COLS>(l_, cnt, cnt - kk);
  int w = 0;
  float mx = 0.f;
  for (int s = 0; s < cnt; ++s) {  // close the gaps
    const uint2 e = l_[s * COLS];
    const float d = __uint_as_float(e.x);
    if (d >= 0.f) {
      l_[w * COLS] = e;
      mx = fmaxf(mx, d);
      ++w;
    }
  }
  *tau = fminf(*tau, mx);
  return w;
}

// Writes the live entries of a list (arrival order = ascending position in the cell arrangement).
template <int COLS>
__device__ __forceinline__ void kc_write(const uint2* __restrict__ l_, int cnt, int kout, int32_t* __restrict__ io,
                                         float* __restrict__ dn) {
  int o = 0;
  for (int s = 0; s < cnt && o < kout; ++s) {
    const uint2 e = l_[s * COLS];
    if (__uint_as_float(e.x) >= 0.f) {
      io[o] = (int32_t)e.y;
      if (dn) dn[o] = __uint_as_float(e.x);
      ++o;
    }
  }
}

template <int K, int T, bool STAGED>
__global__ void __launch_bounds__(T)
knn_cells_kernel(const unsigned char* __restrict__ blobs, int n, int G, int kout, int drop, const int32_t* hint, int hint_k,
                 int32_t* idx_out, float* __restrict__ dist_out, int kk) {
  // kk = min(requested K, n): the list target (the template K is the capacity class)
  constexpr int R = KcCfg<K>::R;
  constexpr int HV = (K - 1) % 4 == 0 ? (K - 1) / 4 : 0;  // hint row as int4 registers when hint_k == K-1
  extern __shared__ __align__(16) unsigned char kc_smem[];
  const int nc = G * G * G;
  const int cloud = blockIdx.y, tid = threadIdx.x;
  const size_t blob_bytes = kc_blob_bytes(n, nc);
  const unsigned char* gblob = blobs + (size_t)cloud * blob_bytes;
  uint2* l_ = reinterpret_cast<uint2*>(kc_smem) + tid;  // this query's list: column tid of [R][T]
  const unsigned char* blob = gblob;
  __shared__ __align__(8) unsigned long long kc_bar;
  if (STAGED) {  // the whole blob is one contiguous, 16-byte sized block: a single TMA bulk copy stages it
    unsigned char* sblob = kc_smem + kc_list_bytes(R, T);
    kc_stage_issue(&kc_bar, sblob, gblob, (unsigned)blob_bytes);
    blob = sblob;
  }
  const float4* g4 = reinterpret_cast<const float4*>(gblob + KC_HDR);
  const int slot = blockIdx.x * T + tid;  // position in cell order
  const bool live = slot < n;
  const float4 q = g4[min(slot, n - 1)];
  const int qo = __float_as_int(q.w);     // ORIGINAL index of the query
  const bool hintv = HV > 0 && hint != nullptr && hint_k == K - 1 && ((reinterpret_cast<uintptr_t>(hint) & 15) == 0);
  int4 hreg[HV > 0 ? HV : 1];
  if (hintv) {  // fetched up front: the latency hides behind the staging
    const int4* h4 = reinterpret_cast<const int4*>(hint + ((size_t)cloud * n + qo) * (K - 1));
#pragma unroll
    for (int t = 0; t < HV; ++t) hreg[t] = h4[t];
  }
  if (STAGED) kc_stage_wait(&kc_bar);
  const float* sgp = reinterpret_cast<const float*>(blob);
  const float4* s4 = reinterpret_cast<const float4*>(blob + KC_HDR);
  const uint16_t* scs = reinterpret_cast<const uint16_t*>(blob + kc_cs_off(n));
  const uint16_t* sip = reinterpret_cast<const uint16_t*>(blob + kc_ip_off(n, nc));
  const float lox = sgp[0], loy = sgp[1], loz = sgp[2], ihx = sgp[3], ihy = sgp[4], ihz = sgp[5];
  const float slack = sgp[9], gm1 = sgp[10];

  float tau = KC_INF;
  if (hint != nullptr && hint_k + 1 >= kk && live) {
    auto cand = [&](int j) {
      const int pos = sip[min((unsigned)j, (unsigned)(n - 1))];
      const float4 c = s4[pos];
      return dist2(c.x, c.y, c.z, q.x, q.y, q.z);
    };
    float mx = 0.f;
    if (hintv) {
#pragma unroll
      for (int t = 0; t < HV; ++t) {
        mx = fmaxf(fmaxf(mx, cand(hreg[t].x)), cand(hreg[t].y));
        mx = fmaxf(fmaxf(mx, cand(hreg[t].z)), cand(hreg[t].w));
      }
    } else {
      const int32_t* h = hint + ((size_t)cloud * n + qo) * hint_k;  // may alias idx_out: own row, read before write
#pragma unroll 1
      for (int t = 0; t < hint_k; ++t) mx = fmaxf(mx, cand(h[t]));
    }
    if (mx < 3.0e38f) tau = mx;
  }
  int32_t* io = idx_out + ((size_t)cloud * n + qo) * kout;
  float* dn = dist_out ? dist_out + ((size_t)cloud * n + qo) * kout : nullptr;

  // ---- fast path (hinted queries, staged blob): every lane walks ITS OWN rows (row offsets relative to the lane's
  // first row are warp-uniform), appends are branch-free, the query itself is skipped (drop == 1: it is the smallest
  // key unless a second point sits at distance 0).  A list overflow, a zero-distance duplicate or a bound that turns
  // out invalid (fewer than K survivors) sends the query to the careful path below.
  bool careful = live;
  const bool fa = STAGED && live && tau < KC_INF && drop <= 1;
  if (STAGED && __any_sync(0xffffffffu, fa)) {
    const int qskip = drop == 1 ? qo : -1;
    int z0 = 0, y0 = 0, nz = -1, ny = -1;
    if (fa) {
      const float r = kc_sqrt(tau * KC_REL) * KC_REL + slack;
      z0 = kc_cell(q.z - r, loz, ihz, gm1);
      nz = kc_cell(q.z + r, loz, ihz, gm1) - z0;
      y0 = kc_cell(q.y - r, loy, ihy, gm1);
      ny = kc_cell(q.y + r, loy, ihy, gm1) - y0;
    }
    const int wz = __reduce_max_sync(0xffffffffu, nz), wy = __reduce_max_sync(0xffffffffu, ny);
    const float r2 = fmaf(tau, KC_REL, 1e-37f);
    const float hy = sgp[7], hz = sgp[8];
    const unsigned a4 = (unsigned)__cvta_generic_to_shared(s4);
    const unsigned acs = (unsigned)__cvta_generic_to_shared(scs);
    const unsigned l0 = (unsigned)__cvta_generic_to_shared(l_);
    const unsigned lend = l0 + R * T * 8;
    unsigned lp = l0;  // next free list slot (keeps advancing past the end: that is how an overflow is seen)
    for (int oz = 0; oz <= wz; ++oz) {
      const int rz = z0 + min(oz, max(nz, 0));
      const float zl = fmaf((float)rz, hz, loz);
      const float ez = fmaxf(fmaxf(zl - q.z, q.z - (zl + hz)) - slack, 0.f);  // lower bound of |c.z - q.z| in this slab
      const float remz = oz <= nz ? r2 - ez * ez : -1.f;
      for (int oy = 0; oy <= wy; ++oy) {
        const int ry = y0 + min(oy, max(ny, 0));
        const float yl = fmaf((float)ry, hy, loy);
        const float ey = fmaxf(fmaxf(yl - q.y, q.y - (yl + hy)) - slack, 0.f);
        const float rem = oy <= ny ? remz - ey * ey : -1.f;  // what is left for (c.x - q.x)^2
        const float rx = kc_sqrt(fmaxf(rem, 0.f)) * KC_REL + slack;
        const unsigned ab = acs + (unsigned)((rz * G + ry) * G) * 2u;
        const int x0 = kc_cell(q.x - rx, lox, ihx, gm1), x1 = kc_cell(q.x + rx, lox, ihx, gm1);
        const int s = (int)kc_lds16(ab + x0 * 2);
        const int len = rem >= 0.f ? (int)kc_lds16(ab + x1 * 2 + 2) - s : 0;
        const int wl = __reduce_max_sync(0xffffffffu, len);
        const unsigned ca = a4 + (unsigned)s * 16u;
        const unsigned cl = ca + (unsigned)max(len - 1, 0) * 16u;  // reads past the lane's own range are clamped to it
        auto eval = [&](int i, const float4 c) {
          const float d = dist2(c.x, c.y, c.z, q.x, q.y, q.z);  // the pinned arithmetic decides
          const int ci = __float_as_int(c.w);
          const bool pass = i < len && d <= tau && ci != qskip;
          if (pass && lp < lend) kc_sts64(lp, d, ci);
          if (pass) lp += T * 8;
        };
        int i = 0;
        for (; i + 1 < wl; i += 2) {
          const float4 c0 = kc_lds128(min(ca + (unsigned)i * 16u, cl));
          const float4 c1 = kc_lds128(min(ca + (unsigned)i * 16u + 16u, cl));
          eval(i, c0);
          eval(i + 1, c1);
        }
        if (i < wl) eval(i, kc_lds128(min(ca + (unsigned)i * 16u, cl)));
      }
    }
    if (fa) {
      const int cnt = (int)(lp - l0) / (T * 8);
      const int kt = kk - (drop == 1 ? 1 : 0);
      bool good = cnt >= kt && cnt <= R;
      if (good && drop == 1) {  // a zero-distance neighbour competes with the query for being the dropped key
        float dm = KC_INF;
        for (int s2 = 0; s2 < cnt; ++s2) dm = fminf(dm, __uint_as_float(l_[s2 * T].x));
        good = dm > 0.f;
      }
      if (good) {
        kc_mark<T>(l_, cnt, cnt - kt);
        kc_write<T>(l_, cnt, kout, io, dn);
        careful = false;
      }
    }
  }

  // ---- careful path: no usable hint, or the fast path gave up.  Generic: the query itself is a candidate, the list
  // is cut when it overflows (tau tightens), rows are the union of the warp's reach, the `drop` smallest keys are
  // removed at the end; an invalid bound makes the warp search again from +inf.
  bool act = careful;
  if (!__any_sync(0xffffffffu, act)) return;
  int cnt = 0;
  bool rescan;
  do {
    int z0, z1, y0, y1;
    {
      const float r = sqrtf(tau * KC_REL) * KC_REL + slack;
      z0 = act ? kc_cell(q.z - r, loz, ihz, gm1) : INT_MAX;
      z1 = act ? kc_cell(q.z + r, loz, ihz, gm1) : -1;
      y0 = act ? kc_cell(q.y - r, loy, ihy, gm1) : INT_MAX;
      y1 = act ? kc_cell(q.y + r, loy, ihy, gm1) : -1;
      z0 = __reduce_min_sync(0xffffffffu, z0); z1 = __reduce_max_sync(0xffffffffu, z1);
      y0 = __reduce_min_sync(0xffffffffu, y0); y1 = __reduce_max_sync(0xffffffffu, y1);
    }
    for (int rz = z0; rz <= z1; ++rz) {
      const float hz = sgp[8];
      const float zl = fmaf((float)rz, hz, loz);
      const float ez = fmaxf(fmaxf(zl - q.z, q.z - (zl + hz)) - slack, 0.f);
      const float ez2 = ez * ez;
      for (int ry = y0; ry <= y1; ++ry) {
        const float hy = sgp[7];
        const float yl = fmaf((float)ry, hy, loy);
        const float ey = fmaxf(fmaxf(yl - q.y, q.y - (yl + hy)) - slack, 0.f);
        const float rem = fmaf(tau, KC_REL, 1e-37f) - ez2 - ey * ey;
        const bool ok = act && rem >= 0.f;
        const float rx = sqrtf(fmaxf(rem, 0.f)) * KC_REL + slack;
        const int base = (rz * G + ry) * G;
        const int x0 = kc_cell(q.x - rx, lox, ihx, gm1), x1 = kc_cell(q.x + rx, lox, ihx, gm1);
        const int s = scs[base + x0];
        const int len = ok ? (int)scs[base + x1 + 1] - s : 0;
        const int wl = __reduce_max_sync(0xffffffffu, len);
        const float4* cp = s4 + s;
#pragma unroll 1
        for (int i = 0; i < wl; ++i) {
          if (i < len) {
            const float4 c = cp[i];
            const float d = dist2(c.x, c.y, c.z, q.x, q.y, q.z);
            if (d <= tau) {
              if (cnt == R) cnt = kc_cut<T>(l_, cnt, kk, &tau);  // makes room and tightens tau
              l_[cnt * T] = make_uint2(__float_as_uint(d), (unsigned)__float_as_int(c.w));
              ++cnt;
            }
          }
        }
      }
    }
    // fewer than K survivors => the hinted bound was not valid for this query: the warp searches again from +inf
    const bool fail = act && cnt < kk;
    rescan = __any_sync(0xffffffffu, fail);
    if (act && !fail) {
      kc_mark<T>(l_, cnt, cnt - kk);
      for (int t = 0; t < drop; ++t) {  // the `drop` smallest (distance, original index) keys are no members
        float md = KC_INF;
        int bs = 0;
        unsigned bo = 0xffffffffu;
        for (int s2 = 0; s2 < cnt; ++s2) {
          const uint2 e = l_[s2 * T];
          const float d = __uint_as_float(e.x);
          if (d >= 0.f && d <= md) {
            if (d < md || e.y < bo) { md = d; bs = s2; bo = e.y; }
          }
        }
        l_[bs * T].x = __float_as_uint(-1.f);
      }
      kc_write<T>(l_, cnt, kout, io, dn);
      act = false;
    }
    if (fail) { tau = KC_INF; cnt = 0; }
  } while (rescan);
}

template <int K>
static int launch_knn_cells_k(const unsigned char* blobs, int b, int n, int G, int kreq, int kout, int drop,
                              const int32_t* hint, int hint_k, int32_t* idx, float* dist, cudaStream_t s) {
  constexpr int T = KcCfg<K>::T;
  constexpr int R = KcCfg<K>::R;
  const int nc = G * G * G;
  const size_t list = kc_list_bytes(R, T), stage = kc_blob_bytes(n, nc);
  const bool staged = list + stage <= 100 * 1024;  // otherwise the arrangement is read through L1 / L2
  dim3 grid(ceil_div(n, T), b, 1);
  cudaError_t e = cudaSuccess;
  if (staged) {
    static PerDeviceOnce once;
    if (once.needed()) {
      e = cudaFuncSetAttribute(knn_cells_kernel<K, T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      if (e != cudaSuccess) return (int)e;
      once.done();
    }
    knn_cells_kernel<K, T, true><<<grid, T, list + stage, s>>>(blobs, n, G, kout, drop, hint, hint_k, idx, dist, min(kreq, n));
  } else {
    knn_cells_kernel<K, T, false><<<grid, T, list, s>>>(blobs, n, G, kout, drop, hint, hint_k, idx, dist, min(kreq, n));
  }
  return GEOA3_LAUNCH_RESULT();
}

}  // namespace geoa3

extern "C" int geoa3_cell_grid_max(int n) {
  // largest G whose sorting pass fits one CTA's shared memory
  int g = 1;
  while (g < 32 && geoa3::kc_sort_smem(n, (g + 1) * (g + 1) * (g + 1)) <= 220 * 1024) ++g;
  return g;
}

extern "C" size_t geoa3_cell_blob_bytes(int n, int G) { return geoa3::kc_blob_bytes(n, G * G * G); }

extern "C" int geoa3_cell_sort(const float* pc, int b, int n, int G, void* blobs, geoa3_stream_t stream) {
  using namespace geoa3;
  GEOA3_CHECK_ARG(pc && blobs && b > 0 && n > 0 && G >= 1);
  if (n > 65535 || b > 65535 || G > geoa3_cell_grid_max(n)) return GEOA3_EUNSUPPORTED;
  GEOA3_CHECK_ARG((reinterpret_cast<uintptr_t>(blobs) & 15) == 0);
  const size_t smem = kc_sort_smem(n, G * G * G);
  static PerDeviceOnce once;
  if (once.needed()) {
    cudaError_t e = cudaFuncSetAttribute(cell_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return (int)e;
    once.done();
  }
  cell_sort_kernel<<<b, KC_SORT_THREADS, smem, (cudaStream_t)stream>>>(pc, n, G, reinterpret_cast<unsigned char*>(blobs));
  return GEOA3_LAUNCH_RESULT();
}

extern "C" int geoa3_knn_cells(const void* blobs, int b, int n, int G, int K, int drop, const int32_t* hint, int hint_k,
                               int32_t* idx, float* dist, geoa3_stream_t stream) {
  using namespace geoa3;
  GEOA3_CHECK_ARG(blobs && idx);
  GEOA3_CHECK_ARG(b > 0 && n > 0 && G >= 1 && K > 0 && drop >= 0 && drop < K && hint_k >= 0);
  GEOA3_CHECK_ARG((reinterpret_cast<uintptr_t>(blobs) & 15) == 0);
  if (K > GEOA3_KNN_MAX_K || b > 65535 || n > 65535 || G > 32) return GEOA3_EUNSUPPORTED;
  if (K > n) return GEOA3_EINVAL;
  const unsigned char* bl = reinterpret_cast<const unsigned char*>(blobs);
  cudaStream_t s = (cudaStream_t)stream;
  const int kout = K - drop;
#define GEOA3_KC_ARGS bl, b, n, G, K, kout, drop, hint, hint_k, idx, dist, s
  if (K <= 3) return launch_knn_cells_k<3>(GEOA3_KC_ARGS);
  if (K <= 5) return launch_knn_cells_k<5>(GEOA3_KC_ARGS);
  if (K <= 9) return launch_knn_cells_k<9>(GEOA3_KC_ARGS);
  if (K <= 17) return launch_knn_cells_k<17>(GEOA3_KC_ARGS);
  return launch_knn_cells_k<33>(GEOA3_KC_ARGS);
#undef GEOA3_KC_ARGS
}
