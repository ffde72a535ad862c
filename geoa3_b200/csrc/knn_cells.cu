// knn_cells.cu — the neighbour SETS of the curvature loss through a per-cloud CELL GRID (self-query, K = k+1 <= 33,
// clouds of up to 65 535 points).  Same members as geoa3_knn / geoa3_knn_set — the K lexicographically smallest
// (pinned distance, ORIGINAL index) pairs minus the `drop` smallest — found by looking only at the cells a query's
// search ball touches instead of streaming the whole cloud past every query.
//
//  geoa3_cell_sort   one CTA per cloud: bounding box, a grid of CUBIC cells (given, or sized per cloud from its surface
//                    density so that a ball of `kref` points is about one cell wide), counting sort by cell with a rank
//                    pass for the order inside a cell (cell-major, ascending original index: bitwise reproducible) ->
//                    one blob per cloud: the cloud as float4 (x, y, z, original index) in cell order, the cell start
//                    table and the inverse permutation (cells.cuh).  Cell index = (cz*gy + cy)*gx + cx, so one (cz, cy)
//                    ROW of cells is one contiguous range of positions.
//  geoa3_knn_cells   one query per thread, a warp's 32 queries are neighbours in cell order; the blob is staged by one
//                    TMA bulk copy.  tau = largest pinned distance to the hinted candidates (previous step's
//                    neighbours): an upper bound of the K-th distance whenever the hints are K-1 distinct points other
//                    than the query (verified at the end: fewer than K survivors => the warp searches again from
//                    tau = +inf).  For every cell row within sqrt(tau) of the query (conservative test on the row's
//                    y/z slab) the x interval that the ball can reach inside that row is mapped to a position range;
//                    every candidate in it is evaluated with the PINNED fma chain and kept when d <= tau (list in
//                    shared memory, arrival = ascending position).  The largest (distance, original index) keys beyond
//                    K are removed at the end (and whenever the list of the careful path overflows).  Members are
//                    written in list order = ascending position in the cell arrangement, a function of the cloud and
//                    its grid alone (never of the hint).
//
// Exactness.  cell(x) = trunc(clamp((x - lo) * inv_h, 0, g-1)) is monotone in x and is the SAME float function on both
// sides (sorting and querying), so |c.x - q.x| <= rx implies cell(q.x - rx) <= cell(c.x) <= cell(q.x + rx) whatever
// the rounding; radii are inflated (relative 1e-5 + `slack` = 8e-6 * max|coordinate|, ~64 ulp of the largest
// coordinate) over the rounding of the pinned chain (<= 5 ulp relative), of MUFU.RSQ, of the cell boundaries and of
// the sums; the y/z slab test treats the LAST cell layer of an axis as unbounded (the clamp puts everything beyond the
// grid there).  The grid only decides which candidates are LOOKED AT; what is kept is decided by the pinned arithmetic.
#include <climits>

#include "cells.cuh"

#ifndef KC_EXTRA
#define KC_EXTRA 11
#endif
#ifndef KC_THREADS
#define KC_THREADS 256
#endif
#ifndef KC_HSCALE
#define KC_HSCALE 1.0f   // cell edge / estimated radius of a kref-point ball (speed only)
#endif
#ifndef KC_ROWS
#define KC_ROWS 24   // row-table slots per query (fast path); a query with more non-empty rows takes the careful path
#endif

namespace geoa3 {

constexpr int KC_SORT_THREADS = 1024;

constexpr int KC_PROBE = 16;  // probe grid (cells per axis) of the surface-density estimate

__host__ __device__ inline size_t kc_sort_smem(int n, int ncap) {
  const int nct = ncap > KC_PROBE * KC_PROBE * KC_PROBE ? ncap : KC_PROBE * KC_PROBE * KC_PROBE;
  return (size_t)n * 4 + (size_t)(ncap + 1) * 4 + (size_t)nct * 4 + (size_t)((n + 1) & ~1) * 2 + 32 * 4 + 32 * 8 * 4 + KC_GP * 4;
}

// ncap: capacity of the cell table (the blob layout depends on it, not on the grid actually used).
// fx, fy, fz > 0: that grid, cubic cells of edge (longest side / largest f).  Otherwise the grid is chosen per cloud:
// cubic cells whose edge is about the radius of a ball holding `kref` points, estimated from the cloud's surface
// density (occupied cells of a 16^3 probe grid ~ surface area), as many cells per axis as the box needs, the edge
// grown until the grid fits ncap.  The choice only affects speed.
__global__ void __launch_bounds__(KC_SORT_THREADS)
cell_sort_kernel(const float* __restrict__ pc, int n, int ncap, float kref, int fx, int fy, int fz,
                 unsigned char* __restrict__ blobs) {
  extern __shared__ __align__(16) unsigned char kc_smem[];
  const int nc = ncap;
  const int nct = ncap > KC_PROBE * KC_PROBE * KC_PROBE ? ncap : KC_PROBE * KC_PROBE * KC_PROBE;
  int* keys = reinterpret_cast<int*>(kc_smem);      // [n]
  int* offs = keys + n;                             // [nc + 1]
  int* cnt = offs + nc + 1;                         // [max(nc, probe cells)]
  int* scan = cnt + nct;                            // [32]
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
      red[0] = lx; red[1] = ly; red[2] = lz; red[3] = hx; red[4] = hy; red[5] = hz;
    }
  }
  __syncthreads();
  const float blo[3] = {red[0], red[1], red[2]}, bhi[3] = {red[3], red[4], red[5]};
  float ext[3], emax = 0.f, ma = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float e = bhi[a] - blo[a];
    ext[a] = (e > 0.f && e < 3e38f) ? e : 0.f;  // empty / single-valued / non-finite axis: one cell layer
    emax = fmaxf(emax, ext[a]);
    ma = fmaxf(ma, fmaxf(fabsf(blo[a]), fabsf(bhi[a])));
  }
  int nocc = 0;
  if (fx <= 0 && emax > 0.f) {  // surface-density probe: how many cells of a 16^3 grid over the box hold points
    const float ip = (float)KC_PROBE / emax, pm1 = (float)(KC_PROBE - 1);
    for (int c = tid; c < KC_PROBE * KC_PROBE * KC_PROBE; c += KC_SORT_THREADS) cnt[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += KC_SORT_THREADS) {
      const int cx = kc_cell(p[i], blo[0], ip, pm1), cy = kc_cell(p[n + i], blo[1], ip, pm1), cz = kc_cell(p[2 * n + i], blo[2], ip, pm1);
      cnt[(cz * KC_PROBE + cy) * KC_PROBE + cx] = 1;
    }
    __syncthreads();
    int mine = 0;
    for (int c = tid; c < KC_PROBE * KC_PROBE * KC_PROBE; c += KC_SORT_THREADS) mine += cnt[c];
    mine = __reduce_add_sync(0xffffffffu, mine);
    if (lane == 0) scan[w] = mine;
    __syncthreads();
    for (int i = 0; i < KC_SORT_THREADS / 32; ++i) nocc += scan[i];
    __syncthreads();
  }
  if (tid == 0) {
    int g[3];
    float h;
    if (fx > 0) {
      const int fs[3] = {fx, fy, fz};
      h = emax / (float)max(fx, max(fy, fz));
#pragma unroll
      for (int a = 0; a < 3; ++a) g[a] = fs[a];
    } else {
      // area ~ nocc * (emax/16)^2, density = n / area, a ball of radius r on the surface holds kref points
      const float hp = emax / (float)KC_PROBE;
      h = sqrtf(kref * (float)max(nocc, 1) * hp * hp / (3.14159265f * (float)n)) * KC_HSCALE;
      h = fmaxf(h, emax / 64.f);
      for (int it = 0; it < 64; ++it) {
        long long tot = 1;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          g[a] = h > 0.f ? min(64, max(1, (int)ceilf(ext[a] / h))) : 1;
          tot *= g[a];
        }
        if (tot <= ncap) break;
        h *= 1.1f;
      }
      if ((long long)g[0] * g[1] * g[2] > ncap) { g[0] = g[1] = g[2] = 1; h = emax; }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const bool ok = ext[a] > 0.f && h > 0.f;
      gp[a] = blo[a] < 3e38f ? blo[a] : 0.f;
      gp[3 + a] = ok ? 1.f / h : 0.f;
      gp[6 + a] = ok ? h : 0.f;
      gp[10 + a] = (float)(g[a] - 1);
    }
    gp[9] = ma < 3e38f ? 8e-6f * ma + 1e-30f : 0.f;
    gp[13] = (float)g[0];
    gp[14] = (float)g[1];
    gp[15] = (float)g[2];
  }
  __syncthreads();
  if (tid < KC_HDR / 4) reinterpret_cast<float*>(blob)[tid] = tid < KC_GP ? gp[tid] : 0.f;
  const int gx = (int)gp[13], gy = (int)gp[14];
  for (int i = tid; i < n; i += KC_SORT_THREADS) {
    const int cx = kc_cell(p[i], gp[0], gp[3], gp[10]), cy = kc_cell(p[n + i], gp[1], gp[4], gp[11]),
              cz = kc_cell(p[2 * n + i], gp[2], gp[5], gp[12]);
    keys[i] = (cz * gy + cy) * gx + cx;
  }
  __syncthreads();
  for (int c = tid; c < nc; c += KC_SORT_THREADS) cnt[c] = 0;
  __syncthreads();
  // counting sort by cell.  Slots inside a cell are handed out by shared atomics in arbitrary order; the order is then
  // made canonical (ascending original index) by RANKING: every point counts the smaller indices of its own cell
  // segment and goes to segment start + rank.  All points of a cell rank in parallel, so a crowded cell costs its
  // length in steps, not its length squared (a single-thread insertion sort took 200 us on flat clouds).
  for (int i = tid; i < n; i += KC_SORT_THREADS) atomicAdd(&cnt[keys[i]], 1);
  __syncthreads();
  {
    const int per = (nc + KC_SORT_THREADS - 1) / KC_SORT_THREADS;
    const int c_beg = min(nc, tid * per), c_end = min(nc, c_beg + per);
    int local = 0;
    for (int c = c_beg; c < c_end; ++c) local += cnt[c];
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) scan[w] = incl;
    __syncthreads();
    int run = incl - local;
    for (int i = 0; i < w; ++i) run += scan[i];
    for (int c = c_beg; c < c_end; ++c) {
      const int k = cnt[c];
      offs[c] = run;
      cnt[c] = run;  // becomes the fill cursor
      run += k;
    }
    if (tid == KC_SORT_THREADS - 1) offs[nc] = n;
  }
  __syncthreads();
  for (int i = tid; i < n; i += KC_SORT_THREADS) ent[atomicAdd(&cnt[keys[i]], 1)] = (uint16_t)i;
  __syncthreads();
  float4* o4 = reinterpret_cast<float4*>(blob + KC_HDR);
  uint16_t* ip = reinterpret_cast<uint16_t*>(blob + kc_ip_off(n, nc));
  for (int t = tid; t < n; t += KC_SORT_THREADS) {
    const int o = ent[t];
    const int c = keys[o];
    const int b0 = offs[c], e0 = offs[c + 1];
    int rank = 0;
    for (int u = b0; u < e0; ++u) rank += (int)ent[u] < o;
    const int pos = b0 + rank;
    o4[pos] = make_float4(p[o], p[n + o], p[2 * n + o], __int_as_float(o));
    ip[o] = (uint16_t)pos;
  }
  uint16_t* cs = reinterpret_cast<uint16_t*>(blob + kc_cs_off(n));
  for (int c = tid; c < ((nc + 1 + 7) & ~7); c += KC_SORT_THREADS) cs[c] = (uint16_t)offs[min(c, nc)];
  for (int t = n + tid; t < ((n + 7) & ~7); t += KC_SORT_THREADS) ip[t] = 0;
}

template <int K>
struct KcCfg {
  static constexpr int R = K + KC_EXTRA;  // list slots per query
  static constexpr int T = K <= 17 ? KC_THREADS : (KC_THREADS > 256 ? 256 : KC_THREADS);  // one query per thread
  static constexpr int W = (KC_ROWS + 1) > R ? (KC_ROWS + 1) : R;  // words of per-query scratch (row table, then distances)
};

// Per-query shared memory: list of ORIGINAL indices (uint16 [R][T]) + a scratch column of 32-bit words [W][T] that
// holds the query's row table while it scans and the members' pinned distances afterwards (dead entries: negative).
__host__ __device__ inline size_t kc_list_bytes(int r, int w, int t) { return ((size_t)r * t * 2 + (size_t)w * t * 4 + 15) & ~(size_t)15; }

// Marks the `rm` largest (distance, original index) keys of a query's list dead (d = -2).  Per thread, loop form:
// it runs under divergence.  d_ / v_: the thread's columns (stride COLS).
template <int COLS>
__device__ __forceinline__ void kc_mark(float* __restrict__ d_, const uint16_t* __restrict__ v_, int cnt, int rm) {
  for (; rm > 0; --rm) {
    float md = -1.f;
    int bs = 0;
    unsigned bi = 0u;
    for (int s = 0; s < cnt; ++s) {
      const float d = d_[s * COLS];
      if (d >= md) {
        const unsigned vi = v_[s * COLS];
        if (d > md || vi > bi) { md = d; bs = s; bi = vi; }
      }
    }
    d_[bs * COLS] = -2.f;
  }
}

// List overflow (careful path only): keeps the kk smallest keys in arrival order, returns the new length and lowers
// *tau to the largest kept distance.
template <int COLS>
__device__ __forceinline__ int kc_cut(float* __restrict__ d_, uint16_t* __restrict__ v_, int cnt, int kk, float* tau) {
  if (cnt <= kk) return cnt;
  kc_mark<COLS>(d_, v_, cnt, cnt - kk);
  int w = 0;
  float mx = 0.f;
  for (int s = 0; s < cnt; ++s) {  // close the gaps
    const float d = d_[s * COLS];
    if (d >= 0.f) {
      d_[w * COLS] = d;
      v_[w * COLS] = v_[s * COLS];
      mx = fmaxf(mx, d);
      ++w;
    }
  }
  *tau = fminf(*tau, mx);
  return w;
}

// Writes the live entries of a list (arrival order = ascending position in the cell arrangement).
template <int COLS>
__device__ __forceinline__ void kc_write(const float* __restrict__ d_, const uint16_t* __restrict__ v_, int cnt, int kout,
                                         int32_t* __restrict__ io, float* __restrict__ dn) {
  int o = 0;
  for (int s = 0; s < cnt && o < kout; ++s) {
    const float d = d_[s * COLS];
    if (d >= 0.f) {
      io[o] = (int32_t)v_[s * COLS];
      if (dn) dn[o] = d;
      ++o;
    }
  }
}

template <int K, int T, bool STAGED>
__global__ void __launch_bounds__(T, 1024 / T)
knn_cells_kernel(const unsigned char* __restrict__ blobs, int n, int ncap, int kout, int drop,
                 const int32_t* hint, int hint_k, int32_t* idx_out, float* __restrict__ dist_out, int kk) {
  // kk = min(requested K, n): the list target (the template K is the capacity class)
  constexpr int R = KcCfg<K>::R;
  constexpr int WS = KcCfg<K>::W;
  constexpr int HV = (K - 1) % 4 == 0 ? (K - 1) / 4 : 0;  // hint row as int4 registers when hint_k == K-1
  extern __shared__ __align__(16) unsigned char kc_smem[];
  const int nc = ncap;
  const int cloud = blockIdx.y, tid = threadIdx.x;
  const size_t blob_bytes = kc_blob_bytes(n, nc);
  const unsigned char* gblob = blobs + (size_t)cloud * blob_bytes;
  float* d_ = reinterpret_cast<float*>(kc_smem) + tid;                               // scratch column of [WS][T]
  uint16_t* v_ = reinterpret_cast<uint16_t*>(kc_smem + (size_t)WS * T * 4) + tid;   // list column of [R][T]
  const unsigned char* blob = gblob;
  __shared__ __align__(8) unsigned long long kc_bar;
  if (STAGED) {  // the whole blob is one contiguous, 16-byte sized block: a single TMA bulk copy stages it
    unsigned char* sblob = kc_smem + kc_list_bytes(R, WS, T);
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
  const float slack = sgp[9], gmx = sgp[10], gmy = sgp[11], gmz = sgp[12];
  const int gx = (int)sgp[13], gy = (int)sgp[14];  // this cloud's grid (cells per axis)

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

  // ---- fast path (hinted queries, staged blob).  Phase 1: the lane's non-empty cell rows within reach of tau go into
  // its row table (row offsets relative to the lane's first row are warp-uniform; the x interval per row comes from
  // what the row's y/z slab distance leaves of tau).  Phase 2: every lane walks ITS OWN candidate stream — the warp
  // runs as long as its longest stream, not as long as the sum of the longest rows — with branch-free appends of the
  // ORIGINAL index; the query itself is skipped (drop == 1: it is the smallest key unless a second point sits at
  // distance 0).  Phase 3: the members' pinned distances are recomputed into the (now free) row-table column, the
  // largest keys beyond K are marked and the rest is written.  A list or row-table overflow, a zero-distance duplicate
  // or a bound that turns out invalid (fewer than K survivors) sends the query to the careful path below.
  bool careful = live;
  const bool fa = STAGED && live && tau < KC_INF && drop <= 1;
  if (STAGED && __any_sync(0xffffffffu, fa)) {
    const int qskip = drop == 1 ? qo : -1;
    int z0 = 0, y0 = 0, nz = -1, ny = -1;
    if (fa) {
      const float r = kc_sqrt(tau * KC_REL) * KC_REL + slack;
      z0 = kc_cell(q.z - r, loz, ihz, gmz);
      nz = kc_cell(q.z + r, loz, ihz, gmz) - z0;
      y0 = kc_cell(q.y - r, loy, ihy, gmy);
      ny = kc_cell(q.y + r, loy, ihy, gmy) - y0;
    }
    const int wz = __reduce_max_sync(0xffffffffu, nz), wy = __reduce_max_sync(0xffffffffu, ny);
    const float r2 = fmaf(tau, KC_REL, 1e-37f);
    const float hy = sgp[7], hz = sgp[8];
    const unsigned a4 = (unsigned)__cvta_generic_to_shared(s4);
    const unsigned acs = (unsigned)__cvta_generic_to_shared(scs);
    const unsigned aip = (unsigned)__cvta_generic_to_shared(sip);
    const unsigned l0 = (unsigned)__cvta_generic_to_shared(v_);
    const unsigned lend = l0 + R * T * 2;
    const unsigned t0 = (unsigned)__cvta_generic_to_shared(d_);
    unsigned tp = t0;
    int tot = 0, nrow = 0;
    bool bad = false;
    for (int oz = 0; oz <= wz; ++oz) {
      const int rz = z0 + min(oz, max(nz, 0));
      const float zl = fmaf((float)rz, hz, loz);
      const float ez = fmaxf(fmaxf(zl - q.z, q.z - ((float)rz >= gmz ? KC_INF : zl + hz)) - slack, 0.f);  // lower bound of |c.z - q.z| in this slab
      const float remz = oz <= nz ? r2 - ez * ez : -1.f;
      for (int oy = 0; oy <= wy; ++oy) {
        const int ry = y0 + min(oy, max(ny, 0));
        const float yl = fmaf((float)ry, hy, loy);
        const float ey = fmaxf(fmaxf(yl - q.y, q.y - ((float)ry >= gmy ? KC_INF : yl + hy)) - slack, 0.f);
        const float rem = oy <= ny ? remz - ey * ey : -1.f;  // what is left for (c.x - q.x)^2
        const float rx = kc_sqrt(fmaxf(rem, 0.f)) * KC_REL + slack;
        const unsigned ab = acs + (unsigned)((rz * gy + ry) * gx) * 2u;
        const int x0 = kc_cell(q.x - rx, lox, ihx, gmx), x1 = kc_cell(q.x + rx, lox, ihx, gmx);
        const int s = (int)kc_lds16(ab + x0 * 2);
        const int len = rem >= 0.f ? (int)kc_lds16(ab + x1 * 2 + 2) - s : 0;
        if (len > 0) {
          if (nrow < KC_ROWS) kc_sts32(tp, (unsigned)s | ((unsigned)len << 16));
          else bad = true;
          tp += T * 4;
          ++nrow;
          tot += len;
        }
      }
    }
    kc_sts32(t0 + min(nrow, KC_ROWS) * (T * 4), 0xffff0000u);  // sentinel: a finished lane idles on it
    const int wtot = __reduce_max_sync(0xffffffffu, bad ? 0 : tot);
    if (bad) tot = 0;
    unsigned lp = l0;  // next free list slot (keeps advancing past the end: that is how an overflow is seen)
    {
      const unsigned aend = a4 + (unsigned)(n - 1) * 16u;
      unsigned ca = a4;
      int left = 0;
      tp = t0;
#pragma unroll 2
      for (int it = 0; it < wtot; ++it) {
        if (left == 0) {
          const unsigned e = kc_lds32(tp);
          tp += T * 4;
          ca = a4 + (e & 0xffffu) * 16u;
          left = (int)(e >> 16);
        }
        const float4 c = kc_lds128(min(ca, aend));
        ca += 16u;
        --left;
        const float d = dist2(c.x, c.y, c.z, q.x, q.y, q.z);  // the pinned arithmetic decides
        const int ci = __float_as_int(c.w);
        const bool pass = it < tot && d <= tau && ci != qskip;
        if (pass && lp < lend) kc_sts16(lp, (unsigned)ci);
        if (pass) lp += T * 2;
      }
    }
    {  // (all lanes of the warp: the cooperative output below shuffles)
      const int cnt = (int)(lp - l0) / (T * 2);
      const int kt = kk - (drop == 1 ? 1 : 0);
      bool good = fa && !bad && cnt >= kt && cnt <= R;
      if (good) {  // the members' distances, into the scratch column (the row table is no longer needed)
        float dm = KC_INF;
        for (int s2 = 0; s2 < cnt; ++s2) {
          const float4 c = kc_lds128(a4 + kc_lds16(aip + (unsigned)v_[s2 * T] * 2u) * 16u);
          const float d = dist2(c.x, c.y, c.z, q.x, q.y, q.z);
          d_[s2 * T] = d;
          dm = fminf(dm, d);
        }
        good = drop != 1 || dm > 0.f;  // a zero-distance neighbour competes with the query for being the dropped key
      }
      if (good) kc_mark<T>(d_, v_, cnt, cnt - kt);
      // Output.  Per-thread stores would touch 32 different rows per instruction (4 bytes of 32 sectors each: the L1
      // store path then costs 32 cycles per instruction, ~25 % of the whole kernel); instead every thread compacts its
      // live members to the front of its list column and the warp writes 8 rows of 64 bytes per instruction (int4 per
      // lane, 4 lanes per row for kout = 16).  Other kout / an optional distance output / short clouds: plain stores.
      const int lpr = kout >> 2;  // lanes per row
      const bool coop = dist_out == nullptr && kk == K && (kout & 3) == 0 && lpr >= 1 && lpr <= 32 && (32 % lpr) == 0 &&
                        ((reinterpret_cast<uintptr_t>(idx_out) & 15) == 0);
      if (good && !coop) kc_write<T>(d_, v_, cnt, kout, io, dn);
      if (coop) {
        if (good) {
          int o = 0;
          for (int s2 = 0; s2 < cnt; ++s2) {
            const uint16_t vi = v_[s2 * T];
            if (d_[s2 * T] >= 0.f) { v_[o * T] = vi; ++o; }  // o <= s2: in place
          }
        }
        __syncwarp();
        const int lane = tid & 31, sub = lane % lpr, rpi = 32 / lpr;  // rows per instruction
        const uint16_t* vw = v_ - lane;  // column of lane 0 of this warp
        for (int r0 = 0; r0 < 32; r0 += rpi) {
          const int row = r0 + lane / lpr;
          const int rqo = __shfl_sync(0xffffffffu, qo, row);
          const bool rgood = __shfl_sync(0xffffffffu, (int)good, row) != 0;
          if (rgood) {
            int4 m4;
            m4.x = vw[(sub * 4 + 0) * T + row]; m4.y = vw[(sub * 4 + 1) * T + row];
            m4.z = vw[(sub * 4 + 2) * T + row]; m4.w = vw[(sub * 4 + 3) * T + row];
            *reinterpret_cast<int4*>(idx_out + ((size_t)cloud * n + rqo) * kout + sub * 4) = m4;
          }
        }
        __syncwarp();
      }
      if (good) careful = false;
    }
  }

  // ---- careful path: no usable hint, or the fast path gave up (list / row-table overflow, zero-distance duplicate,
  // invalid bound, drop > 1).  Such queries are rare in an attack step but expensive (their bound is loose), so the warp
  // takes them ONE AT A TIME with its 32 lanes on 32 consecutive candidates of a row: passes are appended in position
  // order by ballot, the list is cut when it overflows (tau tightens), the query itself is a candidate like any other
  // and the `drop` smallest keys are removed at the end; fewer than K survivors = the bound was not valid: the
  // query is searched again from +inf.  (A lane-per-query walk made the whole warp wait on its one slow lane: up to
  // 17 us of the kernel on late attack states.)
  unsigned todo = __ballot_sync(0xffffffffu, careful);
  const int lane = tid & 31;
  while (todo) {
    const int L = __ffs(todo) - 1;
    todo &= todo - 1;
    const float qx = __shfl_sync(0xffffffffu, q.x, L), qy = __shfl_sync(0xffffffffu, q.y, L),
                qz = __shfl_sync(0xffffffffu, q.z, L);
    float tl = __shfl_sync(0xffffffffu, tau, L);  // the query's bound (warp-uniform from here on)
    float* dc = d_ - lane + L;                    // the query's columns
    uint16_t* vc = v_ - lane + L;
    int cnt = 0;
    for (;;) {
      const float r = (tl < 3e38f ? kc_sqrt(tl * KC_REL) * KC_REL : KC_INF) + slack;
      const int z0 = kc_cell(qz - r, loz, ihz, gmz), z1 = kc_cell(qz + r, loz, ihz, gmz);
      const int y0 = kc_cell(qy - r, loy, ihy, gmy), y1 = kc_cell(qy + r, loy, ihy, gmy);
      for (int rz = z0; rz <= z1; ++rz) {
        const float hz = sgp[8];
        const float zl = fmaf((float)rz, hz, loz);
        const float ez = fmaxf(fmaxf(zl - qz, qz - ((float)rz >= gmz ? KC_INF : zl + hz)) - slack, 0.f);
        for (int ry = y0; ry <= y1; ++ry) {
          const float hy = sgp[7];
          const float yl = fmaf((float)ry, hy, loy);
          const float ey = fmaxf(fmaxf(yl - qy, qy - ((float)ry >= gmy ? KC_INF : yl + hy)) - slack, 0.f);
          const float rem = fmaf(tl, KC_REL, 1e-37f) - ez * ez - ey * ey;  // with the CURRENT bound
          if (!(rem >= 0.f)) continue;
          const float rx = (rem < 3e38f ? kc_sqrt(rem) * KC_REL : KC_INF) + slack;
          const int base = (rz * gy + ry) * gx;
          const int s = scs[base + kc_cell(qx - rx, lox, ihx, gmx)], e = scs[base + kc_cell(qx + rx, lox, ihx, gmx) + 1];
          for (int p0 = s; p0 < e; p0 += 32) {
            const int pos = p0 + lane;
            const float4 c = s4[min(pos, e - 1)];
            const float d = dist2(c.x, c.y, c.z, qx, qy, qz);  // the pinned arithmetic decides
            const int ci = __float_as_int(c.w);
            const unsigned pass = __ballot_sync(0xffffffffu, pos < e && d <= tl);
            if (cnt + __popc(pass) <= R) {
              if ((pass >> lane) & 1u) {
                const int slot = cnt + __popc(pass & ((1u << lane) - 1u));
                dc[slot * T] = d;
                vc[slot * T] = (uint16_t)ci;
              }
              cnt += __popc(pass);
            } else {  // the list overflows inside this group: one entry at a time, cutting when full
              unsigned m2 = pass;
              while (m2) {
                const int bl = __ffs(m2) - 1;
                m2 &= m2 - 1;
                const float db = __shfl_sync(0xffffffffu, d, bl);
                const int cb = __shfl_sync(0xffffffffu, ci, bl);
                if (db <= tl) {  // (tl may have dropped since the ballot)
                  if (cnt == R) {
                    __syncwarp();
                    if (lane == 0) cnt = kc_cut<T>(dc, vc, cnt, kk, &tl);
                    cnt = __shfl_sync(0xffffffffu, cnt, 0);
                    tl = __shfl_sync(0xffffffffu, tl, 0);
                  }
                  if (lane == 0 && db <= tl) { dc[cnt * T] = db; vc[cnt * T] = (uint16_t)cb; }
                  if (db <= tl) ++cnt;
                }
              }
            }
          }
        }
      }
      if (cnt >= kk || !(tl < 3e38f)) break;  // (already searched from +inf: non-finite coordinates, nothing more to find)
      tl = KC_INF;  // fewer than K survivors: the bound was not valid for this query
      cnt = 0;
    }
    __syncwarp();
    if (lane == L) {  // the query's own lane finishes its list (rare path: plain loops)
      kc_mark<T>(d_, v_, cnt, cnt - kk);
      for (int t = 0; t < drop; ++t) {  // the `drop` smallest (distance, original index) keys are no members
        float md = KC_INF;
        int bs = 0;
        unsigned bo = 0xffffffffu;
        for (int s2 = 0; s2 < cnt; ++s2) {
          const float d = d_[s2 * T];
          if (d >= 0.f && d <= md) {
            const unsigned vi = v_[s2 * T];
            if (d < md || vi < bo) { md = d; bs = s2; bo = vi; }
          }
        }
        d_[bs * T] = -1.f;
      }
      kc_write<T>(d_, v_, cnt, kout, io, dn);
    }
    __syncwarp();
  }
}

template <int K>
static int launch_knn_cells_k(const unsigned char* blobs, int b, int n, int ncap, int kreq, int kout, int drop,
                              const int32_t* hint, int hint_k, int32_t* idx, float* dist, cudaStream_t s) {
  constexpr int T = KcCfg<K>::T;
  constexpr int R = KcCfg<K>::R;
  const int nc = ncap;
  const size_t list = kc_list_bytes(R, KcCfg<K>::W, T), stage = kc_blob_bytes(n, nc);
  const bool staged = list + stage <= 226 * 1024;  // otherwise the arrangement is read through L1 / L2 (careful path only)
  dim3 grid(ceil_div(n, T), b, 1);
  cudaError_t e = cudaSuccess;
  if (staged) {
    static PerDeviceOnce once;
    if (once.needed()) {
      e = cudaFuncSetAttribute(knn_cells_kernel<K, T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
      if (e != cudaSuccess) return (int)e;
      once.done();
    }
    knn_cells_kernel<K, T, true><<<grid, T, list + stage, s>>>(blobs, n, ncap, kout, drop, hint, hint_k, idx, dist, min(kreq, n));
  } else {
    static PerDeviceOnce once;
    if (once.needed()) {
      e = cudaFuncSetAttribute(knn_cells_kernel<K, T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
      if (e != cudaSuccess) return (int)e;
      once.done();
    }
    knn_cells_kernel<K, T, false><<<grid, T, list, s>>>(blobs, n, ncap, kout, drop, hint, hint_k, idx, dist, min(kreq, n));
  }
  return GEOA3_LAUNCH_RESULT();
}

}  // namespace geoa3

extern "C" int geoa3_cell_grid_max(int n) {
  // largest cell-table capacity whose sorting pass fits one CTA's shared memory
  if (n < 1) return 0;
  int lo = 1, hi = 65536;
  while (lo < hi) {
    const int mid = (lo + hi + 1) / 2;
    if (geoa3::kc_sort_smem(n, mid) <= 220 * 1024) lo = mid; else hi = mid - 1;
  }
  return lo;
}

extern "C" size_t geoa3_cell_blob_bytes(int n, int ncap) { return geoa3::kc_blob_bytes(n, ncap); }

extern "C" int geoa3_cell_sort(const float* pc, int b, int n, int ncap, float kref, int gx, int gy, int gz, void* blobs,
                               geoa3_stream_t stream) {
  using namespace geoa3;
  GEOA3_CHECK_ARG(pc && blobs && b > 0 && n > 0 && ncap >= 1 && gx >= 0 && gy >= 0 && gz >= 0);
  const bool fixed = gx > 0 || gy > 0 || gz > 0;
  if (fixed) GEOA3_CHECK_ARG(gx > 0 && gy > 0 && gz > 0);
  else GEOA3_CHECK_ARG(kref > 0.f);
  if (n > 65535 || b > 65535 || ncap > geoa3_cell_grid_max(n)) return GEOA3_EUNSUPPORTED;
  if (fixed && (gx > 64 || gy > 64 || gz > 64 || (long long)gx * gy * gz > ncap)) return GEOA3_EUNSUPPORTED;
  GEOA3_CHECK_ARG((reinterpret_cast<uintptr_t>(blobs) & 15) == 0);
  const size_t smem = kc_sort_smem(n, ncap);
  static PerDeviceOnce once;
  if (once.needed()) {
    cudaError_t e = cudaFuncSetAttribute(cell_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return (int)e;
    once.done();
  }
  cell_sort_kernel<<<b, KC_SORT_THREADS, smem, (cudaStream_t)stream>>>(pc, n, ncap, kref, gx, gy, gz,
                                                                      reinterpret_cast<unsigned char*>(blobs));
  return GEOA3_LAUNCH_RESULT();
}

extern "C" int geoa3_knn_cells(const void* blobs, int b, int n, int ncap, int K, int drop, const int32_t* hint, int hint_k,
                               int32_t* idx, float* dist, geoa3_stream_t stream) {
  using namespace geoa3;
  GEOA3_CHECK_ARG(blobs && idx);
  GEOA3_CHECK_ARG(b > 0 && n > 0 && ncap >= 1 && K > 0 && drop >= 0 && drop < K && hint_k >= 0);
  GEOA3_CHECK_ARG((reinterpret_cast<uintptr_t>(blobs) & 15) == 0);
  if (K > GEOA3_KNN_MAX_K || b > 65535 || n > 65535 || ncap > 65536) return GEOA3_EUNSUPPORTED;
  if (K > n) return GEOA3_EINVAL;
  const unsigned char* bl = reinterpret_cast<const unsigned char*>(blobs);
  cudaStream_t s = (cudaStream_t)stream;
  const int kout = K - drop;
#define GEOA3_KC_ARGS bl, b, n, ncap, K, kout, drop, hint, hint_k, idx, dist, s
  if (K <= 3) return launch_knn_cells_k<3>(GEOA3_KC_ARGS);
  if (K <= 5) return launch_knn_cells_k<5>(GEOA3_KC_ARGS);
  if (K <= 9) return launch_knn_cells_k<9>(GEOA3_KC_ARGS);
  if (K <= 17) return launch_knn_cells_k<17>(GEOA3_KC_ARGS);
  return launch_knn_cells_k<33>(GEOA3_KC_ARGS);
#undef GEOA3_KC_ARGS
}
