// csr.cuh — deterministic CSR-by-target construction inside one CTA (shared by the loss backward and
// the pointnet2 scatter-gradients).  It is a stable counting sort: the E sources are split into W
// contiguous chunks, one per warp; a warp walks its chunk 32 sources at a time in ascending order and
// uses match.any to rank equal keys inside the 32-wide group, per-(warp,target) counters carry the rank
// across groups and across warps.  Result: segment p of `ent` lists the sources whose key is p in
// ascending source order, independent of scheduling => gathers that replace float atomics are bitwise
// reproducible.
#pragma once
#include "common.cuh"

namespace geoa3 {

// keys[e], e in [0,E), values in [0,n).  offs[n+1] (shared or global), whist = W*n counters of shared
// scratch (HistT = uint16_t when E < 65536, else int), ent[E] receives e / src_div.
// scan_scratch: THREADS/32 ints of shared memory.  Must be called by all THREADS threads of the CTA;
// ends with a __syncthreads().  Keys are fetched CSR_BATCH groups at a time so the L2 latency of the
// key loads is paid once per batch, not once per 32 sources.
constexpr int CSR_BATCH = 8;

// counter += v, returns the previous value.  uint16 counters are updated through a 32-bit shared atomic on
// the containing word (counts stay < 65536, so no carry crosses the halves).  Only ONE lane per distinct key
// issues the atomic in any instruction and a warp's instructions reach the LSU in program order, so the
// sequence of values a (warp, key) counter takes is fixed: using atomics here does not cost determinism,
// it only removes the load -> add -> store -> __syncwarp round trip from the dependency chain.
__device__ __forceinline__ int hist_fetch_add(uint16_t* h, int key, int v) {
  unsigned* w = reinterpret_cast<unsigned*>(h) + (key >> 1);
  const int sh = (key & 1) * 16;
  const unsigned old = atomicAdd(w, (unsigned)v << sh);
  return (int)((old >> sh) & 0xffffu);
}
__device__ __forceinline__ int hist_fetch_add(int* h, int key, int v) { return atomicAdd(h + key, v); }

template <int THREADS, typename EntT, typename HistT>
__device__ void build_csr(const int32_t* __restrict__ keys, int E, int n, unsigned src_magic, int* offs, HistT* whist,
                          int W, EntT* ent, int* scan_scratch) {
  // src_magic: 0 => ent = e; else ent = e / d computed as umulhi(e, ceil(2^32/d)) (exact for e, d < 2^16)
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int n2 = (n + 1) & ~1;  // row stride: keeps every uint16 row 4-byte aligned for the packed atomics
  for (int i = tid; i < W * n2; i += THREADS) whist[i] = 0;
  __syncthreads();
  const int chunk = ((E + W - 1) / W + 31) & ~31;
  const int e_beg = w * chunk, e_end = min(E, e_beg + chunk);
  if (w < W) {
    HistT* h = whist + w * n2;
    for (int e0 = e_beg; e0 < e_end; e0 += 32 * CSR_BATCH) {
      int kreg[CSR_BATCH];
#pragma unroll
      for (int u = 0; u < CSR_BATCH; ++u) {
        const int e = e0 + u * 32 + lane;
        kreg[u] = e < e_end ? keys[e] : -1 - lane;  // inactive lanes get unique dummy keys
      }
#pragma unroll
      for (int u = 0; u < CSR_BATCH; ++u) {
        const int key = kreg[u];
        const unsigned mask = __match_any_sync(0xffffffffu, key);
        if (key >= 0 && (__ffs(mask) - 1) == lane) hist_fetch_add(h, key, __popc(mask));
      }
    }
  }
  __syncthreads();
  // exclusive scan over targets of the per-target totals; per-(warp,target) cursors written back
  const int per = (n + THREADS - 1) / THREADS;
  const int p_beg = min(n, tid * per), p_end = min(n, p_beg + per);
  int local = 0;
  for (int p = p_beg; p < p_end; ++p)
    for (int ww = 0; ww < W; ++ww) local += whist[ww * n2 + p];
  int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) scan_scratch[w] = incl;
  __syncthreads();
  int base = 0;
  for (int i = 0; i < w; ++i) base += scan_scratch[i];
  int run = base + incl - local;
  for (int p = p_beg; p < p_end; ++p) {
    offs[p] = run;
    for (int ww = 0; ww < W; ++ww) {
      const int c = whist[ww * n2 + p];
      whist[ww * n2 + p] = (HistT)run;
      run += c;
    }
  }
  if (tid == THREADS - 1) offs[n] = E;
  __syncthreads();
  if (w < W) {
    HistT* h = whist + w * n2;
    for (int e0 = e_beg; e0 < e_end; e0 += 32 * CSR_BATCH) {
      int kreg[CSR_BATCH];
#pragma unroll
      for (int u = 0; u < CSR_BATCH; ++u) {
        const int e = e0 + u * 32 + lane;
        kreg[u] = e < e_end ? keys[e] : -1 - lane;
      }
#pragma unroll
      for (int u = 0; u < CSR_BATCH; ++u) {
        const int key = kreg[u];
        const int e = e0 + u * 32 + lane;
        const unsigned mask = __match_any_sync(0xffffffffu, key);
        const int leader = __ffs(mask) - 1;
        int cur = 0;
        if (key >= 0 && leader == lane) cur = hist_fetch_add(h, key, __popc(mask));
        cur = __shfl_sync(0xffffffffu, cur, leader);  // group members take the leader's base
        if (key >= 0)
          ent[cur + __popc(mask & ((1u << lane) - 1u))] = (EntT)(src_magic ? __umulhi((unsigned)e, src_magic) : (unsigned)e);
      }
    }
  }
  __syncthreads();
}


// ---------------------------------------------------------------------------------------------------------
// Variant for SHORT segments (kNN in-degree, Chamfer fan-in): MATCH.ANY costs ~63 SM-cycles per warp-op on
// B200 when the 32 keys are distinct (ubench/match_atoms.cu) whereas a shared atomic costs ~5, so here the
// slots are handed out by plain shared atomics (arrival order is arbitrary) and every target then sorts its
// own segment by source id (insertion sort in shared memory; sources are distinct inside a segment, so the
// sorted order is unique).  Same result as build_csr: ascending sources per target, independent of timing.
// cnt: n ints of shared scratch.  Ends with a __syncthreads().
template <int THREADS, typename EntT>
__device__ void build_csr_sorted(const int32_t* __restrict__ keys, int E, int n, unsigned src_magic, int* offs, int* cnt,
                                 EntT* ent, int* scan_scratch) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int i = tid; i < n; i += THREADS) cnt[i] = 0;
  __syncthreads();
  for (int e = tid; e < E; e += THREADS) atomicAdd(&cnt[keys[e]], 1);
  __syncthreads();
  const int per = (n + THREADS - 1) / THREADS;
  const int p_beg = min(n, tid * per), p_end = min(n, p_beg + per);
  int local = 0;
  for (int p = p_beg; p < p_end; ++p) local += cnt[p];
  int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) scan_scratch[w] = incl;
  __syncthreads();
  int run = incl - local;
  for (int i = 0; i < w; ++i) run += scan_scratch[i];
  for (int p = p_beg; p < p_end; ++p) {
    const int c = cnt[p];
    offs[p] = run;
    cnt[p] = run;  // becomes the fill cursor
    run += c;
  }
  if (tid == THREADS - 1) offs[n] = E;
  __syncthreads();
  // Fill in waves of THREADS consecutive sources with a barrier in between: slots inside a wave are taken in
  // arbitrary order, but waves are ordered, so every segment is already sorted at wave granularity and the
  // insertion sort below only has to fix the (few) inversions inside a wave.
  {
    int e = tid;
    int key = e < E ? keys[e] : -1;
    for (int e0 = 0; e0 < E; e0 += THREADS) {
      const int e_next = e + THREADS;
      const int key_next = e_next < E ? keys[e_next] : -1;  // prefetch: hides the L2 latency behind the barrier
      if (key >= 0) ent[atomicAdd(&cnt[key], 1)] = (EntT)(src_magic ? __umulhi((unsigned)e, src_magic) : (unsigned)e);
      __syncthreads();
      e = e_next;
      key = key_next;
    }
  }
  for (int p = tid; p < n; p += THREADS) {
    const int b0 = offs[p], L = offs[p + 1] - b0;
    for (int a = 1; a < L; ++a) {
      const EntT v = ent[b0 + a];
      int c = a - 1;
      while (c >= 0 && ent[b0 + c] > v) { ent[b0 + c + 1] = ent[b0 + c]; --c; }
      ent[b0 + c + 1] = v;
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------
// Variant for a neighbour table nbr[n][k] (k <= 32) whose ROWS hold distinct targets (true for any kNN list):
// a warp instruction that only touches entries of one row (two rows when k <= 16, issued as two predicated
// halves) never sees the same counter twice, so plain shared atomics on per-(warp,target) counters already
// return deterministic ranks — no MATCH.ANY (63 SM-cycles per op), no sorting.  Warps own ascending row
// ranges and walk them in order, hence segment p lists its sources (rows) in ascending order.
// whist: W * ((n+1)&~1) uint16 counters of shared scratch.  Ends with a __syncthreads().
template <int THREADS>
__device__ void build_csr_rows(const int32_t* __restrict__ nbr, int n, int k, int* offs, uint16_t* whist, int W,
                               uint16_t* ent, int* scan_scratch) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int n2 = (n + 1) & ~1;
  for (int i = tid; i < W * n2; i += THREADS) whist[i] = 0;
  __syncthreads();
  const int rpi = k <= 16 ? 2 : 1;               // rows per instruction group
  const int half = k <= 16 ? (lane >> 4) : 0;    // which of them this lane serves
  const int t = k <= 16 ? (lane & 15) : lane;    // neighbour slot
  const bool lane_on = t < k;
  const int rows_per_warp = (((n + W - 1) / W) + 1) & ~1;
  const int r_beg = w * rows_per_warp, r_end = min(n, r_beg + rows_per_warp);
  constexpr int B = 8;
  if (w < W) {
    uint16_t* h = whist + w * n2;
    for (int r0 = r_beg; r0 < r_end; r0 += rpi * B) {
      int kreg[B];
#pragma unroll
      for (int u = 0; u < B; ++u) {
        const int row = r0 + u * rpi + half;
        kreg[u] = (lane_on && row < r_end) ? nbr[(size_t)row * k + t] : -1;
      }
#pragma unroll
      for (int u = 0; u < B; ++u) {
        if (half == 0 && kreg[u] >= 0) hist_fetch_add(h, kreg[u], 1);
        if (half == 1 && kreg[u] >= 0) hist_fetch_add(h, kreg[u], 1);
      }
    }
  }
  __syncthreads();
  const int per = (n + THREADS - 1) / THREADS;
  const int p_beg = min(n, tid * per), p_end = min(n, p_beg + per);
  int local = 0;
  for (int p = p_beg; p < p_end; ++p)
    for (int ww = 0; ww < W; ++ww) local += whist[ww * n2 + p];
  int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) scan_scratch[w] = incl;
  __syncthreads();
  int run = incl - local;
  for (int i = 0; i < w; ++i) run += scan_scratch[i];
  for (int p = p_beg; p < p_end; ++p) {
    offs[p] = run;
    for (int ww = 0; ww < W; ++ww) {
      const int c = whist[ww * n2 + p];
      whist[ww * n2 + p] = (uint16_t)run;
      run += c;
    }
  }
  if (tid == THREADS - 1) offs[n] = n * k;
  __syncthreads();
  if (w < W) {
    uint16_t* h = whist + w * n2;
    for (int r0 = r_beg; r0 < r_end; r0 += rpi * B) {
      int kreg[B];
#pragma unroll
      for (int u = 0; u < B; ++u) {
        const int row = r0 + u * rpi + half;
        kreg[u] = (lane_on && row < r_end) ? nbr[(size_t)row * k + t] : -1;
      }
#pragma unroll
      for (int u = 0; u < B; ++u) {
        const int row = r0 + u * rpi + half;
        if (half == 0 && kreg[u] >= 0) ent[hist_fetch_add(h, kreg[u], 1)] = (uint16_t)row;
        if (half == 1 && kreg[u] >= 0) ent[hist_fetch_add(h, kreg[u], 1)] = (uint16_t)row;
      }
    }
  }
  __syncthreads();
}

}  // namespace geoa3
