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

template <int THREADS, typename EntT, typename HistT>
__device__ void build_csr(const int32_t* __restrict__ keys, int E, int n, int src_div, int* offs, HistT* whist, int W,
                          EntT* ent, int* scan_scratch) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int i = tid; i < W * n; i += THREADS) whist[i] = 0;
  __syncthreads();
  const int chunk = ((E + W - 1) / W + 31) & ~31;
  const int e_beg = w * chunk, e_end = min(E, e_beg + chunk);
  if (w < W) {
    HistT* h = whist + w * n;
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
        if (key >= 0 && (__ffs(mask) - 1) == lane) h[key] = (HistT)(h[key] + __popc(mask));
        __syncwarp();
      }
    }
  }
  __syncthreads();
  // exclusive scan over targets of the per-target totals; per-(warp,target) cursors written back
  const int per = (n + THREADS - 1) / THREADS;
  const int p_beg = min(n, tid * per), p_end = min(n, p_beg + per);
  int local = 0;
  for (int p = p_beg; p < p_end; ++p)
    for (int ww = 0; ww < W; ++ww) local += whist[ww * n + p];
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
      const int c = whist[ww * n + p];
      whist[ww * n + p] = (HistT)run;
      run += c;
    }
  }
  if (tid == THREADS - 1) offs[n] = E;
  __syncthreads();
  if (w < W) {
    HistT* h = whist + w * n;
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
        const bool ok = key >= 0;
        const unsigned mask = __match_any_sync(0xffffffffu, key);
        const int rank = __popc(mask & ((1u << lane) - 1u));
        const int cur = ok ? (int)h[key] : 0;
        __syncwarp();
        if (ok) {
          ent[cur + rank] = (EntT)(e / src_div);
          if ((__ffs(mask) - 1) == lane) h[key] = (HistT)(cur + __popc(mask));
        }
        __syncwarp();
      }
    }
  }
  __syncthreads();
}

}  // namespace geoa3
