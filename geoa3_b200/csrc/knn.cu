// knn.cu — exact K-nearest-neighbour selection (K = k+1 <= 33) for the curvature losses.
//
// Thread-per-query brute force, candidates streamed through shared memory as SoA float4 broadcasts
// and evaluated two at a time on the packed fp32 pipe (pinned fma-chain arithmetic, see common.cuh).
// Each thread keeps its K best (dist, idx) pairs sorted in registers; tau = current K-th distance.
// The result is the K lexicographically smallest (dist, ORIGINAL index) pairs, ascending — the pinned rule.
//
// One alu op per candidate.  The alu pipe (compares, selects, integer adds) is half as wide as the fma pipe,
// so the per-candidate bookkeeping is reduced to a funnel shift: s = d - next(tau) is formed on the fma pipe
// (its sign is exact) and the sign bit is shifted into a 32-candidate pass mask.  Set bits are expanded into a
// per-thread index queue in shared memory once per 32 candidates; passing candidates are NOT inserted on the
// spot (that would serialise the warp on its slowest lane): the warp drains all queues together through a
// branch-free sorted-insert network when any lane's queue is nearly full.  The exact distance is recomputed at
// drain time and the network orders by (dist, original index), so the visiting order is irrelevant.
//
// Hinted threshold.  Scanning from tau = +inf inserts ~K*ln(N/K) candidates per query, and the insert network
// is what the alu pipe chokes on.  If the caller passes, per query, a list of candidate indices (typically
// the neighbours found one optimisation step earlier), tau starts at the largest exact distance to those
// candidates and to the query's own index — an upper bound of the true K-th distance whenever these are K
// distinct points — so only ~K candidates are ever inserted.  The bound is self-verifying: if fewer than K
// candidates passed it the CTA rescans with tau = +inf.  Any hint yields the exact top-K.
//
// Spatial pruning.  The caller may pass the clouds ALREADY ARRANGED in a visiting order (position t holds original
// point perm[t], e.g. the Morton order of the original cloud computed once per attack; staging stays coalesced):
// a warp's 32 queries and every group of 32 staged candidates are then spatially compact, each group carries its
// bounding box, and a warp skips the group when no lane's box distance can reach its tau.  Indices are reported
// in original numbering (through perm); any permutation is valid.
#include <climits>

#include "common.cuh"

namespace geoa3 {

constexpr int KNN_THREADS = 128;
constexpr int KNN_CHUNK = 1024;  // candidates per shared-memory pass (static smem stays under 48 KB)
constexpr int KNN_QDEPTH = 56;   // queue slots per thread (indices only; distances are recomputed at drain):
                                 // drained when > 24 are pending, and one 32-candidate group adds at most 32
constexpr float KNN_INF = __builtin_huge_valf();

template <int K>
struct TopK {
  float d[K];
  int i[K];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int l = 0; l < K; ++l) { d[l] = KNN_INF; i[l] = -1; }
  }
  // sorted insert of (x, xi); a no-op feed is (+inf, INT_MAX).  LEX: order by (distance, index) — needed when
  // candidates arrive in a permuted order; otherwise arrival order IS index order and "equal stays in front"
  // gives the same lexicographic result with 3 fewer alu ops per slot.
  template <bool LEX>
  __device__ __forceinline__ void insert(float x, int xi) {
    float cd = x;
    int ci = xi;
#pragma unroll
    for (int l = 0; l < K; ++l) {
      const bool p = LEX ? (x < d[l] || (x == d[l] && xi < i[l])) : (x < d[l]);
      const float od = d[l];
      const int oi = i[l];
      d[l] = p ? cd : od;
      i[l] = p ? ci : oi;
      cd = p ? od : cd;
      ci = p ? oi : ci;
    }
  }
  __device__ __forceinline__ float tau() const { return d[K - 1]; }
};

template <int K, bool PRUNE>
__global__ void __launch_bounds__(KNN_THREADS, K <= 17 ? 6 : 4)
knn_kernel(const float* __restrict__ query, const float* __restrict__ ref, int n, int m, int kout, int drop,
           const int32_t* __restrict__ perm_q, const int32_t* __restrict__ perm_c, const int32_t* __restrict__ iperm_c,
           const float* __restrict__ bb_c, const int32_t* hint, int hint_k, int32_t* idx_out,
           float* __restrict__ dist_out) {
  __shared__ __align__(16) float sx[KNN_CHUNK];
  __shared__ __align__(16) float sy[KNN_CHUNK];
  __shared__ __align__(16) float sz[KNN_CHUNK];
  __shared__ int so[PRUNE ? KNN_CHUNK : 1];             // original index of the staged candidate
  __shared__ __align__(16) float sbb[PRUNE ? KNN_CHUNK / 32 : 1][8];  // bounding boxes of the staged 32-candidate groups
  __shared__ uint16_t qj[KNN_QDEPTH][KNN_THREADS];  // positions inside the staged chunk (< KNN_CHUNK)

  const int cloud = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31;
  const int slot = blockIdx.x * KNN_THREADS + tid;  // position in visiting order
  const float* qbase = query + (size_t)cloud * 3 * n;
  const float* cbase = ref + (size_t)cloud * 3 * m;
  const int32_t* pc = perm_c ? perm_c + (size_t)cloud * m : nullptr;
  const int32_t* ipc = iperm_c ? iperm_c + (size_t)cloud * m : nullptr;
  const int sl = min(slot, n - 1);
  const int qq = perm_q ? perm_q[(size_t)cloud * n + sl] : sl;  // ORIGINAL index of this thread's query
  const bool live = slot < n;
  const int qpos = perm_q ? sl : qq;  // the query cloud is stored by position when perm_q is given
  const float qx = qbase[qpos], qy = qbase[n + qpos], qz = qbase[2 * n + qpos];
  const float2 nqx = make_float2(-qx, -qx), nqy = make_float2(-qy, -qy), nqz = make_float2(-qz, -qz);

  // hinted start threshold (evaluated below, once the first candidate chunk sits in shared memory)
  float tau0 = KNN_INF;
  const bool use_hint = hint != nullptr && hint_k + 1 >= K && (!pc || ipc);

  const int G0 = (m + 31) >> 5;  // level-0 boxes (32 candidates), followed by level-1 boxes (KNN_CHUNK candidates)
  const float* bbc = (PRUNE && bb_c) ? bb_c + (size_t)cloud * (G0 + (m + KNN_CHUNK - 1) / KNN_CHUNK) * 8 : nullptr;
  if (PRUNE && use_hint) {
    // chunks may be skipped before they are ever staged, so the hinted bound is evaluated up front from the
    // arranged cloud in global memory (positions through the inverse permutation)
    const int32_t* h = hint + ((size_t)cloud * n + qq) * hint_k;  // may alias idx_out: own row, read before write
    auto cand_d = [&](int j) {
      const int pos = ipc[min(max(j, 0), m - 1)];
      return dist2(cbase[pos], cbase[m + pos], cbase[2 * m + pos], qx, qy, qz);
    };
    float mx = cand_d(qq);
    for (int t = 0; t < hint_k; ++t) mx = fmaxf(mx, cand_d(h[t]));
    if (mx < 3.0e38f) tau0 = mx;
  }

  TopK<K> top;
  bool rescan = false;
  do {
  top.init();
  float tau = tau0;  // candidates with d <= tau are worth an exact look
  bool hint_pending = !PRUNE && use_hint && !rescan;
  // 32-bit shared-window address of this thread's queue column, kept in ONE register: the compiler otherwise
  // re-materialises the base (3 extra predicated instructions per candidate)
  const unsigned q0 = (unsigned)__cvta_generic_to_shared(&qj[0][tid]);
  constexpr unsigned QSTRIDE = KNN_THREADS * sizeof(uint16_t);
  unsigned qa = q0;
  int c0 = 0;

  auto drain = [&]() {
    const int cnt = (int)((qa - q0) / QSTRIDE);
    const int mx = __reduce_max_sync(0xffffffffu, cnt);
    for (int r = 0; r < mx; ++r) {
      // lanes without an r-th entry feed (+inf, INT_MAX), which the network leaves in the carry
      float x = KNN_INF;
      int xi = INT_MAX;
      if (r < cnt) {
        const int l = qj[r][tid];  // position inside the staged chunk (drained before the chunk is replaced)
        x = dist2(sx[l], sy[l], sz[l], qx, qy, qz);  // the pinned arithmetic decides
        xi = PRUNE ? so[l] : c0 + l;
        if (!(x <= tau0)) { x = KNN_INF; xi = INT_MAX; }  // outside the hinted bound: must not count as a pass
      }
      top.template insert<PRUNE>(x, xi);
    }
    qa = q0;
    tau = fminf(tau0, top.tau());
  };

  for (c0 = 0; c0 < m; c0 += KNN_CHUNK) {
    const int cn = min(KNN_CHUNK, m - c0);
    const int cn32 = (cn + 31) & ~31;
    if (PRUNE) {  // whole chunk out of reach of every query of this CTA?  then it is never even staged
      const float* cb = bbc + (size_t)(G0 + c0 / KNN_CHUNK) * 8;
      const float ex = fmaxf(fmaxf(cb[0] - qx, qx - cb[3]), 0.f);
      const float ey = fmaxf(fmaxf(cb[1] - qy, qy - cb[4]), 0.f);
      const float ez = fmaxf(fmaxf(cb[2] - qz, qz - cb[5]), 0.f);
      if (!__syncthreads_or(live && (ex * ex + ey * ey + ez * ez) * 0.9999f <= tau)) continue;
    } else {
      __syncthreads();
    }
    for (int t = tid; t < cn32; t += KNN_THREADS) {
      const bool ok = t < cn;
      sx[t] = ok ? cbase[c0 + t] : KNN_INF;  // +inf padding: d = +inf never passes (ref is stored by position)
      sy[t] = ok ? cbase[m + c0 + t] : 0.f;
      sz[t] = ok ? cbase[2 * m + c0 + t] : 0.f;
      if (PRUNE) so[t] = ok ? pc[c0 + t] : 0;
    }
    if (PRUNE)  // the chunk's level-0 boxes (precomputed by group_bbox_kernel), coalesced
      for (int t = tid; t < (cn32 >> 5) * 8; t += KNN_THREADS) (&sbb[0][0])[t] = bbc[(size_t)(c0 >> 5) * 8 + t];
    __syncthreads();
    if (hint_pending) {
      // tau0 = max exact distance to the hinted candidates and to the point with the query's own index (the
      // self match of a self-query).  Candidates of the staged chunk are read from shared memory (through the
      // inverse permutation when there is one), the rest from global memory.
      hint_pending = false;
      const int32_t* h = hint + ((size_t)cloud * n + qq) * hint_k;  // may alias idx_out: own row, read before write
      auto cand_d = [&](int j) {
        j = min(max(j, 0), m - 1);
        const int pos = pc ? ipc[j] : j;  // position of original index j (hints need iperm_c when permuted)
        return pos < cn ? dist2(sx[pos], sy[pos], sz[pos], qx, qy, qz)
                        : dist2(cbase[pos], cbase[m + pos], cbase[2 * m + pos], qx, qy, qz);
      };
      float mx = cand_d(qq);
      for (int t = 0; t < hint_k; ++t) mx = fmaxf(mx, cand_d(h[t]));
      if (mx < 3.0e38f) tau0 = mx;
      tau = tau0;
    }
    for (int j = 0; j < cn32; j += 32) {
      if (PRUNE) {  // can any candidate of this group be within tau of any query of this warp?
        const float* bb = sbb[j >> 5];
        const float ex = fmaxf(fmaxf(bb[0] - qx, qx - bb[3]), 0.f);
        const float ey = fmaxf(fmaxf(bb[1] - qy, qy - bb[4]), 0.f);
        const float ez = fmaxf(fmaxf(bb[2] - qz, qz - bb[5]), 0.f);
        const float lb = ex * ex + ey * ey + ez * ez;  // d_pin >= lb*(1-1e-6) for every candidate in the box
        if (!__any_sync(0xffffffffu, live && lb * 0.9999f <= tau)) continue;
      }
      unsigned mask = 0u;
      // pass <=> d <= tau <=> d - next(tau) < 0 (sign bit); next(+inf) stays +inf, +inf padding gives +inf / NaN(+)
      const float taun = tau < 3.0e38f ? __uint_as_float(__float_as_uint(tau) + 1u) : KNN_INF;
      const float2 ntau = make_float2(-taun, -taun);
#pragma unroll 2
      for (int u = 0; u < 32; u += 8) {
        const float4 cxa = *reinterpret_cast<const float4*>(sx + j + u), cxb = *reinterpret_cast<const float4*>(sx + j + u + 4);
        const float4 cya = *reinterpret_cast<const float4*>(sy + j + u), cyb = *reinterpret_cast<const float4*>(sy + j + u + 4);
        const float4 cza = *reinterpret_cast<const float4*>(sz + j + u), czb = *reinterpret_cast<const float4*>(sz + j + u + 4);
        const float2 s01 = __fadd2_rn(dist2x2(make_float2(cxa.x, cxa.y), make_float2(cya.x, cya.y), make_float2(cza.x, cza.y), nqx, nqy, nqz), ntau);
        const float2 s23 = __fadd2_rn(dist2x2(make_float2(cxa.z, cxa.w), make_float2(cya.z, cya.w), make_float2(cza.z, cza.w), nqx, nqy, nqz), ntau);
        const float2 s45 = __fadd2_rn(dist2x2(make_float2(cxb.x, cxb.y), make_float2(cyb.x, cyb.y), make_float2(czb.x, czb.y), nqx, nqy, nqz), ntau);
        const float2 s67 = __fadd2_rn(dist2x2(make_float2(cxb.z, cxb.w), make_float2(cyb.z, cyb.w), make_float2(czb.z, czb.w), nqx, nqy, nqz), ntau);
        mask = __funnelshift_l(__float_as_uint(s01.x), mask, 1);
        mask = __funnelshift_l(__float_as_uint(s01.y), mask, 1);
        mask = __funnelshift_l(__float_as_uint(s23.x), mask, 1);
        mask = __funnelshift_l(__float_as_uint(s23.y), mask, 1);
        mask = __funnelshift_l(__float_as_uint(s45.x), mask, 1);
        mask = __funnelshift_l(__float_as_uint(s45.y), mask, 1);
        mask = __funnelshift_l(__float_as_uint(s67.x), mask, 1);
        mask = __funnelshift_l(__float_as_uint(s67.y), mask, 1);
      }
      while (mask) {  // bit 31 is position j, bit 0 is position j+31: ascending position = descending bit
        const int t = __clz(mask);
        mask &= ~(0x80000000u >> t);
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(qa), "h"((unsigned short)(j + t)) : "memory");
        qa += QSTRIDE;
      }
      if (__any_sync(0xffffffffu, qa > q0 + (KNN_QDEPTH - 32) * QSTRIDE)) drain();
    }
    drain();  // queue entries index the staged chunk: empty it before the chunk is replaced
  }
  // fewer than K candidates under the hinted bound => the bound was not valid for this query: redo unhinted
  rescan = __syncthreads_or(live && top.i[K - 1] < 0 && tau0 != KNN_INF);
  tau0 = KNN_INF;
  } while (rescan);

  if (live) {
    int32_t* io = idx_out + ((size_t)cloud * n + qq) * kout;
    float* dn = dist_out ? dist_out + ((size_t)cloud * n + qq) * kout : nullptr;
#pragma unroll
    for (int l = 0; l < K; ++l) {
      const int o = l - drop;
      if (o >= 0 && o < kout) {
        io[o] = top.i[l];
        if (dn) dn[o] = top.d[l];
      }
    }
  }
}

template <int K>
static int launch_knn(const float* query, const float* ref, int b, int n, int m, int kout, int drop,
                      const int32_t* perm_q, const int32_t* perm_c, const int32_t* iperm_c, const float* bb_c,
                      const int32_t* hint, int hint_k, int32_t* idx, float* dist, cudaStream_t s) {
  dim3 grid(ceil_div(n, KNN_THREADS), b, 1);
  if (perm_c)
    knn_kernel<K, true><<<grid, KNN_THREADS, 0, s>>>(query, ref, n, m, kout, drop, perm_q, perm_c, iperm_c, bb_c, hint,
                                                     hint_k, idx, dist);
  else
    knn_kernel<K, false><<<grid, KNN_THREADS, 0, s>>>(query, ref, n, m, kout, drop, perm_q, nullptr, nullptr, nullptr,
                                                      hint, hint_k, idx, dist);
  return GEOA3_LAUNCH_RESULT();
}

// Bounding boxes of an arranged cloud: level 0 = every 32 consecutive positions, level 1 = every KNN_CHUNK.
// bb layout per cloud: [G0 + G1][8] floats = lo xyz, hi xyz, max |p|^2, pad.
// With `perm` the cloud is first ARRANGED (position t takes original point perm[t]) and written to `arranged`:
// one launch replaces the gather + box pass of every pruned search of a step.
__global__ void group_bbox_kernel(const float* __restrict__ pc, const int32_t* __restrict__ perm,
                                  float* __restrict__ arranged, int n, float* __restrict__ bb) {
  const int cloud = blockIdx.y, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int G0 = (n + 31) >> 5, G1 = (n + KNN_CHUNK - 1) / KNN_CHUNK;
  const float* p = pc + (size_t)cloud * 3 * n;
  float* out = bb ? bb + (size_t)cloud * (G0 + G1) * 8 : nullptr;
  const int per = KNN_CHUNK / 32;                 // level-0 groups per level-1 box; one CTA (32 warps) per level-1 box
  const int g1 = blockIdx.x;
  __shared__ float s_box[32][8];
  float lx = 3e38f, ly = 3e38f, lz = 3e38f, hx = -3e38f, hy = -3e38f, hz = -3e38f, w2 = 0.f;
  const int g = g1 * per + wib;
  const int t = g * 32 + lane;
  if (g < G0 && t < n) {
    const int src = perm ? perm[(size_t)cloud * n + t] : t;
    const float x = p[src], y = p[n + src], z = p[2 * n + src];
    if (arranged) {
      float* a = arranged + (size_t)cloud * 3 * n;
      a[t] = x; a[n + t] = y; a[2 * n + t] = z;
    }
    lx = hx = x; ly = hy = y; lz = hz = z; w2 = x * x + y * y + z * z;
  }
  if (!out) return;  // arrangement only (uniform per launch)
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, s)); hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, s));
    ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, s)); hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, s));
    lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, s)); hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, s));
    w2 = fmaxf(w2, __shfl_xor_sync(0xffffffffu, w2, s));
  }
  if (lane == 0) {
    float* o = s_box[wib];
    o[0] = lx; o[1] = ly; o[2] = lz; o[3] = hx; o[4] = hy; o[5] = hz; o[6] = w2; o[7] = 0.f;
    if (g < G0)
      for (int i = 0; i < 8; ++i) out[(size_t)g * 8 + i] = o[i];
  }
  __syncthreads();
  if (wib == 0) {  // level 1: union of this CTA's 32 level-0 boxes
    const float* o = s_box[lane];
    lx = o[0]; ly = o[1]; lz = o[2]; hx = o[3]; hy = o[4]; hz = o[5]; w2 = o[6];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, s)); hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, s));
      ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, s)); hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, s));
      lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, s)); hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, s));
      w2 = fmaxf(w2, __shfl_xor_sync(0xffffffffu, w2, s));
    }
    if (lane == 0) {
      float* o1 = out + (size_t)(G0 + g1) * 8;
      o1[0] = lx; o1[1] = ly; o1[2] = lz; o1[3] = hx; o1[4] = hy; o1[5] = hz; o1[6] = w2; o1[7] = 0.f;
    }
  }
}

}  // namespace geoa3

namespace geoa3 {
int launch_knn_select(const float* query, const float* ref, int b, int n, int m, int K, int kout, int drop,
                      const int32_t* perm_q, const int32_t* perm_c, const int32_t* iperm_c, const float* bb_c,
                      const int32_t* hint, int hint_k, int32_t* idx, float* dist, cudaStream_t s);  // knn_select.cu
}

extern "C" size_t geoa3_group_bbox_floats(int n) {
  return (size_t)(((n + 31) >> 5) + (n + geoa3::KNN_CHUNK - 1) / geoa3::KNN_CHUNK) * 8;
}

extern "C" int geoa3_group_bbox(const float* pc_arranged, int b, int n, float* bb, geoa3_stream_t stream) {
  using namespace geoa3;
  GEOA3_CHECK_ARG(pc_arranged && bb && b > 0 && n > 0);
  if (b > 65535) return GEOA3_EUNSUPPORTED;
  static_assert(KNN_CHUNK == 1024, "one 1024-thread CTA per level-1 box");
  group_bbox_kernel<<<dim3((n + KNN_CHUNK - 1) / KNN_CHUNK, b), 1024, 0, (cudaStream_t)stream>>>(pc_arranged, nullptr,
                                                                                               nullptr, n, bb);
  return GEOA3_LAUNCH_RESULT();
}

extern "C" int geoa3_arrange(const float* pc, const int32_t* perm, int b, int n, float* arranged, float* bb,
                             geoa3_stream_t stream) {
  using namespace geoa3;
  GEOA3_CHECK_ARG(pc && perm && arranged && pc != arranged && b > 0 && n > 0);
  if (b > 65535) return GEOA3_EUNSUPPORTED;
  group_bbox_kernel<<<dim3((n + KNN_CHUNK - 1) / KNN_CHUNK, b), 1024, 0, (cudaStream_t)stream>>>(pc, perm, arranged, n,
                                                                                               bb);
  return GEOA3_LAUNCH_RESULT();
}

extern "C" int geoa3_knn(const float* query, const float* ref, int b, int n, int m, int K, int drop,
                         const int32_t* perm_q, const int32_t* perm_c, const int32_t* iperm_c, const float* bb_c,
                         const int32_t* hint, int hint_k, int32_t* idx, float* dist, geoa3_stream_t stream) {
  using namespace geoa3;
  GEOA3_CHECK_ARG(query && ref && idx);
  GEOA3_CHECK_ARG((perm_c == nullptr) == (bb_c == nullptr));  // an arranged ref comes with its boxes
  GEOA3_CHECK_ARG(b > 0 && n > 0 && m > 0 && K > 0 && drop >= 0 && drop < K && hint_k >= 0);
  if (K > GEOA3_KNN_MAX_K || b > 65535) return GEOA3_EUNSUPPORTED;
  if (K > m) return GEOA3_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  const int kout = K - drop;
  // the list size is a compile-time constant (register arrays); a larger list than requested is still
  // exact: the first K entries of the top-K' (K' >= K) are the top-K.
#define GEOA3_KNN_ARGS query, ref, b, n, m, kout, drop, perm_q, perm_c, iperm_c, bb_c, hint, hint_k, idx, dist, s
  if (K <= 3) return launch_knn<3>(GEOA3_KNN_ARGS);
  if (K <= 5) return launch_knn<5>(GEOA3_KNN_ARGS);
  if (K <= 9) return launch_knn<9>(GEOA3_KNN_ARGS);
  if (K <= 17) return launch_knn<17>(GEOA3_KNN_ARGS);
  return launch_knn<33>(GEOA3_KNN_ARGS);
#undef GEOA3_KNN_ARGS
}

extern "C" int geoa3_knn_set(const float* query, const float* ref, int b, int n, int m, int K, int drop,
                             const int32_t* perm_q, const int32_t* perm_c, const int32_t* iperm_c, const float* bb_c,
                             const int32_t* hint, int hint_k, int32_t* idx, float* dist, geoa3_stream_t stream) {
  using namespace geoa3;
  GEOA3_CHECK_ARG(query && ref && idx);
  GEOA3_CHECK_ARG((perm_c == nullptr) == (bb_c == nullptr));
  GEOA3_CHECK_ARG(b > 0 && n > 0 && m > 0 && K > 0 && drop >= 0 && drop < K && hint_k >= 0);
  if (K > GEOA3_KNN_MAX_K || b > 65535) return GEOA3_EUNSUPPORTED;
  if (K > m) return GEOA3_EINVAL;
  const int r = launch_knn_select(query, ref, b, n, m, K, K - drop, drop, perm_q, perm_c, iperm_c, bb_c, hint, hint_k, idx, dist,
                                  (cudaStream_t)stream);
  if (r != INT_MIN) return r;
  // clouds beyond 65 535 points: the sorted list of geoa3_knn is one valid order of the same members
  return geoa3_knn(query, ref, b, n, m, K, drop, perm_q, perm_c, iperm_c, bb_c, hint, hint_k, idx, dist, stream);
}
