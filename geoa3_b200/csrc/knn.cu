// knn.cu — exact K-nearest-neighbour selection (K = k+1 <= 33) for the curvature losses.
//
// Thread-per-query brute force, candidates streamed through shared memory as SoA float4 broadcasts
// and evaluated two at a time on the packed fp32 pipe (pinned fma-chain arithmetic, see common.cuh).
// Each thread keeps its K best (dist, idx) pairs sorted in registers; tau = current K-th distance.
// A candidate passes only if d < tau (strict; candidates are visited in ascending index, so among
// equal distances the lowest indices survive => lexicographic (dist, idx) order, the pinned rule).
// Passing candidates are NOT inserted on the spot (that would serialise the warp on its slowest
// lane every step): they are appended to a small per-thread queue in shared memory and the warp
// drains all queues together — a branch-free sorted-insert network executed by all 32 lanes — when
// any lane's queue is nearly full.  The queue preserves arrival order, so ties keep their order.
//
// Hinted threshold.  Scanning in index order from tau = +inf inserts ~K*ln(N/K) candidates per query, and the
// insert network (5 alu ops per list slot) is what the half-width alu pipe chokes on.  If the caller passes,
// per query, a list of candidate indices (typically the neighbours found one optimisation step earlier),
// tau starts at the largest exact distance to those candidates — an upper bound of the true K-th distance
// whenever the list (plus the query's own index) holds K distinct points — so only ~K candidates are ever
// inserted.  The bound is self-verifying: if fewer than K candidates passed it the CTA rescans with
// tau = +inf.  Any hint therefore yields the exact lexicographic top-K; a bad one only costs time.
#include "common.cuh"

namespace geoa3 {

constexpr int KNN_THREADS = 128;
constexpr int KNN_CHUNK = 1024;  // candidates per shared-memory pass (12 KB; static smem stays under 48 KB)
constexpr int KNN_QDEPTH = 56;   // queue slots per thread (indices only; distances are recomputed at drain):
                                 // drained when > 24 are pending, and one 32-candidate group adds at most 32

template <int K>
struct TopK {
  float d[K];
  int i[K];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int l = 0; l < K; ++l) { d[l] = __int_as_float(0x7f800000); i[l] = -1; }
  }
  // sorted insert of (x, xi); elements equal to x stay in front of it (arrival order = index order)
  __device__ __forceinline__ void insert(float x, int xi) {
    float cd = x;
    int ci = xi;
#pragma unroll
    for (int l = 0; l < K; ++l) {
      const bool p = x < d[l];
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

template <int K>
__global__ void __launch_bounds__(KNN_THREADS, K <= 17 ? 6 : 4)
knn_kernel(const float* __restrict__ query, const float* __restrict__ ref, int n, int m, int kout, int drop,
           const int32_t* hint, int hint_k, int32_t* idx_out, float* __restrict__ dist_out) {
  __shared__ __align__(16) float sx[KNN_CHUNK];
  __shared__ __align__(16) float sy[KNN_CHUNK];
  __shared__ __align__(16) float sz[KNN_CHUNK];
  __shared__ int qj[KNN_QDEPTH][KNN_THREADS];

  const int cloud = blockIdx.y;
  const int tid = threadIdx.x;
  const int qi = blockIdx.x * KNN_THREADS + tid;
  const float* qbase = query + (size_t)cloud * 3 * n;
  const float* cbase = ref + (size_t)cloud * 3 * m;
  const int qq = min(qi, n - 1);
  const float qx = qbase[qq], qy = qbase[n + qq], qz = qbase[2 * n + qq];
  const float2 nqx = make_float2(-qx, -qx), nqy = make_float2(-qy, -qy), nqz = make_float2(-qz, -qz);

  // hinted start threshold (evaluated below, once the first candidate chunk sits in shared memory)
  float tau0 = __int_as_float(0x7f800000);
  const bool use_hint = hint != nullptr && hint_k + 1 >= K;

  TopK<K> top;
  bool rescan = false;
  do {
  top.init();
  float tau = tau0;
  bool hint_pending = use_hint && !rescan;
  // per-thread queue of passing candidate indices: qp walks down column `tid` of qj; the hot loop only
  // does "compare, predicated store, predicated pointer bump" per candidate
  // 32-bit shared-window address of this thread's queue column, kept in ONE register: the compiler otherwise
  // re-materialises the base (3 extra predicated instructions per candidate)
  const unsigned q0 = (unsigned)__cvta_generic_to_shared(&qj[0][tid]);
  constexpr unsigned QSTRIDE = KNN_THREADS * sizeof(int);
  unsigned qa = q0;
  int c0 = 0;

  auto drain = [&]() {
    const int cnt = (int)((qa - q0) / QSTRIDE);
    const int mx = __reduce_max_sync(0xffffffffu, cnt);
    for (int r = 0; r < mx; ++r) {
      // lanes without an r-th entry feed +inf, which the network leaves in the carry
      float x = __int_as_float(0x7f800000);
      int xi = -1;
      if (r < cnt) {
        xi = qj[r][tid];
        const int l = xi - c0;  // the queue is drained before the staged chunk is replaced
        x = dist2(sx[l], sy[l], sz[l], qx, qy, qz);  // bit-identical to the packed evaluation
      }
      top.insert(x, xi);
    }
    qa = q0;
    tau = fminf(tau0, top.tau());
  };

  for (c0 = 0; c0 < m; c0 += KNN_CHUNK) {
    const int cn = min(KNN_CHUNK, m - c0);
    const int cn32 = (cn + 31) & ~31;
    __syncthreads();
    for (int t = tid; t < cn32; t += KNN_THREADS) {
      const bool ok = t < cn;
      sx[t] = ok ? cbase[c0 + t] : __int_as_float(0x7f800000);  // +inf padding: d = +inf never passes
      sy[t] = ok ? cbase[m + c0 + t] : 0.f;
      sz[t] = ok ? cbase[2 * m + c0 + t] : 0.f;
    }
    __syncthreads();
    if (hint_pending) {
      // tau0 = max exact distance to the hinted candidates and to the point of the same index (the self match
      // of a self-query), nudged one ulp up so that the strict test admits equality.  Candidates of the staged
      // chunk are read from shared memory, the rest (m > KNN_CHUNK only) from global memory.
      hint_pending = false;
      const int32_t* h = hint + ((size_t)cloud * n + qq) * hint_k;  // may alias idx_out: own row, read before write
      auto cand_d = [&](int j) {
        j = min(max(j, 0), m - 1);
        return j < cn ? dist2(sx[j], sy[j], sz[j], qx, qy, qz)
                      : dist2(cbase[j], cbase[m + j], cbase[2 * m + j], qx, qy, qz);
      };
      float mx = cand_d(qq);
      for (int t = 0; t < hint_k; ++t) mx = fmaxf(mx, cand_d(h[t]));
      if (mx < 3.0e38f) tau0 = __uint_as_float(__float_as_uint(mx) + 1u);
      tau = tau0;
    }
    // The alu pipe (compares, selects, integer adds) is half as wide as the fma pipe, so the per-candidate
    // bookkeeping is reduced to ONE alu op: s = d - tau is formed on the fma pipe (its sign is exact) and a
    // funnel shift pushes the sign bit into a 32-candidate pass mask.  Set bits are expanded into the index
    // queue once per 32 candidates (a short divergent loop, ~2% of the candidates pass).
    for (int j = 0; j < cn32; j += 32) {
      unsigned mask = 0u;
      const float2 ntau = make_float2(-tau, -tau);
#pragma unroll 2
      for (int u = 0; u < 32; u += 8) {
        const float4 cxa = *reinterpret_cast<const float4*>(sx + j + u), cxb = *reinterpret_cast<const float4*>(sx + j + u + 4);
        const float4 cya = *reinterpret_cast<const float4*>(sy + j + u), cyb = *reinterpret_cast<const float4*>(sy + j + u + 4);
        const float4 cza = *reinterpret_cast<const float4*>(sz + j + u), czb = *reinterpret_cast<const float4*>(sz + j + u + 4);
        const float2 s01 = __fadd2_rn(dist2x2(make_float2(cxa.x, cxa.y), make_float2(cya.x, cya.y), make_float2(cza.x, cza.y), nqx, nqy, nqz), ntau);
        const float2 s23 = __fadd2_rn(dist2x2(make_float2(cxa.z, cxa.w), make_float2(cya.z, cya.w), make_float2(cza.z, cza.w), nqx, nqy, nqz), ntau);
        const float2 s45 = __fadd2_rn(dist2x2(make_float2(cxb.x, cxb.y), make_float2(cyb.x, cyb.y), make_float2(czb.x, czb.y), nqx, nqy, nqz), ntau);
        const float2 s67 = __fadd2_rn(dist2x2(make_float2(cxb.z, cxb.w), make_float2(cyb.z, cyb.w), make_float2(czb.z, czb.w), nqx, nqy, nqz), ntau);
        // d < tau  <=>  sign(d - tau) set  (d = +inf padding gives +inf or the positive canonical NaN)
        mask = __funnelshift_l(__float_as_uint(s01.x), mask, 1);
        mask = __funnelshift_l(__float_as_uint(s01.y), mask, 1);
        mask = __funnelshift_l(__float_as_uint(s23.x), mask, 1);
        mask = __funnelshift_l(__float_as_uint(s23.y), mask, 1);
        mask = __funnelshift_l(__float_as_uint(s45.x), mask, 1);
        mask = __funnelshift_l(__float_as_uint(s45.y), mask, 1);
        mask = __funnelshift_l(__float_as_uint(s67.x), mask, 1);
        mask = __funnelshift_l(__float_as_uint(s67.y), mask, 1);
      }
      const int jj = c0 + j;
      while (mask) {  // bit 31 is candidate jj, bit 0 is candidate jj+31: ascending index = descending bit
        const int t = __clz(mask);
        mask &= ~(0x80000000u >> t);
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(qa), "r"(jj + t) : "memory");
        qa += QSTRIDE;
      }
      if (__any_sync(0xffffffffu, qa > q0 + (KNN_QDEPTH - 32) * QSTRIDE)) drain();
    }
    drain();  // queue entries index the staged chunk: empty it before the chunk is replaced
  }
  // fewer than K candidates under the hinted bound => the bound was not valid for this query: redo unhinted
  rescan = __syncthreads_or((qi < n) && top.i[K - 1] < 0 && tau0 != __int_as_float(0x7f800000));
  tau0 = __int_as_float(0x7f800000);
  } while (rescan);

  if (qi < n) {
    int32_t* io = idx_out + ((size_t)cloud * n + qi) * kout;
    float* dn = dist_out ? dist_out + ((size_t)cloud * n + qi) * kout : nullptr;
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
                      const int32_t* hint, int hint_k, int32_t* idx, float* dist, cudaStream_t s) {
  dim3 grid(ceil_div(n, KNN_THREADS), b, 1);
  knn_kernel<K><<<grid, KNN_THREADS, 0, s>>>(query, ref, n, m, kout, drop, hint, hint_k, idx, dist);
  return GEOA3_LAUNCH_RESULT();
}

}  // namespace geoa3

extern "C" int geoa3_knn(const float* query, const float* ref, int b, int n, int m, int K, int drop,
                         const int32_t* hint, int hint_k, int32_t* idx, float* dist, geoa3_stream_t stream) {
  using namespace geoa3;
  GEOA3_CHECK_ARG(query && ref && idx);
  GEOA3_CHECK_ARG(b > 0 && n > 0 && m > 0 && K > 0 && drop >= 0 && drop < K && hint_k >= 0);
  if (K > GEOA3_KNN_MAX_K || b > 65535) return GEOA3_EUNSUPPORTED;
  if (K > m) return GEOA3_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  const int kout = K - drop;
  // the list size is a compile-time constant (register arrays); a larger list than requested is still
  // exact: the first K entries of the top-K' (K' >= K) are the top-K.
  if (K <= 3) return launch_knn<3>(query, ref, b, n, m, kout, drop, hint, hint_k, idx, dist, s);
  if (K <= 5) return launch_knn<5>(query, ref, b, n, m, kout, drop, hint, hint_k, idx, dist, s);
  if (K <= 9) return launch_knn<9>(query, ref, b, n, m, kout, drop, hint, hint_k, idx, dist, s);
  if (K <= 17) return launch_knn<17>(query, ref, b, n, m, kout, drop, hint, hint_k, idx, dist, s);
  return launch_knn<33>(query, ref, b, n, m, kout, drop, hint, hint_k, idx, dist, s);
}
