// knn_select.cu — the neighbour SETS of the curvature loss: exact K-nearest-neighbour membership in
// "filter, collect, select" form (K = k+1 <= 33, clouds of up to 65 535 points).  Same members as knn.cu's kernel
// (the K lexicographically smallest (pinned distance, ORIGINAL index) pairs, minus the `drop` smallest), but written
// in ascending VISITING order instead of by distance: kappa and its gradient sum over the neighbourhood, they never
// look at the order, and not sorting is what makes this ~2x cheaper than the register top-K scan per query.
//
//  scan     one query per thread; candidates stream through shared memory as SoA float4 broadcasts.  The per-pair
//           work is a CONSERVATIVE FILTER, not the pinned distance:
//               s = |c|^2 - 2 q.c - (tau - |q|^2 + margin)     3 FFMA2 + 1 FADD2 per two candidates
//           (|c|^2 is staged next to the coordinates), and the sign bit of s is funnel-shifted into a 32-candidate
//           pass mask (1 alu op per pair).  tau = current upper bound of the query's K-th distance;
//           margin = 2^-16 * max(|q|^2, max|c|^2) >= 4x the worst-case rounding of both formulas
//           (DESIGN.md section 5), so every candidate whose PINNED distance is <= tau passes.
//  collect  set bits are expanded into a per-query list in shared memory (positions).
//  select   the list is made exact with the pinned fma chain (entries with d > tau are dropped); while it holds more
//           than K entries the largest (distance, index) key is removed; tau then drops to the K-th distance.  This
//           happens when the list overflows (rare) and once at the end.  The `drop` smallest keys are marked and
//           the members are written in list (= visiting) order.
//
// tau starts at the largest exact distance to the hinted candidates (previous step's neighbours) and to the point
// with the query's own index — K distinct points, hence a valid bound — so in steady state the list ends with
// K..K+3 entries.  Without a hint (or with an invalid one: fewer than K entries survive and the warp searches again)
// tau starts at +inf and tightens as the list overflows.  Any hint yields the exact member set, in the same order.
//
// Spatial pruning (clouds passed ARRANGED in a visiting order together with the group boxes of geoa3_arrange): a
// warp's 32 queries are neighbours in that order; it skips every 128-candidate block and 32-candidate group whose
// box no lane's tau reaches.
#include <climits>

#include "common.cuh"

#ifndef KS_THREADS_17
#define KS_THREADS_17 128
#endif
#ifndef KS_EXTRA
#define KS_EXTRA 11
#endif

namespace geoa3 {

constexpr int KS_MAXCHUNK = 2048;  // candidates staged per pass
constexpr float KS_INF = __builtin_huge_valf();
constexpr float KS_MARGIN = 1.52587890625e-5f * 1.001f;  // 2^-16

template <int K>
struct KsCfg {
  static constexpr int R = K + KS_EXTRA;                      // list slots per query
  static constexpr int THREADS = K <= 17 ? KS_THREADS_17 : 256;  // one query per thread
  static constexpr int MINB = K <= 17 ? 640 / THREADS : 2;     // resident CTAs that shared memory allows at n = 1024
};

__host__ __device__ inline size_t ks_smem_bytes(int cs, int r, int threads, bool prune) {
  return (size_t)cs * 16 + (size_t)((cs >> 5) + (cs >> 7) + 1) * 32 + (size_t)r * threads * 6 + (prune ? (size_t)cs * 4 : 0);
}

struct KsChunk {
  const float *sx, *sy, *sz;
  const uint16_t* so;  // original index of a staged position (nullptr: c0 + position)
  int c0, cn;
};

struct KsList {
  int cnt, nex;  // list length, exact prefix length
  float tau0;    // current upper bound of the K-th distance
};

// List maintenance of one query (one shared-memory column), loop form — the RARE paths only (the list overflows
// in mid-scan: unhinted / stale-hint searches; end of a non-final chunk of a multi-chunk cloud).
// Entries [0, nex) are exact: qd = pinned distance, qv = ORIGINAL index; entries [nex, cnt) are raw: qv = position
// inside the staged chunk.  Makes every entry exact (dropping d > tau0 and padding), removes the largest
// (distance, index) keys until at most kk remain and lowers tau0 to the kk-th distance.  Per thread, no warp-level
// primitives: it may run under divergence.
template <int COLS>
__device__ __forceinline__ KsList ks_compact(float* __restrict__ d_, uint16_t* __restrict__ v_, int cnt, int nex, float tau0,
                                             float qx, float qy, float qz, const KsChunk& ch, int kk) {
  int w = nex;
  for (int s = nex; s < cnt; ++s) {
    const int l = v_[s * COLS];
    const float d = dist2(ch.sx[l], ch.sy[l], ch.sz[l], qx, qy, qz);  // the pinned arithmetic decides
    if (l < ch.cn && d <= tau0) {
      d_[w * COLS] = d;
      v_[w * COLS] = (uint16_t)(ch.so ? ch.so[l] : ch.c0 + l);
      ++w;
    }
  }
  while (w > kk) {  // remove the largest (distance, original index) key: one pass, ties on the distance are rare
    float md = d_[0];
    int bs = 0;
    for (int s = 1; s < w; ++s) {
      const float d = d_[s * COLS];
      if (d >= md) {
        if (d > md || v_[s * COLS] > v_[bs * COLS]) { md = d; bs = s; }
      }
    }
    --w;
    for (int s = bs; s < w; ++s) {  // close the gap: the survivors stay in arrival (= visiting) order
      d_[s * COLS] = d_[(s + 1) * COLS];
      v_[s * COLS] = v_[(s + 1) * COLS];
    }
  }
  if (w == kk) {
    float md = d_[0];
    for (int s = 1; s < w; ++s) md = fmaxf(md, d_[s * COLS]);
    tau0 = fminf(tau0, md);
  }
  KsList r;
  r.cnt = r.nex = w;
  r.tau0 = tau0;
  return r;
}

template <int K, int T, bool PRUNE>
__global__ void __launch_bounds__(T, KsCfg<K>::MINB)
knn_select_kernel(const float* __restrict__ query, const float* __restrict__ ref, int n, int m, int kout, int drop,
                  const int32_t* __restrict__ perm_q, const int32_t* __restrict__ perm_c,
                  const int32_t* __restrict__ iperm_c, const float* __restrict__ bb_c, const int32_t* hint, int hint_k,
                  int32_t* idx_out, float* __restrict__ dist_out, int cs, int kk) {
  // kk = min(requested K, m): the list target (the template K is the capacity class; a cloud may hold fewer points)
  constexpr int R = KsCfg<K>::R;
  extern __shared__ __align__(16) unsigned char ks_smem[];
  float* sx = reinterpret_cast<float*>(ks_smem);
  float* sy = sx + cs;
  float* sz = sy + cs;
  float* sw = sz + cs;                    // |c|^2
  float* sbb = sw + cs;                   // PRUNE: [cs/32][8] group boxes (lo xyz, hi xyz, max |c|^2, pad)
  float* sb2 = sbb + (cs >> 5) * 8;       // PRUNE: [cs/128][8] boxes of 4 groups; else [0] = max |c|^2 of the chunk
  float* qd = sb2 + ((cs >> 7) + 1) * 8;  // [R][T] exact distances
  uint16_t* qv = reinterpret_cast<uint16_t*>(qd + R * T);  // [R][T] position (raw) / original index (exact)
  uint16_t* so = qv + R * T;              // PRUNE: [cs] original index of the staged candidate
  uint16_t* sip = so + cs;                // PRUNE, single chunk: [cs] position of an original index (hint look-up)

  const int cloud = blockIdx.y, tid = threadIdx.x;
  const float* qbase = query + (size_t)cloud * 3 * n;
  const float* cbase = ref + (size_t)cloud * 3 * m;
  const int32_t* pc = PRUNE ? perm_c + (size_t)cloud * m : nullptr;
  const int32_t* ipc = (PRUNE && iperm_c) ? iperm_c + (size_t)cloud * m : nullptr;
  const int G0 = (m + 31) >> 5;
  const float* bbc = PRUNE ? bb_c + (size_t)cloud * (G0 + (m + 1023) / 1024) * 8 : nullptr;  // layout of geoa3_arrange
  const bool one_chunk = m <= cs;

  const int slot = blockIdx.x * T + tid;  // position in visiting order: a warp's 32 queries are neighbours
  const int sl = min(slot, n - 1);
  const int qo = perm_q ? perm_q[(size_t)cloud * n + sl] : sl;  // ORIGINAL index of the query
  const bool live = slot < n;
  bool act = live;                        // still searching
  const bool use_hint = hint != nullptr && hint_k + 1 >= kk && (!PRUNE || ipc);
  const bool hint16 = use_hint && one_chunk && hint_k == 16 && ((reinterpret_cast<uintptr_t>(hint) & 15) == 0);
  int4 hreg[4];                           // hint row fetched up front: its DRAM/L2 latency hides behind the staging
  if (hint16) {
    const int4* h4 = reinterpret_cast<const int4*>(hint + ((size_t)cloud * n + qo) * 16);
#pragma unroll
    for (int t = 0; t < 4; ++t) hreg[t] = h4[t];
  }
  const float qx = qbase[sl], qy = qbase[n + sl], qz = qbase[2 * n + sl];
  const float2 ax = make_float2(-2.f * qx, -2.f * qx), ay = make_float2(-2.f * qy, -2.f * qy),
               az = make_float2(-2.f * qz, -2.f * qz);
  const float qq2 = qx * qx + qy * qy + qz * qz;
  float tau0 = KS_INF;
  int cnt = 0, nex = 0;
  float* d_ = qd + tid;
  uint16_t* v_ = qv + tid;
  bool first_pass = true;
  KsChunk ch;
  ch.sx = sx; ch.sy = sy; ch.sz = sz; ch.so = PRUNE ? so : nullptr; ch.c0 = 0; ch.cn = 0;

  bool rescan;
  do {
    for (int c0 = 0; c0 < m; c0 += cs) {
      const int cn = min(cs, m - c0);
      const int cn32 = (cn + 31) & ~31;
      ch.c0 = c0; ch.cn = cn;
      if (first_pass || !one_chunk) {  // (a rescan of a single-chunk cloud finds everything still staged)
        __syncthreads();  // the previous chunk (and every raw entry pointing into it) is done with
        float wmax = 0.f;
        for (int t = tid; t < cn32; t += T) {
          const bool ok = t < cn;
          // padding candidates sit far away (finite, so the filter never sees inf - inf) and fail the `l < cn` test
          const float x = ok ? cbase[c0 + t] : 1e18f, y = ok ? cbase[m + c0 + t] : 0.f, z = ok ? cbase[2 * m + c0 + t] : 0.f;
          int o = 0, ip = 0;
          if (PRUNE && ok) {
            o = pc[c0 + t];
            if (one_chunk && ipc) ip = ipc[t];
          }
          const float w2 = x * x + y * y + z * z;
          sx[t] = x; sy[t] = y; sz[t] = z; sw[t] = w2;
          if (PRUNE) {
            so[t] = (uint16_t)o;
            if (one_chunk) sip[t] = (uint16_t)ip;
          } else if (ok) {
            wmax = fmaxf(wmax, w2);
          }
        }
        if (PRUNE) {  // the chunk's group boxes, precomputed by geoa3_arrange (coalesced copy)
          for (int t = tid; t < (cn32 >> 5) * 8; t += T) sbb[t] = bbc[(size_t)(c0 >> 5) * 8 + t];
        } else {      // one bound for the whole chunk: max |c|^2 (float bits of non-negative values order like uints)
          if (tid == 0) sb2[0] = 0.f;
          __syncthreads();
          const unsigned wm = __reduce_max_sync(0xffffffffu, __float_as_uint(wmax));
          if ((tid & 31) == 0) atomicMax(reinterpret_cast<unsigned*>(sb2), wm);
        }
        __syncthreads();
        if (PRUNE) {  // boxes of 4 consecutive groups (128 candidates): the first level of the pruning test
          const int g4 = (cn32 + 127) >> 7;
          if (tid < g4) {
            float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f}, w2m = 0.f;
            for (int g = tid * 4; g < min(tid * 4 + 4, cn32 >> 5); ++g) {
              const float* b = sbb + g * 8;
#pragma unroll
              for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], b[a]); hi[a] = fmaxf(hi[a], b[3 + a]); }
              w2m = fmaxf(w2m, b[6]);
            }
            float* o2 = sb2 + tid * 8;
            o2[0] = lo[0]; o2[1] = lo[1]; o2[2] = lo[2]; o2[3] = hi[0]; o2[4] = hi[1]; o2[5] = hi[2]; o2[6] = w2m; o2[7] = 0.f;
          }
          __syncthreads();
        }
      }

      if (first_pass && c0 == 0 && use_hint && act) {
        // tau0 = max exact distance to the hinted candidates and to the point with the query's own index: an upper
        // bound of the K-th distance whenever these are K distinct points (verified at the end: >= K survivors).
        float mx;
        if (hint16) {  // every candidate is staged: shared-memory look-ups only, the row is already in registers
          auto cand_s = [&](int j) {
            j = min(max(j, 0), m - 1);
            const int pos = PRUNE ? (int)sip[j] : j;
            return dist2(sx[pos], sy[pos], sz[pos], qx, qy, qz);
          };
          mx = cand_s(qo);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            mx = fmaxf(fmaxf(mx, cand_s(hreg[t].x)), cand_s(hreg[t].y));
            mx = fmaxf(fmaxf(mx, cand_s(hreg[t].z)), cand_s(hreg[t].w));
          }
        } else {
          auto cand_d = [&](int j) {
            j = min(max(j, 0), m - 1);
            const int pos = PRUNE ? ipc[j] : j;
            return pos < cn ? dist2(sx[pos], sy[pos], sz[pos], qx, qy, qz)
                            : dist2(cbase[pos], cbase[m + pos], cbase[2 * m + pos], qx, qy, qz);
          };
          const int32_t* h = hint + ((size_t)cloud * n + qo) * hint_k;  // may alias idx_out: own row, read before write
          mx = cand_d(qo);
#pragma unroll 1
          for (int t = 0; t < hint_k; ++t) mx = fmaxf(mx, cand_d(h[t]));
        }
        if (mx < 3.0e38f) tau0 = mx;
      }

      for (int j4 = 0; j4 < cn32; j4 += 128) {
        if (PRUNE) {  // can any candidate of these 4 groups be within tau of any query of this warp?
          const float* b2 = sb2 + (j4 >> 7) * 8;
          const float ex = fmaxf(fmaxf(b2[0] - qx, qx - b2[3]), 0.f);
          const float ey = fmaxf(fmaxf(b2[1] - qy, qy - b2[4]), 0.f);
          const float ez = fmaxf(fmaxf(b2[2] - qz, qz - b2[5]), 0.f);
          if (!__any_sync(0xffffffffu, act && (ex * ex + ey * ey + ez * ez) * 0.9999f <= tau0)) continue;
        }
        for (int j = j4; j < min(j4 + 128, cn32); j += 32) {
          const float* bb = sbb + (j >> 5) * 8;
          if (PRUNE) {
            const float ex = fmaxf(fmaxf(bb[0] - qx, qx - bb[3]), 0.f);
            const float ey = fmaxf(fmaxf(bb[1] - qy, qy - bb[4]), 0.f);
            const float ez = fmaxf(fmaxf(bb[2] - qz, qz - bb[5]), 0.f);
            const float lb = ex * ex + ey * ey + ez * ez;  // d_pin >= lb*(1-1e-6) for every candidate in the box
            if (!__any_sync(0xffffffffu, act && lb * 0.9999f <= tau0)) continue;
          }
          // pass <=> t - thr < 0 with thr = tau0 - |q|^2 + margin; finished / empty slots get thr = -inf
          const float w2max = PRUNE ? bb[6] : sb2[0];
          const float thr = act ? (tau0 - qq2) + KS_MARGIN * fmaxf(qq2, w2max) : -KS_INF;
          const float2 nthr = make_float2(-thr, -thr);
          unsigned mk = 0u;
#pragma unroll
          for (int u = 0; u < 32; u += 8) {
            const float4 cxa = *reinterpret_cast<const float4*>(sx + j + u), cxb = *reinterpret_cast<const float4*>(sx + j + u + 4);
            const float4 cya = *reinterpret_cast<const float4*>(sy + j + u), cyb = *reinterpret_cast<const float4*>(sy + j + u + 4);
            const float4 cza = *reinterpret_cast<const float4*>(sz + j + u), czb = *reinterpret_cast<const float4*>(sz + j + u + 4);
            const float4 cwa = *reinterpret_cast<const float4*>(sw + j + u), cwb = *reinterpret_cast<const float4*>(sw + j + u + 4);
            float2 s01 = __ffma2_rn(ax, make_float2(cxa.x, cxa.y), nthr);
            float2 s23 = __ffma2_rn(ax, make_float2(cxa.z, cxa.w), nthr);
            float2 s45 = __ffma2_rn(ax, make_float2(cxb.x, cxb.y), nthr);
            float2 s67 = __ffma2_rn(ax, make_float2(cxb.z, cxb.w), nthr);
            s01 = __ffma2_rn(ay, make_float2(cya.x, cya.y), s01);
            s23 = __ffma2_rn(ay, make_float2(cya.z, cya.w), s23);
            s45 = __ffma2_rn(ay, make_float2(cyb.x, cyb.y), s45);
            s67 = __ffma2_rn(ay, make_float2(cyb.z, cyb.w), s67);
            s01 = __ffma2_rn(az, make_float2(cza.x, cza.y), s01);
            s23 = __ffma2_rn(az, make_float2(cza.z, cza.w), s23);
            s45 = __ffma2_rn(az, make_float2(czb.x, czb.y), s45);
            s67 = __ffma2_rn(az, make_float2(czb.z, czb.w), s67);
            s01 = __fadd2_rn(s01, make_float2(cwa.x, cwa.y));
            s23 = __fadd2_rn(s23, make_float2(cwa.z, cwa.w));
            s45 = __fadd2_rn(s45, make_float2(cwb.x, cwb.y));
            s67 = __fadd2_rn(s67, make_float2(cwb.z, cwb.w));
            mk = __funnelshift_l(__float_as_uint(s01.x), mk, 1);
            mk = __funnelshift_l(__float_as_uint(s01.y), mk, 1);
            mk = __funnelshift_l(__float_as_uint(s23.x), mk, 1);
            mk = __funnelshift_l(__float_as_uint(s23.y), mk, 1);
            mk = __funnelshift_l(__float_as_uint(s45.x), mk, 1);
            mk = __funnelshift_l(__float_as_uint(s45.y), mk, 1);
            mk = __funnelshift_l(__float_as_uint(s67.x), mk, 1);
            mk = __funnelshift_l(__float_as_uint(s67.y), mk, 1);
          }
          while (mk) {  // bit 31 is position j, bit 0 is position j+31: ascending position = descending bit
            const int t = __clz(mk);
            mk &= ~(0x80000000u >> t);
            if (cnt == R) {  // rare: makes room (<= kk entries remain) and tightens tau0
              const KsList r = ks_compact<T>(d_, v_, cnt, nex, tau0, qx, qy, qz, ch, kk);
              cnt = r.cnt; nex = r.nex; tau0 = r.tau0;
            }
            v_[cnt * T] = (uint16_t)(j + t);
            ++cnt;
          }
        }
      }
      // raw entries index the staged chunk: before it is replaced they have to become exact
      if (act && c0 + cs < m) {
        const KsList r = ks_compact<T>(d_, v_, cnt, nex, tau0, qx, qy, qz, ch, kk);
        cnt = r.cnt; nex = r.nex; tau0 = r.tau0;
      }
    }

    // -- select: make the list exact (the last chunk is still staged) and cut it down to the kk smallest keys
    if (act) {
      const KsList r = ks_compact<T>(d_, v_, cnt, nex, tau0, qx, qy, qz, ch, kk);
      cnt = r.cnt; nex = r.nex; tau0 = r.tau0;
    }
    // fewer than K survivors => the hinted bound was not valid for this query: search again from tau = +inf
    const bool fail = act && cnt < kk;
    rescan = one_chunk ? __any_sync(0xffffffffu, fail) : __syncthreads_or(fail);
    if (!fail && act) {
      // drop the `drop` smallest (distance, index) keys (the self match of a self-query), then write the members
      // in list order = ascending visiting position (ks_compact keeps the survivors in arrival order)
      for (int t = 0; t < drop; ++t) {
        float md = KS_INF;
        int bs = 0, bo = INT_MAX;
        for (int s2 = 0; s2 < cnt; ++s2) {
          const float d = d_[s2 * T];
          if (d >= 0.f && d <= md) {
            const int o = v_[s2 * T];
            if (d < md || o < bo) { md = d; bs = s2; bo = o; }
          }
        }
        d_[bs * T] = -1.f;  // marked: no longer a member
      }
      int32_t* io = idx_out + ((size_t)cloud * n + qo) * kout;
      float* dn = dist_out ? dist_out + ((size_t)cloud * n + qo) * kout : nullptr;
      int o = 0;
      for (int s2 = 0; s2 < cnt && o < kout; ++s2) {
        const float d = d_[s2 * T];
        if (d >= 0.f) {
          io[o] = (int32_t)v_[s2 * T];
          if (dn) dn[o] = d;
          ++o;
        }
      }
      act = false;  // finished: takes no further part if the warp / CTA has to rescan
    }
    if (fail) { tau0 = KS_INF; cnt = nex = 0; }
    first_pass = false;
  } while (rescan);
}

template <int K>
static int launch_knn_select_k(const float* query, const float* ref, int b, int n, int m, int kreq, int kout, int drop,
                               const int32_t* perm_q, const int32_t* perm_c, const int32_t* iperm_c, const float* bb_c,
                               const int32_t* hint, int hint_k, int32_t* idx, float* dist, cudaStream_t s) {
  constexpr int T = KsCfg<K>::THREADS;
  const int cs = min((m + 31) & ~31, KS_MAXCHUNK);
  const bool prune = perm_c != nullptr;
  const size_t smem = ks_smem_bytes(cs, KsCfg<K>::R, T, prune);
  dim3 grid(ceil_div(n, T), b, 1);
  cudaError_t e = cudaSuccess;
  if (prune) {
    static PerDeviceOnce once;
    if (once.needed()) {
      e = cudaFuncSetAttribute(knn_select_kernel<K, T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return (int)e;
      once.done();
    }
    knn_select_kernel<K, T, true><<<grid, T, smem, s>>>(query, ref, n, m, kout, drop, perm_q, perm_c, iperm_c, bb_c, hint,
                                                      hint_k, idx, dist, cs, min(kreq, m));
  } else {
    static PerDeviceOnce once;
    if (once.needed()) {
      e = cudaFuncSetAttribute(knn_select_kernel<K, T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return (int)e;
      once.done();
    }
    knn_select_kernel<K, T, false><<<grid, T, smem, s>>>(query, ref, n, m, kout, drop, perm_q, nullptr, nullptr, nullptr,
                                                       hint, hint_k, idx, dist, cs, min(kreq, m));
  }
  return GEOA3_LAUNCH_RESULT();
}

// Dispatch for geoa3_knn_set (knn.cu).  Returns INT_MIN when this kernel does not apply (clouds beyond 65 535 points).
int launch_knn_select(const float* query, const float* ref, int b, int n, int m, int K, int kout, int drop,
                      const int32_t* perm_q, const int32_t* perm_c, const int32_t* iperm_c, const float* bb_c,
                      const int32_t* hint, int hint_k, int32_t* idx, float* dist, cudaStream_t s) {
  if (n > 65535 || m > 65535) return INT_MIN;
#define GEOA3_KS_ARGS query, ref, b, n, m, K, kout, drop, perm_q, perm_c, iperm_c, bb_c, hint, hint_k, idx, dist, s
  if (K <= 3) return launch_knn_select_k<3>(GEOA3_KS_ARGS);
  if (K <= 5) return launch_knn_select_k<5>(GEOA3_KS_ARGS);
  if (K <= 9) return launch_knn_select_k<9>(GEOA3_KS_ARGS);
  if (K <= 17) return launch_knn_select_k<17>(GEOA3_KS_ARGS);
  return launch_knn_select_k<33>(GEOA3_KS_ARGS);
#undef GEOA3_KS_ARGS
}

}  // namespace geoa3
