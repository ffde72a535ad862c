// nn_pair.cu — fused bidirectional 1-NN (Chamfer / Hausdorff / normal borrowing front end).
//
// One launch covers both directions: each CTA owns a tile of queries of one cloud and streams the
// cloud's candidates through shared memory (SoA x[],y[],z[]: one LDS.128 feeds four candidates to every
// lane as a broadcast).  Each thread keeps Q queries in registers and evaluates two candidates per
// instruction on the packed fp32 pipe (FADD2/FMUL2/FFMA2).  The N x M matrix only ever exists as registers.
//
// Seeded exact search.  On B200 the alu pipe (FSETP/FSEL/SEL/FMNMX, 16 lanes/clk/SMSP) is half as wide
// as the fma pipe, so a per-pair compare+select+index update (3 alu ops) costs as much as the distance
// itself.  Instead every query starts from a seed candidate (caller hint, default: the point with the
// same index — in the attack adv_i is a perturbed ori_i) whose exact distance initialises `best`; the
// hot loop then only folds 8 distances into one FMNMX tree and asks "min <= best?" (1 alu op per pair).
// Only when that fires does the thread walk the 8 candidates with the exact lexicographic rule
// (d < best) or (d == best and j < argbest), so the result is the same (min, lowest index) the plain
// ascending scan produces — for ANY seed; a bad seed only costs speed.
//
// Conservative filter.  The question "can any of these 8 candidates reach best?" does not need the pinned
// arithmetic, only a value that never under-reports by more than a known bound.  The hot loop therefore
// evaluates t = |c|^2 - 2 q.c (three FFMA2 per two candidates instead of six packed ops; |c|^2 is staged
// next to the coordinates) and tests  min t <= best - |q|^2 + margin,  margin = 2^-17 * max(|q|^2, max|c|^2)
// (~4x the worst-case sum of the rounding errors of both formulas).  Everything that
// passes is re-evaluated with the pinned fma chain before it may touch (best, argbest), so the outputs are
// bit-identical to the plain scan; the filter only decides what is worth evaluating exactly.
//
// Spatial pruning.  The caller may pass the clouds ALREADY ARRANGED in a visiting order (position t holds original
// point perm[t]; e.g. the Morton order of the original cloud, computed once per attack, applied to both clouds —
// staging stays coalesced that way).  A warp's 32*Q queries and every group of 32 staged candidates are then
// spatially compact; each group carries its bounding box and a warp skips the whole group when no lane's box
// distance can reach its current best (strictly, with a relative slack far above the rounding of the box test).
// Indices are reported in ORIGINAL numbering (through perm) and ties are resolved on original indices, so the
// result does not depend on the order — any permutation is valid.
#include "common.cuh"

namespace geoa3 {

constexpr int NN_THREADS = 256;
constexpr int NN_Q = 2;          // queries per thread
constexpr int NN_CHUNK = 2048;   // candidates staged per shared-memory pass
constexpr int NN_GROUPS = NN_CHUNK / 32;

template <int Q, bool PRUNE>
__global__ void __launch_bounds__(NN_THREADS)
nn_pair_kernel(const float* __restrict__ adv, const float* __restrict__ ori, int n, int m,
               const int32_t* __restrict__ perm_a, const int32_t* __restrict__ perm_o,
               const int32_t* __restrict__ iperm_a, const int32_t* __restrict__ iperm_o,
               const int32_t* hint_a2o, const int32_t* hint_o2a,
               float* __restrict__ d_a2o, int32_t* jstar, float* __restrict__ d_o2a, int32_t* istar, int tiles_a) {
  __shared__ __align__(16) float sx[NN_CHUNK];
  __shared__ __align__(16) float sy[NN_CHUNK];
  __shared__ __align__(16) float sz[NN_CHUNK];
  __shared__ __align__(16) float sw[NN_CHUNK];  // |c|^2
  __shared__ int so[NN_CHUNK];                  // original index of the staged candidate
  __shared__ float sbb[NN_GROUPS][8];           // per 32-candidate group: box lo xyz, hi xyz, max |c|^2

  const int cloud = blockIdx.y;
  int tile = blockIdx.x;
  // direction 0: queries = adv (n), candidates = ori (m); direction 1: swapped.
  const bool dir1 = tile >= tiles_a;
  if (dir1) tile -= tiles_a;
  const float* qbase = dir1 ? ori + (size_t)cloud * 3 * m : adv + (size_t)cloud * 3 * n;
  const float* cbase = dir1 ? adv + (size_t)cloud * 3 * n : ori + (size_t)cloud * 3 * m;
  const int nq = dir1 ? m : n;
  const int nc = dir1 ? n : m;
  const int32_t* pq = dir1 ? perm_o : perm_a;     // position -> original index (clouds are stored by position)
  const int32_t* pc = dir1 ? perm_a : perm_o;
  const int32_t* ipc = dir1 ? iperm_a : iperm_o;  // original index -> position (only needed to honour hints)
  if (pq) pq += (size_t)cloud * nq;
  if (pc) pc += (size_t)cloud * nc;
  if (ipc) ipc += (size_t)cloud * nc;
  float* dout = dir1 ? d_o2a + (size_t)cloud * m : d_a2o + (size_t)cloud * n;
  int32_t* iout = dir1 ? istar + (size_t)cloud * m : jstar + (size_t)cloud * n;
  const int32_t* hint = dir1 ? hint_o2a : hint_a2o;  // may alias iout: each thread reads its own slot first
  if (hint) hint += (size_t)cloud * nq;
  const int lane = threadIdx.x & 31;

  float qx[Q], qy[Q], qz[Q];
  float2 ax[Q], ay[Q], az[Q];  // -2q, both halves
  float qq2[Q], thr[Q];
  float best[Q];
  int bi[Q];   // ORIGINAL index of the best candidate
  int qo[Q];   // ORIGINAL index of the query (-1: this slot is past the end)
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    // position in visiting order: a thread's Q queries are neighbours, a warp covers 32*Q consecutive positions
    const int slot = tile * (NN_THREADS * Q) + threadIdx.x * Q + q;
    const int sl = min(slot, nq - 1);
    const int o = pq ? pq[sl] : sl;
    qo[q] = slot < nq ? o : -1;
    const float x = qbase[sl], y = qbase[nq + sl], z = qbase[2 * nq + sl];
    qx[q] = x; qy[q] = y; qz[q] = z;
    ax[q] = make_float2(-2.f * x, -2.f * x);
    ay[q] = make_float2(-2.f * y, -2.f * y);
    az[q] = make_float2(-2.f * z, -2.f * z);
    qq2[q] = x * x + y * y + z * z;
    // seed: hinted original index (needs the inverse permutation when the clouds are permuted), else the
    // candidate at the same POSITION (both clouds share the visiting order in the attack: adv_t ~ ori_t)
    int spos = min(sl, nc - 1);
    if (hint && (!pc || ipc)) {
      const int h = min(max(hint[o], 0), nc - 1);
      spos = pc ? ipc[h] : h;
    }
    best[q] = dist2(cbase[spos], cbase[nc + spos], cbase[2 * nc + spos], x, y, z);
    bi[q] = pc ? pc[spos] : spos;
    thr[q] = 0.f;  // set per candidate group below (the margin depends on the group's max |c|^2)
  }

  for (int c0 = 0; c0 < nc; c0 += NN_CHUNK) {
    const int cn = min(NN_CHUNK, nc - c0);
    const int cn32 = (cn + 31) & ~31;
    __syncthreads();
    for (int t = threadIdx.x; t < cn32; t += NN_THREADS) {  // a warp stages 32 consecutive positions = one group
      const bool ok = t < cn;
      const int o = ok ? (pc ? pc[c0 + t] : c0 + t) : 0;
      // padding candidates sit far away (finite, so the filter never sees inf*0): they can never pass
      const float x = ok ? cbase[c0 + t] : 1e18f, y = ok ? cbase[nc + c0 + t] : 0.f, z = ok ? cbase[2 * nc + c0 + t] : 0.f;
      sx[t] = x; sy[t] = y; sz[t] = z;
      sw[t] = x * x + y * y + z * z;
      so[t] = o;
      float lx = ok ? x : 3e38f, ly = ok ? y : 3e38f, lz = ok ? z : 3e38f;
      float hx = ok ? x : -3e38f, hy = ok ? y : -3e38f, hz = ok ? z : -3e38f;
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, s)); hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, s));
        ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, s)); hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, s));
        lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, s)); hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, s));
      }
      float w2 = ok ? sw[t] : 0.f;
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) w2 = fmaxf(w2, __shfl_xor_sync(0xffffffffu, w2, s));
      if (lane == 0) {
        float* bb = sbb[t >> 5];
        bb[0] = lx; bb[1] = ly; bb[2] = lz; bb[3] = hx; bb[4] = hy; bb[5] = hz; bb[6] = w2;
      }
    }
    __syncthreads();
    for (int g = 0; g < cn32; g += 32) {
      const float* bb = sbb[g >> 5];
      if (PRUNE) {
        // can any candidate of this group reach the best of any query of this warp?  (box distance, conservative)
        bool need = false;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const float ex = fmaxf(fmaxf(bb[0] - qx[q], qx[q] - bb[3]), 0.f);
          const float ey = fmaxf(fmaxf(bb[1] - qy[q], qy[q] - bb[4]), 0.f);
          const float ez = fmaxf(fmaxf(bb[2] - qz[q], qz[q] - bb[5]), 0.f);
          const float lb = ex * ex + ey * ey + ez * ez;
          need |= lb * 0.9999f <= best[q];  // d_pin >= lb*(1-1e-6) for every candidate in the box
        }
        if (!__any_sync(0xffffffffu, need)) continue;
      }
      float margin[Q];
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        margin[q] = 7.62939453125e-6f * 1.001f * fmaxf(qq2[q], bb[6]);  // 2^-17 * R^2 for THIS group
        thr[q] = (best[q] - qq2[q]) + margin[q];
      }
      for (int j = g; j < g + 32; j += 8) {
        const float4 cxa = *reinterpret_cast<const float4*>(sx + j), cxb = *reinterpret_cast<const float4*>(sx + j + 4);
        const float4 cya = *reinterpret_cast<const float4*>(sy + j), cyb = *reinterpret_cast<const float4*>(sy + j + 4);
        const float4 cza = *reinterpret_cast<const float4*>(sz + j), czb = *reinterpret_cast<const float4*>(sz + j + 4);
        const float4 cwa = *reinterpret_cast<const float4*>(sw + j), cwb = *reinterpret_cast<const float4*>(sw + j + 4);
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          // t = |c|^2 - 2 q.c  for 8 candidates: 12 FFMA2
          float2 t01 = __ffma2_rn(az[q], make_float2(cza.x, cza.y), make_float2(cwa.x, cwa.y));
          float2 t23 = __ffma2_rn(az[q], make_float2(cza.z, cza.w), make_float2(cwa.z, cwa.w));
          float2 t45 = __ffma2_rn(az[q], make_float2(czb.x, czb.y), make_float2(cwb.x, cwb.y));
          float2 t67 = __ffma2_rn(az[q], make_float2(czb.z, czb.w), make_float2(cwb.z, cwb.w));
          t01 = __ffma2_rn(ay[q], make_float2(cya.x, cya.y), t01);
          t23 = __ffma2_rn(ay[q], make_float2(cya.z, cya.w), t23);
          t45 = __ffma2_rn(ay[q], make_float2(cyb.x, cyb.y), t45);
          t67 = __ffma2_rn(ay[q], make_float2(cyb.z, cyb.w), t67);
          t01 = __ffma2_rn(ax[q], make_float2(cxa.x, cxa.y), t01);
          t23 = __ffma2_rn(ax[q], make_float2(cxa.z, cxa.w), t23);
          t45 = __ffma2_rn(ax[q], make_float2(cxb.x, cxb.y), t45);
          t67 = __ffma2_rn(ax[q], make_float2(cxb.z, cxb.w), t67);
          const float mn = fminf(fminf(fminf(t01.x, t01.y), fminf(t23.x, t23.y)), fminf(fminf(t45.x, t45.y), fminf(t67.x, t67.y)));
          if (mn <= thr[q]) {  // rare: pinned arithmetic + exact lexicographic update over the 8 candidates
            const float cx8[8] = {cxa.x, cxa.y, cxa.z, cxa.w, cxb.x, cxb.y, cxb.z, cxb.w};
            const float cy8[8] = {cya.x, cya.y, cya.z, cya.w, cyb.x, cyb.y, cyb.z, cyb.w};
            const float cz8[8] = {cza.x, cza.y, cza.z, cza.w, czb.x, czb.y, czb.z, czb.w};
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              const float d = dist2(cx8[t], cy8[t], cz8[t], qx[q], qy[q], qz[q]);
              const int o = so[j + t];
              if (j + t < cn && (d < best[q] || (d == best[q] && o < bi[q]))) { best[q] = d; bi[q] = o; }
            }
            thr[q] = (best[q] - qq2[q]) + margin[q];
          }
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < Q; ++q)
    if (qo[q] >= 0) {
      dout[qo[q]] = best[q];
      iout[qo[q]] = bi[q];
    }
}

}  // namespace geoa3

extern "C" int geoa3_nn_pair(const float* adv, const float* ori, int b, int n, int m, const int32_t* perm_a,
                             const int32_t* perm_o, const int32_t* iperm_a, const int32_t* iperm_o,
                             const int32_t* hint_a2o, const int32_t* hint_o2a, float* d_a2o, int32_t* jstar,
                             float* d_o2a, int32_t* istar, geoa3_stream_t stream) {
  using namespace geoa3;
  GEOA3_CHECK_ARG(adv && ori && d_a2o && jstar);
  GEOA3_CHECK_ARG(b > 0 && n > 0 && m > 0);
  GEOA3_CHECK_ARG((d_o2a == nullptr) == (istar == nullptr));
  GEOA3_CHECK_ARG((perm_a == nullptr) == (perm_o == nullptr));  // either both clouds are pre-arranged or none
  if (b > 65535) return GEOA3_EUNSUPPORTED;
  const int per = NN_THREADS * NN_Q;
  const int tiles_a = ceil_div(n, per);
  const int tiles_b = d_o2a ? ceil_div(m, per) : 0;
  dim3 grid(tiles_a + tiles_b, b, 1);
  if (perm_a || perm_o)
    nn_pair_kernel<NN_Q, true><<<grid, NN_THREADS, 0, (cudaStream_t)stream>>>(
        adv, ori, n, m, perm_a, perm_o, iperm_a, iperm_o, hint_a2o, hint_o2a, d_a2o, jstar, d_o2a, istar, tiles_a);
  else
    nn_pair_kernel<NN_Q, false><<<grid, NN_THREADS, 0, (cudaStream_t)stream>>>(
        adv, ori, n, m, nullptr, nullptr, nullptr, nullptr, hint_a2o, hint_o2a, d_a2o, jstar, d_o2a, istar, tiles_a);
  return GEOA3_LAUNCH_RESULT();
}
