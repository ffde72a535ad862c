// kappa_loss.cu — curvature (kappa), per-cloud loss reductions and the fused deterministic backward.
//
// Both kernels run one CTA per cloud with the cloud resident in shared memory:
//   forward : kappa_i, borrowed normals, CD / HD(+argmax) / curvature-loss reductions in fixed order;
//   backward: d(g_cd*CD + g_hd*HD + g_cu*CUR + <g_kappa,kappa>)/d adv.  Scatter terms (Chamfer column
//             term through istar, curvature neighbour term through nbr) are turned into gathers: a
//             CSR-by-target list is built in shared memory (integer shared atomics hand out the slots,
//             then every target sorts its short segment by source id — csr.cuh), so every target sums
//             its contributions in ascending source order — no float atomics, bitwise reproducible.
#include <cstdlib>

#include "common.cuh"
#include "csr.cuh"

namespace geoa3 {

#ifndef KL_THREADS_N
#define KL_THREADS_N 512
#endif
constexpr int KL_THREADS = KL_THREADS_N;   // forward: one CTA per cloud
constexpr int KL_WARPS = KL_THREADS / 32;
constexpr int BW_THREADS = 1024;  // backward: one CTA per cloud, one target per thread at n = 1024
constexpr int BW_WARPS = BW_THREADS / 32;
constexpr float KL_EPS = 1e-12f;  // utility.py:30 _normalize eps

// ---------------------------------------------------------------- block helpers (fixed order)
__device__ __forceinline__ float block_sum(float v, float* scratch /*KL_WARPS*/) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) scratch[w] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < KL_WARPS; ++i) t += scratch[i];
  return t;
}

// max with lowest index on ties
__device__ __forceinline__ void block_argmax(float& v, int& idx, float* sv, int* si) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) { sv[w] = v; si[w] = idx; }
  __syncthreads();
  v = sv[0]; idx = si[0];
#pragma unroll
  for (int i = 1; i < KL_WARPS; ++i)
    if (sv[i] > v || (sv[i] == v && si[i] < idx)) { v = sv[i]; idx = si[i]; }
}

// ---------------------------------------------------------------- forward
__global__ void __launch_bounds__(KL_THREADS)
kappa_loss_fwd_kernel(const float* __restrict__ pc, const float* __restrict__ normal, int n_normal,
                      const int32_t* __restrict__ jstar, const int32_t* __restrict__ nbr, int k,
                      const float* __restrict__ d_a2o, const float* __restrict__ d_o2a, int m,
                      const float* __restrict__ kappa_ori, int n, float* __restrict__ kappa,
                      float* __restrict__ nrm_out, float* __restrict__ cd, float* __restrict__ hd,
                      int32_t* __restrict__ hd_arg, float* __restrict__ curv) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* pts = reinterpret_cast<float4*>(smem_raw);
  __shared__ float s_f[KL_WARPS];
  __shared__ int s_i[KL_WARPS];

  const int cloud = blockIdx.x;
  const int tid = threadIdx.x;
  const float* p = pc + (size_t)cloud * 3 * n;
  const bool do_kappa = nbr != nullptr && k > 0;
  if (do_kappa) {
    for (int i = tid; i < n; i += KL_THREADS) pts[i] = make_float4(p[i], p[n + i], p[2 * n + i], 0.f);
  }
  __syncthreads();

  float s1 = 0.f, sc = 0.f, mx = -1.f;
  int am = 0x7fffffff;
  const float inv_k = do_kappa ? 1.f / (float)k : 0.f;
  for (int i = tid; i < n; i += KL_THREADS) {
    const size_t gi = (size_t)cloud * n + i;
    const int js = jstar ? jstar[gi] : i;
    float kap = 0.f;
    if (do_kappa) {
      const float* nb = normal + (size_t)cloud * 3 * n_normal;
      const float nx = nb[js], ny = nb[n_normal + js], nz = nb[2 * n_normal + js];
      if (nrm_out) {
        float* no = nrm_out + (size_t)cloud * 3 * n;
        no[i] = nx; no[n + i] = ny; no[2 * n + i] = nz;
      }
      const float4 pi = pts[i];
      const int32_t* nbi = nbr + gi * k;
      float acc = 0.f;
      auto term = [&](int j) {
        const float4 pj = pts[j];
        const float vx = pj.x - pi.x, vy = pj.y - pi.y, vz = pj.z - pi.z;
        // v / max(|v|, 1e-12): |v| >= 1e-12 <=> |v|^2 >= 1e-24; MUFU.RSQ is accurate to ~1e-7 relative
        const float s2 = vx * vx + vy * vy + vz * vz;
        const float inv = s2 >= 1e-24f ? rsqrtf(s2) : 1e12f;
        acc += fabsf((vx * nx + vy * ny + vz * nz) * inv);
      };
      if ((k & 3) == 0) {
        // 16-byte index rows: up to 16 neighbour indices are in flight before the first one is used (a scalar
        // load per neighbour put k dependent global round trips on every thread's critical path)
        const int4* nb4 = reinterpret_cast<const int4*>(nbi);
        const int k4 = k >> 2;
        for (int t = 0; t < k4; t += 4) {
          int4 q[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) q[u] = t + u < k4 ? nb4[t + u] : make_int4(0, 0, 0, 0);
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (t + u < k4) { term(q[u].x); term(q[u].y); term(q[u].z); term(q[u].w); }
        }
      } else {
        for (int t = 0; t < k; ++t) term(nbi[t]);
      }
      kap = acc * inv_k;
      if (kappa) kappa[gi] = kap;
      if (curv) {
        const float e = kap - kappa_ori[(size_t)cloud * n_normal + js];
        sc += e * e;
      }
    }
    if (d_a2o) {
      const float d = d_a2o[gi];
      s1 += d;
      if (d > mx) { mx = d; am = i; }  // ascending i per thread + strict '>' keeps the lowest index
    }
  }
  if (cd || hd || hd_arg) {
    float s2 = 0.f;
    if (d_o2a)
      for (int j = tid; j < m; j += KL_THREADS) s2 += d_o2a[(size_t)cloud * m + j];
    const float t1 = block_sum(s1, s_f);
    const float t2 = d_o2a ? block_sum(s2, s_f) : 0.f;
    block_argmax(mx, am, s_f, s_i);
    if (tid == 0) {
      if (cd) cd[cloud] = t1 / (float)n + (d_o2a ? t2 / (float)m : 0.f);
      if (hd) hd[cloud] = mx;
      if (hd_arg) hd_arg[cloud] = am;
    }
  }
  if (curv && do_kappa) {
    const float tc = block_sum(sc, s_f);
    if (tid == 0) curv[cloud] = tc / (float)n;
  }
}

// d|<n, v/|v|>| / dv scaled by f0 = gk/k  (clamp at 1e-12 passes no grad to the norm)
__device__ __forceinline__ float3 dkappa_dv(float vx, float vy, float vz, float nx, float ny, float nz, float f0) {
  const float s2 = vx * vx + vy * vy + vz * vz;
  float3 r;
  if (s2 >= 1e-24f) {  // |v| >= 1e-12
    const float inv = rsqrtf(s2);
    const float ux = vx * inv, uy = vy * inv, uz = vz * inv;
    const float s = ux * nx + uy * ny + uz * nz;
    const float sg = (s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f);
    const float f = f0 * sg * inv;
    r.x = f * (nx - ux * s); r.y = f * (ny - uy * s); r.z = f * (nz - uz * s);
  } else {
    const float s = (vx * nx + vy * ny + vz * nz) * 1e12f;
    const float sg = (s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f);
    const float f = f0 * sg * 1e12f;
    r.x = f * nx; r.y = f * ny; r.z = f * nz;
  }
  return r;
}

struct BwdLayout {  // byte offsets into dynamic shared memory
  int pts, nrm, offs1, offs2, whist, ent1, ent2, total, W;
  unsigned k_magic;  // ceil(2^32 / k)
};

__global__ void __launch_bounds__(BW_THREADS, 2)
loss_bwd_kernel(const float* __restrict__ adv, const float* __restrict__ ori, const float* __restrict__ nrm_adv,
                const float* __restrict__ kappa_adv, const float* __restrict__ kappa_ori,
                const int32_t* __restrict__ jstar, const int32_t* __restrict__ istar,
                const int32_t* __restrict__ nbr, const int32_t* __restrict__ hd_arg,
                const float* __restrict__ g_cd, const float* __restrict__ g_hd, const float* __restrict__ g_cu,
                const float* __restrict__ g_kappa, int n, int m, int k, float* __restrict__ grad_adv,
                BwdLayout lay) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* pts = reinterpret_cast<float4*>(smem_raw + lay.pts);
  float4* nrm = reinterpret_cast<float4*>(smem_raw + lay.nrm);
  int* offs1 = reinterpret_cast<int*>(smem_raw + lay.offs1);
  int* offs2 = reinterpret_cast<int*>(smem_raw + lay.offs2);
  int* whist = reinterpret_cast<int*>(smem_raw + lay.whist);  // n counters / fill cursors
  uint16_t* ent1 = reinterpret_cast<uint16_t*>(smem_raw + lay.ent1);
  uint16_t* ent2 = reinterpret_cast<uint16_t*>(smem_raw + lay.ent2);
  __shared__ int scan_scratch[BW_WARPS + 1];

  const int cloud = blockIdx.x, tid = threadIdx.x;
  const float* a = adv + (size_t)cloud * 3 * n;
  const float* o = ori + (size_t)cloud * 3 * m;
  const bool do_curv = k > 0 && nbr != nullptr && (g_cu != nullptr || g_kappa != nullptr);
  const bool do_col = istar != nullptr && g_cd != nullptr;
  const float gcd = g_cd ? g_cd[cloud] : 0.f;
  const float ghd = g_hd ? g_hd[cloud] : 0.f;
  const float gcu = g_cu ? g_cu[cloud] : 0.f;

  for (int i = tid; i < n; i += BW_THREADS) {
    const size_t gi = (size_t)cloud * n + i;
    float gk = 0.f;
    if (do_curv) {
      if (g_cu) gk = gcu * (2.f / (float)n) * (kappa_adv[gi] - kappa_ori[(size_t)cloud * m + jstar[gi]]);
      if (g_kappa) gk += g_kappa[gi];
      const float* nn = nrm_adv + (size_t)cloud * 3 * n;
      nrm[i] = make_float4(nn[i], nn[n + i], nn[2 * n + i], 0.f);
    }
    pts[i] = make_float4(a[i], a[n + i], a[2 * n + i], gk);
  }
  __syncthreads();
  if (do_curv)
    build_csr_sorted<BW_THREADS, uint16_t>(nbr + (size_t)cloud * n * k, n * k, n, lay.k_magic, offs1, whist, ent1,
                                           scan_scratch);
  if (do_col)
    build_csr_sorted<BW_THREADS, uint16_t>(istar + (size_t)cloud * m, m, n, 0u, offs2, whist, ent2, scan_scratch);

  const int ha = (g_hd && hd_arg) ? hd_arg[cloud] : -1;
  const float w_row = gcd * (2.f / (float)n), w_col = gcd * (2.f / (float)m);
  const float inv_k = k > 0 ? 1.f / (float)k : 0.f;
  for (int p = tid; p < n; p += BW_THREADS) {
    const size_t gp = (size_t)cloud * n + p;
    const float4 ap = pts[p];
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (g_cd || ha == p) {
      const int js = jstar[gp];
      const float rx = ap.x - o[js], ry = ap.y - o[m + js], rz = ap.z - o[2 * m + js];
      if (g_cd) { gx = w_row * rx; gy = w_row * ry; gz = w_row * rz; }
      if (do_col) {
        const int e1 = offs2[p + 1];
        for (int e = offs2[p]; e < e1; ++e) {
          const int j = ent2[e];
          gx += w_col * (ap.x - o[j]); gy += w_col * (ap.y - o[m + j]); gz += w_col * (ap.z - o[2 * m + j]);
        }
      }
      if (ha == p) { gx += ghd * 2.f * rx; gy += ghd * 2.f * ry; gz += ghd * 2.f * rz; }
    }
    if (do_curv) {
      // own term: -sum_m d kappa_p / d v_pm
      const float4 np_ = nrm[p];
      const float f0 = ap.w * inv_k;
      const int32_t* nbp = nbr + gp * k;
      float ox = 0.f, oy = 0.f, oz = 0.f;
      if ((k & 3) == 0) {  // 16-byte rows: fetch four neighbour indices per load
        for (int t = 0; t < k; t += 4) {
          const int4 j4 = *reinterpret_cast<const int4*>(nbp + t);
          const int js4[4] = {j4.x, j4.y, j4.z, j4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 aj = pts[js4[u]];
            const float3 dv = dkappa_dv(aj.x - ap.x, aj.y - ap.y, aj.z - ap.z, np_.x, np_.y, np_.z, f0);
            ox += dv.x; oy += dv.y; oz += dv.z;
          }
        }
      } else {
        for (int t = 0; t < k; ++t) {
          const float4 aj = pts[nbp[t]];
          const float3 dv = dkappa_dv(aj.x - ap.x, aj.y - ap.y, aj.z - ap.z, np_.x, np_.y, np_.z, f0);
          ox += dv.x; oy += dv.y; oz += dv.z;
        }
      }
      gx -= ox; gy -= oy; gz -= oz;
      // incoming terms: every i that lists p as a neighbour, ascending i
      const int e1 = offs1[p + 1];
      for (int e = offs1[p]; e < e1; ++e) {
        const int i = ent1[e];
        const float4 ai = pts[i];
        const float4 ni = nrm[i];
        const float3 dv = dkappa_dv(ap.x - ai.x, ap.y - ai.y, ap.z - ai.z, ni.x, ni.y, ni.z, ai.w * inv_k);
        gx += dv.x; gy += dv.y; gz += dv.z;
      }
    }
    float* g = grad_adv + (size_t)cloud * 3 * n;
    g[p] = gx; g[n + p] = gy; g[2 * n + p] = gz;
  }
}

static bool plan_bwd_layout(int n, int m, int k, bool do_curv, bool do_col, BwdLayout* L) {
  const int budget = 227 * 1024 - 256;
  L->k_magic = k > 1 ? (unsigned)(((1ull << 32) + (unsigned)k - 1) / (unsigned)k) : 0u;
  int off = 0;
  auto take = [&](size_t bytes) { int o = off; off += (int)((bytes + 15) & ~(size_t)15); return o; };
  L->pts = take((size_t)16 * n);
  L->nrm = take(do_curv ? (size_t)16 * n : 0);
  L->offs1 = take(do_curv ? (size_t)4 * (n + 1) : 0);
  L->offs2 = take(do_col ? (size_t)4 * (n + 1) : 0);
  // build_csr_sorted() uses this as `int cnt[n]` (counters, then fill cursors) for BOTH lists, one after the other
  L->whist = take((do_curv || do_col) ? (size_t)4 * n : 0);
  L->ent1 = take(do_curv ? (size_t)2 * n * k : 0);
  L->ent2 = take(do_col ? (size_t)2 * m : 0);
  L->total = off;
  L->W = 1;
  return off <= budget;
}

// ---------------------------------------------------------------------------------------------------------
// Large clouds (n*k > 65535 edges, or a shared-memory plan that does not fit): the same gather with the two
// CSRs and the per-point float4 arrays in a caller-provided GLOBAL workspace.  Same summation order as the fused
// kernel (own terms, then incoming edges by ascending source), so the results are bit-identical to it.
// Workspace per cloud (ints/floats, 16-byte aligned): P[n] float4 (a_i, gk_i), N[n] float4 (borrowed normal),
// offs1[n+1], ent1[n*k], offs2[n+1], ent2[m].
constexpr int BWL_THREADS = 512;

struct BwdLargeLayout { size_t P, N, offs1, ent1, offs2, ent2, per_cloud; };  // offsets in 4-byte words

static BwdLargeLayout plan_bwd_large(int n, int m, int k) {
  BwdLargeLayout L;
  size_t o = 0;
  auto take = [&](size_t words) { size_t r = o; o += (words + 3) & ~(size_t)3; return r; };
  L.P = take((size_t)4 * n); L.N = take((size_t)4 * n);
  L.offs1 = take((size_t)n + 1); L.ent1 = take((size_t)n * (k > 0 ? k : 0));
  L.offs2 = take((size_t)n + 1); L.ent2 = take((size_t)m);
  L.per_cloud = o;
  return L;
}

__global__ void __launch_bounds__(BWL_THREADS)
bwd_large_csr_kernel(const int32_t* __restrict__ keys, int E, int n, int W, float* __restrict__ ws, size_t per_cloud,
                     size_t off_offs, size_t off_ent) {
  extern __shared__ __align__(16) int s_whist_l[];
  __shared__ int scan_scratch[BWL_THREADS / 32 + 1];
  int* base = reinterpret_cast<int*>(ws + (size_t)blockIdx.x * per_cloud);
  build_csr<BWL_THREADS, int, int>(keys + (size_t)blockIdx.x * E, E, n, 0u, base + off_offs, s_whist_l, W,
                                   base + off_ent, scan_scratch);
}

__global__ void bwd_large_pack_kernel(const float* __restrict__ adv, const float* __restrict__ nrm_adv,
                                      const float* __restrict__ kappa_adv, const float* __restrict__ kappa_ori,
                                      const int32_t* __restrict__ jstar, const float* __restrict__ g_cu,
                                      const float* __restrict__ g_kappa, int n, int m, bool do_curv,
                                      float* __restrict__ ws, BwdLargeLayout L) {
  const int cloud = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t gi = (size_t)cloud * n + i;
  const float* a = adv + (size_t)cloud * 3 * n;
  float4* P = reinterpret_cast<float4*>(ws + (size_t)cloud * L.per_cloud + L.P);
  float4* N = reinterpret_cast<float4*>(ws + (size_t)cloud * L.per_cloud + L.N);
  float gk = 0.f;
  if (do_curv) {
    if (g_cu) gk = g_cu[cloud] * (2.f / (float)n) * (kappa_adv[gi] - kappa_ori[(size_t)cloud * m + jstar[gi]]);
    if (g_kappa) gk += g_kappa[gi];
    const float* nn = nrm_adv + (size_t)cloud * 3 * n;
    N[i] = make_float4(nn[i], nn[n + i], nn[2 * n + i], 0.f);
  }
  P[i] = make_float4(a[i], a[n + i], a[2 * n + i], gk);
}

__global__ void bwd_large_gather_kernel(const float* __restrict__ ori, const int32_t* __restrict__ jstar,
                                        const int32_t* __restrict__ nbr, const int32_t* __restrict__ hd_arg,
                                        const float* __restrict__ g_cd, const float* __restrict__ g_hd, int n, int m,
                                        int k, bool do_curv, bool do_col, const float* __restrict__ ws,
                                        BwdLargeLayout L, float* __restrict__ grad_adv) {
  const int cloud = blockIdx.y, p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const float* base = ws + (size_t)cloud * L.per_cloud;
  const float4* P = reinterpret_cast<const float4*>(base + L.P);
  const float4* N = reinterpret_cast<const float4*>(base + L.N);
  const int* offs1 = reinterpret_cast<const int*>(base + L.offs1);
  const int* ent1 = reinterpret_cast<const int*>(base + L.ent1);
  const int* offs2 = reinterpret_cast<const int*>(base + L.offs2);
  const int* ent2 = reinterpret_cast<const int*>(base + L.ent2);
  const float* o = ori ? ori + (size_t)cloud * 3 * m : nullptr;
  const float gcd = g_cd ? g_cd[cloud] : 0.f, ghd = g_hd ? g_hd[cloud] : 0.f;
  const int ha = (g_hd && hd_arg) ? hd_arg[cloud] : -1;
  const float w_row = gcd * (2.f / (float)n), w_col = gcd * (2.f / (float)m);
  const float inv_k = k > 0 ? 1.f / (float)k : 0.f;
  const size_t gp = (size_t)cloud * n + p;
  const float4 ap = P[p];
  float gx = 0.f, gy = 0.f, gz = 0.f;
  if (g_cd || ha == p) {
    const int js = jstar[gp];
    const float rx = ap.x - o[js], ry = ap.y - o[m + js], rz = ap.z - o[2 * m + js];
    if (g_cd) { gx = w_row * rx; gy = w_row * ry; gz = w_row * rz; }
    if (do_col) {
      const int e1 = offs2[p + 1];
      for (int e = offs2[p]; e < e1; ++e) {
        const int j = ent2[e];
        gx += w_col * (ap.x - o[j]); gy += w_col * (ap.y - o[m + j]); gz += w_col * (ap.z - o[2 * m + j]);
      }
    }
    if (ha == p) { gx += ghd * 2.f * rx; gy += ghd * 2.f * ry; gz += ghd * 2.f * rz; }
  }
  if (do_curv) {
    const float4 np_ = N[p];
    const float f0 = ap.w * inv_k;
    const int32_t* nbp = nbr + gp * k;
    float ox = 0.f, oy = 0.f, oz = 0.f;
    for (int t = 0; t < k; ++t) {
      const float4 aj = P[nbp[t]];
      const float3 dv = dkappa_dv(aj.x - ap.x, aj.y - ap.y, aj.z - ap.z, np_.x, np_.y, np_.z, f0);
      ox += dv.x; oy += dv.y; oz += dv.z;
    }
    gx -= ox; gy -= oy; gz -= oz;
    const int e1 = offs1[p + 1];
    for (int e = offs1[p]; e < e1; ++e) {
      const int i = ent1[e] / k;  // edge id -> source point
      const float4 ai = P[i];
      const float4 ni = N[i];
      const float3 dv = dkappa_dv(ap.x - ai.x, ap.y - ai.y, ap.z - ai.z, ni.x, ni.y, ni.z, ai.w * inv_k);
      gx += dv.x; gy += dv.y; gz += dv.z;
    }
  }
  float* g = grad_adv + (size_t)cloud * 3 * n;
  g[p] = gx; g[n + p] = gy; g[2 * n + p] = gz;
}

static bool bwd_fused_ok(int n, int m, int k, bool do_curv, bool do_col, BwdLayout* L) {
  if (getenv("GEOA3_BWD_LARGE")) return false;  // test knob: route every call through the large-cloud path
  if (n > 65535 || m > 65535 || (size_t)n * (size_t)(k > 0 ? k : 1) > 65535) return false;  // uint16 CSR entries
  return plan_bwd_layout(n, m, k, do_curv, do_col, L);
}

// ---------------------------------------------------------------------------------------------------------
// One kernel per attack step for everything that follows the two searches: kappa_i with borrowed normals, the CD / HD /
// curvature reductions AND the gradient of  w_cd*CD + w_hd*HD + w_cu*CUR  for a unit upstream gradient per cloud (the
// loss is linear in its upstream gradient, so autograd's backward is a scale by g[b]).  Same arithmetic as
// kappa_loss_fwd_kernel followed by loss_bwd_kernel — the cloud, the normals and the neighbour rows are staged / fetched
// once, and one launch (with its per-cloud ramp-up) disappears.  Requires the fused shared-memory layout (plan_bwd_layout).
template <int WARPS>
__device__ __forceinline__ float block_sum_w(float v, float* scratch) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) scratch[w] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < WARPS; ++i) t += scratch[i];
  return t;
}


__global__ void __launch_bounds__(BW_THREADS, 2)
geo_fwd_bwd_kernel(const float* __restrict__ adv, const float* __restrict__ ori, const float* __restrict__ normal,
                   const float* __restrict__ kappa_ori, const int32_t* __restrict__ jstar,
                   const int32_t* __restrict__ istar, const int32_t* __restrict__ nbr, int k,
                   const float* __restrict__ d_a2o, const float* __restrict__ d_o2a, float w_cd, float w_hd, float w_cu,
                   int n, int m, float* __restrict__ cd, float* __restrict__ hd, float* __restrict__ curv,
                   float* __restrict__ kappa_out, float* __restrict__ nrm_out, int32_t* __restrict__ hd_arg_out,
                   float* __restrict__ grad_adv, BwdLayout lay) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* pts = reinterpret_cast<float4*>(smem_raw + lay.pts);
  float4* nrm = reinterpret_cast<float4*>(smem_raw + lay.nrm);
  int* offs1 = reinterpret_cast<int*>(smem_raw + lay.offs1);
  int* offs2 = reinterpret_cast<int*>(smem_raw + lay.offs2);
  int* whist = reinterpret_cast<int*>(smem_raw + lay.whist);
  uint16_t* ent1 = reinterpret_cast<uint16_t*>(smem_raw + lay.ent1);
  uint16_t* ent2 = reinterpret_cast<uint16_t*>(smem_raw + lay.ent2);
  __shared__ int scan_scratch[BW_WARPS + 1];
  __shared__ float s_f[BW_WARPS];
  __shared__ int s_i[BW_WARPS];
  __shared__ int s_ha;

  const int cloud = blockIdx.x, tid = threadIdx.x;
  const float* a = adv + (size_t)cloud * 3 * n;
  const float* o = ori + (size_t)cloud * 3 * m;
  const bool do_curv = k > 0 && nbr != nullptr && w_cu != 0.f;
  const bool do_col = istar != nullptr && w_cd != 0.f;
  const float* nb = normal + (size_t)cloud * 3 * m;

  // ---- stage the cloud and the borrowed normals
  for (int i = tid; i < n; i += BW_THREADS) {
    const size_t gi = (size_t)cloud * n + i;
    if (do_curv) {
      const int js = jstar[gi];
      const float nx = nb[js], ny = nb[m + js], nz = nb[2 * m + js];
      nrm[i] = make_float4(nx, ny, nz, 0.f);
      if (nrm_out) {
        float* no = nrm_out + (size_t)cloud * 3 * n;
        no[i] = nx; no[n + i] = ny; no[2 * n + i] = nz;
      }
    }
    pts[i] = make_float4(a[i], a[n + i], a[2 * n + i], 0.f);
  }
  __syncthreads();

  // ---- forward: kappa_i, per-cloud sums (the arithmetic of kappa_loss_fwd_kernel)
  float* gkf = reinterpret_cast<float*>(whist);  // per-point curvature factors (the CSR build reuses this array later)
  float s1 = 0.f, sc = 0.f, mx = -1.f;
  int am = 0x7fffffff;
  const float inv_k = k > 0 ? 1.f / (float)k : 0.f;
#pragma unroll 1
  for (int i = tid; i < n; i += BW_THREADS) {
    float gki = 0.f;
    {
      const size_t gi = (size_t)cloud * n + i;
      if (do_curv) {
        const float4 pi = pts[i];
        const float4 ni = nrm[i];
        const int32_t* nbi = nbr + gi * k;
        float acc = 0.f;
        auto term = [&](int j) {
          const float4 pj = pts[j];
          const float vx = pj.x - pi.x, vy = pj.y - pi.y, vz = pj.z - pi.z;
          const float s2 = vx * vx + vy * vy + vz * vz;
          const float inv = s2 >= 1e-24f ? rsqrtf(s2) : 1e12f;
          acc += fabsf((vx * ni.x + vy * ni.y + vz * ni.z) * inv);
        };
        if ((k & 3) == 0) {
          const int4* nb4 = reinterpret_cast<const int4*>(nbi);
          const int k4 = k >> 2;
          for (int t = 0; t < k4; t += 2) {  // (32 registers per thread: two index quads in flight)
            const int4 q0 = nb4[t];
            const int4 q1 = t + 1 < k4 ? nb4[t + 1] : make_int4(0, 0, 0, 0);
            term(q0.x); term(q0.y); term(q0.z); term(q0.w);
            if (t + 1 < k4) { term(q1.x); term(q1.y); term(q1.z); term(q1.w); }
          }
        } else {
          for (int t = 0; t < k; ++t) term(nbi[t]);
        }
        const float kap = acc * inv_k;
        if (kappa_out) kappa_out[gi] = kap;
        const float e = kap - kappa_ori[(size_t)cloud * m + jstar[gi]];
        sc += e * e;
        gki = w_cu * (2.f / (float)n) * e;
      }
      const float d = d_a2o[gi];
      s1 += d;
      if (d > mx) { mx = d; am = i; }  // ascending i per thread + strict '>' keeps the lowest index
    }
    if (do_curv) gkf[i] = gki;  // (no curvature term: the layout has no counter array, the factors stay 0)
  }
  float s2 = 0.f;
  if (d_o2a)
    for (int j = tid; j < m; j += BW_THREADS) s2 += d_o2a[(size_t)cloud * m + j];
  const float t1 = block_sum_w<BW_WARPS>(s1, s_f);
  const float t2 = d_o2a ? block_sum_w<BW_WARPS>(s2, s_f) : 0.f;
  const float tc = do_curv ? block_sum_w<BW_WARPS>(sc, s_f) : 0.f;
  {  // arg-max with the lowest index on ties
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, mx, of);
      const int oi = __shfl_xor_sync(0xffffffffu, am, of);
      if (ov > mx || (ov == mx && oi < am)) { mx = ov; am = oi; }
    }
    const int w = tid >> 5, l = tid & 31;
    __syncthreads();
    if (l == 0) { s_f[w] = mx; s_i[w] = am; }
    __syncthreads();
    if (tid == 0) {
      float v = s_f[0];
      int ix = s_i[0];
      for (int i = 1; i < BW_WARPS; ++i)
        if (s_f[i] > v || (s_f[i] == v && s_i[i] < ix)) { v = s_f[i]; ix = s_i[i]; }
      s_ha = ix;
      if (cd) cd[cloud] = t1 / (float)n + (d_o2a ? t2 / (float)m : 0.f);
      if (hd) hd[cloud] = v;
      if (hd_arg_out) hd_arg_out[cloud] = ix;
      if (curv) curv[cloud] = do_curv ? tc / (float)n : 0.f;
    }
  }
  // every kappa gather is done: the per-point curvature factors may now sit next to the coordinates
  if (do_curv)
    for (int i = tid; i < n; i += BW_THREADS) pts[i].w = gkf[i];
  __syncthreads();

  // ---- backward (the gather of loss_bwd_kernel, upstream gradient 1 per cloud)
  if (do_curv)
    build_csr_sorted<BW_THREADS, uint16_t>(nbr + (size_t)cloud * n * k, n * k, n, lay.k_magic, offs1, whist, ent1,
                                           scan_scratch);
  if (do_col)
    build_csr_sorted<BW_THREADS, uint16_t>(istar + (size_t)cloud * m, m, n, 0u, offs2, whist, ent2, scan_scratch);
  const int ha = w_hd != 0.f ? s_ha : -1;
  const float w_row = w_cd * (2.f / (float)n), w_col = w_cd * (2.f / (float)m);
  for (int p = tid; p < n; p += BW_THREADS) {
    const size_t gp = (size_t)cloud * n + p;
    const float4 ap = pts[p];
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (w_cd != 0.f || ha == p) {
      const int js = jstar[gp];
      const float rx = ap.x - o[js], ry = ap.y - o[m + js], rz = ap.z - o[2 * m + js];
      if (w_cd != 0.f) { gx = w_row * rx; gy = w_row * ry; gz = w_row * rz; }
      if (do_col) {
        const int e1 = offs2[p + 1];
        for (int e = offs2[p]; e < e1; ++e) {
          const int j = ent2[e];
          gx += w_col * (ap.x - o[j]); gy += w_col * (ap.y - o[m + j]); gz += w_col * (ap.z - o[2 * m + j]);
        }
      }
      if (ha == p) { gx += w_hd * 2.f * rx; gy += w_hd * 2.f * ry; gz += w_hd * 2.f * rz; }
    }
    if (do_curv) {
      const float4 np_ = nrm[p];
      const float f0 = ap.w * inv_k;
      const int32_t* nbp = nbr + gp * k;
      float ox = 0.f, oy = 0.f, oz = 0.f;
      if ((k & 3) == 0) {
        for (int t = 0; t < k; t += 4) {
          const int4 j4 = *reinterpret_cast<const int4*>(nbp + t);
          const int js4[4] = {j4.x, j4.y, j4.z, j4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 aj = pts[js4[u]];
            const float3 dv = dkappa_dv(aj.x - ap.x, aj.y - ap.y, aj.z - ap.z, np_.x, np_.y, np_.z, f0);
            ox += dv.x; oy += dv.y; oz += dv.z;
          }
        }
      } else {
        for (int t = 0; t < k; ++t) {
          const float4 aj = pts[nbp[t]];
          const float3 dv = dkappa_dv(aj.x - ap.x, aj.y - ap.y, aj.z - ap.z, np_.x, np_.y, np_.z, f0);
          ox += dv.x; oy += dv.y; oz += dv.z;
        }
      }
      gx -= ox; gy -= oy; gz -= oz;
      const int e1 = offs1[p + 1];
      for (int e = offs1[p]; e < e1; ++e) {
        const int i = ent1[e];
        const float4 ai = pts[i];
        const float4 ni = nrm[i];
        const float3 dv = dkappa_dv(ap.x - ai.x, ap.y - ai.y, ap.z - ai.z, ni.x, ni.y, ni.z, ai.w * inv_k);
        gx += dv.x; gy += dv.y; gz += dv.z;
      }
    }
    float* g = grad_adv + (size_t)cloud * 3 * n;
    g[p] = gx; g[n + p] = gy; g[2 * n + p] = gz;
  }
}

}  // namespace geoa3


extern "C" int geoa3_kappa_loss_fwd(const float* pc, const float* normal, const int32_t* jstar,
                                    const int32_t* nbr, int k, const float* d_a2o, const float* d_o2a,
                                    const float* kappa_ori, int b, int n, int m, float* kappa, float* nrm_out,
                                    float* cd, float* hd, int32_t* hd_arg, float* curv, geoa3_stream_t stream) {
  using namespace geoa3;
  GEOA3_CHECK_ARG(pc && b > 0 && n > 0 && m > 0);
  const bool do_kappa = nbr != nullptr && k > 0;
  if (do_kappa) GEOA3_CHECK_ARG(normal);
  if (!jstar && do_kappa) GEOA3_CHECK_ARG(m == n);
  if (curv) GEOA3_CHECK_ARG(do_kappa && kappa_ori && jstar);
  if (cd || hd || hd_arg) GEOA3_CHECK_ARG(d_a2o);
  const size_t smem = do_kappa ? (size_t)16 * n : 0;
  if (smem > 227 * 1024 - 1024) return GEOA3_EUNSUPPORTED;
  static PerDeviceOnce attr_done;
  if (attr_done.needed()) {
    cudaError_t e = cudaFuncSetAttribute(kappa_loss_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024 - 1024);
    if (e != cudaSuccess) return (int)e;
    attr_done.done();
  }
  kappa_loss_fwd_kernel<<<b, KL_THREADS, smem, (cudaStream_t)stream>>>(
      pc, normal, m, jstar, nbr, k, d_a2o, d_o2a, m, kappa_ori, n, kappa, nrm_out, cd, hd, hd_arg, curv);
  return GEOA3_LAUNCH_RESULT();
}

extern "C" size_t geoa3_loss_bwd_workspace_bytes(int b, int n, int m, int k) {
  using namespace geoa3;
  if (b <= 0 || n <= 0 || m <= 0 || k < 0) return 0;
  BwdLayout L;
  if (bwd_fused_ok(n, m, k, k > 0, true, &L)) return 0;  // the single-kernel path keeps everything in shared memory
  return (size_t)b * plan_bwd_large(n, m, k).per_cloud * 4;
}

extern "C" int geoa3_loss_bwd(const float* adv, const float* ori, const float* nrm_adv, const float* kappa_adv,
                              const float* kappa_ori, const int32_t* jstar, const int32_t* istar,
                              const int32_t* nbr, const int32_t* hd_arg, const float* g_cd, const float* g_hd,
                              const float* g_cu, const float* g_kappa, int b, int n, int m, int k,
                              float* grad_adv, void* workspace, size_t workspace_bytes, geoa3_stream_t stream) {
  using namespace geoa3;
  GEOA3_CHECK_ARG(adv && grad_adv && b > 0 && n > 0 && m > 0 && k >= 0);
  if (g_cd || g_hd) GEOA3_CHECK_ARG(ori && jstar);
  if (g_hd) GEOA3_CHECK_ARG(hd_arg);
  const bool do_curv = k > 0 && nbr && (g_cu || g_kappa);
  if (do_curv) GEOA3_CHECK_ARG(nrm_adv);
  if (do_curv && k > 32) return GEOA3_EUNSUPPORTED;
  if (g_cu && do_curv) GEOA3_CHECK_ARG(kappa_adv && kappa_ori && jstar);
  const bool do_col = istar && g_cd;
  if (b > 65535) return GEOA3_EUNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  BwdLayout L;
  if (bwd_fused_ok(n, m, k, do_curv, do_col, &L)) {
    static PerDeviceOnce attr_done;
    if (attr_done.needed()) {
      cudaError_t e = cudaFuncSetAttribute(loss_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           227 * 1024 - 256);
      if (e != cudaSuccess) return (int)e;
      attr_done.done();
    }
    loss_bwd_kernel<<<b, BW_THREADS, L.total, s>>>(adv, ori, nrm_adv, kappa_adv, kappa_ori, jstar, istar, nbr, hd_arg,
                                                   g_cd, g_hd, g_cu, g_kappa, n, m, k, grad_adv, L);
    return GEOA3_LAUNCH_RESULT();
  }
  // large-cloud path
  const BwdLargeLayout G = plan_bwd_large(n, m, k);
  if (!workspace || workspace_bytes < (size_t)b * G.per_cloud * 4 || ((uintptr_t)workspace & 15)) return GEOA3_EWORKSPACE;
  float* ws = (float*)workspace;
  int W = 16;
  while (W > 1 && (size_t)W * ((n + 1) & ~1) * 4 > 160 * 1024) W >>= 1;
  const size_t hsm = (size_t)W * ((n + 1) & ~1) * 4;
  if (hsm > 200 * 1024) return GEOA3_EUNSUPPORTED;
  static PerDeviceOnce attr_l;
  if (attr_l.needed()) {
    cudaError_t e = cudaFuncSetAttribute(bwd_large_csr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_l.done();
  }
  int err;
  if (do_curv) {
    bwd_large_csr_kernel<<<b, BWL_THREADS, hsm, s>>>(nbr, n * k, n, W, ws, G.per_cloud, G.offs1, G.ent1);
    if ((err = GEOA3_LAUNCH_RESULT())) return err;
  }
  if (do_col) {
    bwd_large_csr_kernel<<<b, BWL_THREADS, hsm, s>>>(istar, m, n, W, ws, G.per_cloud, G.offs2, G.ent2);
    if ((err = GEOA3_LAUNCH_RESULT())) return err;
  }
  const dim3 grid(ceil_div(n, 256), b);
  bwd_large_pack_kernel<<<grid, 256, 0, s>>>(adv, nrm_adv, kappa_adv, kappa_ori, jstar, do_curv ? g_cu : nullptr,
                                            do_curv ? g_kappa : nullptr, n, m, do_curv, ws, G);
  if ((err = GEOA3_LAUNCH_RESULT())) return err;
  bwd_large_gather_kernel<<<grid, 256, 0, s>>>(ori, jstar, nbr, hd_arg, g_cd, g_hd, n, m, k, do_curv, do_col, ws, G,
                                              grad_adv);
  return GEOA3_LAUNCH_RESULT();
}

extern "C" int geoa3_geo_fwd_bwd_supported(int n, int m, int k) {
  using namespace geoa3;
  BwdLayout L;
  return n > 0 && m > 0 && k >= 0 && k <= 32 && bwd_fused_ok(n, m, k, k > 0, true, &L) && L.total <= 227 * 1024 - 1024 ? 1 : 0;
}

extern "C" int geoa3_geo_fwd_bwd(const float* adv, const float* ori, const float* normal, const float* kappa_ori,
                                 const int32_t* jstar, const int32_t* istar, const int32_t* nbr, int k, const float* d_a2o,
                                 const float* d_o2a, float w_cd, float w_hd, float w_cu, int b, int n, int m, float* cd,
                                 float* hd, float* curv, float* kappa, float* nrm_out, int32_t* hd_arg, float* grad_adv,
                                 geoa3_stream_t stream) {
  using namespace geoa3;
  GEOA3_CHECK_ARG(adv && ori && jstar && d_a2o && grad_adv && b > 0 && n > 0 && m > 0 && k >= 0);
  GEOA3_CHECK_ARG((istar == nullptr) == (d_o2a == nullptr));
  const bool do_curv = k > 0 && nbr != nullptr && w_cu != 0.f;
  if (do_curv) GEOA3_CHECK_ARG(normal && kappa_ori);
  if (b > 65535 || !geoa3_geo_fwd_bwd_supported(n, m, do_curv ? k : 0)) return GEOA3_EUNSUPPORTED;
  BwdLayout L;
  if (!bwd_fused_ok(n, m, do_curv ? k : 0, do_curv, istar != nullptr && w_cd != 0.f, &L)) return GEOA3_EUNSUPPORTED;
  if (L.total > 227 * 1024 - 1024) return GEOA3_EUNSUPPORTED;  // (this kernel has 400 bytes of static shared memory)
  static PerDeviceOnce attr_done;
  if (attr_done.needed()) {
    cudaError_t e = cudaFuncSetAttribute(geo_fwd_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    if (e != cudaSuccess) return (int)e;
    attr_done.done();
  }
  geo_fwd_bwd_kernel<<<b, BW_THREADS, L.total, (cudaStream_t)stream>>>(adv, ori, normal, kappa_ori, jstar, istar, nbr,
                                                                      do_curv ? k : 0, d_a2o, d_o2a, w_cd, w_hd, w_cu, n, m,
                                                                      cd, hd, curv, kappa, nrm_out, hd_arg, grad_adv, L);
  return GEOA3_LAUNCH_RESULT();
}
