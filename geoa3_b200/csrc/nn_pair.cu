// nn_pair.cu — fused bidirectional 1-NN (Chamfer / Hausdorff / normal borrowing front end).
//
// One launch covers both directions (blockIdx.z): each CTA owns a tile of queries of one cloud and
// streams the cloud's candidates through shared memory (SoA x[],y[],z[] so that one LDS.128 feeds
// four candidates to every lane as a broadcast).  Each thread keeps Q queries in registers and
// evaluates two candidates per instruction with the packed fp32 pipe (FADD2/FMUL2/FFMA2).  The
// N x M matrix only ever exists as registers; per query the running (min, argmin) is updated with
// strict '<' while candidates are visited in ascending index => ties resolve to the lowest index.
#include "common.cuh"

namespace geoa3 {

constexpr int NN_THREADS = 256;
constexpr int NN_Q = 2;          // queries per thread
constexpr int NN_CHUNK = 2048;   // candidates staged per shared-memory pass (24 KB)

template <int Q>
__global__ void __launch_bounds__(NN_THREADS)
nn_pair_kernel(const float* __restrict__ adv, const float* __restrict__ ori, int n, int m,
               float* __restrict__ d_a2o, int32_t* __restrict__ jstar,
               float* __restrict__ d_o2a, int32_t* __restrict__ istar, int tiles_a) {
  __shared__ __align__(16) float sx[NN_CHUNK];
  __shared__ __align__(16) float sy[NN_CHUNK];
  __shared__ __align__(16) float sz[NN_CHUNK];

  const int cloud = blockIdx.y;
  int tile = blockIdx.x;
  // direction 0: queries = adv (n), candidates = ori (m); direction 1: swapped.
  const bool dir1 = tile >= tiles_a;
  if (dir1) tile -= tiles_a;
  const float* qbase = dir1 ? ori + (size_t)cloud * 3 * m : adv + (size_t)cloud * 3 * n;
  const float* cbase = dir1 ? adv + (size_t)cloud * 3 * n : ori + (size_t)cloud * 3 * m;
  const int nq = dir1 ? m : n;
  const int nc = dir1 ? n : m;
  float* dout = dir1 ? d_o2a + (size_t)cloud * m : d_a2o + (size_t)cloud * n;
  int32_t* iout = dir1 ? istar + (size_t)cloud * m : jstar + (size_t)cloud * n;

  float2 nqx[Q], nqy[Q], nqz[Q];
  float best[Q];
  int bi[Q];
  int qi[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    qi[q] = tile * (NN_THREADS * Q) + q * NN_THREADS + threadIdx.x;
    const int qq = min(qi[q], nq - 1);
    const float x = -qbase[qq], y = -qbase[nq + qq], z = -qbase[2 * nq + qq];
    nqx[q] = make_float2(x, x);
    nqy[q] = make_float2(y, y);
    nqz[q] = make_float2(z, z);
    best[q] = __int_as_float(0x7f800000);
    bi[q] = 0;
  }

  for (int c0 = 0; c0 < nc; c0 += NN_CHUNK) {
    const int cn = min(NN_CHUNK, nc - c0);
    const int cn4 = (cn + 3) & ~3;
    __syncthreads();
    for (int t = threadIdx.x; t < cn4; t += NN_THREADS) {
      const bool ok = t < cn;
      // padding candidates sit at +inf: their distance is +inf and can never win a strict '<'
      sx[t] = ok ? cbase[c0 + t] : __int_as_float(0x7f800000);
      sy[t] = ok ? cbase[nc + c0 + t] : 0.f;
      sz[t] = ok ? cbase[2 * nc + c0 + t] : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int j = 0; j < cn4; j += 4) {
      const float4 cx = *reinterpret_cast<const float4*>(sx + j);
      const float4 cy = *reinterpret_cast<const float4*>(sy + j);
      const float4 cz = *reinterpret_cast<const float4*>(sz + j);
      const int jj = c0 + j;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const float2 d01 = dist2x2(make_float2(cx.x, cx.y), make_float2(cy.x, cy.y), make_float2(cz.x, cz.y),
                                   nqx[q], nqy[q], nqz[q]);
        const float2 d23 = dist2x2(make_float2(cx.z, cx.w), make_float2(cy.z, cy.w), make_float2(cz.z, cz.w),
                                   nqx[q], nqy[q], nqz[q]);
        if (d01.x < best[q]) { best[q] = d01.x; bi[q] = jj; }
        if (d01.y < best[q]) { best[q] = d01.y; bi[q] = jj + 1; }
        if (d23.x < best[q]) { best[q] = d23.x; bi[q] = jj + 2; }
        if (d23.y < best[q]) { best[q] = d23.y; bi[q] = jj + 3; }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < Q; ++q)
    if (qi[q] < nq) {
      dout[qi[q]] = best[q];
      iout[qi[q]] = bi[q];
    }
}

}  // namespace geoa3

extern "C" int geoa3_nn_pair(const float* adv, const float* ori, int b, int n, int m, float* d_a2o,
                             int32_t* jstar, float* d_o2a, int32_t* istar, geoa3_stream_t stream) {
  using namespace geoa3;
  GEOA3_CHECK_ARG(adv && ori && d_a2o && jstar);
  GEOA3_CHECK_ARG(b > 0 && n > 0 && m > 0);
  GEOA3_CHECK_ARG((d_o2a == nullptr) == (istar == nullptr));
  if (b > 65535) return GEOA3_EUNSUPPORTED;
  const int per = NN_THREADS * NN_Q;
  const int tiles_a = ceil_div(n, per);
  const int tiles_b = d_o2a ? ceil_div(m, per) : 0;
  dim3 grid(tiles_a + tiles_b, b, 1);
  nn_pair_kernel<NN_Q><<<grid, NN_THREADS, 0, (cudaStream_t)stream>>>(adv, ori, n, m, d_a2o, jstar, d_o2a, istar,
                                                                      tiles_a);
  return GEOA3_LAUNCH_RESULT();
}
