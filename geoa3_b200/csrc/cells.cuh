// cells.cuh — the per-cloud cell-grid blob shared by geoa3_cell_sort (writer) and the searches that read it
// (geoa3_knn_cells, geoa3_nn_pair_cells): layout, the monotone cell mapping, staging by one TMA bulk copy and the
// shared-memory access helpers of the hot loops.
#pragma once
#include "common.cuh"

namespace geoa3 {

constexpr int KC_GP = 16;            // grid parameters per cloud: lo xyz, inv_h xyz, h xyz, slack, (cells per axis - 1) xyz, cells per axis xyz
constexpr float KC_INF = __builtin_huge_valf();
constexpr float KC_REL = 1.00001f;

__device__ __forceinline__ int kc_cell(float x, float lo, float inv_h, float gm1) {
  return (int)fminf(fmaxf(__fmul_rn(__fsub_rn(x, lo), inv_h), 0.f), gm1);
}

// One blob per cloud (every part 16-byte aligned, so a kNN CTA stages it with ONE bulk copy):
//   [0, 64) grid parameters | float4 cl[n] | uint16 cstart[nc+1 (padded to 8)] | uint16 ipos[n (padded to 8)]
constexpr int KC_HDR = 64;
__host__ __device__ inline size_t kc_cs_off(int n) { return KC_HDR + (size_t)n * 16; }
__host__ __device__ inline size_t kc_ip_off(int n, int nc) { return kc_cs_off(n) + (size_t)((nc + 1 + 7) & ~7) * 2; }
__host__ __device__ inline size_t kc_blob_bytes(int n, int nc) { return kc_ip_off(n, nc) + (size_t)((n + 7) & ~7) * 2; }


// Shared-memory accesses of the hot loop by 32-bit shared address (the generic -> shared window arithmetic the
// compiler would otherwise redo per access costs four uniform-datapath instructions per candidate).
__device__ __forceinline__ float4 kc_lds128(unsigned a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ unsigned kc_lds16(unsigned a) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ unsigned kc_lds32(unsigned a) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void kc_sts32(unsigned a, unsigned v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void kc_sts16(unsigned a, unsigned v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void kc_sts64(unsigned a, float d, int i) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(__float_as_uint(d)), "r"(i) : "memory");
}
// sqrt for search radii: rounded by at most 2 ulp (MUFU.RSQ), always used with the KC_REL inflation
__device__ __forceinline__ float kc_sqrt(float x) { return x * rsqrtf(fmaxf(x, 1e-30f)); }


// Stages one blob (contiguous, 16-byte sized and aligned) into shared memory with a single TMA bulk copy; `bar` is a
// CTA-shared mbarrier.  kc_stage_issue is called by every thread (thread 0 issues), kc_stage_wait after whatever
// independent work should overlap the copy; it contains the __syncthreads that publishes the barrier's init.
__device__ __forceinline__ void kc_stage_issue(unsigned long long* bar_, unsigned char* dst, const unsigned char* src,
                                               unsigned bytes) {
  const unsigned bar = (unsigned)__cvta_generic_to_shared(bar_);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the init must be visible to the async proxy
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"(bar) : "memory");
  }
}
__device__ __forceinline__ void kc_stage_wait(unsigned long long* bar_) {
  const unsigned bar = (unsigned)__cvta_generic_to_shared(bar_);
  __syncthreads();  // everybody sees the initialised barrier before polling it
  unsigned done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar) : "memory");
  } while (!done);
}

}  // namespace geoa3
