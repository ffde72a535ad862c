// fp32_peak.cu — measures the non-tensor FP32 peak of the B200 (needed as a roofline denominator for
// the N x N distance kernels): dependent-free FFMA chains, scalar (FFMA) vs packed
// (FFMA2), plus a mixed FFMA2 + FMNMX stream (do the alu-pipe ops ride along for free?).
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
  float a[16];
  float2 p[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = make_float2(a[2 * i], a[2 * i + 1]);
  const float m = 1.000001f, c = 1e-7f;
  const float2 m2 = make_float2(m, m), c2 = make_float2(c, c);
  float mn = 1e30f;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = __fmaf_rn(a[i], m, c);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = __ffma2_rn(p[i], m2, c2);
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        p[i] = __ffma2_rn(p[i], m2, c2);
        mn = fminf(mn, p[i].x);
        mn = fminf(mn, p[i].y);
      }
    } else if (MODE == 3) {  // the distance mix: 3 FADD2 + FMUL2 + 2 FFMA2 per 2 pairs (12 "fma-equivalents")
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const float2 dx = __fadd2_rn(p[i], c2), dy = __fadd2_rn(p[i + 1], m2), dz = __fadd2_rn(p[i], m2);
        float2 t = __fmul2_rn(dx, dx);
        t = __ffma2_rn(dy, dy, t);
        t = __ffma2_rn(dz, dz, t);
        p[i] = t;
      }
    } else {  // distance mix + one FMNMX per pair (the seeded nn_pair hot loop)
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const float2 dx = __fadd2_rn(p[i], c2), dy = __fadd2_rn(p[i + 1], m2), dz = __fadd2_rn(p[i], m2);
        float2 t = __fmul2_rn(dx, dx);
        t = __ffma2_rn(dy, dy, t);
        t = __ffma2_rn(dz, dz, t);
        p[i] = t;
        mn = fminf(mn, fminf(t.x, t.y));
      }
    }
  }
  float s = mn;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += p[i].x + p[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(const char* name, float* d, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = 148 * 8;
  k<MODE><<<blocks, 256>>>(d, 16);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(d, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fma = (double)blocks * 256 * iters * 16;
  const double tflops = 2 * fma / (ms * 1e-3) / 1e12;
  printf("{\"ubench\":\"%s\",\"ms\":%.3f,\"fp32_tflops\":%.2f}\n", name, ms, tflops);
  return tflops;
}

int main() {
  float* d;
  cudaMalloc(&d, 148 * 8 * 256 * 4);
  for (int r = 0; r < 2; ++r) {
    run<0>("ffma_scalar", d, 1 << 16);
    run<1>("ffma2_packed", d, 1 << 16);
    run<2>("ffma2_plus_2fmnmx_per_pair", d, 1 << 16);
    run<3>("dist_mix_4x(3fadd2+fmul2+2ffma2)_per_iter[tflops_field=x16/24_of_real]", d, 1 << 16);
    run<4>("dist_mix_plus_fmnmx[same]", d, 1 << 16);
  }
  return 0;
}
