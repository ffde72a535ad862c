// fps_probe.cu — where a furthest-point-sampling round of the single-warp kernel spends its cycles
// (phase timestamps with clock64; inputs to the multi-warp FPS design).  nvcc -arch=sm_100a -O3 -I.. fps_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../geoa3_b200/csrc/common.cuh"
using namespace geoa3;

template <int P>
__global__ void __launch_bounds__(32) probe(const float* __restrict__ xyz, int n, int m, int* out, long long* cyc) {
  extern __shared__ __align__(16) float s_xyz[];
  const int lane = threadIdx.x;
  const float* p = xyz + (size_t)blockIdx.x * n * 3;
  constexpr int H = P / 2, G = P / 8;
  float2 px[H], py[H], pz[H], temp[H];
  unsigned tk[P];
#pragma unroll
  for (int t = 0; t < P; ++t) {
    const int k = lane + t * 32;
    const bool ok = k < n;
    const float x = ok ? p[k * 3] : 0.f, y = ok ? p[k * 3 + 1] : 0.f, z = ok ? p[k * 3 + 2] : 0.f;
    if (ok) { s_xyz[k] = x; s_xyz[n + k] = y; s_xyz[2 * n + k] = z; }
    tk[t] = ok ? ~(unsigned)k : 0u;
    const float t0 = ok ? 1e10f : 0.f;
    if (t & 1) { px[t >> 1].y = x; py[t >> 1].y = y; pz[t >> 1].y = z; temp[t >> 1].y = t0; }
    else { px[t >> 1].x = x; py[t >> 1].x = y; pz[t >> 1].x = z; temp[t >> 1].x = t0; }
  }
  __syncwarp();
  int old = 0;
  long long acc[5] = {0, 0, 0, 0, 0};
  for (int j = 1; j < m; ++j) {
    const long long c0 = clock64();
    const float x1 = -s_xyz[old], y1 = -s_xyz[n + old], z1 = -s_xyz[2 * n + old];
    const float2 nx = make_float2(x1, x1), ny = make_float2(y1, y1), nz = make_float2(z1, z1);
    const long long c1 = clock64() + (long long)(x1 + y1 + z1 == 12345.f);
    float hg[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float mx = 0.f;
#pragma unroll
      for (int h = g * 4; h < g * 4 + 4; ++h) {
        const float2 d = dist2x2_pn2(px[h], py[h], pz[h], nx, ny, nz);
        temp[h].x = fminf(d.x, temp[h].x);
        temp[h].y = fminf(d.y, temp[h].y);
        mx = fmaxf(mx, fmaxf(temp[h].x, temp[h].y));
      }
      hg[g] = mx;
    }
    float hmax = hg[0];
#pragma unroll
    for (int g = 1; g < G; ++g) hmax = fmaxf(hmax, hg[g]);
    const long long c2 = clock64() + (long long)(hmax == 12345.f);
    const float gmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(hmax)));
    const long long c3 = clock64() + (long long)(gmax == 12345.f);
    unsigned bl = 0u;
    if (hmax == gmax) {
#pragma unroll
      for (int g = 0; g < G; ++g)
        if (hg[g] == gmax) {
#pragma unroll
          for (int h = g * 4; h < g * 4 + 4; ++h) {
            bl = max(bl, temp[h].x == gmax ? tk[2 * h] : 0u);
            bl = max(bl, temp[h].y == gmax ? tk[2 * h + 1] : 0u);
          }
        }
    }
    const long long c4 = clock64() + (long long)(bl == 0x12345u);
    const unsigned lo = __reduce_max_sync(0xffffffffu, bl);
    old = lo != 0u ? (int)((~lo) & 0xFFFFFu) : 0;
    if (lane == 0) out[blockIdx.x * m + j] = old;
    const long long c5 = clock64() + (long long)(old == 0x7654321);
    acc[0] += c1 - c0; acc[1] += c2 - c1; acc[2] += c3 - c2; acc[3] += c4 - c3; acc[4] += c5 - c4;
  }
  if (lane == 0 && blockIdx.x == 0)
    for (int i = 0; i < 5; ++i) cyc[i] = acc[i];
}

template <int P, int MODE>
__global__ void __launch_bounds__(128) probe_quad(const float* __restrict__ xyz, int n, int m, int* out, long long* cyc) {
  extern __shared__ __align__(16) float s_xyz[];
  __shared__ __align__(16) unsigned long long s_part[2][4];
  constexpr int T = 128, H = P / 2;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float* p = xyz + (size_t)blockIdx.x * n * 3;
  float2 px[H], py[H], pz[H], temp[H];
  unsigned tk[P];
#pragma unroll
  for (int t = 0; t < P; ++t) {
    const int k = tid + t * T;
    const bool ok = k < n;
    const float x = ok ? p[k * 3] : 0.f, y = ok ? p[k * 3 + 1] : 0.f, z = ok ? p[k * 3 + 2] : 0.f;
    if (ok) { s_xyz[k] = x; s_xyz[n + k] = y; s_xyz[2 * n + k] = z; }
    tk[t] = ok ? ~(unsigned)k : 0u;
    const float t0 = ok ? 1e10f : 0.f;
    if (t & 1) { px[t >> 1].y = x; py[t >> 1].y = y; pz[t >> 1].y = z; temp[t >> 1].y = t0; }
    else { px[t >> 1].x = x; py[t >> 1].x = y; pz[t >> 1].x = z; temp[t >> 1].x = t0; }
  }
  __syncthreads();
  int old = 0;
  long long acc[6] = {0, 0, 0, 0, 0, 0};
  for (int j = 1; j < m; ++j) {
    const long long c0 = clock64();
    const float x1 = -s_xyz[old], y1 = -s_xyz[n + old], z1 = -s_xyz[2 * n + old];
    const float2 nx = make_float2(x1, x1), ny = make_float2(y1, y1), nz = make_float2(z1, z1);
    const long long c1 = clock64() + (long long)(x1 + y1 + z1 == 12345.f);
    float hm[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float2 d = dist2x2_pn2(px[h], py[h], pz[h], nx, ny, nz);
      temp[h].x = fminf(d.x, temp[h].x);
      temp[h].y = fminf(d.y, temp[h].y);
      hm[h] = fmaxf(temp[h].x, temp[h].y);
    }
#pragma unroll
    for (int st = 1; st < H; st <<= 1)
#pragma unroll
      for (int h = 0; h + st < H; h += 2 * st) hm[h] = fmaxf(hm[h], hm[h + st]);
    const float hmax = hm[0];
    const long long c2 = clock64() + (long long)(hmax == 12345.f);
    const unsigned gw = __reduce_max_sync(0xffffffffu, __float_as_uint(hmax));
    unsigned kk[H];
#pragma unroll
    for (int h = 0; h < H; ++h)
      kk[h] = max(temp[h].x == hmax ? tk[2 * h] : 0u, temp[h].y == hmax ? tk[2 * h + 1] : 0u);
#pragma unroll
    for (int st = 1; st < H; st <<= 1)
#pragma unroll
      for (int h = 0; h + st < H; h += 2 * st) kk[h] = max(kk[h], kk[h + st]);
    const long long c3 = clock64() + (long long)(gw + kk[0] == 0x12345u);
    const unsigned lw = __reduce_max_sync(0xffffffffu, __float_as_uint(hmax) == gw ? kk[0] : 0u);
    const long long c4 = clock64() + (long long)(lw == 0x12345u);
    if (lane == 0) s_part[j & 1][w] = ((unsigned long long)gw << 32) | lw;
    __syncthreads();
    const long long c5 = clock64();
    const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(&s_part[j & 1][0]);
    const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(&s_part[j & 1][2]);
    const unsigned long long ab = a.x > a.y ? a.x : a.y, cd = b.x > b.y ? b.x : b.y;
    const unsigned lo = (unsigned)(ab > cd ? ab : cd);
    old = lo != 0u ? (int)((~lo) & 0xFFFFFu) : 0;
    if (MODE == 0 && tid == 0) out[blockIdx.x * m + j] = old;
    if (MODE == 1 && w == 1 && lane == 0) out[blockIdx.x * m + j] = old;   // store from a thread that is not timed
    if (MODE == 2) s_xyz[3 * n + j] = (float)old;                          // results to shared memory only
    const long long c6 = clock64() + (long long)(old == 0x7654321);
    acc[0] += c1 - c0; acc[1] += c2 - c1; acc[2] += c3 - c2; acc[3] += c4 - c3; acc[4] += c5 - c4; acc[5] += c6 - c5;
  }
  if (tid == 0 && blockIdx.x == 0)
    for (int i = 0; i < 6; ++i) cyc[i] = acc[i];
}

template <int P, int MODE> void run_quad(int n, int m, int blocks) {
  float* h = (float*)malloc((size_t)blocks * n * 3 * 4);
  srand(1);
  for (size_t i = 0; i < (size_t)blocks * n * 3; ++i) h[i] = rand() / (float)RAND_MAX - 0.5f;
  float* d; int* o; long long* c;
  cudaMalloc(&d, (size_t)blocks * n * 12); cudaMalloc(&o, (size_t)blocks * m * 4); cudaMalloc(&c, 48);
  cudaMemcpy(d, h, (size_t)blocks * n * 12, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe_quad<P, MODE><<<blocks, 128, n * 12 + m * 4>>>(d, n, m, o, c);
  cudaEventRecord(e0);
  probe_quad<P, MODE><<<blocks, 128, n * 12 + m * 4>>>(d, n, m, o, c);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long hc[6]; cudaMemcpy(hc, c, 48, cudaMemcpyDeviceToHost);
  const double r = m - 1;
  printf("{\"what\":\"fps_probe_quad\",\"mode\":%d,\"P\":%d,\"n\":%d,\"m\":%d,\"blocks\":%d,\"us\":%.1f,\"cycles_per_round\":{\"lds_old\":%.1f,\"dist_minmax\":%.1f,\"redux1_keys\":%.1f,\"redux2\":%.1f,\"sts_barrier\":%.1f,\"merge_store\":%.1f}}\n",
         MODE, P, n, m, blocks, ms * 1e3, hc[0] / r, hc[1] / r, hc[2] / r, hc[3] / r, hc[4] / r, hc[5] / r);
  free(h); cudaFree(d); cudaFree(o); cudaFree(c);
}

template <int P> void run(int n, int m, int blocks) {
  float* h = (float*)malloc((size_t)blocks * n * 3 * 4);
  srand(1);
  for (size_t i = 0; i < (size_t)blocks * n * 3; ++i) h[i] = rand() / (float)RAND_MAX - 0.5f;
  float* d; int* o; long long* c;
  cudaMalloc(&d, (size_t)blocks * n * 12); cudaMalloc(&o, (size_t)blocks * m * 4); cudaMalloc(&c, 40);
  cudaMemcpy(d, h, (size_t)blocks * n * 12, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<P><<<blocks, 32, n * 12>>>(d, n, m, o, c);
  cudaEventRecord(e0);
  probe<P><<<blocks, 32, n * 12>>>(d, n, m, o, c);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long hc[5]; cudaMemcpy(hc, c, 40, cudaMemcpyDeviceToHost);
  printf("{\"what\":\"fps_probe\",\"P\":%d,\"n\":%d,\"m\":%d,\"blocks\":%d,\"us\":%.1f,\"cycles_per_round\":{\"lds_old\":%.1f,\"dist_minmax\":%.1f,\"redux1\":%.1f,\"tie\":%.1f,\"redux2_store\":%.1f}}\n",
         P, n, m, blocks, ms * 1e3, hc[0] / (double)(m - 1), hc[1] / (double)(m - 1), hc[2] / (double)(m - 1),
         hc[3] / (double)(m - 1), hc[4] / (double)(m - 1));
  free(h); cudaFree(d); cudaFree(o); cudaFree(c);
}
int main() {
  run<32>(1024, 512, 250); run<32>(1024, 512, 1);
  run<16>(512, 128, 250);
  run_quad<8, 0>(1024, 512, 250); run_quad<8, 1>(1024, 512, 250); run_quad<8, 2>(1024, 512, 250); run_quad<4, 0>(512, 128, 250);
  return 0;
}
