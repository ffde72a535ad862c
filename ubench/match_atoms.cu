// match_atoms.cu — cost of MATCH.ANY and shared-memory atomics (inputs to the CSR-by-target design).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512) k(unsigned* out, int iters, int distinct, long long* cyc) {
  __shared__ unsigned h[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) h[i] = 0;
  __syncthreads();
  unsigned x = threadIdx.x * 2654435761u + blockIdx.x;
  unsigned acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    x = x * 1664525u + 1013904223u;
    const int key = (x >> 8) % distinct;
    if (MODE == 0) acc += __match_any_sync(0xffffffffu, key);
    if (MODE == 1) acc += atomicAdd(&h[key], 1u);            // returning shared atomic
    if (MODE == 2) atomicAdd(&h[key], 1u);                   // result unused
    if (MODE == 3) { unsigned v = h[key]; __syncwarp(); h[key] = v + 1; __syncwarp(); acc += v; }  // plain RMW
    if (MODE == 4) acc += __reduce_max_sync(0xffffffffu, key);
    if (MODE == 5) acc += __ballot_sync(0xffffffffu, key & 1);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + h[threadIdx.x];
}
template <int MODE> void run(const char* name, unsigned* d, long long* c, int distinct, int warps) {
  const int iters = 4096;
  k<MODE><<<148, warps * 32>>>(d, iters, distinct, c);
  cudaDeviceSynchronize();
  long long hc; cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
  printf("{\"op\":\"%s\",\"distinct\":%d,\"warps_per_sm\":%d,\"cycles_per_warp_op_at_this_occupancy\":%.1f,\"sm_cycles_per_op\":%.2f}\n",
         name, distinct, warps, (double)hc / iters, (double)hc / iters / warps);
}
int main() {
  unsigned* d; long long* c;
  cudaMalloc(&d, 148 * 512 * 4); cudaMalloc(&c, 8);
  for (int warps : {1, 16}) for (int distinct : {1024, 8}) {
    run<0>("match_any", d, c, distinct, warps);
    run<1>("atoms_add_ret", d, c, distinct, warps);
    run<2>("atoms_add_noret", d, c, distinct, warps);
    run<3>("lds_sts_rmw_syncwarp", d, c, distinct, warps);
  }
  run<4>("redux_max", d, c, 1024, 1); run<4>("redux_max", d, c, 1024, 16);
  run<5>("ballot", d, c, 1024, 1); run<5>("ballot", d, c, 1024, 16);
  return 0;
}
