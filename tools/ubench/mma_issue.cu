// Microbenchmark: cycles per tcgen05.mma (kind::f16, M = 128, K = 16) issued back to back by one thread, as a function of N
// and of the A operand's source (shared-memory descriptor or TMEM).  Operand contents are irrelevant (uninitialised smem).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../semantic-abstraction_b200/csrc -o mma_issue mma_issue.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace sb;
template <int N, int TS>
__global__ void k(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    const uint32_t leader = elect_one() ? 1u : 0u;
    const uint32_t sbase = smem_u32(smem);
    const uint64_t dA = make_smem_desc(sbase, 16, 1024, SW_128B), dB = make_smem_desc(sbase + 32768, 16, 1024, SW_128B);
    constexpr uint32_t idesc = make_idesc_f16(128, N);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (TS) umma_f16_ts_elect(tm + 256, tm + uint32_t(8 * kk), dB + uint64_t(2 * kk), idesc, 1, leader);
        else umma_f16_elect(tm + 256, dA + uint64_t(2 * kk), dB + uint64_t(2 * kk), idesc, 1, leader);
      }
    }
    const long long t1 = clock64();
    umma_commit_elect(&bar, leader);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (leader) out[0] = t1 - t0, out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}
template <int N, int TS>
void run() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2000;
  k<N, TS><<<1, 64, 100 * 1024>>>(10, d);
  k<N, TS><<<1, 64, 100 * 1024>>>(iters, d);
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("N %3d A from %s: issue %6.1f cycles / MMA, issue + drain %6.1f cycles / MMA  (%s)\n", N, TS ? "TMEM" : "smem", double(h[0]) / (4 * iters),
         double(h[1]) / (4 * iters), cudaGetErrorString(cudaDeviceSynchronize()));
  cudaFree(d);
}
int main() {
  run<16, 0>(); run<64, 0>(); run<80, 0>(); run<128, 0>(); run<144, 0>(); run<256, 0>();
  run<16, 1>(); run<64, 1>(); run<128, 1>(); run<256, 1>();
  return 0;
}
