// Microbenchmark: cycles per tcgen05.mma for the shapes the decode / training kernels use (one CTA per SM).
// NOTE: this first version issues the MMAs under `if (threadIdx.x == 0)`, which makes the compiler wrap every mma in a
// waterfall loop (ELECT / R2UR.BROADCAST / BRA.U.ANY): the ~95 cycles per MMA it reports for every N <= 128 are an ISSUE
// cost of that code shape, not a tensor-pipe floor.  mma_rate2.cu measures both shapes side by side (48 cycles SS /
// 32 cycles TS for M128 N64 K16 when issued from uniform control flow).  Kept for the record.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../neural_image_compression_v2_b200/csrc -o mma_rate mma_rate.cu
#include <cstdio>
#include "nic_tc_common.cuh"
using namespace nic;

__global__ void __launch_bounds__(128, 1) k(int mode, int iters, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t a = smem_u32(smem), b = a + 16384;
  if (threadIdx.x == 0) {
    int M = 128, N = 64, ts = 0, ndst = 1;
    if (mode == 1) ts = 1;
    if (mode == 2) N = 256;
    if (mode == 3) { M = 64; N = 256; }
    if (mode == 4) N = 16;
    if (mode == 5) { ts = 1; ndst = 4; }        // TS, rotating over 4 accumulators (independent chains)
    if (mode == 6) { ndst = 4; }                // SS, rotating over 4 accumulators
    if (mode == 7) { N = 128; }
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t da = make_smem_desc(a, 128, 256), db = make_smem_desc(b, (N / 8) * 128, 128);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tmem + (ndst > 1 ? (i & 3) * 64 : 0);
      if (ts) mma_ts(d, tmem + 256, db, idesc, 1);
      else mma_ss(d, da, db, idesc, 1);
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const char* names[] = {"SS M128 N64 K16", "TS M128 N64 K16", "SS M128 N256 K16", "SS M64 N256 K16", "SS M128 N16 K16",
                         "TS M128 N64 x4 accumulators", "SS M128 N64 x4 accumulators", "SS M128 N128 K16"};
  for (int mode = 0; mode < 8; ++mode) {
    const int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) k<<<148, 128, 64 * 1024>>>(mode, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    printf("%-32s %7.1f cycles/MMA  (%s)\n", names[mode], avg / 148 / iters, cudaGetErrorString(e));
  }
  return 0;
}
