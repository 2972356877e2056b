// Microbenchmark 2: is the ~95-cycle cost of a small-N tcgen05.mma an issue-side or an operand-fetch-side limit?
//   modes vary (a) the number of warps issuing concurrently, (b) the shared-memory layout (no swizzle vs 128-byte swizzle).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../neural_image_compression_v2_b200/csrc -o mma_rate2 mma_rate2.cu
#include <cstdio>
#include "nic_tc_common.cuh"
using namespace nic;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {      // K-major, SWIZZLE_128B: rows of 128 B, 8-row groups of 1 KB
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// Same measurement with the issue code the way the compiler wants it: the WHOLE warp runs the loop (warp index made
// provably uniform with a shuffle) and one elected lane issues -> UTCHMMA operands stay in uniform registers and the
// per-MMA "waterfall" loop (ELECT / R2UR.BROADCAST / BRA.U.ANY) that `if (lane == 0)` produces disappears.
__global__ void __launch_bounds__(128, 1) kclean(int nissue, int N, int swz, int ts, int ndst, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  __shared__ long long tmax[4];
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  if (threadIdx.x < 4) mbar_init(&bar[threadIdx.x], 1);
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (warp < nissue) {
    const uint32_t a = base, b = base + 65536;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t da = swz ? desc_sw128(a) : make_smem_desc(a, 128, 256);
    const uint64_t db = swz ? desc_sw128(b) : make_smem_desc(b, (N / 8) * 128, 128);
    const uint32_t dbase = tmem + warp * 128 * (N <= 128 ? 1 : 0);
    long long t0 = clock64();
    if (elect_one()) {
      if (ts) {
#pragma unroll 8
        for (int i = 0; i < iters; ++i) mma_ts(dbase + (ndst > 1 ? (i & 1) * 64 : 0), tmem + 448, db + (swz ? (uint64_t)((i & 3) * 2) : 0), idesc, 1);
      } else {
#pragma unroll 8
        for (int i = 0; i < iters; ++i)
          mma_ss(dbase + (ndst > 1 ? (i & 1) * 64 : 0), da + (swz ? (uint64_t)((i & 3) * 2) : 0), db + (swz ? (uint64_t)((i & 3) * 2) : 0), idesc, 1);
      }
      tc_commit(&bar[warp]);
    }
    mbar_wait(&bar[warp], 0);
    if ((threadIdx.x & 31) == 0) tmax[warp] = clock64() - t0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long m = 0;
    for (int w = 0; w < nissue; ++w) m = tmax[w] > m ? tmax[w] : m;
    out[blockIdx.x] = m;
  }
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

__global__ void __launch_bounds__(128, 1) k(int nissue, int N, int swz, int ts, int ndst, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  __shared__ long long tmax[4];
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  if (threadIdx.x < 4) mbar_init(&bar[threadIdx.x], 1);
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && warp < nissue) {
    const uint32_t a = base + warp * 16384 * 0, b = base + 65536;          // all issuers share the operand tiles
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t da = swz ? desc_sw128(a) : make_smem_desc(a, 128, 256);
    const uint64_t db = swz ? desc_sw128(b) : make_smem_desc(b, (N / 8) * 128, 128);
    const uint32_t dbase = tmem + warp * 128 * (N <= 128 ? 1 : 0);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = dbase + (ndst > 1 ? (i & 1) * 64 : 0);
      // step the K offset like a real K loop does (4 MMAs per 64-element K block)
      const uint64_t koff = swz ? (uint64_t)((i & 3) * 2) : 0;
      if (ts) mma_ts(d, tmem + 448, db + koff, idesc, 1);
      else mma_ss(d, da + koff, db + koff, idesc, 1);
    }
    tc_commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
    tmax[warp] = clock64() - t0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long m = 0;
    for (int w = 0; w < nissue; ++w) m = tmax[w] > m ? tmax[w] : m;
    out[blockIdx.x] = m;
  }
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Mode { const char* name; int nissue, N, swz, ts, ndst; };
  const Mode modes[] = {
      {"SS N64 noswz 1 issuer", 1, 64, 0, 0, 1},   {"SS N64 noswz 2 issuers", 2, 64, 0, 0, 1},
      {"SS N64 noswz 4 issuers", 4, 64, 0, 0, 1},  {"SS N64 sw128 1 issuer", 1, 64, 1, 0, 1},
      {"SS N64 sw128 2 accum", 1, 64, 1, 0, 2},    {"SS N64 sw128 4 issuers", 4, 64, 1, 0, 1},
      {"TS N64 sw128 1 issuer", 1, 64, 1, 1, 1},   {"SS N128 sw128 1 issuer", 1, 128, 1, 0, 1},
      {"SS N256 sw128 1 issuer", 1, 256, 1, 0, 1}, {"SS N256 noswz 1 issuer", 1, 256, 0, 0, 1},
      {"SS N16 sw128 1 issuer", 1, 16, 1, 0, 1},   {"SS N32 noswz 1 issuer", 1, 32, 0, 0, 1},
      {"TS N64 noswz 4 issuers", 4, 64, 0, 1, 1},
  };
  cudaFuncSetAttribute(kclean, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int clean = 0; clean < 2; ++clean)
  for (const Mode& m : modes) {
    const int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) {
      if (clean) kclean<<<148, 128, 200 * 1024>>>(m.nissue, m.N, m.swz, m.ts, m.ndst, iters, d);
      else k<<<148, 128, 200 * 1024>>>(m.nissue, m.N, m.swz, m.ts, m.ndst, iters, d);
    }
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    printf("%s %-28s %7.1f cycles per MMA per SM (total %d MMAs/SM)  (%s)\n", clean ? "[elect]" : "[lane0]", m.name, avg / 148 / (iters * m.nissue), iters * m.nissue,
           cudaGetErrorString(e));
  }
  return 0;
}
