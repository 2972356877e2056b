// Pure-write bandwidth ceilings for K1's output pattern on one B200 (no reads, no row build):
//   plain   : coalesced 16-byte st.global, grid-stride
//   bulk<B> : cp.async.bulk.global.shared::cta of B bytes per store, two stores in flight per CTA (K1's skeleton)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/write_bw tools/ubench/write_bw.cu ; ./tools/ubench/write_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void plain_kernel(uint4* dst, size_t n16) {
  const uint4 v = make_uint4(1, 2, 3, 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}

__global__ void __launch_bounds__(128) bulk_kernel(uint8_t* dst, unsigned nchunks, unsigned bytes) {
  extern __shared__ __align__(128) uint8_t smem[];
  for (unsigned i = threadIdx.x; i < 2 * bytes / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  unsigned n = 0;
  for (unsigned c = blockIdx.x; c < nchunks; c += gridDim.x, ++n) {
    if (threadIdx.x == 0) {
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem + (n & 1) * bytes);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (size_t)c * bytes), "r"(s), "r"(bytes)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  const size_t total = (size_t)4096 * 4096 * 146;          // K1's f16 X for a 4096^2 frame: 2.45 GB
  uint8_t* d;
  cudaMalloc(&d, total);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  auto time = [&](auto launch, const char* name) {
    for (int i = 0; i < 2; ++i) launch();
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    printf("WRITE_BW %-28s %.3f ms  %.0f GB/s  (%s)\n", name, ms / 5, total / (ms / 5 * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
  };
  time([&] { plain_kernel<<<148 * 8, 256>>>((uint4*)d, total / 16); }, "plain 16-byte stores");
  for (unsigned bytes : {18688u, 37376u}) {
    for (int per_sm : {2, 4, 6}) {
      if (2 * bytes * per_sm > 220 * 1024) continue;
      cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * bytes);
      char name[64];
      snprintf(name, sizeof(name), "bulk %u B, %d CTAs/SM", bytes, per_sm);
      time([&] { bulk_kernel<<<148 * per_sm, 128, 2 * bytes>>>(d, (unsigned)(total / bytes), bytes); }, name);
    }
  }
  return 0;
}
