// compile-only: which source pattern gives a UTCHMMA without a waterfall (BRA.U.ANY) loop?
#include "nic_tc_common.cuh"
using namespace nic;
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
extern "C" __global__ void __launch_bounds__(256) patA(uint32_t* slot_g, int iters) {   // current style: tid == 0
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  __syncthreads();
  const uint32_t tmem = slot, a = smem_u32(smem);
  if (threadIdx.x == 0)
    for (int i = 0; i < iters; ++i) mma_ss(tmem, make_smem_desc(a + i * 256, 128, 256), make_smem_desc(a + 16384, 1024, 128), make_idesc(0, 128, 64), 1);
}
extern "C" __global__ void __launch_bounds__(256) patB(uint32_t* slot_g, int iters) {   // uniform warp branch + elect
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  __syncthreads();
  const uint32_t tmem = slot, a = smem_u32(smem);
  const int warp_u = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (warp_u == 0) {
    for (int i = 0; i < iters; ++i)
      if (elect_one()) mma_ss(tmem, make_smem_desc(a + i * 256, 128, 256), make_smem_desc(a + 16384, 1024, 128), make_idesc(0, 128, 64), 1);
  }
}
extern "C" __global__ void __launch_bounds__(256) patC(uint32_t* slot_g, int iters) {   // warp-dependent operands via shfl
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  __syncthreads();
  const int warp_u = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t tmem = slot + warp_u * 64, a = smem_u32(smem) + warp_u * 4096;
  if ((warp_u & 3) == 0) {
    for (int i = 0; i < iters; ++i)
      if (elect_one()) mma_ss(tmem, make_smem_desc(a + i * 256, 128, 256), make_smem_desc(a + 16384, 1024, 128), make_idesc(0, 128, 64), 1);
  }
}
extern "C" __global__ void __launch_bounds__(256) patD(uint32_t* slot_g, int iters) {   // elect once around the batch
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  __syncthreads();
  const int warp_u = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t tmem = slot + warp_u * 64, a = smem_u32(smem) + warp_u * 4096;
  if ((warp_u & 3) == 0) {
    if (elect_one()) {
      for (int i = 0; i < iters; ++i)
        mma_ss(tmem, make_smem_desc(a + i * 256, 128, 256), make_smem_desc(a + 16384, 1024, 128), make_idesc(0, 128, 64), 1);
    }
  }
}
