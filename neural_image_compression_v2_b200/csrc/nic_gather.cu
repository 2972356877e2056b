// nic_gather.cu — K1, the tile-staged decoder-input kernel for the common 2-D shape (C = 12, PE = 6, step <= 1,
// runs of 128 texels along the fast block axis): X [N, 73] in fp32 (bit-exact with the reference), f16 or bf16
// (each value rounded once from the fp32 value).  Reference: create_decoder_input_2d / finally_decode_input_2d
// (Projects/image_compression.py:71-100, 170-181) over fp_def.create_g0_g1 (Projects/fp_def.py:115-145).
//
// The kernel is HBM-WRITE-bound: 73 * sizeof(X) bytes per texel leave the SM, ~3.7 bytes of grid come in
// (295.8 B/texel for fp32 X, 149.8 B/texel for 16-bit X).  Design:
//   * a tile = 128 consecutive samples n (one ix, 128 consecutive iy) = ONE contiguous span of 128 * 73 elements of X;
//     it is assembled in shared memory and leaves with a single TMA bulk store (cp.async.bulk.global.shared::cta),
//     double-buffered so the store of tile i overlaps the build of tile i+1;
//   * the grid nodes a tile touches (2 x <=130 nodes of G0, 2 x <=67 of G1) are staged once into shared memory in
//     channel-LAST order, so a corner is three conflict-free LDS.128 instead of 12 strided global loads;
//   * thread = texel; texel parity is warp-uniform (warp w owns texels 2*lane + (w & 1) + 64*(w >> 1)), so the
//     16-bit row (146 bytes: rows alternate 4-byte alignment) is written with aligned 32-bit shared stores whose
//     pairing depends only on the warp, and lanes stride 73 words (== 9 mod 32): bank-conflict free.
// Everything else (3-D, other C / PE, step > 1, short rows) takes the flat kernel in nic_f32.cu.
#include "nic_internal.cuh"

namespace nic {

constexpr int GT_T = 128;        // texels per tile
constexpr int GT_C = 12, GT_PE = 6, GT_CIN = 73;

__device__ __forceinline__ uint32_t gt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <typename OutT> struct Pack16;
template <> struct Pack16<__half> {
  static __device__ __forceinline__ uint32_t two(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ uint16_t one(float a) {
    __half h = __float2half_rn(a);
    return *reinterpret_cast<uint16_t*>(&h);
  }
};
template <> struct Pack16<__nv_bfloat16> {
  static __device__ __forceinline__ uint32_t two(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ uint16_t one(float a) {
    __nv_bfloat16 h = __float2bfloat16_rn(a);
    return *reinterpret_cast<uint16_t*>(&h);
  }
};

// Row of texel `t` (73 fp32 values in registers) -> staging buffer.
template <typename OutT>
__device__ __forceinline__ void write_row(OutT* stage, int t, int par, const float* v);

template <>
__device__ __forceinline__ void write_row<float>(float* stage, int t, int, const float* v) {
  float* row = stage + t * GT_CIN;
#pragma unroll
  for (int c = 0; c < GT_CIN; ++c) row[c] = v[c];
}

template <typename OutT>
__device__ __forceinline__ void write_row16(OutT* stage, int t, int par, const float* v) {
  using P = Pack16<OutT>;
  uint32_t* words = reinterpret_cast<uint32_t*>(stage) + (t >> 1) * GT_CIN;     // texel pair (2m, 2m+1) = 73 words
  if (par == 0) {
#pragma unroll
    for (int k = 0; k < 36; ++k) words[k] = P::two(v[2 * k], v[2 * k + 1]);
    reinterpret_cast<uint16_t*>(words + 36)[0] = P::one(v[72]);
  } else {
    reinterpret_cast<uint16_t*>(words + 36)[1] = P::one(v[0]);
#pragma unroll
    for (int k = 0; k < 36; ++k) words[37 + k] = P::two(v[2 * k + 1], v[2 * k + 2]);
  }
}
template <>
__device__ __forceinline__ void write_row<__half>(__half* stage, int t, int par, const float* v) { write_row16(stage, t, par, v); }
template <>
__device__ __forceinline__ void write_row<__nv_bfloat16>(__nv_bfloat16* stage, int t, int par, const float* v) {
  write_row16(stage, t, par, v);
}

// Triangular encoding rows of coordinate u for PE = 6 (utils.py:211-223):
// [tri(u/4, 0), tri(u/4, .5), tri(u/2, 0), tri(u/2, .5), tri(u, 0), 0]; the divisions are exact (powers of two).
__device__ __forceinline__ void pe6_triangular(float u, float* out) {
  const float u2 = __fmul_rn(u, 0.5f), u4 = __fmul_rn(u, 0.25f);
  out[0] = tri_wave(u4, 0.0f);
  out[1] = tri_wave(u4, 0.5f);
  out[2] = tri_wave(u2, 0.0f);
  out[3] = tri_wave(u2, 0.5f);
  out[4] = tri_wave(u, 0.0f);
  out[5] = 0.0f;
}

template <typename OutT>
__global__ void __launch_bounds__(GT_T) gather_tile_kernel(DevGeom g, const float* __restrict__ g0,
                                                           const float* __restrict__ g1,
                                                           const long long* __restrict__ origins, OutT* __restrict__ x,
                                                           int np0, int np1, unsigned ntiles, unsigned tiles_per_row) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int STAGE_BYTES = GT_T * GT_CIN * (int)sizeof(OutT);
  float* patch0 = reinterpret_cast<float*>(smem_raw + 2 * STAGE_BYTES);      // [2][np0][C]
  float* patch1 = patch0 + 2 * np0 * GT_C;                                  // [2][np1][C]
  float* pex = patch1 + 2 * np1 * GT_C;                                     // [8]: encoding of the tile's x coordinate

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // 16-bit rows: warp-uniform texel parity (see header); fp32 rows: thread = texel, lanes stride 73 words (== 9 mod 32)
  const int par = sizeof(OutT) == 2 ? (warp & 1) : 0;
  const int t = sizeof(OutT) == 2 ? 64 * (warp >> 1) + 2 * lane + par : tid;
  const long long ps0 = (long long)g.n0[0] * g.n0[1], ps1 = (long long)g.n1[0] * g.n1[1];
  const unsigned rows_per_block = (unsigned)g.B[0] * tiles_per_row;

  // Patch staging: thread -> one (channel, x-node) column of the patch (24 columns), walking the y nodes with stride 5
  // (threads 120..127 idle here).  No runtime division; the first PRE loads per grid are PREFETCHED into registers one
  // tile ahead (the whole patch at step 1/4: 34 and 18 nodes), hiding their latency behind the previous tile's build.
  constexpr int PRE0 = 7, PRE1 = 4, YS = 5;
  const int combo = tid % 24, yn0 = tid / 24, pc = combo >> 1, pxn = combo & 1;
  const bool stager = tid < 24 * YS;
  float pre0[PRE0], pre1[PRE1];
  auto tile_coords = [&](unsigned tile, AxisCoord& ax, AxisCoord& ay_first, int& py0) {
    const unsigned b = tile / rows_per_block, r = tile - b * rows_per_block;
    const unsigned ix = r / tiles_per_row, iy0 = (r - ix * tiles_per_row) * GT_T;
    int ox, oy;
    if (origins) {
      ox = (int)origins[2 * (long long)b];
      oy = (int)origins[2 * (long long)b + 1];
    } else {
      ox = g.origin0[0];
      oy = g.origin0[1];
    }
    py0 = oy + (int)iy0;
    ax = axis_coord(ox + (int)ix, g.step);
    ay_first = axis_coord(py0, g.step);
  };
  auto load0 = [&](int yn, const AxisCoord& ax, const AxisCoord& ayf) -> float {
    const int gx = clampi(ax.i0 + pxn, 0, g.n0[0] - 1), gy = clampi(ayf.i0 + yn, 0, g.n0[1] - 1);
    return __ldg(g0 + pc * ps0 + (long long)gy * g.n0[0] + gx);
  };
  auto load1 = [&](int yn, const AxisCoord& ax, const AxisCoord& ayf) -> float {
    const int gx = clampi(ax.i1 + pxn, 0, g.n1[0] - 1), gy = clampi(ayf.i1 + yn, 0, g.n1[1] - 1);
    return __ldg(g1 + pc * ps1 + (long long)gy * g.n1[0] + gx);
  };
  auto prefetch = [&](unsigned tile) {
    if (tile >= ntiles || !stager) return;
    AxisCoord ax, ayf;
    int py0;
    tile_coords(tile, ax, ayf, py0);
#pragma unroll
    for (int k = 0; k < PRE0; ++k)
      if (yn0 + k * YS < np0) pre0[k] = load0(yn0 + k * YS, ax, ayf);
#pragma unroll
    for (int k = 0; k < PRE1; ++k)
      if (yn0 + k * YS < np1) pre1[k] = load1(yn0 + k * YS, ax, ayf);
  };

  prefetch(blockIdx.x);
  unsigned it = 0;
  for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    OutT* stage = reinterpret_cast<OutT*>(smem_raw + (it & 1) * STAGE_BYTES);
    AxisCoord ax, ay_first;
    int py0;
    tile_coords(tile, ax, ay_first, py0);
    // the bulk store that used this staging buffer two tiles ago must have finished READING it; every thread must be
    // done with the previous tile's patches
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncthreads();
    // ---- stage the grid nodes of this tile, channel-last: patch[(xn * np + yn) * C + c]
    if (stager) {
      float* col0 = patch0 + pxn * np0 * GT_C + pc;
      float* col1 = patch1 + pxn * np1 * GT_C + pc;
#pragma unroll
      for (int k = 0; k < PRE0; ++k)
        if (yn0 + k * YS < np0) col0[(yn0 + k * YS) * GT_C] = pre0[k];
#pragma unroll
      for (int k = 0; k < PRE1; ++k)
        if (yn0 + k * YS < np1) col1[(yn0 + k * YS) * GT_C] = pre1[k];
      for (int yb = yn0 + PRE0 * YS; yb < np0; yb += 8 * YS) {       // larger patches (step > 1/4): unrolled chunks
        float tmp[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (yb + k * YS < np0) tmp[k] = load0(yb + k * YS, ax, ay_first);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (yb + k * YS < np0) col0[(yb + k * YS) * GT_C] = tmp[k];
      }
      for (int yb = yn0 + PRE1 * YS; yb < np1; yb += 8 * YS) {
        float tmp[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (yb + k * YS < np1) tmp[k] = load1(yb + k * YS, ax, ay_first);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (yb + k * YS < np1) col1[(yb + k * YS) * GT_C] = tmp[k];
      }
    } else if (tid < 24 * YS + GT_PE) {
      pex[tid - 24 * YS] = pe_value(g, ax.u1, tid - 24 * YS);     // the tile's x encoding, shared by its 128 texels
    }
    __syncthreads();
    prefetch(tile + gridDim.x);        // next tile's nodes: in flight while this tile's rows are built
    // ---- this thread's row, in reference column order (image_compression.py:94-96)
    {
      const AxisCoord ay = axis_coord(py0 + t, g.step);
      const int y0 = ay.i0 - ay_first.i0, y1 = ay.i1 - ay_first.i1;
      float v[GT_CIN];
      // G0 corners (dy,dx) = (0,0) (1,0) (0,1) (1,1): raw copies
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int dy = j & 1, dx = j >> 1;
        const float4* node = reinterpret_cast<const float4*>(patch0 + (dx * np0 + y0 + dy) * GT_C);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          float4 f = node[q];
          v[12 * j + 4 * q] = f.x;
          v[12 * j + 4 * q + 1] = f.y;
          v[12 * j + 4 * q + 2] = f.z;
          v[12 * j + 4 * q + 3] = f.w;
        }
      }
      // G1: ((g*wx)*wy) per corner, summed ((c0+c1)+c2)+c3 (fp_def.py:136-144, image_compression.py:95)
      {
        const float kx = ax.k, ky = ay.k;
        const float wx0 = __fsub_rn(1.0f, kx), wy0 = __fsub_rn(1.0f, ky);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int dy = j & 1, dx = j >> 1;
          const float wx = dx ? kx : wx0, wy = dy ? ky : wy0;
          const float4* node = reinterpret_cast<const float4*>(patch1 + (dx * np1 + y1 + dy) * GT_C);
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const float4 f = node[q];
            float e0 = f.x, e1 = f.y, e2 = f.z, e3 = f.w;
            if (g.interp) {
              e0 = __fmul_rn(__fmul_rn(e0, wx), wy);
              e1 = __fmul_rn(__fmul_rn(e1, wx), wy);
              e2 = __fmul_rn(__fmul_rn(e2, wx), wy);
              e3 = __fmul_rn(__fmul_rn(e3, wx), wy);
            }
            v[48 + 4 * q] = j == 0 ? e0 : __fadd_rn(v[48 + 4 * q], e0);
            v[48 + 4 * q + 1] = j == 0 ? e1 : __fadd_rn(v[48 + 4 * q + 1], e1);
            v[48 + 4 * q + 2] = j == 0 ? e2 : __fadd_rn(v[48 + 4 * q + 2], e2);
            v[48 + 4 * q + 3] = j == 0 ? e3 : __fadd_rn(v[48 + 4 * q + 3], e3);
          }
        }
      }
      // positional encodings of u1 along x (shared) then y, then the LOD column
#pragma unroll
      for (int rr = 0; rr < GT_PE; ++rr) v[60 + rr] = pex[rr];
      if (g.pe_kind == NIC_PE_TRIANGULAR) {
        pe6_triangular(ay.u1, v + 66);
      } else {
#pragma unroll
        for (int rr = 0; rr < GT_PE; ++rr) v[66 + rr] = pe_sinusoidal(ay.u1, rr, g.pe_div);
      }
      v[72] = g.lod;
      write_row<OutT>(stage, t, par, v);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the bulk copy
    __syncthreads();
    if (tid == 0) {
      OutT* dst = x + (size_t)tile * GT_T * GT_CIN;                    // tiles are consecutive spans of X
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(gt_smem_u32(stage)),
                   "r"(STAGE_BYTES)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

bool gather_tile_eligible(const DevGeom& g, const void* x) {
  return g.method == NIC_METHOD_2D && g.C == GT_C && g.PE == GT_PE && g.step <= 1.0f && g.step >= 1.0f / 1024.0f &&
         g.B[1] > 0 && g.B[1] % GT_T == 0 && g.N > 0 && g.N < (1ll << 31) && (((uintptr_t)x) & 15) == 0;
}

template <typename OutT>
static int launch_gather_tile_t(Handle* h, const DevGeom& g, const float* g0, const float* g1, const long long* origins,
                                OutT* x, cudaStream_t st) {
  // nodes a run of T texels can touch along y: floor((T-1)*step) + 2 cells' corners, +1 for an unaligned start
  const int np0 = (int)floorf((GT_T - 1) * g.step) + 3, np1 = (int)floorf((GT_T - 1) * g.step * 0.5f) + 3;
  const size_t smem = 2 * (size_t)GT_T * GT_CIN * sizeof(OutT) + ((size_t)2 * (np0 + np1) * GT_C + 8) * sizeof(float);
  auto kern = gather_tile_kernel<OutT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const unsigned tiles_per_row = (unsigned)(g.B[1] / GT_T);
  const unsigned ntiles = (unsigned)(g.N / GT_T);
  int per_sm = (int)((200 * 1024) / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
  long long cap = (long long)h->sms * per_sm;
  int grid = (int)(ntiles < cap ? ntiles : cap);
  {
    KernelTimer timer(h, st);
    kern<<<grid, GT_T, smem, st>>>(g, g0, g1, origins, x, np0, np1, ntiles, tiles_per_row);
  }
  h->launches++;
  return (int)cudaGetLastError();
}

int launch_gather_tile(Handle* h, const DevGeom& g, const float* g0, const float* g1, const long long* origins, void* x,
                       int x_dtype, cudaStream_t st) {
  switch (x_dtype) {
    case NIC_DT_F32: return launch_gather_tile_t<float>(h, g, g0, g1, origins, (float*)x, st);
    case NIC_DT_F16: return launch_gather_tile_t<__half>(h, g, g0, g1, origins, (__half*)x, st);
    case NIC_DT_BF16: return launch_gather_tile_t<__nv_bfloat16>(h, g, g0, g1, origins, (__nv_bfloat16*)x, st);
  }
  return NIC_ERR_ARG;
}

}  // namespace nic
