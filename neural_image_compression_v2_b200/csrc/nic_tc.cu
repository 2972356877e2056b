// nic_tc.cu — K2, the fused tensor-core decode: gather -> Linear/GELU/Linear/GELU/Linear/Sigmoid -> output,
// with all three layer GEMMs on tcgen05.mma (sm_100a) and every activation resident in TMEM / registers.
// Reference behaviour: finally_decode_input_* + ColorDecoder.forward inside decode_image
// (Projects/image_compression.py:54-68, 170-211, 313-327).
//
// Shape of the computation per CTA (128 threads, one thread per texel row, 4 CTAs resident per SM so one
// CTA's epilogue overlaps another's MMA):
//   tile of 128 texels:
//     gather   : each thread builds its decoder-input row (fp32), packs it to 16-bit pairs and writes it with
//                tcgen05.st into TMEM columns [A, A+KX/2)  -> operand A of layer 1 (A-from-TMEM, "TS" MMA).
//     layer 1  : KX/16 x tcgen05.mma (M=128, N=64, K=16)  A = TMEM, B = W1' in shared memory, D1 -> TMEM.
//     epilogue : tcgen05.ld D1 -> packed tanh-GELU (x + x*tanh(u); the 1/2 is folded into W2') -> tcgen05.st H1.
//     layer 2  : 5 x tcgen05.mma, A = H1 (TMEM, K = 64 + bias column block), B = W2'.
//     epilogue : same -> H2.
//     layer 3  : 5 x tcgen05.mma (N = 16), B = W3'.   epilogue: sigmoid, quantise, store.
//   Biases ride in the K dimension: the row carries a constant 1 and the matching column of W' holds the bias
//   (for layer 1 that column also absorbs the LOD input, which is constant for a launch).
//   TMEM map (128 columns per CTA): [0,64) accumulator D (fp32), [64,64+KX/2) operand A (16-bit pairs).
//
// Algorithmic work: 2*(Cin*64 + 64*64 + 64*Cout) FLOP/texel (17,920 for the 2-D default); tensor-bound roofline.
#include "nic_tc_common.cuh"

namespace nic {

// ------------------------------------------------------------------------------------------------ weight images
// Operand-B images in global memory, already in the shared-memory UMMA layout (K-major, no swizzle):
//   byte offset of element (n, k) = (k/8)*LBO + (n/8)*128 + (n%8)*16 + (k%8)*2,  LBO = (Nrows/8)*128.
// W1': [64 x KX]  col k < cin-1: W1[n][k];  col cin-1: b1[n] + lod*W1[n][cin-1];  rest 0.
// W2': [64 x 80]  col k < 64: W2[n][k]/2;   col 64: b2[n];  rest 0.
// W3': [16 x 80]  row n < cout: col k < 64: W3[n][k]/2; col 64: b3[n];  rest 0.
template <int FMT>
__global__ void pack_weights_kernel(MlpDev m, float lod, int KX, uint16_t* __restrict__ img) {
  const int H = 64;
  const int n1 = H * KX, n2 = H * 80, n3 = 16 * 80;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2 + n3; i += gridDim.x * blockDim.x) {
    int which = i < n1 ? 0 : (i < n1 + n2 ? 1 : 2);
    int local = which == 0 ? i : (which == 1 ? i - n1 : i - n1 - n2);
    int nrows = which == 2 ? 16 : H;
    // decode the position inside the image: [k/8][n/8][n%8][k%8]
    int kc = local / (nrows * 8);
    int rem = local - kc * nrows * 8;
    int n = rem / 8, ke = rem - n * 8;
    int k = kc * 8 + ke;
    float v = 0.f;
    if (which == 0) {
      if (k < m.cin - 1) v = m.w1[n * m.cin + k];
      else if (k == m.cin - 1) v = m.b1[n] + lod * m.w1[n * m.cin + k];
    } else if (which == 1) {
      if (k < H) v = 0.5f * m.w2[n * H + k];
      else if (k == H) v = m.b2[n];
    } else if (n < m.cout) {
      if (k < H) v = 0.5f * m.w3[n * H + k];
      else if (k == H) v = m.b3[n];
    }
    img[i] = to16<FMT>(v);
  }
}

// ------------------------------------------------------------------------------------------------ gather to registers
// Decoder-input row of one texel for C = 12, PE = 6, built directly as packed 16-bit pairs (the operand-A registers):
//   pairs [6j, 6j+6)         corner j of G0 (raw copy of the shadow node)
//   pairs [6*NC0, 6*NC0+6)   sum_j w_j * G1 corner j   (packed fma; w_j exact in f16)
//   then 3 pairs per axis    positional encoding (shared-memory LUT for the triangular kind)
//   half CIN-1               constant 1 carrying the layer-1 bias (+ LOD), remaining halves 0.
template <int METHOD>
struct RowShape {
  static constexpr int DIM = METHOD == NIC_METHOD_2D ? 2 : 3;
  static constexpr int NC0 = METHOD == NIC_METHOD_3D ? 8 : 4;
  static constexpr int NC1 = METHOD == NIC_METHOD_2D ? 4 : 8;
  static constexpr int C = 12, PE = 6;
  static constexpr int CIN = C * (NC0 + 1) + PE * DIM + 1;    // 73 / 127 / 79
  static constexpr int KX = (CIN + 15) / 16 * 16;            // 80 / 128 / 80
};

struct ShadowGeom {
  const uint2* s0;   // shadow of G0, 3 x uint2 per node
  const uint2* s1;   // shadow of G1
};

__device__ __forceinline__ int node_lin(const int* n, int dim, int x, int y, int z) {
  // shadow order: x slowest, then y, then z (2-D: x, y)
  return dim == 2 ? x * n[1] + y : (x * n[1] + y) * n[2] + z;
}

template <int METHOD, int FMT>
__device__ __forceinline__ void gather_row_packed(const DevGeom& g, const ShadowGeom& sg, const int* p,
                                                  const uint4* __restrict__ pe_lut, int lut_mask, uint32_t* xp) {
  using S = RowShape<METHOD>;
  using P = Pair<FMT>;
  AxisCoord ax[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    ax[a] = axis_coord(p[a], g.step);
    // memory safety for device-supplied origins (host-known origins are validated): keep i, i+1 inside the grid
    ax[a].i0 = clampi(ax[a].i0, 0, g.n0[a] - 2 < 0 ? 0 : g.n0[a] - 2);
    ax[a].i1 = clampi(ax[a].i1, 0, g.n1[a] - 2 < 0 ? 0 : g.n1[a] - 2);
  }
  // ---- G0 corners: raw copies
#pragma unroll
  for (int j = 0; j < S::NC0; ++j) {
    const int8_t* d = S::DIM == 2 ? kCorner2D[j] : (METHOD == NIC_METHOD_3D ? kCorner3D[j] : kCorner3Dv2[j]);
    int dz = S::DIM == 3 ? d[0] : 0;
    const uint2* node = sg.s0 + 3 * node_lin(g.n0, S::DIM, ax[0].i0 + d[2], ax[1].i0 + d[1], ax[2].i0 + dz);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      uint2 v = __ldg(node + q);
      xp[6 * j + 2 * q] = v.x;
      xp[6 * j + 2 * q + 1] = v.y;
    }
  }
  // ---- G1: weighted sum of corners
  typename P::T2 acc[6];
#pragma unroll
  for (int j = 0; j < S::NC1; ++j) {
    const int8_t* d = S::DIM == 2 ? kCorner2D[j] : kCorner3D[j];
    int dz = S::DIM == 3 ? d[0] : 0;
    const uint2* node = sg.s1 + 3 * node_lin(g.n1, S::DIM, ax[0].i1 + d[2], ax[1].i1 + d[1], ax[2].i1 + dz);
    float f[3];
    g1_factors(g, j, ax, f);
    float w = f[0] * f[1];
    if (S::DIM == 3) w *= f[2];
    typename P::T2 w2 = P::cst(w);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      uint2 v = __ldg(node + q);
      typename P::T2 a = *reinterpret_cast<typename P::T2*>(&v.x), b = *reinterpret_cast<typename P::T2*>(&v.y);
      acc[2 * q] = j == 0 ? __hmul2(a, w2) : __hfma2(a, w2, acc[2 * q]);
      acc[2 * q + 1] = j == 0 ? __hmul2(b, w2) : __hfma2(b, w2, acc[2 * q + 1]);
    }
  }
#pragma unroll
  for (int q = 0; q < 6; ++q) xp[6 * S::NC0 + q] = *reinterpret_cast<uint32_t*>(&acc[q]);
  // ---- positional encoding: 3 pairs per axis
  constexpr int PE0 = 6 * (S::NC0 + 1);
#pragma unroll
  for (int a = 0; a < S::DIM; ++a) {
    if (g.pe_kind == NIC_PE_TRIANGULAR) {
      uint4 e = pe_lut[p[a] & lut_mask];       // the encoding is periodic in the texel coordinate (period 16/step)
      xp[PE0 + 3 * a] = e.x;
      xp[PE0 + 3 * a + 1] = e.y;
      xp[PE0 + 3 * a + 2] = e.z;
    } else {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        float arg = __fmul_rn(ax[a].u1, g.pe_div[q]);
        auto v = P::pack(sinf(arg), cosf(arg));
        xp[PE0 + 3 * a + q] = *reinterpret_cast<uint32_t*>(&v);
      }
    }
  }
  // ---- bias carrier and padding.  CIN - 1 is even for all three methods: the 1 sits in the low half of its pair.
  static_assert((S::CIN - 1) % 2 == 0, "bias column must be the low half of a pair");
  {
    auto one = P::pack(1.0f, 0.0f);
    xp[(S::CIN - 1) / 2] = *reinterpret_cast<uint32_t*>(&one);
#pragma unroll
    for (int i = (S::CIN - 1) / 2 + 1; i < S::KX / 2; ++i) xp[i] = 0u;
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
constexpr int TC_ROWS = 128;    // texels per tile = MMA M = TMEM lanes
constexpr int TC_THREADS = 256; // 8 warps: warps w and w+4 share TMEM lane quarter w and split the columns
constexpr int TC_TMEM_COLS = 128;
constexpr int TC_COL_D = 0;     // accumulator columns [0, 64)
constexpr int TC_COL_A = 64;    // operand-A columns   [64, 64 + KX/2)
constexpr int TC_K2 = 80;       // K of layers 2/3: 64 hidden + the bias block
constexpr int TC_LUT_MAX = 256; // positional-encoding LUT entries (period 16/step texels)

template <int METHOD, int FMT, typename OutT>
__global__ void __launch_bounds__(TC_THREADS, 4) decode_tc_kernel(DevGeom g, ShadowGeom sg,
                                                                  const long long* __restrict__ origins,
                                                                  const uint4* __restrict__ wimg, int cout,
                                                                  int lut_n, OutT* __restrict__ out) {
  using S = RowShape<METHOD>;
  using P = Pair<FMT>;
  constexpr int KX = S::KX;
  constexpr int HALF = KX / 4;   // operand-A registers per warp group
  constexpr int W1_BYTES = b_image_bytes(64, KX), W2_BYTES = b_image_bytes(64, TC_K2), W3_BYTES = b_image_bytes(16, TC_K2);
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* sW1 = smem_raw;
  uint8_t* sW2 = sW1 + W1_BYTES;
  uint8_t* sW3 = sW2 + W2_BYTES;
  uint4* sLut = reinterpret_cast<uint4*>(sW3 + W3_BYTES);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sLut + TC_LUT_MAX);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);

  const int tid = threadIdx.x, warp = uniform_warp_index();
  const int grp = warp >> 2;                 // 0: columns [0, 32) of D / first half of the row; 1: the rest
  const int row = tid & (TC_ROWS - 1);       // texel row of the tile = TMEM lane
  // ---- one-time setup: TMEM allocation, barrier, weight images and PE LUT -> shared memory
  if (warp == 0) tmem_alloc(tmem_slot, TC_TMEM_COLS);
  if (tid == 0) mbar_init(mbar, 1);
  {
    uint4* dst = reinterpret_cast<uint4*>(smem_raw);
    constexpr int NV = (W1_BYTES + W2_BYTES + W3_BYTES) / 16;
    for (int i = tid; i < NV; i += TC_THREADS) dst[i] = __ldg(wimg + i);
    for (int i = tid; i < lut_n; i += TC_THREADS) {
      float u1 = __fmul_rn(__fmul_rn((float)i, g.step), 0.5f);
      uint4 e;
      auto v0 = P::pack(pe_triangular(u1, 0, 6), pe_triangular(u1, 1, 6));
      auto v1 = P::pack(pe_triangular(u1, 2, 6), pe_triangular(u1, 3, 6));
      auto v2 = P::pack(pe_triangular(u1, 4, 6), pe_triangular(u1, 5, 6));
      e.x = *reinterpret_cast<uint32_t*>(&v0);
      e.y = *reinterpret_cast<uint32_t*>(&v1);
      e.z = *reinterpret_cast<uint32_t*>(&v2);
      e.w = 0u;
      sLut[i] = e;
    }
  }
  fence_async_smem();            // weights (generic-proxy stores) -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t tD = tmem + TC_COL_D, tA = tmem + TC_COL_A;

  constexpr uint32_t IDESC_64 = make_idesc(FMT, 128, 64), IDESC_16 = make_idesc(FMT, 128, 16);
  constexpr uint32_t LBO_64 = (64 / 8) * 128, LBO_16 = (16 / 8) * 128, SBO = 128;
  const uint32_t aW1 = smem_u32(sW1), aW2 = smem_u32(sW2), aW3 = smem_u32(sW3);
  uint32_t phase = 0;

  const unsigned ntiles = (unsigned)((g.N + TC_ROWS - 1) / TC_ROWS);
  for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const unsigned n = tile * TC_ROWS + row;
    const bool live = n < (unsigned)g.N;
    // ---- gather: each warp group builds and stores its half of the row (the other half is dead code per branch)
    {
      Texel t = texel_of_fast(g, live ? n : (unsigned)g.N - 1, origins);
      uint32_t xp[KX / 2];
      if (grp == 0) {
        gather_row_packed<METHOD, FMT>(g, sg, t.p, sLut, lut_n - 1, xp);
        store_row_part<HALF>(tA + lane_base, xp);
      } else {
        gather_row_packed<METHOD, FMT>(g, sg, t.p, sLut, lut_n - 1, xp);
        store_row_part<HALF>(tA + lane_base + HALF, xp + HALF);
      }
    }
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    // ---- layer 1
    if (warp == 0) {               // one warp issues and waits for the MMAs; the others sleep at the barrier below
      if (elect_one()) {
        tc_fence_after();
#pragma unroll
        for (int kc = 0; kc < KX / 16; ++kc)
          mma_ts(tD, tA + kc * 8, make_smem_desc(aW1 + kc * 2 * LBO_64, LBO_64, SBO), IDESC_64, kc > 0);
        tc_commit(mbar);
      }
      __syncwarp();
      mbar_wait(mbar, phase);
    }
    phase ^= 1;
    __syncthreads();
    tc_fence_after();
    // ---- epilogue of layers 1 and 2: D -> 2*gelu -> H (operand A of the next layer), bias block [1, 0, ...]
#pragma unroll 1
    for (int layer = 0; layer < 2; ++layer) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint32_t acc[16];
        tmem_ld16(tD + lane_base + grp * 32 + q * 16, acc);
        tc_wait_ld();
        uint32_t hp[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) hp[i] = gelu2x_pair<FMT>(__uint_as_float(acc[2 * i]), __uint_as_float(acc[2 * i + 1]));
        tmem_st8(tA + lane_base + grp * 16 + q * 8, hp);
      }
      if (layer == 0 && grp == 1) {   // the bias block survives layer 2's epilogue, which rewrites columns [0, 32) only
        uint32_t ones[8];
        auto one = P::pack(1.0f, 0.0f);
        ones[0] = *reinterpret_cast<uint32_t*>(&one);
#pragma unroll
        for (int i = 1; i < 8; ++i) ones[i] = 0u;
        tmem_st8(tA + lane_base + 32, ones);
      }
      tc_wait_st();
      tc_fence_before();
      __syncthreads();
      if (warp == 0) {
        if (elect_one()) {
          tc_fence_after();
          const uint32_t aW = layer == 0 ? aW2 : aW3, lbo = layer == 0 ? LBO_64 : LBO_16;
          const uint32_t idesc = layer == 0 ? IDESC_64 : IDESC_16;
#pragma unroll
          for (int kc = 0; kc < TC_K2 / 16; ++kc)
            mma_ts(tD, tA + kc * 8, make_smem_desc(aW + kc * 2 * lbo, lbo, SBO), idesc, kc > 0);
          tc_commit(mbar);
        }
        __syncwarp();
        mbar_wait(mbar, phase);
      }
      phase ^= 1;
      __syncthreads();
      tc_fence_after();
    }
    // ---- output (warp group 0): sigmoid, optional 8-bit quantisation.  Group 1 moves on to the next gather.
    if (grp == 0) {
      uint32_t acc[16];
      tmem_ld16(tD + lane_base, acc);
      tc_wait_ld();
      if (live) {
#pragma unroll
        for (int c = 0; c < 16; ++c)
          if (c < cout) {
            float z = __uint_as_float(acc[c]);
            float v = __fdividef(1.0f, 1.0f + __expf(-z));
            store_out(out + (size_t)n * cout + c, v);
          }
      }
    }
    // the next tile's tcgen05.st / mma reuse the A and D columns: order them after this tile's tcgen05.ld
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TC_TMEM_COLS);
}

// ================================================================================================ 2-D fast path
// Full-resolution 2-D decode (step 1/4, interpolation on, triangular PE, block aligned to 8 x 16 texels): layer 1 is
// re-associated so that NO per-texel input is ever built.  For a tile of 8 x 16 texels:
//   z1 = W1[:, 0:48] * (G0 corners of the texel's cell)          -> the 128 rows alias 8 cell rows (descriptor SBO = 0)
//      + sum_nodes tent(node, texel) * R[node]                   R[node] = W1[:, 48:60] * G1[node]   (per-node table)
//      + LUTx[px mod 64] + LUTy[py mod 64]                       LUT = W1[:, 60:72] * PE(p) (+ bias and LOD in LUTx)
// The last two lines are ONE constant 128 x 32 selector/weight matrix (tent weights, one-hot x, one-hot y) that stays
// in TMEM for the life of the CTA, times per-tile table rows that are addressed in shared memory as an MN-major B
// operand.  Row m of a tile is texel (cell = m % 8, within-cell index = m / 8).
constexpr int F_TX = 8, F_TY = 16;                 // tile extent in texels (x = first image axis)
constexpr int F_COL_SEL = 104;                     // TMEM columns [104, 120): the selector matrix (32 halves / row)
constexpr int F_W1G0 = b_image_bytes(64, 48);      // 6144
constexpr int F_LUT = 64 * 64 * 2;                 // 8192 per axis
constexpr int F_IMG = F_W1G0 + b_image_bytes(64, TC_K2) + b_image_bytes(16, TC_K2) + 2 * F_LUT;

// R[node][n] = sum_c W1[n][4C + c] * G1[c][node], stored [x][y][64] 16-bit (128 B per node).  One thread per node:
// 12 coalesced loads, 64 x 12 fma against weights broadcast from shared memory, one 128-byte row out.
template <int FMT>
__global__ void __launch_bounds__(128) g1_rows_kernel(MlpDev m, const float* __restrict__ g1, int nx, int ny,
                                                      uint16_t* __restrict__ R, int code_bits) {
  constexpr int C = 12;
  __shared__ float w[C * 64];                 // [c][n]
  for (int i = threadIdx.x; i < 64 * C; i += blockDim.x) w[(i % C) * 64 + i / C] = m.w1[(i / C) * m.cin + 4 * C + (i % C)];
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= nx) return;
  const size_t nodes = (size_t)nx * ny, node = (size_t)y * nx + x;
  float acc[64];
#pragma unroll
  for (int n = 0; n < 64; ++n) acc[n] = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float g = grid_value(g1, (long long)(c * nodes + node), code_bits);
    const float4* wr = reinterpret_cast<const float4*>(w + c * 64);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      float4 ww = wr[q];
      acc[4 * q] = fmaf(g, ww.x, acc[4 * q]);
      acc[4 * q + 1] = fmaf(g, ww.y, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(g, ww.z, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(g, ww.w, acc[4 * q + 3]);
    }
  }
  uint4* dst = reinterpret_cast<uint4*>(R + ((size_t)x * ny + y) * 64);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    uint4 v;
    v.x = to16<FMT>(acc[8 * q]) | ((uint32_t)to16<FMT>(acc[8 * q + 1]) << 16);
    v.y = to16<FMT>(acc[8 * q + 2]) | ((uint32_t)to16<FMT>(acc[8 * q + 3]) << 16);
    v.z = to16<FMT>(acc[8 * q + 4]) | ((uint32_t)to16<FMT>(acc[8 * q + 5]) << 16);
    v.w = to16<FMT>(acc[8 * q + 6]) | ((uint32_t)to16<FMT>(acc[8 * q + 7]) << 16);
    dst[q] = v;
  }
}

// Weight images of the fast path: [W1g0 K-major 64x48][W2' 64x80][W3' 16x80][LUTx][LUTy]; LUT element (n, k = p mod 64)
// at (k/8)*1024 + (n/8)*128 + (k%8)*16 + (n%8)*2  (MN-major B operand, 8 k-rows x 8 n per 128-byte block).
template <int FMT>
__global__ void pack_fast_kernel(MlpDev m, float lod, float step, uint16_t* __restrict__ img) {
  const int H = 64, C = 12, PE = 6;
  const int n1 = H * 48, n2 = H * 80, n3 = 16 * 80, nl = 64 * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2 + n3 + 2 * nl; i += gridDim.x * blockDim.x) {
    float v = 0.f;
    if (i < n1 + n2 + n3) {
      int which = i < n1 ? 0 : (i < n1 + n2 ? 1 : 2);
      int local = which == 0 ? i : (which == 1 ? i - n1 : i - n1 - n2);
      int nrows = which == 2 ? 16 : H;
      int kc = local / (nrows * 8), rem = local - kc * nrows * 8;
      int n = rem / 8, k = kc * 8 + (rem - n * 8);
      if (which == 0) v = m.w1[n * m.cin + k];
      else if (which == 1) v = k < H ? 0.5f * m.w2[n * H + k] : (k == H ? m.b2[n] : 0.f);
      else if (n < m.cout) v = k < H ? 0.5f * m.w3[n * H + k] : (k == H ? m.b3[n] : 0.f);
    } else {
      int local = i - (n1 + n2 + n3);
      int axis = local / nl;
      local -= axis * nl;
      int kg = local / 512, rem = local - kg * 512;       // 512 elements per k-group: [n/8][k%8][n%8]
      int ng = rem / 64, r2 = rem - ng * 64;
      int k = kg * 8 + r2 / 8, n = ng * 8 + (r2 & 7);
      float u1 = __fmul_rn(__fmul_rn((float)k, step), 0.5f);
      float acc = axis == 0 ? m.b1[n] + lod * m.w1[n * m.cin + (m.cin - 1)] : 0.f;
      for (int r = 0; r < PE; ++r) acc = fmaf(m.w1[n * m.cin + 5 * C + axis * PE + r], pe_triangular(u1, r, PE), acc);
      v = acc;
    }
    img[i] = to16<FMT>(v);
  }
}

// Fast-path kernel: 512 threads (16 warps: lane quarter = warp % 4, 16-column slice = warp / 4), two tiles in flight
// per CTA in ping-pong so one tile's MMAs run under the other tile's epilogue; 2 CTAs per SM (256 TMEM columns each).
//   TMEM: D0 [0,64) D1 [64,128) A0 [128,168) A1 [168,208) SEL [208,224).
constexpr int F_THREADS = 512;
constexpr int F_TMEM_COLS = 256;
constexpr int F_COL_D = 0, F_COL_A = 128, F_COL_S = 208;
constexpr int F_STAGE = 4096;                      // per-slot staging: [slot][Ag0 1 KB | G1 rows 1 KB]

template <int FMT, typename OutT>
__global__ void __launch_bounds__(F_THREADS, 2) decode_tc2d_kernel(DevGeom g, const uint2* __restrict__ shadow0,
                                                                   const uint4* __restrict__ R,
                                                                   const uint4* __restrict__ wimg, int cout,
                                                                   unsigned tiles_y, unsigned fd_mul, unsigned fd_shift,
                                                                   OutT* __restrict__ out) {
  using P = Pair<FMT>;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* sW = smem_raw + F_STAGE;                 // weight images + LUTs sit ABOVE the staging buffers (LBO = distance)
  uint8_t* sW1 = sW;
  uint8_t* sW2 = sW1 + F_W1G0;
  uint8_t* sW3 = sW2 + b_image_bytes(64, TC_K2);
  uint8_t* sLx = sW3 + b_image_bytes(16, TC_K2);
  uint8_t* sLy = sLx + F_LUT;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sLy + F_LUT);       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 2);

  const int tid = threadIdx.x, warp = uniform_warp_index();
  const int cs = warp >> 2;                          // which 16 accumulator columns this thread owns
  const int row = tid & (TC_ROWS - 1);
  if (warp == 0) tmem_alloc(tmem_slot, F_TMEM_COLS);
  if (tid == 0) {
    mbar_init(mbar, 1);
    mbar_init(mbar + 1, 1);
  }
  {
    uint4* dst = reinterpret_cast<uint4*>(sW);
    for (int i = tid; i < F_IMG / 16; i += F_THREADS) dst[i] = __ldg(wimg + i);
    for (int i = tid; i < F_STAGE / 16; i += F_THREADS) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;

  // ---- this thread's texel inside a tile: row m = cell + 8 * within
  const int cell = row & 7, within = row >> 3;
  const int lx = 4 * (cell >> 2) + (within >> 2), ly = 4 * (cell & 3) + (within & 3);
  // ---- constants in TMEM: selector rows (slice-1 warps), bias blocks of both A slots (slice-0 warps)
  if (cs == 1) {
    float kx = (float)lx * 0.125f, ky = (float)(ly & 7) * 0.125f;
    int cy1 = ly >> 3;
    float sel[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) sel[i] = 0.f;
#pragma unroll
    for (int ix = 0; ix < 2; ++ix)
#pragma unroll
      for (int iy = 0; iy < 3; ++iy) {
        float wx = ix ? kx : 1.0f - kx;
        float wy = iy == cy1 ? 1.0f - ky : (iy == cy1 + 1 ? ky : 0.f);
        sel[ix * 3 + iy] = wx * wy;
      }
#pragma unroll
    for (int j = 0; j < 8; ++j) sel[8 + j] = j == lx ? 1.f : 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) sel[16 + j] = j == ly ? 1.f : 0.f;
    uint32_t sp[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      auto v = P::pack(sel[2 * i], sel[2 * i + 1]);
      sp[i] = *reinterpret_cast<uint32_t*>(&v);
    }
    tmem_st16(tmem + F_COL_S + lane_base, sp);
  } else if (cs == 0) {
    uint32_t ones[8];
    auto one = P::pack(1.0f, 0.0f);
    ones[0] = *reinterpret_cast<uint32_t*>(&one);
#pragma unroll
    for (int i = 1; i < 8; ++i) ones[i] = 0u;
    tmem_st8(tmem + F_COL_A + lane_base + 32, ones);
    tmem_st8(tmem + F_COL_A + 40 + lane_base + 32, ones);
  }
  tc_wait_st();

  constexpr uint32_t IDESC_64 = make_idesc(FMT, 128, 64), IDESC_16 = make_idesc(FMT, 128, 16);
  constexpr uint32_t IDESC_64_BMN = make_idesc(FMT, 128, 64, 1);
  constexpr uint32_t LBO_64 = (64 / 8) * 128, LBO_16 = (16 / 8) * 128, SBO = 128;
  const uint32_t aStage = smem_u32(smem_raw), aW1 = smem_u32(sW1), aW2 = smem_u32(sW2), aW3 = smem_u32(sW3);
  const uint32_t aLx = smem_u32(sLx), aLy = smem_u32(sLy);
  const int ny0 = g.n0[1], ny1 = g.n1[1];
  const unsigned ntiles = (unsigned)(g.B[0] / F_TX) * tiles_y;
  const unsigned npairs = (ntiles + 1) >> 1;
  uint32_t ph0 = 0, ph1 = 0;

  // ---- staging: thread t < 288 fetches one piece of tile (t >= 144) of the pair
  const int st_slot = tid >= 144, st_t = tid - 144 * st_slot;
  const bool st_active = tid < 288;
  uint4 pre = make_uint4(0, 0, 0, 0);
  auto tile_origin = [&](unsigned tile, int& px0, int& py0, int& bx0, int& by0) {
    unsigned tx = fastdiv31(tile, fd_mul, fd_shift), ty = tile - tx * tiles_y;
    bx0 = (int)tx * F_TX;
    by0 = (int)ty * F_TY;
    px0 = g.origin0[0] + bx0;
    py0 = g.origin0[1] + by0;
  };
  auto prefetch = [&](unsigned pair) {
    unsigned tile = 2 * pair + st_slot;
    if (!st_active || tile >= ntiles) return;
    int px0, py0, bx0, by0;
    tile_origin(tile, px0, py0, bx0, by0);
    if (st_t < 96) {
      int c8 = st_t / 12, piece = st_t - c8 * 12;
      int seg = piece >= 6, off = piece - seg * 6;
      int nx_ = (px0 >> 2) + (c8 >> 2) + seg, ny_ = (py0 >> 2) + (c8 & 3);
      uint2 v = __ldg(shadow0 + ((size_t)nx_ * ny0 + ny_) * 3 + off);      // nodes (x, y) and (x, y + 1) are contiguous
      pre.x = v.x;
      pre.y = v.y;
    } else {
      int t2 = st_t - 96, r6 = t2 >> 3, piece = t2 & 7;
      int nx_ = (px0 >> 3) + r6 / 3, ny_ = (py0 >> 3) + r6 % 3;
      pre = __ldg(R + ((size_t)nx_ * ny1 + ny_) * 8 + piece);
    }
  };
  auto commit_stage = [&]() {
    if (!st_active) return;
    uint8_t* base = smem_raw + st_slot * 2048;
    if (st_t < 96) {
      int c8 = st_t / 12, piece = st_t - c8 * 12;
      int seg = piece >= 6, off = piece - seg * 6;
      int k = seg * 24 + off * 4;
      *reinterpret_cast<uint2*>(base + (k >> 3) * 128 + c8 * 16 + (k & 7) * 2) = make_uint2(pre.x, pre.y);
    } else {
      int t2 = st_t - 96, r6 = t2 >> 3, piece = t2 & 7;
      *reinterpret_cast<uint4*>(base + 1024 + piece * 128 + r6 * 16) = pre;
    }
  };
  // layer 1 of the tile in `slot`: 3 aliased-cell MMAs (G0) + [G1 rows | LUTx] + [LUTy]
  auto issue_layer1 = [&](int slot, unsigned tile) {
    int px0, py0, bx0, by0;
    tile_origin(tile, px0, py0, bx0, by0);
    const uint32_t tD = tmem + F_COL_D + slot * 64, tS = tmem + F_COL_S;
    const uint32_t aAg0 = aStage + slot * 2048, aG1 = aAg0 + 1024;
#pragma unroll
    for (int kc = 0; kc < 3; ++kc)
      mma_ss(tD, make_smem_desc(aAg0 + kc * 256, 128, 0), make_smem_desc(aW1 + kc * 2 * LBO_64, LBO_64, SBO), IDESC_64, kc > 0);
    const uint32_t lx_grp = aLx + ((px0 & 63) >> 3) * 1024;
    mma_ts(tD, tS, make_smem_desc(aG1, lx_grp - aG1, SBO), IDESC_64_BMN, 1);
    mma_ts(tD, tS + 8, make_smem_desc(aLy + ((py0 & 63) >> 3) * 1024, 1024, SBO), IDESC_64_BMN, 1);
    tc_commit(mbar + slot);
  };
  auto issue_layer23 = [&](int slot, int layer) {
    const uint32_t tD = tmem + F_COL_D + slot * 64, tA = tmem + F_COL_A + slot * 40;
    const uint32_t aW = layer == 0 ? aW2 : aW3, lbo = layer == 0 ? LBO_64 : LBO_16;
    const uint32_t idesc = layer == 0 ? IDESC_64 : IDESC_16;
#pragma unroll
    for (int kc = 0; kc < TC_K2 / 16; ++kc)
      mma_ts(tD, tA + kc * 8, make_smem_desc(aW + kc * 2 * lbo, lbo, SBO), idesc, kc > 0);
    tc_commit(mbar + slot);
  };
  // epilogue of one hidden layer: this thread's 16 accumulator columns -> 8 packed activations
  auto epilogue = [&](int slot) {
    uint32_t acc[16];
    tmem_ld16(tmem + F_COL_D + slot * 64 + lane_base + cs * 16, acc);
    tc_wait_ld();
    uint32_t hp[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) hp[i] = gelu2x_pair<FMT>(__uint_as_float(acc[2 * i]), __uint_as_float(acc[2 * i + 1]));
    tmem_st8(tmem + F_COL_A + slot * 40 + lane_base + cs * 8, hp);
    tc_wait_st();
    tc_fence_before();
  };
  auto output = [&](int slot, unsigned tile) {
    if (cs != 0) return;
    int px0, py0, bx0, by0;
    tile_origin(tile, px0, py0, bx0, by0);
    uint32_t acc[16];
    tmem_ld16(tmem + F_COL_D + slot * 64 + lane_base, acc);
    tc_wait_ld();
    const size_t n = (size_t)(bx0 + lx) * g.B[1] + (by0 + ly);
#pragma unroll
    for (int c = 0; c < 16; ++c)
      if (c < cout) store_out(out + n * cout + c, __fdividef(1.0f, 1.0f + __expf(-__uint_as_float(acc[c]))));
    tc_fence_before();
  };

  prefetch(blockIdx.x);
  for (unsigned pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
    const unsigned tA_ = 2 * pair, tB_ = 2 * pair + 1;
    const bool hasB = tB_ < ntiles;
    commit_stage();
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
        tc_fence_after();
        issue_layer1(0, tA_);
        if (hasB) issue_layer1(1, tB_);
      }
      __syncwarp();
    }
    prefetch(pair + gridDim.x);                 // next pair's operands: in flight during this pair's epilogues
#pragma unroll 1
    for (int layer = 0; layer < 2; ++layer) {
      mbar_wait_sleep(mbar, ph0);
      ph0 ^= 1;
      tc_fence_after();
      epilogue(0);
      __syncthreads();
      if (warp == 0) {
        if (elect_one()) {
          tc_fence_after();
          issue_layer23(0, layer);
        }
        __syncwarp();
      }
      if (hasB) {
        mbar_wait_sleep(mbar + 1, ph1);
        ph1 ^= 1;
        tc_fence_after();
        epilogue(1);
      }
      __syncthreads();
      if (hasB && warp == 0) {
        if (elect_one()) {
          tc_fence_after();
          issue_layer23(1, layer);
        }
        __syncwarp();
      }
    }
    mbar_wait_sleep(mbar, ph0);
    ph0 ^= 1;
    tc_fence_after();
    output(0, tA_);
    if (hasB) {
      mbar_wait_sleep(mbar + 1, ph1);
      ph1 ^= 1;
      tc_fence_after();
      output(1, tB_);
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, F_TMEM_COLS);
}

// ================================================================================================ 2-D fast path, v2
// Warp-specialised persistent form of the kernel above.  Knock-out experiments (NIC option 100, tools/run_decode.py)
// showed the first versions were bound by the LATENCY of the per-tile chain (epilogue -> barrier -> MMA issue ->
// MMAs (then ~95 cycles each: they were still issued from divergent control flow, see elect_one() in nic_tc_common.cuh)
// -> commit -> wake-up -> tcgen05.ld), i.e. by how many tiles are in flight, not by any pipe.
// TMEM caps that number: with the next layer's A operand in TMEM a tile needs 104 columns (4 tiles).  Here the
// activations go to SHARED memory instead (st.shared in the UMMA K-major core-matrix layout, SS-form MMAs), a tile
// needs only its 64 accumulator columns, and EIGHT tiles are in flight per SM:
//   * ONE CTA of 1024 threads per SM = eight independent warp-groups of 4 warps; group s owns TMEM columns [64 s, +64),
//     a 16 KB activation buffer and a double-buffered 2 KB operand staging area;
//   * a group runs its tiles as a private pipeline: tcgen05.ld 16 columns -> packed 2*gelu -> one 16-byte st.shared per
//     8 columns; after a 128-thread named barrier its elected lane issues the next tcgen05.mma batch (layer 1: 3
//     aliased-cell MMAs + selector x [G1 rows | LUTx] + selector x LUTy; layers 2, 3: 4 MMAs + 1 bias MMA against a
//     shared constant ones block) and the group waits on its own mbarrier (tcgen05.commit).  Groups never meet at a
//     CTA-wide barrier, so one group's MMAs and hand-offs run under the other groups' epilogues;
//   * the group stages its own operands: its threads fetch the pieces (8 B of a G0 cell row / 16 B of a G1 node row)
//     of the tile after next into registers while the current tile is processed.
constexpr int WS_SLOTS = 8, WS_GROUP = 128, WS_THREADS = WS_SLOTS * WS_GROUP;
constexpr int WS_TMEM_COLS = 512;
constexpr int WS_STAGE = WS_SLOTS * 2 * 2048;       // [slot][buffer][Ag0 1 KB | G1 rows 1 KB]
constexpr int WS_KG = 16 * 128;                     // bytes of one k-group (8 columns) of a 128-row K-major A operand
constexpr int WS_OFF_SEL = WS_STAGE + F_IMG;        // selector matrix, 128 x 32 (4 k-groups)
constexpr int WS_OFF_ONE = WS_OFF_SEL + 4 * WS_KG;  // ones block, 128 x 16: column 0 = 1
constexpr int WS_OFF_ACT = WS_OFF_ONE + 2 * WS_KG;  // [slot] 128 x 64 activations (8 k-groups)
constexpr int WS_OFF_BAR = WS_OFF_ACT + WS_SLOTS * 8 * WS_KG;
constexpr int WS_SMEM = WS_OFF_BAR + 256;

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}

template <int FMT, typename OutT>
__global__ void __launch_bounds__(WS_THREADS, 1) decode_tc2d_ws_kernel(DevGeom g, const uint2* __restrict__ shadow0,
                                                                       const uint4* __restrict__ R,
                                                                       const uint4* __restrict__ wimg, int cout,
                                                                       unsigned tiles_y, unsigned fd_mul, unsigned fd_shift,
                                                                       OutT* __restrict__ out, int dbg) {
  using P = Pair<FMT>;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* sW = smem_raw + WS_STAGE;                // weight images + LUTs sit ABOVE the staging buffers (LBO = distance)
  uint8_t* sW1 = sW;
  uint8_t* sW2 = sW1 + F_W1G0;
  uint8_t* sW3 = sW2 + b_image_bytes(64, TC_K2);
  uint8_t* sLx = sW3 + b_image_bytes(16, TC_K2);
  uint8_t* sLy = sLx + F_LUT;
  uint8_t* sSel = smem_raw + WS_OFF_SEL;
  uint8_t* sOne = smem_raw + WS_OFF_ONE;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_raw + WS_OFF_BAR);     // [8]: tensor core -> group
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_full + WS_SLOTS);

  const int tid = threadIdx.x, warp = uniform_warp_index(), lane = tid & 31;     // warp: provably uniform (MMA operands)
  const int slot = warp >> 2;                        // warp-group = tile slot
  const int row = 32 * (warp & 3) + lane;            // texel row of the tile = TMEM lane
  const int cell = row & 7, within = row >> 3;
  const int lx = 4 * (cell >> 2) + (within >> 2), ly = 4 * (cell & 3) + (within & 3);
  const int roff = (row >> 3) * 128 + (row & 7) * 16;      // this row's 16-byte chunk inside a k-group block
  if (warp == 0) tmem_alloc(tmem_slot, WS_TMEM_COLS);
  if (tid == 0)
    for (int s = 0; s < WS_SLOTS; ++s) mbar_init(bar_full + s, 1);
  {
    uint4* dst = reinterpret_cast<uint4*>(sW);
    for (int i = tid; i < F_IMG / 16; i += WS_THREADS) dst[i] = __ldg(wimg + i);
    for (int i = tid; i < WS_STAGE / 16; i += WS_THREADS) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
  }
  if (slot == 0) {
    // selector row of this texel: [6 tent weights of the 2 x 3 G1 nodes, 0, 0 | one-hot x (8) | one-hot y (16)]
    float kx = (float)lx * 0.125f, ky = (float)(ly & 7) * 0.125f;
    int cy1 = ly >> 3;
    float sel[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) sel[i] = 0.f;
#pragma unroll
    for (int ix = 0; ix < 2; ++ix)
#pragma unroll
      for (int iy = 0; iy < 3; ++iy) {
        float wx = ix ? kx : 1.0f - kx;
        float wy = iy == cy1 ? 1.0f - ky : (iy == cy1 + 1 ? ky : 0.f);
        sel[ix * 3 + iy] = wx * wy;
      }
#pragma unroll
    for (int j = 0; j < 8; ++j) sel[8 + j] = j == lx ? 1.f : 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) sel[16 + j] = j == ly ? 1.f : 0.f;
#pragma unroll
    for (int kg = 0; kg < 4; ++kg) {
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        auto v = P::pack(sel[8 * kg + 2 * i], sel[8 * kg + 2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&v);
      }
      *reinterpret_cast<uint4*>(sSel + kg * WS_KG + roff) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    auto one = P::pack(1.0f, 0.0f);
    *reinterpret_cast<uint4*>(sOne + roff) = make_uint4(*reinterpret_cast<uint32_t*>(&one), 0, 0, 0);
    *reinterpret_cast<uint4*>(sOne + WS_KG + roff) = make_uint4(0, 0, 0, 0);
  }
  fence_async_smem();
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const unsigned ntiles = (unsigned)(g.B[0] / F_TX) * tiles_y;
  const unsigned my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto tile_origin = [&](unsigned tile, int& px0, int& py0, int& bx0, int& by0) {
    unsigned tx = fastdiv31(tile, fd_mul, fd_shift), ty = tile - tx * tiles_y;
    bx0 = (int)tx * F_TX;
    by0 = (int)ty * F_TY;
    px0 = g.origin0[0] + bx0;
    py0 = g.origin0[1] + by0;
  };

  const int gt = tid & (WS_GROUP - 1);               // thread index inside the group
  const uint32_t tDs = tmem + slot * 64;             // this slot's accumulator columns
  const uint32_t tD = tDs + ((uint32_t)((warp & 3) * 32) << 16);
  uint8_t* sAct = smem_raw + WS_OFF_ACT + slot * 8 * WS_KG;
  const bool issuer_warp = (warp & 3) == 0;          // its elected lane issues the group's MMAs (uniform control flow)
  constexpr uint32_t IDESC_64 = make_idesc(FMT, 128, 64), IDESC_16 = make_idesc(FMT, 128, 16);
  constexpr uint32_t IDESC_64_BMN = make_idesc(FMT, 128, 64, 1);
  constexpr uint32_t LBO_64 = (64 / 8) * 128, LBO_16 = (16 / 8) * 128, SBO = 128;
  const uint32_t aStage = smem_u32(smem_raw), aW1 = smem_u32(sW1), aW2 = smem_u32(sW2), aW3 = smem_u32(sW3);
  const uint32_t aLx = smem_u32(sLx), aLy = smem_u32(sLy), aSel = smem_u32(sSel), aOne = smem_u32(sOne), aAct = smem_u32(sAct);
  const int ny0 = g.n0[1], ny1 = g.n1[1];
  auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "n"(WS_GROUP) : "memory"); };
  // wait for the group's MMA batch: ONE warp polls the mbarrier, the other three park at the group's named barrier
  // (a hardware wait: no issue slots) — polling by all four warps was 10 % of the kernel's issued instructions
  auto group_wait = [&](uint32_t phase) {
    if (issuer_warp) mbar_wait(bar_full + slot, phase);
    group_sync();
  };
  // ---- operand staging: 144 pieces per tile over 128 threads (threads 0..15 take a second piece)
  uint4 pre0 = make_uint4(0, 0, 0, 0), pre1 = make_uint4(0, 0, 0, 0);
  auto fetch_piece = [&](int p, int px0, int py0) -> uint4 {
    if (p < 96) {
      int c8 = p / 12, piece = p - c8 * 12;
      int seg = piece >= 6, off = piece - seg * 6;
      int nx_ = (px0 >> 2) + (c8 >> 2) + seg, ny_ = (py0 >> 2) + (c8 & 3);
      uint2 v = __ldg(shadow0 + ((size_t)nx_ * ny0 + ny_) * 3 + off);      // nodes (x, y) and (x, y + 1) are contiguous
      return make_uint4(v.x, v.y, 0, 0);
    }
    int t2 = p - 96, r6 = t2 >> 3, piece = t2 & 7;
    int nx_ = (px0 >> 3) + r6 / 3, ny_ = (py0 >> 3) + r6 % 3;
    return __ldg(R + ((size_t)nx_ * ny1 + ny_) * 8 + piece);
  };
  auto store_piece = [&](uint8_t* base, int p, const uint4& v) {
    if (p < 96) {
      int c8 = p / 12, piece = p - c8 * 12;
      int seg = piece >= 6, off = piece - seg * 6;
      int k = seg * 24 + off * 4;
      *reinterpret_cast<uint2*>(base + (k >> 3) * 128 + c8 * 16 + (k & 7) * 2) = make_uint2(v.x, v.y);
    } else {
      int t2 = p - 96, r6 = t2 >> 3, piece = t2 & 7;
      *reinterpret_cast<uint4*>(base + 1024 + piece * 128 + r6 * 16) = v;
    }
  };
  auto fetch = [&](unsigned i) {
    if (i >= my_tiles) return;
    int px0, py0, bx0, by0;
    tile_origin(blockIdx.x + i * gridDim.x, px0, py0, bx0, by0);
    pre0 = fetch_piece(gt, px0, py0);
    if (gt < 16) pre1 = fetch_piece(128 + gt, px0, py0);
  };
  auto stage = [&](unsigned i) {          // store the fetched pieces of tile i into its staging buffer
    if (i >= my_tiles) return;
    uint8_t* base = smem_raw + (slot * 2 + ((i / WS_SLOTS) & 1)) * 2048;
    store_piece(base, gt, pre0);
    if (gt < 16) store_piece(base, 128 + gt, pre1);
    fence_async_smem();
  };
  auto issue_layer1 = [&](unsigned i) {          // elected lane only; tile i's operands are staged
    int px0, py0, bx0, by0;
    tile_origin(blockIdx.x + i * gridDim.x, px0, py0, bx0, by0);
    const uint32_t aAg0 = aStage + (slot * 2 + ((i / WS_SLOTS) & 1)) * 2048, aG1 = aAg0 + 1024;
#pragma unroll
    for (int kc = 0; kc < 3; ++kc)
      if (!((dbg & 4) && kc > 0))
        mma_ss(tDs, make_smem_desc(aAg0 + kc * 256, 128, 0), make_smem_desc(aW1 + kc * 2 * LBO_64, LBO_64, SBO), IDESC_64, kc > 0);
    const uint32_t lx_grp = aLx + ((px0 & 63) >> 3) * 1024;
    if (!(dbg & 4)) {
      mma_ss(tDs, make_smem_desc(aSel, WS_KG, SBO), make_smem_desc(aG1, lx_grp - aG1, SBO), IDESC_64_BMN, 1);
      mma_ss(tDs, make_smem_desc(aSel + 2 * WS_KG, WS_KG, SBO), make_smem_desc(aLy + ((py0 & 63) >> 3) * 1024, 1024, SBO),
             IDESC_64_BMN, 1);
    }
    tc_commit(bar_full + slot);
  };
  auto issue_layer23 = [&](int layer) {          // elected lane only: A = this slot's activations (+ the ones block: bias)
    const uint32_t aW = layer == 0 ? aW2 : aW3, lbo = layer == 0 ? LBO_64 : LBO_16;
    const uint32_t idesc = layer == 0 ? IDESC_64 : IDESC_16;
#pragma unroll
    for (int kc = 0; kc < 4; ++kc)
      if (!((dbg & 4) && kc > 0)) mma_ss(tDs, make_smem_desc(aAct + kc * 2 * WS_KG, WS_KG, SBO), make_smem_desc(aW + kc * 2 * lbo, lbo, SBO), idesc, kc > 0);
    if (!(dbg & 4)) mma_ss(tDs, make_smem_desc(aOne, WS_KG, SBO), make_smem_desc(aW + 4 * 2 * lbo, lbo, SBO), idesc, 1);
    tc_commit(bar_full + slot);
  };

  // ---- prologue: first tile of this slot
  fetch(slot);
  stage(slot);
  tc_fence_before();
  group_sync();
  if (issuer_warp && (unsigned)slot < my_tiles) {
    if (elect_one()) {
      tc_fence_after();
      issue_layer1(slot);
    }
    __syncwarp();
  }
  fetch(slot + WS_SLOTS);
  uint32_t ph = 0;
  for (unsigned i = slot; i < my_tiles; i += WS_SLOTS) {
    const unsigned tile = blockIdx.x + i * gridDim.x;
#pragma unroll 1
    for (int layer = 0; layer < 2; ++layer) {
      group_wait(ph);
      ph ^= 1;
      tc_fence_after();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t acc[16];
        tmem_ld16(tD + 16 * q, acc);
        tc_wait_ld();
        uint32_t hp[8];
        if (dbg & 2) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            auto v = P::pack(__uint_as_float(acc[2 * k]), __uint_as_float(acc[2 * k + 1]));
            hp[k] = *reinterpret_cast<uint32_t*>(&v);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) hp[k] = gelu2x_pair<FMT>(__uint_as_float(acc[2 * k]), __uint_as_float(acc[2 * k + 1]));
        }
        *reinterpret_cast<uint4*>(sAct + (2 * q) * WS_KG + roff) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
        *reinterpret_cast<uint4*>(sAct + (2 * q + 1) * WS_KG + roff) = make_uint4(hp[4], hp[5], hp[6], hp[7]);
      }
      fence_async_smem();                           // the activations (generic-proxy stores) -> visible to the tensor core
      tc_fence_before();
      group_sync();
      if (issuer_warp) {
        if (elect_one()) {
          tc_fence_after();
          issue_layer23(layer);
        }
        __syncwarp();
      }
    }
    group_wait(ph);
    ph ^= 1;
    tc_fence_after();
    uint32_t acc[16];
    if (cout > 4) tmem_ld16(tD, acc);               // multi-channel outputs (material stacks): up to 16 channels
    else tmem_ld4(tD, acc);
    tc_wait_ld();
    stage(i + WS_SLOTS);                          // operands of this slot's next tile (fetched one tile ago)
    tc_fence_before();
    group_sync();                                 // D has been read and the next operands are staged
    if (issuer_warp && i + WS_SLOTS < my_tiles) {
      if (elect_one()) {
        tc_fence_after();
        issue_layer1(i + WS_SLOTS);
      }
      __syncwarp();
    }
    fetch(i + 2 * WS_SLOTS);
    {
      int px0, py0, bx0, by0;
      tile_origin(tile, px0, py0, bx0, by0);
      const size_t n = (size_t)(bx0 + lx) * g.B[1] + (by0 + ly);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < cout && !(dbg & 1)) store_out(out + n * cout + c, __fdividef(1.0f, 1.0f + __expf(-__uint_as_float(acc[c]))));
      if (cout > 4) {
#pragma unroll
        for (int c = 4; c < 16; ++c)
          if (c < cout) store_out(out + n * cout + c, __fdividef(1.0f, 1.0f + __expf(-__uint_as_float(acc[c]))));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, WS_TMEM_COLS);
}

// ================================================================================================ general kernel, v2
// The general decode (any mip / origin / block shape, 2-D and both 3-D methods, random-access queries) in the same
// warp-specialised form as decode_tc2d_ws_kernel: ONE persistent CTA per SM made of NG independent groups of 4 warps
// (NG = 8 for a decoder input of <= 80 columns, 5 for method 3's 128), group s owning TMEM columns [64 s, +64) and one
// [128 x KX] operand buffer in shared memory.  A group's thread = one texel of its tile:
//   gather : the texel's decoder-input row goes STRAIGHT from the 16-bit channel-last shadow grids to the operand
//            buffer — a G0 corner is three 8-byte loads and three 8-byte shared stores, G1 is interpolated in packed
//            16-bit math (6 registers), the triangular PE comes from a shared LUT — so no thread ever holds the row;
//   layers : SS-form tcgen05.mma issued by the group's elected lane after a 128-thread named barrier (layer 1: KX/16
//            MMAs; layers 2, 3: 4 MMAs + 1 bias MMA against a shared constant ones block), the epilogue overwrites the
//            operand buffer with the next layer's activations;  the groups never meet at a CTA-wide barrier.
constexpr int GW_GROUP = 128;

template <int METHOD, int FMT, typename OutT>
__global__ void __launch_bounds__(RowShape<METHOD>::KX > 80 ? 5 * GW_GROUP : 8 * GW_GROUP, 1)
    decode_tc_gws_kernel(DevGeom g, ShadowGeom sg, const long long* __restrict__ origins, const uint4* __restrict__ wimg,
                         int cout, int lut_n, OutT* __restrict__ out) {
  using S = RowShape<METHOD>;
  using P = Pair<FMT>;
  constexpr int KX = S::KX, NG = KX > 80 ? 5 : 8, THREADS = NG * GW_GROUP;
  constexpr int W1_BYTES = b_image_bytes(64, KX), W2_BYTES = b_image_bytes(64, TC_K2), W3_BYTES = b_image_bytes(16, TC_K2);
  constexpr int KG = 16 * 128;                       // bytes of one k-group (8 columns) of a 128-row K-major operand
  constexpr int OFF_LUT = W1_BYTES + W2_BYTES + W3_BYTES, OFF_ONE = OFF_LUT + TC_LUT_MAX * 16, OFF_ACT = OFF_ONE + 2 * KG;
  constexpr int ACT_BYTES = (KX / 8) * KG, OFF_BAR = OFF_ACT + NG * ACT_BYTES;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* sW1 = smem_raw;
  uint8_t* sW2 = sW1 + W1_BYTES;
  uint8_t* sW3 = sW2 + W2_BYTES;
  uint4* sLut = reinterpret_cast<uint4*>(smem_raw + OFF_LUT);
  uint8_t* sOne = smem_raw + OFF_ONE;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_raw + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_full + NG);

  const int tid = threadIdx.x, warp = uniform_warp_index(), lane = tid & 31;     // warp: provably uniform (MMA operands)
  const int slot = warp >> 2;
  const int row = 32 * (warp & 3) + lane;
  const int roff = (row >> 3) * 128 + (row & 7) * 16;
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (tid == 0)
    for (int s = 0; s < NG; ++s) mbar_init(bar_full + s, 1);
  {
    uint4* dst = reinterpret_cast<uint4*>(smem_raw);
    for (int i = tid; i < OFF_LUT / 16; i += THREADS) dst[i] = __ldg(wimg + i);
    for (int i = tid; i < lut_n; i += THREADS) {
      float u1 = __fmul_rn(__fmul_rn((float)i, g.step), 0.5f);
      uint4 e;
      auto v0 = P::pack(pe_triangular(u1, 0, 6), pe_triangular(u1, 1, 6));
      auto v1 = P::pack(pe_triangular(u1, 2, 6), pe_triangular(u1, 3, 6));
      auto v2 = P::pack(pe_triangular(u1, 4, 6), pe_triangular(u1, 5, 6));
      e.x = *reinterpret_cast<uint32_t*>(&v0);
      e.y = *reinterpret_cast<uint32_t*>(&v1);
      e.z = *reinterpret_cast<uint32_t*>(&v2);
      e.w = 0u;
      sLut[i] = e;
    }
    if (slot == 0) {
      auto one = P::pack(1.0f, 0.0f);
      *reinterpret_cast<uint4*>(sOne + roff) = make_uint4(*reinterpret_cast<uint32_t*>(&one), 0, 0, 0);
      *reinterpret_cast<uint4*>(sOne + KG + roff) = make_uint4(0, 0, 0, 0);
    }
  }
  fence_async_smem();
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tDs = tmem + slot * 64;
  const uint32_t tD = tDs + ((uint32_t)((warp & 3) * 32) << 16);
  uint8_t* sAct = smem_raw + OFF_ACT + slot * ACT_BYTES;
  const bool issuer_warp = (warp & 3) == 0;          // its elected lane issues the group's MMAs (uniform control flow)
  constexpr uint32_t IDESC_64 = make_idesc(FMT, 128, 64), IDESC_16 = make_idesc(FMT, 128, 16);
  constexpr uint32_t LBO_64 = (64 / 8) * 128, LBO_16 = (16 / 8) * 128, SBO = 128;
  const uint32_t aW1 = smem_u32(sW1), aW2 = smem_u32(sW2), aW3 = smem_u32(sW3), aOne = smem_u32(sOne), aAct = smem_u32(sAct);
  auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "n"(GW_GROUP) : "memory"); };
  auto group_wait = [&](uint32_t phase) {          // one polling warp per group, the others park at the named barrier
    if (issuer_warp) mbar_wait(bar_full + slot, phase);
    group_sync();
  };
  // this thread's 8-byte piece holding features [f, f + 4) of its row (f a multiple of 4)
  auto piece = [&](int f) -> uint2* { return reinterpret_cast<uint2*>(sAct + (f >> 3) * KG + roff + (f & 7) * 2); };

  const unsigned ntiles = (unsigned)((g.N + TC_ROWS - 1) / TC_ROWS);
  const int lut_mask = lut_n - 1;
  uint32_t ph = 0;
  for (unsigned tile = blockIdx.x + gridDim.x * slot; tile < ntiles; tile += gridDim.x * NG) {
    const unsigned n = tile * TC_ROWS + row;
    const bool live = n < (unsigned)g.N;
    // ------------------------------------------------------------------------------------ gather -> operand buffer
    {
      Texel t = texel_of_fast(g, live ? n : (unsigned)g.N - 1, origins);
      AxisCoord ax[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        ax[a] = axis_coord(t.p[a], g.step);
        ax[a].i0 = clampi(ax[a].i0, 0, g.n0[a] - 2 < 0 ? 0 : g.n0[a] - 2);
        ax[a].i1 = clampi(ax[a].i1, 0, g.n1[a] - 2 < 0 ? 0 : g.n1[a] - 2);
      }
#pragma unroll
      for (int j = 0; j < S::NC0; ++j) {             // G0 corners: raw copies
        const int8_t* d = S::DIM == 2 ? kCorner2D[j] : (METHOD == NIC_METHOD_3D ? kCorner3D[j] : kCorner3Dv2[j]);
        const int dz = S::DIM == 3 ? d[0] : 0;
        const uint2* node = sg.s0 + 3 * node_lin(g.n0, S::DIM, ax[0].i0 + d[2], ax[1].i0 + d[1], ax[2].i0 + dz);
#pragma unroll
        for (int q = 0; q < 3; ++q) *piece(12 * j + 4 * q) = __ldg(node + q);
      }
      typename P::T2 acc[6];                          // G1: weighted sum of corners
#pragma unroll
      for (int j = 0; j < S::NC1; ++j) {
        const int8_t* d = S::DIM == 2 ? kCorner2D[j] : kCorner3D[j];
        const int dz = S::DIM == 3 ? d[0] : 0;
        const uint2* node = sg.s1 + 3 * node_lin(g.n1, S::DIM, ax[0].i1 + d[2], ax[1].i1 + d[1], ax[2].i1 + dz);
        float f[3];
        g1_factors(g, j, ax, f);
        float w = f[0] * f[1];
        if (S::DIM == 3) w *= f[2];
        typename P::T2 w2 = P::cst(w);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          uint2 v = __ldg(node + q);
          typename P::T2 va = *reinterpret_cast<typename P::T2*>(&v.x), vb = *reinterpret_cast<typename P::T2*>(&v.y);
          acc[2 * q] = j == 0 ? __hmul2(va, w2) : __hfma2(va, w2, acc[2 * q]);
          acc[2 * q + 1] = j == 0 ? __hmul2(vb, w2) : __hfma2(vb, w2, acc[2 * q + 1]);
        }
      }
#pragma unroll
      for (int q = 0; q < 3; ++q)
        *piece(12 * S::NC0 + 4 * q) = make_uint2(*reinterpret_cast<uint32_t*>(&acc[2 * q]), *reinterpret_cast<uint32_t*>(&acc[2 * q + 1]));
      // positional encoding (3 pairs per axis), the bias carrier (feature CIN - 1 = 1) and zero padding up to KX: the
      // tail of the row, [12 (NC0 + 1), KX), is assembled as packed pairs and stored in 8-byte pieces
      constexpr int T0 = 12 * (S::NC0 + 1), NTAIL = (KX - T0) / 2;      // pairs in the tail
      uint32_t tail[NTAIL];
#pragma unroll
      for (int i = 0; i < NTAIL; ++i) tail[i] = 0u;
#pragma unroll
      for (int a = 0; a < S::DIM; ++a) {
        if (g.pe_kind == NIC_PE_TRIANGULAR) {
          uint4 e = sLut[t.p[a] & lut_mask];
          tail[3 * a] = e.x;
          tail[3 * a + 1] = e.y;
          tail[3 * a + 2] = e.z;
        } else {
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            float arg = __fmul_rn(ax[a].u1, g.pe_div[q]);
            auto v = P::pack(sinf(arg), cosf(arg));
            tail[3 * a + q] = *reinterpret_cast<uint32_t*>(&v);
          }
        }
      }
      {
        auto one = P::pack(1.0f, 0.0f);
        tail[(S::CIN - 1 - T0) / 2] = *reinterpret_cast<uint32_t*>(&one);
      }
#pragma unroll
      for (int i = 0; i < NTAIL / 2; ++i) *piece(T0 + 4 * i) = make_uint2(tail[2 * i], tail[2 * i + 1]);
    }
    fence_async_smem();
    tc_fence_before();
    group_sync();
    if (issuer_warp) {
      if (elect_one()) {
        tc_fence_after();
#pragma unroll
        for (int kc = 0; kc < KX / 16; ++kc)
          mma_ss(tDs, make_smem_desc(aAct + kc * 2 * KG, KG, SBO), make_smem_desc(aW1 + kc * 2 * LBO_64, LBO_64, SBO), IDESC_64, kc > 0);
        tc_commit(bar_full + slot);
      }
      __syncwarp();
    }
    // ------------------------------------------------------------------------------------ layers 1, 2: epilogues
#pragma unroll 1
    for (int layer = 0; layer < 2; ++layer) {
      group_wait(ph);
      ph ^= 1;
      tc_fence_after();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t acc[16];
        tmem_ld16(tD + 16 * q, acc);
        tc_wait_ld();
        uint32_t hp[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) hp[k] = gelu2x_pair<FMT>(__uint_as_float(acc[2 * k]), __uint_as_float(acc[2 * k + 1]));
        *reinterpret_cast<uint4*>(sAct + (2 * q) * KG + roff) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
        *reinterpret_cast<uint4*>(sAct + (2 * q + 1) * KG + roff) = make_uint4(hp[4], hp[5], hp[6], hp[7]);
      }
      fence_async_smem();
      tc_fence_before();
      group_sync();
      if (issuer_warp) {
        if (elect_one()) {
          tc_fence_after();
          const uint32_t aW = layer == 0 ? aW2 : aW3, lbo = layer == 0 ? LBO_64 : LBO_16;
          const uint32_t idesc = layer == 0 ? IDESC_64 : IDESC_16;
#pragma unroll
          for (int kc = 0; kc < 4; ++kc)
            mma_ss(tDs, make_smem_desc(aAct + kc * 2 * KG, KG, SBO), make_smem_desc(aW + kc * 2 * lbo, lbo, SBO), idesc, kc > 0);
          mma_ss(tDs, make_smem_desc(aOne, KG, SBO), make_smem_desc(aW + 4 * 2 * lbo, lbo, SBO), idesc, 1);
          tc_commit(bar_full + slot);
        }
        __syncwarp();
      }
    }
    // ------------------------------------------------------------------------------------ output
    group_wait(ph);
    ph ^= 1;
    tc_fence_after();
    uint32_t acc[16];
    if (cout > 4) tmem_ld16(tD, acc);
    else tmem_ld4(tD, acc);
    tc_wait_ld();
    tc_fence_before();
    if (live) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < cout) store_out(out + (size_t)n * cout + c, __fdividef(1.0f, 1.0f + __expf(-__uint_as_float(acc[c]))));
      if (cout > 4) {
#pragma unroll
        for (int c = 4; c < 16; ++c)
          if (c < cout) store_out(out + (size_t)n * cout + c, __fdividef(1.0f, 1.0f + __expf(-__uint_as_float(acc[c]))));
      }
    }
    // (every thread passes its own tcgen05.wait::ld above before it reaches the next tile's pre-MMA barrier, so the next
    //  layer-1 MMAs cannot overwrite D early)
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ launcher
// NIC_OPT_REUSE_PREPARED: do the private tables already describe these inputs?
static Handle::PreparedKey make_key(const DevGeom& g, const MlpDev& m, const float* g0, const float* g1, int fmt, int fast,
                                    int code_bits = 0) {
  Handle::PreparedKey k;
  memset(&k, 0, sizeof(k));
  k.g0 = g0; k.g1 = g1; k.w1 = m.w1; k.b1 = m.b1; k.w2 = m.w2; k.b2 = m.b2; k.w3 = m.w3; k.b3 = m.b3;
  for (int a = 0; a < 3; ++a) { k.n0[a] = g.n0[a]; k.n1[a] = g.n1[a]; }
  k.method = g.method; k.pe_kind = g.pe_kind; k.mip = g.mip; k.fmt = fmt; k.fast = fast; k.valid = 1;
  k.code_bits = code_bits;
  k.step = g.step;
  return k;
}
static bool prepared_matches(Handle* h, const Handle::PreparedKey& k) {
  return h->reuse_prepared && h->prepared.valid && memcmp(&h->prepared, &k, sizeof(k)) == 0;
}

static bool fast2d_eligible(const DevGeom& g, const long long* origins) {
  return g.method == NIC_METHOD_2D && g.step == 0.25f && g.interp && g.pe_kind == NIC_PE_TRIANGULAR && !origins &&
         g.nblocks == 1 && g.B[0] % F_TX == 0 && g.B[1] % F_TY == 0 && g.origin0[0] % F_TX == 0 &&
         g.origin0[1] % F_TY == 0 && g.B[0] > 0 && g.B[1] > 0;
}

template <int FMT, typename OutT>
static int launch_fast2d(Handle* h, const DevGeom& g, const MlpDev& m, const float* g0, const float* g1, OutT* out,
                         cudaStream_t st) {
  int rc = ensure_scratch(&h->tc_weights, &h->tc_weights_bytes, 64 * 1024);
  if (rc) return rc;
  const long long nodes0 = plane_size_host(g.n0, 2), nodes1 = plane_size_host(g.n1, 2);
  const size_t b0 = ((size_t)nodes0 * g.C * 2 + 255) & ~(size_t)255, b1 = (size_t)nodes1 * 128;
  rc = ensure_scratch(&h->tc_shadow, &h->tc_shadow_bytes, b0 + b1);
  if (rc) return rc;
  uint16_t* s0 = (uint16_t*)h->tc_shadow;
  uint16_t* R = (uint16_t*)((uint8_t*)h->tc_shadow + b0);
  cudaError_t e = cudaSuccess;
  const Handle::PreparedKey key = make_key(g, m, g0, g1, FMT, 1, h->src_code_bits);
  if (!prepared_matches(h, key)) {
    h->prepared.valid = 0;
    e = (cudaError_t)launch_relayout<FMT>(h, g, g0, g.n0, s0, st);
    if (e != cudaSuccess) return (int)e;
    dim3 grid_r((g.n1[0] + 127) / 128, g.n1[1]);
    g1_rows_kernel<FMT><<<grid_r, 128, 0, st>>>(m, g1, g.n1[0], g.n1[1], R, h->src_code_bits);
    h->launches++;
    pack_fast_kernel<FMT><<<32, 256, 0, st>>>(m, g.lod, g.step, (uint16_t*)h->tc_weights);
    h->launches++;
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    memcpy(&h->prepared, &key, sizeof(key));
  }
  unsigned tiles_y = (unsigned)(g.B[1] / F_TY);
  long long ntiles = (long long)(g.B[0] / F_TX) * tiles_y;
  if (ntiles >= (1ll << 31)) return NIC_ERR_UNSUPPORTED;
  unsigned sh = 0;
  while ((1ull << sh) < tiles_y) ++sh;
  unsigned mul = (unsigned)(((1ull << (31 + sh)) + tiles_y - 1) / tiles_y);
  if (!h->legacy_fast2d) {
    // warp-specialised persistent kernel: one CTA per SM (all 512 TMEM columns)
    static_assert(WS_SMEM <= 227 * 1024, "shared memory budget");
    auto kern = decode_tc2d_ws_kernel<FMT, OutT>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_SMEM);
    if (e != cudaSuccess) return (int)e;
    int grid = (int)(ntiles < h->sms ? ntiles : h->sms);
    {
      KernelTimer timer(h, st);
      kern<<<grid, WS_THREADS, WS_SMEM, st>>>(g, (const uint2*)s0, (const uint4*)R, (const uint4*)h->tc_weights, m.cout,
                                               tiles_y, mul, sh, out, h->debug_flags);
    }
    h->launches++;
    return (int)cudaGetLastError();
  }
  // 100 KB of dynamic shared memory per CTA caps residency at 2 CTAs/SM = 2 x 256 TMEM columns = all 512.
  size_t smem = 100 * 1024;
  static_assert(F_STAGE + F_IMG + 64 <= 100 * 1024, "shared memory budget");
  auto kern = decode_tc2d_kernel<FMT, OutT>;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  long long npairs = (ntiles + 1) / 2, cap = (long long)h->sms * 2;
  int grid = (int)(npairs < cap ? npairs : cap);
  {
    KernelTimer timer(h, st);
    kern<<<grid, F_THREADS, smem, st>>>(g, (const uint2*)s0, (const uint4*)R, (const uint4*)h->tc_weights, m.cout,
                                        tiles_y, mul, sh, out);
  }
  h->launches++;
  return (int)cudaGetLastError();
}

template <int METHOD, int FMT, typename OutT>
static int launch_tc_t(Handle* h, const DevGeom& g, const MlpDev& m, const float* g0, const float* g1,
                       const long long* origins, OutT* out, cudaStream_t st) {
  using S = RowShape<METHOD>;
  if constexpr (METHOD == NIC_METHOD_2D) {
    if (fast2d_eligible(g, origins) && !h->disable_fast2d) return launch_fast2d<FMT, OutT>(h, g, m, g0, g1, out, st);
  }
  constexpr int IMG = b_image_bytes(64, S::KX) + b_image_bytes(64, TC_K2) + b_image_bytes(16, TC_K2);
  int rc = ensure_scratch(&h->tc_weights, &h->tc_weights_bytes, 64 * 1024);
  if (rc) return rc;
  // 16-bit channel-last shadows of the two active grids
  const long long nodes0 = plane_size_host(g.n0, g.dim), nodes1 = plane_size_host(g.n1, g.dim);
  const size_t b0 = ((size_t)nodes0 * g.C * 2 + 255) & ~(size_t)255, b1 = ((size_t)nodes1 * g.C * 2 + 255) & ~(size_t)255;
  rc = ensure_scratch(&h->tc_shadow, &h->tc_shadow_bytes, b0 + b1);
  if (rc) return rc;
  uint16_t* s0 = (uint16_t*)h->tc_shadow;
  uint16_t* s1 = (uint16_t*)((uint8_t*)h->tc_shadow + b0);
  cudaError_t e = cudaSuccess;
  const Handle::PreparedKey key = make_key(g, m, g0, g1, FMT, 0, h->src_code_bits);
  if (!prepared_matches(h, key)) {
    h->prepared.valid = 0;
    e = (cudaError_t)launch_relayout<FMT>(h, g, g0, g.n0, s0, st);
    if (e != cudaSuccess) return (int)e;
    e = (cudaError_t)launch_relayout<FMT>(h, g, g1, g.n1, s1, st);
    if (e != cudaSuccess) return (int)e;
    pack_weights_kernel<FMT><<<16, 256, 0, st>>>(m, g.lod, S::KX, (uint16_t*)h->tc_weights);
    h->launches++;
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    memcpy(&h->prepared, &key, sizeof(key));
  }
  // positional-encoding LUT length: the triangular encoding has period 16/step texels
  int lut_n = 1;
  {
    float period = 16.0f / g.step;
    while (lut_n < period && lut_n < TC_LUT_MAX) lut_n <<= 1;
    if (g.pe_kind == NIC_PE_TRIANGULAR && (float)lut_n < period) return NIC_ERR_UNSUPPORTED;
  }
  if (g.N >= (1ll << 31) - TC_ROWS) return NIC_ERR_UNSUPPORTED;     // 32-bit sample index in the kernel
  long long ntiles = (g.N + TC_ROWS - 1) / TC_ROWS;
  ShadowGeom sg = {(const uint2*)s0, (const uint2*)s1};
  if (!h->legacy_fast2d) {
    // warp-specialised persistent kernel: one CTA per SM, NG groups of 4 warps
    constexpr int NG = S::KX > 80 ? 5 : 8;
    constexpr int GSMEM = IMG + TC_LUT_MAX * 16 + 2 * 2048 + NG * (S::KX / 8) * 2048 + 256;
    static_assert(GSMEM <= 227 * 1024, "shared memory budget");
    auto kern = decode_tc_gws_kernel<METHOD, FMT, OutT>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GSMEM);
    if (e != cudaSuccess) return (int)e;
    long long groups = (ntiles + NG - 1) / NG;
    int grid = (int)(groups < h->sms ? groups : h->sms);
    {
      KernelTimer timer(h, st);
      kern<<<grid, NG * GW_GROUP, GSMEM, st>>>(g, sg, origins, (const uint4*)h->tc_weights, m.cout, lut_n, out);
    }
    h->launches++;
    return (int)cudaGetLastError();
  }
  // first-generation kernel: 56 KB of dynamic shared memory per CTA caps residency at 4 CTAs/SM = 4 x 128 TMEM columns.
  size_t smem = 56 * 1024;
  static_assert(IMG + TC_LUT_MAX * 16 + 64 <= 56 * 1024, "shared memory budget");
  auto kern = decode_tc_kernel<METHOD, FMT, OutT>;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  long long cap = (long long)h->sms * 4;
  int grid = (int)(ntiles < cap ? ntiles : cap);
  {
    KernelTimer timer(h, st);
    kern<<<grid, TC_THREADS, smem, st>>>(g, sg, origins, (const uint4*)h->tc_weights, m.cout, lut_n, out);
  }
  h->launches++;
  return (int)cudaGetLastError();
}

template <int METHOD>
static int launch_tc_m(Handle* h, const DevGeom& g, const MlpDev& m, const float* g0, const float* g1,
                       const long long* origins, void* out, int out_dtype, int precision, cudaStream_t st) {
  if (precision == NIC_PREC_F16) {
    if (out_dtype == NIC_DT_U8) return launch_tc_t<METHOD, 0, uint8_t>(h, g, m, g0, g1, origins, (uint8_t*)out, st);
    return launch_tc_t<METHOD, 0, float>(h, g, m, g0, g1, origins, (float*)out, st);
  }
  if (out_dtype == NIC_DT_U8) return launch_tc_t<METHOD, 1, uint8_t>(h, g, m, g0, g1, origins, (uint8_t*)out, st);
  return launch_tc_t<METHOD, 1, float>(h, g, m, g0, g1, origins, (float*)out, st);
}

int launch_decode_tc(Handle* h, const DevGeom& g, const MlpDev& m, const float* g0, const float* g1,
                     const long long* origins, void* out, int out_dtype, int precision, cudaStream_t st) {
  if (g.N == 0) return NIC_OK;
  if (g.C != 12 || g.PE != 6 || m.hidden != 64 || m.cout > 16) return NIC_ERR_UNSUPPORTED;
  switch (g.method) {
    case NIC_METHOD_2D: return launch_tc_m<NIC_METHOD_2D>(h, g, m, g0, g1, origins, out, out_dtype, precision, st);
    case NIC_METHOD_3D: return launch_tc_m<NIC_METHOD_3D>(h, g, m, g0, g1, origins, out, out_dtype, precision, st);
    case NIC_METHOD_3D_V2: return launch_tc_m<NIC_METHOD_3D_V2>(h, g, m, g0, g1, origins, out, out_dtype, precision, st);
  }
  return NIC_ERR_UNSUPPORTED;
}

}  // namespace nic
