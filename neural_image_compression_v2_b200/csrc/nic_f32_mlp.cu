// nic_f32_mlp.cu — fp32 (reference-exact, erf GELU) decoder MLP kernels: forward on a materialised input or fused
// with the gather (K2 fp32), and backward / fused training step (K3 fp32 + K4 scatter).  Compiled once per hidden
// width and direction (-DNIC_H=64|32 -DNIC_PART=0 forward | 1 backward) so the heavy unrolled bodies build in parallel.
// Reference: Projects/image_compression.py:54-68 (ColorDecoder), :239-265 (training step).
#include "nic_internal.cuh"

#ifndef NIC_H
#error "compile with -DNIC_H=<hidden width> -DNIC_PART=<0|1>"
#endif
#define NIC_CAT2(a, b) a##b
#define NIC_CAT(a, b) NIC_CAT2(a, b)

namespace nic {

// ===================================================================================================== MLP (fp32)
// Shared-memory layout of the weights for the fp32 kernels: W1t [cin][H], W2t [H][H] (transposed so a thread
// reads the H outputs of one input as broadcast float4s), W3 [cout][H], b1, b2, b3.
struct SmemMlp {
  float* w1t;
  float* w2t;
  float* w3;
  float* b1;
  float* b2;
  float* b3;
};

template <int H>
__device__ __forceinline__ SmemMlp carve_mlp(float* base, int cin, int cout) {
  SmemMlp s;
  s.w1t = base;
  s.w2t = s.w1t + (size_t)cin * H;
  s.w3 = s.w2t + H * H;
  s.b1 = s.w3 + cout * H;
  s.b2 = s.b1 + H;
  s.b3 = s.b2 + H;
  return s;
}

template <int H>
__device__ __forceinline__ void load_mlp(const SmemMlp& s, const MlpDev& m) {
  for (int i = threadIdx.x; i < m.cin * H; i += blockDim.x) {
    int k = i / H, j = i - k * H;
    s.w1t[i] = m.w1[j * m.cin + k];
  }
  for (int i = threadIdx.x; i < H * H; i += blockDim.x) {
    int k = i / H, j = i - k * H;
    s.w2t[i] = m.w2[j * H + k];
  }
  for (int i = threadIdx.x; i < m.cout * H; i += blockDim.x) s.w3[i] = m.w3[i];
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    s.b1[i] = m.b1[i];
    s.b2[i] = m.b2[i];
  }
  for (int i = threadIdx.x; i < m.cout; i += blockDim.x) s.b3[i] = m.b3[i];
}

template <int H>
__device__ __forceinline__ void axpy_row(float* __restrict__ z, float v, const float* __restrict__ wrow) {
  const float4* w = reinterpret_cast<const float4*>(wrow);
#pragma unroll
  for (int q = 0; q < H / 4; ++q) {
    float4 ww = w[q];
    z[4 * q + 0] = fmaf(v, ww.x, z[4 * q + 0]);
    z[4 * q + 1] = fmaf(v, ww.y, z[4 * q + 1]);
    z[4 * q + 2] = fmaf(v, ww.z, z[4 * q + 2]);
    z[4 * q + 3] = fmaf(v, ww.w, z[4 * q + 3]);
  }
}

template <int H>
__device__ __forceinline__ float dot_row(const float* __restrict__ z, const float* __restrict__ wrow) {
  const float4* w = reinterpret_cast<const float4*>(wrow);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int q = 0; q < H / 4; ++q) {
    float4 ww = w[q];
    a0 = fmaf(z[4 * q + 0], ww.x, a0);
    a1 = fmaf(z[4 * q + 1], ww.y, a1);
    a2 = fmaf(z[4 * q + 2], ww.z, a2);
    a3 = fmaf(z[4 * q + 3], ww.w, a3);
  }
  return (a0 + a1) + (a2 + a3);
}

// Layers 2 and 3 given z1 (pre-activation of layer 1, bias included).  Returns outputs in o[].
template <int H>
__device__ __forceinline__ void mlp_tail(const SmemMlp& s, int cout, const float* z1, float* z2, float* o) {
#pragma unroll
  for (int j = 0; j < H; ++j) z2[j] = s.b2[j];
#pragma unroll
  for (int k = 0; k < H; ++k) axpy_row<H>(z2, gelu_erf(z1[k]), s.w2t + k * H);
  float h2[H];
#pragma unroll
  for (int k = 0; k < H; ++k) h2[k] = gelu_erf(z2[k]);
#pragma unroll
  for (int c = 0; c < NIC_MAX_COUT; ++c)
    if (c < cout) o[c] = sigmoidf_exact(dot_row<H>(h2, s.w3 + c * H) + s.b3[c]);
}

#if NIC_PART == 0
// Forward on a materialised X (ColorDecoder.forward) or fused with the gather (decode).
//   FUSED = 0: x read from global [N, ldx];  FUSED = 1: x produced per texel by for_each_input.
template <int H, int FUSED>
__global__ void __launch_bounds__(128) mlp_forward_kernel(DevGeom g, MlpDev m, const float* __restrict__ g0,
                                                          const float* __restrict__ g1,
                                                          const long long* __restrict__ origins,
                                                          const float* __restrict__ x, long long ldx, long long N,
                                                          void* __restrict__ out, int out_u8,
                                                          float* __restrict__ z1_out, float* __restrict__ z2_out) {
  extern __shared__ __align__(16) float smem[];
  SmemMlp s = carve_mlp<H>(smem, m.cin, m.cout);
  load_mlp<H>(s, m);
  __syncthreads();
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N;
       n += (long long)gridDim.x * blockDim.x) {
    float z1[H];
#pragma unroll
    for (int j = 0; j < H; ++j) z1[j] = s.b1[j];
    if (FUSED) {
      Texel t = texel_of(g, n, origins);
      AxisCoord ax[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) ax[a] = axis_coord(t.p[a], g.step);
      for_each_input(g, g0, g1, ax, [&](int col, float v) { axpy_row<H>(z1, v, s.w1t + col * H); });
    } else {
      const float* xr = x + n * ldx;
      for (int k = 0; k < m.cin; ++k) axpy_row<H>(z1, __ldg(xr + k), s.w1t + k * H);
    }
    float z2[H], o[NIC_MAX_COUT];
    mlp_tail<H>(s, m.cout, z1, z2, o);
    if (z1_out) {
#pragma unroll
      for (int j = 0; j < H; j += 4)
        *reinterpret_cast<float4*>(z1_out + n * H + j) = make_float4(z1[j], z1[j + 1], z1[j + 2], z1[j + 3]);
    }
    if (z2_out) {
#pragma unroll
      for (int j = 0; j < H; j += 4)
        *reinterpret_cast<float4*>(z2_out + n * H + j) = make_float4(z2[j], z2[j + 1], z2[j + 2], z2[j + 3]);
    }
#pragma unroll
    for (int c = 0; c < NIC_MAX_COUT; ++c)
      if (c < m.cout) {
        if (out_u8) store_out((uint8_t*)out + n * m.cout + c, o[c]);
        else store_out((float*)out + n * m.cout + c, o[c]);
      }
  }
}

#else
// ===================================================================================================== backward
// Per-CTA tile of T = 128 samples.  Thread = sample for the forward / delta propagation; thread = slice of a
// weight matrix for the outer-product sums over the tile (staged through shared memory), which are flushed
// with red.global.add once per tile.  Phases are ordered so that at most two H-wide register arrays are live.
//   FUSED = 0: standalone backward of ColorDecoder (x, z1, z2, out, dout given; dx optional).
//   FUSED = 1: training step: gather + noise + forward + MSE + backward + grid-gradient scatter.
template <int H, int FUSED>
__global__ void __launch_bounds__(128, 1) mlp_backward_kernel(
    DevGeom g, MlpDev m, MlpGradDev gm, const float* __restrict__ g0, const float* __restrict__ g1,
    const long long* __restrict__ origins, const float* __restrict__ x, long long ldx, long long N,
    const float* __restrict__ z1_in, const float* __restrict__ z2_in, const float* __restrict__ out_in,
    const float* __restrict__ dout_in, float* __restrict__ dx_out, const float* __restrict__ targets,
    const float* __restrict__ noise, int noise_bits, unsigned long long seed, unsigned long long step,
    float grad_scale, float* __restrict__ dg0, float* __restrict__ dg1, float* __restrict__ loss_sum,
    float* __restrict__ out_save) {
  static_assert(H % 4 == 0 && H <= 64 && 128 % H == 0, "hidden width");
  constexpr int T = 128;
  constexpr int PARTS = T / H;                       // weight matrices are split PARTS ways along their inputs
  constexpr int SPAN1_MAX = (NIC_MAX_CIN + PARTS - 1) / PARTS;
  constexpr int SPAN2 = H / PARTS;
  extern __shared__ __align__(16) float smem[];
  SmemMlp s = carve_mlp<H>(smem, m.cin, m.cout);
  const int cin = m.cin, cout = m.cout;
  const int ldxs = cin | 1;                          // odd pitches: conflict-free per-thread rows
  float* sX = s.b3 + ((cout + 3) & ~3);              // x~ tile   [T][ldxs]
  float* sH = sX + (size_t)T * ldxs;                 // h tile    [T][H+1]  (h2, then h1)
  float* sD = sH + (size_t)T * (H + 1);              // delta     [T][H+1]  (dz2, then dz1)
  float* sD3 = sD + (size_t)T * (H + 1);             // dz3       [T][NIC_MAX_COUT]
  float* sRed = sD3 + (size_t)T * NIC_MAX_COUT;      // block reduction scratch
  load_mlp<H>(s, m);
  __syncthreads();

  const int tid = threadIdx.x;
  const int jrow = tid % H, part = tid / H;
  const int span1 = (cin + PARTS - 1) / PARTS, k1base = part * span1;
  const int k2base = part * SPAN2;
  float loss_local = 0.f, sse8_local = 0.f;

  const long long ntiles = (N + T - 1) / T;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long n = tile * T + tid;
    const bool live = n < N;
    float z1[H], z2[H];
    float dz3[NIC_MAX_COUT];
    AxisCoord ax[3];
    // ---------------- F: forward (fused) or reload (standalone) ----------------
    if (FUSED) {
#pragma unroll
      for (int j = 0; j < H; ++j) z1[j] = s.b1[j];
      if (live) {
        Texel t = texel_of(g, n, origins);
#pragma unroll
        for (int a = 0; a < 3; ++a) ax[a] = axis_coord(t.p[a], g.step);
        for_each_input(g, g0, g1, ax, [&](int col, float v) {
          if (noise) v += noise[n * cin + col];
          else if (noise_bits > 0) v += philox_noise(seed, step, (unsigned long long)n * cin + col, noise_bits);
          sX[tid * ldxs + col] = v;
          axpy_row<H>(z1, v, s.w1t + col * H);
        });
      } else {
        for (int k = 0; k < cin; ++k) sX[tid * ldxs + k] = 0.f;
      }
      float o[NIC_MAX_COUT];
      mlp_tail<H>(s, cout, z1, z2, o);
#pragma unroll
      for (int c = 0; c < NIC_MAX_COUT; ++c) {
        float d = 0.f, oc = c < cout ? o[c] : 0.f;
        if (live && c < cout) {
          const float tc = targets[n * cout + c];
          d = oc - tc;
          loss_local += d * d;
          if (g.metrics) {          // calculate_psnr(quantize_to_bit(out), quantize_to_bit(target)), image_compression.py:260-261
            const float d8 = quant_round(oc, 255.0f) - quant_round(tc, 255.0f);
            sse8_local += d8 * d8;
          }
          if (out_save) out_save[n * cout + c] = oc;
        }
        dz3[c] = 2.0f * d * grad_scale * oc * (1.0f - oc);
      }
    } else {
      for (int k = 0; k < cin; ++k) sX[tid * ldxs + k] = live ? __ldg(x + n * ldx + k) : 0.f;
#pragma unroll
      for (int j = 0; j < H; ++j) {
        z1[j] = live ? z1_in[n * H + j] : 0.f;
        z2[j] = live ? z2_in[n * H + j] : 0.f;
      }
#pragma unroll
      for (int c = 0; c < NIC_MAX_COUT; ++c) {
        float oc = (live && c < cout) ? out_in[n * cout + c] : 0.f;
        dz3[c] = (live && c < cout) ? dout_in[n * cout + c] * oc * (1.0f - oc) : 0.f;
      }
    }
    // ---------------- B1: h2 -> sH, dz3 -> sD3, z2 <- dz2 ----------------
#pragma unroll
    for (int c = 0; c < NIC_MAX_COUT; ++c) sD3[tid * NIC_MAX_COUT + c] = dz3[c];
#pragma unroll
    for (int k = 0; k < H; ++k) {
      sH[tid * (H + 1) + k] = gelu_erf(z2[k]);
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < NIC_MAX_COUT; ++c)
        if (c < cout) d = fmaf(s.w3[c * H + k], dz3[c], d);
      z2[k] = d * gelu_erf_grad(z2[k]);
    }
    __syncthreads();
    // ---------------- (c) dW3 += dz3^T h2 ; db3 ----------------
    if (tid < H) {
      float a3[NIC_MAX_COUT];
#pragma unroll
      for (int c = 0; c < NIC_MAX_COUT; ++c) a3[c] = 0.f;
      for (int si = 0; si < T; ++si) {
        float hv = sH[si * (H + 1) + tid];
#pragma unroll
        for (int c = 0; c < NIC_MAX_COUT; ++c) a3[c] = fmaf(sD3[si * NIC_MAX_COUT + c], hv, a3[c]);
      }
#pragma unroll
      for (int c = 0; c < NIC_MAX_COUT; ++c)
        if (c < cout) atomicAdd(gm.w3 + c * H + tid, a3[c]);
    } else if (tid - H < cout) {
      float a = 0.f;
      for (int si = 0; si < T; ++si) a += sD3[si * NIC_MAX_COUT + (tid - H)];
      atomicAdd(gm.b3 + (tid - H), a);
    }
    __syncthreads();
    // ---------------- B2: h1 -> sH, dz2 -> sD, z1 <- dz1 ----------------
#pragma unroll
    for (int k = 0; k < H; ++k) {
      sH[tid * (H + 1) + k] = gelu_erf(z1[k]);
      sD[tid * (H + 1) + k] = z2[k];
    }
#pragma unroll
    for (int k = 0; k < H; ++k) z1[k] = dot_row<H>(z2, s.w2t + k * H) * gelu_erf_grad(z1[k]);
    // ---------------- B3: dX = dz1 W1 -> grid scatter (fused) or dx_out ----------------
    if (FUSED) {
      if (live && dg0) {
        const int gcols = (g.ncorner0 + 1) * g.C;
        for (int col = 0; col < gcols; ++col) scatter_column(g, dg0, dg1, ax, col, dot_row<H>(z1, s.w1t + col * H));
      }
    } else if (dx_out && live) {
      for (int col = 0; col < cin; ++col) dx_out[n * cin + col] = dot_row<H>(z1, s.w1t + col * H);
    }
    __syncthreads();
    // ---------------- (b) dW2 += dz2^T h1 ; db2 ----------------
    {
      float a2[SPAN2];
#pragma unroll
      for (int i = 0; i < SPAN2; ++i) a2[i] = 0.f;
      float ab = 0.f;
      for (int si = 0; si < T; ++si) {
        float d = sD[si * (H + 1) + jrow];
        const float* hr = sH + si * (H + 1) + k2base;
#pragma unroll
        for (int i = 0; i < SPAN2; ++i) a2[i] = fmaf(d, hr[i], a2[i]);
        ab += d;
      }
#pragma unroll
      for (int i = 0; i < SPAN2; ++i) atomicAdd(gm.w2 + jrow * H + k2base + i, a2[i]);
      if (part == 0) atomicAdd(gm.b2 + jrow, ab);
    }
    __syncthreads();
    // ---------------- (a) dW1 += dz1^T x~ ; db1 ----------------
#pragma unroll
    for (int k = 0; k < H; ++k) sD[tid * (H + 1) + k] = z1[k];
    __syncthreads();
    {
      float a1[SPAN1_MAX];
#pragma unroll
      for (int i = 0; i < SPAN1_MAX; ++i) a1[i] = 0.f;
      float ab = 0.f;
      for (int si = 0; si < T; ++si) {
        float d = sD[si * (H + 1) + jrow];
        const float* xr = sX + si * ldxs + k1base;
#pragma unroll
        for (int i = 0; i < SPAN1_MAX; ++i)
          if (i < span1 && k1base + i < cin) a1[i] = fmaf(d, xr[i], a1[i]);
        ab += d;
      }
#pragma unroll
      for (int i = 0; i < SPAN1_MAX; ++i)
        if (i < span1 && k1base + i < cin) atomicAdd(gm.w1 + jrow * cin + k1base + i, a1[i]);
      if (part == 0) atomicAdd(gm.b1 + jrow, ab);
    }
    __syncthreads();
  }
  if (FUSED && loss_sum) {
    float v = loss_local;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((tid & 31) == 0) sRed[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) atomicAdd(loss_sum, (sRed[0] + sRed[1]) + (sRed[2] + sRed[3]));
    if (g.metrics) {
      __syncthreads();
      v = sse8_local;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if ((tid & 31) == 0) sRed[tid >> 5] = v;
      __syncthreads();
      if (tid == 0) atomicAdd(loss_sum + 1, (sRed[0] + sRed[1]) + (sRed[2] + sRed[3]));
    }
  }
}


#endif
static size_t mlp_smem_bytes(int H, int cin, int cout) {
  return sizeof(float) * ((size_t)cin * H + (size_t)H * H + (size_t)cout * H + 2 * H + ((cout + 3) & ~3));
}

#if NIC_PART == 0
template <int FUSED>
static int launch_fwd_t(Handle* h, const DevGeom& g, const MlpDev& m, const float* g0, const float* g1,
                        const long long* origins, const float* x, long long ldx, long long N, void* out, int out_u8,
                        float* z1, float* z2, cudaStream_t st) {
  constexpr int H = NIC_H;
  size_t smem = mlp_smem_bytes(H, m.cin, m.cout);
  auto kern = mlp_forward_kernel<H, FUSED>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  long long blocks = (N + 127) / 128, cap = (long long)h->sms * 3;
  {
    KernelTimer timer(h, st);
    kern<<<(int)(blocks > cap ? cap : blocks), 128, smem, st>>>(g, m, g0, g1, origins, x, ldx, N, out, out_u8, z1, z2);
  }
  h->launches++;
  return (int)cudaGetLastError();
}

int NIC_CAT(launch_mlp_forward_f32_h, NIC_H)(Handle* h, const DevGeom* g, const MlpDev& m, const float* g0,
                                             const float* g1, const long long* origins, const float* x, long long ldx,
                                             long long N, void* out, int out_dtype, float* z1, float* z2,
                                             cudaStream_t st) {
  DevGeom dummy = {};
  int u8 = out_dtype == NIC_DT_U8;
  if (g) return launch_fwd_t<1>(h, *g, m, g0, g1, origins, x, ldx, N, out, u8, z1, z2, st);
  return launch_fwd_t<0>(h, dummy, m, g0, g1, origins, x, ldx, N, out, u8, z1, z2, st);
}
#else
template <int FUSED>
static int launch_bwd_t(Handle* h, const DevGeom& g, const MlpDev& m, const MlpGradDev& gm, const float* g0,
                        const float* g1, const long long* origins, const float* x, long long ldx, long long N,
                        const float* z1, const float* z2, const float* out, const float* dout, float* dx,
                        const float* targets, const float* noise, int noise_bits, unsigned long long seed,
                        unsigned long long step, float grad_scale, float* dg0, float* dg1, float* loss_sum,
                        float* out_save, cudaStream_t st) {
  constexpr int H = NIC_H;
  const int T = 128;
  int ldxs = m.cin | 1;
  size_t smem = mlp_smem_bytes(H, m.cin, m.cout) +
                sizeof(float) * ((size_t)T * ldxs + 2 * (size_t)T * (H + 1) + (size_t)T * NIC_MAX_COUT + 8);
  if (smem > 227 * 1024) return NIC_ERR_UNSUPPORTED;
  auto kern = mlp_backward_kernel<H, FUSED>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  long long ntiles = (N + T - 1) / T;
  int grid = (int)(ntiles < h->sms ? ntiles : h->sms);
  {
    KernelTimer timer(h, st);
    kern<<<grid, T, smem, st>>>(g, m, gm, g0, g1, origins, x, ldx, N, z1, z2, out, dout, dx, targets, noise,
                                noise_bits, seed, step, grad_scale, dg0, dg1, loss_sum, out_save);
  }
  h->launches++;
  return (int)cudaGetLastError();
}

int NIC_CAT(launch_mlp_backward_f32_h, NIC_H)(Handle* h, const MlpDev& m, const MlpGradDev& gm, const float* x,
                                              long long ldx, long long N, const float* z1, const float* z2,
                                              const float* out, const float* dout, float* dx, cudaStream_t st) {
  DevGeom dummy = {};
  return launch_bwd_t<0>(h, dummy, m, gm, nullptr, nullptr, nullptr, x, ldx, N, z1, z2, out, dout, dx, nullptr, nullptr,
                         0, 0, 0, 1.0f, nullptr, nullptr, nullptr, nullptr, st);
}

int NIC_CAT(launch_train_f32_h, NIC_H)(Handle* h, const DevGeom& g, const MlpDev& m, const MlpGradDev& gm,
                                       const float* g0, const float* g1, const long long* origins,
                                       const float* targets, const float* noise, int noise_bits,
                                       unsigned long long seed, unsigned long long step, float grad_scale, float* dg0,
                                       float* dg1, float* loss_sum, float* out_save, cudaStream_t st) {
  return launch_bwd_t<1>(h, g, m, gm, g0, g1, origins, nullptr, 0, g.N, nullptr, nullptr, nullptr, nullptr, nullptr,
                         targets, noise, noise_bits, seed, step, grad_scale, dg0, dg1, loss_sum, out_save, st);
}
#endif

}  // namespace nic
