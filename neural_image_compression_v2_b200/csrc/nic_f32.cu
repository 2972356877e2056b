// nic_f32.cu — fp32 gather (K1) and scatter (K4) kernels plus the dispatch to the per-width fp32 MLP objects
// (nic_f32_mlp.cu).  These are the reference-exact (1e-5 parity) paths; the tcgen05 path lives in nic_tc.cu.
// Reference: Projects/image_compression.py:71-211; Projects/fp_def.py:81-223.
#include "nic_internal.cuh"

namespace nic {

// ===================================================================================================== K1 gather
// One thread per element of X (flat index): perfectly coalesced stores, grid reads through L1/L2.
template <typename OutT>
__global__ void __launch_bounds__(256) gather_flat_kernel(DevGeom g, const float* __restrict__ g0,
                                                          const float* __restrict__ g1,
                                                          const long long* __restrict__ origins,
                                                          OutT* __restrict__ x, long long total) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    long long n = e / g.cin;
    int col = (int)(e - n * g.cin);
    Texel t = texel_of(g, n, origins);
    AxisCoord ax[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) ax[a] = axis_coord(t.p[a], g.step);
    store_as(x + e, gather_column(g, g0, g1, ax, col));
  }
}

// ===================================================================================================== K4 scatter
// Transpose of the gather: one thread per (sample, grid column); red.global.add.f32 into the grid gradients.
__global__ void __launch_bounds__(256) scatter_flat_kernel(DevGeom g, const float* __restrict__ dx,
                                                           const long long* __restrict__ origins,
                                                           float* __restrict__ dg0, float* __restrict__ dg1,
                                                           long long total) {
  const int gcols = (g.ncorner0 + 1) * g.C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    long long n = e / gcols;
    int col = (int)(e - n * gcols);
    float v = dx[n * g.cin + col];
    Texel t = texel_of(g, n, origins);
    AxisCoord ax[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) ax[a] = axis_coord(t.p[a], g.step);
    scatter_column(g, dg0, dg1, ax, col, v);
  }
}

// ===================================================================================================== crop targets
// random_crop_dataset's target gather (image_compression.py:44-47): image [Ci, S0, S1(, S2)] -> targets
// [num_crops, crop^D, Ci] (sample-major, channels last), crop origins read from the device `coord` tensor.
// (the target gather of random_crop_dataset lives in nic_data.cu: sample_crops_tile_kernel)

// ===================================================================================================== stand-alone PE
struct PeDiv { float d[NIC_MAX_PE]; };
__global__ void __launch_bounds__(256) pe_kernel(const float* __restrict__ coord, int dim, long long n, int PE, int kind,
                                                 PeDiv div, float* __restrict__ out) {
  long long total = (long long)dim * PE * n;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    long long row = e / n, i = e - row * n;
    int a = (int)(row / PE), r = (int)(row - (long long)a * PE);
    float u = coord[(long long)a * n + i];
    out[e] = kind == NIC_PE_TRIANGULAR ? pe_triangular(u, r, PE) : pe_sinusoidal(u, r, div.d);
  }
}

int launch_pe(Handle* h, const float* coord, int dim, long long n, int PE, int kind, const float* div_host, float* out,
              cudaStream_t st) {
  long long total = (long long)dim * PE * n;
  if (total == 0) return NIC_OK;
  PeDiv d = {};
  if (div_host) for (int i = 0; i < PE / 2 && i < NIC_MAX_PE; ++i) d.d[i] = div_host[i];
  long long blocks = (total + 255) / 256, cap = (long long)h->sms * 8;
  pe_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, st>>>(coord, dim, n, PE, kind, d, out);
  h->launches++;
  return (int)cudaGetLastError();
}

// ===================================================================================================== launchers
static int grid_for(long long work, int block, int sms, int per_sm) {
  long long b = (work + block - 1) / block;
  long long cap = (long long)sms * per_sm;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

int launch_gather(Handle* h, const DevGeom& g, const float* g0, const float* g1, const long long* origins, void* x,
                  int x_dtype, cudaStream_t st) {
  long long total = g.N * g.cin;
  if (total == 0) return NIC_OK;
  if (gather_tile_eligible(g, x) && !h->disable_fast2d) return launch_gather_tile(h, g, g0, g1, origins, x, x_dtype, st);
  int grid = grid_for(total, 256, h->sms, 8);
  KernelTimer timer(h, st);
  switch (x_dtype) {
    case NIC_DT_F32: gather_flat_kernel<float><<<grid, 256, 0, st>>>(g, g0, g1, origins, (float*)x, total); break;
    case NIC_DT_F16: gather_flat_kernel<__half><<<grid, 256, 0, st>>>(g, g0, g1, origins, (__half*)x, total); break;
    case NIC_DT_BF16:
      gather_flat_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(g, g0, g1, origins, (__nv_bfloat16*)x, total);
      break;
    default: return NIC_ERR_ARG;
  }
  h->launches++;
  return (int)cudaGetLastError();
}

int launch_scatter(Handle* h, const DevGeom& g, const float* dx, const long long* origins, float* dg0, float* dg1,
                   cudaStream_t st) {
  long long total = g.N * (long long)((g.ncorner0 + 1) * g.C);
  if (total == 0) return NIC_OK;
  scatter_flat_kernel<<<grid_for(total, 256, h->sms, 8), 256, 0, st>>>(g, dx, origins, dg0, dg1, total);
  h->launches++;
  return (int)cudaGetLastError();
}

// per-width entry points live in nic_f32_mlp.cu (one object per width and direction)
#define NIC_DECL(HH)                                                                                               \
  int launch_mlp_forward_f32_h##HH(Handle*, const DevGeom*, const MlpDev&, const float*, const float*,             \
                                   const long long*, const float*, long long, long long, void*, int, float*,       \
                                   float*, cudaStream_t);                                                          \
  int launch_mlp_backward_f32_h##HH(Handle*, const MlpDev&, const MlpGradDev&, const float*, long long, long long, \
                                    const float*, const float*, const float*, const float*, float*, cudaStream_t); \
  int launch_train_f32_h##HH(Handle*, const DevGeom&, const MlpDev&, const MlpGradDev&, const float*, const float*, \
                             const long long*, const float*, const float*, int, unsigned long long,                \
                             unsigned long long, float, float*, float*, float*, float*, cudaStream_t);
NIC_DECL(64)
NIC_DECL(32)
#undef NIC_DECL

int launch_mlp_forward_f32(Handle* h, const DevGeom* g, const MlpDev& m, const float* g0, const float* g1,
                           const long long* origins, const float* x, long long ldx, long long N, void* out,
                           int out_dtype, float* z1, float* z2, cudaStream_t st) {
  if (N == 0) return NIC_OK;
  if (m.hidden == 64) return launch_mlp_forward_f32_h64(h, g, m, g0, g1, origins, x, ldx, N, out, out_dtype, z1, z2, st);
  if (m.hidden == 32) return launch_mlp_forward_f32_h32(h, g, m, g0, g1, origins, x, ldx, N, out, out_dtype, z1, z2, st);
  return NIC_ERR_UNSUPPORTED;
}

int launch_mlp_backward_f32(Handle* h, const MlpDev& m, const MlpGradDev& gm, const float* x, long long ldx,
                            long long N, const float* z1, const float* z2, const float* out, const float* dout,
                            float* dx, cudaStream_t st) {
  if (N == 0) return NIC_OK;
  if (m.hidden == 64) return launch_mlp_backward_f32_h64(h, m, gm, x, ldx, N, z1, z2, out, dout, dx, st);
  if (m.hidden == 32) return launch_mlp_backward_f32_h32(h, m, gm, x, ldx, N, z1, z2, out, dout, dx, st);
  return NIC_ERR_UNSUPPORTED;
}

int launch_train_f32(Handle* h, const DevGeom& g, const MlpDev& m, const MlpGradDev& gm, const float* g0,
                     const float* g1, const long long* origins, const float* targets, const float* noise,
                     int noise_bits, unsigned long long seed, unsigned long long step, float grad_scale, float* dg0,
                     float* dg1, float* loss_sum, float* out_save, cudaStream_t st) {
  if (g.N == 0) return NIC_OK;
  if (m.hidden == 64)
    return launch_train_f32_h64(h, g, m, gm, g0, g1, origins, targets, noise, noise_bits, seed, step, grad_scale, dg0, dg1,
                                loss_sum, out_save, st);
  if (m.hidden == 32)
    return launch_train_f32_h32(h, g, m, gm, g0, g1, origins, targets, noise, noise_bits, seed, step, grad_scale, dg0, dg1,
                                loss_sum, out_save, st);
  return NIC_ERR_UNSUPPORTED;
}

}  // namespace nic
