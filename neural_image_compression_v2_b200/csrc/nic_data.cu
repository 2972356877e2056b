// nic_data.cu — the data front end of the training step, on the device:
//   * random crop sampler: crop origins drawn with Philox on the device + the target gather, ONE launch
//     (replaces the host RNG + per-crop slicing of random_crop_dataset, Projects/image_compression.py:26-50);
//   * mip-pyramid builder: transforms.Resize on the PIL image (Projects/image_compression.py:433-442, 462-469) =
//     PIL's two-pass separable BILINEAR resample with antialiasing on 8-bit pixels, reproduced bit for bit
//     (fixed-point coefficients, 8-bit intermediate), followed by ToTensor (u8 -> float32 / 255, [C, H, W]);
//   * method-2 atlas: frames of a [T, S, S, C] volume tiled into one 2-D image and back
//     (Projects/image_compression.py:453-460 and :410-421).
// All three are HBM-bound byte / float moves; none is on the per-step critical path except the sampler (24 B/sample).
#include <cmath>
#include <vector>

#include "nic_internal.cuh"

namespace nic {

// ===================================================================================================== crop sampler
// origin[a] of crop b = floor(U * (size[a] - crop[a] + 1)) with U the a-th 32-bit word of Philox4x32-10(key = seed,
// counter = (b, step)): uniform over the integers [0, size - crop], the distribution of torch.randint(0, size - crop + 1)
// (image_compression.py:40-41).  Every block recomputes the (few) origins into shared memory; block 0 publishes them.
// The same kernel serves nic_sample_crops (origins given by the caller).
// One block = 256 consecutive samples of the flattened [crop][i0][i1][i2] order: channel plane by channel plane the image
// reads are coalesced along the fast image axis, the [sample][channel] tile is transposed in shared memory and leaves as
// one linear, fully coalesced run of 256 * channels floats.  (One thread per OUTPUT element reads `channels` different
// planes from neighbouring threads: 32-byte sectors for 4 useful bytes — 32 us per step for a 9-channel 2048^2 stack.)
template <int RANDOM>
__global__ void __launch_bounds__(256) sample_crops_tile_kernel(const float* __restrict__ img, int dim, int ci, int s0, int s1,
                                                                int s2, int ncrops, int c0, int c1, int c2,
                                                                unsigned long long seed, unsigned long long step,
                                                                long long nsamples, const long long* __restrict__ origins_in,
                                                                long long* __restrict__ origins_out, float* __restrict__ out) {
  extern __shared__ int s_mem[];
  int* s_org = s_mem;                                        // [ncrops][3]
  float* tile = reinterpret_cast<float*>(s_mem + 3 * ncrops);      // [256][ci | 1]
  const int pitch = ci | 1;
  for (int b = threadIdx.x; b < ncrops; b += blockDim.x) {
    int o0, o1, o2 = 0;
    if (RANDOM) {
      const uint4 r = philox4x32(seed, step, (unsigned long long)b);
      o0 = (int)__umulhi(r.x, (unsigned)(s0 - c0 + 1));
      o1 = (int)__umulhi(r.y, (unsigned)(s1 - c1 + 1));
      if (dim == 3) o2 = (int)__umulhi(r.z, (unsigned)(s2 - c2 + 1));
      if (blockIdx.x == 0) {
        origins_out[(long long)b * dim] = o0;
        origins_out[(long long)b * dim + 1] = o1;
        if (dim == 3) origins_out[(long long)b * dim + 2] = o2;
      }
    } else {                                                 // caller's origins, clamped into the image
      o0 = clampi((int)origins_in[(long long)b * dim], 0, s0 - c0);
      o1 = clampi((int)origins_in[(long long)b * dim + 1], 0, s1 - c1);
      if (dim == 3) o2 = clampi((int)origins_in[(long long)b * dim + 2], 0, s2 - c2);
    }
    s_org[3 * b] = o0;
    s_org[3 * b + 1] = o1;
    s_org[3 * b + 2] = o2;
  }
  __syncthreads();
  const long long per = (long long)c0 * c1 * c2, plane = (long long)s0 * s1 * s2;
  const long long nchunks = (nsamples + 255) / 256;
  for (long long chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const long long n = chunk * 256 + threadIdx.x;
    if (n < nsamples) {
      int b, i0, i1, i2;
      if (nsamples < (1ll << 31)) {            // 32-bit index math (three 64-bit divisions were most of this kernel's instructions)
        const unsigned un = (unsigned)n, uper = (unsigned)per;
        const unsigned ub = un / uper;
        unsigned r = un - ub * uper;
        b = (int)ub;
        if (c2 == 1) {
          i2 = 0;
        } else {
          const unsigned q = r / (unsigned)c2;
          i2 = (int)(r - q * (unsigned)c2);
          r = q;
        }
        const unsigned q1 = r / (unsigned)c1;
        i1 = (int)(r - q1 * (unsigned)c1);
        i0 = (int)q1;
      } else {
        b = (int)(n / per);
        long long r = n - (long long)b * per;
        i2 = (int)(r % c2);
        r /= c2;
        i1 = (int)(r % c1);
        i0 = (int)(r / c1);
      }
      const float* p = img + ((long long)(s_org[3 * b] + i0) * s1 + (s_org[3 * b + 1] + i1)) * s2 + (s_org[3 * b + 2] + i2);
      for (int c = 0; c < ci; ++c) tile[threadIdx.x * pitch + c] = __ldg(p + c * plane);
    }
    __syncthreads();
    const long long left = nsamples - chunk * 256;
    const int cnt = (int)(left < 256 ? left : 256) * ci;
    float* dst = out + chunk * 256 * ci;
    for (int j = threadIdx.x; j < cnt; j += 256) {
      const int sidx = j / ci;
      dst[j] = tile[sidx * pitch + (j - sidx * ci)];
    }
    __syncthreads();
  }
}

static int launch_sample_tile(Handle* h, const float* img, int dim, int ci, const int* size, int ncrops, const int* crop,
                              int random, unsigned long long seed, unsigned long long step, const long long* origins_in,
                              long long* origins_out, float* out, cudaStream_t st) {
  const int s2 = dim == 3 ? size[2] : 1, c2 = dim == 3 ? crop[2] : 1;
  const long long nsamples = (long long)ncrops * crop[0] * crop[1] * c2;
  if (ncrops == 0) return NIC_OK;
  const size_t smem = (size_t)ncrops * 3 * sizeof(int) + (size_t)256 * (ci | 1) * sizeof(float);
  if (smem > 48 * 1024) return NIC_ERR_UNSUPPORTED;
  long long blocks = (nsamples + 255) / 256, cap = (long long)h->sms * 8;
  if (blocks < 1) blocks = 1;
  const int grid = (int)(blocks > cap ? cap : blocks);
  if (random)
    sample_crops_tile_kernel<1><<<grid, 256, smem, st>>>(img, dim, ci, size[0], size[1], s2, ncrops, crop[0], crop[1], c2, seed, step,
                                                         nsamples, nullptr, origins_out, out);
  else
    sample_crops_tile_kernel<0><<<grid, 256, smem, st>>>(img, dim, ci, size[0], size[1], s2, ncrops, crop[0], crop[1], c2, 0, 0,
                                                         nsamples, origins_in, nullptr, out);
  h->launches++;
  return (int)cudaGetLastError();
}

int launch_sample_crops_random(Handle* h, const float* img, int dim, int ci, const int* size, int ncrops, const int* crop,
                               unsigned long long seed, unsigned long long step, long long* origins_out, float* out,
                               cudaStream_t st) {
  return launch_sample_tile(h, img, dim, ci, size, ncrops, crop, 1, seed, step, nullptr, origins_out, out, st);
}

int launch_sample_crops(Handle* h, const float* img, int dim, int ci, const int* size, const long long* origins, int ncrops,
                        const int* crop, float* out, cudaStream_t st) {
  return launch_sample_tile(h, img, dim, ci, size, ncrops, crop, 0, 0, 0, origins, nullptr, out, st);
}

// ===================================================================================================== PIL bilinear resample
// Pillow's ImagingResample for 8-bit images (src/libImaging/Resample.c), which is what transforms.Resize does to the PIL
// image the reference opens: per output index a window [xmin, xmin + n) of input pixels and n normalised triangle-filter
// weights (support = max(scale, 1)), converted to fixed point with PRECISION_BITS = 22; the horizontal pass rounds to
// 8 bits before the vertical pass.  The coefficient tables are evaluated on the host in double precision with the same
// expressions in the same order as Pillow (so they round identically) and uploaded; the passes run on the device.
constexpr int PIL_PRECISION_BITS = 32 - 8 - 2;

struct ResampleTable {
  int ksize;
  std::vector<int> bounds;      // [out][2]: xmin, count
  std::vector<int> kk;          // [out][ksize] fixed-point weights
};

static double pil_bilinear_filter(double x) {
  if (x < 0.0) x = -x;
  if (x < 1.0) return 1.0 - x;
  return 0.0;
}

static ResampleTable pil_precompute_coeffs(int in_size, int out_size) {
  ResampleTable t;
  const double in0 = 0.0, in1 = (double)in_size;
  double scale, filterscale;
  filterscale = scale = (in1 - in0) / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 1.0 * filterscale;            // bilinear: support 1.0
  t.ksize = (int)ceil(support) * 2 + 1;
  t.bounds.resize((size_t)out_size * 2);
  t.kk.assign((size_t)out_size * t.ksize, 0);
  std::vector<double> k((size_t)t.ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = in0 + (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    int x;
    for (x = 0; x < xmax; ++x) {
      const double w = pil_bilinear_filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (; x < t.ksize; ++x) k[x] = 0.0;
    t.bounds[2 * xx] = xmin;
    t.bounds[2 * xx + 1] = xmax;
    for (x = 0; x < t.ksize; ++x) {
      if (k[x] < 0) t.kk[(size_t)xx * t.ksize + x] = (int)(-0.5 + k[x] * (1 << PIL_PRECISION_BITS));
      else t.kk[(size_t)xx * t.ksize + x] = (int)(0.5 + k[x] * (1 << PIL_PRECISION_BITS));
    }
  }
  return t;
}

__device__ __forceinline__ uint8_t pil_clip8(int v) {
  v >>= PIL_PRECISION_BITS;                            // arithmetic shift, as Pillow's lookup index
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// One pass along an axis of an [A, B, C]-shaped byte image (C interleaved channels).  VERTICAL = 0: resample axis B
// (length b_in -> b_out); VERTICAL = 1: resample axis A (a_in -> a_out).  out_f (optional, last pass only): the ToTensor
// result [C, A_out, B_out] = u8 / 255 in float32.
template <int VERTICAL>
__global__ void __launch_bounds__(256) pil_resample_kernel(const uint8_t* __restrict__ src, int a_in, int b_in, int C, int a_out,
                                                           int b_out, const int* __restrict__ bounds, const int* __restrict__ kk,
                                                           int ksize, uint8_t* __restrict__ dst, float* __restrict__ out_f) {
  const long long total = (long long)a_out * b_out * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const long long ab = e / C;
    const int bb = (int)(ab % b_out), aa = (int)(ab / b_out);
    const int o = VERTICAL ? aa : bb;
    const int lo = bounds[2 * o], n = bounds[2 * o + 1];
    const int* k = kk + (long long)o * ksize;
    int ss = 1 << (PIL_PRECISION_BITS - 1);
    if (VERTICAL) {
      const uint8_t* p = src + ((long long)lo * b_in + bb) * C + c;
      for (int i = 0; i < n; ++i) ss += (int)p[(long long)i * b_in * C] * k[i];
    } else {
      const uint8_t* p = src + ((long long)aa * b_in + lo) * C + c;
      for (int i = 0; i < n; ++i) ss += (int)p[(long long)i * C] * k[i];
    }
    const uint8_t v = pil_clip8(ss);
    if (dst) dst[e] = v;
    if (out_f) out_f[((long long)c * a_out + aa) * b_out + bb] = __fdiv_rn((float)v, 255.0f);
  }
}

// u8 image [H, W, C] -> float32 [C, H, W] / 255 (ToTensor of an image that needs no resampling: mip 0)
__global__ void __launch_bounds__(256) to_tensor_kernel(const uint8_t* __restrict__ src, int H, int W, int C, uint8_t* __restrict__ dst,
                                                        float* __restrict__ out_f) {
  const long long total = (long long)H * W * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const long long hw = e / C;
    const uint8_t v = src[e];
    if (dst) dst[e] = v;
    if (out_f) out_f[(long long)c * H * W + hw] = __fdiv_rn((float)v, 255.0f);
  }
}

static int ensure_bytes(void** p, size_t* have, size_t need) {
  if (*have >= need && *p) return NIC_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *have = 0;
  cudaError_t e = cudaMalloc(p, need);
  if (e != cudaSuccess) return NIC_ERR_SCRATCH;
  *have = need;
  return NIC_OK;
}

int launch_resize_bilinear_u8(Handle* h, const uint8_t* src, int H, int W, int C, int out_h, int out_w, uint8_t* dst_u8,
                              float* dst_f32, cudaStream_t st) {
  const bool need_h = out_w != W, need_v = out_h != H;
  auto grid_for = [&](long long total) {
    long long blocks = (total + 255) / 256, cap = (long long)h->sms * 16;
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
  };
  if (!need_h && !need_v) {
    to_tensor_kernel<<<grid_for((long long)H * W * C), 256, 0, st>>>(src, H, W, C, dst_u8, dst_f32);
    h->launches++;
    return (int)cudaGetLastError();
  }
  // coefficient tables (host, double) -> device scratch: [bounds_h | kk_h | bounds_v | kk_v], then the 8-bit intermediate
  ResampleTable th, tv;
  if (need_h) th = pil_precompute_coeffs(W, out_w);
  if (need_v) tv = pil_precompute_coeffs(H, out_h);
  const size_t nh = th.bounds.size() + th.kk.size(), nv = tv.bounds.size() + tv.kk.size();
  const size_t tab_bytes = ((nh + nv) * sizeof(int) + 255) & ~(size_t)255;
  const size_t tmp_bytes = need_h && need_v ? (size_t)H * out_w * C : 0;
  int rc = ensure_bytes(&h->data_scratch, &h->data_scratch_bytes, tab_bytes + tmp_bytes + 256);
  if (rc) return rc;
  std::vector<int> host(nh + nv);
  size_t o = 0;
  for (int v : th.bounds) host[o++] = v;
  for (int v : th.kk) host[o++] = v;
  for (int v : tv.bounds) host[o++] = v;
  for (int v : tv.kk) host[o++] = v;
  int* dtab = (int*)h->data_scratch;
  // pageable source: the runtime stages the bytes before cudaMemcpyAsync returns, so `host` may go out of scope
  cudaError_t e = cudaMemcpyAsync(dtab, host.data(), host.size() * sizeof(int), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaStreamSynchronize(st);          // (an init-time call; keeps the scratch table safe against re-entry)
  if (e != cudaSuccess) return (int)e;
  const int* bh = dtab;
  const int* kh = dtab + th.bounds.size();
  const int* bv = dtab + nh;
  const int* kv = bv + tv.bounds.size();
  uint8_t* tmp = (uint8_t*)h->data_scratch + tab_bytes;
  const uint8_t* cur = src;
  if (need_h) {       // Pillow runs the horizontal pass first
    uint8_t* d = need_v ? tmp : dst_u8;
    pil_resample_kernel<0><<<grid_for((long long)H * out_w * C), 256, 0, st>>>(cur, H, W, C, H, out_w, bh, kh, th.ksize, d,
                                                                               need_v ? nullptr : dst_f32);
    h->launches++;
    cur = tmp;
  }
  if (need_v) {
    const int win = need_h ? out_w : W;
    pil_resample_kernel<1><<<grid_for((long long)out_h * win * C), 256, 0, st>>>(cur, H, win, C, out_h, win, bv, kv, tv.ksize,
                                                                                 dst_u8, dst_f32);
    h->launches++;
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  return (int)cudaStreamSynchronize(st);  // the scratch tables / intermediate may be reused by the next call
}

// ===================================================================================================== method-2 atlas
// frames [T, S, S, C] (u8) <-> atlas [A, A, C]: frame i sits at rows [r S, (r+1) S), columns [q S, (q+1) S) with
// r = i / (A / S), q = i % (A / S)  (image_compression.py:453-460; the inverse at :413-419).  Atlas cells without a frame
// are zero (np.zeros).
template <int UNPACK>
__global__ void __launch_bounds__(256) atlas_kernel(uint8_t* __restrict__ frames, uint8_t* __restrict__ atlas, int T, int S, int C,
                                                    int A) {
  const int per_row = A / S;
  const long long total = (long long)A * A * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const long long yx = e / C;
    const int x = (int)(yx % A), y = (int)(yx / A);
    const int r = y / S, q = x / S;
    const long long i = (long long)r * per_row + q;
    const bool has = q < per_row && i < T;
    const long long f = ((i * S + (y - r * S)) * S + (x - q * S)) * C + c;
    if (UNPACK) {
      if (has) frames[f] = atlas[e];
    } else {
      atlas[e] = has ? frames[f] : (uint8_t)0;
    }
  }
}

int launch_atlas(Handle* h, uint8_t* frames, uint8_t* atlas, int T, int S, int C, int A, int unpack, cudaStream_t st) {
  const long long total = (long long)A * A * C;
  if (total == 0) return NIC_OK;
  long long blocks = (total + 255) / 256, cap = (long long)h->sms * 16;
  const int grid = (int)(blocks > cap ? cap : blocks);
  if (unpack) atlas_kernel<1><<<grid, 256, 0, st>>>(frames, atlas, T, S, C, A);
  else atlas_kernel<0><<<grid, 256, 0, st>>>(frames, atlas, T, S, C, A);
  h->launches++;
  return (int)cudaGetLastError();
}

}  // namespace nic
