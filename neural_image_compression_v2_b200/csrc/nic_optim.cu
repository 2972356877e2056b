// nic_optim.cu — K5 fused multi-tensor Adam (+ clamp, + grad zeroing) and K6 quantise / pack / unpack kernels.
// Reference: torch.optim.Adam as configured at Projects/image_compression.py:361-365, fp_quantize_clamp
// (Projects/fp_def.py:227-232) and the quantiser family of Projects/models.py:29-71.
// HBM-bound elementwise work: 28 B/param for Adam (read p,g,m,v; write p,m,v), vectorised 16-byte accesses.
#include <cmath>

#include "nic_internal.cuh"

namespace nic {

#define NIC_ADAM_BATCH 24
struct AdamBatch {
  NicAdamTensor t[NIC_ADAM_BATCH];
  float step_size[NIC_ADAM_BATCH];     // lr / (1 - beta1^t) and sqrt(1 - beta2^t): evaluated in double ON THE HOST, as torch does
  float bc2_sqrt[NIC_ADAM_BATCH];      // (a double pow() per thread made this kernel 11 us for 260 k parameters)
  int count;
  float beta1, beta2, eps, grad_scale;
  int zero_grad;
  float* loss_sum;      // optional: loss_out[0] = loss_sum[0] * loss_scale; loss_sum[0] = 0  (done by one thread)
  float* loss_out;
  float loss_scale;
  int metrics;          // loss_sum / loss_out hold two values: [0] squared error, [1] squared error of the 8-bit outputs
};

__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v, float beta1, float beta2, float eps,
                                         float gscale, float step_size, float bc2_sqrt, int clamp, float lo, float hi) {
  float gr = g * gscale;
  m = m + (gr - m) * (1.0f - beta1);                 // exp_avg.lerp_(grad, 1 - beta1)
  v = v * beta2 + (1.0f - beta2) * gr * gr;          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  float denom = sqrtf(v) / bc2_sqrt + eps;           // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
  float np = p - step_size * (m / denom);            // param.addcdiv_(exp_avg, denom, value=-step_size)
  if (clamp) np = fminf(fmaxf(np, lo), hi);          // fp_quantize_clamp after the step
  p = np;
}

__global__ void __launch_bounds__(256) adam_kernel(AdamBatch b) {
  pdl_wait();                    // launched programmatically behind the kernel that produced the gradients
  if (b.loss_sum && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    b.loss_out[0] = b.loss_sum[0] * b.loss_scale;
    b.loss_sum[0] = 0.f;
    if (b.metrics) {
      b.loss_out[1] = b.loss_sum[1] * b.loss_scale;
      b.loss_sum[1] = 0.f;
    }
  }
  const NicAdamTensor& t = b.t[blockIdx.y];
  const float step_size = b.step_size[blockIdx.y], bc2_sqrt = b.bc2_sqrt[blockIdx.y];
  const long long n4 = ((((uintptr_t)t.p | (uintptr_t)t.g | (uintptr_t)t.m | (uintptr_t)t.v) & 15) == 0) ? t.numel / 4 : 0;
  float4* p4 = reinterpret_cast<float4*>(t.p);
  float4* g4 = reinterpret_cast<float4*>(t.g);
  float4* m4 = reinterpret_cast<float4*>(t.m);
  float4* v4 = reinterpret_cast<float4*>(t.v);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 p = p4[i], g = g4[i], m = m4[i], v = v4[i];
    adam_one(p.x, g.x, m.x, v.x, b.beta1, b.beta2, b.eps, b.grad_scale, step_size, bc2_sqrt, t.clamp, t.clamp_lo, t.clamp_hi);
    adam_one(p.y, g.y, m.y, v.y, b.beta1, b.beta2, b.eps, b.grad_scale, step_size, bc2_sqrt, t.clamp, t.clamp_lo, t.clamp_hi);
    adam_one(p.z, g.z, m.z, v.z, b.beta1, b.beta2, b.eps, b.grad_scale, step_size, bc2_sqrt, t.clamp, t.clamp_lo, t.clamp_hi);
    adam_one(p.w, g.w, m.w, v.w, b.beta1, b.beta2, b.eps, b.grad_scale, step_size, bc2_sqrt, t.clamp, t.clamp_lo, t.clamp_hi);
    p4[i] = p; m4[i] = m; v4[i] = v;
    if (b.zero_grad) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < t.numel; i += stride) {
    float p = t.p[i], g = t.g[i], m = t.m[i], v = t.v[i];
    adam_one(p, g, m, v, b.beta1, b.beta2, b.eps, b.grad_scale, step_size, bc2_sqrt, t.clamp, t.clamp_lo, t.clamp_hi);
    t.p[i] = p; t.m[i] = m; t.v[i] = v;
    if (b.zero_grad) t.g[i] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------------ fused exchange + Adam
// The exchange step of data-parallel training inside the optimiser kernel (see nic.h: NicExchange).  Every block
//   1. (one warp of block 0,0) pushes this rank's token into every peer's flag array: the gradients were written by the
//      previous kernel of the stream, a system-scope release store makes them visible to the peers before the flag;
//   2. waits (local acquire loads; timeout NIC_OPT_EXCHANGE_TIMEOUT_MS, default 10 s, FATAL: see the kernel) until its own
//      flag array shows the token of every rank;
//   3. forms each gradient element as the sum of the `world` peer buffers in RANK ORDER — identical bits on every rank —
//      with cache-volatile 16-byte loads (L1 is not coherent with peer writes), and applies Adam;
//   4. (row 0 of the grid) clears the other-parity buffer of this rank for its next use.
struct XchDev {
  int world, rank;
  unsigned token;
  const float* peer_flat[NIC_MAX_PEERS];
  unsigned* peer_flag[NIC_MAX_PEERS];
  float* zero_buf;
  long long zero_numel;
  unsigned* err;              // device word, sticky: an exchange timed out; no update is applied from then on
  unsigned* host_err;         // the same flag in mapped pinned host memory: the next API call reads it without a sync
  unsigned* go;               // local word: block (0,0) releases the other blocks once every peer has arrived
  unsigned go_token;          // ... by storing this launch's sequence number (per handle, monotonic)
  long long timeout_cycles;   // SM clock cycles (clock64: %globaltimer reads by every block cost ~20 us per launch)
  // Work decomposition: "virtual blocks" of 256 items (an item = one float4 of a tensor, or one float of its unaligned tail /
  // of an unaligned tensor), tensor i owning virtual blocks [vb_start[i], vb_start[i + 1]); the blocks after vb_start[count]
  // clear zero_buf.  Every thread handles ONE item of ONE tensor per round, so the peer loads of ALL tensors are in flight at
  // once (a block that walks the tensors one after the other pays the ~2 us NVLink round trip once per tensor: +17 us per
  // step on 2 GPUs).
  int vb_start[NIC_ADAM_BATCH + 2];
  unsigned char vec[NIC_ADAM_BATCH];
  const float* loss_sum;      // inside peer_flat[rank]
  int dbg;                    // knock-outs (timing experiments only): bit 5 no flag wait, bit 6 read the own buffer only
  // Sliced (two-phase) mode for buffers too large to read world times: rank r sums slice r of every buffer and writes
  // the sums back into ALL buffers; after a second handshake every rank applies Adam from its own, now reduced, buffer.
  int sliced;
  long long flat_numel;       // floats in one flat buffer (multiple of 4)
  unsigned* go2;              // local release word of the second handshake
  unsigned* arrive;           // local counter: blocks of this grid that finished their part of the slice
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void xch_fail(const XchDev& x) {
  atomicExch(x.err, 1u);
  if (x.host_err) {
    *reinterpret_cast<volatile unsigned*>(x.host_err) = 1u;
    __threadfence_system();
  }
}

// A timeout is FATAL for the exchange: the flag is sticky, this launch and every later one leave the parameters, the
// Adam state and the loss untouched, and the next nic_adam_step_exchange / nic_exchange_status call on the handle returns
// NIC_ERR_EXCHANGE — the replicas stay consistent (nobody applied a partial sum) until the caller re-synchronises them.
// The grid is sized to the resident block count (launch_adam_exchange), so block (0,0) is always scheduled.
// One cross-rank handshake of the grid (see the kernel's header): block 0's first warp exchanges tokens with the peers
// through flag slots [slot0, slot0 + world) and releases the other blocks through `go`; on return (after the caller's
// __syncthreads) every rank's writes before ITS handshake are visible to this block's cache-volatile loads.
__device__ __forceinline__ void xch_handshake(const XchDev& x, int slot0, unsigned* go) {
  if (blockIdx.x == 0) {
    // ONE warp of the grid talks to the peers.  Flags are PUSHED: lane p stores this rank's token into slot `rank` of
    // peer p's flag array (a fire-and-forget NVLink store), then polls slot p of the LOCAL array — no remote polling,
    // and all peers are awaited in parallel.  Lane 0 then releases this rank's other blocks through a local word.
    if (threadIdx.x < 32) {
      const int p = threadIdx.x;
      bool bad = ld_acquire_gpu(x.err) != 0u;
      if (p < x.world && !bad) {          // a rank that has already failed stays silent: its peers time out too
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(x.peer_flag[p] + slot0 + x.rank), "r"(x.token) : "memory");
        const unsigned* mine = x.peer_flag[x.rank] + slot0 + p;
        const long long t0 = clock64();
        while (!bad && (int)(ld_acquire_sys(mine) - x.token) < 0) {
          if (clock64() - t0 > x.timeout_cycles) bad = true;     // the peer is gone
        }
      }
      bad = __any_sync(0xffffffffu, bad);
      if (p == 0) {
        if (bad) xch_fail(x);
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(go), "r"(x.go_token) : "memory");
      }
    }
  } else if (threadIdx.x == 0) {
    const long long t0 = clock64();
    while ((int)(ld_acquire_gpu(go) - x.go_token) < 0) {
      if (clock64() - t0 > x.timeout_cycles + 2000000000ll) {
        xch_fail(x);
        break;
      }
    }
  }
}

__global__ void __launch_bounds__(256) adam_exchange_kernel(AdamBatch b, XchDev x) {
  pdl_wait();
  __shared__ unsigned s_err;
  if (!(x.dbg & 32)) xch_handshake(x, 0, x.go);
  if (x.sliced) {
    // ---- phase 1: reduce-scatter by peer reads + all-gather by peer writes.  Rank r owns the float4 items
    // [n4 * r / world, n4 * (r + 1) / world) of the flat buffer: it forms their sum over the ranks IN RANK ORDER and stores
    // it into every rank's buffer, so that all replicas consume the same bits.  Nobody else touches slice r of any buffer
    // between the two handshakes (slice q of this rank's buffer is read and rewritten by rank q alone).
    __syncthreads();
    if (threadIdx.x == 0) s_err = ld_acquire_gpu(x.err);
    __syncthreads();
    if (!s_err) {
      const long long n4 = x.flat_numel / 4, lo = n4 * x.rank / x.world, hi = n4 * (x.rank + 1) / x.world;
      // Consecutive 16-byte items per warp (512-byte runs per peer); UN items per thread and round keep UN * world
      // independent peer loads in flight per thread (knock-out bit 7 of NIC_OPT_DEBUG_KNOCKOUT: one item per round).
      constexpr int UN = 4;
      const long long stride = (long long)gridDim.x * 256;
      const int un = (x.dbg & 128) ? 1 : UN;
      for (long long it0 = lo + (long long)blockIdx.x * 256 + threadIdx.x; it0 < hi; it0 += stride * un) {
        float4 g[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < x.world; ++r) {
          const float4* src = reinterpret_cast<const float4*>(x.peer_flat[r]);
#pragma unroll
          for (int u = 0; u < UN; ++u) {
            const long long it = it0 + u * stride;
            if (u < un && it < hi) {
              const float4 q = __ldcv(src + it);
              g[u].x += q.x; g[u].y += q.y; g[u].z += q.z; g[u].w += q.w;
            }
          }
        }
        for (int r = 0; r < x.world; ++r) {
          float4* dst = reinterpret_cast<float4*>(const_cast<float*>(x.peer_flat[r]));
#pragma unroll
          for (int u = 0; u < UN; ++u) {
            const long long it = it0 + u * stride;
            if (u < un && it < hi) dst[it] = g[u];
          }
        }
      }
    }
    // every block of this grid has stored its sums -> second handshake -> the whole reduced gradient is in the LOCAL buffer
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      if (blockIdx.x != 0) {
        atomicAdd(x.arrive, 1u);
      } else {
        const long long t0 = clock64();
        while (ld_acquire_gpu(x.arrive) < gridDim.x - 1) {
          if (clock64() - t0 > x.timeout_cycles + 2000000000ll) {
            xch_fail(x);
            break;
          }
        }
        *x.arrive = 0u;            // for the next launch (nobody adds again before this grid has completed)
        __threadfence_system();
      }
    }
    __syncthreads();
    xch_handshake(x, NIC_MAX_PEERS, x.go2);
  }
  __syncthreads();
  if (threadIdx.x == 0) s_err = ld_acquire_gpu(x.err);
  __syncthreads();
  if (s_err) return;             // sticky failure: nothing is updated (see above)
  // one-shot: every element is the sum over the peers' buffers; sliced: the own buffer already holds the reduced gradient
  const bool own = x.sliced || (x.dbg & 64);
  const int nsrc = x.sliced ? 1 : x.world;
  if (b.loss_out && blockIdx.x == 0 && threadIdx.x == 0) {
    const long long off = x.loss_sum - x.peer_flat[x.rank];
    float s = 0.f;
    for (int r = 0; r < nsrc; ++r) s += __ldcv(x.peer_flat[own ? x.rank : r] + off);
    b.loss_out[0] = s * b.loss_scale;
    if (b.metrics) {
      float s8 = 0.f;
      for (int r = 0; r < nsrc; ++r) s8 += __ldcv(x.peer_flat[own ? x.rank : r] + off + 1);
      b.loss_out[1] = s8 * b.loss_scale;
    }
  }
  const int total_vb = x.vb_start[b.count + 1];
  for (int vb = blockIdx.x; vb < total_vb; vb += gridDim.x) {
    if (vb >= x.vb_start[b.count]) {               // zero_buf: 16-byte aligned, multiple of 4 floats (FusedTrainer's layout)
      const long long i = (long long)(vb - x.vb_start[b.count]) * 256 + threadIdx.x;
      if (i < x.zero_numel / 4) reinterpret_cast<float4*>(x.zero_buf)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    int ti = 0;
    while (vb >= x.vb_start[ti + 1]) ++ti;
    const NicAdamTensor& t = b.t[ti];
    const float step_size = b.step_size[ti], bc2_sqrt = b.bc2_sqrt[ti];
    const long long goff = t.g - x.peer_flat[x.rank];          // this tensor's offset inside every rank's buffer
    const long long n4 = x.vec[ti] ? t.numel / 4 : 0;
    const long long item = (long long)(vb - x.vb_start[ti]) * 256 + threadIdx.x;
    if (item < n4) {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < nsrc; ++r) {
        const float4 q = __ldcv(reinterpret_cast<const float4*>(x.peer_flat[own ? x.rank : r] + goff) + item);
        g.x += q.x; g.y += q.y; g.z += q.z; g.w += q.w;
      }
      float4* p4 = reinterpret_cast<float4*>(t.p);
      float4* m4 = reinterpret_cast<float4*>(t.m);
      float4* v4 = reinterpret_cast<float4*>(t.v);
      float4 p = p4[item], m = m4[item], v = v4[item];
      adam_one(p.x, g.x, m.x, v.x, b.beta1, b.beta2, b.eps, b.grad_scale, step_size, bc2_sqrt, t.clamp, t.clamp_lo, t.clamp_hi);
      adam_one(p.y, g.y, m.y, v.y, b.beta1, b.beta2, b.eps, b.grad_scale, step_size, bc2_sqrt, t.clamp, t.clamp_lo, t.clamp_hi);
      adam_one(p.z, g.z, m.z, v.z, b.beta1, b.beta2, b.eps, b.grad_scale, step_size, bc2_sqrt, t.clamp, t.clamp_lo, t.clamp_hi);
      adam_one(p.w, g.w, m.w, v.w, b.beta1, b.beta2, b.eps, b.grad_scale, step_size, bc2_sqrt, t.clamp, t.clamp_lo, t.clamp_hi);
      p4[item] = p; m4[item] = m; v4[item] = v;
    } else {
      const long long i = n4 * 4 + (item - n4);
      if (i < t.numel) {
        float g = 0.f;
        for (int r = 0; r < nsrc; ++r) g += __ldcv(x.peer_flat[own ? x.rank : r] + goff + i);
        float p = t.p[i], m = t.m[i], v = t.v[i];
        adam_one(p, g, m, v, b.beta1, b.beta2, b.eps, b.grad_scale, step_size, bc2_sqrt, t.clamp, t.clamp_lo, t.clamp_hi);
        t.p[i] = p; t.m[i] = m; t.v[i] = v;
      }
    }
  }
}

int launch_adam_exchange(Handle* h, const NicAdamTensor* tensors, int count, float beta1, float beta2, float eps,
                         float grad_scale, const NicExchange& xc, const float* loss_sum, float* loss_out, float loss_scale,
                         cudaStream_t st) {
  if (count > NIC_ADAM_BATCH) return NIC_ERR_UNSUPPORTED;      // one launch: the flag protocol runs once per step
  if (!h->xch_err) {          // two words: [0] sticky timeout flag, [1] the local release word; + the host-visible copy of [0]
    cudaError_t e = cudaMalloc(&h->xch_err, 4 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemsetAsync(h->xch_err, 0, 4 * sizeof(unsigned), st);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&h->xch_host_err, sizeof(unsigned), cudaHostAllocMapped);
    if (e == cudaSuccess) {
      *h->xch_host_err = 0u;
      e = cudaHostGetDevicePointer((void**)&h->xch_host_err_dev, h->xch_host_err, 0);
    }
    if (e == cudaSuccess) {
      int occ = 0;
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, adam_exchange_kernel, 256, 0);
      h->xch_resident_blocks = (occ > 0 ? occ : 1) * h->sms;
    }
    if (e != cudaSuccess) return (int)e;
  }
  if (*reinterpret_cast<volatile unsigned*>(h->xch_host_err)) return NIC_ERR_EXCHANGE;     // sticky until nic_exchange_status
  AdamBatch b;
  b.loss_sum = nullptr;
  b.loss_out = loss_out;
  b.loss_scale = loss_scale;
  b.metrics = h->step_metrics;
  b.count = count;
  b.beta1 = beta1; b.beta2 = beta2; b.eps = eps; b.grad_scale = grad_scale; b.zero_grad = 0;
  for (int i = 0; i < count; ++i) {
    b.t[i] = tensors[i];
    const double bc1 = 1.0 - pow((double)beta1, (double)b.t[i].t), bc2 = 1.0 - pow((double)beta2, (double)b.t[i].t);
    b.step_size[i] = (float)((double)b.t[i].lr / bc1);
    b.bc2_sqrt[i] = (float)sqrt(bc2);
  }
  XchDev x;
  memset(&x, 0, sizeof(x));
  x.world = xc.world;
  x.rank = xc.rank;
  x.token = xc.token;
  for (int r = 0; r < xc.world; ++r) {
    x.peer_flat[r] = xc.peer_flat[r];
    x.peer_flag[r] = (unsigned*)xc.peer_flag[r];
  }
  x.zero_buf = xc.zero_buf;
  x.zero_numel = xc.zero_numel;
  x.err = h->xch_err;
  x.host_err = h->xch_host_err_dev;
  x.go = h->xch_err + 1;
  x.go_token = ++h->xch_seq;
  x.sliced = xc.reserved == NIC_EXCHANGE_SLICED;
  x.flat_numel = xc.zero_numel;
  x.go2 = h->xch_err + 2;
  x.arrive = h->xch_err + 3;
  if (x.sliced && (xc.zero_numel & 3)) return NIC_ERR_ARG;
  x.timeout_cycles = (long long)(h->xch_timeout_ms > 0 ? h->xch_timeout_ms : 10000) * 2000000ll;      // at <= 2 GHz: >= the requested time
  x.dbg = h->debug_flags;
  x.loss_sum = loss_sum;
  // Every block but (0,0) spins until block (0,0) has met the peers, so ALL blocks must be co-resident: the grid is
  // capped at the occupancy-bounded resident block count (the loops are grid-stride).
  long long vb = 0;
  for (int i = 0; i < count; ++i) {
    const NicAdamTensor& t = b.t[i];
    const long long goff = t.g - xc.peer_flat[xc.rank];
    const bool vec = ((((uintptr_t)t.p | (uintptr_t)t.m | (uintptr_t)t.v) & 15) == 0) && (goff & 3) == 0;
    const long long items = vec ? t.numel / 4 + t.numel % 4 : t.numel;
    x.vec[i] = vec ? 1 : 0;
    x.vb_start[i] = (int)vb;
    vb += (items + 255) / 256;
  }
  x.vb_start[count] = (int)vb;
  vb += (xc.zero_numel / 4 + 255) / 256;
  x.vb_start[count + 1] = (int)vb;
  if (vb >= (1ll << 31)) return NIC_ERR_UNSUPPORTED;
  long long blocks = vb < 1 ? 1 : vb;
  long long cap = h->xch_resident_blocks;
  dim3 grid((unsigned)(blocks > cap ? cap : blocks));
  cudaError_t e = launch_pdl(adam_exchange_kernel, grid, dim3(256), 0, st, b, x);
  h->launches++;
  if (e == cudaSuccess) e = cudaGetLastError();
  return (int)e;
}

int launch_adam(Handle* h, const NicAdamTensor* tensors, int count, float beta1, float beta2, float eps,
                float grad_scale, int zero_grad, float* loss_sum, float* loss_out, float loss_scale, cudaStream_t st) {
  for (int base = 0; base < count; base += NIC_ADAM_BATCH) {
    AdamBatch b;
    b.loss_sum = base == 0 ? loss_sum : nullptr;
    b.loss_out = loss_out;
    b.loss_scale = loss_scale;
    b.metrics = h->step_metrics;
    b.count = count - base < NIC_ADAM_BATCH ? count - base : NIC_ADAM_BATCH;
    b.beta1 = beta1; b.beta2 = beta2; b.eps = eps; b.grad_scale = grad_scale; b.zero_grad = zero_grad;
    long long maxn = 0;
    for (int i = 0; i < b.count; ++i) {
      b.t[i] = tensors[base + i];
      const double bc1 = 1.0 - pow((double)beta1, (double)b.t[i].t), bc2 = 1.0 - pow((double)beta2, (double)b.t[i].t);
      b.step_size[i] = (float)((double)b.t[i].lr / bc1);
      b.bc2_sqrt[i] = (float)sqrt(bc2);
      if (b.t[i].numel > maxn) maxn = b.t[i].numel;
    }
    if (maxn == 0) continue;
    long long blocks = (maxn / 4 + 255) / 256 + 1;
    long long cap = (long long)h->sms * 8;
    dim3 grid((unsigned)(blocks > cap ? cap : blocks), (unsigned)b.count);
    cudaError_t e = launch_pdl(adam_kernel, grid, dim3(256), 0, st, b);
    h->launches++;
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  return NIC_OK;
}

// ------------------------------------------------------------------------------------------------ quantisers
enum { Q_QUANT = 0, Q_PACK = 1, Q_UNPACK = 2, Q_CLAMP = 3, Q_OUT8 = 4 };

template <int OP>
__global__ void __launch_bounds__(256) quant_kernel(const float* __restrict__ src, const uint8_t* __restrict__ csrc,
                                                    float* __restrict__ dst, uint8_t* __restrict__ cdst, long long n,
                                                    float scale, float offset, float lo, float hi) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (OP == Q_QUANT) {
      dst[i] = __fdiv_rn(quant_round(src[i], scale), scale);                 // models.py:55-57
    } else if (OP == Q_PACK) {
      float c = __fsub_rn(__fadd_rn(quant_round(src[i], scale), offset), 1.0f);   // models.py:61-64: + 2^(b-1) - 1
      cdst[i] = (uint8_t)fminf(fmaxf(c, 0.f), 255.f);
    } else if (OP == Q_UNPACK) {
      float z = __fadd_rn(__fsub_rn((float)csrc[i], offset), 1.0f);          // models.py:68-71: - 2^(b-1) + 1
      dst[i] = __fdiv_rn(z, scale);
    } else if (OP == Q_CLAMP) {
      dst[i] = fminf(fmaxf(dst[i], lo), hi);                                 // models.py:48-51
    } else {
      float q = quant_round(src[i], scale);                                  // models.py:29-40 (+ astype uint8)
      cdst[i] = (uint8_t)fminf(fmaxf(q, 0.f), scale);
    }
  }
}

template <int OP>
static int launch_q(Handle* h, const float* src, const uint8_t* csrc, float* dst, uint8_t* cdst, long long n,
                    float scale, float offset, float lo, float hi, cudaStream_t st) {
  if (n == 0) return NIC_OK;
  long long blocks = (n + 255) / 256, cap = (long long)h->sms * 16;
  quant_kernel<OP><<<(unsigned)(blocks > cap ? cap : blocks), 256, 0, st>>>(src, csrc, dst, cdst, n, scale, offset, lo, hi);
  h->launches++;
  return (int)cudaGetLastError();
}

int launch_quantize4fp(Handle* h, const float* src, float* dst, long long n, int bits, cudaStream_t st) {
  return launch_q<Q_QUANT>(h, src, nullptr, dst, nullptr, n, (float)((1 << bits) - 1), 0.f, 0.f, 0.f, st);
}
int launch_quantize_pack(Handle* h, const float* src, uint8_t* codes, long long n, int bits, cudaStream_t st) {
  return launch_q<Q_PACK>(h, src, nullptr, nullptr, codes, n, (float)((1 << bits) - 1), (float)(1 << (bits - 1)), 0.f, 0.f, st);
}
int launch_unpack(Handle* h, const uint8_t* codes, float* dst, long long n, int bits, cudaStream_t st) {
  return launch_q<Q_UNPACK>(h, nullptr, codes, dst, nullptr, n, (float)((1 << bits) - 1), (float)(1 << (bits - 1)), 0.f, 0.f, st);
}
int launch_clamp(Handle* h, float* p, long long n, float lo, float hi, cudaStream_t st) {
  return launch_q<Q_CLAMP>(h, nullptr, nullptr, p, nullptr, n, 0.f, 0.f, lo, hi, st);
}
int launch_output_to_u8(Handle* h, const float* src, uint8_t* dst, long long n, int bits, cudaStream_t st) {
  return launch_q<Q_OUT8>(h, src, nullptr, nullptr, dst, n, (float)((1 << bits) - 1), 0.f, 0.f, 0.f, st);
}

// True sub-byte storage of the grid codes (the reference spends one byte per code even for
// FP_BITS 4 / 2, models.py:61-64).  bits in {1, 2, 4, 8}: 8/bits codes per byte, code i in bits [(i % per) * bits, ...).
__global__ void __launch_bounds__(256) pack_bits_kernel(const uint8_t* __restrict__ codes, uint8_t* __restrict__ packed,
                                                        long long n, int bits) {
  const int per = 8 / bits;
  const long long nbytes = (n + per - 1) / per, stride = (long long)gridDim.x * blockDim.x;
  for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nbytes; b += stride) {
    unsigned v = 0;
    for (int k = 0; k < per; ++k) {
      const long long i = b * per + k;
      if (i < n) v |= (unsigned)(codes[i] & ((1u << bits) - 1u)) << (k * bits);
    }
    packed[b] = (uint8_t)v;
  }
}
__global__ void __launch_bounds__(256) unpack_bits_kernel(const uint8_t* __restrict__ packed, uint8_t* __restrict__ codes,
                                                          long long n, int bits) {
  const int per = 8 / bits;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    codes[i] = (uint8_t)((packed[i / per] >> ((int)(i % per) * bits)) & ((1u << bits) - 1u));
}

int launch_pack_bits(Handle* h, const uint8_t* codes, uint8_t* packed, long long n, int bits, int unpack, cudaStream_t st) {
  if (n == 0) return NIC_OK;
  long long work = unpack ? n : (n + 8 / bits - 1) / (8 / bits);
  long long blocks = (work + 255) / 256, cap = (long long)h->sms * 16;
  unsigned grid = (unsigned)(blocks > cap ? cap : blocks);
  if (unpack) unpack_bits_kernel<<<grid, 256, 0, st>>>(codes, packed, n, bits);      // (packed, codes) swapped by the caller
  else pack_bits_kernel<<<grid, 256, 0, st>>>(codes, packed, n, bits);
  h->launches++;
  return (int)cudaGetLastError();
}

// sum of squared differences of two 8-bit images (integer-exact per thread, double at the end)
__global__ void __launch_bounds__(256) sse_u8_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                     long long n, double* __restrict__ sse) {
  unsigned long long acc = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int d = (int)a[i] - (int)b[i];
    acc += (unsigned long long)(d * d);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  __shared__ unsigned long long red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(sse, (double)t);
  }
}

int launch_sse_u8(Handle* h, const uint8_t* a, const uint8_t* b, long long n, double* sse, cudaStream_t st) {
  if (n == 0) return NIC_OK;
  long long blocks = (n + 255) / 256, cap = (long long)h->sms * 8;
  sse_u8_kernel<<<(unsigned)(blocks > cap ? cap : blocks), 256, 0, st>>>(a, b, n, sse);
  h->launches++;
  return (int)cudaGetLastError();
}

}  // namespace nic
