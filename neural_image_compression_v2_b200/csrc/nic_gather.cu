// nic_gather.cu — K1, the tile-staged decoder-input kernels for the common 2-D shape (C = 12, PE = 6, step <= 1,
// runs of 128 texels along the fast block axis): X [N, 73] in fp32 (bit-exact with the reference), f16 or bf16
// (each value rounded once from the fp32 value).  Reference: create_decoder_input_2d / finally_decode_input_2d
// (Projects/image_compression.py:71-100, 170-181) over fp_def.create_g0_g1 (Projects/fp_def.py:115-145).
//
// The kernels are HBM-WRITE-bound: 73 * sizeof(X) bytes per texel leave the SM, ~3.7 bytes of grid come in
// (295.8 B/texel for fp32 X, 149.8 B/texel for 16-bit X; a pure bulk-store loop reaches 6.3 TB/s on a
// B200, tools/ubench/write_bw.cu).  Common design:
//   * a tile = 128 consecutive samples n (one ix, 128 consecutive iy) = ONE contiguous span of 128 * 73 elements of X;
//     it is assembled in shared memory and leaves with a single TMA bulk store (cp.async.bulk.global.shared::cta),
//     double-buffered so the store of tile i overlaps the build of tile i+1;
//   * the grid nodes a super-tile of several x-rows touches are staged once into shared memory in channel-LAST order, so
//     a corner is a few conflict-free vector loads instead of 12 strided global loads; the next super-tile's nodes are
//     prefetched into registers while the current rows are built;
//   * node loads carry an L2 evict_last policy, row stores evict_first: the grids stay L2-resident under the write stream;
//   * gather_tile_kernel (fp32 X; 16-bit X at steps outside {1/4, 1/2, 1}): thread = texel, run-time patch geometry, rows
//     written piece by piece through RowWriter (16-bit rows are 146 bytes: rows alternate 4-byte alignment, texel parity
//     is warp-uniform, lanes stride 73 words == 9 mod 32: bank-conflict free);
//   * gather_tile16_kernel (16-bit X): patch staged already rounded in two alignments, x-weighted G1 rows, compile-time
//     geometry — see its header below.
// Everything else (3-D, other C / PE, step > 1, short rows) takes the flat kernel in nic_f32.cu.
#include "nic_internal.cuh"

namespace nic {

constexpr int GT_T = 128;        // texels per tile
constexpr int GT_C = 12, GT_PE = 6, GT_CIN = 73;

__device__ __forceinline__ uint32_t gt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// L2 residency hints: the grids (63 MB at 4096^2) are re-read by neighbouring super-tiles while 2.4-4.9 GB of X stream
// through the same L2 — node loads ask to be evicted last, the row stores first (profiles/r02k: without the hints the
// kernel read 209 MB from DRAM for 66 MB of grids, and reads interleaved with a write stream cost more than their bytes).
__device__ __forceinline__ uint64_t gt_policy_keep() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t gt_policy_stream() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float gt_ldg_keep(const float* a, uint64_t pol) {
  float v;
  asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ void gt_bulk_store(void* dst, uint32_t src_smem, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src_smem),
               "r"(bytes), "l"(pol)
               : "memory");
}

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

// Streaming row writers: pieces of the row (12 channels of a corner, 6 encoding rows, the LOD) go to the staging
// buffer as soon as they are produced, so a thread never holds the whole 73-value row in registers.
template <typename OutT> struct RowWriter;
template <> struct RowWriter<float> {
  float* row;
  __device__ __forceinline__ RowWriter(float* stage, int t, int) : row(stage + t * GT_CIN) {}
  template <int N> __device__ __forceinline__ void put(const float* v, int base) {
#pragma unroll
    for (int i = 0; i < N; ++i) row[base + i] = v[i];
  }
  __device__ __forceinline__ void last(float v) { row[GT_CIN - 1] = v; }
};
// 16-bit rows are 146 bytes: the rows of a texel pair (2m, 2m+1) form 73 aligned words.  Even texels pair columns
// (even, even+1) and end with a lone low half (column 72); odd texels start with a lone high half (column 0) and pair
// (even-1, even), which needs a one-value carry between pieces.  Parity is warp-uniform.
template <typename OutT> struct RowWriter16 {
  using P = Pack16<OutT>;
  uint32_t* words;
  int par;
  float carry;
  __device__ __forceinline__ RowWriter16(OutT* stage, int t, int par_)
      : words(reinterpret_cast<uint32_t*>(stage) + (t >> 1) * GT_CIN), par(par_), carry(0.f) {}
  template <int N> __device__ __forceinline__ void put(const float* v, int base) {   // N and base even
    if (par == 0) {
#pragma unroll
      for (int i = 0; i < N / 2; ++i) words[base / 2 + i] = P::two(v[2 * i], v[2 * i + 1]);
    } else {
      if (base == 0) reinterpret_cast<uint16_t*>(words + 36)[1] = P::one(v[0]);
      else words[36 + base / 2] = P::two(carry, v[0]);
#pragma unroll
      for (int i = 1; i < N / 2; ++i) words[36 + base / 2 + i] = P::two(v[2 * i - 1], v[2 * i]);
      carry = v[N - 1];
    }
  }
  __device__ __forceinline__ void last(float v) {
    if (par == 0) reinterpret_cast<uint16_t*>(words + 36)[0] = P::one(v);
    else words[GT_CIN - 1] = P::two(carry, v);
  }
};
template <> struct RowWriter<__half> : RowWriter16<__half> { using RowWriter16<__half>::RowWriter16; };
template <> struct RowWriter<__nv_bfloat16> : RowWriter16<__nv_bfloat16> { using RowWriter16<__nv_bfloat16>::RowWriter16; };

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

// One CTA iteration = a SUPER-TILE of GT_R x-rows by 128 y-texels: the GT_R rows share one staged patch of grid nodes
// (NX0 x NP0 nodes of G0, NX1 x NP1 of G1), which divides the global-load / L1-lookup cost per texel by GT_R; each
// row of 128 texels is then built in a double-buffered staging area and leaves with its own bulk store.
constexpr int GT_R = 4;

template <typename OutT>
__global__ void __launch_bounds__(GT_T) gather_tile_kernel(DevGeom g, const float* __restrict__ g0,
                                                           const float* __restrict__ g1,
                                                           const long long* __restrict__ origins, OutT* __restrict__ x,
                                                           int nx0, int nx1, int np0, int np1, unsigned ntiles,
                                                           unsigned tiles_per_row, int /*dbg*/) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int STAGE_BYTES = GT_T * GT_CIN * (int)sizeof(OutT);
  float* patch0 = reinterpret_cast<float*>(smem_raw + 2 * STAGE_BYTES);      // [nx0][np0][C]
  float* patch1 = patch0 + nx0 * np0 * GT_C;                                // [nx1][np1][C]
  float* pex = patch1 + nx1 * np1 * GT_C;                                   // [GT_R][8]: x encodings of the rows

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // 16-bit rows: warp-uniform texel parity (see header); fp32 rows: thread = texel, lanes stride 73 words (== 9 mod 32)
  const int par = sizeof(OutT) == 2 ? (warp & 1) : 0;
  const int t = sizeof(OutT) == 2 ? 64 * (warp >> 1) + 2 * lane + par : tid;
  const uint64_t pol_keep = gt_policy_keep(), pol_stream = gt_policy_stream();
  const long long ps0 = (long long)g.n0[0] * g.n0[1], ps1 = (long long)g.n1[0] * g.n1[1];
  const unsigned xgroups = (unsigned)g.B[0] / GT_R;
  const unsigned tiles_per_block = xgroups * tiles_per_row;

  // Patch staging: thread -> one (channel, x-node) column of the patch, walking the y nodes with a fixed stride (no
  // runtime division).  Lanes run over x-nodes first (adjacent addresses), then channels.  The first PRE loads per grid
  // are PREFETCHED into registers one super-tile ahead, hiding their latency behind the previous rows' build.
  constexpr int PRE0 = 12, PRE1 = 6;
  const int ncol0 = GT_C * nx0, ncol1 = GT_C * nx1;            // <= 128 (launcher guarantees)
  const int ys0 = GT_T / ncol0, ys1 = GT_T / ncol1;            // y stride of a staging thread
  const int col0 = tid % ncol0, yA0 = tid / ncol0, c0_ = col0 / nx0, xn0_ = col0 - c0_ * nx0;
  const int col1 = tid % ncol1, yA1 = tid / ncol1, c1_ = col1 / nx1, xn1_ = col1 - c1_ * nx1;
  const bool stager0 = yA0 < ys0, stager1 = yA1 < ys1;
  float pre0[PRE0], pre1[PRE1];
  auto tile_coords = [&](unsigned tile, int& px0, int& py0) {
    const unsigned b = tile / tiles_per_block, r = tile - b * tiles_per_block;
    const unsigned xg = r / tiles_per_row, iy0 = (r - xg * tiles_per_row) * GT_T;
    int ox, oy;
    if (origins) {
      ox = (int)origins[2 * (long long)b];
      oy = (int)origins[2 * (long long)b + 1];
    } else {
      ox = g.origin0[0];
      oy = g.origin0[1];
    }
    px0 = ox + (int)xg * GT_R;
    py0 = oy + (int)iy0;
  };
  auto load0 = [&](int yn, const AxisCoord& axf, const AxisCoord& ayf) -> float {
    const int gx = clampi(axf.i0 + xn0_, 0, g.n0[0] - 1), gy = clampi(ayf.i0 + yn, 0, g.n0[1] - 1);
    return gt_ldg_keep(g0 + c0_ * ps0 + (long long)gy * g.n0[0] + gx, pol_keep);
  };
  auto load1 = [&](int yn, const AxisCoord& axf, const AxisCoord& ayf) -> float {
    const int gx = clampi(axf.i1 + xn1_, 0, g.n1[0] - 1), gy = clampi(ayf.i1 + yn, 0, g.n1[1] - 1);
    return gt_ldg_keep(g1 + c1_ * ps1 + (long long)gy * g.n1[0] + gx, pol_keep);
  };
  auto prefetch = [&](unsigned tile) {
    if (tile >= ntiles) return;
    int px0, py0;
    tile_coords(tile, px0, py0);
    const AxisCoord axf = axis_coord(px0, g.step), ayf = axis_coord(py0, g.step);
    if (stager0) {
#pragma unroll
      for (int k = 0; k < PRE0; ++k)
        if (yA0 + k * ys0 < np0) pre0[k] = load0(yA0 + k * ys0, axf, ayf);
    }
    if (stager1) {
#pragma unroll
      for (int k = 0; k < PRE1; ++k)
        if (yA1 + k * ys1 < np1) pre1[k] = load1(yA1 + k * ys1, axf, ayf);
    }
  };

  prefetch(blockIdx.x);
  unsigned nstore = 0;                   // bulk stores issued so far by this CTA (staging buffer = nstore & 1)
  for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int px0, py0;
    tile_coords(tile, px0, py0);
    const AxisCoord ax_first = axis_coord(px0, g.step), ay_first = axis_coord(py0, g.step);
    __syncthreads();                     // every thread is done with the previous super-tile's patches
    // ---- stage the grid nodes of this super-tile, channel-last: patch[(xn * np + yn) * C + c]
    if (stager0) {
      float* colp = patch0 + xn0_ * np0 * GT_C + c0_;
#pragma unroll
      for (int k = 0; k < PRE0; ++k)
        if (yA0 + k * ys0 < np0) colp[(yA0 + k * ys0) * GT_C] = pre0[k];
      for (int yb = yA0 + PRE0 * ys0; yb < np0; yb += 8 * ys0) {        // larger patches (step > 1/4): unrolled chunks
        float tmp[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (yb + k * ys0 < np0) tmp[k] = load0(yb + k * ys0, ax_first, ay_first);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (yb + k * ys0 < np0) colp[(yb + k * ys0) * GT_C] = tmp[k];
      }
    }
    if (stager1) {
      float* colp = patch1 + xn1_ * np1 * GT_C + c1_;
#pragma unroll
      for (int k = 0; k < PRE1; ++k)
        if (yA1 + k * ys1 < np1) colp[(yA1 + k * ys1) * GT_C] = pre1[k];
      for (int yb = yA1 + PRE1 * ys1; yb < np1; yb += 8 * ys1) {
        float tmp[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (yb + k * ys1 < np1) tmp[k] = load1(yb + k * ys1, ax_first, ay_first);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (yb + k * ys1 < np1) colp[(yb + k * ys1) * GT_C] = tmp[k];
      }
    }
    if (tid >= GT_T - GT_R * GT_PE) {      // the last 24 threads: x encodings of the GT_R rows, shared by their texels
      const int e = tid - (GT_T - GT_R * GT_PE), row = e / GT_PE, rr = e - row * GT_PE;
      pex[row * 8 + rr] = pe_value(g, axis_coord(px0 + row, g.step).u1, rr);
    }
    __syncthreads();
    prefetch(tile + gridDim.x);        // next super-tile's nodes: in flight while this one's rows are built
    const AxisCoord ay = axis_coord(py0 + t, g.step);
    const int y0 = ay.i0 - ay_first.i0, y1 = ay.i1 - ay_first.i1;
    const float ky = ay.k, wy0 = __fsub_rn(1.0f, ky);
    float pey[GT_PE];
    if (g.pe_kind == NIC_PE_TRIANGULAR) {
      pe6_triangular(ay.u1, pey);
    } else {
#pragma unroll
      for (int rr = 0; rr < GT_PE; ++rr) pey[rr] = pe_sinusoidal(ay.u1, rr, g.pe_div);
    }
    const unsigned b = tile / tiles_per_block, r = tile - b * tiles_per_block;
    const unsigned xg = r / tiles_per_row, run = r - xg * tiles_per_row;
#pragma unroll 1
    for (int row = 0; row < GT_R; ++row, ++nstore) {
      OutT* stage = reinterpret_cast<OutT*>(smem_raw + (nstore & 1) * STAGE_BYTES);
      // the bulk store that used this staging buffer two rows ago must have finished READING it
      if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncthreads();
      // ---- this thread's row, in reference column order (image_compression.py:94-96), streamed piece by piece
      {
        const AxisCoord ax = axis_coord(px0 + row, g.step);
        const int x0 = ax.i0 - ax_first.i0, x1 = ax.i1 - ax_first.i1;
        RowWriter<OutT> w(stage, t, par);
        // G0 corners (dy,dx) = (0,0) (1,0) (0,1) (1,1): raw copies
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int dy = j & 1, dx = j >> 1;
          const float4* node = reinterpret_cast<const float4*>(patch0 + ((x0 + dx) * np0 + y0 + dy) * GT_C);
          const float4 f0 = node[0], f1 = node[1], f2 = node[2];
          const float v[12] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x, f2.y, f2.z, f2.w};
          w.template put<12>(v, 12 * j);
        }
        // G1: ((g*wx)*wy) per corner, summed ((c0+c1)+c2)+c3 (fp_def.py:136-144, image_compression.py:95)
        {
          const float kx = ax.k, wx0 = __fsub_rn(1.0f, kx);
          float acc[12];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int dy = j & 1, dx = j >> 1;
            const float wx = dx ? kx : wx0, wy = dy ? ky : wy0;
            const float4* node = reinterpret_cast<const float4*>(patch1 + ((x1 + dx) * np1 + y1 + dy) * GT_C);
            const float4 f0 = node[0], f1 = node[1], f2 = node[2];
            float e[12] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x, f2.y, f2.z, f2.w};
#pragma unroll
            for (int c = 0; c < 12; ++c) {
              if (g.interp) e[c] = __fmul_rn(__fmul_rn(e[c], wx), wy);
              acc[c] = j == 0 ? e[c] : __fadd_rn(acc[c], e[c]);
            }
          }
          w.template put<12>(acc, 48);
        }
        // positional encodings of u1 along x (shared by the row) then y, then the LOD column
        {
          float pe[GT_PE];
#pragma unroll
          for (int rr = 0; rr < GT_PE; ++rr) pe[rr] = pex[row * 8 + rr];
          w.template put<GT_PE>(pe, 60);
          w.template put<GT_PE>(pey, 66);
        }
        w.last(g.lod);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the bulk copy
      __syncthreads();
      if (tid == 0) {
        // sample index of this row's first texel: block b, x row xg*GT_R + row, y run `run`
        const size_t n0 = (size_t)b * (size_t)g.per_block + ((size_t)xg * GT_R + row) * (size_t)g.B[1] + (size_t)run * GT_T;
        OutT* dst = x + n0 * GT_CIN;
        gt_bulk_store(dst, gt_smem_u32(stage), STAGE_BYTES, pol_stream);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------- 16-bit X (f16 / bf16)
// The 16-bit rows cost the same HBM time per byte as the fp32 rows but half the bytes per texel, so the row build of the
// kernel above showed: 464 instructions per texel (every value converted and paired per texel, 96 multiplies of the G1
// interpolation, ~180 of index arithmetic for staging 4 rows' worth of nodes), 73 % of the copy peak against 83 % for
// fp32 X (profiles/r01h).  Knock-outs on this kernel's first version (profiles/r02k): stores alone 0.44 ms, build alone
// 0.38 ms, together 0.52 ms — the build has to shrink for the two to overlap.  This kernel removes the per-texel work
// that does not depend on the texel:
//   * the G0 patch is staged ALREADY ROUNDED to 16 bits, channel-last, twice: `pa` aligned and `ps` shifted by one
//     half (ps[h] = pa[h - 1]).  The corners (x, y) and (x, y + 1) are neighbours in the patch, so the 48 G0 columns of
//     a texel are two runs of 24 halves: an even texel copies 2 x 12 aligned words of `pa`, an odd texel (whose row
//     starts in the upper half of a word) copies the same runs from `ps`, already paired (col 2i-1, col 2i) — no
//     conversions, no carries, two byte-permutes per row for the words that straddle two runs;
//   * the x-weights of the G1 interpolation are applied once per (row, node) — gx[dx][yn][c] = g * wx, the first
//     product of the reference's ((g * wx) * wy) — so a texel does 48 multiplies and 36 adds instead of 96 and 36, with
//     the same roundings in the same order (bit-identical to the flat kernel and to the fp32 row rounded once);
//   * the x encodings of a row are staged as packed words for both parities, the y encodings and the LOD are packed
//     per thread once per super-tile;
//   * a super-tile is R = 8 x-rows (4 for steps above 1/2, whose patches would not leave room for two CTAs per SM), and
//     the staging threads walk their node column with one min, one multiply-add and one load per node.
// Patch geometry of a super-tile of R x-rows by 128 y-texels at step 2^-SL (SL = 2, 1, 0: the steps a mip chain produces
// below 2), compile-time so that every shared-memory offset is an immediate and the staging loops unroll completely —
// with run-time patch sizes the compiler re-derived the offsets inside the row loop (90 uniform-datapath instructions
// per row, profiles/r02k).
template <int R, int SL> struct G16 {
  static constexpr int NP0 = (127 >> SL) + 3, NP1 = (127 >> (SL + 1)) + 3;       // nodes a run of 128 texels can touch (+1: unaligned start)
  static constexpr int NX0 = ((R - 1) >> SL) + 3, NX1 = ((R - 1) >> (SL + 1)) + 3;
  static constexpr int NCOL0 = GT_C * NX0, NCOL1 = GT_C * NX1;                   // staging columns (channel, x-node): <= 128
  static constexpr int YS0 = GT_T / NCOL0, YS1 = GT_T / NCOL1;                   // y stride of a staging thread
  static constexpr int NK0 = (NP0 + YS0 - 1) / YS0, NK1 = (NP1 + YS1 - 1) / YS1; // nodes per staging thread (upper bound)
  static constexpr int STAGE_BYTES = GT_T * GT_CIN * 2;
  static constexpr int N0B = (NX0 * NP0 * GT_C * 2 + 24 + 15) & ~15;             // one copy of the 16-bit G0 patch (+ a node of padding)
  static constexpr int HALF1 = NP1 * GT_C;                                       // one x-node column of patch1 / gx
  static constexpr int OFF_PA = 2 * STAGE_BYTES, OFF_PS = OFF_PA + N0B, OFF_P1 = OFF_PS + N0B;
  static constexpr int OFF_GX = OFF_P1 + NX1 * HALF1 * 4, OFF_PEX = OFF_GX + 2 * HALF1 * 4, SMEM = OFF_PEX + R * 8 * 4;
  static_assert(NCOL0 <= GT_T && NCOL1 <= GT_T, "one staging thread per (channel, x-node) column");
};

template <typename OutT, int R, int SL>
__global__ void __launch_bounds__(GT_T, 4) gather_tile16_kernel(DevGeom g, const float* __restrict__ g0,
                                                                const float* __restrict__ g1,
                                                                const long long* __restrict__ origins, OutT* __restrict__ x,
                                                                int, int, int, int, unsigned ntiles, unsigned tiles_per_row,
                                                                int dbg) {
  using P = Pack16<OutT>;
  using G = G16<R, SL>;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int STAGE_BYTES = G::STAGE_BYTES;
  constexpr int np0 = G::NP0, np1 = G::NP1, nx0 = G::NX0, nx1 = G::NX1, half1 = G::HALF1;
  uint16_t* const pa = reinterpret_cast<uint16_t*>(smem_raw + G::OFF_PA);    // [nx0][np0][C] halves
  uint16_t* const ps = reinterpret_cast<uint16_t*>(smem_raw + G::OFF_PS);    // ps[h] = pa[h - 1]
  float* const patch1 = reinterpret_cast<float*>(smem_raw + G::OFF_P1);      // [nx1][np1][C] fp32
  float* const gx = reinterpret_cast<float*>(smem_raw + G::OFF_GX);          // [2][np1][C]: this row's x-weighted G1 nodes
  uint32_t* const pexw = reinterpret_cast<uint32_t*>(smem_raw + G::OFF_PEX); // [R][8] packed x encodings

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int par = warp & 1;                                  // warp-uniform texel parity
  const int t = 64 * (warp >> 1) + 2 * lane + par;
  const uint64_t pol_keep = gt_policy_keep(), pol_stream = gt_policy_stream();
  const unsigned xgroups = (unsigned)g.B[0] / R;
  const unsigned tiles_per_block = xgroups * tiles_per_row;
  constexpr float step = 1.0f / (float)(1 << SL);

  // Patch staging: thread -> one (channel, x-node) column of a patch, walking the y nodes with a fixed stride, all of
  // them prefetched into registers one super-tile ahead.  Texel coordinates are never negative, so only the upper clamp
  // of the node index is needed (nodes past the edge are staged but never used).
  constexpr int PRE0 = G::NK0 < 17 ? G::NK0 : 17, PRE1 = G::NK1 < 9 ? G::NK1 : 9;
  constexpr int ncol0 = G::NCOL0, ncol1 = G::NCOL1, ys0 = G::YS0, ys1 = G::YS1;
  const int col0 = tid % ncol0, yA0 = tid / ncol0, c0_ = col0 / nx0, xn0_ = col0 - c0_ * nx0;
  const int col1 = tid % ncol1, yA1 = tid / ncol1, c1_ = col1 / nx1, xn1_ = col1 - c1_ * nx1;
  const bool stager0 = yA0 < ys0, stager1 = yA1 < ys1;
  const float* const gcol0 = g0 + (long long)c0_ * g.n0[0] * g.n0[1];        // this thread's channel plane
  const float* const gcol1 = g1 + (long long)c1_ * g.n1[0] * g.n1[1];
  const int n0x = g.n0[0], n0y1 = g.n0[1] - 1, n1x = g.n1[0], n1y1 = g.n1[1] - 1;
  constexpr int e0step = ys0 * GT_C, e1step = ys1 * GT_C;
  uint16_t* const pa_t = pa + (xn0_ * np0 + yA0) * GT_C + c0_;               // this thread's node 0 in both copies: node k at + k * e0step
  float* const p1col = patch1 + (xn1_ * np1 + yA1) * GT_C + c1_;
  float pre0[PRE0], pre1[PRE1];
  auto tile_coords = [&](unsigned tile, int& px0, int& py0) {
    const unsigned b = tile / tiles_per_block, r = tile - b * tiles_per_block;
    const unsigned xg = r / tiles_per_row, iy0 = (r - xg * tiles_per_row) * GT_T;
    int ox, oy;
    if (origins) {
      ox = (int)origins[2 * (long long)b];
      oy = (int)origins[2 * (long long)b + 1];
    } else {
      ox = g.origin0[0];
      oy = g.origin0[1];
    }
    px0 = ox + (int)xg * R;
    py0 = oy + (int)iy0;
  };
  auto loads0 = [&](int gx0, int gy0, int k0, int k1, float* dst) {       // nodes k0 .. k1-1 of this thread's G0 column
    const float* colp = gcol0 + min(gx0 + xn0_, n0x - 1);
#pragma unroll
    for (int k = 0; k < PRE0; ++k)
      if (k0 + k < k1) dst[k] = gt_ldg_keep(colp + min(gy0 + yA0 + (k0 + k) * ys0, n0y1) * n0x, pol_keep);
  };
  auto loads1 = [&](int gx1, int gy1, int k0, int k1, float* dst) {
    const float* colp = gcol1 + min(gx1 + xn1_, n1x - 1);
#pragma unroll
    for (int k = 0; k < PRE1; ++k)
      if (k0 + k < k1) dst[k] = gt_ldg_keep(colp + min(gy1 + yA1 + (k0 + k) * ys1, n1y1) * n1x, pol_keep);
  };
  const int nk0 = stager0 ? (np0 - yA0 + ys0 - 1) / ys0 : 0, nk1 = stager1 ? (np1 - yA1 + ys1 - 1) / ys1 : 0;   // nodes per thread
  auto prefetch = [&](unsigned tile) {
    if (tile >= ntiles) return;
    int px0, py0;
    tile_coords(tile, px0, py0);
    loads0(px0 >> SL, py0 >> SL, 0, nk0, pre0);
    loads1(px0 >> (SL + 1), py0 >> (SL + 1), 0, nk1, pre1);
  };
  auto puts0 = [&](int k0, int k1, const float* v) {           // G0 values -> both copies of the 16-bit patch
#pragma unroll
    for (int k = 0; k < PRE0; ++k)
      if (k0 + k < k1) {
        const uint16_t h = P::one(v[k]);
        pa_t[(k0 + k) * e0step] = h;
        pa_t[(k0 + k) * e0step + G::N0B / 2 + 1] = h;          // ps[e + 1]
      }
  };
  auto puts1 = [&](int k0, int k1, const float* v) {
#pragma unroll
    for (int k = 0; k < PRE1; ++k)
      if (k0 + k < k1) p1col[(k0 + k) * e1step] = v[k];
  };

  prefetch(blockIdx.x);
  unsigned nstore = 0;                   // bulk stores issued so far by this CTA (staging buffer = nstore & 1)
  for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int px0, py0;
    tile_coords(tile, px0, py0);
    // p * step and p * step / 2 are exact (dyadic): integer shifts give the node indices, the masked low bits the weights
    constexpr int M1 = (2 << SL) - 1;
    constexpr float HS = 0.5f * step;
    __syncthreads();                     // every thread is done with the previous super-tile's patches
    puts0(0, nk0, pre0);
    puts1(0, nk1, pre1);
    for (int k0 = PRE0; k0 < nk0; k0 += PRE0) {       // larger patches: the rest of the column, loads before stores
      float tmp[PRE0];
      loads0(px0 >> SL, py0 >> SL, k0, nk0, tmp);
      puts0(k0, nk0, tmp);
    }
    for (int k0 = PRE1; k0 < nk1; k0 += PRE1) {
      float tmp[PRE1];
      loads1(px0 >> (SL + 1), py0 >> (SL + 1), k0, nk1, tmp);
      puts1(k0, nk1, tmp);
    }
    if (tid >= GT_T - R) {               // the last R threads: packed x encodings of one row each, both pairings
      const int row = tid - (GT_T - R);
      const float u = (float)(px0 + row) * HS;
      float pe[GT_PE];
#pragma unroll
      for (int rr = 0; rr < GT_PE; ++rr) pe[rr] = pe_value(g, u, rr);
      uint32_t* w = pexw + row * 8;
      w[0] = P::two(pe[0], pe[1]);
      w[1] = P::two(pe[2], pe[3]);
      w[2] = P::two(pe[4], pe[5]);
      w[4] = (uint32_t)P::one(pe[0]) << 16;
      w[5] = P::two(pe[1], pe[2]);
      w[6] = P::two(pe[3], pe[4]);
      w[7] = (uint32_t)P::one(pe[5]);
    }
    __syncthreads();
    prefetch(tile + gridDim.x);        // next super-tile's nodes: in flight while this one's rows are built
    const int py = py0 + t;
    const float ky = (float)(py & M1) * HS, wy0 = __fsub_rn(1.0f, ky), uy1 = (float)py * HS;
    // this texel's first G0 run (words) at x-node 0, its G1 node pair, its two words of the staging row
    constexpr int xrun = (np0 * GT_C) >> 1;
    const int run0 = (((py >> SL) - (py0 >> SL)) * GT_C) >> 1;
    const float* const gxt = gx + ((py >> (SL + 1)) - (py0 >> (SL + 1))) * GT_C;
    const uint32_t* const src = reinterpret_cast<const uint32_t*>(par == 0 ? pa : ps) + run0;
    const int wrow = (t >> 1) * GT_CIN + par * 36;
    // y encodings + LOD of this texel as packed words of its parity (constant over the R rows)
    uint32_t yw0, yw1, yw2, yw3;
    {
      float pey[GT_PE];
      if (g.pe_kind == NIC_PE_TRIANGULAR) {
        pe6_triangular(uy1, pey);
      } else {
#pragma unroll
        for (int rr = 0; rr < GT_PE; ++rr) pey[rr] = pe_sinusoidal(uy1, rr, g.pe_div);
      }
      if (par == 0) {
        yw0 = P::two(pey[0], pey[1]);
        yw1 = P::two(pey[2], pey[3]);
        yw2 = P::two(pey[4], pey[5]);
        yw3 = P::one(g.lod);
      } else {
        yw0 = (uint32_t)P::one(pey[0]) << 16;
        yw1 = P::two(pey[1], pey[2]);
        yw2 = P::two(pey[3], pey[4]);
        yw3 = P::two(pey[5], g.lod);
      }
    }
    const unsigned b = tile / tiles_per_block, r = tile - b * tiles_per_block;
    const unsigned xg = r / tiles_per_row, run = r - xg * tiles_per_row;
    const size_t n_first = (size_t)b * (size_t)g.per_block + (size_t)xg * R * (size_t)g.B[1] + (size_t)run * GT_T;
    // x-weighted G1 nodes: gx has 2 * half1 elements, thread -> elements tid, tid + 128, ... (at most GXK each)
    constexpr int GXK = 4;
#pragma unroll 1
    for (int row = 0; row < R; ++row, ++nstore) {
      uint32_t* stage = reinterpret_cast<uint32_t*>(smem_raw + (nstore & 1) * STAGE_BYTES);
      const int px = px0 + row;
      const int x0 = (px >> SL) - (px0 >> SL);
      {
        // gx[dx][yn][c] = G1[x1 + dx][yn][c] * wx(dx): everybody finished reading the previous row's gx at the barrier
        // that closed its build.  Loads first, then stores (see below).
        const float kx = (float)(px & M1) * HS, wx0 = __fsub_rn(1.0f, kx);
        const float* p1 = patch1 + ((px >> (SL + 1)) - (px0 >> (SL + 1))) * half1;            // columns x1 and x1 + 1 are adjacent: one run
        for (int e0 = tid; e0 < 2 * half1; e0 += GXK * GT_T) {
          float v[GXK];
#pragma unroll
          for (int k = 0; k < GXK; ++k) {
            const int e = e0 + k * GT_T;
            if (e < 2 * half1) v[k] = __fmul_rn(p1[e], e >= half1 ? kx : wx0);
          }
#pragma unroll
          for (int k = 0; k < GXK; ++k)
            if (e0 + k * GT_T < 2 * half1) gx[e0 + k * GT_T] = v[k];
        }
      }
      // the bulk store that used this staging buffer two rows ago must have finished READING it
      if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncthreads();
      if (!(dbg & 1)) {
        // Every shared-memory LOAD of the row comes before its first STORE: patches and staging buffer are one array to
        // the compiler, so a load written after a store would wait for it.
        // G1 sum ((A + B) + C) + D with A = gx0[y1] * wy0, B = gx0[y1 + 1] * ky, C = gx1[y1] * wy0, D = gx1[y1 + 1] * ky
        float acc[12];
        {
          const float4* n0p = reinterpret_cast<const float4*>(gxt);
          const float4* n1p = reinterpret_cast<const float4*>(gxt + half1);
          float4 a[3], bq[3], c[3], d[3];
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            a[q] = n0p[q];
            bq[q] = n0p[3 + q];
            c[q] = n1p[q];
            d[q] = n1p[3 + q];
          }
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            acc[4 * q + 0] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a[q].x, wy0), __fmul_rn(bq[q].x, ky)), __fmul_rn(c[q].x, wy0)), __fmul_rn(d[q].x, ky));
            acc[4 * q + 1] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a[q].y, wy0), __fmul_rn(bq[q].y, ky)), __fmul_rn(c[q].y, wy0)), __fmul_rn(d[q].y, ky));
            acc[4 * q + 2] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a[q].z, wy0), __fmul_rn(bq[q].z, ky)), __fmul_rn(c[q].z, wy0)), __fmul_rn(d[q].z, ky));
            acc[4 * q + 3] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a[q].w, wy0), __fmul_rn(bq[q].w, ky)), __fmul_rn(c[q].w, wy0)), __fmul_rn(d[q].w, ky));
          }
        }
        const uint32_t* xw = pexw + row * 8;
        const uint32_t* s1 = src + x0 * xrun;                  // run (x0, y0), (x0, y0 + 1); the run of x0 + 1 follows at + xrun
        const uint32_t* s2 = s1 + xrun;
        uint2 q1[6], q2[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          q1[i] = reinterpret_cast<const uint2*>(s1)[i];
          q2[i] = reinterpret_cast<const uint2*>(s2)[i];
        }
        uint32_t* w = stage + wrow;
        if (par == 0) {
          const uint32_t x0w = xw[0], x1w = xw[1], x2w = xw[2];
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            w[2 * i] = q1[i].x;
            w[2 * i + 1] = q1[i].y;
            w[12 + 2 * i] = q2[i].x;
            w[12 + 2 * i + 1] = q2[i].y;
          }
#pragma unroll
          for (int i = 0; i < 6; ++i) w[24 + i] = P::two(acc[2 * i], acc[2 * i + 1]);
          w[30] = x0w;
          w[31] = x1w;
          w[32] = x2w;
          w[33] = yw0;
          w[34] = yw1;
          w[35] = yw2;
          reinterpret_cast<uint16_t*>(w + 36)[0] = (uint16_t)yw3;
        } else {
          // w[0].hi = column 0, w[i] = columns (2i-1, 2i); the shifted runs are already paired that way
          const uint32_t c1 = s1[12], c2 = s2[12];           // low halves: the last value of run 1 / run 2
          const uint32_t x4w = xw[4], x5w = xw[5], x6w = xw[6], x7w = xw[7];
          reinterpret_cast<uint16_t*>(w)[1] = (uint16_t)(q1[0].x >> 16);
          w[1] = q1[0].y;
#pragma unroll
          for (int i = 1; i < 6; ++i) {
            w[2 * i] = q1[i].x;
            w[2 * i + 1] = q1[i].y;
          }
          w[12] = __byte_perm(c1, q2[0].x, 0x7610);
          w[13] = q2[0].y;
#pragma unroll
          for (int i = 1; i < 6; ++i) {
            w[12 + 2 * i] = q2[i].x;
            w[12 + 2 * i + 1] = q2[i].y;
          }
          w[24] = (c2 & 0xffffu) | ((uint32_t)P::one(acc[0]) << 16);
#pragma unroll
          for (int i = 1; i < 6; ++i) w[24 + i] = P::two(acc[2 * i - 1], acc[2 * i]);
          w[30] = (uint32_t)P::one(acc[11]) | x4w;
          w[31] = x5w;
          w[32] = x6w;
          w[33] = x7w | yw0;
          w[34] = yw1;
          w[35] = yw2;
          w[36] = yw3;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the bulk copy
      __syncthreads();
      if (tid == 0) {
        OutT* dst = x + (n_first + (size_t)row * (size_t)g.B[1]) * GT_CIN;
        if (!(dbg & 2)) gt_bulk_store(dst, gt_smem_u32(stage), STAGE_BYTES, pol_stream);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

bool gather_tile_eligible(const DevGeom& g, const void* x) {
  // x-nodes a group of GT_R rows can touch: floor((GT_R-1)*step) + 3 (unaligned start); 12 * nx staging columns <= 128
  const int nx0 = (int)floorf((GT_R - 1) * g.step) + 3;
  return g.method == NIC_METHOD_2D && g.C == GT_C && g.PE == GT_PE && g.step <= 1.0f && g.step >= 1.0f / 1024.0f &&
         GT_C * nx0 <= GT_T && g.B[1] > 0 && g.B[1] % GT_T == 0 && g.B[0] % GT_R == 0 && g.N > 0 && g.N < (1ll << 31) &&
         (((uintptr_t)x) & 15) == 0;
}

template <typename OutT>
static int launch_gather_tile_t(Handle* h, const DevGeom& g, const float* g0, const float* g1, const long long* origins,
                                OutT* x, cudaStream_t st) {
  // nodes a run of T texels can touch along y: floor((T-1)*step) + 2 cells' corners, +1 for an unaligned start
  const int np0 = (int)floorf((GT_T - 1) * g.step) + 3, np1 = (int)floorf((GT_T - 1) * g.step * 0.5f) + 3;
  int nx0 = (int)floorf((GT_R - 1) * g.step) + 3, nx1 = (int)floorf((GT_R - 1) * g.step * 0.5f) + 3;
  // 16-bit rows with interpolated G1 (always, for step <= 1) take the kernel that stages the G0 patch already rounded
  constexpr bool ROWS16 = sizeof(OutT) == 2;
  size_t smem = 2 * (size_t)GT_T * GT_CIN * sizeof(OutT) + ((size_t)(nx0 * np0 + nx1 * np1) * GT_C + 8 * GT_R) * sizeof(float);
  // 16-bit rows at the steps of a mip chain (1/4, 1/2, 1): compile-time patch geometry, super-tiles of 8 x-rows up to
  // step 1/2 (4 at step 1, whose patches would leave room for one CTA per SM, or when the block height is not a
  // multiple of 8); other steps keep the run-time kernel above
  const int sl = g.step == 0.25f ? 2 : (g.step == 0.5f ? 1 : (g.step == 1.0f ? 0 : -1));
  const bool k16 = ROWS16 && g.interp && sl >= 0 && !(h->debug_flags & 64);
  const int R = (k16 && sl >= 1 && g.B[0] % 8 == 0 && !(h->debug_flags & 16)) ? 8 : GT_R;
  void (*kern)(DevGeom, const float*, const float*, const long long*, OutT*, int, int, int, int, unsigned, unsigned, int) =
      gather_tile_kernel<OutT>;
  if constexpr (ROWS16) {
    if (k16) {
      if (R == 8 && sl == 2) kern = gather_tile16_kernel<OutT, 8, 2>, smem = G16<8, 2>::SMEM;
      else if (R == 8) kern = gather_tile16_kernel<OutT, 8, 1>, smem = G16<8, 1>::SMEM;
      else if (sl == 2) kern = gather_tile16_kernel<OutT, GT_R, 2>, smem = G16<GT_R, 2>::SMEM;
      else if (sl == 1) kern = gather_tile16_kernel<OutT, GT_R, 1>, smem = G16<GT_R, 1>::SMEM;
      else kern = gather_tile16_kernel<OutT, GT_R, 0>, smem = G16<GT_R, 0>::SMEM;
    }
  }
  if (smem > 200 * 1024) return NIC_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const unsigned tiles_per_row = (unsigned)(g.B[1] / GT_T);
  const unsigned ntiles = (unsigned)(g.N / (GT_T * R));
  int per_sm = 8;
  {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, GT_T, smem) == cudaSuccess && occ > 0) per_sm = occ;
  }
  long long cap = (long long)h->sms * per_sm;
  int grid = (int)(ntiles < cap ? ntiles : cap);
  {
    KernelTimer timer(h, st);
    kern<<<grid, GT_T, smem, st>>>(g, g0, g1, origins, x, nx0, nx1, np0, np1, ntiles, tiles_per_row, h->debug_flags);
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
