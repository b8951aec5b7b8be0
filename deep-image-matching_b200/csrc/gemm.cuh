// gemm.cuh - the tensor-core workhorse of libdimb200.
//
// One kernel template covers every dense contraction on the hot path:
//   C[M,N] = A[M,K] * B[N,K]^T      (LightGlue linears, 1x1 convs, similarity / distance matrices)
//   3x3 convolution as implicit GEMM (SuperPoint VGG encoder and heads): K = 9 taps x Cin, the A tile of a
//   tap is the NHWC activation tile shifted by (dy-1, dx-1), fetched by a 4D TMA whose out-of-bounds
//   fill implements the zero padding.
//
// Structure (sm_100a): 192 threads = 4 epilogue warps + 1 TMA producer warp + 1 MMA-issuer warp.
//   producer : cp.async.bulk.tensor -> 128B-swizzled smem stages, mbarrier full/empty ring
//   issuer   : one thread issues tcgen05.mma (M=128, N=BN, K=16) into an fp32 TMEM accumulator;
//              EXACT mode issues hi*hi + hi*lo + lo*hi per k-step (fp16 split operands)
//   epilogue : tcgen05.ld 32 columns at a time, thread r owns output row r -> fused epilogue functor
//
// A SIMT twin (simt_gemm_kernel) evaluates the same contraction on CUDA cores with the same
// epilogue functors; it is a debug/bisect aid (DIMB_TC=0), never the default.
#pragma once
#include "common.cuh"
#include "tc05.cuh"

struct TileCoord {
  int m0;         // GEMM: first global row of this 128-row tile
  int b, y0, x0;  // CONV: image index and top-left pixel of the 8x16 pixel tile
};

struct GemmArgs {
  int num_kb;      // K / 64
  int M;           // GEMM: valid rows of A
  int N;           // valid rows of B (output columns)
  int cin_blocks;  // CONV: Cin / 64
  int H, W;        // CONV: spatial size
  int tiles_x, tiles_y;
  const __half *Ah, *Al, *Bh, *Bl;  // raw operands (SIMT twin); Al/Bl null in FAST mode
  int lda, ldb;
};

constexpr int kTileM = 128;
constexpr int kConvTH = 8, kConvTW = 16;  // 8 rows x 16 cols of pixels = 128 GEMM rows; warp w owns rows 2w,2w+1

template <bool CONV>
__device__ __forceinline__ TileCoord make_tile_coord(const GemmArgs& g, int t) {
  TileCoord tc;
  if (CONV) {
    int per_img = g.tiles_x * g.tiles_y;
    tc.b = t / per_img;
    int rem = t - tc.b * per_img;
    tc.y0 = (rem / g.tiles_x) * kConvTH;
    tc.x0 = (rem % g.tiles_x) * kConvTW;
    tc.m0 = 0;
  } else {
    tc.m0 = t * kTileM;
    tc.b = tc.y0 = tc.x0 = 0;
  }
  return tc;
}

template <int BN, bool SPLIT>
struct GemmCfg {
  static constexpr int kPlanes = SPLIT ? 2 : 1;
  static constexpr int kABytes = kTileM * 128;  // 128 rows x 64 halfs
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kPlanes * (kABytes + kBBytes);
  static constexpr int kStagesRaw = (200 * 1024) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

template <int BN, bool SPLIT, bool CONV, class Epi>
__global__ void __launch_bounds__(192, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
               const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, GemmArgs g, Epi epi) {
  using Cfg = GemmCfg<BN, SPLIT>;
  using namespace tc05;
  TileCoord tc = make_tile_coord<CONV>(g, blockIdx.x);
  if (!CONV) tc.m0 = epi.m0_of(blockIdx.x);
  if (!epi.tile_active(tc)) return;  // CTA-uniform (device-side early exit / pruned rows)
  const int b_off = epi.b_row_offset(tc);  // B rows may depend on the tile (per-pair layer / other image)

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* tmem_full = empty + Cfg::kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_ptr, BN);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const int n0 = blockIdx.y * BN;

  if (warp == 4) {
    if (lane == 0) {  // ---------------- TMA producer
      tma_prefetch_desc(&tmAh);
      tma_prefetch_desc(&tmBh);
      if (SPLIT) {
        tma_prefetch_desc(&tmAl);
        tma_prefetch_desc(&tmBl);
      }
      for (int kb = 0; kb < g.num_kb; ++kb) {
        const int s = kb % Cfg::kStages;
        const uint32_t ph = (kb / Cfg::kStages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* st = smem + s * Cfg::kStageBytes;
        uint8_t* sb = st + Cfg::kPlanes * Cfg::kABytes;
        mbar_expect_tx(&full[s], Cfg::kStageBytes);
        if (CONV) {
          const int tap = kb / g.cin_blocks, cb = kb - tap * g.cin_blocks;
          const int dy = tap / 3, dx = tap - dy * 3;
          tma_load_4d(st, &tmAh, &full[s], cb * 64, tc.x0 + dx - 1, tc.y0 + dy - 1, tc.b);
          if (SPLIT) tma_load_4d(st + Cfg::kABytes, &tmAl, &full[s], cb * 64, tc.x0 + dx - 1, tc.y0 + dy - 1, tc.b);
        } else {
          tma_load_2d(st, &tmAh, &full[s], kb * 64, tc.m0);
          if (SPLIT) tma_load_2d(st + Cfg::kABytes, &tmAl, &full[s], kb * 64, tc.m0);
        }
        tma_load_2d(sb, &tmBh, &full[s], kb * 64, n0 + b_off);
        if (SPLIT) tma_load_2d(sb + Cfg::kBBytes, &tmBl, &full[s], kb * 64, n0 + b_off);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {  // ---------------- MMA issuer
      constexpr uint32_t idesc = make_idesc_f16(BN);
      for (int kb = 0; kb < g.num_kb; ++kb) {
        const int s = kb % Cfg::kStages;
        const uint32_t ph = (kb / Cfg::kStages) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after_sync();
        const uint32_t sa = smem_u32(smem + s * Cfg::kStageBytes);
        const uint32_t sb = sa + Cfg::kPlanes * Cfg::kABytes;
        const uint64_t a_h = make_sdesc_sw128(sa), b_h = make_sdesc_sw128(sb);
        const uint64_t a_l = make_sdesc_sw128(sa + Cfg::kABytes), b_l = make_sdesc_sw128(sb + Cfg::kBBytes);
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16) {
          mma_f16_ss(tmem_base, sdesc_advance_k(a_h, k16), sdesc_advance_k(b_h, k16), idesc, (kb | k16) != 0);
          if (SPLIT) {
            mma_f16_ss(tmem_base, sdesc_advance_k(a_h, k16), sdesc_advance_k(b_l, k16), idesc, 1);
            mma_f16_ss(tmem_base, sdesc_advance_k(a_l, k16), sdesc_advance_k(b_h, k16), idesc, 1);
          }
        }
        mma_commit(&empty[s]);  // smem stage reusable once these MMAs retire
      }
      mma_commit(tmem_full);  // accumulator complete
    }
  } else {  // ---------------- epilogue warps 0..3: thread r <-> accumulator row r (TMEM lane r)
    mbar_wait(tmem_full, 0);
    tc_fence_after_sync();
    const int r = warp * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      epi(tc, r, n0 + c0, v);
    }
    tc_fence_before_sync();
  }
  __syncthreads();
  if (warp == 5) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, BN);
  }
}

// ------------------------------------------------------------------ SIMT twin (debug path)
template <bool CONV>
__device__ __forceinline__ float simt_load_a(const GemmArgs& g, const TileCoord& tc, int row, int k) {
  size_t off;
  if (CONV) {
    const int cin = g.cin_blocks * 64;
    const int tap = k / cin, c = k - tap * cin;
    const int dy = tap / 3, dx = tap - dy * 3;
    const int y = tc.y0 + row / kConvTW + dy - 1, x = tc.x0 + row % kConvTW + dx - 1;
    if (y < 0 || y >= g.H || x < 0 || x >= g.W) return 0.f;
    off = ((static_cast<size_t>(tc.b) * g.H + y) * g.W + x) * cin + c;
  } else {
    const int gm = tc.m0 + row;
    if (gm >= g.M) return 0.f;
    off = static_cast<size_t>(gm) * g.lda + k;
  }
  float v = __half2float(g.Ah[off]);
  if (g.Al) v += __half2float(g.Al[off]);
  return v;
}

template <bool CONV, class Epi>
__global__ void __launch_bounds__(128) simt_gemm_kernel(GemmArgs g, Epi epi) {
  TileCoord tc = make_tile_coord<CONV>(g, blockIdx.x);
  if (!CONV) tc.m0 = epi.m0_of(blockIdx.x);
  if (!epi.tile_active(tc)) return;
  const int b_off = epi.b_row_offset(tc);
  __shared__ float As[128][33];
  __shared__ float Bs[32][33];
  const int t = threadIdx.x, n0 = blockIdx.y * 32;
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  const int K = g.num_kb * 64;
  for (int k0 = 0; k0 < K; k0 += 32) {
    const int kk = t & 31;
    for (int i = 0; i < 32; ++i) {
      const int row = i * 4 + (t >> 5);
      As[row][kk] = simt_load_a<CONV>(g, tc, row, k0 + kk);
    }
    for (int i = 0; i < 8; ++i) {
      const int nn = i * 4 + (t >> 5);
      float b = 0.f;
      if (n0 + nn < g.N) {
        const size_t off = static_cast<size_t>(n0 + nn + b_off) * g.ldb + k0 + kk;
        b = __half2float(g.Bh[off]);
        if (g.Bl) b += __half2float(g.Bl[off]);
      }
      Bs[nn][kk] = b;
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < 32; ++k) {
      const float a = As[t][k];
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = fmaf(a, Bs[j][k], acc[j]);
    }
    __syncthreads();
  }
  epi(tc, t, n0, acc);
}

// ------------------------------------------------------------------ launch
struct TcOperands {
  CUtensorMap Ah, Al, Bh, Bl;
};

template <int BN, bool SPLIT, bool CONV, class Epi>
int launch_tc(dimb_ctx* ctx, cudaStream_t st, const TcOperands& ops, const GemmArgs& g, const Epi& epi, int m_tiles,
              int n_pad) {
  using Cfg = GemmCfg<BN, SPLIT>;
  static bool attr_set = false;
  auto kern = tc_gemm_kernel<BN, SPLIT, CONV, Epi>;
  if (!attr_set) {
    DIMB_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  dim3 grid(m_tiles, n_pad / BN);
  kern<<<grid, 192, Cfg::kSmemBytes, st>>>(ops.Ah, ops.Al, ops.Bh, ops.Bl, g, epi);
  DIMB_LAUNCH_CHECK(ctx);
  return DIMB_OK;
}

// n_pad: output columns rounded up to a multiple of BN (B operand rows beyond N read as zero via TMA OOB fill).
template <int BN, bool CONV, class Epi>
int launch_gemm(dimb_ctx* ctx, cudaStream_t st, const TcOperands& ops, GemmArgs g, const Epi& epi, int m_tiles, int n_pad,
                const char* tag = "gemm") {
  if (m_tiles <= 0) return DIMB_OK;
  ProfScope prof(ctx, st, tag);
  if (ctx->use_tc) {
    if (ctx->precision == DIMB_PRECISION_EXACT) return launch_tc<BN, true, CONV, Epi>(ctx, st, ops, g, epi, m_tiles, n_pad);
    return launch_tc<BN, false, CONV, Epi>(ctx, st, ops, g, epi, m_tiles, n_pad);
  }
  if (ctx->precision != DIMB_PRECISION_EXACT) g.Al = g.Bl = nullptr;
  dim3 grid(m_tiles, n_pad / 32);
  simt_gemm_kernel<CONV, Epi><<<grid, 128, 0, st>>>(g, epi);
  DIMB_LAUNCH_CHECK(ctx);
  return DIMB_OK;
}

// ------------------------------------------------------------------ generic epilogues
// Every functor: tile_active(tc) (CTA-uniform) and operator()(tc, r, n, v): row r of the tile, v[j] = C[row][n+j].
// Optional hooks (defaults in EpiBase): m0_of(t) maps the tile index to its first A row; b_row_offset(tc) shifts
// the B rows a tile multiplies with (stacked per-layer weights, or "the other image" for similarity matrices).
struct EpiBase {
  __device__ int m0_of(int t) const { return t * kTileM; }
  __device__ int b_row_offset(const TileCoord&) const { return 0; }
  __device__ bool tile_active(const TileCoord&) const { return true; }
};

// fp32 store: out[row][n] = (acc + bias[n]) * scale, columns < n_valid, rows < m_valid.
struct EpiStoreF32 : EpiBase {
  float* out;
  const float* bias;  // may be null
  int ldc, n_valid, m_valid;
  float scale;
  __device__ void operator()(const TileCoord& tc, int r, int n, float (&v)[32]) const {
    const int row = tc.m0 + r;
    if (row >= m_valid) return;
    float* o = out + static_cast<size_t>(row) * ldc + n;
    if (n + 32 <= n_valid && (ldc & 3) == 0) {  // full, 16B-aligned chunk: 8 vector stores instead of 32 scalar ones
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 x;
        x.x = (v[4 * q] + (bias ? bias[n + 4 * q] : 0.f)) * scale;
        x.y = (v[4 * q + 1] + (bias ? bias[n + 4 * q + 1] : 0.f)) * scale;
        x.z = (v[4 * q + 2] + (bias ? bias[n + 4 * q + 2] : 0.f)) * scale;
        x.w = (v[4 * q + 3] + (bias ? bias[n + 4 * q + 3] : 0.f)) * scale;
        reinterpret_cast<float4*>(o)[q] = x;
      }
      return;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (n + j < n_valid) o[j] = (v[j] + (bias ? bias[n + j] : 0.f)) * scale;
  }
};

// 16-byte store of 8 consecutive halfs
__device__ __forceinline__ void store_half8(__half* dst, const __half (&h)[8]) {
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(h);
}

// writes 32 consecutive values as fp16 hi (+ lo) planes; dst pointers must be 16B aligned
__device__ __forceinline__ void store_split32(__half* hi, __half* lo, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    __half h[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split_f32(v[q * 8 + j], h[j], l[j]);
    store_half8(hi + q * 8, h);
    if (lo) store_half8(lo + q * 8, l);
  }
}

// fp16 hi/lo store: out[row][col_off + n] = (acc + bias[n]) * scale   (n_valid multiple of 32)
struct EpiStoreSplit : EpiBase {
  __half *hi, *lo;  // lo may be null (FAST)
  const float* bias;
  int ldc, col_off, n_valid, m_valid;
  float scale;
  __device__ void operator()(const TileCoord& tc, int r, int n, float (&v)[32]) const {
    const int row = tc.m0 + r;
    if (row >= m_valid || n >= n_valid) return;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (v[j] + (bias ? bias[n + j] : 0.f)) * scale;
    const size_t off = static_cast<size_t>(row) * ldc + col_off + n;
    store_split32(hi + off, lo ? lo + off : nullptr, v);
  }
};
