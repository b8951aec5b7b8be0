// conv_pair.cuh - the Cin = Cout = 64 3x3 convolutions of SuperPoint (conv1b / conv2a / conv2b) on CTA PAIRS (cta_group::2).
//
// Why: with 64 output channels a single-CTA tcgen05.mma reads 14 KB of shared memory per 16-deep k-step (A_hi 4 KB + [B_hi;B_lo]
// 4 KB, then A_lo 4 KB + B_hi 2 KB) for 96 tensor cycles of math - the shared-memory pipe (128 B/clk) needs 112 cycles, the layer is
// operand-fetch bound (ncu: tc pipe busy 74 %, math active 52 %).  A CTA pair issues ONE M = 256 instruction per product: each SM
// feeds its own 128 pixels (its own A halo stage) but only HALF of the N rows of B, so the per-SM traffic drops to 11 KB per k-step
// (86 cycles) and the math becomes the bound again.
//
// Layout per CTA (identical in both, so one descriptor serves both):
//   A ring  : SA stages x [hi | lo] halo boxes (18 x 10 px x 32 ch, 64-byte rows, SWIZZLE_64B) - gemm.cuh CONV 2
//   B panel : resident, per K tile kb (tap x 32-channel half block):
//               X_kb (64 rows)  CTA0: W_hi           CTA1: W_lo            <- the two halves of the stacked N = 128 operand [W_hi ; W_lo]
//               Y_kb (32 rows)  CTA0: W_hi[ 0:32]    CTA1: W_hi[32:64]     <- the two halves of the N = 64 operand W_hi
//   per k-step (one thread of the LEADER CTA):  D[:, 0:128] (+)= A_hi x X      (hi.hi | hi.lo)      M 256, N 128
//                                               D[:, 0: 64]  += A_lo x Y      (lo.hi)              M 256, N  64
// Barriers: TMA of both CTAs completes on the LEADER's full barriers (.cta_group::2 + mapa); tcgen05.commit multicasts the stage /
// accumulator hand-offs to both CTAs; the peer's epilogue releases the accumulator by a remote arrive on the leader's barrier.
#pragma once
#include "gemm.cuh"

namespace pairconv {
using namespace tc05;

constexpr int kBN = 64, kAccCols = 128, kEpiWarps = 4;
constexpr int kXBytes = 64 * 64, kYBytes = 32 * 64, kBTile = kXBytes + kYBytes;  // per K tile and CTA
constexpr int kPlane = (kHaloRows * 64 + 1023) / 1024 * 1024, kAStage = 2 * kPlane, kATx = 2 * kHaloRows * 64;
constexpr int kNkb = 18;  // 9 taps x 2 half blocks of 32 channels

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {  // shared::cluster address of `saddr` in CTA `rank`
  uint32_t d;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(d) : "r"(saddr), "r"(rank));
  return d;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Arrive on a barrier of another CTA of the cluster.  Default semantics (.release.cta), as in CUTLASS' ClusterBarrier::arrive: the data
// this orders is either TMEM (tcgen05.wait::ld + tcgen05.fence before it) or this CTA's own shared memory behind fence.proxy.async,
// both complete before the arrive is issued; .release.cluster would put a MEMBAR.ALL.GPU + ERRBAR in front of every arrive
// (13 % of the fused kernel's producer samples, profiles/r2_ncu_conv1ab_fused_8warps_full.txt).
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads whose completion is signalled on a barrier given by its shared::cluster address (the leader's)
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_f16_m256(int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(256 >> 4) << 24);
}
__device__ __forceinline__ void mma2_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair once all previously issued MMAs have completed
__device__ __forceinline__ void mma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}

struct PairArgs {
  int H, W, tiles_x, tiles_y, total;  // total tiles (even)
};

// EPW epilogue warps per CTA: 4 (one per TMEM lane quarter), or 8 (two per quarter, one 32-channel chunk each) for the layer without
// pooling (conv2a), whose epilogue writes four times the bytes per tile and otherwise holds both accumulators of a pair back
template <class Epi, int EPW = kEpiWarps>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((EPW + 2) * 32, 1)
conv64_pair_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl, const __grid_constant__ CUtensorMap tmWh64,
                   const __grid_constant__ CUtensorMap tmWl64, const __grid_constant__ CUtensorMap tmWh32, PairArgs pa, Epi epi, int SA) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by OFFSET from the __shared__ array (not by integer-casting the pointer): the compiler keeps the shared
  // address space and emits LDS / STS instead of generic LD / ST for every access derived from it
  uint8_t* smem = smem_raw + ((1024u - (tc05::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = sA + SA * kAStage;
  uint64_t* fullA = reinterpret_cast<uint64_t*>(sB + kNkb * kBTile);  // [SA]   waited by the leader's issuer only
  uint64_t* emptyA = fullA + SA;                                      // [SA]   per CTA (commit multicast)
  uint64_t* fullB = emptyA + SA;                                      // [1]    leader: both CTAs' resident weights
  uint64_t* tfull = fullB + 1;                                        // [2]    per CTA (commit multicast)
  uint64_t* tempty = tfull + 2;                                       // [2]    leader: both CTAs' epilogue threads
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) {
      mbar_init(&fullA[s], 1);
      mbar_init(&emptyA[s], 1);
    }
    mbar_init(&fullB[0], 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 2 * EPW);
    }
    fence_barrier_init();
  }
  if (warp == EPW + 1) tmem_alloc2(tmem_ptr, 2 * kAccCols);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // barriers of both CTAs are initialised before any remote arrive / peer TMA completion
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const int n_super = pa.total / 2, n_pairs = static_cast<int>(gridDim.x) / 2, pair = static_cast<int>(cluster_id_x());
  GemmArgs g{};
  g.tiles_x = pa.tiles_x;
  g.tiles_y = pa.tiles_y;

  if (warp == EPW) {  // ---------------- TMA producer (every CTA loads its own operands; completion on the LEADER's barriers)
    const uint32_t fullB_leader = mapa(smem_u32(&fullB[0]), 0);
    if (elect_one()) {
      tma_prefetch_desc(&tmAh);
      tma_prefetch_desc(&tmAl);
      if (leader) mbar_expect_tx(&fullB[0], 2 * kNkb * kBTile);
      for (int kb = 0; kb < kNkb; ++kb) {
        uint8_t* x = sB + kb * kBTile;
        tma2_load_2d(x, leader ? &tmWh64 : &tmWl64, fullB_leader, kb * 32, 0);        // X: W_hi (leader) / W_lo (peer), 64 rows
        tma2_load_2d(x + kXBytes, &tmWh32, fullB_leader, kb * 32, leader ? 0 : 32);  // Y: W_hi rows [0,32) / [32,64)
      }
    }
    __syncwarp();
    uint32_t it = 0;
    for (int u = pair; u < n_super; u += n_pairs) {
      const TileCoord tc = make_tile_coord<2>(g, 2 * u + static_cast<int>(rank));
      for (int o = 0; o < 2; ++o) {
        const int s = it % SA;
        mbar_wait(&emptyA[s], ((it / SA) & 1) ^ 1);
        if (elect_one()) {
          if (leader) mbar_expect_tx(&fullA[s], 2 * kATx);
          const uint32_t bar = mapa(smem_u32(&fullA[s]), 0);
          uint8_t* st = sA + s * kAStage;
          tma2_load_4d(st, &tmAh, bar, o * 32, tc.x0 - 1, tc.y0 - 1, tc.b);
          tma2_load_4d(st + kPlane, &tmAl, bar, o * 32, tc.x0 - 1, tc.y0 - 1, tc.b);
        }
        __syncwarp();
        ++it;
      }
    }
  } else if (warp == EPW + 1) {  // ---------------- MMA issuer: leader CTA only
    if (leader) {
      constexpr uint32_t idesc128 = make_idesc_f16_m256(128), idesc64 = make_idesc_f16_m256(64);
      uint32_t it = 0, tcount = 0;
      mbar_wait(&fullB[0], 0);
      tc_fence_after_sync();
      for (int u = pair; u < n_super; u += n_pairs) {
        const uint32_t acc = tcount & 1;
        mbar_wait(&tempty[acc], ((tcount >> 1) & 1) ^ 1);  // both epilogues have drained this accumulator
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * kAccCols;
        uint32_t accumulate = 0;
        for (int o = 0; o < 2; ++o) {
          const int s = it % SA;
          mbar_wait(&fullA[s], (it / SA) & 1);
          tc_fence_after_sync();
          const uint32_t a_base = smem_u32(sA + s * kAStage);
          const uint64_t a0h = make_sdesc(a_base, (kHaloTW + 2) * 64, kLayoutSw64), a0l = make_sdesc(a_base + kPlane, (kHaloTW + 2) * 64, kLayoutSw64);
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int kb = tap * 2 + o;  // weights are [Cout][tap * 64 + c]
              const uint64_t tap16 = static_cast<uint64_t>(((tap / 3) * (kHaloTW + 2) + tap % 3) * 4);
              const uint32_t xb = smem_u32(sB + kb * kBTile);
              const uint64_t bx = make_sdesc(xb, 512, kLayoutSw64), by = make_sdesc(xb + kXBytes, 512, kLayoutSw64);
#pragma unroll
              for (int k16 = 0; k16 < 2; ++k16) {
                mma2_f16_ss(d_tmem, sdesc_advance_k(a0h + tap16, k16), sdesc_advance_k(bx, k16), idesc128, (tap | k16) ? 1u : accumulate);
                mma2_f16_ss(d_tmem, sdesc_advance_k(a0l + tap16, k16), sdesc_advance_k(by, k16), idesc64, 1);
              }
            }
            mma2_commit(&emptyA[s]);
          }
          __syncwarp();
          accumulate = 1;
          ++it;
        }
        if (elect_one()) mma2_commit(&tfull[acc]);
        __syncwarp();
        ++tcount;
      }
    }
  } else {  // ---------------- epilogue warps (each CTA drains its own 128 TMEM lanes = its own pixel tile)
    const uint32_t tempty_leader[2] = {mapa(smem_u32(&tempty[0]), 0), mapa(smem_u32(&tempty[1]), 0)};
    uint32_t tcount = 0;
    const int q = warp & 3, r = q * 32 + lane;
    for (int u = pair; u < n_super; u += n_pairs) {
      const TileCoord tc = make_tile_coord<2>(g, 2 * u + static_cast<int>(rank));
      const uint32_t acc = tcount & 1;
      mbar_wait(&tfull[acc], (tcount >> 1) & 1);
      tc_fence_after_sync();
#pragma unroll 1
      for (int c0 = (warp >> 2) * 32; c0 < kBN; c0 += 32 * (EPW / 4)) {  // EPW = 8: warp / 4 picks the warp's single 32-channel chunk
        float v[32], v2[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kAccCols + c0, v);
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kAccCols + kBN + c0, v2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += v2[j];
        if (c0 + 32 * (EPW / 4) >= kBN) {  // last TMEM read of this warp: release the accumulator on the leader's barrier - ONE remote arrive per
          tc_fence_before_sync();  // warp (a cluster-scope release fence each; 128 of them per tile showed up as 9 % of the samples)
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(tempty_leader[acc]);
        }
        epi(tc, r, c0, v, nullptr);
      }
      ++tcount;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // the leader's MMAs read the peer's shared memory, the peer arrives on the leader's barriers: leave together
  if (warp == EPW + 1) {
    tc_fence_after_sync();
    tmem_dealloc2(tmem_base, 2 * kAccCols);
  }
}

// launch helper: returns DIMB_ERR_UNSUPPORTED if the shape does not fit the pair kernel (odd tile count)
template <class Epi, int EPW = kEpiWarps>
int launch_conv64_pair(dimb_ctx* ctx, cudaStream_t st, const CUtensorMap& Ah, const CUtensorMap& Al, const CUtensorMap& Wh64, const CUtensorMap& Wl64,
                       const CUtensorMap& Wh32, int B, int H, int W, const Epi& epi) {
  PairArgs pa;
  pa.H = H;
  pa.W = W;
  pa.tiles_x = ceil_div(W, kHaloTW);
  pa.tiles_y = ceil_div(H, kHaloTH);
  pa.total = B * pa.tiles_x * pa.tiles_y;
  if (pa.total & 1) return DIMB_ERR_UNSUPPORTED;
  const int budget = 232448 - 1024 - 1024 - kNkb * kBTile;
  int SA = budget / kAStage;
  if (SA > 6) SA = 6;
  const int smem = SA * kAStage + kNkb * kBTile + 1024 + 1024;
  auto kern = conv64_pair_kernel<Epi, EPW>;
  DIMB_TRY(dimb_func_smem(ctx, kern, smem));
  int grid = ctx->num_sms & ~1;
  if (grid > pa.total) grid = pa.total;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3((EPW + 2) * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  DIMB_CUDA_OK(ctx, cudaLaunchKernelEx(&cfg, kern, Ah, Al, Wh64, Wl64, Wh32, pa, epi, SA));
  ctx->launches++;
  return DIMB_OK;
}


// =====================================================================================================================
// conv1a fused into conv1b (SuperPoint's first two layers, superpoint.py:161-162) on the CTA-pair kernel.
//
// conv1a (1 -> 64 channels, 0.6 GMAC per image) exists only to feed conv1b: as a kernel of its own it writes 268 MB per 1024^2 image
// (fp16 hi + lo planes) that conv1b reads straight back - 20 GB per 37-pair step, 4.7 ms at HBM speed.  Here four producer warps
// compute the (16+2) x (8+2)-pixel conv1a halo tile of each conv1b tile on the CUDA cores - same arithmetic, same order, same
// hi/lo split as sp_conv1a_kernel - and write it directly into the SWIZZLE_64B A stages that TMA would have filled (32 channels per
// stage; halo pixels outside the image are conv1b's zero padding, not relu(bias)).  The activation never touches HBM.
// fp32 image rows come in through a small double-buffered patch (20 x 12 pixels) prefetched one tile ahead.
struct Conv1aWeights {
  float v[576 + 64];  // tap-major [9][64] + bias [64], as sp_conv1a_kernel takes them
};

constexpr int kProdWarps = 8, kProdPix = 23, kPatchH = kHaloTH + 4, kPatchW = kHaloTW + 4;  // 8 warps x 23 halo pixels >= 180; 20 x 12 input pixels per tile

template <class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((kEpiWarps + kProdWarps + 1) * 32, 1)
conv1ab_pair_kernel(const __grid_constant__ Conv1aWeights c1, const float* __restrict__ img, const __grid_constant__ CUtensorMap tmWh64,
                    const __grid_constant__ CUtensorMap tmWl64, const __grid_constant__ CUtensorMap tmWh32, PairArgs pa, Epi epi, int SA) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by OFFSET from the __shared__ array (not by integer-casting the pointer): the compiler keeps the shared
  // address space and emits LDS / STS instead of generic LD / ST for every access derived from it
  uint8_t* smem = smem_raw + ((1024u - (tc05::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = sA + SA * kAStage;
  uint64_t* fullA = reinterpret_cast<uint64_t*>(sB + kNkb * kBTile);  // [SA]  leader: 2 CTAs x kProdWarps arrivals
  uint64_t* emptyA = fullA + SA;                                      // [SA]  per CTA (commit multicast)
  uint64_t* fullB = emptyA + SA;                                      // [1]   leader
  uint64_t* tfull = fullB + 1;                                        // [2]   per CTA
  uint64_t* tempty = tfull + 2;                                       // [2]   leader
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  float* patch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(fullA) + 512);  // [2][kPatchH][kPatchW]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  constexpr int kIssuer = kEpiWarps + kProdWarps;
  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) {
      mbar_init(&fullA[s], 2 * kProdWarps);
      mbar_init(&emptyA[s], 1);
    }
    mbar_init(&fullB[0], 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 2 * kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == kIssuer) tmem_alloc2(tmem_ptr, 2 * kAccCols);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const int n_super = pa.total / 2, n_pairs = static_cast<int>(gridDim.x) / 2, pair = static_cast<int>(cluster_id_x());
  GemmArgs g{};
  g.tiles_x = pa.tiles_x;
  g.tiles_y = pa.tiles_y;

  if (warp >= kEpiWarps && warp < kIssuer) {  // ---------------- conv1a producers: 8 warps, 23 halo pixels each (one lane per pixel, 32 channels)
    const int pw = warp - kEpiWarps, pt = pw * 32 + lane;  // 0..255
    if (pw == 0) {  // resident conv1b weights (as conv64_pair_kernel)
      const uint32_t fullB_leader = mapa(smem_u32(&fullB[0]), 0);
      if (elect_one()) {
        if (leader) mbar_expect_tx(&fullB[0], 2 * kNkb * kBTile);
        for (int kb = 0; kb < kNkb; ++kb) {
          uint8_t* x = sB + kb * kBTile;
          tma2_load_2d(x, leader ? &tmWh64 : &tmWl64, fullB_leader, kb * 32, 0);
          tma2_load_2d(x + kXBytes, &tmWh32, fullB_leader, kb * 32, leader ? 0 : 32);
        }
      }
      __syncwarp();
    }
    // input patch of a tile: image rows y0-2 .. y0+17, columns x0-2 .. x0+9, normalised (/255, IEEE division like the reference), 0 outside
    auto load_patch = [&](const TileCoord& tc, float (&pv)[2]) {
#pragma unroll
      for (int k = 0; k < 1; ++k) {
        const int e = pt;
        pv[k] = 0.f;
        if (e < kPatchH * kPatchW) {
          const int yy = tc.y0 - 2 + e / kPatchW, xx = tc.x0 - 2 + e % kPatchW;
          if (yy >= 0 && yy < pa.H && xx >= 0 && xx < pa.W) pv[k] = img[(static_cast<size_t>(tc.b) * pa.H + yy) * pa.W + xx];  // raw: the load stays in flight
        }
      }
    };
    auto store_patch = [&](int buf, const float (&pv)[2]) {
      if (pt < kPatchH * kPatchW) patch[buf * kPatchH * kPatchW + pt] = __fdiv_rn(pv[0], 255.f);  // image / 255, IEEE division like the reference
    };
    float pv[2];
    int u = pair;
    if (u < n_super) {
      load_patch(make_tile_coord<2>(g, 2 * u + static_cast<int>(rank)), pv);
      store_patch(0, pv);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kProdWarps * 32) : "memory");  // producers only
    uint32_t it = 0, tl = 0;
    for (; u < n_super; u += n_pairs, ++tl) {
      const TileCoord tc = make_tile_coord<2>(g, 2 * u + static_cast<int>(rank));
      const int un = u + n_pairs;
      if (un < n_super) load_patch(make_tile_coord<2>(g, 2 * un + static_cast<int>(rank)), pv);  // in flight while this tile is computed
      const float* P = patch + (tl & 1) * kPatchH * kPatchW;
      for (int o = 0; o < 2; ++o) {
        const int s = it % SA;
        mbar_wait(&emptyA[s], ((it / SA) & 1) ^ 1);
        uint8_t* st = sA + s * kAStage;
        {
          const int hp = pw * kProdPix + lane;  // halo pixel of this lane
          if (lane < kProdPix && hp < kHaloRows) {
            const int hy = hp / (kHaloTW + 2), hx = hp - hy * (kHaloTW + 2);
            const int gy = tc.y0 - 1 + hy, gx = tc.x0 - 1 + hx;
            const bool inside = gy >= 0 && gy < pa.H && gx >= 0 && gx < pa.W;
            float a[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) a[t] = P[(hy + t / 3) * kPatchW + hx + t % 3];
            float acc[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[c] = c1.v[576 + o * 32 + c];
#pragma unroll
            for (int t = 0; t < 9; ++t) {
#pragma unroll
              for (int c = 0; c < 32; ++c) acc[c] = fmaf(a[t], c1.v[t * 64 + o * 32 + c], acc[c]);
            }
            const uint32_t rowoff = static_cast<uint32_t>(hp) * 64u, sw = (static_cast<uint32_t>(hp) >> 1) & 3u;
#pragma unroll
            for (int q = 0; q < 4; ++q) {  // 8 channels = one 16-byte chunk, SWIZZLE_64B: chunk ^ ((row >> 1) & 3)
              __half2 h[4], l[4];
#pragma unroll
              for (int j = 0; j < 4; ++j)
                split2_f32(inside ? fmaxf(acc[q * 8 + 2 * j], 0.f) : 0.f, inside ? fmaxf(acc[q * 8 + 2 * j + 1], 0.f) : 0.f, h[j], l[j]);
              const uint32_t off = rowoff + ((static_cast<uint32_t>(q) ^ sw) << 4);
              *reinterpret_cast<uint4*>(st + off) = *reinterpret_cast<uint4*>(h);
              *reinterpret_cast<uint4*>(st + kPlane + off) = *reinterpret_cast<uint4*>(l);
            }
          }
        }
        fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(mapa(smem_u32(&fullA[s]), 0));
        ++it;
      }
      // the next tile's patch goes to the OTHER buffer (last read during the previous tile, which ended with this barrier)
      if (un < n_super) store_patch((tl + 1) & 1, pv);
      asm volatile("bar.sync 1, %0;" ::"n"(kProdWarps * 32) : "memory");
    }
  } else if (warp == kIssuer) {  // ---------------- MMA issuer: leader CTA only (identical to conv64_pair_kernel)
    if (leader) {
      constexpr uint32_t idesc128 = make_idesc_f16_m256(128), idesc64 = make_idesc_f16_m256(64);
      uint32_t it = 0, tcount = 0;
      mbar_wait(&fullB[0], 0);
      tc_fence_after_sync();
      for (int u = pair; u < n_super; u += n_pairs) {
        const uint32_t acc = tcount & 1;
        mbar_wait(&tempty[acc], ((tcount >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * kAccCols;
        uint32_t accumulate = 0;
        for (int o = 0; o < 2; ++o) {
          const int s = it % SA;
          mbar_wait(&fullA[s], (it / SA) & 1);
          tc_fence_after_sync();
          const uint32_t a_base = smem_u32(sA + s * kAStage);
          const uint64_t a0h = make_sdesc(a_base, (kHaloTW + 2) * 64, kLayoutSw64), a0l = make_sdesc(a_base + kPlane, (kHaloTW + 2) * 64, kLayoutSw64);
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int kb = tap * 2 + o;
              const uint64_t tap16 = static_cast<uint64_t>(((tap / 3) * (kHaloTW + 2) + tap % 3) * 4);
              const uint32_t xb = smem_u32(sB + kb * kBTile);
              const uint64_t bx = make_sdesc(xb, 512, kLayoutSw64), by = make_sdesc(xb + kXBytes, 512, kLayoutSw64);
#pragma unroll
              for (int k16 = 0; k16 < 2; ++k16) {
                mma2_f16_ss(d_tmem, sdesc_advance_k(a0h + tap16, k16), sdesc_advance_k(bx, k16), idesc128, (tap | k16) ? 1u : accumulate);
                mma2_f16_ss(d_tmem, sdesc_advance_k(a0l + tap16, k16), sdesc_advance_k(by, k16), idesc64, 1);
              }
            }
            mma2_commit(&emptyA[s]);
          }
          __syncwarp();
          accumulate = 1;
          ++it;
        }
        if (elect_one()) mma2_commit(&tfull[acc]);
        __syncwarp();
        ++tcount;
      }
    }
  } else {  // ---------------- epilogue warps
    const uint32_t tempty_leader[2] = {mapa(smem_u32(&tempty[0]), 0), mapa(smem_u32(&tempty[1]), 0)};
    uint32_t tcount = 0;
    const int q = warp & 3, r = q * 32 + lane;
    for (int u = pair; u < n_super; u += n_pairs) {
      const TileCoord tc = make_tile_coord<2>(g, 2 * u + static_cast<int>(rank));
      const uint32_t acc = tcount & 1;
      mbar_wait(&tfull[acc], (tcount >> 1) & 1);
      tc_fence_after_sync();
#pragma unroll 1
      for (int c0 = 0; c0 < kBN; c0 += 32) {
        float v[32], v2[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kAccCols + c0, v);
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kAccCols + kBN + c0, v2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += v2[j];
        if (c0 + 32 >= kBN) {  // last TMEM read of this warp: release the accumulator on the leader's barrier - ONE remote arrive per
          tc_fence_before_sync();  // warp (a cluster-scope release fence each; 128 of them per tile showed up as 9 % of the samples)
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(tempty_leader[acc]);
        }
        epi(tc, r, c0, v, nullptr);
      }
      ++tcount;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == kIssuer) {
    tc_fence_after_sync();
    tmem_dealloc2(tmem_base, 2 * kAccCols);
  }
}

// =====================================================================================================================
// conv1a on the tensor cores too (conv1ab_mma_pair_kernel, DIMB_FUSE1A=2).
//
// The SIMT producers above spend ~2400 issue cycles per tile and scheduler on the 9 x 64 FMAs of every halo pixel - 70 % of the
// 3456 cycles conv1b's MMAs take - and hold the tensor pipe at 47 %.  conv1a is a 9 -> 64 linear map per pixel: as an im2col GEMM
// [halo pixels x 16] x [16 x 64] (K = 9 taps + a constant 1 that carries the bias + zeros) it is ONE 16-deep k-step - six M = 256 MMAs
// per tile pair (two 128-row blocks x hi.hi + hi.lo + lo.hi), 192 tensor cycles.  What is left for the CUDA cores: writing the
// im2col rows (180 x 10 values) and draining the result - tcgen05.ld, ReLU, zero outside the image (conv1b's padding), hi/lo split,
// store into the SWIZZLE_64B A stages - about a sixth of the instructions.
//   smem per CTA : im2col buffer (one A-stage-sized block: 180 rows x 64 B, K columns 0-15 used, SWIZZLE_64B like every other operand
//                  here; the MMA reads 256 rows, rows >= 180 are whatever follows - never used) | SA = 3 A stages | conv1b panel | W1 tiles
//   TMEM         : conv1b accumulators 2 x 128 columns | D1 block a (halo rows 0-127) 64 | D1 block b (rows 128-255) 64
//   front warps  : 8 (warp % 4 = TMEM lane quarter): all write im2col rows; warps 0-3 drain block a, warps 4-5 block b
//   barriers     : iFull (front -> leader), iEmpty / d1Full (commit, both CTAs), d1Free (drain -> leader), fullA / emptyA as before
//   order        : front  : im2col(0); for t: { [wait iEmpty(t); im2col(t+1)]; wait d1Full(t); drain(t) }
//                  issuer : c1a(0); for t: { [wait iFull(t+1), d1Free(t); c1a(t+1)]; c1b(t) }
constexpr int kFrontWarps = 8, kDrainWarps = 6, kW1Bytes = 2 * 32 * 64;  // per CTA: 32 rows of W1_hi + 32 rows of W1_lo, 64-byte rows

template <class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((kEpiWarps + kFrontWarps + 1) * 32, 1)
conv1ab_mma_pair_kernel(const float* __restrict__ img, const __grid_constant__ CUtensorMap tmW1h, const __grid_constant__ CUtensorMap tmW1l,
                        const __grid_constant__ CUtensorMap tmWh64, const __grid_constant__ CUtensorMap tmWl64,
                        const __grid_constant__ CUtensorMap tmWh32, PairArgs pa, Epi epi, int SA) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc05::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sI = smem;                    // im2col rows [hi plane | lo plane]
  uint8_t* sA = sI + kAStage;
  uint8_t* sB = sA + SA * kAStage;
  uint8_t* sW1 = sB + kNkb * kBTile;     // [W1_hi half | W1_lo half]
  uint64_t* fullA = reinterpret_cast<uint64_t*>(sW1 + kW1Bytes);  // [SA]  leader: 2 CTAs x kDrainWarps arrivals
  uint64_t* emptyA = fullA + SA;         // [SA]  per CTA (commit multicast)
  uint64_t* fullB = emptyA + SA;         // [1]   leader
  uint64_t* tfull = fullB + 1;           // [2]   per CTA
  uint64_t* tempty = tfull + 2;          // [2]   leader
  uint64_t* iFull = tempty + 2;          // [1]   leader: 2 CTAs x kFrontWarps
  uint64_t* iEmpty = iFull + 1;          // [1]   per CTA (commit multicast)
  uint64_t* d1Full = iEmpty + 1;         // [1]   per CTA (commit multicast)
  uint64_t* d1Free = d1Full + 1;         // [1]   leader: 2 CTAs x kDrainWarps
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(d1Free + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  constexpr int kIssuer = kEpiWarps + kFrontWarps;
  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) {
      mbar_init(&fullA[s], 2 * kDrainWarps);
      mbar_init(&emptyA[s], 1);
    }
    mbar_init(&fullB[0], 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 2 * kEpiWarps);
    }
    mbar_init(iFull, 2 * kFrontWarps);
    mbar_init(iEmpty, 1);
    mbar_init(d1Full, 1);
    mbar_init(d1Free, 2 * kDrainWarps);
    fence_barrier_init();
  }
  if (warp == kIssuer) tmem_alloc2(tmem_ptr, 512);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tD1 = tmem_base + 2 * kAccCols;  // block a at +0, block b at +64
  const int n_super = pa.total / 2, n_pairs = static_cast<int>(gridDim.x) / 2, pair = static_cast<int>(cluster_id_x());
  GemmArgs g{};
  g.tiles_x = pa.tiles_x;
  g.tiles_y = pa.tiles_y;

  if (warp >= kEpiWarps && warp < kIssuer) {  // ---------------- front warps: im2col rows in, conv1a result out
    const int fw = warp - kEpiWarps, ft = fw * 32 + lane;  // 0..255: halo pixel whose im2col row this thread writes (if < 180)
    if (fw == 0) {  // resident weights: the conv1b panel (as conv64_pair_kernel) + this CTA's halves of W1_hi / W1_lo
      const uint32_t fullB_leader = mapa(smem_u32(&fullB[0]), 0);
      if (elect_one()) {
        if (leader) mbar_expect_tx(&fullB[0], 2 * (kNkb * kBTile + kW1Bytes));
        for (int kb = 0; kb < kNkb; ++kb) {
          uint8_t* x = sB + kb * kBTile;
          tma2_load_2d(x, leader ? &tmWh64 : &tmWl64, fullB_leader, kb * 32, 0);
          tma2_load_2d(x + kXBytes, &tmWh32, fullB_leader, kb * 32, leader ? 0 : 32);
        }
        tma2_load_2d(sW1, &tmW1h, fullB_leader, 0, leader ? 0 : 32);
        tma2_load_2d(sW1 + kW1Bytes / 2, &tmW1l, fullB_leader, 0, leader ? 0 : 32);
      }
      __syncwarp();
    }
    const int hy = ft / (kHaloTW + 2), hx = ft - hy * (kHaloTW + 2);
    // the 9 taps of halo pixel (hy, hx) of a tile: image rows y0-2+hy .. +2, columns x0-2+hx .. +2; 0 outside the image (conv1a's padding)
    auto load_taps = [&](const TileCoord& tc, float (&a)[9]) {
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int yy = tc.y0 - 2 + hy + t / 3, xx = tc.x0 - 2 + hx + t % 3;
        a[t] = (ft < kHaloRows && yy >= 0 && yy < pa.H && xx >= 0 && xx < pa.W) ? img[(static_cast<size_t>(tc.b) * pa.H + yy) * pa.W + xx] : 0.f;
      }
    };
    auto write_im2col = [&](const float (&a)[9]) {  // K = [9 taps (image / 255, IEEE division like the reference), 1, 0 ...]: chunks 0 and 1 of the 64-byte row
      if (ft < kHaloRows) {
        float v[16];
#pragma unroll
        for (int t = 0; t < 9; ++t) v[t] = __fdiv_rn(a[t], 255.f);
        v[9] = 1.f;
#pragma unroll
        for (int t = 10; t < 16; ++t) v[t] = 0.f;
        const uint32_t rowoff = static_cast<uint32_t>(ft) * 64u, sw = (static_cast<uint32_t>(ft) >> 1) & 3u;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          __half2 h[4], l[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) split2_f32(v[q * 8 + 2 * j], v[q * 8 + 2 * j + 1], h[j], l[j]);
          const uint32_t off = rowoff + ((static_cast<uint32_t>(q) ^ sw) << 4);
          *reinterpret_cast<uint4*>(sI + off) = *reinterpret_cast<uint4*>(h);
          *reinterpret_cast<uint4*>(sI + kPlane + off) = *reinterpret_cast<uint4*>(l);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(mapa(smem_u32(iFull), 0));
    };
    float taps[9];
    int u = pair;
    if (u < n_super) {
      load_taps(make_tile_coord<2>(g, 2 * u + static_cast<int>(rank)), taps);
      write_im2col(taps);
    }
    const bool drains = fw < kDrainWarps;
    const int blk = fw >> 2, hp = blk * 128 + (fw & 3) * 32 + lane;  // halo pixel = D1 row this thread drains
    const uint32_t tRow = tD1 + blk * 64 + (static_cast<uint32_t>((fw & 3) * 32) << 16);
    const uint32_t fullA_leader0 = mapa(smem_u32(&fullA[0]), 0), d1Free_leader = mapa(smem_u32(d1Free), 0);
    uint32_t tl = 0;
    for (; u < n_super; u += n_pairs, ++tl) {
      const TileCoord tc = make_tile_coord<2>(g, 2 * u + static_cast<int>(rank));
      const int un = u + n_pairs;
      if (un < n_super) {  // the next tile's im2col rows, as soon as conv1a of this tile has consumed the buffer
        load_taps(make_tile_coord<2>(g, 2 * un + static_cast<int>(rank)), taps);
        mbar_wait(iEmpty, tl & 1);
        write_im2col(taps);
      }
      if (!drains) continue;
      mbar_wait(d1Full, tl & 1);
      tc_fence_after_sync();
      float v[64];
      tmem_ld32(tRow, v);
      tmem_ld32(tRow + 32, v + 32);
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(d1Free_leader);  // D1 is in registers: conv1a of the next tile may overwrite it
      const int py = hp / (kHaloTW + 2), pxx = hp - py * (kHaloTW + 2);
      const int gy = tc.y0 - 1 + py, gx = tc.x0 - 1 + pxx;
      const bool live = hp < kHaloRows;
      const bool inside = live && gy >= 0 && gy < pa.H && gx >= 0 && gx < pa.W;  // outside the image: conv1b's zero padding, not relu(bias)
      const uint32_t rowoff = static_cast<uint32_t>(hp) * 64u, sw = (static_cast<uint32_t>(hp) >> 1) & 3u;
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        const uint32_t it = 2 * tl + static_cast<uint32_t>(o);
        const int s = it % SA;
        mbar_wait(&emptyA[s], ((it / SA) & 1) ^ 1);
        uint8_t* st = sA + s * kAStage;
        if (live) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            __half2 h[4], l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              split2_f32(inside ? fmaxf(v[o * 32 + q * 8 + 2 * j], 0.f) : 0.f, inside ? fmaxf(v[o * 32 + q * 8 + 2 * j + 1], 0.f) : 0.f, h[j], l[j]);
            const uint32_t off = rowoff + ((static_cast<uint32_t>(q) ^ sw) << 4);
            *reinterpret_cast<uint4*>(st + off) = *reinterpret_cast<uint4*>(h);
            *reinterpret_cast<uint4*>(st + kPlane + off) = *reinterpret_cast<uint4*>(l);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(fullA_leader0 + static_cast<uint32_t>(s) * 8u);
      }
    }
  } else if (warp == kIssuer) {  // ---------------- MMA issuer: leader CTA only
    if (leader) {
      constexpr uint32_t idesc128 = make_idesc_f16_m256(128), idesc64 = make_idesc_f16_m256(64);
      uint32_t it = 0, tcount = 0;
      mbar_wait(&fullB[0], 0);
      tc_fence_after_sync();
      const uint32_t i_base = smem_u32(sI), w1 = smem_u32(sW1);
      const uint64_t w1h = make_sdesc(w1, 512, kLayoutSw64), w1l = make_sdesc(w1 + kW1Bytes / 2, 512, kLayoutSw64);
      auto issue_c1a = [&](uint32_t t) {  // conv1a of tile t: D1 = I_hi W1_hi + I_hi W1_lo + I_lo W1_hi, two 128-row blocks
        mbar_wait(iFull, t & 1);
        if (t > 0) mbar_wait(d1Free, (t - 1) & 1);
        tc_fence_after_sync();
        if (elect_one()) {
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const uint64_t ah = make_sdesc(i_base + b * 8192, 512, kLayoutSw64), al = make_sdesc(i_base + kPlane + b * 8192, 512, kLayoutSw64);
            mma2_f16_ss(tD1 + b * 64, ah, w1h, idesc64, 0);
            mma2_f16_ss(tD1 + b * 64, ah, w1l, idesc64, 1);
            mma2_f16_ss(tD1 + b * 64, al, w1h, idesc64, 1);
          }
          mma2_commit(d1Full);
          mma2_commit(iEmpty);
        }
        __syncwarp();
      };
      uint32_t tl = 0;
      if (pair < n_super) issue_c1a(0);
      for (int u = pair; u < n_super; u += n_pairs, ++tl) {
        if (u + n_pairs < n_super) issue_c1a(tl + 1);  // queued ahead of conv1b(tl): its drain overlaps conv1b's MMAs
        const uint32_t acc = tcount & 1;
        mbar_wait(&tempty[acc], ((tcount >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * kAccCols;
        uint32_t accumulate = 0;
        for (int o = 0; o < 2; ++o) {
          const int s = it % SA;
          mbar_wait(&fullA[s], (it / SA) & 1);
          tc_fence_after_sync();
          const uint32_t a_base = smem_u32(sA + s * kAStage);
          const uint64_t a0h = make_sdesc(a_base, (kHaloTW + 2) * 64, kLayoutSw64), a0l = make_sdesc(a_base + kPlane, (kHaloTW + 2) * 64, kLayoutSw64);
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int kb = tap * 2 + o;
              const uint64_t tap16 = static_cast<uint64_t>(((tap / 3) * (kHaloTW + 2) + tap % 3) * 4);
              const uint32_t xb = smem_u32(sB + kb * kBTile);
              const uint64_t bx = make_sdesc(xb, 512, kLayoutSw64), by = make_sdesc(xb + kXBytes, 512, kLayoutSw64);
#pragma unroll
              for (int k16 = 0; k16 < 2; ++k16) {
                mma2_f16_ss(d_tmem, sdesc_advance_k(a0h + tap16, k16), sdesc_advance_k(bx, k16), idesc128, (tap | k16) ? 1u : accumulate);
                mma2_f16_ss(d_tmem, sdesc_advance_k(a0l + tap16, k16), sdesc_advance_k(by, k16), idesc64, 1);
              }
            }
            mma2_commit(&emptyA[s]);
          }
          __syncwarp();
          accumulate = 1;
          ++it;
        }
        if (elect_one()) mma2_commit(&tfull[acc]);
        __syncwarp();
        ++tcount;
      }
    }
  } else {  // ---------------- epilogue warps (as conv1ab_pair_kernel)
    const uint32_t tempty_leader[2] = {mapa(smem_u32(&tempty[0]), 0), mapa(smem_u32(&tempty[1]), 0)};
    uint32_t tcount = 0;
    const int q = warp & 3, r = q * 32 + lane;
    for (int u = pair; u < n_super; u += n_pairs) {
      const TileCoord tc = make_tile_coord<2>(g, 2 * u + static_cast<int>(rank));
      const uint32_t acc = tcount & 1;
      mbar_wait(&tfull[acc], (tcount >> 1) & 1);
      tc_fence_after_sync();
#pragma unroll 1
      for (int c0 = 0; c0 < kBN; c0 += 32) {
        float v[32], v2[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kAccCols + c0, v);
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kAccCols + kBN + c0, v2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += v2[j];
        if (c0 + 32 >= kBN) {
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(tempty_leader[acc]);
        }
        epi(tc, r, c0, v, nullptr);
      }
      ++tcount;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == kIssuer) {
    tc_fence_after_sync();
    tmem_dealloc2(tmem_base, 512);
  }
}

template <class Epi>
int launch_conv1ab_mma_pair(dimb_ctx* ctx, cudaStream_t st, const float* d_img, const CUtensorMap& W1h, const CUtensorMap& W1l,
                            const CUtensorMap& Wh64, const CUtensorMap& Wl64, const CUtensorMap& Wh32, int B, int H, int W, const Epi& epi) {
  PairArgs pa;
  pa.H = H;
  pa.W = W;
  pa.tiles_x = ceil_div(W, kHaloTW);
  pa.tiles_y = ceil_div(H, kHaloTH);
  pa.total = B * pa.tiles_x * pa.tiles_y;
  if (pa.total & 1) return DIMB_ERR_UNSUPPORTED;
  const int SA = 3;  // im2col buffer + 3 A stages + the resident panels = 214 KB
  const int smem = (1 + SA) * kAStage + kNkb * kBTile + kW1Bytes + 1024 + 1024;
  auto kern = conv1ab_mma_pair_kernel<Epi>;
  DIMB_TRY(dimb_func_smem(ctx, kern, smem));
  int grid = ctx->num_sms & ~1;
  if (grid > pa.total) grid = pa.total;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3((kEpiWarps + kFrontWarps + 1) * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  DIMB_CUDA_OK(ctx, cudaLaunchKernelEx(&cfg, kern, d_img, W1h, W1l, Wh64, Wl64, Wh32, pa, epi, SA));
  ctx->launches++;
  return DIMB_OK;
}

template <class Epi>
int launch_conv1ab_pair(dimb_ctx* ctx, cudaStream_t st, const Conv1aWeights& c1, const float* d_img, const CUtensorMap& Wh64, const CUtensorMap& Wl64,
                        const CUtensorMap& Wh32, int B, int H, int W, const Epi& epi) {
  PairArgs pa;
  pa.H = H;
  pa.W = W;
  pa.tiles_x = ceil_div(W, kHaloTW);
  pa.tiles_y = ceil_div(H, kHaloTH);
  pa.total = B * pa.tiles_x * pa.tiles_y;
  if (pa.total & 1) return DIMB_ERR_UNSUPPORTED;
  const int budget = 232448 - 1024 - 4096 - kNkb * kBTile;  // barriers + the two input patches live in the 4 KB after the B panel
  int SA = budget / kAStage;
  if (SA > 6) SA = 6;
  const int smem = SA * kAStage + kNkb * kBTile + 1024 + 4096;
  auto kern = conv1ab_pair_kernel<Epi>;
  DIMB_TRY(dimb_func_smem(ctx, kern, smem));
  int grid = ctx->num_sms & ~1;
  if (grid > pa.total) grid = pa.total;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3((kEpiWarps + kProdWarps + 1) * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  DIMB_CUDA_OK(ctx, cudaLaunchKernelEx(&cfg, kern, c1, d_img, Wh64, Wl64, Wh32, pa, epi, SA));
  ctx->launches++;
  return DIMB_OK;
}

}  // namespace pairconv
