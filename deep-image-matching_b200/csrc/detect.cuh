// detect.cuh - the dense-score-map -> keypoint-list machinery shared by the SuperPoint and ALIKED extractors:
// simple_nms (identical in both reference models: superpoint.py:47-63, aliked.py:66-89), threshold + border
// compaction in row-major order, and top-k selection (radix select + bitonic sort).
#pragma once
#include "common.cuh"

namespace {

constexpr int kNmsTile = 64;
constexpr int kChunk = 4096;       // pixels per compaction chunk
constexpr int kSelThreads = 1024;
constexpr int kMaxTopK = 16384;

// ------------------------------------------------------------------ simple_nms (superpoint.py:47-63)
// Tile of 64x64 outputs + halo 5r; every stage is a separable (2r+1)^2 window max in shared memory.
// 1024 threads as 32 x 32 (one CTA per SM - shared memory bound - so the warps have to come from the CTA itself): loops run over (row, column) directly - no integer divisions in the hot loops.
template <int RT>  // RT > 0: compile-time radius (window in registers); RT = -1: runtime radius
__device__ __forceinline__ void win_max(const float* src, float* tmp, float* dst, int S, int PS, int k, int r_rt) {
  const int r = RT >= 0 ? RT : r_rt;
  // src valid on margin k-r; writes dst on margin k.  Each thread produces a run of 8 outputs from a register
  // sliding window (8 + 2r loads instead of 8 * (2r+1)).
  const int nthr = blockDim.x;
  {  // row pass: runs of 8 columns; lanes walk down the rows so that a warp's smem accesses hit 32 different rows
    const int rows = S - 2 * (k - r), cols = S - 2 * k, segs = (cols + 7) >> 3;
    for (int t = threadIdx.x; t < rows * segs; t += nthr) {
      const int seg = t / rows, i = k - r + (t - seg * rows);
      const int j0 = k + seg * 8, n = min(8, S - k - j0);
      const float* p = src + i * PS + j0;
      if (RT >= 0 && n == 8) {
        float w[8 + 2 * (RT >= 0 ? RT : 0)];
#pragma unroll
        for (int d = 0; d < 8 + 2 * RT; ++d) w[d] = p[d - RT];
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          float m = w[o];
#pragma unroll
          for (int d = 1; d <= 2 * RT; ++d) m = fmaxf(m, w[o + d]);
          tmp[i * PS + j0 + o] = m;
        }
      } else {
        for (int o = 0; o < n; ++o) {
          float m = p[o];
          for (int d = 1; d <= r; ++d) m = fmaxf(m, fmaxf(p[o - d], p[o + d]));
          tmp[i * PS + j0 + o] = m;
        }
      }
    }
  }
  __syncthreads();
  {  // column pass: runs of 8 rows; lanes along columns (conflict-free)
    const int cols = S - 2 * k, segs = (cols + 7) >> 3;
    for (int t = threadIdx.x; t < cols * segs; t += nthr) {
      const int seg = t / cols, j = k + (t - seg * cols);
      const int i0 = k + seg * 8, n = min(8, S - k - i0);
      const float* p = tmp + i0 * PS + j;
      if (RT >= 0 && n == 8) {
        float w[8 + 2 * (RT >= 0 ? RT : 0)];
#pragma unroll
        for (int d = 0; d < 8 + 2 * RT; ++d) w[d] = p[(d - RT) * PS];
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          float m = w[o];
#pragma unroll
          for (int d = 1; d <= 2 * RT; ++d) m = fmaxf(m, w[o + d]);
          dst[(i0 + o) * PS + j] = m;
        }
      } else {
        for (int o = 0; o < n; ++o) {
          float m = p[o * PS];
          for (int d = 1; d <= r; ++d) m = fmaxf(m, fmaxf(p[(o - d) * PS], p[(o + d) * PS]));
          dst[(i0 + o) * PS + j] = m;
        }
      }
    }
  }
  __syncthreads();
}

template <int RT>
__global__ void __launch_bounds__(1024) sp_nms_kernel(const float* __restrict__ scores, float* __restrict__ out, int H, int W, int r,
                                                     int T) {
  extern __shared__ float nsm[];
  const int S = T + 10 * r, PS = S | 1;  // odd row pitch: the row pass walks lanes down the rows, an even pitch costs 2-way bank conflicts
  float* s0 = nsm;             // scores, -inf outside the image
  float* xa = s0 + S * PS;     // mask-as-float / suppressed scores
  float* tmp = xa + S * PS;
  float* wm = tmp + S * PS;    // window max
  unsigned char* msk = reinterpret_cast<unsigned char*>(wm + S * PS);  // max_mask
  unsigned char* sup = msk + S * PS;                                   // supp_mask of the current round
  const int b = blockIdx.z, ty0 = blockIdx.y * T - 5 * r, tx0 = blockIdx.x * T - 5 * r;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, nty = blockDim.x >> 5;
  const float* sc = scores + static_cast<size_t>(b) * H * W;
  for (int i = ty; i < S; i += nty) {
    const int gy = ty0 + i;
    for (int j = tx; j < S; j += 32) {
      const int gx = tx0 + j;
      const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
      s0[i * PS + j] = in ? sc[static_cast<size_t>(gy) * W + gx] : -INFINITY;
    }
  }
  __syncthreads();
  // max_mask = scores == max_pool(scores)                              margin r
  win_max<RT>(s0, tmp, wm, S, PS, r, r);
  for (int i = ty; i < S; i += nty)
    for (int j = tx; j < S; j += 32) {
      const int e = i * PS + j;
      const bool v = i >= r && i < S - r && j >= r && j < S - r;
      const bool in = s0[e] != -INFINITY;  // inside the image (scores are softmax outputs > 0)
      const unsigned char m = (v && in && s0[e] == wm[e]) ? 1 : 0;
      msk[e] = m;
      xa[e] = m ? 1.f : 0.f;
    }
  __syncthreads();
  for (int round = 0; round < 2; ++round) {
    const int kb = r + 2 * r * round;  // margin on which max_mask is valid: r, then 3r
    // supp_mask = max_pool(max_mask.float()) > 0                        margin kb + r
    win_max<RT>(xa, tmp, wm, S, PS, kb + r, r);
    for (int i = ty; i < S; i += nty)
      for (int j = tx; j < S; j += 32) {
        const int e = i * PS + j, k = kb + r;
        const bool v = i >= k && i < S - k && j >= k && j < S - k;
        const bool in = s0[e] != -INFINITY;
        const unsigned char sp = (v && in && wm[e] > 0.f) ? 1 : 0;
        sup[e] = sp;
        // supp_scores = where(supp_mask, 0, scores); -inf outside the image (max_pool2d padding)
        xa[e] = in ? (sp ? 0.f : s0[e]) : -INFINITY;
      }
    __syncthreads();
    // new_max_mask = supp_scores == max_pool(supp_scores)               margin kb + 2r
    win_max<RT>(xa, tmp, wm, S, PS, kb + 2 * r, r);
    for (int i = ty; i < S; i += nty)
      for (int j = tx; j < S; j += 32) {
        const int e = i * PS + j, k = kb + 2 * r;
        const bool v = i >= k && i < S - k && j >= k && j < S - k;
        unsigned char m = 0;
        if (v && s0[e] != -INFINITY) m = (msk[e] | ((xa[e] == wm[e]) && !sup[e])) ? 1 : 0;  // max_mask | (new_max_mask & ~supp_mask)
        msk[e] = m;
      }
    __syncthreads();
    if (round == 0) {
      for (int i = ty; i < S; i += nty)
        for (int j = tx; j < S; j += 32) xa[i * PS + j] = msk[i * PS + j] ? 1.f : 0.f;
      __syncthreads();
    }
  }
  float* o = out + static_cast<size_t>(b) * H * W;
  for (int i = 5 * r + ty; i < 5 * r + T; i += nty) {
    const int gy = ty0 + i;
    if (gy >= H) break;
    for (int j = 5 * r + tx; j < 5 * r + T; j += 32) {
      const int gx = tx0 + j;
      if (gx < W) o[static_cast<size_t>(gy) * W + gx] = msk[i * PS + j] ? s0[i * PS + j] : 0.f;
    }
  }
}

// ------------------------------------------------------------------ simple_nms, second cut (default for radii 1..5)
// Same arithmetic (exact float equality, -inf outside the image, the reference's five max-pools), a third of the instructions:
//   * max_mask / supp_mask live as BIT rows (one 32-bit word per 32 columns, built with __ballot_sync): the two binary max-pools of
//     the reference (max_pool(max_mask.float()) > 0) become funnel shifts and ORs over a few hundred words per tile;
//   * only the three float max-pools remain; each is a row pass (register sliding window, lanes walking down the rows) into `tmp`
//     and a column pass whose result is consumed in registers - compared, balloted into the mask words, or written out - so no
//     window-max, mask-as-float or suppressed-score array exists;
//   * window maxima by doubling (max of 2, of 4, of 8 neighbours) instead of 2r compares per output.
// Region = 64 x 64 outputs + 5r halo; columns are rounded up to whole mask words and framed by 8 columns of -inf.
template <int R>
struct Nms2 {
  static constexpr int S = kNmsTile + 10 * R;     // region rows / meaningful columns
  static constexpr int NW = (S + 31) / 32;         // mask words per row
  static constexpr int SW = NW * 32;               // computed columns
  static constexpr int PAD = 8;                    // -inf frame (>= R)
  static constexpr int PS = (SW + 2 * PAD) | 1;    // odd pitch: the row passes walk lanes down the rows
  static constexpr int NWP = NW + 2;               // mask row pitch: one zero word on either side
  static constexpr int kThreads = 512;
  static constexpr size_t kSmem = static_cast<size_t>(2) * S * PS * 4 + static_cast<size_t>(4) * S * NWP * 4;
};

// m[o] = max(w[o .. o + 2R]), o = 0..7
template <int R>
__device__ __forceinline__ void window_max8(const float (&w)[8 + 2 * R], float (&m)[8]) {
  constexpr int K = 2 * R + 1, N = 8 + 2 * R;
  if constexpr (R == 0) {
#pragma unroll
    for (int o = 0; o < 8; ++o) m[o] = w[o];
  } else {
    float p2[N - 1];
#pragma unroll
    for (int d = 0; d < N - 1; ++d) p2[d] = fmaxf(w[d], w[d + 1]);
    if constexpr (K == 3) {
#pragma unroll
      for (int o = 0; o < 8; ++o) m[o] = fmaxf(p2[o], w[o + 2]);
    } else {
      float p4[N - 3];
#pragma unroll
      for (int d = 0; d < N - 3; ++d) p4[d] = fmaxf(p2[d], p2[d + 2]);
      if constexpr (K < 8) {
#pragma unroll
        for (int o = 0; o < 8; ++o) m[o] = fmaxf(p4[o], p4[o + K - 4]);
      } else {
        float p8[N - 7];
#pragma unroll
        for (int d = 0; d < N - 7; ++d) p8[d] = fmaxf(p4[d], p4[d + 4]);
#pragma unroll
        for (int o = 0; o < 8; ++o) m[o] = fmaxf(p8[o], p8[o + K - 8]);
      }
    }
  }
}

// row pass of one max-pool over rows [K0 - R, S - K0 + R) (K0 = margin of the pool's output): tmp = window max along x of
// (USE_SUP ? where(supp, 0, s0) : s0).  Only the 8-column segments that overlap the output columns [K0, S - K0) are computed.
template <int R, int K0, bool USE_SUP>
__device__ __forceinline__ void nms2_row_pass(const float* __restrict__ s0, float* __restrict__ tmp, const uint32_t* __restrict__ SUP) {
  using G = Nms2<R>;
  constexpr int S = G::S, PS = G::PS, PAD = G::PAD, NWP = G::NWP;
  constexpr int row_lo = K0 - R, rows = S - 2 * (K0 - R), seg_lo = K0 >> 3, nseg = ((S - K0 - 1) >> 3) + 1 - seg_lo;
  for (int t = threadIdx.x; t < rows * nseg; t += G::kThreads) {
    const int sg = t / rows, i = row_lo + (t - sg * rows), j0 = (seg_lo + sg) * 8;
    const float* p = s0 + i * PS + PAD + j0 - R;
    float w[8 + 2 * R];
#pragma unroll
    for (int d = 0; d < 8 + 2 * R; ++d) w[d] = p[d];
    if (USE_SUP) {
      const int a = j0 - R + 32;  // first window column, shifted by the zero pad word
      const uint32_t bits = __funnelshift_r(SUP[i * NWP + (a >> 5)], SUP[i * NWP + (a >> 5) + 1], a & 31);
#pragma unroll
      for (int d = 0; d < 8 + 2 * R; ++d)
        if ((bits >> d) & 1u) w[d] = 0.f;  // supp_scores = where(supp_mask, 0, scores); supp is never set outside the image
    }
    float m[8];
    window_max8<R>(w, m);
    float* o = tmp + i * PS + PAD + j0;
#pragma unroll
    for (int q = 0; q < 8; ++q) o[q] = m[q];
  }
}

// supp_mask = max_pool(max_mask) > 0 as bit rows: horizontal dilation by funnel shifts, vertical by OR-ing 2R+1 rows; & inside-image
template <int R>
__device__ __forceinline__ void nms2_dilate(const uint32_t* __restrict__ M, uint32_t* __restrict__ HB, uint32_t* __restrict__ SUP,
                                            const uint32_t* __restrict__ IN) {
  using G = Nms2<R>;
  constexpr int S = G::S, NW = G::NW, NWP = G::NWP;
  for (int it = threadIdx.x; it < S * NW; it += G::kThreads) {
    const int i = it / NW, w = it - i * NW;
    const uint32_t l = M[i * NWP + w], c = M[i * NWP + 1 + w], r = M[i * NWP + 2 + w];
    uint32_t h = c;
#pragma unroll
    for (int d = 1; d <= R; ++d) h |= __funnelshift_l(l, c, d) | __funnelshift_r(c, r, d);
    HB[i * NWP + 1 + w] = h;
  }
  __syncthreads();
  for (int it = threadIdx.x; it < S * NW; it += G::kThreads) {
    const int i = it / NW, w = it - i * NW;
    uint32_t v = 0;
#pragma unroll
    for (int d = -R; d <= R; ++d) {
      const int ii = i + d;
      if (ii >= 0 && ii < S) v |= HB[ii * NWP + 1 + w];
    }
    SUP[i * NWP + 1 + w] = v & IN[i * NWP + 1 + w];
  }
  __syncthreads();
}

// column pass of one max-pool with its consumer.  STAGE 0: max_mask = scores == max_pool(scores).  STAGE 1: max_mask |=
// (supp_scores == max_pool(supp_scores)) & ~supp_mask.  STAGE 2: the same, then where(max_mask, scores, 0) -> global (tile centre).
template <int R, int K0, int STAGE>
__device__ __forceinline__ void nms2_col_pass(const float* __restrict__ s0, const float* __restrict__ tmp, uint32_t* __restrict__ M,
                                              const uint32_t* __restrict__ SUP, const uint32_t* __restrict__ IN, float* __restrict__ o,
                                              int H, int W, int ty0, int tx0) {
  using G = Nms2<R>;
  constexpr int S = G::S, PS = G::PS, PAD = G::PAD, NWP = G::NWP;
  constexpr int nchunk = (S - 2 * K0 + 7) / 8, w_lo = K0 >> 5, nw = ((S - K0 - 1) >> 5) + 1 - w_lo;
  const int lane = threadIdx.x & 31;
  for (int it = threadIdx.x >> 5; it < nchunk * nw; it += G::kThreads / 32) {
    const int ch = it / nw, w = w_lo + (it - ch * nw), i0 = K0 + ch * 8, c = w * 32 + lane;
    float win[8 + 2 * R];
#pragma unroll
    for (int d = 0; d < 8 + 2 * R; ++d) win[d] = tmp[min(i0 - R + d, S - 1) * PS + PAD + c];
    float m[8];
    window_max8<R>(win, m);
    const bool colv = c >= K0 && c < S - K0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int i = i0 + q;
      if (i >= S - K0) break;  // warp-uniform
      const float sv = s0[i * PS + PAD + c];
      const bool in = (IN[i * NWP + 1 + w] >> lane) & 1u;
      if (STAGE == 0) {
        const uint32_t bal = __ballot_sync(0xffffffffu, colv && in && sv == m[q]);
        if (lane == 0) M[i * NWP + 1 + w] = bal;
      } else {
        const bool sup = (SUP[i * NWP + 1 + w] >> lane) & 1u;
        const uint32_t fresh = __ballot_sync(0xffffffffu, colv && in && !sup && sv == m[q]);
        if (STAGE == 1) {
          if (lane == 0) M[i * NWP + 1 + w] |= fresh;  // one thread reads and writes the word: no intra-warp hazard
        } else {
          const uint32_t mw = M[i * NWP + 1 + w] | fresh;  // M is read-only in the last stage
          const int gy = ty0 + i, gx = tx0 + c;  // K0 = 5R: rows are the tile centre by construction
          if (colv && gy < H && gx < W) o[static_cast<size_t>(gy) * W + gx] = ((mw >> lane) & 1u) ? sv : 0.f;
        }
      }
    }
  }
}

template <int R>
__global__ void __launch_bounds__(512) sp_nms2_kernel(const float* __restrict__ scores, float* __restrict__ out, int H, int W) {
  using G = Nms2<R>;
  constexpr int S = G::S, NW = G::NW, SW = G::SW, PS = G::PS, PAD = G::PAD, NWP = G::NWP;
  extern __shared__ float nsm2[];
  float* s0 = nsm2;                // scores, -inf outside the image and in the frame
  float* tmp = s0 + S * PS;        // row-pass result of the current pool
  uint32_t* M = reinterpret_cast<uint32_t*>(tmp + S * PS);  // max_mask bit rows
  uint32_t* SUP = M + S * NWP;     // supp_mask of the current round
  uint32_t* IN = SUP + S * NWP;    // inside-the-image bits
  uint32_t* HB = IN + S * NWP;     // horizontally dilated max_mask
  const int b = blockIdx.z, ty0 = blockIdx.y * kNmsTile - 5 * R, tx0 = blockIdx.x * kNmsTile - 5 * R;
  const int tid = threadIdx.x, lane = tid & 31;
  const float* sc = scores + static_cast<size_t>(b) * H * W;
  for (int i = tid; i < 4 * S * NWP; i += G::kThreads) M[i] = 0u;
  for (int i = tid; i < S * 2 * PAD; i += G::kThreads) {
    const int row = i / (2 * PAD), c = i - row * 2 * PAD;
    s0[row * PS + (c < PAD ? c : SW + c)] = -INFINITY;
  }
  __syncthreads();
  for (int it = tid >> 5; it < S * NW; it += G::kThreads / 32) {
    const int i = it / NW, w = it - i * NW, c = w * 32 + lane, gy = ty0 + i, gx = tx0 + c;
    const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
    s0[i * PS + PAD + c] = in ? sc[static_cast<size_t>(gy) * W + gx] : -INFINITY;
    const uint32_t bal = __ballot_sync(0xffffffffu, in);
    if (lane == 0) IN[i * NWP + 1 + w] = bal;
  }
  __syncthreads();
  float* o = out + static_cast<size_t>(b) * H * W;
  // max_mask = scores == max_pool(scores)                                           valid on margin R
  nms2_row_pass<R, R, false>(s0, tmp, SUP);
  __syncthreads();
  nms2_col_pass<R, R, 0>(s0, tmp, M, SUP, IN, o, H, W, ty0, tx0);
  __syncthreads();
  // round 0: supp on margin 2R, new maxima on margin 3R
  nms2_dilate<R>(M, HB, SUP, IN);
  nms2_row_pass<R, 3 * R, true>(s0, tmp, SUP);
  __syncthreads();
  nms2_col_pass<R, 3 * R, 1>(s0, tmp, M, SUP, IN, o, H, W, ty0, tx0);
  __syncthreads();
  // round 1: supp on margin 4R, final mask on margin 5R = the 64 x 64 centre
  nms2_dilate<R>(M, HB, SUP, IN);
  nms2_row_pass<R, 5 * R, true>(s0, tmp, SUP);
  __syncthreads();
  nms2_col_pass<R, 5 * R, 2>(s0, tmp, M, SUP, IN, o, H, W, ty0, tx0);
}

// ------------------------------------------------------------------ candidate compaction in row-major order
__device__ __forceinline__ bool sp_is_cand(float v, int p, int W, int H, float thr, int border) {
  const int y = p / W, x = p - y * W;
  return v > thr && y >= border && y < H - border && x >= border && x < W - border;  // superpoint.py:183,66-71
}

__global__ void __launch_bounds__(256) sp_count_kernel(const float* __restrict__ nms, int* __restrict__ chunk_count, int H, int W,
                                                       float thr, int border, int nchunks, const float* __restrict__ thr_dev) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const float* s = nms + static_cast<size_t>(b) * H * W;
  if (thr_dev) thr = thr_dev[b];  // threshold decided on the device (ALIKED's mean fallback)
  int cnt = 0;
  const int base = chunk * kChunk + threadIdx.x * 16;
  for (int i = 0; i < 16; ++i) {
    const int p = base + i;
    if (p < H * W && sp_is_cand(s[p], p, W, H, thr, border)) ++cnt;
  }
  __shared__ int red[8];
#pragma unroll
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < 8; ++i) t += red[i];
    chunk_count[b * nchunks + chunk] = t;
  }
}

__global__ void sp_scan_kernel(const int* __restrict__ chunk_count, int* __restrict__ chunk_off, int* __restrict__ cand_count,
                               int nchunks) {
  // one thread block per image; nchunks is small (H*W/4096): serial scan by thread 0 is fine
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int i = 0; i < nchunks; ++i) {
      chunk_off[b * nchunks + i] = acc;
      acc += chunk_count[b * nchunks + i];
    }
    cand_count[b] = acc;
  }
}

__global__ void __launch_bounds__(256) sp_compact_kernel(const float* __restrict__ nms, const int* __restrict__ chunk_off,
                                                         int* __restrict__ cand_idx, float* __restrict__ cand_score, int H,
                                                         int W, float thr, int border, int nchunks,
                                                         const float* __restrict__ thr_dev) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const float* s = nms + static_cast<size_t>(b) * H * W;
  if (thr_dev) thr = thr_dev[b];
  const int base = chunk * kChunk + threadIdx.x * 16;
  float v[16];
  int cnt = 0;
  unsigned flags = 0;
  for (int i = 0; i < 16; ++i) {
    const int p = base + i;
    v[i] = p < H * W ? s[p] : 0.f;
    if (p < H * W && sp_is_cand(v[i], p, W, H, thr, border)) {
      flags |= 1u << i;
      ++cnt;
    }
  }
  // block exclusive scan of cnt (256 threads)
  __shared__ int wsum[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[wid] = inc;
  __syncthreads();
  int woff = 0;
  for (int i = 0; i < wid; ++i) woff += wsum[i];
  int pos = chunk_off[b * nchunks + chunk] + woff + inc - cnt;
  int* ci = cand_idx + static_cast<size_t>(b) * H * W;
  float* cs = cand_score + static_cast<size_t>(b) * H * W;
  for (int i = 0; i < 16; ++i)
    if (flags & (1u << i)) {
      ci[pos] = base + i;
      cs[pos] = v[i];
      ++pos;
    }
}

// ------------------------------------------------------------------ top-k (superpoint.py:74-78): radix select + bitonic sort
// One CTA per image.  If count <= K (or K < 0): keep everything in row-major order.  Otherwise pick the K
// largest scores (ties at the cut resolved by smaller pixel index) and order them score-desc, index-asc.
__global__ void __launch_bounds__(kSelThreads) sp_select_kernel(const int* __restrict__ cand_idx, const float* __restrict__ cand_score,
                                                                const int* __restrict__ cand_count, int* __restrict__ sel_idx,
                                                                float* __restrict__ sel_score, int* __restrict__ sel_count,
                                                                int HW, int K, int cap, int sort_cap) {
  extern __shared__ unsigned long long keys[];  // sort_cap entries
  __shared__ unsigned hist[256];
  __shared__ unsigned s_prefix, s_remaining, s_ngreater, s_tiepos;
  const int b = blockIdx.x, t = threadIdx.x;
  const int C = cand_count[b];
  const int* ci = cand_idx + static_cast<size_t>(b) * HW;
  const float* cs = cand_score + static_cast<size_t>(b) * HW;
  int* oi = sel_idx + static_cast<size_t>(b) * cap;
  float* os = sel_score + static_cast<size_t>(b) * cap;
  if (K < 0 || C <= K) {
    if (t == 0) sel_count[b] = C;  // host checks C <= cap
    for (int i = t; i < C && i < cap; i += blockDim.x) {
      oi[i] = ci[i];
      os[i] = cs[i];
    }
    return;
  }
  // ---- radix select of the K-th largest score (scores > 0 -> float bits are order preserving)
  if (t == 0) {
    s_prefix = 0;
    s_remaining = K;
  }
  __syncthreads();
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (t < 256) hist[t] = 0;
    __syncthreads();
    const unsigned prefix = s_prefix;
    const unsigned himask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
    for (int i = t; i < C; i += blockDim.x) {
      const unsigned u = __float_as_uint(cs[i]);
      if ((u & himask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (t == 0) {
      unsigned rem = s_remaining, d = 255;
      for (;; --d) {  // walk digits from large to small
        if (hist[d] >= rem) break;
        rem -= hist[d];
        if (d == 0) break;
      }
      s_prefix = prefix | (d << shift);
      s_remaining = rem;  // how many elements equal to the final threshold are still needed
    }
    __syncthreads();
  }
  const unsigned T = s_prefix;       // bit pattern of the K-th largest score
  const unsigned need_ties = s_remaining;
  if (t == 0) {
    s_ngreater = 0;
    s_tiepos = 0;
  }
  __syncthreads();
  // ---- gather: strictly greater first (any order, sorted below), then the first `need_ties` ties by index.
  // key = (~scorebits << 32) | pixel index : ascending key order == score desc, index asc
  for (int i = t; i < C; i += blockDim.x) {
    const unsigned u = __float_as_uint(cs[i]);
    if (u > T) {
      const unsigned pos = atomicAdd(&s_ngreater, 1u);
      keys[pos] = (static_cast<unsigned long long>(~u) << 32) | static_cast<unsigned>(ci[i]);
    }
  }
  __syncthreads();
  const unsigned G = s_ngreater;  // == K - need_ties
  // ties: candidates are stored in increasing pixel index, so rank among ties = number of earlier ties
  for (int base = 0; base < C; base += blockDim.x) {
    const int i = base + t;
    const bool tie = i < C && __float_as_uint(cs[i]) == T;
    // block-wide ordered rank via ballot + warp counts
    __shared__ unsigned wcnt[32];
    const unsigned bal = __ballot_sync(0xffffffffu, tie);
    if ((t & 31) == 0) wcnt[t >> 5] = __popc(bal);
    __syncthreads();
    unsigned before = s_tiepos;
    for (int wv = 0; wv < (t >> 5); ++wv) before += wcnt[wv];
    before += __popc(bal & ((1u << (t & 31)) - 1u));
    if (tie && before < need_ties)
      keys[G + before] = (static_cast<unsigned long long>(~T) << 32) | static_cast<unsigned>(ci[i]);
    __syncthreads();
    if (t == 0) {
      unsigned tot = 0;
      for (int wv = 0; wv < 32; ++wv) tot += wcnt[wv];
      s_tiepos += tot;
    }
    __syncthreads();
  }
  // ---- bitonic sort of K keys padded to a power of two
  int P = 1;
  while (P < K) P <<= 1;
  for (int i = K + t; i < P; i += blockDim.x) keys[i] = ~0ull;
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = t; i < P; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], c = keys[ixj];
          const bool up = (i & k) == 0;
          if ((a > c) == up) {
            keys[i] = c;
            keys[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  for (int i = t; i < K; i += blockDim.x) {
    oi[i] = static_cast<int>(keys[i] & 0xffffffffull);
    os[i] = __uint_as_float(~static_cast<unsigned>(keys[i] >> 32));
  }
  if (t == 0) sel_count[b] = K;
}


// launches simple_nms on a [B][H][W] score map
inline int launch_nms(dimb_ctx* ctx, cudaStream_t st, const float* scores, float* out, int B, int H, int W, int r) {
  if (ctx->nms_ver == 2 && r >= 1 && r <= 5) {  // bit-mask kernel (DIMB_NMS=1 selects the first cut below)
    dim3 grid2(ceil_div(W, kNmsTile), ceil_div(H, kNmsTile), B);
    auto launch2 = [&](auto kern, size_t smem2) -> int {
      DIMB_TRY(dimb_func_smem(ctx, kern, static_cast<int>(smem2)));
      kern<<<grid2, 512, smem2, st>>>(scores, out, H, W);
      return DIMB_OK;
    };
    switch (r) {
      case 1: DIMB_TRY(launch2(sp_nms2_kernel<1>, Nms2<1>::kSmem)); break;
      case 2: DIMB_TRY(launch2(sp_nms2_kernel<2>, Nms2<2>::kSmem)); break;
      case 3: DIMB_TRY(launch2(sp_nms2_kernel<3>, Nms2<3>::kSmem)); break;
      case 4: DIMB_TRY(launch2(sp_nms2_kernel<4>, Nms2<4>::kSmem)); break;
      default: DIMB_TRY(launch2(sp_nms2_kernel<5>, Nms2<5>::kSmem)); break;
    }
    DIMB_LAUNCH_CHECK(ctx);
    return DIMB_OK;
  }
  int T = kNmsTile;  // 64x64 outputs per CTA unless the 5r halo no longer fits in shared memory
  if (static_cast<size_t>(T + 10 * r) * ((T + 10 * r) | 1) * (4 * sizeof(float) + 2) > 220 * 1024) T = 32;
  const int S = T + 10 * r;
  const size_t smem = static_cast<size_t>(S) * (S | 1) * (4 * sizeof(float) + 2);
  dim3 grid(ceil_div(W, T), ceil_div(H, T), B);
  auto launch = [&](auto kern) -> int {
    DIMB_TRY(dimb_func_smem(ctx, kern, static_cast<int>(smem)));
    kern<<<grid, 1024, smem, st>>>(scores, out, H, W, r, T);
    return DIMB_OK;
  };
  switch (r) {  // the reference's configurations use 2/3 (pipelines), 4 (defaults) and 5 (tile preselection)
    case 2: DIMB_TRY(launch(sp_nms_kernel<2>)); break;
    case 3: DIMB_TRY(launch(sp_nms_kernel<3>)); break;
    case 4: DIMB_TRY(launch(sp_nms_kernel<4>)); break;
    case 5: DIMB_TRY(launch(sp_nms_kernel<5>)); break;
    default: DIMB_TRY(launch(sp_nms_kernel<-1>)); break;
  }
  DIMB_LAUNCH_CHECK(ctx);
  return DIMB_OK;
}

}  // namespace
