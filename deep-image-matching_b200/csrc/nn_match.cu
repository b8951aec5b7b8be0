// nn_match.cu - brute-force descriptor matcher (dimb_nn_match), replacing KorniaMatcher._match_pairs
// (reference matchers/kornia_matcher.py:27-54 -> kornia.feature.DescriptorMatcher nn/mnn/snn/smnn).
//
// The reference materialises the full n0 x n1 distance matrix (torch.cdist, 268 MB at 8192^2) and runs
// min / topk over it.  Here the distance tile never leaves the SM: the tensor-core GEMM of gemm.cuh
// produces a 128 x 128 tile of dot products in TMEM and its epilogue turns it into distances
// (|a|^2 + |b|^2 - 2ab, clamped, sqrt) and a per-row running (best, second best, argbest) over each
// 32-column chunk; a small merge kernel reduces the chunk partials.  The column statistics needed by the
// mutual / symmetric modes are the same kernel with the operands swapped.
#include <algorithm>
#include <vector>

#include "gemm.cuh"

namespace {

// K = D is small (256): a 128 x 256 output tile halves the A re-reads per FLOP compared with 128 x 128 (the kernel is bound by
// L2 -> shared-memory operand traffic, not by the tensor pipe)
constexpr int kNnBN = 256;

struct EpiNNTop2 : EpiBase {
  static constexpr bool kUsesScratch = false;
  static constexpr int kEpiWarps = 8;  // sqrt + running top-2 per element: the epilogue out-lasts the K = 256 MMAs of a tile
  const float *na, *nb;  // squared norms of A rows / B rows
  float *pd1, *pd2;      // [rows][chunks] best / second best distance of each 32-column chunk
  int* pi1;              // [rows][chunks] argbest
  int n_rows, n_cols, chunks;
  __device__ void operator()(const TileCoord& tc, int r, int n, float (&v)[32], float*) const {
    const int row = tc.m0 + r;
    if (row >= n_rows) return;
    const float a2 = na[row];
    float d1 = INFINITY, d2 = INFINITY;
    int i1 = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (n + j < n_cols) {
        const float d = sqrtf(fmaxf(a2 + nb[n + j] - 2.f * v[j], 0.f));
        if (d < d1) {
          d2 = d1;
          d1 = d;
          i1 = n + j;
        } else if (d < d2) {
          d2 = d;
        }
      }
    }
    const size_t o = static_cast<size_t>(row) * chunks + (n >> 5);
    pd1[o] = d1;
    pd2[o] = d2;
    pi1[o] = i1;
  }
};

// (D,n) fp32 -> [n_pad][D] fp16 hi/lo + squared norms; block (32,8) transposing 32x32 tiles
__global__ void nn_prep_kernel(const float* __restrict__ d, int D, int n, __half* __restrict__ hi, __half* __restrict__ lo,
                               float* __restrict__ norm) {
  __shared__ float tile[32][33];
  const int t0 = blockIdx.x * 32, tx = threadIdx.x, ty = threadIdx.y;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c0 = 0; c0 < D; c0 += 32) {
    for (int k = ty; k < 32; k += 8) tile[k][tx] = (t0 + tx < n && c0 + k < D) ? d[static_cast<size_t>(c0 + k) * n + t0 + tx] : 0.f;
    __syncthreads();
    int q = 0;
    for (int k = ty; k < 32; k += 8, ++q) {
      const float v = tile[tx][k];  // token t0+k, channel c0+tx
      if (t0 + k < n && c0 + tx < D) {
        __half h, l;
        split_f32(v, h, l);
        hi[static_cast<size_t>(t0 + k) * D + c0 + tx] = h;
        if (lo) lo[static_cast<size_t>(t0 + k) * D + c0 + tx] = l;
      }
      float sq = v * v;
#pragma unroll
      for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      acc[q] += sq;
    }
    __syncthreads();
  }
  int q = 0;
  for (int k = ty; k < 32; k += 8, ++q)
    if (tx == 0 && t0 + k < n) norm[t0 + k] = acc[q];
}

// warp per row: merge chunk partials -> best, second, arg (first index wins ties)
__global__ void nn_merge_kernel(const float* __restrict__ pd1, const float* __restrict__ pd2, const int* __restrict__ pi1, int rows,
                                int chunks, float* __restrict__ d1, float* __restrict__ d2, int* __restrict__ i1) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  float b1 = INFINITY, b2 = INFINITY;
  int bi = 0x7fffffff;
  for (int c = lane; c < chunks; c += 32) {
    const size_t o = static_cast<size_t>(row) * chunks + c;
    const float x1 = pd1[o], x2 = pd2[o];
    const int xi = pi1[o];
    if (x1 < b1 || (x1 == b1 && xi < bi)) {
      b2 = fminf(b1, x2);
      b1 = x1;
      bi = xi;
    } else {
      b2 = fminf(b2, x1);
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float x1 = __shfl_xor_sync(0xffffffffu, b1, o), x2 = __shfl_xor_sync(0xffffffffu, b2, o);
    const int xi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (x1 < b1 || (x1 == b1 && xi < bi)) {
      b2 = fminf(b1, x2);
      b1 = x1;
      bi = xi;
    } else {
      b2 = fminf(b2, x1);
    }
  }
  if (lane == 0) {
    d1[row] = b1;
    d2[row] = b2;
    i1[row] = bi;
  }
}

// one CTA: apply the kornia mode logic and compact in ascending row order
__global__ void __launch_bounds__(1024)
nn_select_kernel(int mode, float th, int n0, int n1, const float* __restrict__ fd1, const float* __restrict__ fd2,
                 const int* __restrict__ fi1, const float* __restrict__ bd1, const float* __restrict__ bd2, const int* __restrict__ bi1,
                 long long* __restrict__ idx, float* __restrict__ dist, int* __restrict__ count, int cap) {
  __shared__ int wsum[32];
  __shared__ int s_base;
  const int t = threadIdx.x;
  if (t == 0) s_base = 0;
  __syncthreads();
  const int ms = min(n0, n1);
  const bool swapped = (mode == DIMB_NN_MNN) && (n0 > n1);  // kornia match_mnn iterates the smaller side
  const int iters = (mode == DIMB_NN_MNN) ? ms : n0;
  for (int base = 0; base < iters; base += blockDim.x) {
    const int i = base + t;
    bool valid = false;
    long long a = 0, b = 0;
    float dv = 0.f;
    if (i < iters) {
      if (mode == DIMB_NN_NN) {
        valid = true, a = i, b = fi1[i], dv = fd1[i];
      } else if (mode == DIMB_NN_MNN) {
        if (!swapped) {
          const int j = fi1[i];
          valid = bi1[j] == i, a = i, b = j, dv = fd1[i];
        } else {
          const int j = bi1[i];  // i indexes desc2
          valid = fi1[j] == i, a = j, b = i, dv = bd1[i];
        }
      } else if (mode == DIMB_NN_SNN) {
        const float ratio = fd1[i] / fd2[i];
        valid = ratio <= th, a = i, b = fi1[i], dv = ratio;
      } else {  // SMNN
        const float rf = fd1[i] / fd2[i];
        const int j = fi1[i];
        const float rb = bd1[j] / bd2[j];
        valid = (rf <= th) && (rb <= th) && (bi1[j] == i);
        a = i, b = j, dv = fmaxf(rf, rb);
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    if ((t & 31) == 0) wsum[t >> 5] = __popc(bal);
    __syncthreads();
    int before = s_base;
    for (int wv = 0; wv < (t >> 5); ++wv) before += wsum[wv];
    before += __popc(bal & ((1u << (t & 31)) - 1u));
    if (valid && before < cap) {
      idx[2 * before] = a;
      idx[2 * before + 1] = b;
      dist[before] = dv;
    }
    __syncthreads();
    if (t == 0) {
      int tot = 0;
      for (int wv = 0; wv < 32; ++wv) tot += wsum[wv];
      s_base += tot;
    }
    __syncthreads();
  }
  if (t == 0) *count = s_base;
}

struct NNSide {
  __half *hi, *lo;
  float* norm;
  CUtensorMap mA[2], mB[2];
};

int nn_rowtop2(dimb_ctx* ctx, cudaStream_t st, const NNSide& A, int na, const NNSide& B, int nb, int D, float* pd1, float* pd2, int* pi1,
               float* d1, float* d2, int* i1) {
  EpiNNTop2 e;
  e.na = A.norm;
  e.nb = B.norm;
  e.pd1 = pd1;
  e.pd2 = pd2;
  e.pi1 = pi1;
  e.n_rows = na;
  e.n_cols = nb;
  e.chunks = round_up(nb, kNnBN) / 32;
  TcOperands ops;
  ops.Ah = A.mA[0];
  ops.Al = A.mA[1];
  ops.Bh = B.mB[0];
  ops.Bl = B.mB[1];
  GemmArgs g{};
  g.num_kb = D / 64;
  g.M = na;
  g.N = nb;
  g.Ah = A.hi;
  g.Al = A.lo;
  g.Bh = B.hi;
  g.Bl = B.lo;
  g.lda = D;
  g.ldb = D;
  DIMB_TRY((launch_gemm<kNnBN, false>(ctx, st, ops, g, e, ceil_div(na, kTileM), round_up(nb, kNnBN), "nn.top2_gemm")));
  ProfScope prof(ctx, st, "nn.merge");
  nn_merge_kernel<<<ceil_div(na * 32, 256), 256, 0, st>>>(pd1, pd2, pi1, na, e.chunks, d1, d2, i1);
  DIMB_LAUNCH_CHECK(ctx);
  return DIMB_OK;
}

}  // namespace

extern "C" int dimb_nn_match(dimb_ctx* ctx, const float* d0, int n0, const float* d1, int n1, int D, int mode, float th, int64_t* idx,
                             float* dist, int* n, int cap) {
  if (!ctx || !idx || !dist || !n || n0 < 0 || n1 < 0 || D < 64 || D % 64 != 0 || mode < 0 || mode > 3 || cap < 1) return DIMB_ERR_ARG;
  *n = 0;
  // kornia: empty inputs / fewer than two candidates for the ratio tests -> no match
  if (n0 == 0 || n1 == 0) return DIMB_OK;
  if (mode == DIMB_NN_SNN && n1 < 2) return DIMB_OK;
  if (mode == DIMB_NN_SMNN && (n0 < 2 || n1 < 2)) return DIMB_OK;
  DIMB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = 0;
  const bool exact = ctx->precision == DIMB_PRECISION_EXACT;
  const int p0 = round_up(n0, kNnBN), p1 = round_up(n1, kNnBN);
  // scratch lives in grow-only context slots: no cudaMalloc / cudaFree in steady state
  int slot = 0;
  auto alloc = [&](void** p, size_t bytes) -> int { return dimb_scratch(ctx, slot++, bytes, p); };
  auto release = [&]() {};
#define NN_TRY(expr)       \
  do {                     \
    int _r = (expr);       \
    if (_r != DIMB_OK) {   \
      release();           \
      return _r;           \
    }                      \
  } while (0)
  float *raw0, *raw1;
  NNSide s0{}, s1{};
  NN_TRY(alloc(reinterpret_cast<void**>(&raw0), static_cast<size_t>(D) * n0 * sizeof(float)));
  NN_TRY(alloc(reinterpret_cast<void**>(&raw1), static_cast<size_t>(D) * n1 * sizeof(float)));
  for (auto* sp : {&s0, &s1}) {
    const int pn = sp == &s0 ? p0 : p1;
    NN_TRY(alloc(reinterpret_cast<void**>(&sp->hi), static_cast<size_t>(pn) * D * sizeof(__half)));
    NN_TRY(alloc(reinterpret_cast<void**>(&sp->lo), static_cast<size_t>(pn) * D * sizeof(__half)));
    NN_TRY(alloc(reinterpret_cast<void**>(&sp->norm), static_cast<size_t>(pn) * sizeof(float)));
    NN_TRY(dimb_tmap_2d(ctx, &sp->mA[0], sp->hi, pn, D, D, kTileM));
    NN_TRY(dimb_tmap_2d(ctx, &sp->mA[1], sp->lo, pn, D, D, kTileM));
    NN_TRY(dimb_tmap_2d(ctx, &sp->mB[0], sp->hi, pn, D, D, kNnBN));  // as B operand: boxes of kNnBN rows
    NN_TRY(dimb_tmap_2d(ctx, &sp->mB[1], sp->lo, pn, D, D, kNnBN));
  }
  cudaError_t ce = cudaMemcpyAsync(raw0, d0, static_cast<size_t>(D) * n0 * sizeof(float), cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(raw1, d1, static_cast<size_t>(D) * n1 * sizeof(float), cudaMemcpyHostToDevice, st);
  if (ce != cudaSuccess) {
    dimb_set_error(ctx, std::string("dimb_nn_match: H2D copy failed: ") + cudaGetErrorString(ce));
    release();
    return DIMB_ERR_CUDA;
  }
  {
    ProfScope prof(ctx, st, "nn.prep");
    nn_prep_kernel<<<ceil_div(n0, 32), dim3(32, 8), 0, st>>>(raw0, D, n0, s0.hi, exact ? s0.lo : nullptr, s0.norm);
    ctx->launches++;
    nn_prep_kernel<<<ceil_div(n1, 32), dim3(32, 8), 0, st>>>(raw1, D, n1, s1.hi, exact ? s1.lo : nullptr, s1.norm);
    ctx->launches++;
  }
  const size_t ch = static_cast<size_t>(std::max(p0, p1)) / 32;
  float *pd1, *pd2, *fd1, *fd2, *bd1, *bd2, *o_dist;
  int *pi1, *fi1, *bi1, *o_n;
  long long* o_idx;
  NN_TRY(alloc(reinterpret_cast<void**>(&pd1), std::max(p0, p1) * ch * sizeof(float)));
  NN_TRY(alloc(reinterpret_cast<void**>(&pd2), std::max(p0, p1) * ch * sizeof(float)));
  NN_TRY(alloc(reinterpret_cast<void**>(&pi1), std::max(p0, p1) * ch * sizeof(int)));
  NN_TRY(alloc(reinterpret_cast<void**>(&fd1), p0 * sizeof(float)));
  NN_TRY(alloc(reinterpret_cast<void**>(&fd2), p0 * sizeof(float)));
  NN_TRY(alloc(reinterpret_cast<void**>(&fi1), p0 * sizeof(int)));
  NN_TRY(alloc(reinterpret_cast<void**>(&bd1), p1 * sizeof(float)));
  NN_TRY(alloc(reinterpret_cast<void**>(&bd2), p1 * sizeof(float)));
  NN_TRY(alloc(reinterpret_cast<void**>(&bi1), p1 * sizeof(int)));
  NN_TRY(alloc(reinterpret_cast<void**>(&o_idx), static_cast<size_t>(cap) * 2 * sizeof(long long)));
  NN_TRY(alloc(reinterpret_cast<void**>(&o_dist), static_cast<size_t>(cap) * sizeof(float)));
  NN_TRY(alloc(reinterpret_cast<void**>(&o_n), sizeof(int)));
  NN_TRY(nn_rowtop2(ctx, st, s0, n0, s1, n1, D, pd1, pd2, pi1, fd1, fd2, fi1));
  if (mode == DIMB_NN_MNN || mode == DIMB_NN_SMNN) NN_TRY(nn_rowtop2(ctx, st, s1, n1, s0, n0, D, pd1, pd2, pi1, bd1, bd2, bi1));
  {
    ProfScope prof(ctx, st, "nn.select");
    nn_select_kernel<<<1, 1024, 0, st>>>(mode, th, n0, n1, fd1, fd2, fi1, bd1, bd2, bi1, o_idx, o_dist, o_n, cap);
    ctx->launches++;
  }
  int cnt = 0;
  ce = cudaMemcpyAsync(&cnt, o_n, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
  if (ce == cudaSuccess && cnt > 0 && cnt <= cap) {
    ce = cudaMemcpy(idx, o_idx, static_cast<size_t>(cnt) * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess) ce = cudaMemcpy(dist, o_dist, static_cast<size_t>(cnt) * sizeof(float), cudaMemcpyDeviceToHost);
  }
  release();
  if (ce != cudaSuccess) {
    dimb_set_error(ctx, std::string("dimb_nn_match: ") + cudaGetErrorString(ce));
    return DIMB_ERR_CUDA;
  }
  *n = cnt;
  if (cnt > cap) {
    dimb_set_error(ctx, "dimb_nn_match: more matches than cap");
    return DIMB_ERR_CAPACITY;
  }
  return DIMB_OK;
#undef NN_TRY
}
