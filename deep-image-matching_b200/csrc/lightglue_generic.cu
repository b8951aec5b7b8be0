// lightglue_generic.cu - LightGlue for shapes other than (descriptor_dim 256, 4 heads x 64), i.e. the LighterGlue
// checkpoint the reference ships (thirdparty/accelerated_features/modules/lighterglue.py:12-27: descriptor_dim 96,
// one head, 6 layers, input_dim 64; matcher plugin src/deep_image_matching/matchers/lighterglue.py:78-262).
// Same algorithm as lightglue.cu (thirdparty/LightGlue/lightglue/lightglue.py:24-610) in plain fp32 on the CUDA cores:
// the tensor-core kernels of lightglue.cu are specialised for head dim 64 / model dim 256 (TMEM and shared-memory budgets of
// the attention kernel), this file trades speed for generality.  Control flow (early stop, pruning) is decided on the host
// from per-token confidences copied back once per layer - exactly the synchronisation points of the reference
// (lightglue.py:499,503).  One pair at a time.
#include <algorithm>
#include <memory>
#include <cmath>
#include <cstring>
#include <numeric>
#include <vector>

#include "common.cuh"
#include "lightglue_generic.cuh"

namespace {

// ------------------------------------------------------------------ kernels
// C[m][n] = (sum_k A[m*lda + k] * W[n*ldw + k] + bias[n]) * scale (+ resid[m*ldr + n]); 64 x 64 tile, 256 threads, 4 x 4 outputs
// per thread, K streamed through shared memory 16 at a time (k ascending per output: deterministic summation order).
__global__ void __launch_bounds__(256) gx_linear_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                                                        const float* __restrict__ bias, float* __restrict__ C, int ldc, int M, int N, int K,
                                                        float scale, const float* __restrict__ resid, int ldr) {
  __shared__ float sa[16][64 + 4], sb[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4, m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int r = e >> 4, k = e & 15;
      sa[k][r] = (m0 + r < M && k0 + k < K) ? A[static_cast<size_t>(m0 + r) * lda + k0 + k] : 0.f;
      sb[k][r] = (n0 + r < N && k0 + k < K) ? W[static_cast<size_t>(n0 + r) * ldw + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&sa[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sb[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = (acc[i][j] + (bias ? bias[n] : 0.f)) * scale;
      if (resid) v += resid[static_cast<size_t>(m) * ldr + n];
      C[static_cast<size_t>(m) * ldc + n] = v;
    }
  }
}

// keypoint normalisation (lightglue.py:24-34) + Fourier encoding (:57-70): enc [2][N][hd] = cos / sin, each frequency twice
__global__ void gx_posenc_kernel(const float* __restrict__ kpts, int n, float size0, float size1, const float* __restrict__ Wr /*[hd/2][2]*/,
                                 int hd, float* __restrict__ enc, int np) {
  const int i = blockIdx.x, f = threadIdx.x;
  if (i >= n || f >= hd / 2) return;
  const float sc = fmaxf(size0, size1) / 2.f;
  const float x = (kpts[2 * i] - size0 / 2.f) / sc, y = (kpts[2 * i + 1] - size1 / 2.f) / sc;
  const float pr = x * Wr[2 * f] + y * Wr[2 * f + 1];
  const float c = cosf(pr), s = sinf(pr);
  float* e0 = enc + static_cast<size_t>(i) * hd + 2 * f;
  e0[0] = c, e0[1] = c;
  e0[static_cast<size_t>(np) * hd] = s, e0[static_cast<size_t>(np) * hd + 1] = s;
}

// Wqkv output [N][3d] interleaved as (h, hd, 3) (lightglue.py:153-154) -> q, k (rotary applied, :47-54), v, each [N][d]
__global__ void gx_qkv_rotary_kernel(const float* __restrict__ qkv, int n, int d, int hd, const float* __restrict__ enc, int np,
                                     float* __restrict__ q, float* __restrict__ k, float* __restrict__ v) {
  const int i = blockIdx.x, c = threadIdx.x * 2;  // channel pair (c, c+1) of the model dimension
  if (i >= n || c >= d) return;
  const float* r = qkv + static_cast<size_t>(i) * 3 * d;
  const int dd = c % hd;  // position inside the head: the encoding is shared by the heads
  const float c0 = enc[static_cast<size_t>(i) * hd + dd], c1 = enc[static_cast<size_t>(i) * hd + dd + 1];
  const float s0 = enc[(static_cast<size_t>(np) + i) * hd + dd], s1 = enc[(static_cast<size_t>(np) + i) * hd + dd + 1];
  const float q0 = r[c * 3], q1 = r[(c + 1) * 3], k0 = r[c * 3 + 1], k1 = r[(c + 1) * 3 + 1];
  const size_t o = static_cast<size_t>(i) * d + c;
  q[o] = q0 * c0 + (-q1) * s0;  // rotate_half: (x0, x1) -> (-x1, x0)
  q[o + 1] = q1 * c1 + q0 * s1;
  k[o] = k0 * c0 + (-k1) * s0;
  k[o + 1] = k1 * c1 + k0 * s1;
  v[o] = r[c * 3 + 2];
  v[o + 1] = r[(c + 1) * 3 + 2];
}

// softmax(q k^T * hd^-0.5) v, fp32.  CTA = 8 warps = 8 queries of one head sharing 32-key tiles of K and V in shared memory;
// lane = key of the tile for the logits, lane = channel (mod 32) for the output.  Online softmax.  HDP = hd rounded up to 32.
template <int HDP>
__global__ void __launch_bounds__(256) gx_attention_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                                           int nq, int nk, int d, int hd, float* __restrict__ out, int ldo) {
  __shared__ float sk[32][HDP + 1], sv[32][HDP], sq[8][HDP];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, head = blockIdx.y, qi = blockIdx.x * 8 + w;
  const int co = head * hd;
  for (int c = lane; c < HDP; c += 32) sq[w][c] = (qi < nq && c < hd) ? q[static_cast<size_t>(qi) * d + co + c] : 0.f;
  const float scale = 1.f / sqrtf(static_cast<float>(hd));
  float mx = -INFINITY, l = 0.f, o[HDP / 32];
#pragma unroll
  for (int j = 0; j < HDP / 32; ++j) o[j] = 0.f;
  for (int k0 = 0; k0 < nk; k0 += 32) {
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * HDP; e += 256) {
      const int r = e / HDP, c = e - r * HDP;
      const bool ok = k0 + r < nk && c < hd;
      sk[r][c] = ok ? k[static_cast<size_t>(k0 + r) * d + co + c] : 0.f;
      sv[r][c] = ok ? v[static_cast<size_t>(k0 + r) * d + co + c] : 0.f;
    }
    __syncthreads();
    float s = 0.f;
#pragma unroll 8
    for (int c = 0; c < HDP; ++c) s = fmaf(sq[w][c], sk[lane][c], s);
    s = (k0 + lane < nk) ? s * scale : -INFINITY;
    float tm = s;
#pragma unroll
    for (int of = 16; of; of >>= 1) tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, of));
    const float mn = fmaxf(mx, tm), corr = expf(mx - mn), p = expf(s - mn);
    float ps = p;
#pragma unroll
    for (int of = 16; of; of >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, of);
    l = l * corr + ps;
#pragma unroll
    for (int j = 0; j < HDP / 32; ++j) o[j] *= corr;
    for (int r = 0; r < 32; ++r) {
      const float pr = __shfl_sync(0xffffffffu, p, r);
#pragma unroll
      for (int j = 0; j < HDP / 32; ++j) o[j] = fmaf(pr, sv[r][lane + 32 * j], o[j]);
    }
    mx = mn;
  }
  if (qi >= nq) return;
#pragma unroll
  for (int j = 0; j < HDP / 32; ++j) {
    const int c = lane + 32 * j;
    if (c < hd) out[static_cast<size_t>(qi) * ldo + co + c] = nk > 0 ? o[j] / l : 0.f;  // empty key set -> zeros (lightglue.py:103-104)
  }
}

// y = gelu(layer_norm(x)) over the n features of a row, eps 1e-5, exact (erf) GELU; warp per row
__global__ void gx_ln_gelu_kernel(const float* __restrict__ x, int rows, int n, const float* __restrict__ g, const float* __restrict__ b,
                                  float* __restrict__ y) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* r = x + static_cast<size_t>(row) * n;
  float s = 0.f;
  for (int c = lane; c < n; c += 32) s += r[c];
#pragma unroll
  for (int of = 16; of; of >>= 1) s += __shfl_xor_sync(0xffffffffu, s, of);
  const float mean = s / n;
  float vs = 0.f;
  for (int c = lane; c < n; c += 32) {
    const float dlt = r[c] - mean;
    vs = fmaf(dlt, dlt, vs);
  }
#pragma unroll
  for (int of = 16; of; of >>= 1) vs += __shfl_xor_sync(0xffffffffu, vs, of);
  const float inv = rsqrtf(vs / n + 1e-5f);
  for (int c = lane; c < n; c += 32) {
    const float t = (r[c] - mean) * inv * g[c] + b[c];
    y[static_cast<size_t>(row) * n + c] = 0.5f * t * (1.f + erff(t * 0.70710678118654752440f));
  }
}

// z[row] = x[row] . w + b (token confidence / matchability logits); warp per row
__global__ void gx_rowdot_kernel(const float* __restrict__ x, int ldx, int rows, int n, const float* __restrict__ w, const float* __restrict__ b,
                                 float* __restrict__ z) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  float s = 0.f;
  for (int c = lane; c < n; c += 32) s = fmaf(x[static_cast<size_t>(row) * ldx + c], w[c], s);
#pragma unroll
  for (int of = 16; of; of >>= 1) s += __shfl_xor_sync(0xffffffffu, s, of);
  if (lane == 0) z[row] = s + b[0];
}

// pruning gather: dst row i = src row idx[i] for the state (stride ld, d used) and both halves of the encoding
__global__ void gx_gather_kernel(const float* __restrict__ xs, float* __restrict__ xd, int ld, int d, const float* __restrict__ es,
                                 float* __restrict__ ed, int hd, int np, const int* __restrict__ idx, int n) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const int s = idx[i];
  for (int c = threadIdx.x; c < d; c += blockDim.x) xd[static_cast<size_t>(i) * ld + c] = xs[static_cast<size_t>(s) * ld + c];
  for (int c = threadIdx.x; c < hd; c += blockDim.x) {
    ed[static_cast<size_t>(i) * hd + c] = es[static_cast<size_t>(s) * hd + c];
    ed[(static_cast<size_t>(np) + i) * hd + c] = es[(static_cast<size_t>(np) + s) * hd + c];
  }
}

// log-sum-exp of the rows (dir 0) or columns (dir 1) of sim [m][n] (row stride ld); warp per row / column
__global__ void gx_lse_kernel(const float* __restrict__ sim, int ld, int m, int n, int dir, float* __restrict__ lse) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int cnt = dir == 0 ? m : n, len = dir == 0 ? n : m;
  if (i >= cnt) return;
  float mx = -INFINITY;
  for (int j = lane; j < len; j += 32) mx = fmaxf(mx, dir == 0 ? sim[static_cast<size_t>(i) * ld + j] : sim[static_cast<size_t>(j) * ld + i]);
#pragma unroll
  for (int of = 16; of; of >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, of));
  float s = 0.f;
  for (int j = lane; j < len; j += 32) s += expf((dir == 0 ? sim[static_cast<size_t>(i) * ld + j] : sim[static_cast<size_t>(j) * ld + i]) - mx);
#pragma unroll
  for (int of = 16; of; of >>= 1) s += __shfl_xor_sync(0xffffffffu, s, of);
  if (lane == 0) lse[i] = mx + logf(s);
}

__device__ __forceinline__ float log_sigmoid(float z) { return fminf(z, 0.f) - log1pf(expf(-fabsf(z))); }

// row (dir 0) / column (dir 1) maximum and first argmax of scores = (sim - rlse) + (sim - clse) + logsig(z0) + logsig(z1)
// in the association of the reference (lightglue.py:246-256: scores0 + scores1 + certainties)
__global__ void gx_argmax_kernel(const float* __restrict__ sim, int ld, int m, int n, const float* __restrict__ rlse,
                                 const float* __restrict__ clse, const float* __restrict__ z0, const float* __restrict__ z1, int dir,
                                 float* __restrict__ best, int* __restrict__ arg) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int cnt = dir == 0 ? m : n, len = dir == 0 ? n : m;
  if (i >= cnt) return;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = lane; j < len; j += 32) {
    const int r = dir == 0 ? i : j, c = dir == 0 ? j : i;
    const float sv = sim[static_cast<size_t>(r) * ld + c];
    const float val = ((sv - rlse[r]) + (sv - clse[c])) + (log_sigmoid(z0[r]) + log_sigmoid(z1[c]));
    if (val > bv) bv = val, bi = j;
  }
#pragma unroll
  for (int of = 16; of; of >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, of);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, of);
    if (ov > bv || (ov == bv && oi < bi)) bv = ov, bi = oi;
  }
  if (lane == 0) best[i] = bv, arg[i] = bi;
}

struct Lin {
  float *w = nullptr, *b = nullptr;
  int n = 0, k = 0;
};
struct Block {
  Lin qkv, to_qk, to_v, out, ffn0, ffn3;
  float *ln_g = nullptr, *ln_b = nullptr;
};

}  // namespace

struct dimb_lgx {
  dimb_ctx* ctx;
  std::vector<void*> mem;
  dimb_lg_conf conf;
  int d, h, hd, din, L, NP;
  float* Wr;
  Lin input_proj;
  std::vector<Block> self_, cross_;
  std::vector<Lin> matchab, final_proj, token;
  // per-side state (2 sides): cat [NP][2d] = [x | message], encodings [2][NP][hd], ping-pong copies for the pruning gather
  float *cat[2][2], *enc[2][2];
  float *desc_in, *kpts, *qkv, *q[2], *k[2], *v[2], *hid, *hid2, *md[2], *zt[2], *sim, *rlse, *clse, *best0, *best1;
  int *arg0, *arg1, *idx;
};

namespace {

int linear(dimb_lgx* g, cudaStream_t st, const float* A, int lda, const Lin& l, float* C, int ldc, int M, float scale = 1.f,
           const float* resid = nullptr, int ldr = 0) {
  if (M <= 0) return DIMB_OK;
  dim3 grid(ceil_div(l.n, 64), ceil_div(M, 64));
  gx_linear_kernel<<<grid, 256, 0, st>>>(A, lda, l.w, l.k, l.b, C, ldc, M, l.n, l.k, scale, resid, ldr);
  DIMB_LAUNCH_CHECK(g->ctx);
  return DIMB_OK;
}

int attention(dimb_lgx* g, cudaStream_t st, const float* q, const float* k, const float* v, int nq, int nk, float* out, int ldo) {
  if (nq <= 0) return DIMB_OK;
  dim3 grid(ceil_div(nq, 8), g->h);
  if (g->hd <= 32)
    gx_attention_kernel<32><<<grid, 256, 0, st>>>(q, k, v, nq, nk, g->d, g->hd, out, ldo);
  else if (g->hd <= 64)
    gx_attention_kernel<64><<<grid, 256, 0, st>>>(q, k, v, nq, nk, g->d, g->hd, out, ldo);
  else if (g->hd <= 96)
    gx_attention_kernel<96><<<grid, 256, 0, st>>>(q, k, v, nq, nk, g->d, g->hd, out, ldo);
  else
    gx_attention_kernel<128><<<grid, 256, 0, st>>>(q, k, v, nq, nk, g->d, g->hd, out, ldo);
  DIMB_LAUNCH_CHECK(g->ctx);
  return DIMB_OK;
}

// x <- x + ffn3(gelu(ln(ffn0([x | msg]))))  on cat [n][2d]   (lightglue.py:135-143 / 176-184)
int ffn(dimb_lgx* g, cudaStream_t st, float* cat, int n, const Block& b) {
  if (n <= 0) return DIMB_OK;
  const int d = g->d;
  DIMB_TRY(linear(g, st, cat, 2 * d, b.ffn0, g->hid, 2 * d, n));
  gx_ln_gelu_kernel<<<ceil_div(n * 32, 256), 256, 0, st>>>(g->hid, n, 2 * d, b.ln_g, b.ln_b, g->hid2);
  DIMB_LAUNCH_CHECK(g->ctx);
  return linear(g, st, g->hid2, 2 * d, b.ffn3, cat, 2 * d, n, 1.f, cat, 2 * d);
}

float conf_threshold(int i, int L) {  // lightglue.py:581-584
  return static_cast<float>(std::min(std::max(0.8 + 0.1 * std::exp(-4.0 * i / L), 0.0), 1.0));
}

}  // namespace

int lgx_create(dimb_ctx* ctx, const float* weights, size_t n_floats, const dimb_lg_conf* cf, dimb_lgx** out) {
  *out = nullptr;
  const int d = cf->descriptor_dim, h = cf->num_heads, L = cf->n_layers, din = cf->input_dim;
  if (d < 2 || h < 1 || d % h != 0 || (d / h) % 2 != 0 || d / h > 128 || d > 1024 || L < 1 || din < 1 || cf->max_kpts < 1) {
    dimb_set_error(ctx, "dimb_lg_create: unsupported LightGlue shape (head dim must be even and <= 128)");
    return DIMB_ERR_UNSUPPORTED;
  }
  const int hd = d / h;
  size_t need = static_cast<size_t>(hd / 2) * 2;
  if (din != d) need += static_cast<size_t>(d) * din + d;
  const size_t per_layer = (3 * d * d + 3 * d) + (d * d + d) + (4 * d * d + 2 * d) + 4 * d + (2 * d * d + d)  // self
                           + 3 * (static_cast<size_t>(d) * d + d) + (4 * d * d + 2 * d) + 4 * d + (2 * d * d + d);  // cross
  need += per_layer * L + static_cast<size_t>(L) * (d + 1 + d * d + d) + static_cast<size_t>(L - 1) * (d + 1);
  if (n_floats != need) {
    dimb_set_error(ctx, "dimb_lg_create: weight blob has " + std::to_string(n_floats) + " floats, expected " + std::to_string(need));
    return DIMB_ERR_ARG;
  }
  dimb_lgx* g = new dimb_lgx();
  g->ctx = ctx;
  std::unique_ptr<dimb_lgx, void (*)(dimb_lgx*)> guard(g, lgx_destroy);  // a failed create releases what it built
  OwnerScope own(ctx, &g->mem);
  g->conf = *cf;
  g->d = d, g->h = h, g->hd = hd, g->din = din, g->L = L;
  g->NP = cf->max_kpts;
  const float* p = weights;
  auto up = [&](float** dst, size_t n) -> int {
    DIMB_TRY(dimb_alloc_t(ctx, dst, n, false));
    DIMB_CUDA_OK(ctx, cudaMemcpy(*dst, p, n * sizeof(float), cudaMemcpyHostToDevice));
    p += n;
    return static_cast<int>(DIMB_OK);
  };
  auto lin = [&](Lin& l, int n, int k) -> int {
    l.n = n, l.k = k;
    DIMB_TRY(up(&l.w, static_cast<size_t>(n) * k));
    return up(&l.b, n);
  };
  DIMB_TRY(up(&g->Wr, static_cast<size_t>(hd / 2) * 2));
  if (din != d) DIMB_TRY(lin(g->input_proj, d, din));
  g->self_.resize(L), g->cross_.resize(L);
  for (int i = 0; i < L; ++i) {
    Block& s = g->self_[i];
    DIMB_TRY(lin(s.qkv, 3 * d, d));
    DIMB_TRY(lin(s.out, d, d));
    DIMB_TRY(lin(s.ffn0, 2 * d, 2 * d));
    DIMB_TRY(up(&s.ln_g, 2 * d));
    DIMB_TRY(up(&s.ln_b, 2 * d));
    DIMB_TRY(lin(s.ffn3, d, 2 * d));
    Block& c = g->cross_[i];
    DIMB_TRY(lin(c.to_qk, d, d));
    DIMB_TRY(lin(c.to_v, d, d));
    DIMB_TRY(lin(c.out, d, d));
    DIMB_TRY(lin(c.ffn0, 2 * d, 2 * d));
    DIMB_TRY(up(&c.ln_g, 2 * d));
    DIMB_TRY(up(&c.ln_b, 2 * d));
    DIMB_TRY(lin(c.ffn3, d, 2 * d));
  }
  g->matchab.resize(L), g->final_proj.resize(L), g->token.resize(std::max(L - 1, 0));
  for (int i = 0; i < L; ++i) {
    DIMB_TRY(lin(g->matchab[i], 1, d));
    DIMB_TRY(lin(g->final_proj[i], d, d));
  }
  for (int i = 0; i < L - 1; ++i) DIMB_TRY(lin(g->token[i], 1, d));
  const size_t NP = g->NP;
  for (int s = 0; s < 2; ++s)
    for (int b = 0; b < 2; ++b) {
      DIMB_TRY(dimb_alloc_t(ctx, &g->cat[s][b], NP * 2 * d));
      DIMB_TRY(dimb_alloc_t(ctx, &g->enc[s][b], 2 * NP * hd));
    }
  DIMB_TRY(dimb_alloc_t(ctx, &g->desc_in, NP * std::max(din, d)));
  DIMB_TRY(dimb_alloc_t(ctx, &g->kpts, NP * 2));
  DIMB_TRY(dimb_alloc_t(ctx, &g->qkv, NP * 3 * d));
  for (int s = 0; s < 2; ++s) {
    DIMB_TRY(dimb_alloc_t(ctx, &g->q[s], NP * d));
    DIMB_TRY(dimb_alloc_t(ctx, &g->k[s], NP * d));
    DIMB_TRY(dimb_alloc_t(ctx, &g->v[s], NP * d));
    DIMB_TRY(dimb_alloc_t(ctx, &g->md[s], NP * d));
    DIMB_TRY(dimb_alloc_t(ctx, &g->zt[s], NP * 2));
  }
  DIMB_TRY(dimb_alloc_t(ctx, &g->hid, NP * 2 * d));
  DIMB_TRY(dimb_alloc_t(ctx, &g->hid2, NP * 2 * d));
  DIMB_TRY(dimb_alloc_t(ctx, &g->sim, NP * NP));
  DIMB_TRY(dimb_alloc_t(ctx, &g->rlse, NP));
  DIMB_TRY(dimb_alloc_t(ctx, &g->clse, NP));
  DIMB_TRY(dimb_alloc_t(ctx, &g->best0, NP));
  DIMB_TRY(dimb_alloc_t(ctx, &g->best1, NP));
  DIMB_TRY(dimb_alloc_t(ctx, &g->arg0, NP));
  DIMB_TRY(dimb_alloc_t(ctx, &g->arg1, NP));
  DIMB_TRY(dimb_alloc_t(ctx, &g->idx, NP));
  *out = guard.release();
  return DIMB_OK;
}

void lgx_destroy(dimb_lgx* g) {
  if (!g) return;
  dimb_release(g->ctx, g->mem);
  delete g;
}

// one pair; outputs as dimb_lg_match
static int lgx_match_pair(dimb_lgx* g, const dimb_feats& f0, const dimb_feats& f1, int64_t* matches, float* mscores, int* n_matches,
                          int* stop_layer, int cap) {
  dimb_ctx* ctx = g->ctx;
  cudaStream_t st = 0;
  const int d = g->d, hd = g->hd, L = g->L, din = g->din, NP = g->NP;
  const dimb_lg_conf& cf = g->conf;
  const dimb_feats* F[2] = {&f0, &f1};
  int n[2] = {f0.n, f1.n};
  *n_matches = 0;
  if (n[0] > NP || n[1] > NP) {
    dimb_set_error(ctx, "dimb_lg_match: more keypoints than max_kpts");
    return DIMB_ERR_ARG;
  }
  if (n[0] == 0 || n[1] == 0) {  // "no keypoints" return of the reference (lightglue.py:518-538): stop = 1
    *stop_layer = 1;
    return DIMB_OK;
  }
  int cur[2] = {0, 0};  // which ping-pong copy holds the live state of each side
  std::vector<int> ind[2];
  for (int s = 0; s < 2; ++s) {
    const dimb_feats& f = *F[s];
    ind[s].resize(n[s]);
    std::iota(ind[s].begin(), ind[s].end(), 0);
    // descriptors -> [n][din] on the device (layout 0 = (D,N): transpose on the host, this is not a tuned path)
    std::vector<float> tmp;
    const float* src = f.descriptors;
    if (f.desc_layout == 0) {
      const int ld = f.desc_ld ? f.desc_ld : f.n;
      tmp.resize(static_cast<size_t>(n[s]) * din);
      for (int c = 0; c < din; ++c)
        for (int i = 0; i < n[s]; ++i) tmp[static_cast<size_t>(i) * din + c] = f.descriptors[static_cast<size_t>(c) * ld + i];
      src = tmp.data();
    } else if (f.desc_ld && f.desc_ld != din) {
      tmp.resize(static_cast<size_t>(n[s]) * din);
      for (int i = 0; i < n[s]; ++i) std::memcpy(&tmp[static_cast<size_t>(i) * din], f.descriptors + static_cast<size_t>(i) * f.desc_ld, din * sizeof(float));
      src = tmp.data();
    }
    DIMB_CUDA_OK(ctx, cudaMemcpy(g->desc_in, src, static_cast<size_t>(n[s]) * din * sizeof(float), cudaMemcpyHostToDevice));
    DIMB_CUDA_OK(ctx, cudaMemcpy(g->kpts, f.keypoints, static_cast<size_t>(n[s]) * 2 * sizeof(float), cudaMemcpyHostToDevice));
    float s0 = f.size0, s1 = f.size1;
    if (!f.has_size) {  // size = 1 + kpts.max(-2) - kpts.min(-2)   (lightglue.py:26-27)
      float mn0 = INFINITY, mn1 = INFINITY, mx0 = -INFINITY, mx1 = -INFINITY;
      for (int i = 0; i < n[s]; ++i) {
        mn0 = std::min(mn0, f.keypoints[2 * i]), mx0 = std::max(mx0, f.keypoints[2 * i]);
        mn1 = std::min(mn1, f.keypoints[2 * i + 1]), mx1 = std::max(mx1, f.keypoints[2 * i + 1]);
      }
      s0 = 1.f + mx0 - mn0, s1 = 1.f + mx1 - mn1;
    }
    gx_posenc_kernel<<<n[s], std::max(32, hd / 2), 0, st>>>(g->kpts, n[s], s0, s1, g->Wr, hd, g->enc[s][0], NP);
    DIMB_LAUNCH_CHECK(ctx);
    if (din != d) {
      DIMB_TRY(linear(g, st, g->desc_in, din, g->input_proj, g->cat[s][0], 2 * d, n[s]));
    } else {
      DIMB_CUDA_OK(ctx, cudaMemcpy2DAsync(g->cat[s][0], 2 * d * sizeof(float), g->desc_in, d * sizeof(float), d * sizeof(float), n[s],
                                          cudaMemcpyDeviceToDevice, st));
    }
    DIMB_CUDA_OK(ctx, cudaStreamSynchronize(st));  // desc_in / kpts are reused by the other side
  }
  const bool do_stop = cf.depth_confidence > 0, do_prune = cf.width_confidence > 0;
  const int m_total = n[0] + n[1];
  std::vector<float> tok[2], sc;
  bool have_tok = false;
  int i = 0;
  for (i = 0; i < L; ++i) {
    if (n[0] == 0 || n[1] == 0) break;
    const Block &sb = g->self_[i], &cb = g->cross_[i];
    for (int s = 0; s < 2; ++s) {  // self block (lightglue.py:146-159)
      float* cat = g->cat[s][cur[s]];
      DIMB_TRY(linear(g, st, cat, 2 * d, sb.qkv, g->qkv, 3 * d, n[s]));
      gx_qkv_rotary_kernel<<<n[s], std::max(32, d / 2), 0, st>>>(g->qkv, n[s], d, hd, g->enc[s][cur[s]], NP, g->q[s], g->k[s], g->v[s]);
      DIMB_LAUNCH_CHECK(ctx);
      DIMB_TRY(attention(g, st, g->q[s], g->k[s], g->v[s], n[s], n[s], g->hid, d));
      DIMB_TRY(linear(g, st, g->hid, d, sb.out, cat + d, 2 * d, n[s]));
      DIMB_TRY(ffn(g, st, cat, n[s], sb));
    }
    for (int s = 0; s < 2; ++s) {  // cross block (lightglue.py:186-211): shared q/k projection, v projection
      float* cat = g->cat[s][cur[s]];
      DIMB_TRY(linear(g, st, cat, 2 * d, cb.to_qk, g->q[s], d, n[s]));
      DIMB_TRY(linear(g, st, cat, 2 * d, cb.to_v, g->v[s], d, n[s]));
    }
    for (int s = 0; s < 2; ++s) {
      float* cat = g->cat[s][cur[s]];
      DIMB_TRY(attention(g, st, g->q[s], g->q[1 - s], g->v[1 - s], n[s], n[1 - s], g->hid, d));
      DIMB_TRY(linear(g, st, g->hid, d, cb.out, cat + d, 2 * d, n[s]));
    }
    for (int s = 0; s < 2; ++s) DIMB_TRY(ffn(g, st, g->cat[s][cur[s]], n[s], cb));
    if (i == L - 1) continue;
    if (do_stop) {  // token confidence + check_if_stop (lightglue.py:73-83, 593-604)
      const float thr = conf_threshold(i, L);
      int below = 0;
      for (int s = 0; s < 2; ++s) {
        gx_rowdot_kernel<<<ceil_div(n[s] * 32, 256), 256, 0, st>>>(g->cat[s][cur[s]], 2 * d, n[s], d, g->token[i].w, g->token[i].b, g->zt[s]);
        DIMB_LAUNCH_CHECK(ctx);
        tok[s].resize(n[s]);
        DIMB_CUDA_OK(ctx, cudaMemcpyAsync(tok[s].data(), g->zt[s], n[s] * sizeof(float), cudaMemcpyDeviceToHost, st));
      }
      DIMB_CUDA_OK(ctx, cudaStreamSynchronize(st));
      for (int s = 0; s < 2; ++s)
        for (float& z : tok[s]) {
          z = 1.f / (1.f + std::exp(-z));
          below += z < thr;
        }
      have_tok = true;
      const float ratio = 1.0f - static_cast<float>(below) / static_cast<float>(m_total);
      if (ratio > static_cast<float>(cf.depth_confidence)) break;
    }
    for (int s = 0; s < 2 && do_prune; ++s) {  // pruning (lightglue.py:481-516, 586-591)
      if (n[s] <= cf.prune_min_kpts) continue;
      gx_rowdot_kernel<<<ceil_div(n[s] * 32, 256), 256, 0, st>>>(g->cat[s][cur[s]], 2 * d, n[s], d, g->matchab[i].w, g->matchab[i].b, g->zt[s]);
      DIMB_LAUNCH_CHECK(ctx);
      sc.resize(n[s]);
      DIMB_CUDA_OK(ctx, cudaMemcpyAsync(sc.data(), g->zt[s], n[s] * sizeof(float), cudaMemcpyDeviceToHost, st));
      DIMB_CUDA_OK(ctx, cudaStreamSynchronize(st));
      const float thr = conf_threshold(i, L), keep_thr = static_cast<float>(1.0 - cf.width_confidence);
      std::vector<int> kidx;
      for (int j = 0; j < n[s]; ++j) {
        bool keep = 1.f / (1.f + std::exp(-sc[j])) > keep_thr;
        if (have_tok) keep = keep || tok[s][j] <= thr;
        if (keep) kidx.push_back(j);
      }
      const int nn = static_cast<int>(kidx.size());
      if (nn) {
        DIMB_CUDA_OK(ctx, cudaMemcpyAsync(g->idx, kidx.data(), nn * sizeof(int), cudaMemcpyHostToDevice, st));
        gx_gather_kernel<<<nn, 128, 0, st>>>(g->cat[s][cur[s]], g->cat[s][1 - cur[s]], 2 * d, d, g->enc[s][cur[s]], g->enc[s][1 - cur[s]], hd,
                                              NP, g->idx, nn);
        DIMB_LAUNCH_CHECK(ctx);
        DIMB_CUDA_OK(ctx, cudaStreamSynchronize(st));
      }
      std::vector<int> ni(nn);
      for (int j = 0; j < nn; ++j) ni[j] = ind[s][kidx[j]];
      ind[s].swap(ni);
      if (have_tok) {
        std::vector<float> nt(nn);
        for (int j = 0; j < nn; ++j) nt[j] = tok[s][kidx[j]];
        tok[s].swap(nt);
      }
      cur[s] = 1 - cur[s];
      n[s] = nn;
    }
  }
  *stop_layer = std::min(i, L - 1) + 1;
  if (n[0] == 0 || n[1] == 0) return DIMB_OK;
  const int li = std::min(i, L - 1);
  // ---- assignment (lightglue.py:246-275) and filter_matches (:281-297)
  const float inv = 1.f / std::pow(static_cast<float>(d), 0.25f);
  for (int s = 0; s < 2; ++s) {
    DIMB_TRY(linear(g, st, g->cat[s][cur[s]], 2 * d, g->final_proj[li], g->md[s], d, n[s], inv));
    gx_rowdot_kernel<<<ceil_div(n[s] * 32, 256), 256, 0, st>>>(g->cat[s][cur[s]], 2 * d, n[s], d, g->matchab[li].w, g->matchab[li].b, g->zt[s]);
    DIMB_LAUNCH_CHECK(ctx);
  }
  {
    dim3 grid(ceil_div(n[1], 64), ceil_div(n[0], 64));
    gx_linear_kernel<<<grid, 256, 0, st>>>(g->md[0], d, g->md[1], d, nullptr, g->sim, NP, n[0], n[1], d, 1.f, nullptr, 0);
    DIMB_LAUNCH_CHECK(ctx);
  }
  gx_lse_kernel<<<ceil_div(n[0] * 32, 256), 256, 0, st>>>(g->sim, NP, n[0], n[1], 0, g->rlse);
  DIMB_LAUNCH_CHECK(ctx);
  gx_lse_kernel<<<ceil_div(n[1] * 32, 256), 256, 0, st>>>(g->sim, NP, n[0], n[1], 1, g->clse);
  DIMB_LAUNCH_CHECK(ctx);
  gx_argmax_kernel<<<ceil_div(n[0] * 32, 256), 256, 0, st>>>(g->sim, NP, n[0], n[1], g->rlse, g->clse, g->zt[0], g->zt[1], 0, g->best0, g->arg0);
  DIMB_LAUNCH_CHECK(ctx);
  gx_argmax_kernel<<<ceil_div(n[1] * 32, 256), 256, 0, st>>>(g->sim, NP, n[0], n[1], g->rlse, g->clse, g->zt[0], g->zt[1], 1, g->best1, g->arg1);
  DIMB_LAUNCH_CHECK(ctx);
  std::vector<float> b0(n[0]);
  std::vector<int> a0(n[0]), a1(n[1]);
  DIMB_CUDA_OK(ctx, cudaMemcpyAsync(b0.data(), g->best0, n[0] * sizeof(float), cudaMemcpyDeviceToHost, st));
  DIMB_CUDA_OK(ctx, cudaMemcpyAsync(a0.data(), g->arg0, n[0] * sizeof(int), cudaMemcpyDeviceToHost, st));
  DIMB_CUDA_OK(ctx, cudaMemcpyAsync(a1.data(), g->arg1, n[1] * sizeof(int), cudaMemcpyDeviceToHost, st));
  DIMB_CUDA_OK(ctx, cudaStreamSynchronize(st));
  int cnt = 0;
  for (int r = 0; r < n[0]; ++r) {
    const int c = a0[r];
    if (a1[c] != r) continue;  // mutual
    const float e = std::exp(b0[r]);
    if (!(e > static_cast<float>(cf.filter_threshold))) continue;
    if (cnt < cap) {
      matches[2 * cnt] = ind[0][r];
      matches[2 * cnt + 1] = ind[1][c];
      mscores[cnt] = e;
    }
    ++cnt;
  }
  *n_matches = cnt;
  if (cnt > cap) {
    dimb_set_error(ctx, "dimb_lg_match: more matches than cap");
    return DIMB_ERR_CAPACITY;
  }
  return DIMB_OK;
}

int lgx_match(dimb_lgx* g, int P, const dimb_feats* f0, const dimb_feats* f1, int64_t* matches, float* mscores, int* n_matches,
              int* stop_layer, int cap) {
  dimb_ctx* ctx = g->ctx;
  OwnerScope own(ctx, &g->mem);
  DIMB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  for (int p = 0; p < P; ++p)
    DIMB_TRY(lgx_match_pair(g, f0[p], f1[p], matches + static_cast<size_t>(p) * cap * 2, mscores + static_cast<size_t>(p) * cap, n_matches + p,
                            stop_layer + p, cap));
  return DIMB_OK;
}
