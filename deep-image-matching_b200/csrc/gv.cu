// gv.cu - geometric verification on the GPU (dimb_gv_*): fundamental-matrix RANSAC over the matches of a batch of pairs, the step
// that follows _match_pairs in the reference (utils/geometric_verification.py:45-179 -> pydegensac.findFundamentalMatrix /
// cv2.findFundamentalMat, called per pair from matchers/matcher_base.py:298-340).  At hundreds of pairs per second per GPU the
// CPU estimator becomes the bottleneck of the pipeline.
//
// Per pair: Hartley normalisation of the matched keypoints; H hypotheses (8 random correspondences each, counter-based RNG ->
// reproducible), each solved by the normalised 8-point algorithm and scored by its Sampson inlier count over all matches by ONE
// thread (gv_math.cuh); the best hypothesis is refitted twice by least squares on its inliers (local optimisation) and the final
// inlier mask is written.  No adaptive stopping: every hypothesis runs in parallel, H = min(max_iters, 8192).
// RANSAC is stochastic in the reference too (pydegensac's own RNG), so parity is statistical: tests compare inlier sets on data
// with known geometry and against OpenCV on the same matches.
#include <memory>
#include <vector>

#include "common.cuh"
#include "gv_math.cuh"

namespace {

struct GvPair {
  const float *k0, *k1;        // keypoints of image 0 / 1: (N,2) x,y
  const long long* matches;    // [n][2] indices into k0 / k1, or null: k0[i] <-> k1[i]
  const int* n_dev;            // device count (or null: n_host)
  int n_host, cap;
};

__device__ __forceinline__ int gv_count(const GvPair& p) { return min(p.n_dev ? *p.n_dev : p.n_host, p.cap); }
__device__ __forceinline__ void gv_point(const GvPair& p, int i, float& x0, float& y0, float& x1, float& y1) {
  const long long a = p.matches ? p.matches[2 * i] : i, b = p.matches ? p.matches[2 * i + 1] : i;
  x0 = p.k0[2 * a], y0 = p.k0[2 * a + 1], x1 = p.k1[2 * b], y1 = p.k1[2 * b + 1];
}

// one CTA per pair: centroid + mean distance of both point sets -> Hartley normalisations; packed coordinates [cap][4]
__global__ void gv_prepare_kernel(const GvPair* pairs, float* xy, gv::Norm* norms, unsigned long long* best, int cap) {
  const GvPair p = pairs[blockIdx.x];
  const int n = gv_count(p), t = threadIdx.x;
  float* out = xy + static_cast<size_t>(blockIdx.x) * cap * 4;
  __shared__ float red[4][32];
  __shared__ float mean[4];
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = t; i < n; i += blockDim.x) {
    float c[4];
    gv_point(p, i, c[0], c[1], c[2], c[3]);
    for (int k = 0; k < 4; ++k) out[4 * i + k] = c[k], s[k] += c[k];
  }
  for (int k = 0; k < 4; ++k) {
    for (int o = 16; o; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    if ((t & 31) == 0) red[k][t >> 5] = s[k];
  }
  __syncthreads();
  if (t < 4) {
    float a = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) a += red[t][w];
    mean[t] = n ? a / n : 0.f;
  }
  __syncthreads();
  float d[2] = {0.f, 0.f};
  for (int i = t; i < n; i += blockDim.x) {
    d[0] += sqrtf((out[4 * i] - mean[0]) * (out[4 * i] - mean[0]) + (out[4 * i + 1] - mean[1]) * (out[4 * i + 1] - mean[1]));
    d[1] += sqrtf((out[4 * i + 2] - mean[2]) * (out[4 * i + 2] - mean[2]) + (out[4 * i + 3] - mean[3]) * (out[4 * i + 3] - mean[3]));
  }
  __syncthreads();
  for (int k = 0; k < 2; ++k) {
    for (int o = 16; o; o >>= 1) d[k] += __shfl_xor_sync(0xffffffffu, d[k], o);
    if ((t & 31) == 0) red[k][t >> 5] = d[k];
  }
  __syncthreads();
  if (t < 2) {
    float a = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) a += red[t][w];
    const float md = n ? a / n : 1.f;
    norms[2 * blockIdx.x + t] = gv::Norm{mean[2 * t], mean[2 * t + 1], md > 0.f ? 1.41421356f / md : 1.f};
  }
  if (t == 0) best[blockIdx.x] = 0ull;
}

// grid (ceil(H / 128), P): one hypothesis per thread; best = max over (inliers << 32 | ~hyp) (ties -> lowest hypothesis index)
__global__ void gv_hypotheses_kernel(const GvPair* pairs, const float* xy, const gv::Norm* norms, unsigned long long* best, int cap, int H,
                                     float thr2, unsigned seed) {
  const int pi = blockIdx.y, h = blockIdx.x * blockDim.x + threadIdx.x;
  const GvPair p = pairs[pi];
  const int n = gv_count(p);
  if (n < 8 || h >= H) return;
  const float* pts = xy + static_cast<size_t>(pi) * cap * 4;
  int idx[8];
  gv::sample8(seed + 0x9E37u * pi, h, n, idx);
  float k0[16], k1[16];
  int id8[8];
  for (int k = 0; k < 8; ++k) {
    k0[2 * k] = pts[4 * idx[k]], k0[2 * k + 1] = pts[4 * idx[k] + 1], k1[2 * k] = pts[4 * idx[k] + 2], k1[2 * k + 1] = pts[4 * idx[k] + 3];
    id8[k] = k;
  }
  float F[9];
  if (!gv::eight_point(k0, k1, id8, norms[2 * pi], norms[2 * pi + 1], F)) return;
  int cnt = 0;
  const float4* p4 = reinterpret_cast<const float4*>(pts);
  for (int i = 0; i < n; ++i) {
    const float4 c = __ldg(p4 + i);
    cnt += gv::sampson2(F, c.x, c.y, c.z, c.w) < thr2;
  }
  atomicMax(&best[pi], (static_cast<unsigned long long>(cnt) << 32) | (0xffffffffu - static_cast<unsigned>(h)));
}

// one CTA per pair: best hypothesis -> two least-squares refits on its inliers -> mask, count, F
__global__ void __launch_bounds__(256)
gv_finalize_kernel(const GvPair* pairs, const float* xy, const gv::Norm* norms, const unsigned long long* best, int cap, float thr2, unsigned seed,
                   float* Fout, unsigned char* mask, int* n_inl) {
  const int pi = blockIdx.x, t = threadIdx.x;
  const GvPair p = pairs[pi];
  const int n = gv_count(p);
  unsigned char* mk = mask + static_cast<size_t>(pi) * cap;
  float* Fo = Fout + 9 * pi;
  __shared__ float F[9];
  __shared__ float Nm[45];
  __shared__ int ok, cnt;
  if (n < 8 || (best[pi] >> 32) < 8) {  // fewer than 8 matches / no usable model: every match stays (reference :107-111 returns all ones)
    for (int i = t; i < n; i += blockDim.x) mk[i] = 1;
    if (t < 9) Fo[t] = 0.f;
    if (t == 0) n_inl[pi] = n;
    return;
  }
  const gv::Norm n0 = norms[2 * pi], n1 = norms[2 * pi + 1];
  const float* pts = xy + static_cast<size_t>(pi) * cap * 4;
  if (t == 0) {
    const unsigned h = 0xffffffffu - static_cast<unsigned>(best[pi] & 0xffffffffu);
    int idx[8], id8[8];
    float k0[16], k1[16], f[9];
    gv::sample8(seed + 0x9E37u * pi, h, n, idx);
    for (int k = 0; k < 8; ++k) {
      k0[2 * k] = pts[4 * idx[k]], k0[2 * k + 1] = pts[4 * idx[k] + 1], k1[2 * k] = pts[4 * idx[k] + 2], k1[2 * k + 1] = pts[4 * idx[k] + 3];
      id8[k] = k;
    }
    gv::eight_point(k0, k1, id8, n0, n1, f);
    for (int k = 0; k < 9; ++k) F[k] = f[k];
  }
  __syncthreads();
  for (int round = 0; round < 2; ++round) {
    if (t < 45) Nm[t] = 0.f;
    if (t == 0) cnt = 0;
    __syncthreads();
    float acc[45];
#pragma unroll
    for (int k = 0; k < 45; ++k) acc[k] = 0.f;
    int c = 0;
    for (int i = t; i < n; i += blockDim.x) {
      const float x0 = pts[4 * i], y0 = pts[4 * i + 1], x1 = pts[4 * i + 2], y1 = pts[4 * i + 3];
      if (gv::sampson2(F, x0, y0, x1, y1) < thr2) {
        const float u0 = (x0 - n0.cx) * n0.s, v0 = (y0 - n0.cy) * n0.s, u1 = (x1 - n1.cx) * n1.s, v1 = (y1 - n1.cy) * n1.s;
        const float a[9] = {u1 * u0, u1 * v0, u1, v1 * u0, v1 * v0, v1, u0, v0, 1.f};
        int k = 0;
#pragma unroll
        for (int r = 0; r < 9; ++r)
#pragma unroll
          for (int q = r; q < 9; ++q) acc[k++] += a[r] * a[q];
        ++c;
      }
    }
#pragma unroll
    for (int k = 0; k < 45; ++k) {
      float v = acc[k];
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((t & 31) == 0) atomicAdd(&Nm[k], v);
    }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((t & 31) == 0) atomicAdd(&cnt, c);
    __syncthreads();
    if (t == 0) {
      ok = 0;
      if (cnt >= 8) {
        float N9[9][9], f[9];
        int k = 0;
        for (int r = 0; r < 9; ++r)
          for (int q = r; q < 9; ++q) N9[r][q] = N9[q][r] = Nm[k++];
        if (gv::refit_from_normal(N9, n0, n1, f)) {
          // keep the refit only if it does not lose inliers (checked below by the caller loop: count with the new model)
          for (int j = 0; j < 9; ++j) Nm[j] = f[j];
          ok = 1;
        }
      }
    }
    __syncthreads();
    if (ok) {  // candidate model in Nm[0..8]: accept if it explains at least as many matches
      __shared__ int c_new;
      if (t == 0) c_new = 0;
      __syncthreads();
      float f[9];
      for (int j = 0; j < 9; ++j) f[j] = Nm[j];
      int cn = 0;
      for (int i = t; i < n; i += blockDim.x) cn += gv::sampson2(f, pts[4 * i], pts[4 * i + 1], pts[4 * i + 2], pts[4 * i + 3]) < thr2;
      for (int o = 16; o; o >>= 1) cn += __shfl_xor_sync(0xffffffffu, cn, o);
      if ((t & 31) == 0) atomicAdd(&c_new, cn);
      __syncthreads();
      if (t < 9 && c_new >= cnt) F[t] = f[t];
      __syncthreads();
    }
  }
  if (t == 0) cnt = 0;
  __syncthreads();
  int c = 0;
  for (int i = t; i < n; i += blockDim.x) {
    const unsigned char in = gv::sampson2(F, pts[4 * i], pts[4 * i + 1], pts[4 * i + 2], pts[4 * i + 3]) < thr2;
    mk[i] = in;
    c += in;
  }
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((t & 31) == 0) atomicAdd(&cnt, c);
  __syncthreads();
  if (t < 9) Fo[t] = F[t];
  if (t == 0) n_inl[pi] = cnt;
}

int gv_run(dimb_ctx* ctx, cudaStream_t st, const std::vector<GvPair>& hp, int cap, float threshold, int max_iters, unsigned seed, float* d_F,
           unsigned char* d_mask, int* d_ninl) {
  const int P = static_cast<int>(hp.size());
  GvPair* d_pairs;
  float* d_xy;
  gv::Norm* d_norm;
  unsigned long long* d_best;
  DIMB_TRY(dimb_scratch(ctx, 40, P * sizeof(GvPair), reinterpret_cast<void**>(&d_pairs)));
  DIMB_TRY(dimb_scratch(ctx, 41, static_cast<size_t>(P) * cap * 4 * sizeof(float), reinterpret_cast<void**>(&d_xy)));
  DIMB_TRY(dimb_scratch(ctx, 42, 2 * P * sizeof(gv::Norm), reinterpret_cast<void**>(&d_norm)));
  DIMB_TRY(dimb_scratch(ctx, 43, P * sizeof(unsigned long long), reinterpret_cast<void**>(&d_best)));
  DIMB_CUDA_OK(ctx, cudaMemcpyAsync(d_pairs, hp.data(), P * sizeof(GvPair), cudaMemcpyHostToDevice, st));
  const int H = std::max(64, std::min(max_iters, 8192));
  const float thr2 = threshold * threshold;
  ProfScope prof(ctx, st, "gv.ransac");
  gv_prepare_kernel<<<P, 256, 0, st>>>(d_pairs, d_xy, d_norm, d_best, cap);
  DIMB_LAUNCH_CHECK(ctx);
  gv_hypotheses_kernel<<<dim3(ceil_div(H, 128), P), 128, 0, st>>>(d_pairs, d_xy, d_norm, d_best, cap, H, thr2, seed);
  DIMB_LAUNCH_CHECK(ctx);
  gv_finalize_kernel<<<P, 256, 0, st>>>(d_pairs, d_xy, d_norm, d_best, cap, thr2, seed, d_F, d_mask, d_ninl);
  DIMB_LAUNCH_CHECK(ctx);
  return DIMB_OK;
}

}  // namespace

extern "C" {

// Host entry: matched keypoints kpts0[i] <-> kpts1[i] (n,2) float32 pixels.  F [9] row-major with x1^T F x0 = 0 (zeros when n < 8 or
// no model was found - the reference returns F = None and an all-True mask then), mask [n] 0/1, n_inliers.
int dimb_gv_fundamental(dimb_ctx* ctx, const float* kpts0, const float* kpts1, int n, float threshold, int max_iters, unsigned seed, float* F,
                        unsigned char* mask, int* n_inliers) {
  if (!ctx || n < 0 || (n > 0 && (!kpts0 || !kpts1)) || !F || !mask || !n_inliers || threshold <= 0.f) return DIMB_ERR_ARG;
  DIMB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  if (n == 0) {
    for (int i = 0; i < 9; ++i) F[i] = 0.f;
    *n_inliers = 0;
    return DIMB_OK;
  }
  cudaStream_t st = 0;
  float *d_k0, *d_k1, *d_F;
  unsigned char* d_mask;
  int* d_n;
  DIMB_TRY(dimb_scratch(ctx, 44, static_cast<size_t>(n) * 2 * sizeof(float), reinterpret_cast<void**>(&d_k0)));
  DIMB_TRY(dimb_scratch(ctx, 45, static_cast<size_t>(n) * 2 * sizeof(float), reinterpret_cast<void**>(&d_k1)));
  DIMB_TRY(dimb_scratch(ctx, 46, 9 * sizeof(float) + sizeof(int), reinterpret_cast<void**>(&d_F)));
  DIMB_TRY(dimb_scratch(ctx, 47, n, reinterpret_cast<void**>(&d_mask)));
  d_n = reinterpret_cast<int*>(d_F + 9);
  DIMB_CUDA_OK(ctx, cudaMemcpyAsync(d_k0, kpts0, static_cast<size_t>(n) * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
  DIMB_CUDA_OK(ctx, cudaMemcpyAsync(d_k1, kpts1, static_cast<size_t>(n) * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
  std::vector<GvPair> hp(1);
  hp[0] = GvPair{d_k0, d_k1, nullptr, nullptr, n, n};
  DIMB_TRY(gv_run(ctx, st, hp, n, threshold, max_iters, seed, d_F, d_mask, d_n));
  DIMB_CUDA_OK(ctx, cudaMemcpyAsync(F, d_F, 9 * sizeof(float), cudaMemcpyDeviceToHost, st));
  DIMB_CUDA_OK(ctx, cudaMemcpyAsync(n_inliers, d_n, sizeof(int), cudaMemcpyDeviceToHost, st));
  DIMB_CUDA_OK(ctx, cudaMemcpyAsync(mask, d_mask, n, cudaMemcpyDeviceToHost, st));
  DIMB_CUDA_OK(ctx, cudaStreamSynchronize(st));
  return DIMB_OK;
}

// Batch on device buffers, asynchronous on `stream`: pair p verifies d_matches[p][0..d_n_matches[p]) (the output layout of
// dimb_lg_match_dev / dimb_pipe_*: [P][cap][2] int64, [P] counts) against the keypoint arrays d_kpts0[p] / d_kpts1[p] ((N,2) float32).
// Outputs (device): d_F [P][9], d_mask [P][cap] (0/1 per match), d_n_inliers [P].
int dimb_gv_fundamental_batch_dev(dimb_ctx* ctx, int P, const float* const* d_kpts0, const float* const* d_kpts1, const int64_t* d_matches,
                                  const int* d_n_matches, int cap, float threshold, int max_iters, unsigned seed, float* d_F,
                                  unsigned char* d_mask, int* d_n_inliers, void* stream) {
  if (!ctx || P < 1 || !d_kpts0 || !d_kpts1 || !d_matches || !d_n_matches || cap < 1 || !d_F || !d_mask || !d_n_inliers || threshold <= 0.f)
    return DIMB_ERR_ARG;
  DIMB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  std::vector<GvPair> hp(P);
  for (int p = 0; p < P; ++p)
    hp[p] = GvPair{d_kpts0[p], d_kpts1[p], reinterpret_cast<const long long*>(d_matches) + static_cast<size_t>(p) * cap * 2, d_n_matches + p, 0, cap};
  return gv_run(ctx, static_cast<cudaStream_t>(stream), hp, cap, threshold, max_iters, seed, d_F, d_mask, d_n_inliers);
}

}  // extern "C"
