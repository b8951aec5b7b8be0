// selftest.cu - dimb_selftest_gemm: C = A * B^T through the production tensor-core GEMM (or its SIMT twin),
// used by tests/ to validate the tcgen05/TMA plumbing in isolation from the model code.
#include <vector>

#include "gemm.cuh"

namespace {
__global__ void split_rows_kernel(const float* __restrict__ src, __half* __restrict__ hi, __half* __restrict__ lo, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  __half h, l;
  split_f32(src[i], h, l);
  hi[i] = h;
  lo[i] = l;
}
}  // namespace

// A [M][K], B [N][K], C [M][N] host fp32; K multiple of 64. bn: 64, 128 or 256 (CTA tile width).
extern "C" int dimb_selftest_gemm(dimb_ctx* ctx, const float* A, const float* B, float* C, int M, int N, int K, int bn) {
  if (!ctx || !A || !B || !C || K % 64 || M < 1 || N < 1) return DIMB_ERR_ARG;
  DIMB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  const int Mp = round_up(M, 128), Np = round_up(N, 256);
  float *dA, *dB, *dC;
  __half *ah, *al, *bh, *bl;
  std::vector<void*> tmp;
  auto alloc = [&](void** p, size_t b) -> int {
    DIMB_CUDA_OK(ctx, cudaMalloc(p, b));
    tmp.push_back(*p);
    DIMB_CUDA_OK(ctx, cudaMemset(*p, 0, b));
    return static_cast<int>(DIMB_OK);
  };
  int rc = DIMB_OK;
  do {
    if ((rc = alloc((void**)&dA, sizeof(float) * Mp * K))) break;
    if ((rc = alloc((void**)&dB, sizeof(float) * Np * K))) break;
    if ((rc = alloc((void**)&dC, sizeof(float) * Mp * N))) break;
    if ((rc = alloc((void**)&ah, sizeof(__half) * Mp * K))) break;
    if ((rc = alloc((void**)&al, sizeof(__half) * Mp * K))) break;
    if ((rc = alloc((void**)&bh, sizeof(__half) * Np * K))) break;
    if ((rc = alloc((void**)&bl, sizeof(__half) * Np * K))) break;
    cudaMemcpy(dA, A, sizeof(float) * M * K, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B, sizeof(float) * N * K, cudaMemcpyHostToDevice);
    split_rows_kernel<<<ceil_div(Mp * K, 256), 256>>>(dA, ah, al, static_cast<size_t>(Mp) * K);
    split_rows_kernel<<<ceil_div(Np * K, 256), 256>>>(dB, bh, bl, static_cast<size_t>(Np) * K);
    TcOperands ops;
    if ((rc = dimb_tmap_2d(ctx, &ops.Ah, ah, Mp, K, K, kTileM))) break;
    if ((rc = dimb_tmap_2d(ctx, &ops.Al, al, Mp, K, K, kTileM))) break;
    if ((rc = dimb_tmap_2d(ctx, &ops.Bh, bh, Np, K, K, bn))) break;
    if ((rc = dimb_tmap_2d(ctx, &ops.Bl, bl, Np, K, K, bn))) break;
    GemmArgs g{};
    g.num_kb = K / 64;
    g.M = M;
    g.N = N;
    g.Ah = ah;
    g.Al = al;
    g.Bh = bh;
    g.Bl = bl;
    g.lda = K;
    g.ldb = K;
    EpiStoreF32 e;
    e.out = dC;
    e.bias = nullptr;
    e.ldc = N;
    e.n_valid = N;
    e.m_valid = M;
    e.scale = 1.f;
    const int mt = Mp / 128;
    if (bn == 64)
      rc = launch_gemm<64, false>(ctx, 0, ops, g, e, mt, round_up(N, 64));
    else if (bn == 128)
      rc = launch_gemm<128, false>(ctx, 0, ops, g, e, mt, round_up(N, 128));
    else if (bn == 256)
      rc = launch_gemm<256, false>(ctx, 0, ops, g, e, mt, round_up(N, 256));
    else
      rc = DIMB_ERR_ARG;
    if (rc) break;
    cudaError_t ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) {
      dimb_set_error(ctx, std::string("dimb_selftest_gemm: ") + cudaGetErrorString(ce));
      rc = DIMB_ERR_CUDA;
      break;
    }
    cudaMemcpy(C, dC, sizeof(float) * M * N, cudaMemcpyDeviceToHost);
  } while (0);
  for (void* p : tmp) cudaFree(p);
  return rc;
}
