// pipeline.cu - fused per-pair hot path (dimb_pipe_*): SuperPoint on both images of each pair, then LightGlue,
// with keypoints / descriptors staying in HBM between the two stages.  The reference writes every image's
// features to features.h5 (fp16, gzip-9: extractors/extractor_base.py:56-99) and re-reads them per pair
// (matchers/matcher_base.py:221-222); the only value-level effect of that round trip - the fp16 rounding of
// keypoints and descriptors - is reproduced on device (dimb_feats_dev.round_fp16).
#include <memory>
#include <vector>

#include "common.cuh"

extern "C" dimb_ctx* dimb_sp_ctx(dimb_sp* sp);

namespace {
// uint8 gray -> float32 (exact): what `image.astype(np.float32)` does in ExtractorBase.extract (extractor_base.py:201-202)
__global__ void u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, size_t n4) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const uchar4 v = reinterpret_cast<const uchar4*>(src)[i];
  reinterpret_cast<float4*>(dst)[i] = make_float4(v.x, v.y, v.z, v.w);
}
}  // namespace

struct dimb_pipe {
  std::vector<void*> mem;  // device memory owned by this handle
  dimb_ctx* ctx;
  dimb_sp* sp;
  dimb_lg* lg;
  int max_pairs, H, W, cap;
  float *d_img, *d_kpts, *d_scores, *d_desc, *d_ms;
  uint8_t* d_img8;
  int *d_cnt, *d_nm, *d_sl;
  long long* d_m;
  cudaStream_t st = nullptr, st_copy = nullptr;  // compute stream; host->device copy stream of the host-buffer entries
  cudaEvent_t ev_chunk[4] = {nullptr, nullptr, nullptr, nullptr}, ev_free = nullptr;
};

extern "C" {

int dimb_pipe_create(dimb_sp* sp, dimb_lg* lg, int max_pairs, int H, int W, int cap, dimb_pipe** out) {
  if (!sp || !lg || !out || max_pairs < 1 || cap < 1) return DIMB_ERR_ARG;
  dimb_ctx* ctx = dimb_sp_ctx(sp);
  dimb_pipe* p = new dimb_pipe();
  p->ctx = ctx;
  std::unique_ptr<dimb_pipe, void (*)(dimb_pipe*)> guard(p, dimb_pipe_destroy);  // a failed create releases what it built
  OwnerScope own(ctx, &p->mem);
  p->sp = sp;
  p->lg = lg;
  p->max_pairs = max_pairs;
  p->H = H;
  p->W = W;
  p->cap = cap;
  const size_t B = 2 * static_cast<size_t>(max_pairs);
  DIMB_TRY(dimb_alloc_t(ctx, &p->d_img, B * H * W));
  DIMB_TRY(dimb_alloc_t(ctx, &p->d_img8, B * H * W));
  DIMB_TRY(dimb_alloc_t(ctx, &p->d_kpts, B * cap * 2));
  DIMB_TRY(dimb_alloc_t(ctx, &p->d_scores, B * cap));
  DIMB_TRY(dimb_alloc_t(ctx, &p->d_desc, B * 256 * cap));
  DIMB_TRY(dimb_alloc_t(ctx, &p->d_cnt, B));
  DIMB_TRY(dimb_alloc_t(ctx, &p->d_m, static_cast<size_t>(max_pairs) * cap * 2));
  DIMB_TRY(dimb_alloc_t(ctx, &p->d_ms, static_cast<size_t>(max_pairs) * cap));
  DIMB_TRY(dimb_alloc_t(ctx, &p->d_nm, max_pairs));
  DIMB_TRY(dimb_alloc_t(ctx, &p->d_sl, max_pairs));
  DIMB_CUDA_OK(ctx, cudaStreamCreateWithFlags(&p->st, cudaStreamNonBlocking));
  DIMB_CUDA_OK(ctx, cudaStreamCreateWithFlags(&p->st_copy, cudaStreamNonBlocking));
  for (auto& e : p->ev_chunk) DIMB_CUDA_OK(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  DIMB_CUDA_OK(ctx, cudaEventCreateWithFlags(&p->ev_free, cudaEventDisableTiming));
  *out = guard.release();
  return DIMB_OK;
}

void dimb_pipe_destroy(dimb_pipe* p) {
  if (!p) return;
  dimb_release(p->ctx, p->mem);
  if (p->st) cudaStreamDestroy(p->st);
  if (p->st_copy) cudaStreamDestroy(p->st_copy);
  for (auto& e : p->ev_chunk)
    if (e) cudaEventDestroy(e);
  if (p->ev_free) cudaEventDestroy(p->ev_free);
  delete p;
}

static int pipe_extract(dimb_pipe* p, const float* d_images, int i0, int n, void* stream);
static int pipe_match(dimb_pipe* p, int P, void* stream);

// d_images: device float32 [2P][H][W]; results stay in the pipe's device buffers (dimb_pipe_outputs_dev).
int dimb_pipe_match_image_pairs_dev(dimb_pipe* p, const float* d_images, int P, void* stream) {
  if (!p || !d_images || P < 1 || P > p->max_pairs) return DIMB_ERR_ARG;
  DIMB_TRY(pipe_extract(p, d_images, 0, 2 * P, stream));
  return pipe_match(p, P, stream);
}

// SuperPoint on images [i0, i0 + n) of the batch; features land in the pipe's per-image slots
static int pipe_extract(dimb_pipe* p, const float* d_images, int i0, int n, void* stream) {
  const size_t cap = p->cap, px = static_cast<size_t>(p->H) * p->W;
  return dimb_sp_extract_dev(p->sp, d_images + i0 * px, n, p->H, p->W, p->d_kpts + i0 * cap * 2, p->d_scores + i0 * cap,
                             p->d_desc + i0 * 256 * cap, p->d_cnt + i0, p->cap, stream);
}

// LightGlue on the P pairs (image 2i, image 2i+1) whose features are in the pipe's slots
static int pipe_match(dimb_pipe* p, int P, void* stream) {
  const int B = 2 * P, cap = p->cap;
  std::vector<dimb_feats_dev> f0(P), f1(P);
  for (int i = 0; i < B; ++i) {
    dimb_feats_dev& f = (i & 1) ? f1[i >> 1] : f0[i >> 1];
    f.keypoints = p->d_kpts + static_cast<size_t>(i) * cap * 2;
    f.descriptors = p->d_desc + static_cast<size_t>(i) * 256 * cap;
    f.n = p->d_cnt + i;
    f.n_cap = cap;
    f.desc_layout = 0;
    f.desc_ld = cap;
    f.size0 = static_cast<float>(p->H);  // image_size = image.shape[:2] = [H, W] (extractor_base.py:227, quirk A.3)
    f.size1 = static_cast<float>(p->W);
    f.round_fp16 = 1;
  }
  return dimb_lg_match_dev(p->lg, P, f0.data(), f1.data(), reinterpret_cast<int64_t*>(p->d_m), p->d_ms, p->d_nm, p->d_sl, cap, stream);
}

int dimb_pipe_outputs_dev(dimb_pipe* p, int64_t** d_matches, float** d_mscores, int** d_n_matches, int** d_stop, int** d_nkpts,
                          float** d_kpts) {
  if (!p) return DIMB_ERR_ARG;
  if (d_matches) *d_matches = reinterpret_cast<int64_t*>(p->d_m);
  if (d_mscores) *d_mscores = p->d_ms;
  if (d_n_matches) *d_n_matches = p->d_nm;
  if (d_stop) *d_stop = p->d_sl;
  if (d_nkpts) *d_nkpts = p->d_cnt;
  if (d_kpts) *d_kpts = p->d_kpts;
  return DIMB_OK;
}

// the SuperPoint features of the LAST call, still in HBM (layouts of dimb_sp_extract_dev with the pipe's cap): what the
// reference would have written to features.h5 before the fp16 cast
int dimb_pipe_features_dev(dimb_pipe* p, float** d_kpts, float** d_scores, float** d_desc, int** d_counts) {
  if (!p) return DIMB_ERR_ARG;
  if (d_kpts) *d_kpts = p->d_kpts;
  if (d_scores) *d_scores = p->d_scores;
  if (d_desc) *d_desc = p->d_desc;
  if (d_counts) *d_counts = p->d_cnt;
  return DIMB_OK;
}

static int pipe_finish(dimb_pipe* p, int P, int64_t* matches, float* mscores, int* n_matches, int* stop_layer, int* n_kpts, float* kpts) {
  dimb_ctx* ctx = p->ctx;
  const size_t B = 2 * static_cast<size_t>(P), cap = p->cap;
  DIMB_TRY(pipe_match(p, P, p->st));
  DIMB_CUDA_OK(ctx, cudaMemcpyAsync(n_matches, p->d_nm, P * sizeof(int), cudaMemcpyDeviceToHost, p->st));
  DIMB_CUDA_OK(ctx, cudaMemcpyAsync(stop_layer, p->d_sl, P * sizeof(int), cudaMemcpyDeviceToHost, p->st));
  DIMB_CUDA_OK(ctx, cudaMemcpyAsync(n_kpts, p->d_cnt, B * sizeof(int), cudaMemcpyDeviceToHost, p->st));
  DIMB_CUDA_OK(ctx, cudaMemcpyAsync(matches, p->d_m, P * cap * 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, p->st));
  DIMB_CUDA_OK(ctx, cudaMemcpyAsync(mscores, p->d_ms, P * cap * sizeof(float), cudaMemcpyDeviceToHost, p->st));
  if (kpts) DIMB_CUDA_OK(ctx, cudaMemcpyAsync(kpts, p->d_kpts, B * cap * 2 * sizeof(float), cudaMemcpyDeviceToHost, p->st));
  DIMB_CUDA_OK(ctx, cudaStreamSynchronize(p->st));
  return DIMB_OK;
}

// Host -> device staging of a batch of B images in up to four chunks of growing size (B/8, B/8, B/4, B/2 images): the first copy - the
// only one nothing can hide, the call being synchronous - is short, every later one overlaps the extraction of the chunks before it.
static int pipe_chunks(int B, int (&bounds)[5]) {
  const int cand[5] = {0, B / 8, B / 4, B / 2, B};
  int n = 0;
  bounds[0] = 0;
  for (int i = 1; i < 5; ++i)
    if (cand[i] > bounds[n]) bounds[++n] = cand[i];
  return n;  // chunks: [bounds[c], bounds[c + 1])
}

// Same as dimb_pipe_match_image_pairs with 8-bit gray images (what cv2 / rasterio deliver before the reference's
// astype(float32)): a quarter of the host->device traffic; the conversion on device is exact.
int dimb_pipe_match_image_pairs_u8(dimb_pipe* p, const uint8_t* images, int P, int64_t* matches, float* mscores, int* n_matches,
                                   int* stop_layer, int* n_kpts, float* kpts) {
  if (!p || !images || !matches || !mscores || !n_matches || !stop_layer || !n_kpts || P < 1 || P > p->max_pairs) return DIMB_ERR_ARG;
  dimb_ctx* ctx = p->ctx;
  DIMB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  const size_t px = static_cast<size_t>(p->H) * p->W;
  if (px % 4) return DIMB_ERR_ARG;
  int bounds[5];
  const int nc = pipe_chunks(2 * P, bounds);
  DIMB_CUDA_OK(ctx, cudaEventRecord(p->ev_free, p->st));  // the previous call's kernels are done with d_img8 / d_img
  DIMB_CUDA_OK(ctx, cudaStreamWaitEvent(p->st_copy, p->ev_free, 0));
  for (int c = 0; c < nc; ++c) {
    const size_t o = bounds[c] * px, n = (bounds[c + 1] - bounds[c]) * px;
    DIMB_CUDA_OK(ctx, cudaMemcpyAsync(p->d_img8 + o, images + o, n, cudaMemcpyHostToDevice, p->st_copy));
    DIMB_CUDA_OK(ctx, cudaEventRecord(p->ev_chunk[c], p->st_copy));
  }
  for (int c = 0; c < nc; ++c) {
    const size_t o = bounds[c] * px, n = (bounds[c + 1] - bounds[c]) * px;
    DIMB_CUDA_OK(ctx, cudaStreamWaitEvent(p->st, p->ev_chunk[c], 0));
    u8_to_f32_kernel<<<static_cast<unsigned>((n / 4 + 255) / 256), 256, 0, p->st>>>(p->d_img8 + o, p->d_img + o, n / 4);
    DIMB_LAUNCH_CHECK(ctx);
    DIMB_TRY(pipe_extract(p, p->d_img, bounds[c], bounds[c + 1] - bounds[c], p->st));
  }
  return pipe_finish(p, P, matches, mscores, n_matches, stop_layer, n_kpts, kpts);
}

// images: HOST float32 [2P][H][W] (pinned memory makes the copies asynchronous). Outputs HOST: matches [P][cap][2],
// mscores [P][cap], n_matches [P], stop_layer [P], n_kpts [2P], kpts [2P][cap][2] (kpts may be NULL).
int dimb_pipe_match_image_pairs(dimb_pipe* p, const float* images, int P, int64_t* matches, float* mscores, int* n_matches,
                                int* stop_layer, int* n_kpts, float* kpts) {
  if (!p || !images || !matches || !mscores || !n_matches || !stop_layer || !n_kpts || P < 1 || P > p->max_pairs) return DIMB_ERR_ARG;
  dimb_ctx* ctx = p->ctx;
  DIMB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  const size_t px = static_cast<size_t>(p->H) * p->W;
  int bounds[5];
  const int nc = pipe_chunks(2 * P, bounds);
  DIMB_CUDA_OK(ctx, cudaEventRecord(p->ev_free, p->st));
  DIMB_CUDA_OK(ctx, cudaStreamWaitEvent(p->st_copy, p->ev_free, 0));
  for (int c = 0; c < nc; ++c) {
    const size_t o = bounds[c] * px, n = (bounds[c + 1] - bounds[c]) * px;
    DIMB_CUDA_OK(ctx, cudaMemcpyAsync(p->d_img + o, images + o, n * sizeof(float), cudaMemcpyHostToDevice, p->st_copy));
    DIMB_CUDA_OK(ctx, cudaEventRecord(p->ev_chunk[c], p->st_copy));
  }
  for (int c = 0; c < nc; ++c) {
    DIMB_CUDA_OK(ctx, cudaStreamWaitEvent(p->st, p->ev_chunk[c], 0));
    DIMB_TRY(pipe_extract(p, p->d_img, bounds[c], bounds[c + 1] - bounds[c], p->st));
  }
  return pipe_finish(p, P, matches, mscores, n_matches, stop_layer, n_kpts, kpts);
}

}  // extern "C"
