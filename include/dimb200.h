/* dimb200.h - C ABI of libdimb200.so: the B200-native (sm_100a) hot path of
 * 3DOM-FBK/deep-image-matching behind plain pointers and sizes.
 *
 * Each entry point replaces the body of one reference plugin method (paths are
 * relative to src/deep_image_matching/ of the reference at 74d7bd5):
 *
 *   dimb_sp_extract      <- SuperPointExtractor._extract      extractors/superpoint.py:107-132
 *                           (+ model thirdparty/SuperGluePretrainedNetwork/models/superpoint.py:160-227)
 *   dimb_lg_match        <- LightGlueMatcher._match_pairs     matchers/lightglue.py:102-125
 *                           (+ featuresDict2Lightglue :8-66, model thirdparty/LightGlue/lightglue/lightglue.py:424-579)
 *   dimb_nn_match        <- KorniaMatcher._match_pairs        matchers/kornia_matcher.py:27-54
 *                           (kornia.feature.DescriptorMatcher modes nn/mnn/snn/smnn)
 *   *_dev variants       :  same computation on device pointers and a caller stream, so that
 *                           features never leave HBM between extraction and matching
 *                           (the reference round-trips them through features.h5, extractor_base.py:56-99).
 *
 * Conventions: every function returns DIMB_OK (0) or a negative error code; the message is
 * available from dimb_last_error(ctx) and contains "CUDA out of memory" for allocation failures
 * (matchers/matcher_base.py:251-254 keys its tile fallback on that text).  Caller owns all host
 * buffers; the library owns device memory inside its handles.  One ctx per device, not thread-safe.
 * There is NO CPU fallback: without a CUDA device every create call fails.
 */
#ifndef DIMB200_H
#define DIMB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dimb_ctx dimb_ctx;
typedef struct dimb_sp dimb_sp;
typedef struct dimb_lg dimb_lg;

enum {
  DIMB_OK = 0,
  DIMB_ERR_CUDA = -1,
  DIMB_ERR_OOM = -2,
  DIMB_ERR_ARG = -3,
  DIMB_ERR_UNSUPPORTED = -4,
  DIMB_ERR_CAPACITY = -5
};

/* Arithmetic mode of the tensor-core contractions:
 *   EXACT: fp16 hi+lo split operands, 3 MMAs per product, fp32 accumulate -> fp32-class results
 *          (parity mode, graded against the fp32 oracle at 1e-4).
 *   FAST : plain fp16 operands (statistically equivalent to the reference's TF32/fp16 GPU path). */
enum { DIMB_PRECISION_EXACT = 0, DIMB_PRECISION_FAST = 1 };

/* ------------------------------------------------------------------ context */
int dimb_ctx_create(int device, dimb_ctx** out);
void dimb_ctx_destroy(dimb_ctx* ctx);
const char* dimb_last_error(dimb_ctx* ctx);
int dimb_ctx_set_precision(dimb_ctx* ctx, int precision);
/* 1 (default): tcgen05 tensor-core kernels.  0: CUDA-core SIMT kernels with the same epilogues
 * (debug aid to bisect a tensor-path problem; also selectable with env DIMB_TC=0). */
int dimb_ctx_set_tensor_path(dimb_ctx* ctx, int use_tc);
/* Number of kernels this library has launched on ctx (bench.py "gpu_launches"). */
unsigned long long dimb_ctx_launch_count(dimb_ctx* ctx);
const char* dimb_version(void);
/* Synchronising device -> host copy of a buffer the library exposed through a *_dev accessor (tests, debug taps). */
int dimb_read_dev(dimb_ctx* ctx, void* dst, const void* d_src, size_t bytes);
/* Per-kernel-group device timing with CUDA events on the launching stream (bench.py roofline):
 * dimb_ctx_profile(ctx, 1) starts recording, dimb_ctx_profile_read returns the JSON text
 * {"<group>": [total_ms, launches], ...}; dimb_ctx_profile(ctx, 0) stops and clears. */
int dimb_ctx_profile(dimb_ctx* ctx, int enable);
int dimb_ctx_profile_read(dimb_ctx* ctx, char* buf, size_t n);

/* ------------------------------------------------------------------ SuperPoint */
typedef struct {
  int nms_radius;            /* config.py:96  (3)      */
  float keypoint_threshold;  /* config.py:97  (0.0005) */
  int max_keypoints;         /* config.py:98  (2048); -1 = unlimited */
  int remove_borders;        /* superpoint.py default (4) */
  int fix_sampling;          /* 0: thirdparty superpoint.py:81-98, 1: extractors/superpoint.py:16-27 */
  int max_batch;             /* workspace sizing: images per call */
  int max_height, max_width; /* workspace sizing */
} dimb_sp_conf;

/* weights: packed fp32, PyTorch OIHW tensors in this order, each weight followed by its bias:
 * conv1a conv1b conv2a conv2b conv3a conv3b conv4a conv4b convPa convPb convDa convDb
 * (1,300,865 floats for superpoint_v1). */
int dimb_sp_create(dimb_ctx* ctx, const float* weights, size_t n_floats, const dimb_sp_conf* conf, dimb_sp** out);
void dimb_sp_destroy(dimb_sp* sp);

/* images: host float32 [B][H][W], gray 0..255 (what ExtractorBase.extract hands to _extract).
 * Outputs (host, caller allocated): kpts [B][cap][2] (x,y) float32, scores [B][cap],
 * desc [B][256][cap] i.e. (D,N) with row pitch cap, counts [B].  Order: reference order
 * (row-major if <= max_keypoints candidates, else score-descending).  Returns
 * DIMB_ERR_CAPACITY (counts filled) if an image yields more than cap keypoints. */
int dimb_sp_extract(dimb_sp* sp, const float* images, int B, int H, int W, float* kpts, float* scores, float* desc,
                    int* counts, int cap);
/* Same on device pointers, asynchronous on `stream` (a cudaStream_t); counts stay on device. */
int dimb_sp_extract_dev(dimb_sp* sp, const float* d_images, int B, int H, int W, float* d_kpts, float* d_scores,
                        float* d_desc, int* d_counts, int cap, void* stream);
/* Debug taps (device->host copies of intermediates of the LAST extract call, image 0):
 * which: 0 = dense score map [H8*8][W8*8], 1 = nms map, 2 = encoder output [h][w][128] (fp32, NHWC),
 * 3 = dense descriptors (un-normalised convDb output) [h][w][256]. */
int dimb_sp_debug_read(dimb_sp* sp, int which, float* out, size_t n_floats);

/* ------------------------------------------------------------------ LightGlue */
typedef struct {
  int input_dim;           /* 256 superpoint, 128 aliked/disk (lightglue.py:330-359) */
  int descriptor_dim;      /* 256 (tensor-core kernels); any other shape, e.g. LighterGlue's 96, runs the generic fp32 path */
  int n_layers;            /* 9 (LighterGlue: 6) */
  int num_heads;           /* 4 (LighterGlue: 1); head dim = descriptor_dim / num_heads must be even and <= 128 */
  double depth_confidence; /* 0.95, -1 disables early exit (double: compared as float(x), like torch) */
  double width_confidence; /* 0.99, -1 disables point pruning; the keep test uses float(1 - width_confidence) */
  double filter_threshold; /* 0.1 */
  int prune_min_kpts;      /* 1536 = reference CUDA+flash semantics (lightglue.py:318-323,606-610) */
  int max_pairs;           /* workspace sizing: pairs per call */
  int max_kpts;            /* workspace sizing: keypoints per image */
} dimb_lg_conf;

/* weights: packed fp32 in this order (names as in the reference state_dict):
 *   posenc.Wr.weight (hd/2,2); [input_proj.weight (d,din), input_proj.bias] iff din != d;
 *   for i in layers: self_attn.{Wqkv,out_proj,ffn.0}.{weight,bias}, ffn.1.{weight,bias}, ffn.3.{weight,bias},
 *                    cross_attn.{to_qk,to_v,to_out,ffn.0}.{weight,bias}, ffn.1.{weight,bias}, ffn.3.{weight,bias};
 *   for i in layers: log_assignment.i.matchability.{weight,bias}, log_assignment.i.final_proj.{weight,bias};
 *   for i in layers-1: token_confidence.i.token.0.{weight,bias}.
 * Replaces LightGlue.__init__ + load_state_dict (lightglue.py:325-398) and, for descriptor_dim 96 / one head / 6 layers /
 * input_dim 64, LighterGlue.__init__ (thirdparty/accelerated_features/modules/lighterglue.py:29-48). */
int dimb_lg_create(dimb_ctx* ctx, const float* weights, size_t n_floats, const dimb_lg_conf* conf, dimb_lg** out);
void dimb_lg_destroy(dimb_lg* lg);

typedef struct {
  const float* keypoints;   /* (n,2) x,y float32 */
  const float* descriptors; /* float32; layout below */
  int n;                    /* number of keypoints */
  int desc_layout;          /* 0: (D,n) rows of pitch desc_ld (FeaturesDict layout), 1: (n,D) rows of pitch desc_ld */
  int desc_ld;              /* row pitch in floats (0 = dense) */
  int has_size;             /* 0: size := 1 + max(kpts) - min(kpts) (lightglue.py:26-27) */
  float size0, size1;       /* image_size exactly as the caller stores it ([H,W] in DIM; quirk A.3) */
} dimb_feats;

/* P pairs.  Outputs (host): matches [P][cap][2] int64 (ascending in column 0), mscores [P][cap],
 * n_matches [P], stop_layer [P] (1-based layer count executed, the reference's "stop").
 * Replaces LightGlueMatcher._match_pairs (matchers/lightglue.py:102-125) and LighterGlueMatcher._match_pairs
 * (matchers/lighterglue.py:105-262). */
int dimb_lg_match(dimb_lg* lg, int P, const dimb_feats* f0, const dimb_feats* f1, int64_t* matches, float* mscores,
                  int* n_matches, int* stop_layer, int cap);

typedef struct {
  const float* keypoints;   /* device (n_cap,2) */
  const float* descriptors; /* device; layout below */
  const int* n;             /* device scalar: number of valid keypoints (<= n_cap) */
  int n_cap;
  int desc_layout;          /* 0: (D,n) rows of pitch desc_ld, 1: (n,D) rows of pitch desc_ld */
  int desc_ld;
  float size0, size1;
  int round_fp16;           /* 1: round keypoints/descriptors to fp16 first, as the features.h5 round trip does */
  int f16;                  /* 1: keypoints and descriptors ARE float16 arrays (device feature store blocks); round_fp16 is moot */
  const int* size_dev;      /* non-NULL: device int[2] holding image_size ([H,W]); overrides size0 / size1 */
} dimb_feats_dev;

/* Device-resident variant, asynchronous on `stream`; d_matches [P][cap][2] int64, d_mscores [P][cap],
 * d_n_matches [P], d_stop_layer [P] are device buffers.  Only for the tensor-core shape (descriptor_dim 256, 4 heads):
 * DIMB_ERR_UNSUPPORTED otherwise. */
int dimb_lg_match_dev(dimb_lg* lg, int P, const dimb_feats_dev* f0, const dimb_feats_dev* f1, int64_t* d_matches,
                      float* d_mscores, int* d_n_matches, int* d_stop_layer, int cap, void* stream);
/* Debug tap: fp32 descriptors x[side][row][d] after the last executed layer of the LAST call (host copy). */
int dimb_lg_debug_read(dimb_lg* lg, int which, int side, float* out, size_t n_floats);

/* ------------------------------------------------------------------ brute-force descriptor NN */
enum { DIMB_NN_NN = 0, DIMB_NN_MNN = 1, DIMB_NN_SNN = 2, DIMB_NN_SMNN = 3 };

/* d0: (D,n0) float32 host (FeaturesDict layout), d1: (D,n1); any descriptor size D >= 1 (zero-padded to a multiple of 64 on
 * device, which changes no distance).  Outputs: idx [cap][2] int64 sorted by column 0, dist [cap] (distance for nn/mnn, ratio
 * for snn/smnn), n = number of matches.  Descriptors that are exactly fp16-representable (everything read back from
 * features.h5 is, extractor_base.py:56-99) are detected on device and take the single-MMA path, which is then exact. */
int dimb_nn_match(dimb_ctx* ctx, const float* d0, int n0, const float* d1, int n1, int D, int mode, float th,
                  int64_t* idx, float* dist, int* n, int cap);
/* Same on device pointers, asynchronous on `stream`: d_desc0 / d_desc1 are (D,n) arrays of row pitch ld0 / ld1 elements
 * (0 = dense), float32 (desc_f16 = 0) or float16 (desc_f16 = 1: the layout the device feature store keeps, exact single-MMA
 * path); d_idx [cap][2] int64, d_dist [cap], d_n [1] are device buffers.  The sequential-pair workload of
 * pairs_generator.py:22-34 keeps every image's descriptors in HBM and calls this once per pair. */
int dimb_nn_match_dev(dimb_ctx* ctx, const void* d_desc0, int n0, int ld0, const void* d_desc1, int n1, int ld1, int D, int desc_f16,
                      int mode, float th, int64_t* d_idx, float* d_dist, int* d_n, int cap, void* stream);

/* ------------------------------------------------------------------ device feature store (the features.h5 boundary kept in HBM)
 * Replaces, for the hot path, save_features_h5 (extractors/extractor_base.py:56-99: every array cast to float16, gzip-9, one
 * group per image) and get_features (io/h5.py:45-89: re-read per image per pair, matchers/matcher_base.py:221-222).  One
 * fixed-size block per image: int32 header {n, H, W, valid}, then float16 keypoints [cap][2], scores [cap], tile_idx [cap],
 * descriptors [D][cap] - the exact values features.h5 would hold.  Blocks are contiguous so that the multi-GPU path can
 * all-gather them over NCCL and an h5 writer needs one bulk copy. */
typedef struct dimb_fstore dimb_fstore;
int dimb_fstore_create(dimb_ctx* ctx, int n_slots, int cap, int desc_dim, dimb_fstore** out);
void dimb_fstore_destroy(dimb_fstore* fs);
/* Device put (asynchronous on `stream`): float32 features in the layouts of dimb_sp_extract_dev / dimb_aliked_extract_dev
 * (d_desc (D,n) rows of pitch desc_ld, d_count device scalar); d_scores / d_tile_idx may be NULL (ones / zeros, as
 * ExtractorBase.extract fills them, extractor_base.py:226,371-373).  The float16 cast of the h5 writer happens here. */
int dimb_fstore_put_dev(dimb_fstore* fs, int slot, const float* d_kpts, const float* d_scores, const float* d_tile_idx, const float* d_desc,
                        int desc_ld, const int* d_count, int height, int width, void* stream);
/* Host put: kpts (n,2), scores (n,) or NULL, tile_idx (n,) or NULL, desc (D,n) dense. */
int dimb_fstore_put(dimb_fstore* fs, int slot, const float* kpts, const float* scores, const float* tile_idx, const float* desc, int n,
                    int height, int width);
/* n = keypoints stored in the slot (-1: empty); image_size[2] = [H,W].  Synchronises. */
int dimb_fstore_count(dimb_fstore* fs, int slot, int* n, int* image_size);
/* get_features' contract: float32 host arrays whose values are float16-exact, desc (D,n) dense; any output may be NULL. */
int dimb_fstore_get(dimb_fstore* fs, int slot, float* kpts, float* scores, float* tile_idx, float* desc, int* n, int* image_size, int cap);
/* The slot as a dimb_feats_dev with f16 = 1, for dimb_lg_match_dev; its descriptors also feed dimb_nn_match_dev (desc_f16 = 1, ld = cap). */
int dimb_fstore_feats_dev(dimb_fstore* fs, int slot, dimb_feats_dev* out);
/* Raw blocks: base pointer, bytes per slot, slot count, keypoint capacity (slot s starts at base + s * slot_bytes). */
int dimb_fstore_block_dev(dimb_fstore* fs, void** d_base, size_t* slot_bytes, int* n_slots, int* cap);

/* ------------------------------------------------------------------ geometric verification (fundamental-matrix RANSAC)
 * Replaces the estimator inside geometric_verification (utils/geometric_verification.py:45-179: pydegensac.findFundamentalMatrix /
 * cv2.findFundamentalMat), the step after _match_pairs (matchers/matcher_base.py:298-340).  min(max_iters, 8192) 8-point
 * hypotheses per pair in parallel, Sampson inliers (threshold in pixels), two least-squares refits of the best model.  Stochastic
 * like the reference's estimators (seeded, reproducible here): parity is statistical.  F row-major, x1^T F x0 = 0; zeros and an
 * all-ones mask when fewer than 8 matches exist or no model is found (the reference returns F = None, mask all True). */
int dimb_gv_fundamental(dimb_ctx* ctx, const float* kpts0, const float* kpts1, int n, float threshold, int max_iters, unsigned seed, float* F,
                        unsigned char* mask, int* n_inliers);
/* P pairs on device buffers, asynchronous on `stream`: matches in the output layout of dimb_lg_match_dev / dimb_pipe_* ([P][cap][2]
 * int64 + [P] counts) indexing the per-pair keypoint arrays d_kpts0[p] / d_kpts1[p] ((N,2) float32; the pointer ARRAYS are host). */
int dimb_gv_fundamental_batch_dev(dimb_ctx* ctx, int P, const float* const* d_kpts0, const float* const* d_kpts1, const int64_t* d_matches,
                                  const int* d_n_matches, int cap, float threshold, int max_iters, unsigned seed, float* d_F,
                                  unsigned char* d_mask, int* d_n_inliers, void* stream);

/* ------------------------------------------------------------------ fused per-pair path
 * SuperPoint on both images of every pair followed by LightGlue, features kept in HBM in between (the
 * reference's features.h5 round trip, ImageMatcher.extract_features -> match_pairs, image_matching.py:413-494,
 * is reduced to its value-level effect: fp16 rounding of keypoints and descriptors). */
typedef struct dimb_pipe dimb_pipe;
int dimb_pipe_create(dimb_sp* sp, dimb_lg* lg, int max_pairs, int H, int W, int cap, dimb_pipe** out);
void dimb_pipe_destroy(dimb_pipe* pipe);
/* images: HOST float32 [2P][H][W] gray 0..255, pair p = images 2p and 2p+1.  Outputs (HOST): matches [P][cap][2]
 * int64, mscores [P][cap], n_matches [P], stop_layer [P], n_kpts [2P], kpts [2P][cap][2] (may be NULL). */
int dimb_pipe_match_image_pairs(dimb_pipe* pipe, const float* images, int P, int64_t* matches, float* mscores,
                                int* n_matches, int* stop_layer, int* n_kpts, float* kpts);
/* Same with 8-bit gray images (the reference casts them with astype(float32), extractor_base.py:201-202): a quarter of
 * the host->device traffic, exact conversion on device. */
int dimb_pipe_match_image_pairs_u8(dimb_pipe* pipe, const uint8_t* images, int P, int64_t* matches, float* mscores,
                                   int* n_matches, int* stop_layer, int* n_kpts, float* kpts);
/* Same with the images already in device memory, asynchronous on `stream`; results stay in device buffers owned
 * by the pipe, exposed by dimb_pipe_outputs_dev (layouts as above). */
int dimb_pipe_match_image_pairs_dev(dimb_pipe* pipe, const float* d_images, int P, void* stream);
int dimb_pipe_outputs_dev(dimb_pipe* pipe, int64_t** d_matches, float** d_mscores, int** d_n_matches, int** d_stop,
                          int** d_nkpts, float** d_kpts);
/* Device buffers holding the SuperPoint features of the last call: kpts [2P][cap][2], scores [2P][cap], desc [2P][256][cap]
 * ((D,N) rows of pitch cap), counts [2P] - the arrays ExtractorBase.extract would hand to save_features_h5
 * (extractor_base.py:223-229) before the float16 cast. */
int dimb_pipe_features_dev(dimb_pipe* pipe, float** d_kpts, float** d_scores, float** d_desc, int** d_counts);
dimb_ctx* dimb_sp_ctx(dimb_sp* sp);

/* ---------------------------------------------------------------------------------------------------------
 * ALIKED extraction.  Replaces AlikedExtractor._extract (reference src/deep_image_matching/extractors/aliked.py:45-64)
 * and the model it drives (thirdparty/LightGlue/lightglue/aliked.py:560-693: encoder with deformable blocks :367-449,
 * DKD detector :92-244, SDDH descriptor head :452-558).  Supported: aliked-n16 / aliked-n16rot (dim 128, K 3, M 16),
 * threshold detection mode (detection_threshold > 0, the reference's only configured mode).
 *
 * weights: fp32 blob, the model's state_dict tensors in state_dict order without num_batches_tracked
 * (block1.conv1.weight ... desc_head.sf_conv.weight; 678316 floats).
 * Reproduced quirk: `scores` are the DKD score *dispersities* (aliked.py:682 swaps the names). */
typedef struct dimb_aliked dimb_aliked;
typedef struct dimb_aliked_conf {
  int max_num_keypoints;      /* n_limit; <= 0 -> 20000 (aliked.py:585) */
  float detection_threshold;  /* 0.2 */
  int nms_radius;             /* 2 */
  int max_height, max_width;  /* workspace size */
} dimb_aliked_conf;
int dimb_aliked_create(dimb_ctx* ctx, const float* weights, size_t n_floats, const dimb_aliked_conf* conf, dimb_aliked** out);
void dimb_aliked_destroy(dimb_aliked* al);
/* image: host fp32 (H,W,channels) 0..255, channels 3 (RGB, ExtractorBase with grayscale=False) or 1 (replicated).
 * Out (host): kpts [cap][2] sub-pixel (x,y); scores [cap]; desc [128][cap] ((D,N) FeaturesDict layout, ld = cap); count. */
int dimb_aliked_extract(dimb_aliked* al, const float* image, int H, int W, int channels, float* kpts, float* scores, float* desc,
                        int* count, int cap);
/* Device-pointer variant (same layouts, all pointers in device memory, count [1]); no host synchronisation - the caller
 * checks count <= cap.  Feeds dimb_lg_match_dev (desc_layout 0, desc_ld = cap) without leaving HBM. */
int dimb_aliked_extract_dev(dimb_aliked* al, const float* d_image, int H, int W, int channels, float* d_kpts, float* d_scores,
                            float* d_desc, int* d_count, int cap, void* stream);
/* debug taps of the last call: 0 = score map [H][W], 1 = L2-normalised feature map [128][H][W] */
int dimb_aliked_debug_read(dimb_aliked* al, int which, float* out, size_t n_floats);

/* ---------------------------------------------------------------------------------------------------------
 * SuperGlue matching.  Replaces SuperGlueMatcher._match_pairs (reference src/deep_image_matching/matchers/superglue.py:75-106,
 * adapter features_2_sg :8-41) and the model it drives (thirdparty/SuperGluePretrainedNetwork/models/superglue.py:51-305).
 * descriptor_dim 256, 4 heads, keypoint encoder [32,64,128,256].  First cut on the plain fp32 kernels (csrc/superglue.cu).
 *
 * weights: fp32 blob in THIS order (names of the reference state_dict; BatchNorm = weight, bias, running_mean, running_var):
 *   kenc.encoder.{0,3,6,9}.{weight,bias} each followed by its BatchNorm kenc.encoder.{1,4,7,10}; kenc.encoder.12.{weight,bias};
 *   for i in layers: gnn.layers.i.attn.merge.{weight,bias}, attn.proj.{0,1,2}.{weight,bias}, mlp.0.{weight,bias}, mlp.1 (BatchNorm),
 *                    mlp.3.{weight,bias};
 *   final_proj.{weight,bias}; bin_score. */
typedef struct dimb_sg dimb_sg;
typedef struct {
  int n_layers;                   /* 18 */
  unsigned long long cross_mask;  /* bit i set: GNN layer i is a cross layer (["self","cross"] * 9 -> 0x2AAAA) */
  int sinkhorn_iterations;        /* 100 (superglue.py:218; the plugin's own 20 never reaches the model, see matchers/superglue.py) */
  float match_threshold;          /* 0.2 */
  int max_kpts;                   /* workspace sizing */
} dimb_sg_conf;
typedef struct {
  const float* keypoints;   /* (n,2) x,y */
  const float* descriptors; /* (256,n) rows of pitch desc_ld (0 = n): the FeaturesDict layout */
  const float* scores;      /* (n,) */
  int n, desc_ld;
  int height, width;        /* image_size = [H,W] */
} dimb_sg_feats;
size_t dimb_sg_weight_count(int n_layers);
int dimb_sg_create(dimb_ctx* ctx, const float* weights, size_t n_floats, const dimb_sg_conf* conf, dimb_sg** out);
void dimb_sg_destroy(dimb_sg* sg);
/* One pair.  Out (host): matches [cap][2] int64 ascending in column 0, mscores [cap] (matching_scores0 of the matched rows). */
int dimb_sg_match(dimb_sg* sg, const dimb_sg_feats* f0, const dimb_sg_feats* f1, int64_t* matches, float* mscores, int* n_matches,
                  int cap);

#ifdef __cplusplus
}
#endif
#endif /* DIMB200_H */
