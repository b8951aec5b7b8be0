// lightglue_generic.cuh - interface of the shape-generic fp32 LightGlue (lightglue_generic.cu) used by dimb_lg_* for
// configurations other than descriptor_dim 256 / 4 heads (e.g. the LighterGlue checkpoint: 96 / 1 head / 6 layers).
#pragma once
#include "common.cuh"

struct dimb_lgx;
int lgx_create(dimb_ctx* ctx, const float* weights, size_t n_floats, const dimb_lg_conf* conf, dimb_lgx** out);
void lgx_destroy(dimb_lgx* g);
int lgx_match(dimb_lgx* g, int P, const dimb_feats* f0, const dimb_feats* f1, int64_t* matches, float* mscores, int* n_matches,
              int* stop_layer, int cap);
