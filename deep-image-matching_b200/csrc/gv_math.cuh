// gv_math.cuh - the arithmetic of fundamental-matrix RANSAC, written once for host and device: the CUDA kernels of gv.cu call it per
// thread, the self-test library drives the very same functions on the CPU (tests without a GPU).
// Replaces the estimator inside the reference's geometric_verification (utils/geometric_verification.py:45-179: pydegensac /
// OpenCV findFundamentalMat); RANSAC is stochastic, so parity is statistical (inlier sets on data with known geometry).
#pragma once
#include <math.h>
#include <stdint.h>
#ifndef __CUDACC__
#define __host__
#define __device__
#endif

namespace gv {

struct Norm {  // Hartley normalisation x' = s (x - c)
  float cx, cy, s;
};

__host__ __device__ inline uint32_t hash3(uint32_t a, uint32_t b, uint32_t c) {  // counter-based RNG (no state to carry)
  uint32_t h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du;
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 12;
  h *= 0x297A2D39u;
  h ^= h >> 15;
  return h;
}

// 8 distinct indices in [0, n) for hypothesis `hyp` (n >= 8)
__host__ __device__ inline void sample8(uint32_t seed, uint32_t hyp, int n, int idx[8]) {
  uint32_t ctr = 0;
  for (int k = 0; k < 8; ++k) {
    while (true) {
      const int c = static_cast<int>(hash3(seed, hyp, ctr++) % static_cast<uint32_t>(n));
      bool dup = false;
      for (int j = 0; j < k; ++j) dup |= idx[j] == c;
      if (!dup) {
        idx[k] = c;
        break;
      }
    }
  }
}

// Null vector of an 8 x 9 system by Gauss-Jordan elimination with partial (row) pivoting and a free column chosen as the worst
// pivot column: returns false for (near-)degenerate samples.  a: row-major [8][9], destroyed.
__host__ __device__ inline bool null9(float a[8][9], float f[9]) {
  int piv_col[8];
  bool used[9] = {false, false, false, false, false, false, false, false, false};
  for (int r = 0; r < 8; ++r) {
    // pivot = largest |a[i][c]| over rows i >= r and unused columns c
    int pr = r, pc = -1;
    float best = 0.f;
    for (int i = r; i < 8; ++i)
      for (int c = 0; c < 9; ++c)
        if (!used[c] && fabsf(a[i][c]) > best) best = fabsf(a[i][c]), pr = i, pc = c;
    if (pc < 0 || best < 1e-7f) return false;
    if (pr != r)
      for (int c = 0; c < 9; ++c) {
        const float t = a[r][c];
        a[r][c] = a[pr][c];
        a[pr][c] = t;
      }
    used[pc] = true;
    piv_col[r] = pc;
    const float inv = 1.f / a[r][pc];
    for (int c = 0; c < 9; ++c) a[r][c] *= inv;
    for (int i = 0; i < 8; ++i)
      if (i != r) {
        const float m = a[i][pc];
        if (m != 0.f)
          for (int c = 0; c < 9; ++c) a[i][c] -= m * a[r][c];
      }
  }
  int fc = 0;
  while (used[fc]) ++fc;  // the free column
  f[fc] = 1.f;
  for (int r = 0; r < 8; ++r) f[piv_col[r]] = -a[r][fc];
  float nrm = 0.f;
  for (int c = 0; c < 9; ++c) nrm += f[c] * f[c];
  nrm = 1.f / sqrtf(nrm);
  for (int c = 0; c < 9; ++c) f[c] *= nrm;
  return true;
}

// symmetric 3x3 eigen-decomposition by cyclic Jacobi: A = V diag(w) V^T (A destroyed; V columns = eigenvectors)
__host__ __device__ inline void jacobi3(float A[3][3], float V[3][3], float w[3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = i == j ? 1.f : 0.f;
  for (int sweep = 0; sweep < 12; ++sweep) {
    const float off = fabsf(A[0][1]) + fabsf(A[0][2]) + fabsf(A[1][2]);
    if (off < 1e-12f) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (fabsf(A[p][q]) < 1e-20f) continue;
        const float theta = (A[q][q] - A[p][p]) / (2.f * A[p][q]);
        const float t = (theta >= 0.f ? 1.f : -1.f) / (fabsf(theta) + sqrtf(theta * theta + 1.f));
        const float c = 1.f / sqrtf(t * t + 1.f), s = t * c;
        for (int k = 0; k < 3; ++k) {  // A <- A J
          const float akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {  // A <- J^T A
          const float apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const float vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < 3; ++i) w[i] = A[i][i];
}

// rank-2 projection: F <- F - (F v)(v^T) with v the right singular vector of the smallest singular value (= the closest rank-2
// matrix in Frobenius norm, what the SVD clamp of the 8-point algorithm computes)
__host__ __device__ inline void rank2(float F[9]) {
  float M[3][3], V[3][3], w[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M[i][j] = F[0 + i] * F[0 + j] + F[3 + i] * F[3 + j] + F[6 + i] * F[6 + j];  // F^T F
  jacobi3(M, V, w);
  int k = 0;
  if (w[1] < w[k]) k = 1;
  if (w[2] < w[k]) k = 2;
  const float v[3] = {V[0][k], V[1][k], V[2][k]};
  for (int r = 0; r < 3; ++r) {
    const float fv = F[3 * r] * v[0] + F[3 * r + 1] * v[1] + F[3 * r + 2] * v[2];
    for (int c = 0; c < 3; ++c) F[3 * r + c] -= fv * v[c];
  }
}

// normalised 8-point algorithm on 8 correspondences given in ORIGINAL pixels; F maps image 0 -> epipolar lines of image 1
// (x1^T F x0 = 0).  n0 / n1: Hartley normalisations of the whole match set.
__host__ __device__ inline bool eight_point(const float* k0, const float* k1, const int idx[8], Norm n0, Norm n1, float F[9]) {
  float a[8][9];
  for (int r = 0; r < 8; ++r) {
    const float x0 = (k0[2 * idx[r]] - n0.cx) * n0.s, y0 = (k0[2 * idx[r] + 1] - n0.cy) * n0.s;
    const float x1 = (k1[2 * idx[r]] - n1.cx) * n1.s, y1 = (k1[2 * idx[r] + 1] - n1.cy) * n1.s;
    a[r][0] = x1 * x0, a[r][1] = x1 * y0, a[r][2] = x1;
    a[r][3] = y1 * x0, a[r][4] = y1 * y0, a[r][5] = y1;
    a[r][6] = x0, a[r][7] = y0, a[r][8] = 1.f;
  }
  float f[9];
  if (!null9(a, f)) return false;
  rank2(f);
  // denormalise: F = T1^T Fn T0 with T = [[s,0,-s cx],[0,s,-s cy],[0,0,1]]
  float G[9];  // Fn T0
  for (int r = 0; r < 3; ++r) {
    G[3 * r] = f[3 * r] * n0.s;
    G[3 * r + 1] = f[3 * r + 1] * n0.s;
    G[3 * r + 2] = f[3 * r + 2] - n0.s * (f[3 * r] * n0.cx + f[3 * r + 1] * n0.cy);
  }
  for (int c = 0; c < 3; ++c) {
    F[c] = n1.s * G[c];
    F[3 + c] = n1.s * G[3 + c];
    F[6 + c] = G[6 + c] - n1.s * (n1.cx * G[c] + n1.cy * G[3 + c]);
  }
  float nrm = 0.f;
  for (int c = 0; c < 9; ++c) nrm += F[c] * F[c];
  if (!(nrm > 0.f)) return false;
  nrm = 1.f / sqrtf(nrm);
  for (int c = 0; c < 9; ++c) F[c] *= nrm;
  return true;
}

// squared Sampson distance of a correspondence (first-order geometric error, pixels^2)
__host__ __device__ inline float sampson2(const float F[9], float x0, float y0, float x1, float y1) {
  const float l0 = F[0] * x0 + F[1] * y0 + F[2], l1 = F[3] * x0 + F[4] * y0 + F[5], l2 = F[6] * x0 + F[7] * y0 + F[8];  // F x0
  const float m0 = F[0] * x1 + F[3] * y1 + F[6], m1 = F[1] * x1 + F[4] * y1 + F[7];                                      // F^T x1
  const float e = x1 * l0 + y1 * l1 + l2;
  const float d = l0 * l0 + l1 * l1 + m0 * m0 + m1 * m1;
  return d > 0.f ? e * e / d : 3.4e38f;
}

// least-squares refit on a set of correspondences: smallest eigenvector of the 9x9 normal matrix (cyclic Jacobi), rank 2, denormalise.
// N: upper-triangular-complete symmetric 9x9 sum of a^T a over the NORMALISED inlier correspondences.
__host__ __device__ inline bool refit_from_normal(float N[9][9], Norm n0, Norm n1, float F[9]) {
  float V[9][9];
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 9; ++j) V[i][j] = i == j ? 1.f : 0.f;
  for (int sweep = 0; sweep < 30; ++sweep) {
    float off = 0.f;
    for (int p = 0; p < 8; ++p)
      for (int q = p + 1; q < 9; ++q) off += fabsf(N[p][q]);
    if (off < 1e-9f) break;
    for (int p = 0; p < 8; ++p)
      for (int q = p + 1; q < 9; ++q) {
        if (fabsf(N[p][q]) < 1e-30f) continue;
        const float theta = (N[q][q] - N[p][p]) / (2.f * N[p][q]);
        const float t = (theta >= 0.f ? 1.f : -1.f) / (fabsf(theta) + sqrtf(theta * theta + 1.f));
        const float c = 1.f / sqrtf(t * t + 1.f), s = t * c;
        for (int k = 0; k < 9; ++k) {
          const float akp = N[k][p], akq = N[k][q];
          N[k][p] = c * akp - s * akq;
          N[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 9; ++k) {
          const float apk = N[p][k], aqk = N[q][k];
          N[p][k] = c * apk - s * aqk;
          N[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 9; ++k) {
          const float vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int k = 0;
  for (int i = 1; i < 9; ++i)
    if (N[i][i] < N[k][k]) k = i;
  float f[9];
  for (int i = 0; i < 9; ++i) f[i] = V[i][k];
  rank2(f);
  float G[9];
  for (int r = 0; r < 3; ++r) {
    G[3 * r] = f[3 * r] * n0.s;
    G[3 * r + 1] = f[3 * r + 1] * n0.s;
    G[3 * r + 2] = f[3 * r + 2] - n0.s * (f[3 * r] * n0.cx + f[3 * r + 1] * n0.cy);
  }
  for (int c = 0; c < 3; ++c) {
    F[c] = n1.s * G[c];
    F[3 + c] = n1.s * G[3 + c];
    F[6 + c] = G[6 + c] - n1.s * (n1.cx * G[c] + n1.cy * G[3 + c]);
  }
  float nrm = 0.f;
  for (int c = 0; c < 9; ++c) nrm += F[c] * F[c];
  if (!(nrm > 0.f)) return false;
  nrm = 1.f / sqrtf(nrm);
  for (int c = 0; c < 9; ++c) F[c] *= nrm;
  return true;
}

}  // namespace gv
