/*
 * frontend_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's per-frame-pair hot path
 * (felixendres/rgbdslam_v2 @ d345352f), used ONLY by tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() as the
 * checker for the CUDA path.  Nothing under rgbdslam_v2_b200/ links, imports
 * or calls this file.
 *
 * PARITY STATUS: "parity unpinned" for everything except oracle_brute_force_
 * search_orb -- the reference ships no golden vectors / known-answer tests for
 * this path and cannot be compiled here (ROS / Qt /
 * PCL / Eigen / g2o / OpenCV-C++ absent).  bruteForceSearchORB is the one
 * function that compiles from the reference tree as-is; oracle/Makefile builds
 * it from /root/reference/src/features.cpp:163-182 into oracle/_ref/ and
 * tests/test_oracle.py checks this restatement against it bit for bit.
 *
 * Each function cites the reference file:line it follows.  Where the reference
 * delegates to an un-vendored library the published algorithm is restated:
 *   - pcl::TransformationFromCorrespondences (PCL 1.7, common/
 *     transformation_from_correspondences.hpp): running weighted mean /
 *     covariance + JacobiSVD, float32.
 *   - Eigen::JacobiSVD<Matrix3f> / Eigen::LLT<Matrix3d>: any correct SVD /
 *     Cholesky agrees to rounding; a Hestenes one-sided Jacobi SVD and a plain
 *     3x3 Cholesky are used.
 * The reference draws from the global rand(); the oracle uses the counter-based
 * generator documented in DESIGN.md (the same stream the CUDA path uses), so a
 * given (seed, pair, hypothesis) selects the same sample on both sides.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_MAX_MATCHES_CAP 512

typedef struct {
  int32_t queryIdx, trainIdx, imgIdx;
  float distance;
} oracle_dmatch; /* == cv::DMatch */

typedef struct {
  int32_t min_matches;       /* parameter_server.cpp:85  */
  int32_t max_matches;       /* :86 */
  int32_t ransac_iterations; /* :101 */
  int32_t pad_;
  double max_dist_for_inliers; /* :100 */
  double sigma_depth;          /* :46 */
  double depth_cov_z0;         /* >0: static-cache quirk of misc2.h:30-35 with this first depth; <0: per point */
} oracle_params;

typedef struct {
  int32_t id1, id2;
  int32_t n_all_matches, n_inliers;
  float rmse;
  int32_t valid_iterations;
  float ransac_trafo[16]; /* column-major Matrix4f */
  double info_scale;
  int32_t used_identity;
  int32_t real_iterations;
} oracle_pair_result;

/* ------------------------------------------------------------------------ */
/* counter-based RNG (DESIGN.md "Random numbers")                            */

static inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

/* 31-bit value standing in for one rand() call (RAND_MAX = 2^31-1). */
uint32_t oracle_rand31(uint64_t seed, uint64_t pair, uint32_t stream, uint32_t ctr) {
  uint64_t k = mix64(seed ^ mix64(pair));
  return (uint32_t)(mix64(k ^ (((uint64_t)stream << 32) | ctr)) >> 33);
}

/* ------------------------------------------------------------------------ */
/* features.cpp:163-166                                                      */
static inline int hamming256(const uint64_t* a, const uint64_t* b) {
  return (__builtin_popcountll(a[0] ^ b[0]) + __builtin_popcountll(a[1] ^ b[1])) +
         (__builtin_popcountll(a[2] ^ b[2]) + __builtin_popcountll(a[3] ^ b[3]));
}

/* features.cpp:168-182.  NB the loop bound `i < size-1` on unsigned size: the last
 * train row is never examined; size==0 would underflow in the reference (it walks
 * off the array) -- the oracle treats size<=1 as "no candidate" (257, -1). */
int oracle_brute_force_search_orb(const uint64_t* v, const uint64_t* search_array, unsigned size,
                                  int* result_index) {
  *result_index = -1;
  int min_distance = 1 + 256;
  if (size == 0) return min_distance;
  for (unsigned i = 0; i < size - 1; i++, search_array += 4) {
    int d = hamming256(v, search_array);
    if (d < min_distance) {
      min_distance = d;
      *result_index = (int)i;
    }
  }
  return min_distance;
}

/* nq independent searches: the loop of node.cpp:567-575 without the filter. */
void oracle_brute_force_orb_batch(const uint64_t* q, int nq, const uint64_t* t, int nt, int32_t* idx,
                                  int32_t* hd) {
  for (int i = 0; i < nq; i++) {
    int r;
    hd[i] = oracle_brute_force_search_orb(q + 4 * (size_t)i, t, (unsigned)nt, &r);
    idx[i] = r;
  }
}

/* ------------------------------------------------------------------------ */
/* node.cpp:561-576 (ORB branch of featureMatching) + keepStrongestMatches
 * (node.cpp:519-531, called at :674) + the std::sort of node.cpp:1127.
 * Deterministic tie rule: (distance, queryIdx) ascending.
 * distance = hd/256.0 + (float)rand()/(1000.0*RAND_MAX)  (node.cpp:573).      */

static int cmp_match(const void* a, const void* b) {
  const oracle_dmatch* x = (const oracle_dmatch*)a;
  const oracle_dmatch* y = (const oracle_dmatch*)b;
  if (x->distance < y->distance) return -1;
  if (x->distance > y->distance) return 1;
  return (x->queryIdx > y->queryIdx) - (x->queryIdx < y->queryIdx);
}

float oracle_match_distance(int hd, uint32_t r31) {
  double d = (double)hd / 256.0 + (double)(float)r31 / (1000.0 * 2147483647.0);
  return (float)d;
}

int oracle_feature_matching_orb(const uint64_t* q, int nq, const uint64_t* t, int nt, int max_matches,
                                uint64_t seed, uint64_t pair, oracle_dmatch* out /* >= nq entries */) {
  int n = 0;
  for (int i = 0; i < nq; i++) {
    int idx;
    int hd = oracle_brute_force_search_orb(q + 4 * (size_t)i, t, (unsigned)nt, &idx);
    if (hd >= 128) continue; /* node.cpp:572 */
    out[n].queryIdx = i;
    out[n].trainIdx = idx;
    out[n].imgIdx = -1;
    out[n].distance = oracle_match_distance(hd, oracle_rand31(seed, pair, 0u, (uint32_t)i));
    n++;
  }
  qsort(out, (size_t)n, sizeof(oracle_dmatch), cmp_match);
  if (n > max_matches) n = max_matches; /* keepStrongestMatches */
  return n;
}

/* ------------------------------------------------------------------------ */
/* 3x3 float SVD (stand-in for Eigen::JacobiSVD<Matrix3f>): Hestenes one-sided
 * Jacobi, singular values sorted descending.  a, u, v are row-major.          */
static void svd3f(const float a_in[9], float u[9], float s[3], float v[9]) {
  float a[9];
  memcpy(a, a_in, sizeof(a));
  for (int i = 0; i < 9; i++) v[i] = (i % 4 == 0) ? 1.f : 0.f;
  for (int sweep = 0; sweep < 12; sweep++) {
    int rotated = 0;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        float alpha = 0, beta = 0, gamma = 0;
        for (int r = 0; r < 3; r++) {
          alpha += a[3 * r + p] * a[3 * r + p];
          beta += a[3 * r + q] * a[3 * r + q];
          gamma += a[3 * r + p] * a[3 * r + q];
        }
        if (fabsf(gamma) <= 1e-9f * sqrtf(alpha * beta) || gamma == 0.f) continue;
        rotated = 1;
        float zeta = (beta - alpha) / (2.f * gamma);
        float tt = (zeta >= 0 ? 1.f : -1.f) / (fabsf(zeta) + sqrtf(1.f + zeta * zeta));
        float c = 1.f / sqrtf(1.f + tt * tt), sn = c * tt;
        for (int r = 0; r < 3; r++) {
          float ap = a[3 * r + p], aq = a[3 * r + q];
          a[3 * r + p] = c * ap - sn * aq;
          a[3 * r + q] = sn * ap + c * aq;
          float vp = v[3 * r + p], vq = v[3 * r + q];
          v[3 * r + p] = c * vp - sn * vq;
          v[3 * r + q] = sn * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  float nrm[3];
  for (int j = 0; j < 3; j++)
    nrm[j] = sqrtf(a[j] * a[j] + a[3 + j] * a[3 + j] + a[6 + j] * a[6 + j]);
  /* sort columns descending */
  int ord[3] = {0, 1, 2};
  for (int i = 0; i < 2; i++)
    for (int j = i + 1; j < 3; j++)
      if (nrm[ord[j]] > nrm[ord[i]]) {
        int tmp = ord[i];
        ord[i] = ord[j];
        ord[j] = tmp;
      }
  float as[9], vs[9];
  for (int j = 0; j < 3; j++) {
    s[j] = nrm[ord[j]];
    for (int r = 0; r < 3; r++) {
      as[3 * r + j] = a[3 * r + ord[j]];
      vs[3 * r + j] = v[3 * r + ord[j]];
    }
  }
  memcpy(v, vs, sizeof(vs));
  /* U columns = normalised A V columns; complete rank-deficient columns orthonormally */
  for (int j = 0; j < 3; j++) {
    if (s[j] > 0.f && s[j] > 1e-30f) {
      for (int r = 0; r < 3; r++) u[3 * r + j] = as[3 * r + j] / s[j];
    } else {
      for (int r = 0; r < 3; r++) u[3 * r + j] = 0.f;
    }
  }
  /* rank 2: third column = u0 x u1 (a full orthogonal U, as ComputeFullU gives) */
  if (!(s[2] > 1e-12f * s[0])) {
    u[2] = u[3] * u[7] - u[6] * u[4];
    u[5] = u[6] * u[1] - u[0] * u[7];
    u[8] = u[0] * u[4] - u[3] * u[1];
  }
}

static float det3f(const float m[9]) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
         m[2] * (m[3] * m[7] - m[4] * m[6]);
}

/* transformation_estimation_euclidean.cpp:7-61 + pcl::TransformationFromCorrespondences
 * (add(): running weighted mean/covariance; getTransformation(): SVD, R = U S V^T,
 * t = mean2 - R mean1).  Points are Vector4f (x,y,z,1); T out is column-major 4x4.
 * sel: indices into matches[] of the correspondences to use, in that order. */
void oracle_get_transform_from_matches(const float* xyz_newer, const float* xyz_earlier,
                                       const oracle_dmatch* matches, const int* sel, int nsel,
                                       float T[16]) {
  float W = 0.f;
  float m1[3] = {0, 0, 0}, m2[3] = {0, 0, 0};
  float C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; /* row-major, C += d2 * d1^T */
  for (int k = 0; k < nsel; k++) {
    const oracle_dmatch* m = &matches[sel ? sel[k] : k];
    const float* from = xyz_newer + 4 * (size_t)m->queryIdx;
    const float* to = xyz_earlier + 4 * (size_t)m->trainIdx;
    if (isnan(from[2]) || isnan(to[2])) continue; /* :22 */
    float weight = (float)(1.0 / (double)(from[2] * to[2])); /* :25 (1.0 is a double literal) */
    if (weight == 0.f) continue;                              /* pcl add(): if (weight==0) return */
    W += weight;
    float alpha = weight / W;
    float d1[3], d2[3];
    for (int r = 0; r < 3; r++) {
      d1[r] = from[r] - m1[r];
      d2[r] = to[r] - m2[r];
    }
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) C[3 * r + c] = (1.0f - alpha) * (C[3 * r + c] + alpha * (d2[r] * d1[c]));
    for (int r = 0; r < 3; r++) {
      m1[r] += alpha * d1[r];
      m2[r] += alpha * d2[r];
    }
  }
  float U[9], S[3], V[9];
  svd3f(C, U, S, V);
  float sgn = (det3f(U) * det3f(V) < 0.0f) ? -1.f : 1.f;
  float R[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++)
      R[3 * r + c] = U[3 * r + 0] * V[3 * c + 0] + U[3 * r + 1] * V[3 * c + 1] + sgn * U[3 * r + 2] * V[3 * c + 2];
  float t[3];
  for (int r = 0; r < 3; r++) t[r] = m2[r] - (R[3 * r + 0] * m1[0] + R[3 * r + 1] * m1[1] + R[3 * r + 2] * m1[2]);
  for (int c = 0; c < 3; c++) {
    for (int r = 0; r < 3; r++) T[4 * c + r] = R[3 * r + c];
    T[4 * c + 3] = 0.f;
  }
  T[12] = t[0];
  T[13] = t[1];
  T[14] = t[2];
  T[15] = 1.f;
}

/* ------------------------------------------------------------------------ */
/* misc2.h:20-35 with the static-cache quirk made explicit. */
static inline double depth_cov(const oracle_params* p, double z) {
  double zz = p->depth_cov_z0 > 0 ? p->depth_cov_z0 : z;
  double sd = p->sigma_depth * zz * zz;
  return sd * sd;
}

/* misc.cpp:697-770 (errorFunction2).  T is the column-major double copy of the
 * float transform (node.cpp:984 transformation4f.cast<double>()). */
double oracle_error_function2(const oracle_params* p, const float* x1, const float* x2, const double T[16]) {
  const double cam_angle_x = 58.0 / 180.0 * M_PI;
  const double cam_angle_y = 45.0 / 180.0 * M_PI;
  const double raster_stddev_x = 3 * tan(cam_angle_x / 640);
  const double raster_stddev_y = 3 * tan(cam_angle_y / 480);
  const double raster_cov_x = raster_stddev_x * raster_stddev_x;
  const double raster_cov_y = raster_stddev_y * raster_stddev_y;
  if (isnan(x1[2]) || isnan(x2[2])) return DBL_MAX;
  double a[4] = {x1[0], x1[1], x1[2], x1[3]};
  double mu2[3] = {x2[0], x2[1], x2[2]};
  double mu1in2[3];
  for (int r = 0; r < 3; r++) mu1in2[r] = T[r] * a[0] + T[4 + r] * a[1] + T[8 + r] * a[2] + T[12 + r] * a[3];
  double d[3] = {mu1in2[0] - mu2[0], mu1in2[1] - mu2[1], mu1in2[2] - mu2[2]};
  {
    double dsq = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    double s1 = fmax(raster_cov_x, depth_cov(p, a[2]));
    double s2 = fmax(raster_cov_x, depth_cov(p, mu2[2]));
    if (dsq > 2.0 * (s1 + s2)) return DBL_MAX; /* :726-735 */
  }
  double c1[3] = {raster_cov_x * a[2], raster_cov_y * a[2], depth_cov(p, a[2])};
  double c2[3] = {raster_cov_x * mu2[2], raster_cov_y * mu2[2], depth_cov(p, mu2[2])};
  /* S = R^T cov1 R + cov2  (:751-757), R(r,c) = T[4c+r] */
  double S[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double acc = 0;
      for (int k = 0; k < 3; k++) acc += T[4 * i + k] * c1[k] * T[4 * j + k];
      S[3 * i + j] = acc + (i == j ? c2[i] : 0.0);
    }
  if (isnan(d[2])) return DBL_MAX;
  /* LLT solve (:763) */
  double l00 = sqrt(S[0]);
  double l10 = S[3] / l00, l20 = S[6] / l00;
  double l11 = sqrt(S[4] - l10 * l10);
  double l21 = (S[7] - l20 * l10) / l11;
  double l22 = sqrt(S[8] - l20 * l20 - l21 * l21);
  double y0 = d[0] / l00;
  double y1 = (d[1] - l10 * y0) / l11;
  double y2 = (d[2] - l20 * y0 - l21 * y1) / l22;
  double m = y0 * y0 + y1 * y1 + y2 * y2; /* d^T S^-1 d */
  if (!(m >= 0.0)) return DBL_MAX;
  return m;
}

/* node.cpp:968-1020.  inl[i]=1 marks all_matches[i] as inlier.  Returns count. */
int oracle_compute_inliers_and_error(const oracle_params* p, const oracle_dmatch* all, int n_all,
                                     const float T4f[16], const float* origins, const float* earlier,
                                     uint8_t* inl, double* return_mean_error, double sq_max_dist) {
  double T[16];
  for (int i = 0; i < 16; i++) T[i] = (double)T4f[i];
  double mean_error = 0.0;
  int cnt = 0;
  for (int i = 0; i < n_all; i++) {
    inl[i] = 0;
    const float* o = origins + 4 * (size_t)all[i].queryIdx;
    const float* t = earlier + 4 * (size_t)all[i].trainIdx;
    if (o[2] == 0.0f || t[2] == 0.0f) continue; /* :994 */
    double md = oracle_error_function2(p, o, t, T);
    if (md > sq_max_dist) continue;
    if (!(md >= 0.0)) continue;
    mean_error += md;
    inl[i] = 1;
    cnt++;
  }
  if (cnt < 3)
    *return_mean_error = 1e9;
  else
    *return_mean_error = sqrt(mean_error / cnt);
  return cnt;
}

/* node.cpp:1024-1047: id = min(rand()%n, rand()%n) until sample_size unique ids
 * (<= 10000 tries); output ascending (std::set order).  Returns #ids. */
int oracle_sample_matches_prefer_by_distance(int sample_size, int n, uint64_t seed, uint64_t pair,
                                             uint32_t hypothesis, int* ids) {
  int cnt = 0, safety = 0;
  uint32_t ctr = 0;
  while (cnt < sample_size && n >= sample_size) {
    int id1 = (int)(oracle_rand31(seed, pair, 1u + hypothesis, ctr++) % (uint32_t)n);
    int id2 = (int)(oracle_rand31(seed, pair, 1u + hypothesis, ctr++) % (uint32_t)n);
    if (id1 > id2) id1 = id2;
    int dup = 0;
    for (int k = 0; k < cnt; k++) dup |= (ids[k] == id1);
    if (!dup) {
      int k = cnt++;
      while (k > 0 && ids[k - 1] > id1) {
        ids[k] = ids[k - 1];
        k--;
      }
      ids[k] = id1;
    }
    if (++safety > 10000) break;
  }
  return cnt;
}

static int has_nan16(const float* T) {
  for (int i = 0; i < 16; i++)
    if (T[i] != T[i]) return 1;
  return 0;
}

/* node.cpp:1074-1277 (getRelativeTransformationTo), g2o refinement off
 * (g2o_transformation_refinement=0, parameter_server.cpp:103).
 * matches: sorted initial matches (n_all).  Outputs: T (col-major), rmse,
 * inlier flags over matches[].  Returns 1 if enough inliers (node.cpp:1275). */
int oracle_get_relative_transformation_to(const oracle_params* p, const float* xyz_newer,
                                          const float* xyz_earlier, const oracle_dmatch* matches, int n_all,
                                          uint64_t seed, uint64_t pair, float T_out[16], float* rmse_out,
                                          uint8_t* inl_out, int* n_inl_out, int* valid_iterations_out,
                                          int* used_identity_out, int* real_iterations_out) {
  static const float I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  *n_inl_out = 0;
  *valid_iterations_out = 0;
  *used_identity_out = 0;
  *real_iterations_out = 0;
  memset(inl_out, 0, (size_t)n_all);
  if ((unsigned)n_all <= (unsigned)p->min_matches) return 0; /* :1087 (rmse / trafo untouched: 0 / identity) */
  memcpy(T_out, I4, sizeof(I4)); /* :1109 */
  *rmse_out = 1e6f;              /* :1110 */
  unsigned min_inlier_threshold = (unsigned)p->min_matches;
  if (min_inlier_threshold > 0.75 * n_all) min_inlier_threshold = (unsigned)(0.75 * n_all); /* :1094-1099 */
  const float max_dist_m = (float)p->max_dist_for_inliers;
  const double sq_max = (double)(max_dist_m * max_dist_m);
  const unsigned sample_size = 4;
  float rmse = 1e6f; /* float& rmse */
  int best_cnt = 0;
  unsigned valid_iterations = 0;
  int real_iterations = 0;
  uint8_t* inl = (uint8_t*)malloc((size_t)n_all);
  uint8_t* refined = (uint8_t*)malloc((size_t)n_all);
  int* sel = (int*)malloc(sizeof(int) * (size_t)n_all);

  for (int n = 0; n < p->ransac_iterations && (unsigned)n_all >= sample_size; n++) {
    double refined_error = 1e6;
    int refined_cnt = 0;
    float refined_T[16];
    memcpy(refined_T, I4, sizeof(I4));
    int nsel = oracle_sample_matches_prefer_by_distance((int)sample_size, n_all, seed, pair, (uint32_t)n, sel);
    real_iterations++;
    for (int refinements = 1; refinements < 20; refinements++) {
      float T[16];
      oracle_get_transform_from_matches(xyz_newer, xyz_earlier, matches, sel, nsel, T);
      if (has_nan16(T)) break; /* :1144 */
      double inlier_error;
      int cnt = oracle_compute_inliers_and_error(p, matches, n_all, T, xyz_newer, xyz_earlier, inl, &inlier_error, sq_max);
      nsel = 0;
      for (int i = 0; i < n_all; i++)
        if (inl[i]) sel[nsel++] = i;
      if ((unsigned)cnt < min_inlier_threshold || inlier_error > max_dist_m) break; /* :1154 */
      if (cnt >= refined_cnt && inlier_error <= refined_error) {                     /* :1160 */
        int prev = refined_cnt;
        memcpy(refined_T, T, sizeof(T));
        memcpy(refined, inl, (size_t)n_all);
        refined_cnt = cnt;
        refined_error = inlier_error;
        if (cnt == prev) break;
      } else
        break;
    }
    if (refined_cnt > 0) { /* :1170 */
      valid_iterations++;
      if (refined_error <= rmse && refined_cnt >= best_cnt && (unsigned)refined_cnt >= min_inlier_threshold) { /* :1177-1179 */
        rmse = (float)refined_error;
        memcpy(T_out, refined_T, sizeof(refined_T));
        memcpy(inl_out, refined, (size_t)n_all);
        best_cnt = refined_cnt;
        if (refined_cnt > n_all * 0.5) n += 10;  /* :1186 */
        if (refined_cnt > n_all * 0.75) n += 10; /* :1187 */
        if (refined_cnt > n_all * 0.8) break;    /* :1188 */
      }
    }
  }
  if (valid_iterations == 0) { /* :1192-1215 identity as last resort */
    double inlier_error;
    int cnt = oracle_compute_inliers_and_error(p, matches, n_all, I4, xyz_newer, xyz_earlier, inl, &inlier_error, sq_max);
    if ((unsigned)cnt > min_inlier_threshold && inlier_error < max_dist_m) {
      memcpy(T_out, I4, sizeof(I4));
      memcpy(inl_out, inl, (size_t)n_all);
      best_cnt = cnt;
      rmse = (float)inlier_error;
      valid_iterations++;
      *used_identity_out = 1;
    }
  }
  free(inl);
  free(refined);
  free(sel);
  *rmse_out = rmse;
  *n_inl_out = best_cnt;
  *valid_iterations_out = (int)valid_iterations;
  *real_iterations_out = real_iterations;
  return (unsigned)best_cnt >= min_inlier_threshold; /* :1275 */
}

/* node.cpp:1305-1429 (matchNodePair), ORB branch, EMM / ICP off (defaults). */
void oracle_match_node_pair(const oracle_params* p, const uint8_t* desc_newer, const float* xyz_newer, int n_newer,
                            int id_newer, const uint8_t* desc_older, const float* xyz_older, int n_older,
                            int id_older, uint64_t seed, uint64_t pair, oracle_pair_result* res,
                            oracle_dmatch* all_matches /* max_matches */, oracle_dmatch* inlier_matches /* max_matches */) {
  static const float I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  memset(res, 0, sizeof(*res));
  res->id1 = res->id2 = -1;
  memcpy(res->ransac_trafo, I4, sizeof(I4));
  oracle_dmatch* tmp = (oracle_dmatch*)malloc(sizeof(oracle_dmatch) * (size_t)(n_newer > 0 ? n_newer : 1));
  int n_all = oracle_feature_matching_orb((const uint64_t*)desc_newer, n_newer, (const uint64_t*)desc_older, n_older,
                                          p->max_matches, seed, pair, tmp);
  res->n_all_matches = n_all;
  if (all_matches) memcpy(all_matches, tmp, sizeof(oracle_dmatch) * (size_t)n_all);
  int found = 0;
  if ((unsigned)n_all >= (unsigned)p->min_matches) { /* :1319 */
    uint8_t* inl = (uint8_t*)malloc((size_t)(n_all > 0 ? n_all : 1));
    int n_inl, vi, ui, ri;
    found = oracle_get_relative_transformation_to(p, xyz_newer, xyz_older, tmp, n_all, seed, pair, res->ransac_trafo,
                                                  &res->rmse, inl, &n_inl, &vi, &ui, &ri);
    res->valid_iterations = vi;
    res->used_identity = ui;
    res->real_iterations = ri;
    res->n_inliers = n_inl;
    if (inlier_matches) {
      int k = 0;
      for (int i = 0; i < n_all; i++)
        if (inl[i]) inlier_matches[k++] = tmp[i];
    }
    free(inl);
    if (found) {
      res->info_scale = (double)((float)n_inl / (res->rmse * res->rmse)); /* :1335 size_t/(float*float) is float */
      res->id1 = id_older; /* :1337 */
      res->id2 = id_newer; /* :1338 */
    }
  }
  if (!found) { /* :1420 */
    res->id1 = res->id2 = -1;
  }
  free(tmp);
}

/* Batch driver (the QtConcurrent::blockingMapped fan-out, graph_manager.cpp:548);
 * threads>1 uses OpenMP when compiled with -fopenmp. */
void oracle_match_pairs(const oracle_params* p, const uint8_t* desc_newer, const float* xyz_newer, const int32_t* n_newer,
                        const uint8_t* desc_older, const float* xyz_older, const int32_t* n_older,
                        const int32_t* id_newer, const int32_t* id_older, int npairs, uint64_t seed,
                        int64_t first_pair_index, oracle_pair_result* results, oracle_dmatch* all_matches,
                        oracle_dmatch* inlier_matches, int threads) {
  int64_t* offn = (int64_t*)malloc(sizeof(int64_t) * (size_t)(npairs + 1));
  int64_t* offo = (int64_t*)malloc(sizeof(int64_t) * (size_t)(npairs + 1));
  offn[0] = offo[0] = 0;
  for (int i = 0; i < npairs; i++) {
    offn[i + 1] = offn[i] + n_newer[i];
    offo[i + 1] = offo[i] + n_older[i];
  }
  (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1)
#endif
  for (int i = 0; i < npairs; i++) {
    oracle_match_node_pair(p, desc_newer + 32 * offn[i], xyz_newer + 4 * offn[i], n_newer[i], id_newer ? id_newer[i] : i,
                           desc_older + 32 * offo[i], xyz_older + 4 * offo[i], n_older[i], id_older ? id_older[i] : i,
                           seed, (uint64_t)(first_pair_index + i), &results[i],
                           all_matches ? all_matches + (size_t)i * p->max_matches : NULL,
                           inlier_matches ? inlier_matches + (size_t)i * p->max_matches : NULL);
  }
  free(offn);
  free(offo);
}

/* ------------------------------------------------------------------------ */
/* node.cpp:67-97 (removeDepthless, use_feature_min_depth=false) on (x,y) pairs:
 * keep[i]=1 if the keypoint survives.  depth: row-major h x w float, NaN invalid. */
void oracle_remove_depthless(const float* xy, int n, const float* depth, int w, int h, uint8_t* keep) {
  for (int i = 0; i < n; i++) {
    float x = xy[2 * i], y = xy[2 * i + 1];
    keep[i] = 0;
    if (x >= w || x < 0 || y >= h || y < 0 || isnan(x) || isnan(y)) continue;
    float Z = depth[(size_t)lroundf(y) * w + lroundf(x)]; /* round(): half away from zero */
    if (isnan(Z)) continue;
    keep[i] = 1;
  }
}

/* node.cpp:900-965 (projectTo3D, depth-image overload) + misc2.h:49-65.
 * Returns number of points written (<= max_keyp); xyz1 gets (x,y,z,1). */
int oracle_project_to_3d(const float* xy, int n, const float* depth, int w, int h, double fx, double fy, double cx_,
                         double cy_, double depth_scaling, int max_keyp, float* xyz1, uint8_t* keep) {
  float fxinv = (float)(1. / fx), fyinv = (float)(1. / fy);
  float cx = (float)cx_, cy = (float)cy_;
  int cnt = 0;
  for (int i = 0; i < n; i++) {
    float x = xy[2 * i], y = xy[2 * i + 1];
    keep[i] = 0;
    if (cnt >= max_keyp) continue; /* :959 break */
    if (x >= w || x < 0 || y >= h || y < 0 || isnan(x) || isnan(y)) continue;
    float Z = (float)((double)depth[(size_t)lroundf(y) * w + lroundf(x)] * depth_scaling);
    if (isnan(Z)) continue;
    xyz1[4 * cnt + 0] = (x - cx) * Z * fxinv;
    xyz1[4 * cnt + 1] = (y - cy) * Z * fyinv;
    xyz1[4 * cnt + 2] = Z;
    xyz1[4 * cnt + 3] = 1.0f;
    keep[i] = 1;
    cnt++;
  }
  return cnt;
}

/* ------------------------------------------------------------------------ */
/* matchNodePair (node.cpp:1305-1429) AFTER featureMatching, for callers that computed the sorted match list
 * themselves (float-descriptor branch, restated in oracle/sift_oracle.py). */
void oracle_match_node_pair_from_matches(const oracle_params* p, const float* xyz_newer, int id_newer, const float* xyz_older,
                                         int id_older, const oracle_dmatch* matches, int n_all, uint64_t seed, uint64_t pair,
                                         oracle_pair_result* res, oracle_dmatch* inlier_matches) {
  static const float I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  memset(res, 0, sizeof(*res));
  res->id1 = res->id2 = -1;
  memcpy(res->ransac_trafo, I4, sizeof(I4));
  res->n_all_matches = n_all;
  int found = 0;
  if ((unsigned)n_all >= (unsigned)p->min_matches) {
    uint8_t* inl = (uint8_t*)malloc((size_t)(n_all > 0 ? n_all : 1));
    int n_inl, vi, ui, ri;
    found = oracle_get_relative_transformation_to(p, xyz_newer, xyz_older, matches, n_all, seed, pair, res->ransac_trafo, &res->rmse,
                                                  inl, &n_inl, &vi, &ui, &ri);
    res->valid_iterations = vi;
    res->used_identity = ui;
    res->real_iterations = ri;
    res->n_inliers = n_inl;
    if (inlier_matches) {
      int k = 0;
      for (int i = 0; i < n_all; i++)
        if (inl[i]) inlier_matches[k++] = matches[i];
    }
    free(inl);
    if (found) {
      res->info_scale = (double)((float)n_inl / (res->rmse * res->rmse));
      res->id1 = id_older;
      res->id2 = id_newer;
    }
  }
  if (!found) res->id1 = res->id2 = -1;
}
