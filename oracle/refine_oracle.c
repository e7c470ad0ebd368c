/*
 * refine_oracle.c -- CPU ORACLE of the pairwise g2o refinement.  TEST INFRASTRUCTURE ONLY
 * (used by tests/ and by nothing in the product).
 *
 * Restates, from the reference sources:
 *   getTransformFromMatchesG2O   src/transformation_estimation.cpp:126-170  (optimizerSetup :36-61, sensorVerticesSetup
 *                                :64-90, edgeToFeature :91-124), point_information_matrix src/misc2.h:37-47
 *   the accept / re-refine logic  src/node.cpp:1225-1268
 * and, from upstream g2o (fork felixendres/g2o, branch c++03 -- not under /root/reference, no pinned commit; published
 * semantics of types/slam3d): VertexSE3 (oplus: X <- X * fromVectorMQT(d)), VertexPointXYZ (p <- p + d),
 * EdgeSE3PointXYZDepth (error (u, v, z) - measurement with Kcam, analytic Jacobian), OptimizationAlgorithmGaussNewton
 * (exactly `iterations` undamped steps, H dx = -b) over BlockSolverX WITHOUT marginalisation: one sparse system of
 * 6 + 3 n unknowns.  The oracle solves that full system densely (Cholesky) on purpose -- the CUDA path uses the Schur
 * complement; both must agree to rounding.  Parity unpinned (g2o absent): cross-checked against a numerical-Jacobian
 * Gauss-Newton in tests/test_refine_oracle.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int32_t queryIdx, trainIdx, imgIdx;
  float distance;
} oracle_dmatch;

typedef struct {
  int32_t min_matches, max_matches, ransac_iterations, pad_;
  double max_dist_for_inliers, sigma_depth, depth_cov_z0;
} oracle_params;

int oracle_compute_inliers_and_error(const oracle_params* p, const oracle_dmatch* all, int n_all, const float T4f[16],
                                     const float* origins, const float* earlier, uint8_t* inl, double* return_mean_error,
                                     double sq_max_dist);

static const double KFX = 521.0, KFY = 521.0, KCX = 319.5, KCY = 239.5; /* transformation_estimation.cpp:56 */

static double depth_cov(const oracle_params* p, double z) { /* misc2.h:20-35 incl. the static cache */
  double zz = p->depth_cov_z0 > 0 ? p->depth_cov_z0 : z;
  double sd = p->sigma_depth * zz * zz;
  return sd * sd;
}

/* Eigen::Quaterniond(Matrix3d) + normalisation (SE3Quat ctor) + toRotationMatrix: R, t column-major 4x4 float in */
static void pose_from_matrix4f(const float T[16], double R[9], double t[3]) {
  double m[3][3];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) m[r][c] = (double)T[4 * c + r];
  double q[4]; /* x y z w */
  double tr = m[0][0] + m[1][1] + m[2][2];
  if (tr > 0) {
    double s = sqrt(tr + 1.0);
    q[3] = 0.5 * s;
    s = 0.5 / s;
    q[0] = (m[2][1] - m[1][2]) * s;
    q[1] = (m[0][2] - m[2][0]) * s;
    q[2] = (m[1][0] - m[0][1]) * s;
  } else {
    int i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    double s = sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
    q[i] = 0.5 * s;
    s = 0.5 / s;
    q[3] = (m[k][j] - m[j][k]) * s;
    q[j] = (m[j][i] + m[i][j]) * s;
    q[k] = (m[k][i] + m[i][k]) * s;
  }
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; i++) q[i] /= n;
  double x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
  for (int r = 0; r < 3; r++) t[r] = (double)T[12 + r];
}

/* X <- X * fromVectorMQT(d): translation d[0..2], quaternion vector part d[3..5] (g2o isometry3d_mappings) */
static void pose_oplus(double R[9], double t[3], const double d[6]) {
  double vx = d[3], vy = d[4], vz = d[5];
  double w = 1.0 - (vx * vx + vy * vy + vz * vz);
  double dR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (w >= 0) {
    w = sqrt(w);
    dR[0] = 1 - 2 * (vy * vy + vz * vz); dR[1] = 2 * (vx * vy - vz * w);     dR[2] = 2 * (vx * vz + vy * w);
    dR[3] = 2 * (vx * vy + vz * w);     dR[4] = 1 - 2 * (vx * vx + vz * vz); dR[5] = 2 * (vy * vz - vx * w);
    dR[6] = 2 * (vx * vz - vy * w);     dR[7] = 2 * (vy * vz + vx * w);     dR[8] = 1 - 2 * (vx * vx + vy * vy);
  }
  double nt[3], nR[9];
  for (int r = 0; r < 3; r++) nt[r] = t[r] + R[3 * r] * d[0] + R[3 * r + 1] * d[1] + R[3 * r + 2] * d[2];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) nR[3 * r + c] = R[3 * r] * dR[c] + R[3 * r + 1] * dR[3 + c] + R[3 * r + 2] * dR[6 + c];
  memcpy(R, nR, sizeof(nR));
  memcpy(t, nt, sizeof(nt));
}

/* EdgeSE3PointXYZDepth: error and Jacobians for camera (R, t) and world point pw.  Jc 3x6, Jp 3x3 (row-major). */
static void edge_depth(const double R[9], const double t[3], const double pw[3], const double meas[3], double e[3], double Jc[18],
                       double Jp[9]) {
  double d[3] = {pw[0] - t[0], pw[1] - t[1], pw[2] - t[2]};
  double zc[3]; /* w2l * p = R^T (p - t) */
  for (int c = 0; c < 3; c++) zc[c] = R[c] * d[0] + R[3 + c] * d[1] + R[6 + c] * d[2];
  double J[3][9];
  memset(J, 0, sizeof(J));
  J[0][0] = J[1][1] = J[2][2] = -1.0;
  J[0][4] = -2 * zc[2]; J[0][5] = 2 * zc[1];
  J[1][3] = 2 * zc[2];  J[1][5] = -2 * zc[0];
  J[2][3] = -2 * zc[1]; J[2][4] = 2 * zc[0];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) J[r][6 + c] = R[3 * c + r]; /* w2l rotation = R^T */
  /* Kcam_inverseOffsetR = K (offset = identity) */
  double Jprime[3][9];
  for (int c = 0; c < 9; c++) {
    Jprime[0][c] = KFX * J[0][c] + KCX * J[2][c];
    Jprime[1][c] = KFY * J[1][c] + KCY * J[2][c];
    Jprime[2][c] = J[2][c];
  }
  double zp[3] = {KFX * zc[0] + KCX * zc[2], KFY * zc[1] + KCY * zc[2], zc[2]}; /* w2i * p */
  double iz2 = 1.0 / (zp[2] * zp[2]);
  double Jh[3][9];
  for (int c = 0; c < 9; c++) {
    Jh[0][c] = iz2 * (Jprime[0][c] * zp[2] - zp[0] * Jprime[2][c]);
    Jh[1][c] = iz2 * (Jprime[1][c] * zp[2] - zp[1] * Jprime[2][c]);
    Jh[2][c] = Jprime[2][c];
  }
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 6; c++) Jc[6 * r + c] = Jh[r][c];
    for (int c = 0; c < 3; c++) Jp[3 * r + c] = Jh[r][6 + c];
  }
  e[0] = zp[0] / zp[2] - meas[0];
  e[1] = zp[1] / zp[2] - meas[1];
  e[2] = zp[2] - meas[2];
}

/* error only (used by the numerical-Jacobian cross-check in the tests) */
void oracle_edge_depth_error(const double R[9], const double t[3], const double pw[3], const double meas[3], double e[3]) {
  double Jc[18], Jp[9];
  edge_depth(R, t, pw, meas, e, Jc, Jp);
}
void oracle_edge_depth_jacobians(const double R[9], const double t[3], const double pw[3], const double meas[3], double Jc[18],
                                 double Jp[9]) {
  double e[3];
  edge_depth(R, t, pw, meas, e, Jc, Jp);
}
void oracle_pose_oplus(double R[9], double t[3], const double d[6]) { pose_oplus(R, t, d); }

static int cholesky_solve(double* A, double* b, int n) { /* A = L L^T in place (lower), b <- solution */
  for (int j = 0; j < n; j++) {
    double s = A[(size_t)j * n + j];
    for (int k = 0; k < j; k++) s -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(s > 0)) return 0;
    double l = sqrt(s);
    A[(size_t)j * n + j] = l;
    for (int i = j + 1; i < n; i++) {
      double v = A[(size_t)i * n + j];
      for (int k = 0; k < j; k++) v -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      A[(size_t)i * n + j] = v / l;
    }
  }
  for (int i = 0; i < n; i++) {
    double v = b[i];
    for (int k = 0; k < i; k++) v -= A[(size_t)i * n + k] * b[k];
    b[i] = v / A[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double v = b[i];
    for (int k = i + 1; k < n; k++) v -= A[(size_t)k * n + i] * b[k];
    b[i] = v / A[(size_t)i * n + i];
  }
  return 1;
}

/* getTransformFromMatchesG2O: `sel` = indices into matches; kp_* = (u, v) per feature; xyz_* = Vector4f per feature. */
void oracle_get_transform_from_matches_g2o(const oracle_params* p, const float* xyz_newer, const float* kp_newer,
                                           const float* xyz_earlier, const float* kp_earlier, const oracle_dmatch* matches,
                                           const int* sel, int nsel, float T[16] /* in: estimate, out */, int iterations) {
  double R1[9], t1[3];
  pose_from_matrix4f(T, R1, t1); /* cam1 (earlier node) starts at the estimate; cam2 (newer) is fixed at identity */
  const double R2[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t2[3] = {0, 0, 0};
  double* pts = (double*)malloc(sizeof(double) * 3 * (size_t)(nsel > 0 ? nsel : 1));
  double* m1 = (double*)malloc(sizeof(double) * 3 * (size_t)(nsel > 0 ? nsel : 1));
  double* m2 = (double*)malloc(sizeof(double) * 3 * (size_t)(nsel > 0 ? nsel : 1));
  double* w1 = (double*)malloc(sizeof(double) * (size_t)(nsel > 0 ? nsel : 1));
  double* w2 = (double*)malloc(sizeof(double) * (size_t)(nsel > 0 ? nsel : 1));
  for (int k = 0; k < nsel; k++) {
    const oracle_dmatch* m = &matches[sel[k]];
    const float* pe = xyz_earlier + 4 * (size_t)m->trainIdx;
    const float* pn = xyz_newer + 4 * (size_t)m->queryIdx;
    /* edgeToFeature(earlier, trainIdx, cam1, v) then edgeToFeature(newer, queryIdx, cam2, v): the second call's
     * setEstimate wins, so the point starts at the NEWER node's position (= world frame). */
    float d1 = pe[2], d2 = pn[2];
    m1[3 * k] = kp_earlier[2 * (size_t)m->trainIdx]; m1[3 * k + 1] = kp_earlier[2 * (size_t)m->trainIdx + 1];
    m2[3 * k] = kp_newer[2 * (size_t)m->queryIdx];   m2[3 * k + 1] = kp_newer[2 * (size_t)m->queryIdx + 1];
    if (!isnan(d1)) { m1[3 * k + 2] = d1; w1[k] = 1.0 / depth_cov(p, d1); } else { m1[3 * k + 2] = 10.0; w1[k] = 1.0 / depth_cov(p, 10.0); }
    if (!isnan(d2)) {
      m2[3 * k + 2] = d2; w2[k] = 1.0 / depth_cov(p, d2);
      pts[3 * k] = pn[0]; pts[3 * k + 1] = pn[1]; pts[3 * k + 2] = pn[2];
    } else {
      m2[3 * k + 2] = 10.0; w2[k] = 1.0 / depth_cov(p, 10.0);
      pts[3 * k] = (double)pn[0] * 10; pts[3 * k + 1] = (double)pn[1] * 10; pts[3 * k + 2] = 10.0;
    }
  }
  const int dim = 6 + 3 * nsel;
  double* H = (double*)malloc(sizeof(double) * (size_t)dim * dim);
  double* b = (double*)malloc(sizeof(double) * (size_t)dim);
  for (int it = 0; it < iterations && nsel > 0; it++) {
    memset(H, 0, sizeof(double) * (size_t)dim * dim);
    memset(b, 0, sizeof(double) * (size_t)dim);
    for (int k = 0; k < nsel; k++) {
      for (int cam = 0; cam < 2; cam++) {
        double e[3], Jc[18], Jp[9];
        const double om[3] = {1.0, 1.0, cam == 0 ? w1[k] : w2[k]}; /* point_information_matrix: diag(1, 1, 1/cov_z) */
        edge_depth(cam == 0 ? R1 : R2, cam == 0 ? t1 : t2, pts + 3 * k, cam == 0 ? m1 + 3 * k : m2 + 3 * k, e, Jc, Jp);
        const int po = 6 + 3 * k;
        for (int r = 0; r < 3; r++) {
          for (int i = 0; i < 3; i++) {
            b[po + i] -= Jp[3 * r + i] * om[r] * e[r];
            for (int j = 0; j < 3; j++) H[(size_t)(po + i) * dim + po + j] += Jp[3 * r + i] * om[r] * Jp[3 * r + j];
          }
          if (cam == 0) { /* cam2 is fixed */
            for (int i = 0; i < 6; i++) {
              b[i] -= Jc[6 * r + i] * om[r] * e[r];
              for (int j = 0; j < 6; j++) H[(size_t)i * dim + j] += Jc[6 * r + i] * om[r] * Jc[6 * r + j];
              for (int j = 0; j < 3; j++) {
                H[(size_t)i * dim + po + j] += Jc[6 * r + i] * om[r] * Jp[3 * r + j];
                H[(size_t)(po + j) * dim + i] += Jc[6 * r + i] * om[r] * Jp[3 * r + j];
              }
            }
          }
        }
      }
    }
    if (!cholesky_solve(H, b, dim)) break; /* solver failure ends the optimisation */
    pose_oplus(R1, t1, b);
    for (int k = 0; k < 3 * nsel; k++) pts[k] += b[6 + k];
  }
  /* cams.first->estimate().cast<float>().inverse().matrix() */
  float Rf[9], tf[3];
  for (int i = 0; i < 9; i++) Rf[i] = (float)R1[i];
  for (int i = 0; i < 3; i++) tf[i] = (float)t1[i];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) T[4 * c + r] = Rf[3 * c + r]; /* R^T */
  for (int r = 0; r < 3; r++) T[12 + r] = -(Rf[r] * tf[0] + Rf[3 + r] * tf[1] + Rf[6 + r] * tf[2]);
  T[3] = T[7] = T[11] = 0.f;
  T[15] = 1.f;
  free(pts); free(m1); free(m2); free(w1); free(w2); free(H); free(b);
}

/* node.cpp:1225-1268: refine the RANSAC result, keep it when it is at least as good.  inl[i] marks matches[i] as inlier. */
void oracle_refine_g2o(const oracle_params* p, int g2o_iterations, const float* xyz_newer, const float* kp_newer,
                       const float* xyz_earlier, const float* kp_earlier, const oracle_dmatch* matches, int n_all, float T[16],
                       float* rmse, uint8_t* inl, int* n_inl, int* valid_iterations) {
  unsigned min_inlier_threshold = (unsigned)p->min_matches;
  if (min_inlier_threshold > 0.75 * n_all) min_inlier_threshold = (unsigned)(0.75 * n_all);
  if (!(g2o_iterations > 0 && (unsigned)*n_inl > min_inlier_threshold)) return; /* :1226 */
  const float max_dist_m = (float)p->max_dist_for_inliers;
  const double sq_max = (double)(max_dist_m * max_dist_m);
  int* sel = (int*)malloc(sizeof(int) * (size_t)n_all);
  uint8_t* inl2 = (uint8_t*)malloc((size_t)n_all);
  int nsel = 0;
  for (int i = 0; i < n_all; i++)
    if (inl[i]) sel[nsel++] = i;
  float T1[16];
  memcpy(T1, T, sizeof(T1));
  oracle_get_transform_from_matches_g2o(p, xyz_newer, kp_newer, xyz_earlier, kp_earlier, matches, sel, nsel, T1, g2o_iterations);
  double err;
  int cnt = oracle_compute_inliers_and_error(p, matches, n_all, T1, xyz_newer, xyz_earlier, inl2, &err, sq_max);
  if (cnt >= *n_inl || ((unsigned)cnt >= min_inlier_threshold && err < (double)*rmse)) { /* :1241 */
    if (cnt > *n_inl) { /* :1243-1251 refine again with the new inliers */
      nsel = 0;
      for (int i = 0; i < n_all; i++)
        if (inl2[i]) sel[nsel++] = i;
      oracle_get_transform_from_matches_g2o(p, xyz_newer, kp_newer, xyz_earlier, kp_earlier, matches, sel, nsel, T1, g2o_iterations);
      cnt = oracle_compute_inliers_and_error(p, matches, n_all, T1, xyz_newer, xyz_earlier, inl2, &err, sq_max);
    }
    if (cnt >= *n_inl) { /* :1254-1263 */
      memcpy(T, T1, sizeof(T1));
      memcpy(inl, inl2, (size_t)n_all);
      *n_inl = cnt;
      *rmse = (float)err;
      (*valid_iterations)++;
    }
  }
  free(sel);
  free(inl2);
}
