// emm.cu -- Environment Measurement Model (parameter observability_threshold > 0):
//   k_build_cloud        the z-plane of createXYZRGBPointCloud (misc.cpp:467-556): every cloud_creation_skip_step-th pixel,
//                        NaN where !(Z >= minimum_depth); x / y are recomputed from the pixel grid (backProject, misc2.h:49-65)
//   k_emm_pairs          pairwiseObservationLikelihood (node.cpp:1520-1554) = observationLikelihood (misc.cpp:814-969) in both
//                        directions + observation_criterion_met (misc.cpp:1136-1148) applied to the pair result
// Dense projective data association on the sub-sampled clouds: embarrassingly parallel per sampled pixel, HBM / latency bound
// (<= 9 random depth reads per sample, 2 x (W/s/k) x (H/s/k) samples per pair: 2400 at the defaults s = 2, k = 8).
#include "kernels.h"

namespace rb200 {


__global__ void __launch_bounds__(256) k_build_cloud(const float* __restrict__ depth, int w, int h, int step, float scaling,
                                                     float min_depth, float* __restrict__ cloud_z, int cw, int ch) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= cw * ch) return;
  const int rx = i % cw, ry = i / cw;
  const int u = rx * step, v = ry * step;
  float z = __int_as_float(0x7fc00000);
  if (u < w && v < h) {
    const float Z = depth[(size_t)v * w + u] * scaling;  // misc.cpp:520
    if (Z >= min_depth) z = Z;                            // :523 (also rejects NaN)
  }
  cloud_z[i] = z;
}

cudaError_t launch_build_cloud(const float* d_depth, int w, int h, int step, float scaling, float min_depth, float* cloud_z, int cw,
                               int ch, cudaStream_t stream) {
  if (cw <= 0 || ch <= 0) return cudaSuccess;
  k_build_cloud<<<(cw * ch + 255) / 256, 256, 0, stream>>>(d_depth, w, h, step, scaling, min_depth, cloud_z, cw, ch);
  return cudaGetLastError();
}

struct EmmView {
  const float* z;   // cloud z-plane
  int cw, ch;
  float fx, fy, cx, cy;  // full-resolution intrinsics of the camera that took it
};

__device__ __forceinline__ int round_like_ref(float d) { return (int)floor((double)d + 0.5); }  // misc.cpp:804-807

// One direction of the model: the `src` cloud transformed by T (row-major R, t: src frame -> dst frame) and projected into
// the `dst` depth raster.  Each thread takes samples; returns this thread's (good, bad, occluded, all).
__device__ void emm_direction(const EmmView& src, const EmmView& dst, const float R[9], const float t[3], int cloud_step,
                              int skip_step, double cov_z_const, double sigma_depth, unsigned& good, unsigned& bad, unsigned& occl,
                              unsigned& all) {
  const float sfxinv = (float)(1.0 / (double)src.fx), sfyinv = (float)(1.0 / (double)src.fy);  // misc.cpp:64-69
  // "downsampled cloud?" branch (misc.cpp:854-861): intrinsics of the raster the old cloud lives on
  const float fx = dst.fx / cloud_step, fy = dst.fy / cloud_step, cx = dst.cx / cloud_step, cy = dst.cy / cloud_step;
  const int nsx = (src.cw + skip_step - 1) / skip_step, nsy = (src.ch + skip_step - 1) / skip_step;
  for (int sidx = threadIdx.x; sidx < nsx * nsy; sidx += blockDim.x) {
    all++;  // the loop header increments `all` for every sampled raster cell (:872)
    const int rx = (sidx % nsx) * skip_step, ry = (sidx / nsx) * skip_step;
    const float Z = src.z[(size_t)ry * src.cw + rx];
    // the source point: NaN depth keeps x / y of the 1 m ray (misc.cpp:525-529) -> the transformed z is NaN as well
    const float u = (float)(rx * cloud_step), v = (float)(ry * cloud_step);
    float px, py, pz;
    if (isnan(Z)) {
      px = (u - src.cx) * 1.0f * sfxinv;
      py = (v - src.cy) * 1.0f * sfyinv;
      pz = Z;
    } else {
      px = (u - src.cx) * Z * sfxinv;
      py = (v - src.cy) * Z * sfyinv;
      pz = Z;
    }
    const float qx = R[0] * px + R[1] * py + R[2] * pz + t[0];
    const float qy = R[3] * px + R[4] * py + R[5] * pz + t[1];
    const float qz = R[6] * px + R[7] * py + R[8] * pz + t[2];
    if (qz != qz) continue;   // NaN
    if (qz < 0) continue;     // behind the camera
    const int ocx = round_like_ref((qx / qz) * fx + cx);
    const int ocy = round_like_ref((qy / qz) * fy + cy);
    if (ocx >= dst.cw || ocx < 0 || ocy >= dst.ch || ocy < 0) continue;
    const int nbhd = 2;
    bool good_point = false, occluded_point = false, bad_point = false;
    const int startx = max(0, ocx - nbhd), starty = max(0, ocy - nbhd);
    const int endx = min(dst.cw, ocx + nbhd + 1), endy = min(dst.ch, ocy + nbhd + 1);
    for (int oy = starty; oy < endy; oy += 2)
      for (int ox = startx; ox < endx; ox += 2) {
        const float oz = dst.z[(size_t)oy * dst.cw + ox];
        if (oz != oz) continue;
        const double old_sigma = cloud_step * (cov_z_const >= 0.0 ? cov_z_const : (sigma_depth * (double)oz * (double)oz) * (sigma_depth * (double)oz * (double)oz));
        const double new_sigma = cloud_step * (cov_z_const >= 0.0 ? cov_z_const : (sigma_depth * (double)qz * (double)qz) * (sigma_depth * (double)qz * (double)qz));
        const double joint_sigma = old_sigma + new_sigma;
        // cdf(old_p.z, p.z, sqrt(joint_sigma)) with the reference's truncated SQRT_2 (misc.cpp:801, 809-812)
        const double p_new_in_front = 0.5 * (1 + erf(((double)oz - (double)qz) / (sqrt(joint_sigma) * 1.41421)));
        if (p_new_in_front < 0.001) occluded_point = true;
        else if (p_new_in_front < 0.999) good_point = true;
        else bad_point = true;
      }
    if (good_point) good++;
    else if (occluded_point) occl++;
    else if (bad_point) bad++;
  }
}

__device__ __forceinline__ unsigned block_sum(unsigned v, unsigned* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  unsigned s = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += scratch[w];
  return s;
}

struct EmmArgs {
  int cloud_step, skip_step;
  double cov_z_const, sigma_depth, observability_threshold;
};

// counts[4] = inlier, outlier, occluded, all points (MatchingResult, matching_result.h:40-42)
__device__ void emm_pair(const EmmView& newer, const EmmView& older, const float* T16 /* column-major, newer -> older */,
                         const EmmArgs& a, unsigned counts[4], unsigned* scratch) {
  float R[9], t[3], Ri[9], ti[3];
#pragma unroll
  for (int r = 0; r < 3; r++) {
#pragma unroll
    for (int c = 0; c < 3; c++) R[3 * r + c] = T16[4 * c + r];
    t[r] = T16[12 + r];
  }
  // mr.final_trafo.inverse(): cofactor inverse of the affine matrix in float
  const float c00 = R[4] * R[8] - R[5] * R[7], c01 = R[5] * R[6] - R[3] * R[8], c02 = R[3] * R[7] - R[4] * R[6];
  const float det = R[0] * c00 + R[1] * c01 + R[2] * c02;
  const float id = 1.0f / det;
  Ri[0] = c00 * id; Ri[1] = (R[2] * R[7] - R[1] * R[8]) * id; Ri[2] = (R[1] * R[5] - R[2] * R[4]) * id;
  Ri[3] = c01 * id; Ri[4] = (R[0] * R[8] - R[2] * R[6]) * id; Ri[5] = (R[2] * R[3] - R[0] * R[5]) * id;
  Ri[6] = c02 * id; Ri[7] = (R[1] * R[6] - R[0] * R[7]) * id; Ri[8] = (R[0] * R[4] - R[1] * R[3]) * id;
#pragma unroll
  for (int r = 0; r < 3; r++) ti[r] = -(Ri[3 * r] * t[0] + Ri[3 * r + 1] * t[1] + Ri[3 * r + 2] * t[2]);
  unsigned g = 0, b = 0, o = 0, al = 0;
  emm_direction(newer, older, R, t, a.cloud_step, a.skip_step, a.cov_z_const, a.sigma_depth, g, b, o, al);   // node.cpp:1527-1535
  emm_direction(older, newer, Ri, ti, a.cloud_step, a.skip_step, a.cov_z_const, a.sigma_depth, g, b, o, al);  // :1538-1548
  counts[0] = block_sum(g, scratch);
  counts[1] = block_sum(b, scratch);
  counts[2] = block_sum(o, scratch);
  counts[3] = block_sum(al, scratch);
}

__global__ void __launch_bounds__(256) k_emm_pairs(const PairDesc* __restrict__ pairs, EmmArgs a,
                                                   rgbdslam_b200_pair_result* __restrict__ results) {
  __shared__ unsigned scratch[8];
  const int p = blockIdx.x;
  rgbdslam_b200_pair_result res = results[p];
  if (res.id1 < 0) return;  // the model only judges transformations RANSAC accepted (node.cpp:1336-1345)
  const PairDesc pd = pairs[p];
  EmmView nv{pd.q_cloud, pd.q_cw, pd.q_ch, pd.q_K[0], pd.q_K[1], pd.q_K[2], pd.q_K[3]};
  EmmView ov{pd.t_cloud, pd.t_cw, pd.t_ch, pd.t_K[0], pd.t_K[1], pd.t_K[2], pd.t_K[3]};
  unsigned counts[4];
  emm_pair(nv, ov, res.ransac_trafo, a, counts, scratch);
  if (threadIdx.x == 0) {
    res.inlier_points = counts[0];
    res.outlier_points = counts[1];
    res.occluded_points = counts[2];
    res.all_points = counts[3];
    // observation_criterion_met(inliers, outliers, occluded + inliers + outliers, quality) (misc.cpp:1136-1148)
    const double quality = counts[0] / (double)(counts[0] + counts[1]);
    const double certainty = counts[0] / (double)(counts[2] + counts[0] + counts[1]);
    if (!(quality > a.observability_threshold && certainty > 0.25)) res.id1 = res.id2 = -1;  // node.cpp:1420
    results[p] = res;
  }
}

cudaError_t launch_emm_pairs(const PairDesc* pairs, int npairs, int cloud_step, int skip_step, double cov_z_const,
                             double sigma_depth, double observability_threshold, rgbdslam_b200_pair_result* results,
                             cudaStream_t stream) {
  if (npairs <= 0) return cudaSuccess;
  EmmArgs a{cloud_step, skip_step, cov_z_const, sigma_depth, observability_threshold};
  k_emm_pairs<<<npairs, 256, 0, stream>>>(pairs, a, results);
  return cudaGetLastError();
}

// rgbdslam_b200_observation_likelihood: one pair, explicit transformation, counts only
__global__ void __launch_bounds__(256) k_emm_single(EmmView newer, EmmView older, const float* __restrict__ T16, EmmArgs a,
                                                    unsigned* __restrict__ counts_out) {
  __shared__ unsigned scratch[8];
  __shared__ float sT[16];
  if (threadIdx.x < 16) sT[threadIdx.x] = T16[threadIdx.x];
  __syncthreads();
  unsigned counts[4];
  emm_pair(newer, older, sT, a, counts, scratch);
  if (threadIdx.x == 0)
    for (int k = 0; k < 4; k++) counts_out[k] = counts[k];
}

cudaError_t launch_emm_single(const float* q_cloud, int q_cw, int q_ch, const float* qK, const float* t_cloud, int t_cw, int t_ch,
                              const float* tK, const float* d_T16, int cloud_step, int skip_step, double cov_z_const,
                              double sigma_depth, unsigned* d_counts, cudaStream_t stream) {
  EmmArgs a{cloud_step, skip_step, cov_z_const, sigma_depth, 0.0};
  EmmView nv{q_cloud, q_cw, q_ch, qK[0], qK[1], qK[2], qK[3]};
  EmmView ov{t_cloud, t_cw, t_ch, tK[0], tK[1], tK[2], tK[3]};
  k_emm_single<<<1, 256, 0, stream>>>(nv, ov, d_T16, a, d_counts);
  return cudaGetLastError();
}

}  // namespace rb200
