// common.cuh -- shared device/host definitions for the sm_100a front-end kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rgbdslam_b200.h"

namespace rb200 {

constexpr int kMaxMatchesCap = RGBDSLAM_B200_MAX_MATCHES_CAP;  // 512
constexpr int kMaxMaskWords = kMaxMatchesCap / 32;             // 16
constexpr int kMaxFeatures = 4096;                             // per node (SiftGPU limit too: sift_gpu_wrapper.cpp:231)

// One frame pair as the kernels see it (device pointers).
struct PairDesc {
  const uint32_t* q_desc;  // newer node descriptors, nq x 8 u32 (16 B aligned)
  const uint32_t* t_desc;  // older node descriptors, nt x 8 u32
  const float4* q_xyz;     // newer node points (x,y,z,1)
  const float4* t_xyz;     // older node points
  int32_t nq, nt;
  int32_t id_q, id_t;  // node ids (newer, older)
};

// Constant-memory copy of the parameters the kernels read.
struct DevParams {
  int32_t min_matches;
  int32_t max_matches;
  int32_t ransac_iterations;
  int32_t pad_;
  float max_dist_m;      // (float) max_dist_for_inliers        node.cpp:1105
  double sq_max_dist;    // (double)(max_dist_m*max_dist_m)     node.cpp:1152
  double sigma_depth;    // misc2.h:23
  double cov_z_const;    // (sigma*z0^2)^2 if the static-cache quirk is emulated, else <0
  double raster_cov_x;   // misc.cpp:702-709
  double raster_cov_y;
};

// Per-hypothesis record written by the RANSAC kernel and replayed by the selection kernel.
struct HypResult {
  double err;     // refined_error (1e6 if the hypothesis never produced a model)
  int32_t count;  // refined_matches.size()
  int32_t pad_;
  float T[12];  // R row-major (9) + t (3)
};

// ---- counter-based RNG (DESIGN.md "Random numbers"; same stream as oracle_rand31) -------------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t pair_key(uint64_t seed, uint64_t pair) { return mix64(seed ^ mix64(pair)); }
__host__ __device__ __forceinline__ uint32_t rand31(uint64_t key, uint32_t stream, uint32_t ctr) {
  return (uint32_t)(mix64(key ^ (((uint64_t)stream << 32) | ctr)) >> 33);
}

}  // namespace rb200
