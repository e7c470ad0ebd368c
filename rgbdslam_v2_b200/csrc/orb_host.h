// orb_host.h -- host-side launchers of the ORB kernels (orb.cu).
#pragma once
#include <cuda_runtime.h>

#include "../../include/rgbdslam_b200.h"
#include "orb.cuh"

namespace rb200 {

// 1: detect with the unfused k_fast_score / k_nms_collect / k_resize kernels (keeps the FAST score map in global memory)
void orb_set_legacy_detect(int on);

cudaError_t orb_upload_constants(const OrbGeom& g, const int* umax, cudaStream_t st);

// d_depth_for_mask != nullptr: detection mask = depthToCV8UC1(depth) != 0 (misc.cpp:414-418), d_mask ignored
cudaError_t orb_run_detect(const OrbGeom& g, const OrbTables& tab, int nframes, const uint8_t* d_gray, const uint8_t* d_mask,
                           const float* d_depth_for_mask, uint8_t* d_cell_img, uint8_t* d_cell_mask, uint8_t* d_score, OrbCand* d_cand,
                           int* d_cand_count, int* d_hist, int* d_mask_any, cudaStream_t st, int* launches);

// the adaptive-threshold recurrence of the F frames of a chunk, on the device (no host round trip)
cudaError_t orb_run_adapt(const OrbGeom& g, int nframes, const int* d_hist, const int* d_cand_count, const int* d_mask_any,
                          double* d_state, int* d_thr, int min_features, int max_features, int max_iters, int* d_err,
                          cudaStream_t st, int* launches);

cudaError_t orb_run_select(const OrbGeom& g, int nframes, int mode, int max_per_cell, int max_keypoints,
                           const uint8_t* d_cell_img, const OrbCand* d_cand, const int* d_cand_count, const int* d_thr,
                           float* d_resp, unsigned long long* d_cell_out, int* d_cell_out_count, const float* d_depth,
                           float depth_scaling, float4 Kinv, void* d_scratch, rgbdslam_b200_keypoint* d_kp, float4* d_xyz,
                           float2* d_trig /* mode 1: (cos, sin) of every keypoint's orientation for orb_run_describe, may be NULL */,
                           int* d_n, int kp_stride, cudaStream_t st, int* launches);

cudaError_t orb_run_describe(const OrbGeom& g, const OrbTables& tab, int nframes, const uint8_t* d_gray, uint8_t* d_pyr_raw,
                             uint8_t* d_pyr_blur, const rgbdslam_b200_keypoint* d_kp, const int* d_n, int kp_stride, int max_kp,
                             const float2* d_trig, uint8_t* d_desc, cudaStream_t st, int* launches);

constexpr int kOrbFrameCap = 4096;       // == kFrameCap in orb.cu
constexpr int kOrbFrameKpBytes = 20;     // sizeof(FrameKp)

}  // namespace rb200
