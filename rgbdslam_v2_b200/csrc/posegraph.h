// posegraph.h -- host entry of the GPU pose-graph optimiser (posegraph.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rb200 {
// poses nv x 7 (t, q) in/out; returns 0 or an RGBDSLAM_B200_ERR_* code.  optimize=false: only chi2 / per-edge chi2.
int posegraph_optimize(int nv, double* poses, const uint8_t* fixed, int ne, const int32_t* ij, const double* meas,
                       const double* info, double stop, double huber_delta, double* chi2_out, int* iters_out,
                       int* cg_iters_out, double* per_edge_chi2, bool optimize);
void posegraph_release();
int posegraph_reserve(int nv, int ne);  // pre-size the cached device buffers
// pose-pose constraints for other solvers (landmark_ba.cu): per-edge normal-equation blocks [A 36 | B 36 | C 36 | gi 6 | gj 6]
// (A = Ji'WJi, B = Jj'WJj, C = Ji'WJj, g = J'We, all scaled by the Huber weight), pose update X <- X * fromVectorMQT(d),
// per-block partial sums of (robust, plain) chi2 (2 doubles per 256 edges)
constexpr int kPgEdgeBlk = 120;
cudaError_t pg_launch_linearize(int ne, const double* x, const int32_t* ij, const double* meas, const double* info, double delta, double* blk,
                                cudaStream_t st);
cudaError_t pg_launch_update(int nv, const double* xin, const double* dlt, const uint8_t* fixed, double* xout, cudaStream_t st);
cudaError_t pg_launch_chi2(int ne, const double* x, const int32_t* ij, const double* meas, const double* info, double delta, double* part,
                           cudaStream_t st);
int landmark_ba(int n_cams, double* poses7, const uint8_t* fixed, int n_points, double* points3, int n_obs, const int32_t* obs_cam,
                const int32_t* obs_point, const double* obs_uvd, const double* obs_info3, const double K4[4], int n_edges,
                const int32_t* ij, const double* meas7, const double* info36, int iterations, double huber_delta, double* chi2_before,
                double* chi2_after, int* lm_iterations, int* pcg_iterations);
int landmark_ba_release();  // frees the cached solver buffers (rgbdslam_b200_shutdown)
}  // namespace rb200
