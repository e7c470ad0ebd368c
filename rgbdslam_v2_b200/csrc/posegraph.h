// posegraph.h -- host entry of the GPU pose-graph optimiser (posegraph.cu).
#pragma once
#include <stdint.h>

namespace rb200 {
// poses nv x 7 (t, q) in/out; returns 0 or an RGBDSLAM_B200_ERR_* code.  optimize=false: only chi2 / per-edge chi2.
int posegraph_optimize(int nv, double* poses, const uint8_t* fixed, int ne, const int32_t* ij, const double* meas,
                       const double* info, double stop, double huber_delta, double* chi2_out, int* iters_out,
                       int* cg_iters_out, double* per_edge_chi2, bool optimize);
void posegraph_release();  // frees the cached solver buffers (rgbdslam_b200_shutdown)
}  // namespace rb200
