// api_posegraph.cu -- C ABI of the pose-graph solve (declared in include/rgbdslam_b200.h).
#include <cmath>
#include <mutex>
#include <vector>

#include "posegraph.h"
#include "state.h"

using namespace rb200;

extern "C" {

int rgbdslam_b200_posegraph_optimize(int nv, double* poses, const uint8_t* fixed, int ne, const int32_t* ij,
                                     const double* meas, const double* info, double stop, double huber_delta,
                                     double* chi2, int* iters, int* cg_iters) {
  std::lock_guard<std::mutex> lk(g_state.mu);
  int rc = check_inited();
  if (rc) return rc;
  if (nv <= 0 || ne < 0 || !poses || !fixed || (ne > 0 && (!ij || !meas || !info)) || !(stop > 0) || !(huber_delta > 0)) {
    set_error("posegraph_optimize: bad arguments (nv > 0, stop > 0, huber_delta > 0)");
    return RGBDSLAM_B200_ERR_ARG;
  }
  for (size_t k = 0; k < (size_t)ne * 36; k++)
    if (!std::isfinite(info[k])) {  // 0 * inf = NaN in the linearisation: the solve would make no progress
      set_error("posegraph_optimize: non-finite entry in an information matrix");
      return RGBDSLAM_B200_ERR_ARG;
    }
  return posegraph_optimize(nv, poses, fixed, ne, ij, meas, info, stop, huber_delta, chi2, iters, cg_iters, nullptr, true);
}

int rgbdslam_b200_posegraph_chi2(int nv, const double* poses, int ne, const int32_t* ij, const double* meas,
                                 const double* info, double huber_delta, double* chi2, double* per_edge_chi2) {
  std::lock_guard<std::mutex> lk(g_state.mu);
  int rc = check_inited();
  if (rc) return rc;
  if (nv <= 0 || ne < 0 || !poses || (ne > 0 && (!ij || !meas || !info)) || !(huber_delta > 0)) {
    set_error("posegraph_chi2: bad arguments");
    return RGBDSLAM_B200_ERR_ARG;
  }
  std::vector<uint8_t> fixed(nv, 0);
  return posegraph_optimize(nv, const_cast<double*>(poses), fixed.data(), ne, ij, meas, info, 1.0, huber_delta, chi2,
                            nullptr, nullptr, per_edge_chi2, false);
}

}  // extern "C"
