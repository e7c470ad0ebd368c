// Drives the GraphManager shim (include/rgbdslam_b200/graph_manager.hpp) the way OpenNIListener drives the reference:
// one Node per frame, addNode in arrival order, optimizeGraph, pruneEdgesWithErrorAbove, saveTrajectory.
// CPU: compile + link only; GPU: run (exit code 77 = no GPU).
#include <cmath>
#include <cstdio>
#include <vector>

#include "rgbdslam_b200/graph_manager.hpp"

using namespace rgbdslam_b200;

static uint64_t s = 0x243F6A8885A308D3ull;
static uint32_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 32); }
static float frand() { return (rnd() & 0xFFFFFF) / 16777216.f; }

int main() {
  rgbdslam_b200_params p;
  rgbdslam_b200_default_params(&p);
  p.depth_cov_z0 = 2.0;
  if (rgbdslam_b200_init(0, &p) != 0) {
    std::printf("init failed (expected without a GPU): %s\n", rgbdslam_b200_last_error());
    return 77;
  }
  // landmarks on a wall 2.5-3.5 m in front, the camera slides along x by 4 cm per frame
  const int L = 2500, F = 14;
  const float step = 0.04f;
  std::vector<float> lx(L), ly(L), lz(L);
  std::vector<uint8_t> ld((size_t)L * 32);
  for (int i = 0; i < L; i++) {
    lx[i] = -2.0f + 5.0f * frand(); ly[i] = -1.0f + 2.0f * frand(); lz[i] = 2.5f + frand();
    for (int k = 0; k < 32; k++) ld[(size_t)i * 32 + k] = (uint8_t)rnd();
  }
  GraphManager gm;
  gm.seed = 5;
  for (int f = 0; f < F; f++) {
    std::vector<uint8_t> desc;
    std::vector<Vector4f> xyz;
    for (int i = 0; i < L && xyz.size() < 900; i++) {
      const float x = lx[i] - step * f, y = ly[i], z = lz[i];
      const float u = 525.f * x / z + 319.5f, v = 525.f * y / z + 239.5f;
      if (u < 20 || u > 620 || v < 20 || v > 460) continue;
      for (int k = 0; k < 32; k++) desc.push_back(ld[(size_t)i * 32 + k] ^ (uint8_t)((k % 8 == f % 8) ? (1u << (rnd() % 8)) : 0));
      const float nz = z + 0.002f * (frand() - 0.5f);
      xyz.push_back(Vector4f{x * nz / z, y * nz / z, nz, 1.f});
    }
    Node* n = new Node(-1, desc, xyz);  // the id is assigned by addNode like in the reference
    n->stamp_ = f / 30.0;
    const bool added = gm.addNode(n);
    if (!added) { std::printf("frame %d not added\n", f); delete n; }
  }
  const double chi2 = gm.optimizeGraph();
  // optimizeGraph(0.0) means "use the parameter" like the reference (graph_manager.cpp:942: break_criterion > 0.0 ? ... : param)
  const double chi2_again = gm.optimizeGraph(0.0);
  int ok = gm.graph_.size() == (size_t)F && gm.estimates_.size() == (size_t)F && gm.edges_.size() >= (size_t)(3 * F - 10);
  double max_err = 0;
  for (auto& kv : gm.estimates_) {
    const double* e = kv.second.v;
    const double ex = e[0] - step * kv.first, ey = e[1], ez = e[2];
    max_err = std::max(max_err, std::sqrt(ex * ex + ey * ey + ez * ez));
  }
  std::printf("nodes %zu edges %zu keyframes %zu chi2 %.4f max position error %.4f m\n", gm.graph_.size(), gm.edges_.size(),
              gm.keyframe_ids_.size(), chi2, max_err);
  ok = ok && max_err < 0.01 && chi2 >= 0;
  const unsigned pruned = gm.pruneEdgesWithErrorAbove(1e9f);  // nothing is that bad
  ok = ok && pruned == 0;
  gm.saveTrajectory("/tmp/rgbdslam_b200_traj_estimate.txt");
  FILE* f = std::fopen("/tmp/rgbdslam_b200_traj_estimate.txt", "r");
  int lines = 0;
  for (int c; f && (c = std::fgetc(f)) != EOF;) lines += c == '\n';
  if (f) std::fclose(f);
  ok = ok && lines == F + 1;
  // the other fixation strategies (graph_manager.cpp:911-937) give the same relative geometry
  double spread = 0;
  for (const char* strategy : {"previous", "largest_loop", "inaffected", "first"}) {
    gm.params.pose_relative_to = strategy;
    const double c2 = gm.optimizeGraph(5.0);
    spread = std::max(spread, std::fabs(c2 - chi2_again));
    const double* a = gm.estimates_[0].v;
    const double* b = gm.estimates_[F - 1].v;
    const double d = std::sqrt((b[0] - a[0]) * (b[0] - a[0]) + (b[1] - a[1]) * (b[1] - a[1]) + (b[2] - a[2]) * (b[2] - a[2]));
    if (std::fabs(d - step * (F - 1)) > 0.01) { std::printf("strategy %s: end-to-end distance %.4f\n", strategy, d); ok = 0; }
  }
  std::printf("chi2 %.6f, again %.6f, spread over fixation strategies %.3g\n", chi2, chi2_again, spread);
  rgbdslam_b200_shutdown();
  std::printf(ok ? "GRAPH MANAGER SHIM OK\n" : "GRAPH MANAGER SHIM FAILED\n");
  return ok ? 0 : 1;
}
