// Compiles the reference-shaped call sites against the shim (CPU: compile+link; GPU: run).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "rgbdslam_b200/node.hpp"

using namespace rgbdslam_b200;

static uint64_t s = 88172645463325252ull;
static uint32_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 32); }

int main() {
  rgbdslam_b200_params p;
  rgbdslam_b200_default_params(&p);
  p.depth_cov_z0 = 2.0;
  if (rgbdslam_b200_init(0, &p) != 0) {
    std::printf("init failed (expected without a GPU): %s\n", rgbdslam_b200_last_error());
    return 77;
  }
  const int n = 500;
  std::vector<uint8_t> d_old(n * 32), d_new(n * 32);
  std::vector<Vector4f> x_old(n), x_new(n);
  for (int i = 0; i < n; i++) {
    for (int k = 0; k < 32; k++) d_old[i * 32 + k] = (uint8_t)rnd();
    float z = 1.f + (rnd() % 1000) / 400.f, u = 40.f + rnd() % 560, v = 40.f + rnd() % 400;
    x_old[i] = {(u - 319.5f) * z / 525.f, (v - 239.5f) * z / 525.f, z, 1.f};
  }
  // newer frame = older frame shifted by (+2 cm, -1 cm, 3 cm): p_new = p_old - t  =>  T(new->old) = +t
  for (int i = 0; i < n; i++) {
    for (int k = 0; k < 32; k++) d_new[i * 32 + k] = d_old[i * 32 + k] ^ (uint8_t)(1u << (rnd() % 8)) * (k % 4 == 0);
    x_new[i] = {x_old[i].x - 0.02f, x_old[i].y + 0.01f, x_old[i].z - 0.03f, 1.f};
  }
  Node* older_node = new Node(3, d_old, x_old);
  Node* new_node = new Node(4, d_new, x_new);
  MatchingResult mr = new_node->matchNodePair(older_node);  // graph_manager.cpp:466
  std::printf("edge %d -> %d, inliers %zu / %zu, rmse %.3f, t = %.4f %.4f %.4f, info %.1f\n", mr.edge.id1, mr.edge.id2,
              mr.inlier_matches.size(), mr.all_matches.size(), mr.rmse, mr.final_trafo(0, 3), mr.final_trafo(1, 3),
              mr.final_trafo(2, 3), mr.edge.informationMatrix.m[0]);
  int ok = mr.edge.id1 == 3 && mr.edge.id2 == 4 && mr.inlier_matches.size() > 250;
  ok = ok && std::abs(mr.final_trafo(0, 3) - 0.02f) < 1e-3f && std::abs(mr.final_trafo(1, 3) + 0.01f) < 1e-3f &&
       std::abs(mr.final_trafo(2, 3) - 0.03f) < 1e-3f;
  int idx = -5;
  int hd = bruteForceSearchORB(reinterpret_cast<const uint64_t*>(d_new.data()), reinterpret_cast<const uint64_t*>(d_old.data()), n, idx);
  std::printf("bruteForceSearchORB: hd %d idx %d\n", hd, idx);
  ok = ok && idx == 0 && hd <= 8;
  delete new_node;
  delete older_node;
  {
    // the reference's factory + constructor call sites (openni_listener.cpp:130-132, 779) on a synthetic textured frame
    Ptr<Feature2D> detector_(createDetector("ORB"));
    Ptr<DescriptorExtractor> extractor_ = createDescriptorExtractor("ORB");
    const int W = 640, H = 480;
    std::vector<uint8_t> img((size_t)W * H), msk((size_t)W * H, 255);
    std::vector<float> dep((size_t)W * H, 2.0f);
    for (int y = 0; y < H; y++)
      for (int x = 0; x < W; x++) img[(size_t)y * W + x] = (uint8_t)(((x / 9 + y / 7) % 2) * 140 + (rnd() % 60));
    Mat visual(H, W, RB_8UC1, img.data()), depth(H, W, RB_32FC1, dep.data()), detection_mask(H, W, RB_8UC1, msk.data());
    CameraInfoConstPtr cam_info(new CameraInfo());
    myHeader depth_header;
    depth_header.stamp = 12.5;
    Node* n = new Node(visual, depth, detection_mask, cam_info, depth_header, detector_, extractor_);
    std::printf("Node(visual, depth, mask, cam_info, header, detector, extractor): %zu features, stamp %.1f\n",
                n->feature_locations_2d_.size(), n->stamp_);
    ok = ok && n->feature_locations_2d_.size() > 100 && n->feature_locations_2d_.size() == n->feature_locations_3d_.size() &&
         n->feature_descriptors_.size() == 32 * n->feature_locations_2d_.size() && n->stamp_ == 12.5;
    // detect() / compute() as separate calls (node.cpp:160,202)
    std::vector<KeyPoint> kps;
    detector_->detect(visual, kps, detection_mask);
    std::vector<uint8_t> desc;
    const size_t n_det = kps.size();
    extractor_->compute(visual, kps, desc);
    std::printf("detect: %zu keypoints, compute kept %zu\n", n_det, kps.size());
    ok = ok && n_det > 100 && kps.size() <= n_det && desc.size() == 32 * kps.size();
    bool threw = false;
    try { createDetector("SURF"); } catch (const std::invalid_argument&) { threw = true; }
    ok = ok && threw && createDetector("SIFTGPU") == nullptr;
    delete n;
  }
  rgbdslam_b200_shutdown();
  std::printf(ok ? "SHIM OK\n" : "SHIM FAILED\n");
  return ok ? 0 : 1;
}
