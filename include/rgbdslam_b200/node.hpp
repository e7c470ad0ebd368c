// node.hpp -- header-only C++ shim that keeps the reference's call sites compiling against the C ABI.
//
// Mirrors, for the frame-pair hot path only:
//   LoadedEdge3D                 src/edge.h:24-32
//   MatchingResult               src/matching_result.h:24-46
//   Node (feature members, matchNodePair, featureMatching-free ctor from features)   src/node.h:64-178
//   bruteForceSearchORB          src/features.h:13, src/features.cpp:168-182
// The reference types Eigen::Matrix4f / Eigen::Isometry3d / cv::DMatch / cv::KeyPoint are replaced by
// layout-compatible PODs (column-major float[16] etc.) so this header has no third-party dependency; a
// maintainer who has Eigen/OpenCV maps them with Eigen::Map / reinterpret_cast.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../rgbdslam_b200.h"
#include "features.hpp"

namespace rgbdslam_b200 {

struct Matrix4f {  // == Eigen::Matrix4f storage (column-major)
  float m[16];
  static Matrix4f Identity() {
    Matrix4f r;
    for (int i = 0; i < 16; i++) r.m[i] = (i % 5 == 0) ? 1.f : 0.f;
    return r;
  }
  float operator()(int row, int col) const { return m[4 * col + row]; }
};
struct Isometry3d {  // == Eigen::Isometry3d storage (4x4 column-major double)
  double m[16];
};
struct Matrix6d {
  double m[36];
};
struct Vector4f {
  float x, y, z, w;
};
typedef rgbdslam_b200_dmatch DMatch;  // == cv::DMatch  (KeyPoint: features.hpp)

struct LoadedEdge3D {  // src/edge.h:24-32
  int id1, id2;
  Isometry3d transform;
  Matrix6d informationMatrix;
};

class MatchingResult {  // src/matching_result.h:24-46
 public:
  MatchingResult()
      : rmse(0.0), ransac_trafo(Matrix4f::Identity()), final_trafo(Matrix4f::Identity()), icp_trafo(Matrix4f::Identity()),
        inlier_points(0), outlier_points(0), occluded_points(0), all_points(0) {
    edge.id1 = edge.id2 = -1;
    std::memset(&edge.transform, 0, sizeof(edge.transform));
    std::memset(&edge.informationMatrix, 0, sizeof(edge.informationMatrix));
  }
  std::vector<DMatch> inlier_matches;
  std::vector<DMatch> all_matches;
  LoadedEdge3D edge;
  float rmse;
  Matrix4f ransac_trafo, final_trafo, icp_trafo;
  unsigned int inlier_points, outlier_points, occluded_points, all_points;
};

inline void check(int rc, const char* what) {
  if (rc != 0) throw std::runtime_error(std::string(what) + ": " + rgbdslam_b200_last_error());
}

// Fill a MatchingResult from the flat ABI record (what node.cpp:1334-1339 does with the RANSAC output).
inline MatchingResult to_matching_result(const rgbdslam_b200_pair_result& r, const DMatch* all, const DMatch* inl) {
  MatchingResult mr;
  mr.all_matches.assign(all, all + r.n_all_matches);
  mr.inlier_matches.assign(inl, inl + r.n_inliers);
  mr.rmse = r.rmse;
  std::memcpy(mr.ransac_trafo.m, r.ransac_trafo, sizeof(r.ransac_trafo));
  mr.edge.id1 = r.id1;
  mr.edge.id2 = r.id2;
  mr.inlier_points = r.inlier_points;  // environment measurement model (node.cpp:1551-1554), 0 when it is off
  mr.outlier_points = r.outlier_points;
  mr.occluded_points = r.occluded_points;
  mr.all_points = r.all_points;
  if (r.id1 >= 0) {
    mr.final_trafo = mr.ransac_trafo;                                               // node.cpp:1334
    for (int i = 0; i < 16; i++) mr.edge.transform.m[i] = (double)r.ransac_trafo[i];  // node.cpp:1339
    for (int i = 0; i < 6; i++) mr.edge.informationMatrix.m[7 * i] = r.info_scale;    // node.cpp:1335
  }
  return mr;
}

class Node {  // src/node.h: the members the hot path reads + matchNodePair
 public:
  int id_ = -1, seq_id_ = -1, vertex_id_ = -1;
  double stamp_ = 0.0;           // header_.stamp in seconds
  bool matchable_ = true, valid_tf_estimate_ = true;  // node.h:160,178
  mutable int initial_node_matches_ = 0;              // node.h:207: accepted transformations of this node (max_connections)
  std::vector<KeyPoint> feature_locations_2d_;  // node.h:167
  std::vector<Vector4f> feature_locations_3d_;  // node.h:174
  std::vector<uint8_t> feature_descriptors_;    // N x 32 (cv::Mat CV_8U rows), node.h:169

  Node() {}
  // The reference's constructor, argument for argument (node.h:64-70, call site openni_listener.cpp:779):
  //   Node(visual, depth, detection_mask, cam_info, depth_header, detector, extractor)
  // visual CV_8UC1, depth CV_32FC1 metres (NaN = invalid), detection_mask CV_8UC1 non-zero at potential keypoint locations
  // (node.h:61-63).  detector / extractor: createDetector / createDescriptorExtractor (features.hpp); the whole constructor
  // -- detect, removeDepthless, retainBest, compute, projectTo3D -- runs as one rgbdslam_b200_nodes_create call, the
  // extractor argument only documents the pairing (ORB with ORB).  id_ stays -1 until GraphManager::addNode assigns it.
  Node(const Mat& visual, const Mat& depth, const Mat& detection_mask, const CameraInfoConstPtr& cam_info, myHeader depth_header,
       Ptr<Feature2D> detector, Ptr<DescriptorExtractor> extractor)
      : stamp_(depth_header.stamp) {
    if (!detector || !detector->handle()) throw std::invalid_argument("Node: detector must come from createDetector(\"ORB\")");
    if (!extractor) throw std::invalid_argument("Node: null extractor");
    if (visual.type() != RB_8UC1 || depth.type() != RB_32FC1 || depth.rows != visual.rows || depth.cols != visual.cols)
      throw std::invalid_argument("Node: visual must be CV_8UC1 and depth CV_32FC1 of the same size");
    seq_id_ = (int)depth_header.seq;
    std::vector<uint8_t> tg, tm;
    std::vector<float> td;
    const uint8_t* g = detail::packed<uint8_t>(visual, tg);
    const float* d = detail::packed<float>(depth, td);
    const uint8_t* m = detection_mask.empty() ? nullptr : detail::packed<uint8_t>(detection_mask, tm);
    const CameraInfo ci = cam_info ? *cam_info : CameraInfo();
    const float K4[4] = {(float)ci.K[0], (float)ci.K[4], (float)ci.K[2], (float)ci.K[5]};  // node.cpp:913-916
    construct(g, d, m, visual.cols, visual.rows, K4, detector->handle());
  }
  // Node(visual, depth, detection_mask, cam_info, depth_header, detector, extractor) (node.cpp:101-240): detect, filter,
  // describe, back-project on the device; the public feature members are filled from the result.  `detector` is the handle of
  // rgbdslam_b200_detector_create -- the counterpart of the detector_ / extractor_ pair OpenNIListener keeps
  // (openni_listener.cpp:130-132), whose adaptive per-cell thresholds live across frames.
  Node(const uint8_t* gray, const float* depth_m, const uint8_t* detection_mask, int w, int h, const float K4[4], int id,
       uint64_t detector, double stamp = 0.0)
      : id_(id), stamp_(stamp) {
    construct(gray, depth_m, detection_mask, w, h, K4, detector);
  }
  // Construct from already extracted features (what the reference ctor node.cpp:101-240 leaves behind).
  Node(int id, const std::vector<uint8_t>& desc, const std::vector<Vector4f>& xyz) : id_(id) {
    feature_descriptors_ = desc;
    feature_locations_3d_ = xyz;
    upload();
  }
  ~Node() {  // Node::~Node (node.cpp:371)
    if (handle_) rgbdslam_b200_node_destroy(handle_);
  }
  Node(const Node&) = delete;
  Node& operator=(const Node&) = delete;

  void construct(const uint8_t* gray, const float* depth_m, const uint8_t* detection_mask, int w, int h, const float K4[4],
                 uint64_t detector) {
    int32_t n = 0, id32 = id_ < 0 ? 0 : id_;
    check(rgbdslam_b200_nodes_create(detector, 1, gray, depth_m, detection_mask, w, h, K4, &id32, &handle_, &n), "nodes_create");
    feature_locations_2d_.resize(n);
    feature_locations_3d_.resize(n);
    feature_descriptors_.resize((size_t)n * 32);
    if (n > 0) {
      check(rgbdslam_b200_node_download_keypoints(handle_, feature_locations_2d_.data()), "node_download_keypoints");
      check(rgbdslam_b200_node_download(handle_, feature_descriptors_.data(), reinterpret_cast<float*>(feature_locations_3d_.data())),
            "node_download");
    }
  }

  void upload() {
    if (handle_) rgbdslam_b200_node_destroy(handle_);
    handle_ = 0;
    // the device copy only needs a non-negative id (negative ids in a result mean "no transformation"); the ids of the
    // MatchingResult come from the host objects, which addNode may renumber (graph_manager.cpp:434)
    check(rgbdslam_b200_node_create_from_features(id_ < 0 ? 0 : id_, feature_descriptors_.data(),
                                                  reinterpret_cast<const float*>(feature_locations_3d_.data()),
                                                  (int)feature_locations_3d_.size(), &handle_),
          "node_create_from_features");
  }
  uint64_t handle() const { return handle_; }

  // MatchingResult Node::matchNodePair(const Node* older_node)  (node.cpp:1305).  Never throws on a failed
  // match: failure is edge.id1 == edge.id2 == -1 (node.cpp:1420), as in the reference.
  MatchingResult matchNodePair(const Node* older_node, uint64_t seed = 0, int64_t pair_index = 0) const {
    std::vector<MatchingResult> v = matchNodePairs(this, std::vector<const Node*>(1, older_node), seed, pair_index);
    return v[0];
  }

  // The QtConcurrent::blockingMapped(nodes_to_comp, bind(&Node::matchNodePair, new_node, _1)) fan-out of
  // graph_manager.cpp:548 as ONE batched launch.
  static std::vector<MatchingResult> matchNodePairs(const Node* newer, const std::vector<const Node*>& older,
                                                    uint64_t seed = 0, int64_t first_pair_index = 0) {
    const int n = (int)older.size();
    std::vector<MatchingResult> out(n);
    if (n == 0) return out;
    // the match arrays of the C ABI hold params.max_matches entries per pair: ask the library, never assume the default
    rgbdslam_b200_params prm;
    if (rgbdslam_b200_get_params(&prm) != 0) return out;  // not initialised: invalid edges, matchNodePair never throws
    const int mm = prm.max_matches;
    std::vector<uint64_t> a(n, newer->handle_), b(n);
    for (int i = 0; i < n; i++) b[i] = older[i]->handle_;
    std::vector<rgbdslam_b200_pair_result> res(n);
    std::vector<DMatch> all((size_t)n * mm), inl((size_t)n * mm);
    int rc = rgbdslam_b200_match_pairs(a.data(), b.data(), n, seed, first_pair_index, res.data(), all.data(), inl.data());
    if (rc != 0) return out;  // invalid edges (-1,-1): matchNodePair never throws (node.cpp:1308,1424)
    for (int i = 0; i < n; i++) {
      // "enough is enough" (node.cpp:1310-1312), in the order the reference's sequential loop would meet the candidates:
      // once the node has more than max_connections accepted transformations the remaining comparisons return empty
      if (max_connections() > 0 && newer->initial_node_matches_ > max_connections()) continue;
      out[i] = to_matching_result(res[i], &all[(size_t)i * mm], &inl[(size_t)i * mm]);
      if (res[i].id1 >= 0) {  // node.cpp:1337-1338: edge.id1 = older_node->id_, edge.id2 = this->id_
        out[i].edge.id1 = older[i]->id_;
        out[i].edge.id2 = newer->id_;
        ++newer->initial_node_matches_;  // node.cpp:1417
      }
    }
    return out;
  }

  static int& max_connections() {  // parameter max_connections (parameter_server.cpp:104), -1 = unlimited
    static int v = -1;
    return v;
  }

 private:
  uint64_t handle_ = 0;
};

}  // namespace rgbdslam_b200

// int bruteForceSearchORB(const uint64_t* v, const uint64_t* search_array, const unsigned int& size, int& result_index)
// -- src/features.h:13.  Single-query form kept for source compatibility; batch callers should use
// rgbdslam_b200_brute_force_orb directly (one launch for all query rows).
inline int bruteForceSearchORB(const uint64_t* v, const uint64_t* search_array, const unsigned int& size, int& result_index) {
  int32_t idx = -1, hd = 257;
  if (rgbdslam_b200_brute_force_orb(v, 1, search_array, (int)size, &idx, &hd) != 0)
    throw std::runtime_error(rgbdslam_b200_last_error());
  result_index = idx;
  return hd;
}
