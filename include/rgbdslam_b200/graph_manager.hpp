// graph_manager.hpp -- header-only C++ shim of the reference's GraphManager call surface over the C ABI:
//   bool   GraphManager::addNode(Node*)                                  src/graph_manager.cpp:681-782 (firstNode :361-409)
//   bool   nodeComparisons(Node*, ...)                                   :421-658   (the <= 12 comparisons = ONE batched call)
//   QList<int> getPotentialEdgeTargetsWithDijkstra(...)                  :204-324
//   bool   addEdgeToG2O(const LoadedEdge3D&, Node*, Node*, bool, bool)   :811-898
//   double optimizeGraph(double break_criterion = -1, bool nonthreaded)  :900-1066  -> rgbdslam_b200_posegraph_optimize
//   unsigned pruneEdgesWithErrorAbove(float)                             :1106-1246 -> rgbdslam_b200_posegraph_chi2
//   void   saveTrajectory(filename)                                      graph_mgr_io.cpp:615-677 / logTransform misc.cpp:90-93
// Host logic only; every compute step is a C-ABI call.  The reference draws from the global rand(); here every draw comes
// from the library's counter-based generator keyed by (seed, node id).  g2o's HyperDijkstra (not under /root/reference) is
// restated in geodesicBall().  The Python mirror rgbdslam_v2_b200/graph_manager.py is the tested twin of this file.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <map>
#include <set>
#include <string>

#include "node.hpp"

namespace rgbdslam_b200 {

struct Pose7 {  // t (x, y, z) + unit quaternion (x, y, z, w): VertexSE3 estimate / EdgeSE3 measurement
  double v[7];
  static Pose7 Identity() { return Pose7{{0, 0, 0, 0, 0, 0, 1}}; }
};

inline void quatToRot(const double* q, double R[9]) {
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double x = q[0] / n, y = q[1] / n, z = q[2] / n, w = q[3] / n;
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}
inline void rotToQuat(const double R[9], double* q) {  // Eigen::Quaternion(Matrix3), normalised
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    double s = std::sqrt(tr + 1.0);
    q[3] = 0.5 * s;
    s = 0.5 / s;
    q[0] = (R[7] - R[5]) * s; q[1] = (R[2] - R[6]) * s; q[2] = (R[3] - R[1]) * s;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double s = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    q[i] = 0.5 * s;
    s = 0.5 / s;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * s;
    q[j] = (R[3 * j + i] + R[3 * i + j]) * s;
    q[k] = (R[3 * k + i] + R[3 * i + k]) * s;
  }
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int a = 0; a < 4; a++) q[a] /= n;
}
inline Pose7 poseFromIsometry(const Isometry3d& T) {  // column-major 4x4
  double R[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) R[3 * r + c] = T.m[4 * c + r];
  Pose7 p;
  p.v[0] = T.m[12]; p.v[1] = T.m[13]; p.v[2] = T.m[14];
  rotToQuat(R, p.v + 3);
  return p;
}
inline Pose7 compose(const Pose7& a, const Pose7& b) {  // a * b
  double Ra[9], Rb[9], R[9];
  quatToRot(a.v + 3, Ra);
  quatToRot(b.v + 3, Rb);
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) R[3 * r + c] = Ra[3 * r] * Rb[c] + Ra[3 * r + 1] * Rb[3 + c] + Ra[3 * r + 2] * Rb[6 + c];
  Pose7 p;
  for (int r = 0; r < 3; r++) p.v[r] = a.v[r] + Ra[3 * r] * b.v[0] + Ra[3 * r + 1] * b.v[1] + Ra[3 * r + 2] * b.v[2];
  rotToQuat(R, p.v + 3);
  return p;
}
inline Pose7 inverse(const Pose7& a) {
  double R[9];
  quatToRot(a.v + 3, R);
  Pose7 p;
  for (int r = 0; r < 3; r++) p.v[r] = -(R[r] * a.v[0] + R[3 + r] * a.v[1] + R[6 + r] * a.v[2]);
  p.v[3] = -a.v[3]; p.v[4] = -a.v[4]; p.v[5] = -a.v[5]; p.v[6] = a.v[6];
  return p;
}

// misc.cpp:272-315
inline void trafoSize(const Isometry3d& t, double& angle, double& dist) {
  angle = std::acos((t.m[0] + t.m[5] + t.m[10] - 1) / 2) * 180.0 / M_PI;
  dist = std::sqrt(t.m[12] * t.m[12] + t.m[13] * t.m[13] + t.m[14] * t.m[14]);
}

class GraphManager {
 public:
  struct Params {  // the ParameterServer entries this logic reads (parameter_server.cpp:85-123)
    int min_matches = 20, predecessor_candidates = 4, neighbor_candidates = 4, min_sampled_candidates = 4, geodesic_depth = 3;
    double min_translation_meter = 0.0, min_rotation_degree = 0.0, max_translation_meter = 1e10, max_rotation_degree = 360.0;
    bool keep_all_nodes = false, keep_good_nodes = false;
    int optimizer_skip_step = 1;
    double optimizer_iterations = 0.01, huber_delta = 1.0;
    bool valid_odometry = false;  // !odom_frame_name.empty()
    // fixationOfVertices strategy (graph_manager.cpp:911-937, parameter_server.cpp:118): "first" (default), "previous",
    // "largest_loop", "inaffected" (the reference's benchmark setting, test/test_settings.launch:94)
    std::string pose_relative_to = "first";
  };
  Params params;
  uint64_t seed = 0;

  std::map<int, Node*> graph_;                 // graph_manager.h:156 (owned after addNode, like the reference)
  std::vector<int> keyframe_ids_;
  std::vector<std::pair<int, int>> edges_;     // (id1 = older, id2 = newer) of cam_cam_edges_
  std::vector<Pose7> meas_;
  std::vector<Matrix6d> info_;
  std::vector<bool> active_;                   // false: removed by pruneEdgesWithErrorAbove
  std::map<int, Pose7> estimates_;             // VertexSE3 estimates by node id (vertex id == node id here)
  MatchingResult curr_best_result_;
  unsigned loop_closures_edges = 0, sequential_edges = 0;
  double last_chi2 = 0.0;
  int earliest_loop_closure_node_ = 0;         // graph_manager.h:356
  std::set<int> fixed_ids_;                    // vertices with setFixed(true) (persist between optimisations like g2o's flags)

  ~GraphManager() {
    for (auto& kv : graph_) delete kv.second;
  }

  bool isBigTrafo(const Isometry3d& t) const {
    double a, d;
    trafoSize(t, a, d);
    return d > params.min_translation_meter || a > params.min_rotation_degree;
  }
  bool isSmallTrafo(const Isometry3d& t, double seconds) const {
    if (seconds <= 0.0) return true;
    double a, d;
    trafoSize(t, a, d);
    return d / seconds < params.max_translation_meter && a / seconds < params.max_rotation_degree;
  }

  // ---- graph_manager.cpp:681-782
  bool addNode(Node* new_node) {
    // the reference gates on feature_locations_2d_ only (:687); nodes built from bare features (no 2-D keypoints) fall back
    // to their 3-D feature count
    const size_t nfeat = new_node->feature_locations_2d_.empty() ? new_node->feature_locations_3d_.size()
                                                                 : new_node->feature_locations_2d_.size();
    if ((int)nfeat < params.min_matches) return false;
    if (graph_.empty()) {
      firstNode(new_node);
      return true;
    }
    bool edge_to_last_keyframe_found = false;
    const bool found_match = nodeComparisons(new_node, edge_to_last_keyframe_found);
    if (found_match) {
      graph_[new_node->id_] = new_node;
      if (!edge_to_last_keyframe_found && earliest_loop_closure_node_ > keyframe_ids_.back())  // :731
        keyframe_ids_.push_back(new_node->id_ - 1);
      if (params.optimizer_skip_step > 0 && (int)estimates_.size() % params.optimizer_skip_step == 0) optimizeGraph();
    } else if (graph_.size() == 1) {
      // only one node so far and it has fewer features: the new node replaces it (:760-767)
      Node* first = graph_.begin()->second;
      const size_t nfirst = first->feature_locations_2d_.empty() ? first->feature_locations_3d_.size() : first->feature_locations_2d_.size();
      if (nfeat > nfirst) {
        resetGraph();
        firstNode(new_node);
        return true;
      }
    }
    return found_match;
  }

  void resetGraph() {  // GraphManager::resetGraph as far as this shim keeps state
    for (auto& kv : graph_) delete kv.second;
    graph_.clear(); keyframe_ids_.clear(); edges_.clear(); meas_.clear(); info_.clear(); active_.clear(); estimates_.clear();
    adj_.clear(); fixed_ids_.clear();
    curr_best_result_ = MatchingResult();
    loop_closures_edges = sequential_edges = 0;
    earliest_loop_closure_node_ = 0;
  }

  // ---- graph_manager.cpp:204-324
  std::vector<int> getPotentialEdgeTargetsWithDijkstra(const Node* new_node, int sequential_targets, int geodesic_targets,
                                                       int sampled_targets, int predecessor_id = -1, bool include_predecessor = false) {
    Rand rnd(seed, (uint64_t)new_node->id_);
    std::vector<int> ids;  // QList: push_front == insert at begin
    const int gsize = (int)graph_.size();
    if (predecessor_id < 0) predecessor_id = gsize - 1;
    if ((int)estimates_.size() <= sequential_targets + geodesic_targets + sampled_targets || estimates_.size() <= 1) {
      sequential_targets += geodesic_targets + sampled_targets;
      geodesic_targets = sampled_targets = 0;
      predecessor_id = gsize - 1;
    }
    for (int i = 1; i < sequential_targets + 1 && predecessor_id - i >= 0; i++) ids.push_back(predecessor_id - i);
    if (geodesic_targets > 0) {
      std::map<int, int> weights;
      int sum = 0;
      for (int vid : geodesicBall(predecessor_id, params.geodesic_depth)) {
        if (vid == new_node->id_) continue;  // (the reference can pick the new node itself here: a self-edge)
        if (!graph_.at(vid)->matchable_) continue;
        if (vid < predecessor_id - sequential_targets || (vid > predecessor_id && vid <= gsize - 1)) {
          weights[vid] = std::abs(predecessor_id - vid);
          sum += weights[vid];
        }
      }
      while ((int)ids.size() < sequential_targets + geodesic_targets && !weights.empty()) {
        const int pick = (int)(rnd() % (uint32_t)sum);
        int acc = 0;
        for (auto it = weights.begin(); it != weights.end(); ++it) {
          acc += it->second;
          if (acc > pick) {
            ids.insert(ids.begin(), it->first);
            sum -= it->second;
            weights.erase(it);
            break;
          }
        }
      }
    }
    if (sampled_targets > 0) {
      std::vector<int> pool;
      for (int k : keyframe_ids_)
        if (std::find(ids.begin(), ids.end(), k) == ids.end() && graph_.at(k)->matchable_) pool.push_back(k);
      while ((int)ids.size() < geodesic_targets + sampled_targets + sequential_targets && !pool.empty()) {
        const int i = (int)(rnd() % (uint32_t)pool.size());
        ids.insert(ids.begin(), pool[i]);
        pool[i] = pool.back();
        pool.pop_back();
      }
    }
    if (include_predecessor) ids.push_back(predecessor_id);
    return ids;
  }

  // ---- graph_manager.cpp:811-898
  bool addEdgeToG2O(const LoadedEdge3D& edge, Node* n1, Node* n2, bool largeEdge, bool set_estimate) {
    if (edge.id1 == edge.id2) return false;
    const bool v1 = estimates_.count(n1->id_) != 0, v2 = estimates_.count(n2->id_) != 0;
    if ((!v1 || !v2) && !largeEdge) return false;
    if (!v1 && !v2) return false;
    const Pose7 z = poseFromIsometry(edge.transform);
    if (!v1 && v2) {
      estimates_[n1->id_] = compose(estimates_[n2->id_], inverse(z));
      n1->vertex_id_ = n1->id_;
    } else if (!v2 && v1) {
      estimates_[n2->id_] = compose(estimates_[n1->id_], z);
      n2->vertex_id_ = n2->id_;
    } else if (set_estimate) {
      estimates_[n2->id_] = compose(estimates_[n1->id_], z);
    }
    edges_.push_back(std::make_pair(edge.id1, edge.id2));
    meas_.push_back(z);
    info_.push_back(edge.informationMatrix);
    active_.push_back(true);
    adj_[edge.id1].insert(edge.id2);
    adj_[edge.id2].insert(edge.id1);
    if (std::abs(edge.id1 - edge.id2) > params.predecessor_candidates) loop_closures_edges++;
    else sequential_edges++;
    earliest_loop_closure_node_ = std::min(earliest_loop_closure_node_, std::min(edge.id1, edge.id2));  // :894-895
    return true;
  }

  // ---- graph_manager.cpp:900-1066.  break_criterion: >= 1 iterations, (0, 1) relative chi2 improvement, < 0 the parameter
  double optimizeGraph(double break_criterion = -1.0, bool /*nonthreaded*/ = false) {
    std::vector<int> ids;
    std::vector<double> poses, meas, info;
    std::vector<uint8_t> fixed;
    std::vector<int32_t> ij;
    gather(ids, poses, fixed, ij, meas, info);
    if (ij.empty()) return 0.0;
    fixationOfVertices(ids, fixed);
    const double stop = break_criterion > 0.0 ? break_criterion : params.optimizer_iterations;  // :942
    double chi2 = 0;
    int it = 0, cg = 0;
    check(rgbdslam_b200_posegraph_optimize((int)ids.size(), poses.data(), fixed.data(), (int)ij.size() / 2, ij.data(), meas.data(),
                                           info.data(), stop, params.huber_delta, &chi2, &it, &cg),
          "posegraph_optimize");
    for (size_t k = 0; k < ids.size(); k++) std::memcpy(estimates_[ids[k]].v, &poses[7 * k], sizeof(double) * 7);
    // after the optimisation (:1031-1037): "inaffected" leaves every camera vertex fixed -- only vertices added afterwards are
    // free in the next run --, every other strategy un-fixes them all
    fixed_ids_.clear();
    if (params.pose_relative_to == "inaffected") fixed_ids_.insert(ids.begin(), ids.end());
    last_chi2 = chi2;
    return chi2;
  }

  // fixationOfVertices (graph_manager.cpp:911-937).  `fixed` comes in with the persistent flags (fixed_ids_ + the first node,
  // which firstNode fixes at the origin, :381).  "inaffected" has no branch there: the flags stay as the previous
  // optimisation left them; its Dijkstra-selected vertex subset (:969-977) is overridden by the second
  // initializeOptimization(cam_cam_edges_) (:989-992), so all edges are always active.
  void fixationOfVertices(const std::vector<int>& ids, std::vector<uint8_t>& fixed) const {
    const std::string& strategy = params.pose_relative_to;
    auto index_of = [&](int id) { return (int)(std::lower_bound(ids.begin(), ids.end(), id) - ids.begin()); };
    if (strategy == "previous" && graph_.size() > 2) {
      std::fill(fixed.begin(), fixed.end(), 0);
      fixed[index_of(graph_.at((int)graph_.size() - 2)->id_)] = 1;
    } else if (strategy == "largest_loop") {
      for (size_t k = 0; k < ids.size(); k++) fixed[k] = ids[k] < earliest_loop_closure_node_ ? 1 : 0;
    } else if (strategy == "first") {
      std::fill(fixed.begin(), fixed.end(), 0);
      fixed[index_of(graph_.at(0)->id_)] = 1;
    }
    // an optimisation without any fixed vertex has a gauge freedom; g2o then fixes nothing either, but its damped LM
    // still runs -- the PCG here needs one anchor
    if (std::find(fixed.begin(), fixed.end(), 1) == fixed.end() && !fixed.empty()) fixed[0] = 1;
  }

  // ---- graph_manager.cpp:1106-1246
  unsigned pruneEdgesWithErrorAbove(float thresh) {
    std::vector<int> ids;
    std::vector<double> poses, meas, info;
    std::vector<uint8_t> fixed;
    std::vector<int32_t> ij;
    std::vector<size_t> which;
    gather(ids, poses, fixed, ij, meas, info, &which);
    if (ij.empty()) return 0;
    std::vector<double> per_edge(which.size());
    double chi2 = 0;
    check(rgbdslam_b200_posegraph_chi2((int)ids.size(), poses.data(), (int)which.size(), ij.data(), meas.data(), info.data(),
                                       params.huber_delta, &chi2, per_edge.data()),
          "posegraph_chi2");
    std::map<int, int> degree;  // v->edges().size(): every edge ever added counts
    for (auto& e : edges_) { degree[e.first]++; degree[e.second]++; }
    unsigned counter = 0;
    for (size_t k = 0; k < which.size(); k++) {
      if (!(per_edge[k] > thresh)) continue;
      counter++;
      const size_t e = which[k];
      meas_[e] = Pose7::Identity();
      const int a = edges_[e].first, b = edges_[e].second;
      Matrix6d I;
      std::memset(&I, 0, sizeof(I));
      if (std::abs(a - b) != 1) {
        if (degree[a] > 1 && degree[b] > 1) { active_[e] = false; continue; }
        for (int i = 0; i < 6; i++) I.m[7 * i] = 1e-100;
      } else {
        for (int i = 0; i < 6; i++) I.m[7 * i] = 1.0;
      }
      info_[e] = I;
    }
    return counter;
  }

  // TUM trajectory "timestamp tx ty tz qx qy qz qw" (logTransform, misc.cpp:90-93)
  void saveTrajectory(const std::string& filename) const {
    FILE* f = std::fopen(filename.c_str(), "w");
    if (!f) throw std::runtime_error("cannot open " + filename);
    std::fprintf(f, "# TF Coordinate Frame ID: (data: )\n");
    for (auto& kv : estimates_) {
      const double* p = kv.second.v;
      std::fprintf(f, "%f %f %f %f %f %f %f %f\n", graph_.at(kv.first)->stamp_, p[0], p[1], p[2], p[3], p[4], p[5], p[6]);
    }
    std::fclose(f);
  }

 private:
  std::map<int, std::set<int>> adj_;

  struct Rand {  // rand() stand-in: the library's splitmix64 counter generator, stream 0xC0
    uint64_t key;
    uint32_t ctr = 0;
    static uint64_t mix(uint64_t z) {
      z += 0x9E3779B97F4A7C15ull;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      return z ^ (z >> 31);
    }
    Rand(uint64_t seed, uint64_t node) : key(mix(seed ^ mix(node))) {}
    uint32_t operator()() { return (uint32_t)(mix(key ^ ((0xC0ull << 32) | ctr++)) >> 33); }
  };

  // g2o::HyperDijkstra::shortestPaths(v, UniformCostFunction, maxDistance) + visited(): the source and every vertex whose
  // hop count is < maxDistance
  std::set<int> geodesicBall(int source, double max_distance) const {
    std::map<int, int> dist;
    dist[source] = 0;
    std::vector<int> frontier(1, source);
    while (!frontier.empty()) {
      std::vector<int> next;
      for (int u : frontier) {
        auto it = adj_.find(u);
        if (it == adj_.end()) continue;
        for (int z : it->second) {
          const int d = dist[u] + 1;
          if (!dist.count(z) && d < max_distance) {
            dist[z] = d;
            next.push_back(z);
          }
        }
      }
      frontier.swap(next);
    }
    std::set<int> out;
    for (auto& kv : dist) out.insert(kv.first);
    return out;
  }

  void firstNode(Node* n) {  // :361-409
    n->id_ = (int)graph_.size();
    n->vertex_id_ = n->id_;
    graph_[n->id_] = n;
    estimates_[n->id_] = Pose7::Identity();
    adj_[n->id_];
    keyframe_ids_.push_back(n->id_);
  }

  // ---- graph_manager.cpp:421-658
  bool nodeComparisons(Node* new_node, bool& edge_to_keyframe) {
    const int num_keypoints = (int)std::max(new_node->feature_locations_2d_.size(), new_node->feature_locations_3d_.size());
    if (num_keypoints < params.min_matches && !params.keep_all_nodes) return false;
    new_node->id_ = (int)graph_.size();
    earliest_loop_closure_node_ = new_node->id_;  // :444
    const size_t num_edges_before = edges_.size();
    edge_to_keyframe = false;
    const int sequentially_previous_id = graph_.rbegin()->second->id_;
    curr_best_result_ = MatchingResult();
    bool predecessor_matched = false;
    if (params.min_translation_meter > 0.0 || params.min_rotation_degree > 0.0) {  // initial comparison :458-513
      Node* prev = graph_[(int)graph_.size() - 1];
      MatchingResult mr = new_node->matchNodePair(prev, seed, 64 * (int64_t)new_node->id_ + 63);
      if (mr.edge.id1 >= 0 && mr.edge.id2 >= 0) {
        const double dt = new_node->stamp_ - prev->stamp_;
        if (!isBigTrafo(mr.edge.transform) || !isSmallTrafo(mr.edge.transform, dt)) {
          curr_best_result_ = mr;
          return false;
        }
        if (!addEdgeToG2O(mr.edge, prev, new_node, true, true)) return false;
        graph_[new_node->id_] = new_node;
        if (std::find(keyframe_ids_.begin(), keyframe_ids_.end(), mr.edge.id1) != keyframe_ids_.end()) edge_to_keyframe = true;
        prev->valid_tf_estimate_ = true;
        curr_best_result_ = mr;
        predecessor_matched = true;
      }
    }
    const int seq_cand = params.predecessor_candidates - 1, geod_cand = params.neighbor_candidates,
              samp_cand = params.min_sampled_candidates;
    std::vector<int> targets = predecessor_matched
        ? getPotentialEdgeTargetsWithDijkstra(new_node, seq_cand, geod_cand, samp_cand, curr_best_result_.edge.id1)
        : getPotentialEdgeTargetsWithDijkstra(new_node, seq_cand, geod_cand, samp_cand, sequentially_previous_id, true);
    std::vector<const Node*> olds;
    for (int t : targets) olds.push_back(graph_[t]);
    // QtConcurrent::blockingMapped(nodes_to_comp, &Node::matchNodePair) (:548) as one batched call
    std::vector<MatchingResult> results = Node::matchNodePairs(new_node, olds, seed, 64 * (int64_t)new_node->id_);
    for (size_t i = 0; i < results.size(); i++) {
      MatchingResult& mr = results[i];
      if (mr.edge.id1 < 0) continue;
      Node* old = graph_[mr.edge.id1];
      const double dt = new_node->stamp_ - old->stamp_;
      const bool more = mr.inlier_matches.size() > curr_best_result_.inlier_matches.size();
      if (isSmallTrafo(mr.edge.transform, dt) && addEdgeToG2O(mr.edge, old, new_node, isBigTrafo(mr.edge.transform), more)) {
        graph_[new_node->id_] = new_node;
        if (mr.edge.id1 == mr.edge.id2 - 1) predecessor_matched = true;
        old->valid_tf_estimate_ = true;
        if (more) curr_best_result_ = mr;
        if (std::find(keyframe_ids_.begin(), keyframe_ids_.end(), mr.edge.id1) != keyframe_ids_.end()) edge_to_keyframe = true;
      }
    }
    const bool found_trafo = edges_.size() != num_edges_before;
    const bool keep_anyway = params.keep_all_nodes || ((int)new_node->feature_locations_3d_.size() > params.min_matches && params.keep_good_nodes);
    const double time_delta_sec = std::fabs(new_node->stamp_ - graph_[sequentially_previous_id]->stamp_);
    if ((!found_trafo && params.valid_odometry) || (!found_trafo && keep_anyway) || (!predecessor_matched && time_delta_sec < 0.1)) {
      LoadedEdge3D odom_edge;  // constant position assumption :636-655
      odom_edge.id1 = sequentially_previous_id;
      odom_edge.id2 = new_node->id_;
      std::memset(&odom_edge.transform, 0, sizeof(odom_edge.transform));
      std::memset(&odom_edge.informationMatrix, 0, sizeof(odom_edge.informationMatrix));
      for (int i = 0; i < 4; i++) odom_edge.transform.m[5 * i] = 1.0;
      // information = I / time_delta_sec (:647); nodes without stamps (dt == 0) would get an infinite information matrix
      // (NaN in the linearisation): the time delta is clamped to 1 ms
      for (int i = 0; i < 6; i++) odom_edge.informationMatrix.m[7 * i] = 1.0 / std::max(time_delta_sec, 1e-3);
      addEdgeToG2O(odom_edge, graph_[sequentially_previous_id], new_node, true, true);
      graph_[new_node->id_] = new_node;
      new_node->valid_tf_estimate_ = false;
      MatchingResult mr;
      mr.edge = odom_edge;
      curr_best_result_ = mr;
    }
    return edges_.size() > num_edges_before;
  }

  void gather(std::vector<int>& ids, std::vector<double>& poses, std::vector<uint8_t>& fixed, std::vector<int32_t>& ij,
              std::vector<double>& meas, std::vector<double>& info, std::vector<size_t>* which = nullptr) const {
    std::map<int, int> index;
    for (auto& kv : estimates_) {
      index[kv.first] = (int)ids.size();
      ids.push_back(kv.first);
      poses.insert(poses.end(), kv.second.v, kv.second.v + 7);
    }
    fixed.assign(ids.size(), 0);
    if (!fixed.empty()) fixed[0] = 1;  // firstNode: reference_pose->setFixed(true) (:381)
    for (size_t k = 0; k < ids.size(); k++)
      if (fixed_ids_.count(ids[k])) fixed[k] = 1;
    for (size_t e = 0; e < edges_.size(); e++) {
      if (!active_[e]) continue;
      ij.push_back(index[edges_[e].first]);
      ij.push_back(index[edges_[e].second]);
      meas.insert(meas.end(), meas_[e].v, meas_[e].v + 7);
      info.insert(info.end(), info_[e].m, info_[e].m + 36);
      if (which) which->push_back(e);
    }
  }
};

}  // namespace rgbdslam_b200
