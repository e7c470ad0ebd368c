// api_orb.cu -- C ABI of the Node constructor path (detect / describe / back-project), host orchestration.
//   rgbdslam_b200_detector_*   == createDetector("ORB") + its persistent adaptive-threshold state
//                                 (features.cpp:63-113, feature_adjuster.cpp:131-150, openni_listener.cpp:130-132)
//   rgbdslam_b200_orb_detect   == detector->detect(gray, keypoints, mask)             (node.cpp:160)
//   rgbdslam_b200_orb_compute  == extractor->compute(gray, keypoints, descriptors)    (node.cpp:202)
//   rgbdslam_b200_nodes_create == Node::Node(visual, depth, mask, cam_info, ...)      (node.cpp:101-240), batched
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "kernels.h"
#include "orb_host.h"
#include "state.h"

namespace rb200 {

struct Detector {
  static constexpr uint32_t kMagic = 0x44455443u;  // 'DETC'
  uint32_t magic = kMagic;
  double thresh[kOrbMaxCells];
  Detector() {
    for (int i = 0; i < kOrbMaxCells; i++) thresh[i] = 20.0;  // new DetectorAdjuster("ORB", 20)  features.cpp:92
  }
};

struct OrbCtx {
  bool ready = false;
  int W = 0, H = 0, grid = 0, max_kp = 0;
  OrbGeom g;
  std::vector<int16_t> h_ofs;
  std::vector<uint16_t> h_w1;
  DevBuf d_ofs, d_w1;
  OrbTables tab;
  int max_per_cell = 0, min_cell = 0, max_cell = 0, kp_stride = 0;
  int chunk = 0;  // frames per pass
  DevBuf gray, mask, depth, cell_img, cell_mask, score, cand, cand_count, hist, thr, resp, cell_out, cell_out_count, scratch, kp,
      xyz, n, pyr_raw, pyr_blur, desc;
  void release() {
    DevBuf* all[] = {&d_ofs, &d_w1, &gray, &mask, &depth, &cell_img, &cell_mask, &score, &cand, &cand_count, &hist, &thr, &resp,
                     &cell_out, &cell_out_count, &scratch, &kp, &xyz, &n, &pyr_raw, &pyr_blur, &desc};
    for (DevBuf* b : all) b->release();
    ready = false;
  }
};
static OrbCtx g_orb;

static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline float layer_scale(int level) { return (float)std::pow((double)1.2f, (double)level); }  // ORB getScale()

// resize tables src_n -> dst_n (INTER_LINEAR_EXACT): first tap index and weight of the second tap (x256)
static void build_table(int src_n, int dst_n, std::vector<int16_t>& ofs, std::vector<uint16_t>& w1) {
  const double scale = (double)src_n / dst_n;
  for (int d = 0; d < dst_n; d++) {
    const double f = (d + 0.5) * scale - 0.5;
    int i = (int)std::floor(f);
    double fr = f - i;
    if (i < 0) { i = 0; fr = 0; }
    if (i >= src_n - 1) { i = src_n - 1; fr = 0; }
    ofs.push_back((int16_t)i);
    w1.push_back((uint16_t)std::lrint(fr * 256.0));  // cvRound: half to even
  }
}

static int orb_prepare(int W, int H, int nframes_hint) {
  State& s = g_state;
  OrbCtx& o = g_orb;
  const int grid = s.params.detector_grid_resolution > 1 ? s.params.detector_grid_resolution : 1;
  const int K = s.params.max_keypoints;
  if (o.ready && o.W == W && o.H == H && o.grid == grid && o.max_kp == K) return 0;
  if (grid * grid > kOrbMaxCells) {
    set_error("detector_grid_resolution > 4 is not supported");
    return RGBDSLAM_B200_ERR_ARG;
  }
  if (W < 96 || H < 96 || W > 1023 || H > 1023) {
    set_error("image size must be within [96, 1023] in both dimensions");
    return RGBDSLAM_B200_ERR_ARG;
  }
  o.release();
  OrbGeom& g = o.g;
  memset(&g, 0, sizeof(g));
  g.W = W; g.H = H; g.grid = grid; g.ncells = grid * grid;
  o.h_ofs.clear(); o.h_w1.clear();
  // table cache per (base length) chain
  struct Chain { int n0; int off[kOrbLevels]; };
  std::vector<Chain> chains;
  auto chain_for = [&](int n0) -> const Chain& {
    for (const Chain& c : chains) if (c.n0 == n0) return c;
    Chain c;
    c.n0 = n0;
    int prev = n0;
    c.off[0] = 0;
    for (int l = 1; l < kOrbLevels; l++) {
      const int nl = cv_round_f((float)n0 / layer_scale(l));
      c.off[l] = (int)o.h_ofs.size();
      build_table(prev, nl, o.h_ofs, o.h_w1);
      prev = nl;
    }
    chains.push_back(c);
    return chains.back();
  };
  // grid cells (feature_adjuster.cpp:286-303; edgeThreshold = 31: feature_adjuster.h:110)
  const int edge = 31;
  int off = 0;
  for (int i = 0; i < grid; i++)
    for (int j = 0; j < grid; j++) {
      const int c = j + i * grid;
      int y0 = 0, y1 = H, x0 = 0, x1 = W;
      if (grid > 1) {
        y0 = std::max((i * H) / grid - edge, 0);
        y1 = std::min(H, ((i + 1) * H) / grid + edge);
        x0 = std::max((j * W) / grid - edge, 0);
        x1 = std::min(W, ((j + 1) * W) / grid + edge);
      }
      g.cell_x0[c] = x0;
      g.cell_y0[c] = y0;
      const int w0 = x1 - x0, h0 = y1 - y0;
      const Chain cx = chain_for(w0), cy = chain_for(h0);
      for (int l = 0; l < kOrbLevels; l++) {
        OrbPlane& p = g.cell[c][l];
        p.scale = layer_scale(l);
        p.w = l == 0 ? w0 : cv_round_f((float)w0 / p.scale);
        p.h = l == 0 ? h0 : cv_round_f((float)h0 / p.scale);
        p.off = off;
        p.tx = cx.off[l];
        p.ty = cy.off[l];
        off += (p.w * p.h + 15) / 16 * 16;
        if (p.w < 40 || p.h < 40) {
          set_error("image too small for an 8-level ORB pyramid per grid cell");
          return RGBDSLAM_B200_ERR_ARG;
        }
      }
    }
  g.cell_bytes = off;
  {
    const Chain cx = chain_for(W), cy = chain_for(H);
    int foff = 0;
    for (int l = 0; l < kOrbLevels; l++) {
      OrbPlane& p = g.full[l];
      p.scale = layer_scale(l);
      p.w = l == 0 ? W : cv_round_f((float)W / p.scale);
      p.h = l == 0 ? H : cv_round_f((float)H / p.scale);
      p.off = foff;
      p.tx = cx.off[l];
      p.ty = cy.off[l];
      foff += (p.w * p.h + 15) / 16 * 16;
    }
    g.full_bytes = foff;
  }
  {  // ORB per-level quotas for nfeatures = 10000 (only used to detect a binding retainBest)
    const float factor = 1.f / 1.2f;
    float nd = 10000 * (1 - factor) / (1 - (float)std::pow((double)factor, (double)kOrbLevels));
    int sum = 0;
    for (int l = 0; l < kOrbLevels - 1; l++) {
      g.n_per_level[l] = cv_round_f(nd);
      sum += g.n_per_level[l];
      nd *= factor;
    }
    g.n_per_level[kOrbLevels - 1] = std::max(10000 - sum, 0);
  }
  int umax[kOrbHalfPatch + 2];
  {
    const int hp = kOrbHalfPatch;
    const int vmax = (int)std::floor(hp * std::sqrt(2.f) / 2 + 1), vmin = (int)std::ceil(hp * std::sqrt(2.f) / 2);
    for (int v = 0; v <= hp + 1; v++) umax[v] = 0;
    for (int v = 0; v <= vmax; ++v) umax[v] = (int)std::lrint(std::sqrt((double)hp * hp - v * v));
    for (int v = hp, v0 = 0; v >= vmin; --v) {
      while (umax[v0] == umax[v0 + 1]) ++v0;
      umax[v] = v0;
      ++v0;
    }
  }
  // adjustedGridWrapper (features.cpp:43-60)
  const int mn = K, mx = (int)(K * 1.5);
  if (grid > 1) {
    o.min_cell = (int)std::lround(mn / (float)g.ncells);
    o.max_cell = (int)std::lround(mx / (float)g.ncells);
    o.max_per_cell = mx / g.ncells;  // maxTotalKeypoints / (gridRows*gridCols), feature_adjuster.cpp:292
  } else {
    o.min_cell = mn;
    o.max_cell = mx;
    o.max_per_cell = kOrbFrameCap;  // no keepStrongest without the grid wrapper
  }
  if (o.max_per_cell * g.ncells > kOrbFrameCap && grid > 1) {
    set_error("max_keypoints too large for the per-frame staging buffer (1.5 * max_keypoints <= 4096)");
    return RGBDSLAM_B200_ERR_ARG;
  }
  o.kp_stride = std::min(kOrbFrameCap, o.max_per_cell * g.ncells);
  o.W = W; o.H = H; o.grid = grid; o.max_kp = K;
  o.chunk = std::max(1, std::min(nframes_hint > 0 ? nframes_hint : 1, 64));
  int rc;
  if ((rc = o.d_ofs.ensure(o.h_ofs.size() * 2 + 16)) || (rc = o.d_w1.ensure(o.h_w1.size() * 2 + 16))) return rc;
  cudaStream_t st = s.stream;
  cudaError_t e = cudaMemcpyAsync(o.d_ofs.ptr, o.h_ofs.data(), o.h_ofs.size() * 2, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(o.d_w1.ptr, o.h_w1.data(), o.h_w1.size() * 2, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = orb_upload_constants(g, umax, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cuda_fail(e, "orb_prepare upload");
  o.tab.ofs = (const int16_t*)o.d_ofs.ptr;
  o.tab.w1 = (const uint16_t*)o.d_w1.ptr;
  o.ready = true;
  return 0;
}

static int orb_ensure_buffers(int F) {
  OrbCtx& o = g_orb;
  const OrbGeom& g = o.g;
  const size_t px = (size_t)g.W * g.H, z = (size_t)F * g.ncells;
  int rc;
  if ((rc = o.gray.ensure(px * F)) || (rc = o.mask.ensure(px * F)) || (rc = o.depth.ensure(px * F * 4)) ||
      (rc = o.cell_img.ensure((size_t)g.cell_bytes * F)) || (rc = o.cell_mask.ensure((size_t)g.cell_bytes * F)) ||
      (rc = o.score.ensure((size_t)g.cell_bytes * F)) || (rc = o.cand.ensure(z * kOrbCandCap * sizeof(OrbCand))) ||
      (rc = o.cand_count.ensure(z * 4)) || (rc = o.hist.ensure(z * 256 * 4)) || (rc = o.thr.ensure(z * 4)) ||
      (rc = o.resp.ensure(z * kOrbCandCap * 4)) || (rc = o.cell_out.ensure(z * (size_t)o.max_per_cell * 8)) ||
      (rc = o.cell_out_count.ensure(z * 4)) || (rc = o.scratch.ensure((size_t)F * 2 * kOrbFrameCap * 24)) ||
      (rc = o.kp.ensure((size_t)F * o.kp_stride * sizeof(rgbdslam_b200_keypoint))) ||
      (rc = o.xyz.ensure((size_t)F * o.kp_stride * 16)) || (rc = o.n.ensure((size_t)F * 4)) ||
      (rc = o.pyr_raw.ensure((size_t)g.full_bytes * F)) || (rc = o.pyr_blur.ensure((size_t)g.full_bytes * F)) ||
      (rc = o.desc.ensure((size_t)F * o.kp_stride * 32)))
    return rc;
  return 0;
}

// VideoDynamicAdaptedFeatureDetector::detect (feature_adjuster.cpp:185-224) on the histogram of corner scores:
// returns the FAST threshold of the LAST detection call and updates the persistent threshold.
static int adapt_threshold(double& thresh, const int* hist, bool mask_nonzero, int min_features, int max_features, int max_iters) {
  int cnt_ge[257];
  cnt_ge[256] = 0;
  for (int t = 255; t >= 0; t--) cnt_ge[t] = cnt_ge[t + 1] + hist[t];
  int iter = max_iters, used = 0;
  bool checked = false;
  do {
    int t = (int)thresh;  // static_cast<int>(thresh_) feature_adjuster.cpp:94
    used = t;
    const int found = t > 255 ? 0 : cnt_ge[t < 0 ? 0 : t];
    if (found < min_features) {
      thresh *= 0.7;  // tooFew
      if (thresh < 2.0) thresh = 2.0;
      if (found == 0 && !checked) {
        checked = true;
        if (!mask_nonzero) break;
      }
    } else if (found > max_features) {
      thresh *= 1.3;  // tooMany
      if (thresh > 10000.0) thresh = 10000.0;
      break;
    } else
      break;
    iter--;
  } while (iter > 0 && thresh > 2.0 && thresh < 10000.0);
  return used;
}

static Detector* get_detector(uint64_t h) {
  Detector* d = (Detector*)(uintptr_t)h;
  if (!d || d->magic != Detector::kMagic) {
    set_error("invalid detector handle");
    return nullptr;
  }
  return d;
}

// detection stage for F frames already resident in o.gray / o.mask (mask_present) -> thresholds decided, o.thr uploaded
static int orb_detect_stage(Detector* det, int F, const uint8_t* h_mask, int* launches) {
  const bool mask_present = h_mask != nullptr;
  State& s = g_state;
  OrbCtx& o = g_orb;
  const OrbGeom& g = o.g;
  cudaStream_t st = s.stream;
  cudaError_t e = orb_run_detect(g, o.tab, F, (const uint8_t*)o.gray.ptr, mask_present ? (const uint8_t*)o.mask.ptr : nullptr,
                                 (uint8_t*)o.cell_img.ptr, (uint8_t*)o.cell_mask.ptr, (uint8_t*)o.score.ptr, (OrbCand*)o.cand.ptr,
                                 (int*)o.cand_count.ptr, (int*)o.hist.ptr, st, launches);
  if (e != cudaSuccess) return cuda_fail(e, "orb detect kernels");
  const size_t z = (size_t)F * g.ncells;
  std::vector<int> hist(z * 256), cnt(z), thr(z);
  e = cudaMemcpyAsync(hist.data(), o.hist.ptr, z * 256 * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(cnt.data(), o.cand_count.ptr, z * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cuda_fail(e, "orb histogram download");
  for (size_t i = 0; i < z; i++)
    if (cnt[i] > kOrbCandCap) {
      set_error("ORB candidate buffer overflow (more than 12288 FAST corners in one grid cell)");
      return RGBDSLAM_B200_ERR_STATE;
    }
  // mask-all-zero test (hasNonZero, feature_adjuster.cpp:176-183): a cell with candidates has a non-zero mask; for a
  // cell without any candidate the distinction only affects how far the threshold decays, decided from the mask itself
  for (int f = 0; f < F; f++)
    for (int c = 0; c < g.ncells; c++) {
      const size_t i = (size_t)f * g.ncells + c;
      bool nz = cnt[i] > 0 || !mask_present;
      if (!nz) {  // scan the cell rectangle of the host mask
        const OrbPlane& p0 = g.cell[c][0];
        const uint8_t* m = h_mask + (size_t)f * g.W * g.H;
        for (int y = 0; y < p0.h && !nz; y++)
          for (int x = 0; x < p0.w; x++)
            if (m[(size_t)(g.cell_y0[c] + y) * g.W + g.cell_x0[c] + x]) { nz = true; break; }
      }
      thr[i] = adapt_threshold(det->thresh[c], &hist[i * 256], nz, o.min_cell, o.max_cell, s.params.adjuster_max_iterations);
    }
  e = cudaMemcpyAsync(o.thr.ptr, thr.data(), z * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cuda_fail(e, "orb threshold upload");
  return 0;
}

}  // namespace rb200

using namespace rb200;

extern "C" {

int rgbdslam_b200_detector_create(uint64_t* detector) {
  if (!detector) return RGBDSLAM_B200_ERR_ARG;
  *detector = (uint64_t)(uintptr_t) new Detector();
  return 0;
}
int rgbdslam_b200_detector_destroy(uint64_t detector) {
  Detector* d = get_detector(detector);
  if (!d) return RGBDSLAM_B200_ERR_ARG;
  d->magic = 0;
  delete d;
  return 0;
}
int rgbdslam_b200_detector_thresholds(uint64_t detector, double* thresholds16, int set) {
  Detector* d = get_detector(detector);
  if (!d || !thresholds16) return RGBDSLAM_B200_ERR_ARG;
  for (int i = 0; i < kOrbMaxCells; i++) {
    if (set) d->thresh[i] = thresholds16[i];
    else thresholds16[i] = d->thresh[i];
  }
  return 0;
}

static int upload_frames(int F, const uint8_t* gray, const float* depth, const uint8_t* mask) {
  OrbCtx& o = g_orb;
  const size_t px = (size_t)o.g.W * o.g.H;
  cudaStream_t st = g_state.stream;
  cudaError_t e = cudaMemcpyAsync(o.gray.ptr, gray, px * F, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && mask) e = cudaMemcpyAsync(o.mask.ptr, mask, px * F, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && depth) e = cudaMemcpyAsync(o.depth.ptr, depth, px * F * 4, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return cuda_fail(e, "frame upload");
  return 0;
}

int rgbdslam_b200_orb_detect(uint64_t detector, const uint8_t* gray, const uint8_t* mask, int w, int h,
                             rgbdslam_b200_keypoint* kp_out, int capacity, int* n_out) {
  std::lock_guard<std::mutex> lk(g_state.mu);
  int rc = check_inited();
  if (rc) return rc;
  Detector* det = get_detector(detector);
  if (!det || !gray || !kp_out || !n_out || capacity < 0) {
    set_error("orb_detect: bad arguments");
    return RGBDSLAM_B200_ERR_ARG;
  }
  if ((rc = orb_prepare(w, h, 1)) || (rc = orb_ensure_buffers(1)) || (rc = upload_frames(1, gray, nullptr, mask))) return rc;
  OrbCtx& o = g_orb;
  int launches = 0;
  if ((rc = orb_detect_stage(det, 1, mask, &launches))) return rc;
  cudaStream_t st = g_state.stream;
  cudaError_t e = orb_run_select(o.g, 1, 0, o.max_per_cell, g_state.params.max_keypoints, (const uint8_t*)o.cell_img.ptr,
                                 (const OrbCand*)o.cand.ptr, (const int*)o.cand_count.ptr, (const int*)o.thr.ptr, (float*)o.resp.ptr,
                                 (unsigned long long*)o.cell_out.ptr, (int*)o.cell_out_count.ptr, nullptr, 1.f,
                                 make_float4(0, 0, 0, 0), o.scratch.ptr, (rgbdslam_b200_keypoint*)o.kp.ptr, (float4*)o.xyz.ptr,
                                 (int*)o.n.ptr, o.kp_stride, st, &launches);
  if (e != cudaSuccess) return cuda_fail(e, "orb select kernels");
  int n = 0;
  e = cudaMemcpyAsync(&n, o.n.ptr, 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cuda_fail(e, "orb_detect download count");
  *n_out = n;
  const int m = n < capacity ? n : capacity;
  if (m > 0) {
    e = cudaMemcpyAsync(kp_out, o.kp.ptr, sizeof(rgbdslam_b200_keypoint) * m, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "orb_detect download keypoints");
  }
  g_state.launches += launches;
  return 0;
}

int rgbdslam_b200_orb_compute(const uint8_t* gray, int w, int h, const rgbdslam_b200_keypoint* kp_in, int n_in,
                              rgbdslam_b200_keypoint* kp_out, uint8_t* desc_out, int* n_out) {
  std::lock_guard<std::mutex> lk(g_state.mu);
  int rc = check_inited();
  if (rc) return rc;
  if (!gray || n_in < 0 || (n_in > 0 && (!kp_in || !kp_out || !desc_out)) || !n_out) {
    set_error("orb_compute: bad arguments");
    return RGBDSLAM_B200_ERR_ARG;
  }
  if ((rc = orb_prepare(w, h, 1)) || (rc = orb_ensure_buffers(1))) return rc;
  OrbCtx& o = g_orb;
  // cv::ORB::compute: runByImageBorder(31) on cvRound'ed coordinates, then group by octave (stable)
  std::vector<rgbdslam_b200_keypoint> kept;
  kept.reserve(n_in);
  for (int i = 0; i < n_in; i++) {
    const rgbdslam_b200_keypoint& k = kp_in[i];
    if (k.octave < 0 || k.octave >= kOrbLevels) {
      set_error("orb_compute: keypoint octave outside [0,8)");
      return RGBDSLAM_B200_ERR_ARG;
    }
    const long rx = lrintf(k.x), ry = lrintf(k.y);
    if (rx >= 31 && rx < w - 31 && ry >= 31 && ry < h - 31) kept.push_back(k);
  }
  std::stable_sort(kept.begin(), kept.end(),
                   [](const rgbdslam_b200_keypoint& a, const rgbdslam_b200_keypoint& b) { return a.octave < b.octave; });
  const int n = (int)kept.size();
  *n_out = n;
  if (n == 0) return 0;
  DevBuf dk, dd;  // keypoint counts are caller-sized here, not bounded by kp_stride
  if ((rc = dk.ensure(sizeof(rgbdslam_b200_keypoint) * (size_t)n)) || (rc = dd.ensure(32 * (size_t)n))) { dk.release(); dd.release(); return rc; }
  cudaStream_t st = g_state.stream;
  int launches = 0;
  cudaError_t e = cudaMemcpyAsync(o.gray.ptr, gray, (size_t)w * h, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dk.ptr, kept.data(), sizeof(rgbdslam_b200_keypoint) * (size_t)n, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(o.n.ptr, &n, 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess)
    e = orb_run_describe(o.g, o.tab, 1, (const uint8_t*)o.gray.ptr, (uint8_t*)o.pyr_raw.ptr, (uint8_t*)o.pyr_blur.ptr,
                         (const rgbdslam_b200_keypoint*)dk.ptr, (const int*)o.n.ptr, n, n, (uint8_t*)dd.ptr, st, &launches);
  if (e == cudaSuccess) e = cudaMemcpyAsync(desc_out, dd.ptr, 32 * (size_t)n, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  dk.release();
  dd.release();
  if (e != cudaSuccess) return cuda_fail(e, "orb_compute");
  memcpy(kp_out, kept.data(), sizeof(rgbdslam_b200_keypoint) * (size_t)n);
  g_state.launches += launches;
  return 0;
}

int rgbdslam_b200_nodes_create(uint64_t detector, int nframes, const uint8_t* gray, const float* depth, const uint8_t* mask,
                               int w, int h, const float* K4, const int32_t* ids, uint64_t* node_handles, int32_t* n_features) {
  std::lock_guard<std::mutex> lk(g_state.mu);
  int rc = check_inited();
  if (rc) return rc;
  Detector* det = get_detector(detector);
  if (!det || nframes < 0 || (nframes > 0 && (!gray || !depth || !K4 || !node_handles))) {
    set_error("nodes_create: bad arguments");
    return RGBDSLAM_B200_ERR_ARG;
  }
  if (nframes == 0) return 0;
  State& s = g_state;
  if ((rc = orb_prepare(w, h, nframes))) return rc;
  OrbCtx& o = g_orb;
  const int chunk = std::min(nframes, 64);
  if ((rc = orb_ensure_buffers(chunk))) return rc;
  const size_t px = (size_t)w * h;
  cudaStream_t st = s.stream;
  // projectTo3D intrinsics (node.cpp:913-916): fxinv, fyinv as float(1./fx)
  const float4 Kinv = make_float4((float)(1. / (double)K4[0]), (float)(1. / (double)K4[1]), K4[2], K4[3]);
  for (int f0 = 0; f0 < nframes; f0 += chunk) {
    const int F = std::min(chunk, nframes - f0);
    int launches = 0;
    if ((rc = upload_frames(F, gray + px * f0, depth + px * f0, mask ? mask + px * f0 : nullptr))) return rc;
    if ((rc = orb_detect_stage(det, F, mask ? mask + px * f0 : nullptr, &launches))) return rc;
    cudaError_t e = orb_run_select(o.g, F, 1, o.max_per_cell, s.params.max_keypoints, (const uint8_t*)o.cell_img.ptr,
                                   (const OrbCand*)o.cand.ptr, (const int*)o.cand_count.ptr, (const int*)o.thr.ptr,
                                   (float*)o.resp.ptr, (unsigned long long*)o.cell_out.ptr, (int*)o.cell_out_count.ptr,
                                   (const float*)o.depth.ptr, (float)s.params.depth_scaling_factor, Kinv, o.scratch.ptr,
                                   (rgbdslam_b200_keypoint*)o.kp.ptr, (float4*)o.xyz.ptr, (int*)o.n.ptr, o.kp_stride, st, &launches);
    if (e != cudaSuccess) return cuda_fail(e, "orb select kernels");
    e = orb_run_describe(o.g, o.tab, F, (const uint8_t*)o.gray.ptr, (uint8_t*)o.pyr_raw.ptr, (uint8_t*)o.pyr_blur.ptr,
                         (const rgbdslam_b200_keypoint*)o.kp.ptr, (const int*)o.n.ptr, o.kp_stride,
                         std::min(o.kp_stride, s.params.max_keypoints), (uint8_t*)o.desc.ptr, st, &launches);
    if (e != cudaSuccess) return cuda_fail(e, "orb describe kernels");
    std::vector<int> n(F);
    e = cudaMemcpyAsync(n.data(), o.n.ptr, 4 * (size_t)F, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "nodes_create count download");
    std::vector<ExpandJob> jobs;
    for (int f = 0; f < F; f++) {
      NodeDev* nd = new NodeDev();
      nd->magic = NodeDev::kMagic;
      nd->id = ids ? ids[f0 + f] : f0 + f;
      nd->n = n[f];
      nd->n_pad = ((n[f] > 0 ? n[f] : 1) + 255) / 256 * 256;
      const size_t na = (size_t)(n[f] > 0 ? n[f] : 1);
      e = cudaMalloc(&nd->desc, 32 * na);
      if (e == cudaSuccess) e = cudaMalloc(&nd->xyz, 16 * na);
      if (e == cudaSuccess) e = cudaMalloc(&nd->kp, sizeof(rgbdslam_b200_keypoint) * na);
      if (e == cudaSuccess) e = cudaMalloc(&nd->desc_i8, 256 * (size_t)nd->n_pad);
      if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(node)");
      if (n[f] > 0) {
        cudaMemcpyAsync(nd->desc, (const uint8_t*)o.desc.ptr + (size_t)f * o.kp_stride * 32, 32 * (size_t)n[f], cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(nd->xyz, (const float4*)o.xyz.ptr + (size_t)f * o.kp_stride, 16 * (size_t)n[f], cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(nd->kp, (const rgbdslam_b200_keypoint*)o.kp.ptr + (size_t)f * o.kp_stride,
                        sizeof(rgbdslam_b200_keypoint) * (size_t)n[f], cudaMemcpyDeviceToDevice, st);
      }
      if (s.params.observability_threshold > 0.0 &&
          (rc = node_build_cloud(nd, (const float*)o.depth.ptr + (size_t)f * px, w, h, K4, st)))  // Node::pc_col for the EMM
        return rc;
      jobs.push_back({nd->desc, nd->desc_i8, nd->n, nd->n_pad});
      node_handles[f0 + f] = (uint64_t)(uintptr_t)nd;
      if (n_features) n_features[f0 + f] = n[f];
    }
    if ((rc = expand_nodes_public(jobs))) return rc;
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "nodes_create finish");
    s.launches += launches;
  }
  return 0;
}

/* Debug/inspection hook (used by tools/debug_orb.py and the tests): the FAST/NMS candidates of grid cell `cell` of
 * frame 0 of the last detect / nodes_create call: 8-byte records {u16 x, u16 y, u8 level, u8 score, u16 0} and their
 * Harris responses (NaN = below the cell's final threshold). */
int rgbdslam_b200_orb_debug_candidates(int cell, void* cand_out, float* resp_out, int capacity, int* n_out, int* thr_out) {
  std::lock_guard<std::mutex> lk(g_state.mu);
  int rc = check_inited();
  if (rc) return rc;
  OrbCtx& o = g_orb;
  if (!o.ready || cell < 0 || cell >= o.g.ncells || !n_out) {
    set_error("orb_debug_candidates: no detection has run / bad cell");
    return RGBDSLAM_B200_ERR_STATE;
  }
  int n = 0, thr = 0;
  cudaStream_t st = g_state.stream;
  cudaMemcpyAsync(&n, (const int*)o.cand_count.ptr + cell, 4, cudaMemcpyDeviceToHost, st);
  cudaMemcpyAsync(&thr, (const int*)o.thr.ptr + cell, 4, cudaMemcpyDeviceToHost, st);
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cuda_fail(e, "orb_debug_candidates");
  *n_out = n;
  if (thr_out) *thr_out = thr;
  const int m = std::min(std::min(n, capacity), kOrbCandCap);
  if (m > 0 && cand_out) cudaMemcpyAsync(cand_out, (const OrbCand*)o.cand.ptr + (size_t)cell * kOrbCandCap, 8 * (size_t)m, cudaMemcpyDeviceToHost, st);
  if (m > 0 && resp_out) cudaMemcpyAsync(resp_out, (const float*)o.resp.ptr + (size_t)cell * kOrbCandCap, 4 * (size_t)m, cudaMemcpyDeviceToHost, st);
  e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cuda_fail(e, "orb_debug_candidates copy");
  return 0;
}

int rgbdslam_b200_orb_debug_plane(int which, int cell, int level, uint8_t* out, int capacity, int* w_out, int* h_out) {
  std::lock_guard<std::mutex> lk(g_state.mu);
  int rc = check_inited();
  if (rc) return rc;
  OrbCtx& o = g_orb;
  if (!o.ready || level < 0 || level >= kOrbLevels || !w_out || !h_out) return RGBDSLAM_B200_ERR_ARG;
  const OrbPlane* p;
  const uint8_t* base;
  if (which <= 2) {
    if (cell < 0 || cell >= o.g.ncells) return RGBDSLAM_B200_ERR_ARG;
    p = &o.g.cell[cell][level];
    base = (const uint8_t*)(which == 0 ? o.cell_img.ptr : which == 1 ? o.cell_mask.ptr : o.score.ptr);
  } else {
    p = &o.g.full[level];
    base = (const uint8_t*)(which == 3 ? o.pyr_raw.ptr : o.pyr_blur.ptr);
  }
  *w_out = p->w;
  *h_out = p->h;
  if (out && capacity >= p->w * p->h) {
    cudaError_t e = cudaMemcpyAsync(out, base + p->off, (size_t)p->w * p->h, cudaMemcpyDeviceToHost, g_state.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g_state.stream);
    if (e != cudaSuccess) return cuda_fail(e, "orb_debug_plane");
  }
  return 0;
}

int rgbdslam_b200_node_download_keypoints(uint64_t node_handle, rgbdslam_b200_keypoint* kp_out) {
  std::lock_guard<std::mutex> lk(g_state.mu);
  int rc = check_inited();
  if (rc) return rc;
  NodeDev* nd = (NodeDev*)(uintptr_t)node_handle;
  if (!nd || nd->magic != NodeDev::kMagic || !kp_out) {
    set_error("node_download_keypoints: bad arguments");
    return RGBDSLAM_B200_ERR_ARG;
  }
  if (!nd->kp) {
    set_error("node has no keypoints (it was created from features)");
    return RGBDSLAM_B200_ERR_STATE;
  }
  if (nd->n == 0) return 0;
  cudaError_t e = cudaMemcpyAsync(kp_out, nd->kp, sizeof(rgbdslam_b200_keypoint) * (size_t)nd->n, cudaMemcpyDeviceToHost, g_state.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(g_state.stream);
  if (e != cudaSuccess) return cuda_fail(e, "node_download_keypoints");
  return 0;
}

}  // extern "C"
