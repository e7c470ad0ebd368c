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

#include "comm.h"
#include "kernels.h"
#include "orb_host.h"
#include "state.h"

namespace rb200 {

struct Detector {
  static constexpr uint32_t kMagic = 0x44455443u;  // 'DETC'
  uint32_t magic = kMagic;
  double thresh[kOrbMaxCells];  // host mirror of the persistent per-cell thresholds
  DevBuf d_state;               // the same on the device: the recurrence runs there (k_adapt_thresholds)
  bool host_valid = true, dev_valid = false;
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
  DevBuf in_gray[2], in_mask[2], in_depth[2];  // double-buffered chunk inputs (upload of chunk k+1 under the kernels of chunk k)
  PinBuf stage[2];                             // pinned staging for callers that pass pageable memory
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  DevBuf cell_img, cell_mask, score, cand, cand_count, hist, mask_any, thr, resp, cell_out, cell_out_count, scratch, kp, xyz, n,
      pyr_raw, pyr_blur, desc, err, trig;
  // rgbdslam_b200_nodes_create_sharded: what must survive between the detection pass and the finishing pass of ALL own frames,
  // and the per-(frame, cell) tables every rank holds for ALL frames of the sequence
  DevBuf sh_gray, sh_depth, sh_mask, sh_cell_img, sh_cand, all_hist, all_cnt, all_many, all_thr;
  const uint8_t* last_gray = nullptr;  // device pointers of frame 0 of the last call (debug hooks)
  void release() {
    DevBuf* all[] = {&d_ofs, &d_w1, &in_gray[0], &in_gray[1], &in_mask[0], &in_mask[1], &in_depth[0], &in_depth[1], &cell_img,
                     &cell_mask, &score, &cand, &cand_count, &hist, &mask_any, &thr, &resp, &cell_out, &cell_out_count, &scratch,
                     &kp, &xyz, &n, &pyr_raw, &pyr_blur, &desc, &err, &trig, &sh_gray, &sh_depth, &sh_mask, &sh_cell_img, &sh_cand,
                     &all_hist, &all_cnt, &all_many, &all_thr};
    for (DevBuf* b : all) b->release();
    stage[0].release();
    stage[1].release();
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
  (void)nframes_hint;
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

constexpr int kOrbChunk = 64;  // frames per pass of nodes_create

static int orb_ensure_streams() {
  OrbCtx& o = g_orb;
  if (o.copy_stream) return 0;
  cudaError_t e = cudaStreamCreateWithFlags(&o.copy_stream, cudaStreamNonBlocking);
  for (int i = 0; i < 2 && e == cudaSuccess; i++) {
    e = cudaEventCreateWithFlags(&o.ev_ready[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&o.ev_free[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&o.ev_copied[i], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) return cuda_fail(e, "orb streams / events");
  return 0;
}

// work buffers for F frames per pass; nbuf input buffers (1: synchronous single-frame entry points, 2: nodes_create)
static int orb_ensure_buffers(int F, int nbuf, bool want_mask) {
  OrbCtx& o = g_orb;
  const OrbGeom& g = o.g;
  const size_t px = (size_t)g.W * g.H, z = (size_t)F * g.ncells;
  int rc;
  for (int b = 0; b < nbuf; b++)
    if ((rc = o.in_gray[b].ensure(px * F)) || (want_mask && (rc = o.in_mask[b].ensure(px * F))) || (rc = o.in_depth[b].ensure(px * F * 4)))
      return rc;
  if ((rc = o.cell_img.ensure((size_t)g.cell_bytes * F)) || (rc = o.cell_mask.ensure((size_t)g.cell_bytes * F)) ||
      (rc = o.score.ensure((size_t)g.cell_bytes * F)) || (rc = o.cand.ensure(z * kOrbCandCap * sizeof(OrbCand))) ||
      (rc = o.cand_count.ensure(z * 4)) || (rc = o.hist.ensure(z * 256 * 4)) || (rc = o.mask_any.ensure(z * 4)) ||
      (rc = o.thr.ensure(z * 4)) || (rc = o.resp.ensure(z * kOrbCandCap * 4)) ||
      (rc = o.cell_out.ensure(z * (size_t)o.max_per_cell * 8)) || (rc = o.cell_out_count.ensure(z * 4)) ||
      (rc = o.scratch.ensure((size_t)F * 2 * kOrbFrameCap * 24)) ||
      (rc = o.kp.ensure((size_t)F * o.kp_stride * sizeof(rgbdslam_b200_keypoint))) ||
      (rc = o.xyz.ensure((size_t)F * o.kp_stride * 16)) || (rc = o.n.ensure((size_t)F * 4)) ||
      (rc = o.pyr_raw.ensure((size_t)g.full_bytes * F)) || (rc = o.pyr_blur.ensure((size_t)g.full_bytes * F)) ||
      (rc = o.desc.ensure((size_t)F * o.kp_stride * 32)) || (rc = o.err.ensure(16)) ||
      (rc = o.trig.ensure((size_t)F * o.kp_stride * 8)))
    return rc;
  return 0;
}

static Detector* get_detector(uint64_t h) {
  Detector* d = (Detector*)(uintptr_t)h;
  if (!d || d->magic != Detector::kMagic) {
    set_error("invalid detector handle");
    return nullptr;
  }
  return d;
}

// the detector's thresholds live on the device while frames are being processed; the host mirror is refreshed on demand
static int detector_to_device(Detector* det, cudaStream_t st) {
  int rc;
  if ((rc = det->d_state.ensure(sizeof(double) * kOrbMaxCells))) return rc;
  if (!det->dev_valid) {
    cudaError_t e = cudaMemcpyAsync(det->d_state.ptr, det->thresh, sizeof(double) * kOrbMaxCells, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "detector state upload");
    det->dev_valid = true;
  }
  return 0;
}
static int detector_to_host(Detector* det) {
  if (det->host_valid) return 0;
  cudaError_t e = cudaMemcpy(det->thresh, det->d_state.ptr, sizeof(double) * kOrbMaxCells, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return cuda_fail(e, "detector state download");
  det->host_valid = true;
  return 0;
}

// detection stage for F frames resident at d_gray / d_mask (or the mask derived from d_depth_for_mask): candidates,
// histograms, and the adaptive-threshold recurrence of the F frames in order -- all queued on `st`, no host round trip
static int orb_detect_stage(Detector* det, int F, const uint8_t* d_gray, const uint8_t* d_mask, const float* d_depth_for_mask,
                            cudaStream_t st, int* launches) {
  State& s = g_state;
  OrbCtx& o = g_orb;
  const OrbGeom& g = o.g;
  int rc;
  if ((rc = detector_to_device(det, st))) return rc;
  cudaError_t e = orb_run_detect(g, o.tab, F, d_gray, d_mask, d_depth_for_mask, (uint8_t*)o.cell_img.ptr, (uint8_t*)o.cell_mask.ptr,
                                 (uint8_t*)o.score.ptr, (OrbCand*)o.cand.ptr, (int*)o.cand_count.ptr, (int*)o.hist.ptr,
                                 (int*)o.mask_any.ptr, st, launches);
  if (e != cudaSuccess) return cuda_fail(e, "orb detect kernels");
  e = orb_run_adapt(g, F, (const int*)o.hist.ptr, (const int*)o.cand_count.ptr, (const int*)o.mask_any.ptr, (double*)det->d_state.ptr,
                    (int*)o.thr.ptr, o.min_cell, o.max_cell, s.params.adjuster_max_iterations, (int*)o.err.ptr, st, launches);
  if (e != cudaSuccess) return cuda_fail(e, "orb threshold kernel");
  det->host_valid = false;
  return 0;
}

static int orb_check_err_flag(int flag) {
  if (flag & 1) {
    set_error("ORB candidate buffer overflow (more than 12288 FAST corners in one grid cell)");
    return RGBDSLAM_B200_ERR_STATE;
  }
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
  std::lock_guard<std::mutex> lk(g_state.mu);
  Detector* d = get_detector(detector);
  if (!d) return RGBDSLAM_B200_ERR_ARG;
  if (g_state.inited) {
    cudaSetDevice(g_state.device);
    cudaStreamSynchronize(g_state.stream);
  }
  d->d_state.release();
  d->magic = 0;
  delete d;
  return 0;
}
int rgbdslam_b200_detector_thresholds(uint64_t detector, double* thresholds16, int set) {
  std::lock_guard<std::mutex> lk(g_state.mu);
  Detector* d = get_detector(detector);
  if (!d || !thresholds16) return RGBDSLAM_B200_ERR_ARG;
  if (!d->host_valid) {
    int rc = check_inited();
    if (rc) return rc;
    cudaError_t e = cudaStreamSynchronize(g_state.stream);
    if (e != cudaSuccess) return cuda_fail(e, "detector_thresholds");
    if ((rc = detector_to_host(d))) return rc;
  }
  for (int i = 0; i < kOrbMaxCells; i++) {
    if (set) d->thresh[i] = thresholds16[i];
    else thresholds16[i] = d->thresh[i];
  }
  if (set) d->dev_valid = false;
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
  if ((rc = orb_prepare(w, h, 1)) || (rc = orb_ensure_buffers(1, 1, true))) return rc;
  OrbCtx& o = g_orb;
  cudaStream_t st = g_state.stream;
  const size_t px = (size_t)w * h;
  cudaError_t e = cudaMemsetAsync(o.err.ptr, 0, 4, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(o.in_gray[0].ptr, gray, px, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && mask) e = cudaMemcpyAsync(o.in_mask[0].ptr, mask, px, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return cuda_fail(e, "orb_detect upload");
  o.last_gray = (const uint8_t*)o.in_gray[0].ptr;
  int launches = 0;
  if ((rc = orb_detect_stage(det, 1, (const uint8_t*)o.in_gray[0].ptr, mask ? (const uint8_t*)o.in_mask[0].ptr : nullptr, nullptr, st,
                             &launches)))
    return rc;
  e = orb_run_select(o.g, 1, 0, o.max_per_cell, g_state.params.max_keypoints, (const uint8_t*)o.cell_img.ptr,
                     (const OrbCand*)o.cand.ptr, (const int*)o.cand_count.ptr, (const int*)o.thr.ptr, (float*)o.resp.ptr,
                     (unsigned long long*)o.cell_out.ptr, (int*)o.cell_out_count.ptr, nullptr, 1.f, make_float4(0, 0, 0, 0),
                     o.scratch.ptr, (rgbdslam_b200_keypoint*)o.kp.ptr, (float4*)o.xyz.ptr, nullptr, (int*)o.n.ptr, o.kp_stride, st,
                     &launches);
  if (e != cudaSuccess) return cuda_fail(e, "orb select kernels");
  int n = 0, flag = 0;
  e = cudaMemcpyAsync(&n, o.n.ptr, 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, o.err.ptr, 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cuda_fail(e, "orb_detect download count");
  if ((rc = orb_check_err_flag(flag))) return rc;
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
  if ((rc = orb_prepare(w, h, 1)) || (rc = orb_ensure_buffers(1, 1, true))) return rc;
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
  cudaError_t e = cudaMemcpyAsync(o.in_gray[0].ptr, gray, (size_t)w * h, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dk.ptr, kept.data(), sizeof(rgbdslam_b200_keypoint) * (size_t)n, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(o.n.ptr, &n, 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess)
    e = orb_run_describe(o.g, o.tab, 1, (const uint8_t*)o.in_gray[0].ptr, (uint8_t*)o.pyr_raw.ptr, (uint8_t*)o.pyr_blur.ptr,
                         (const rgbdslam_b200_keypoint*)dk.ptr, (const int*)o.n.ptr, n, n, nullptr, (uint8_t*)dd.ptr, st, &launches);
  if (e == cudaSuccess) e = cudaMemcpyAsync(desc_out, dd.ptr, 32 * (size_t)n, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  dk.release();
  dd.release();
  if (e != cudaSuccess) return cuda_fail(e, "orb_compute");
  memcpy(kp_out, kept.data(), sizeof(rgbdslam_b200_keypoint) * (size_t)n);
  g_state.launches += launches;
  return 0;
}

static bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// Node::Node for nframes frames in order.  Pipeline per chunk of kOrbChunk frames:
//   copy stream    : host -> device of the chunk's gray / depth / mask into one of two input buffers (straight from the caller's
//                    buffers when they are pinned, else through pinned staging filled by this thread)
//   compute stream : detect kernels -> threshold recurrence (device) -> Harris / keepStrongest / finalize -> describe, writing
//                    straight into the slab that holds all nodes of the call (no per-node allocation, no device->device copy)
// The only host synchronisation is the download of the feature counts at the end.
int rgbdslam_b200_nodes_create_ex(uint64_t detector, int nframes, const uint8_t* gray, const float* depth, const uint8_t* mask,
                                  int w, int h, const float* K4, const int32_t* ids, int flags, uint64_t* node_handles,
                                  int32_t* n_features) {
  std::lock_guard<std::mutex> lk(g_state.mu);
  int rc = check_inited();
  if (rc) return rc;
  Detector* det = get_detector(detector);
  if (!det || nframes < 0 || (nframes > 0 && (!gray || !depth || !K4 || !node_handles)) ||
      (flags & ~RGBDSLAM_B200_MASK_FROM_DEPTH)) {
    set_error("nodes_create: bad arguments");
    return RGBDSLAM_B200_ERR_ARG;
  }
  if (nframes == 0) return 0;
  State& s = g_state;
  const bool mask_from_depth = (flags & RGBDSLAM_B200_MASK_FROM_DEPTH) != 0;
  if (mask_from_depth) mask = nullptr;
  if ((rc = orb_prepare(w, h, nframes)) || (rc = orb_ensure_streams())) return rc;
  OrbCtx& o = g_orb;
  const int chunk = std::min(nframes, kOrbChunk);
  if ((rc = orb_ensure_buffers(chunk, 2, mask != nullptr))) return rc;
  const size_t px = (size_t)w * h;
  cudaStream_t st = s.stream, cs = o.copy_stream;
  const int K = std::min(o.kp_stride, s.params.max_keypoints);  // features per node (finalize mode 1 emits <= max_keypoints)
  const int Kpad = ((K > 0 ? K : 1) + 255) / 256 * 256;
  // slab: [desc F x K x 32][xyz F x K x 16][kp F x K x 28][n F x 4]
  const size_t b_desc = ((size_t)nframes * K * 32 + 255) / 256 * 256, b_xyz = ((size_t)nframes * K * 16 + 255) / 256 * 256;
  const size_t b_kp = ((size_t)nframes * K * sizeof(rgbdslam_b200_keypoint) + 255) / 256 * 256;
  const size_t b_n = ((size_t)nframes * 4 + 255) / 256 * 256;
  NodeSlab* slab = new NodeSlab();
  cudaError_t e = cudaMalloc(&slab->base, b_desc + b_xyz + b_kp + b_n);
  if (e != cudaSuccess) {
    delete slab;
    return cuda_fail(e, "cudaMalloc(node slab)");
  }
  uint8_t* sl_desc = (uint8_t*)slab->base;
  float4* sl_xyz = (float4*)(sl_desc + b_desc);
  rgbdslam_b200_keypoint* sl_kp = (rgbdslam_b200_keypoint*)((uint8_t*)sl_xyz + b_xyz);
  int* sl_n = (int*)((uint8_t*)sl_kp + b_kp);
  auto fail = [&](int code) {
    cudaStreamSynchronize(cs);
    cudaStreamSynchronize(st);
    cudaFree(slab->base);
    delete slab;
    return code;
  };
  const bool pinned = is_pinned(gray) && is_pinned(depth) && (!mask || is_pinned(mask));
  const size_t stage_bytes = (px + px * 4 + (mask ? px : 0)) * chunk;
  if (!pinned && ((rc = o.stage[0].ensure(stage_bytes)) || (rc = o.stage[1].ensure(stage_bytes)))) return fail(rc);
  // projectTo3D intrinsics (node.cpp:913-916): fxinv, fyinv as float(1./fx)
  const float4 Kinv = make_float4((float)(1. / (double)K4[0]), (float)(1. / (double)K4[1]), K4[2], K4[3]);
  e = cudaMemsetAsync(o.err.ptr, 0, 4, st);
  if (e != cudaSuccess) return fail(cuda_fail(e, "nodes_create"));
  // the copy stream must not run ahead of work already queued on the compute stream that still reads the input buffers
  e = cudaEventRecord(o.ev_free[0], st);
  if (e == cudaSuccess) e = cudaEventRecord(o.ev_free[1], st);
  if (e != cudaSuccess) return fail(cuda_fail(e, "nodes_create events"));
  int launches = 0, ci = 0;
  std::vector<NodeDev*> made;
  for (int f0 = 0; f0 < nframes; f0 += chunk, ci++) {
    const int F = std::min(chunk, nframes - f0), b = ci & 1;
    const uint8_t* hg = gray + px * f0;
    const float* hd = depth + px * f0;
    const uint8_t* hm = mask ? mask + px * f0 : nullptr;
    if (!pinned) {  // stage through pinned memory (the previous copy out of this staging buffer must have finished)
      if (ci >= 2 && (e = cudaEventSynchronize(o.ev_copied[b])) != cudaSuccess) return fail(cuda_fail(e, "staging wait"));
      uint8_t* sp = (uint8_t*)o.stage[b].ptr;
      memcpy(sp, hg, px * F);
      memcpy(sp + px * chunk, hd, px * 4 * F);
      if (hm) memcpy(sp + px * 5 * chunk, hm, px * F);
      hg = sp;
      hd = (const float*)(sp + px * chunk);
      if (hm) hm = sp + px * 5 * chunk;
    }
    e = cudaStreamWaitEvent(cs, o.ev_free[b], 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(o.in_gray[b].ptr, hg, px * F, cudaMemcpyHostToDevice, cs);
    if (e == cudaSuccess) e = cudaMemcpyAsync(o.in_depth[b].ptr, hd, px * 4 * F, cudaMemcpyHostToDevice, cs);
    if (e == cudaSuccess && hm) e = cudaMemcpyAsync(o.in_mask[b].ptr, hm, px * F, cudaMemcpyHostToDevice, cs);
    if (e == cudaSuccess) e = cudaEventRecord(o.ev_ready[b], cs);
    if (e == cudaSuccess) e = cudaEventRecord(o.ev_copied[b], cs);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, o.ev_ready[b], 0);
    if (e != cudaSuccess) return fail(cuda_fail(e, "frame upload"));
    const uint8_t* dg = (const uint8_t*)o.in_gray[b].ptr;
    const float* dd = (const float*)o.in_depth[b].ptr;
    if (f0 == 0) o.last_gray = dg;
    if ((rc = orb_detect_stage(det, F, dg, hm ? (const uint8_t*)o.in_mask[b].ptr : nullptr, mask_from_depth ? dd : nullptr, st,
                               &launches)))
      return fail(rc);
    e = orb_run_select(o.g, F, 1, o.max_per_cell, s.params.max_keypoints, (const uint8_t*)o.cell_img.ptr,
                       (const OrbCand*)o.cand.ptr, (const int*)o.cand_count.ptr, (const int*)o.thr.ptr, (float*)o.resp.ptr,
                       (unsigned long long*)o.cell_out.ptr, (int*)o.cell_out_count.ptr, dd, (float)s.params.depth_scaling_factor, Kinv,
                       o.scratch.ptr, sl_kp + (size_t)f0 * K, sl_xyz + (size_t)f0 * K, (float2*)o.trig.ptr, sl_n + f0, K, st, &launches);
    if (e != cudaSuccess) return fail(cuda_fail(e, "orb select kernels"));
    e = orb_run_describe(o.g, o.tab, F, dg, (uint8_t*)o.pyr_raw.ptr, (uint8_t*)o.pyr_blur.ptr, sl_kp + (size_t)f0 * K, sl_n + f0, K, K,
                         (const float2*)o.trig.ptr, sl_desc + (size_t)f0 * K * 32, st, &launches);
    if (e != cudaSuccess) return fail(cuda_fail(e, "orb describe kernels"));
    if (s.params.observability_threshold > 0.0) {  // Node::pc_col for the environment measurement model
      for (int f = 0; f < F; f++) {
        NodeDev* nd = new NodeDev();
        made.push_back(nd);
        if ((rc = node_build_cloud(nd, dd + (size_t)f * px, w, h, K4, st))) {
          for (NodeDev* x : made) { if (x->cloud_z) cudaFree(x->cloud_z); delete x; }
          return fail(rc);
        }
      }
    }
    e = cudaEventRecord(o.ev_free[b], st);
    if (e != cudaSuccess) return fail(cuda_fail(e, "nodes_create events"));
  }
  std::vector<int> n(nframes);
  int flag = 0;
  e = cudaMemcpyAsync(n.data(), sl_n, 4 * (size_t)nframes, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, o.err.ptr, 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(cs);
  if (e != cudaSuccess || (rc = orb_check_err_flag(flag))) {
    for (NodeDev* x : made) { if (x->cloud_z) cudaFree(x->cloud_z); delete x; }
    return fail(e != cudaSuccess ? cuda_fail(e, "nodes_create finish") : rc);
  }
  for (int f = 0; f < nframes; f++) {
    NodeDev* nd = made.empty() ? new NodeDev() : made[f];
    nd->magic = NodeDev::kMagic;
    nd->id = ids ? ids[f] : f;
    nd->n = n[f];
    nd->n_pad = Kpad;
    nd->desc = sl_desc + (size_t)f * K * 32;
    nd->xyz = sl_xyz + (size_t)f * K;
    nd->kp = sl_kp + (size_t)f * K;
    nd->slab = slab;
    slab->refs++;
    node_handles[f] = (uint64_t)(uintptr_t)nd;
    if (n_features) n_features[f] = n[f];
  }
  s.launches += launches;
  return 0;
}

// Frame-sharded Node construction: see include/rgbdslam_b200.h.  Two passes over the rank's own frames around ONE exchange of
// the score histograms; then one exchange of the finished features.
int rgbdslam_b200_nodes_create_sharded(uint64_t detector, uint64_t comm_handle, int total_frames, const uint8_t* gray,
                                       const float* depth, const uint8_t* mask, int w, int h, const float* K4, const int32_t* ids,
                                       int flags, uint64_t* node_handles, int32_t* n_features) {
  std::lock_guard<std::mutex> lk(g_state.mu);
  int rc = check_inited();
  if (rc) return rc;
  Detector* det = get_detector(detector);
  Comm* cm = get_comm(comm_handle);
  if (!det || !cm || total_frames < 0 || (total_frames > 0 && (!K4 || !node_handles)) || (flags & ~RGBDSLAM_B200_MASK_FROM_DEPTH)) {
    set_error("nodes_create_sharded: bad arguments");
    return RGBDSLAM_B200_ERR_ARG;
  }
  if (total_frames == 0) return 0;
  State& s = g_state;
  if (s.params.observability_threshold > 0.0) {
    set_error("nodes_create_sharded: the environment measurement model needs every node's depth cloud on every rank (not exchanged)");
    return RGBDSLAM_B200_ERR_STATE;
  }
  const int world = cm->world, rank = cm->rank;
  const int per = (total_frames + world - 1) / world;
  const int f0 = std::min(rank * per, total_frames), f1 = std::min((rank + 1) * per, total_frames);
  const int own = f1 - f0, Wp = world * per;
  if (own > 0 && (!gray || !depth)) {
    set_error("nodes_create_sharded: null image buffers");
    return RGBDSLAM_B200_ERR_ARG;
  }
  const bool mask_from_depth = (flags & RGBDSLAM_B200_MASK_FROM_DEPTH) != 0;
  if (mask_from_depth) mask = nullptr;
  if ((rc = orb_prepare(w, h, own)) || (rc = orb_ensure_streams())) return rc;
  OrbCtx& o = g_orb;
  const OrbGeom& g = o.g;
  const int chunk = std::max(1, std::min(own, kOrbChunk));
  if ((rc = orb_ensure_buffers(chunk, 0, false))) return rc;
  const size_t px = (size_t)w * h, nc = (size_t)g.ncells;
  const size_t own_ = (size_t)std::max(own, 1);
  if ((rc = o.sh_gray.ensure(px * own_)) || (rc = o.sh_depth.ensure(px * 4 * own_)) || (mask && (rc = o.sh_mask.ensure(px * own_))) ||
      (rc = o.sh_cell_img.ensure((size_t)g.cell_bytes * own_)) || (rc = o.sh_cand.ensure(own_ * nc * kOrbCandCap * sizeof(OrbCand))) ||
      (rc = o.all_hist.ensure((size_t)Wp * nc * 256 * 4)) || (rc = o.all_cnt.ensure((size_t)Wp * nc * 4)) ||
      (rc = o.all_many.ensure((size_t)Wp * nc * 4)) || (rc = o.all_thr.ensure((size_t)Wp * nc * 4)))
    return rc;
  cudaStream_t st = s.stream, cs = o.copy_stream;
  const int K = std::min(o.kp_stride, s.params.max_keypoints);
  const int Kpad = ((K > 0 ? K : 1) + 255) / 256 * 256;
  // slab: [desc Wp x K x 32][xyz Wp x K x 16][n Wp x 4][kp own x K x 28]
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t b_desc = up((size_t)Wp * K * 32), b_xyz = up((size_t)Wp * K * 16), b_n = up((size_t)Wp * 4);
  const size_t b_kp = up(own_ * K * sizeof(rgbdslam_b200_keypoint));
  NodeSlab* slab = new NodeSlab();
  cudaError_t e = cudaMalloc(&slab->base, b_desc + b_xyz + b_n + b_kp);
  if (e != cudaSuccess) {
    delete slab;
    return cuda_fail(e, "cudaMalloc(node slab)");
  }
  uint8_t* sl_desc = (uint8_t*)slab->base;
  float4* sl_xyz = (float4*)(sl_desc + b_desc);
  int* sl_n = (int*)((uint8_t*)sl_xyz + b_xyz);
  rgbdslam_b200_keypoint* sl_kp = (rgbdslam_b200_keypoint*)((uint8_t*)sl_n + b_n);
  auto fail = [&](int code) {
    cudaStreamSynchronize(cs);
    cudaStreamSynchronize(st);
    cudaFree(slab->base);
    delete slab;
    return code;
  };
  const bool pinned = own == 0 || (is_pinned(gray) && is_pinned(depth) && (!mask || is_pinned(mask)));
  const size_t stage_bytes = (px + px * 4 + (mask ? px : 0)) * chunk;
  if (!pinned && ((rc = o.stage[0].ensure(stage_bytes)) || (rc = o.stage[1].ensure(stage_bytes)))) return fail(rc);
  const float4 Kinv = make_float4((float)(1. / (double)K4[0]), (float)(1. / (double)K4[1]), K4[2], K4[3]);
  int* hist_all = (int*)o.all_hist.ptr;
  int* cnt_all = (int*)o.all_cnt.ptr;
  int* many_all = (int*)o.all_many.ptr;
  int* thr_all = (int*)o.all_thr.ptr;
  e = cudaMemsetAsync(o.err.ptr, 0, 4, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(sl_n, 0, b_n, st);
  // frames of the padding (Wp > total_frames) and of ranks without frames must read as "no candidates"
  if (e == cudaSuccess) e = cudaMemsetAsync(hist_all, 0, (size_t)Wp * nc * 256 * 4, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(cnt_all, 0, (size_t)Wp * nc * 4, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(many_all, 0, (size_t)Wp * nc * 4, st);
  if (e == cudaSuccess) e = cudaEventRecord(o.ev_free[0], st);  // uploads start after everything queued so far
  if (e == cudaSuccess) e = cudaStreamWaitEvent(cs, o.ev_free[0], 0);
  if (e != cudaSuccess) return fail(cuda_fail(e, "nodes_create_sharded setup"));
  int launches = 0, ci = 0;
  // ---- pass A: upload + pyramids + FAST / NMS candidates + score histograms of the own frames
  for (int c0 = 0; c0 < own; c0 += chunk, ci++) {
    const int F = std::min(chunk, own - c0), b = ci & 1;
    const uint8_t* hg = gray + px * c0;
    const float* hd = depth + px * c0;
    const uint8_t* hm = mask ? mask + px * c0 : nullptr;
    if (!pinned) {
      if (ci >= 2 && (e = cudaEventSynchronize(o.ev_copied[b])) != cudaSuccess) return fail(cuda_fail(e, "staging wait"));
      uint8_t* sp = (uint8_t*)o.stage[b].ptr;
      memcpy(sp, hg, px * F);
      memcpy(sp + px * chunk, hd, px * 4 * F);
      if (hm) memcpy(sp + px * 5 * chunk, hm, px * F);
      hg = sp;
      hd = (const float*)(sp + px * chunk);
      if (hm) hm = sp + px * 5 * chunk;
    }
    uint8_t* dg = (uint8_t*)o.sh_gray.ptr + px * c0;
    float* dd = (float*)o.sh_depth.ptr + px * c0;
    uint8_t* dm = hm ? (uint8_t*)o.sh_mask.ptr + px * c0 : nullptr;
    e = cudaMemcpyAsync(dg, hg, px * F, cudaMemcpyHostToDevice, cs);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dd, hd, px * 4 * F, cudaMemcpyHostToDevice, cs);
    if (e == cudaSuccess && hm) e = cudaMemcpyAsync(dm, hm, px * F, cudaMemcpyHostToDevice, cs);
    if (e == cudaSuccess) e = cudaEventRecord(o.ev_ready[b], cs);
    if (e == cudaSuccess) e = cudaEventRecord(o.ev_copied[b], cs);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, o.ev_ready[b], 0);
    if (e != cudaSuccess) return fail(cuda_fail(e, "frame upload"));
    if (c0 == 0) o.last_gray = dg;
    const size_t gf = (size_t)(f0 + c0);  // global index of the chunk's first frame
    e = orb_run_detect(g, o.tab, F, dg, dm, mask_from_depth ? dd : nullptr, (uint8_t*)o.sh_cell_img.ptr + (size_t)g.cell_bytes * c0,
                       (uint8_t*)o.cell_mask.ptr, (uint8_t*)o.score.ptr, (OrbCand*)o.sh_cand.ptr + (size_t)c0 * nc * kOrbCandCap,
                       cnt_all + gf * nc, hist_all + gf * nc * 256, many_all + gf * nc, st, &launches);
    if (e != cudaSuccess) return fail(cuda_fail(e, "orb detect kernels"));
  }
  // ---- the exchange that makes the frames independent: every rank gets every frame's score histograms, replays the
  //      threshold recurrence of the whole sequence (feature_adjuster.cpp:131-150, 185-224) and keeps its own frames' thresholds
  if (world > 1) {
    ncclResult_t r = g_nccl.AllGather(hist_all + (size_t)rank * per * nc * 256, hist_all, (size_t)per * nc * 256 * 4, 0, cm->comm, st);
    if (r == 0) r = g_nccl.AllGather(cnt_all + (size_t)rank * per * nc, cnt_all, (size_t)per * nc * 4, 0, cm->comm, st);
    if (r == 0) r = g_nccl.AllGather(many_all + (size_t)rank * per * nc, many_all, (size_t)per * nc * 4, 0, cm->comm, st);
    if (r != 0) return fail(nccl_fail(r, "ncclAllGather(histograms)"));
  }
  if ((rc = detector_to_device(det, st))) return fail(rc);
  e = orb_run_adapt(g, total_frames, hist_all, cnt_all, many_all, (double*)det->d_state.ptr, thr_all, o.min_cell, o.max_cell,
                    s.params.adjuster_max_iterations, (int*)o.err.ptr, st, &launches);
  if (e != cudaSuccess) return fail(cuda_fail(e, "orb threshold kernel"));
  det->host_valid = false;
  // ---- pass B: Harris / keepStrongest / finalize / describe of the own frames, straight into the slab
  for (int c0 = 0; c0 < own; c0 += chunk) {
    const int F = std::min(chunk, own - c0);
    const size_t gf = (size_t)(f0 + c0);
    const uint8_t* dg = (const uint8_t*)o.sh_gray.ptr + px * c0;
    const float* dd = (const float*)o.sh_depth.ptr + px * c0;
    const uint8_t* cimg = (const uint8_t*)o.sh_cell_img.ptr + (size_t)g.cell_bytes * c0;
    e = orb_run_select(g, F, 1, o.max_per_cell, s.params.max_keypoints, cimg, (const OrbCand*)o.sh_cand.ptr + (size_t)c0 * nc * kOrbCandCap,
                       cnt_all + gf * nc, thr_all + gf * nc, (float*)o.resp.ptr, (unsigned long long*)o.cell_out.ptr,
                       (int*)o.cell_out_count.ptr, dd, (float)s.params.depth_scaling_factor, Kinv, o.scratch.ptr, sl_kp + (size_t)c0 * K,
                       sl_xyz + gf * K, (float2*)o.trig.ptr, sl_n + gf, K, st, &launches);
    if (e != cudaSuccess) return fail(cuda_fail(e, "orb select kernels"));
    e = orb_run_describe(g, o.tab, F, dg, (uint8_t*)o.pyr_raw.ptr, (uint8_t*)o.pyr_blur.ptr, sl_kp + (size_t)c0 * K, sl_n + gf, K, K,
                         (const float2*)o.trig.ptr, sl_desc + gf * K * 32, st, &launches);
    if (e != cudaSuccess) return fail(cuda_fail(e, "orb describe kernels"));
  }
  // ---- every rank gets every node's features (48 KB per 1000-keypoint frame over NVLink)
  if (world > 1) {
    ncclResult_t r = g_nccl.AllGather(sl_desc + (size_t)rank * per * K * 32, sl_desc, (size_t)per * K * 32, 0, cm->comm, st);
    if (r == 0) r = g_nccl.AllGather((uint8_t*)(sl_xyz + (size_t)rank * per * K), sl_xyz, (size_t)per * K * 16, 0, cm->comm, st);
    if (r == 0) r = g_nccl.AllGather((uint8_t*)(sl_n + (size_t)rank * per), sl_n, (size_t)per * 4, 0, cm->comm, st);
    if (r != 0) return fail(nccl_fail(r, "ncclAllGather(features)"));
  }
  std::vector<int> n(total_frames);
  int flag = 0;
  e = cudaMemcpyAsync(n.data(), sl_n, 4 * (size_t)total_frames, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, o.err.ptr, 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(cs);
  if (e != cudaSuccess) return fail(cuda_fail(e, "nodes_create_sharded finish"));
  if ((rc = orb_check_err_flag(flag))) return fail(rc);
  for (int f = 0; f < total_frames; f++) {
    NodeDev* nd = new NodeDev();
    nd->magic = NodeDev::kMagic;
    nd->id = ids ? ids[f] : f;
    nd->n = n[f];
    nd->n_pad = Kpad;
    nd->desc = sl_desc + (size_t)f * K * 32;
    nd->xyz = sl_xyz + (size_t)f * K;
    nd->kp = (f >= f0 && f < f1) ? sl_kp + (size_t)(f - f0) * K : nullptr;  // 2-D keypoints stay on the rank that built the node
    nd->slab = slab;
    slab->refs++;
    node_handles[f] = (uint64_t)(uintptr_t)nd;
    if (n_features) n_features[f] = n[f];
  }
  s.launches += launches;
  return 0;
}

int rgbdslam_b200_nodes_create(uint64_t detector, int nframes, const uint8_t* gray, const float* depth, const uint8_t* mask,
                               int w, int h, const float* K4, const int32_t* ids, uint64_t* node_handles, int32_t* n_features) {
  return rgbdslam_b200_nodes_create_ex(detector, nframes, gray, depth, mask, w, h, K4, ids, 0, node_handles, n_features);
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

int rgbdslam_b200_orb_debug_detect_path(int unfused) {
  std::lock_guard<std::mutex> lk(g_state.mu);
  orb_set_legacy_detect(unfused);
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
