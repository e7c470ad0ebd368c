// orb.cu -- ORB keypoint detection / description kernels for sm_100a (see orb.cuh for provenance).
//
// Node constructor path (src/node.cpp:101-240), batched over frames:
//   detector->detect(gray, kp, mask)      VideoGridAdaptedFeatureDetector (feature_adjuster.cpp:286-317) over
//                                         VideoDynamicAdaptedFeatureDetector (:185-224) over cv::ORB (:94)
//     k_cell_extract / k_resize           per-cell 8-level pyramids, chained INTER_LINEAR_EXACT; mask pyramid
//     k_fast_score / k_nms_collect        threshold-free FAST-9/16 corner score S, strict 3x3 NMS, mask + 15 px border
//                                         -> candidate list + histogram of S per (frame, cell)
//     (host) adaptive thresholds          the x0.7 / x1.3 recurrence only needs #candidates(S >= t): a histogram lookup
//     k_harris                            Harris response of the candidates that pass the final threshold
//     k_cell_select                       keepStrongest(maxTotal / cells) by |response|   (feature_adjuster.cpp:247-255)
//   removeDepthless / retainBest / compute() border filter + octave sort / projectTo3D      (node.cpp:186-210)
//     k_frame_finalize                    one CTA per frame
//   extractor->compute(gray, kp, desc)    cv::ORB::create() defaults (features.cpp:117-119)
//     k_resize / k_blur / k_describe      full-image pyramid, 7-tap float Gaussian, steered BRIEF (one warp / keypoint)
#include "orb.cuh"

#include <cuda_runtime.h>

#include "orb_tables_generated.h"
#include "orb_host.h"

namespace rb200 {

__constant__ OrbGeom c_geom;
__constant__ float c_gauss[7];
__constant__ int8_t c_pattern[256][4];
__constant__ int c_umax[kOrbHalfPatch + 2];

cudaError_t orb_upload_constants(const OrbGeom& g, const int* umax, cudaStream_t st) {
  cudaError_t e = cudaMemcpyToSymbolAsync(c_geom, &g, sizeof(OrbGeom), 0, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbolAsync(c_gauss, kOrbGaussBits, sizeof(float) * 7, 0, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbolAsync(c_pattern, kOrbPattern, sizeof(kOrbPattern), 0, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  return cudaMemcpyToSymbolAsync(c_umax, umax, sizeof(int) * (kOrbHalfPatch + 2), 0, cudaMemcpyHostToDevice, st);
}

// -------------------------------------------------------------------------------------------------
// level 0 of the per-cell pyramids: sub-image copy; mask binarised (cv2: any non-zero mask pixel is valid)
__global__ void __launch_bounds__(256) k_cell_extract(const uint8_t* __restrict__ gray, const uint8_t* __restrict__ mask,
                                                      uint8_t* __restrict__ cell_img, uint8_t* __restrict__ cell_mask) {
  const int f = blockIdx.z / c_geom.ncells, c = blockIdx.z % c_geom.ncells;
  const OrbPlane& p = c_geom.cell[c][0];
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= p.w || y >= p.h) return;
  const size_t src = (size_t)f * c_geom.W * c_geom.H + (size_t)(c_geom.cell_y0[c] + y) * c_geom.W + c_geom.cell_x0[c] + x;
  const size_t dst = (size_t)f * c_geom.cell_bytes + p.off + (size_t)y * p.w + x;
  cell_img[dst] = gray[src];
  cell_mask[dst] = mask ? (mask[src] ? 255 : 0) : 255;
}

// dst plane = INTER_LINEAR_EXACT resize of the previous level (8.8 fixed-point taps, round to nearest at the end).
// which: 0 = cell pyramids (plane index blockIdx.z % ncells), 1 = full-image pyramid.  tozero: THRESH_TOZERO(254)
// applied to the result (mask pyramid).
__global__ void __launch_bounds__(256) k_resize(uint8_t* __restrict__ buf, int frame_stride, int which, int level, int tozero,
                                                OrbTables tab) {
  const int per = which == 0 ? c_geom.ncells : 1;
  const int f = blockIdx.z / per, c = blockIdx.z % per;
  const OrbPlane& d = which == 0 ? c_geom.cell[c][level] : c_geom.full[level];
  const OrbPlane& s = which == 0 ? c_geom.cell[c][level - 1] : c_geom.full[level - 1];
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= d.w || y >= d.h) return;
  const uint8_t* sp = buf + (size_t)f * frame_stride + s.off;
  const int ox = tab.ofs[d.tx + x], ax1 = tab.w1[d.tx + x], ax0 = 256 - ax1;
  const int oy = tab.ofs[d.ty + y], ay1 = tab.w1[d.ty + y], ay0 = 256 - ay1;
  const int x1 = min(ox + 1, s.w - 1), y1 = min(oy + 1, s.h - 1);
  const int h0 = sp[oy * s.w + ox] * ax0 + sp[oy * s.w + x1] * ax1;
  const int h1 = sp[y1 * s.w + ox] * ax0 + sp[y1 * s.w + x1] * ax1;
  int v = (h0 * ay0 + h1 * ay1 + (1 << 15)) >> 16;
  if (tozero && v <= 254) v = 0;
  buf[(size_t)f * frame_stride + d.off + (size_t)y * d.w + x] = (uint8_t)v;
}

// -------------------------------------------------------------------------------------------------
// FAST-9/16 corner score without a threshold: S = A - 1, A = max over the 16 arcs of 9 contiguous circle pixels
// of the minimum signed difference (centre brighter: v - p; centre darker: p - v).  A pixel is a FAST corner for
// threshold t exactly when S >= t, and its cv::FAST response is S (cornerScore<16>).
__global__ void __launch_bounds__(256) k_fast_score(const uint8_t* __restrict__ cell_img, uint8_t* __restrict__ score, int level) {
  const int f = blockIdx.z / c_geom.ncells, c = blockIdx.z % c_geom.ncells;
  const OrbPlane& p = c_geom.cell[c][level];
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= p.w || y >= p.h) return;
  const size_t base = (size_t)f * c_geom.cell_bytes + p.off;
  int S = 0;
  if (x >= 3 && y >= 3 && x < p.w - 3 && y < p.h - 3) {
    const uint8_t* im = cell_img + base;
    const int w = p.w;
    const int v = im[y * w + x];
    const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = v - im[(y + dy[k]) * w + x + dx[k]];
    // corner(t): 9 contiguous circle pixels all darker than v - t (d > t) or all brighter than v + t (d < -t).
    // S = largest t for which the pixel is still a corner (binary search; only corners pay for it).
    auto corner = [&](int t) -> bool {
      unsigned hi = 0, lo = 0;
#pragma unroll
      for (int k = 0; k < 16; k++) {
        hi |= (unsigned)(d[k] > t) << k;
        lo |= (unsigned)(d[k] < -t) << k;
      }
      hi |= hi << 16;  // wrap the circle
      lo |= lo << 16;
      unsigned a = hi & (hi >> 1), c2 = lo & (lo >> 1);
      a &= a >> 2; c2 &= c2 >> 2;
      a &= a >> 4; c2 &= c2 >> 4;   // runs of 8
      a &= hi >> 8; c2 &= lo >> 8;  // runs of 9
      return ((a | c2) & 0xFFFFu) != 0u;
    };
    if (corner(1)) {  // the adaptive threshold never goes below 2, scores < 2 are irrelevant
      int lo_t = 1, hi_t = 254;
      while (lo_t < hi_t) {
        const int mid = (lo_t + hi_t + 1) >> 1;
        if (corner(mid)) lo_t = mid;
        else hi_t = mid - 1;
      }
      S = lo_t;
    }
  }
  score[base + (size_t)y * p.w + x] = (uint8_t)min(S, 255);
}

// strict 3x3 non-maximum suppression on S, runByPixelsMask, runByImageBorder(15)  -> candidates + histogram
__global__ void __launch_bounds__(256) k_nms_collect(const uint8_t* __restrict__ score, const uint8_t* __restrict__ cell_mask,
                                                     int level, OrbCand* __restrict__ cand, int* __restrict__ cand_count,
                                                     int* __restrict__ hist) {
  const int f = blockIdx.z / c_geom.ncells, c = blockIdx.z % c_geom.ncells;
  const OrbPlane& p = c_geom.cell[c][level];
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int edge = 15;  // ORB::create(..., edgeThreshold = 15, ...)  feature_adjuster.cpp:94
  if (x < edge || y < edge || x >= p.w - edge || y >= p.h - edge) return;
  const size_t base = (size_t)f * c_geom.cell_bytes + p.off;
  const uint8_t* s = score + base;
  const int w = p.w;
  const int v = s[y * w + x];
  if (v < 2) return;  // the adaptive threshold never drops below 2 (DetectorAdjuster min_thresh)
  if (cell_mask[base + (size_t)y * w + x] == 0) return;
  const bool mx = v > s[(y - 1) * w + x - 1] && v > s[(y - 1) * w + x] && v > s[(y - 1) * w + x + 1] && v > s[y * w + x - 1] &&
                  v > s[y * w + x + 1] && v > s[(y + 1) * w + x - 1] && v > s[(y + 1) * w + x] && v > s[(y + 1) * w + x + 1];
  if (!mx) return;
  const int fc = f * c_geom.ncells + c;
  atomicAdd(&hist[fc * 256 + v], 1);
  const int slot = atomicAdd(&cand_count[fc], 1);
  if (slot < kOrbCandCap) {
    OrbCand cd;
    cd.x = (uint16_t)x;
    cd.y = (uint16_t)y;
    cd.level = (uint8_t)level;
    cd.score = (uint8_t)v;
    cd.pad_ = 0;
    cand[(size_t)fc * kOrbCandCap + slot] = cd;
  }
}

// HarrisResponses(img, pts, blockSize 7, k 0.04): Sobel-3 sums over 7x7, float formula evaluated in the same order
__device__ __forceinline__ float harris_response(const uint8_t* __restrict__ im, int w, int x0, int y0) {
  int a = 0, b = 0, c = 0;
  for (int dy = -3; dy <= 3; dy++) {
#pragma unroll
    for (int dx = -3; dx <= 3; dx++) {
      const uint8_t* p = im + (y0 + dy) * w + x0 + dx;
      const int Ix = (p[1] - p[-1]) * 2 + (p[-w + 1] - p[-w - 1]) + (p[w + 1] - p[w - 1]);
      const int Iy = (p[w] - p[-w]) * 2 + (p[w - 1] - p[-w - 1]) + (p[w + 1] - p[-w + 1]);
      a += Ix * Ix;
      b += Iy * Iy;
      c += Ix * Iy;
    }
  }
  const float scale = __fdiv_rn(1.f, 4.f * 7.f * 255.f);
  const float s4 = __fmul_rn(__fmul_rn(__fmul_rn(scale, scale), scale), scale);
  const float fa = (float)a, fb = (float)b, fc = (float)c;
  const float sum = __fadd_rn(fa, fb);
  const float r = __fsub_rn(__fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc)), __fmul_rn(__fmul_rn(0.04f, sum), sum));
  return __fmul_rn(r, s4);
}

// cv::fastAtan2 (degrees)
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
  const float k = 57.29577951308232f;  // (float)(180/CV_PI)
  const float p1 = __fmul_rn(0.9997878412794807f, k), p3 = __fmul_rn(-0.3258083974640975f, k);
  const float p5 = __fmul_rn(0.1555786518463281f, k), p7 = __fmul_rn(-0.04432655554792128f, k);
  const float ax = fabsf(x), ay = fabsf(y);
  const float eps = 2.220446049250313e-16f;
  float a, c, c2;
  if (ax >= ay) {
    c = __fdiv_rn(ay, __fadd_rn(ax, eps));
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = __fdiv_rn(ax, __fadd_rn(ay, eps));
    c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

// IC_Angle: intensity-centroid orientation over the radius-15 disc; one warp per keypoint
__device__ float ic_angle_warp(const uint8_t* __restrict__ im, int w, int x0, int y0, int lane) {
  int m01 = 0, m10 = 0;
  const uint8_t* ctr = im + y0 * w + x0;
  if (lane < 31) m10 += (lane - 15) * ctr[lane - 15];  // v = 0 row
  for (int v = 1; v <= kOrbHalfPatch; v++) {
    const int d = c_umax[v];
    const int u = lane - d;  // lanes cover u = -d .. d (d <= 15 -> <= 31 lanes)
    if (u <= d) {
      const int vp = ctr[u + v * w], vm = ctr[u - v * w];
      m01 += v * (vp - vm);
      m10 += u * (vp + vm);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    m10 += __shfl_xor_sync(0xffffffffu, m10, o);
  }
  return fast_atan2_deg((float)m01, (float)m10);
}

// Harris response for candidates with S >= the cell's final threshold; others get NaN (excluded).
__global__ void __launch_bounds__(256) k_harris(const uint8_t* __restrict__ cell_img, const OrbCand* __restrict__ cand,
                                                const int* __restrict__ cand_count, const int* __restrict__ thr,
                                                float* __restrict__ resp) {
  const int fc = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(cand_count[fc], kOrbCandCap);
  if (i >= n) return;
  const OrbCand cd = cand[(size_t)fc * kOrbCandCap + i];
  float r = __int_as_float(0x7fc00000);
  if (cd.score >= thr[fc]) {
    const int f = fc / c_geom.ncells, c = fc % c_geom.ncells;
    const OrbPlane& p = c_geom.cell[c][cd.level];
    r = harris_response(cell_img + (size_t)f * c_geom.cell_bytes + p.off, p.w, cd.x, cd.y);
  }
  resp[(size_t)fc * kOrbCandCap + i] = r;
}

__device__ __forceinline__ uint32_t f32_ordered(float f) {  // ascending unsigned order == ascending float order
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ void bitonic_sort_u64(unsigned long long* keys, int N) {
  for (int k = 2; k <= N; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (N >> 1); t += blockDim.x) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int hi = lo | j;
        const bool up = (lo & k) == 0;
        const unsigned long long a = keys[lo], b = keys[hi];
        if ((a > b) == up) {
          keys[lo] = b;
          keys[hi] = a;
        }
      }
      __syncthreads();
    }
  }
}

// keepStrongest(maxPerCell) by |response| (feature_adjuster.cpp:247-255; nth_element order is unspecified in the
// reference -> canonical tie rule: (level, y, x) ascending).  One CTA per (frame, cell).
// key = [~ordered(|resp|) : 32][level:3 y:10 x:10 : 23][0 : 9]
__global__ void __launch_bounds__(1024) k_cell_select(const OrbCand* __restrict__ cand, const int* __restrict__ cand_count,
                                                      const float* __restrict__ resp, int max_per_cell,
                                                      unsigned long long* __restrict__ cell_out, int* __restrict__ cell_out_count,
                                                      int out_stride) {
  extern __shared__ unsigned long long keys[];
  __shared__ int s_n;
  const int fc = blockIdx.x;
  const int n = min(cand_count[fc], kOrbCandCap);
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float r = resp[(size_t)fc * kOrbCandCap + i];
    if (r == r) {
      const OrbCand cd = cand[(size_t)fc * kOrbCandCap + i];
      const int slot = atomicAdd(&s_n, 1);
      const uint32_t pos = ((uint32_t)cd.level << 20) | ((uint32_t)cd.y << 10) | cd.x;
      keys[slot] = ((unsigned long long)(~f32_ordered(fabsf(r))) << 32) | ((unsigned long long)pos << 9);
      // the signed response is recovered later from the candidate position (recomputed) -- keep sign in bit 0
      keys[slot] |= (r < 0.f) ? 1ull : 0ull;
    }
  }
  __syncthreads();
  const int cnt = s_n;
  int N = 2;
  while (N < cnt) N <<= 1;
  for (int i = cnt + threadIdx.x; i < N; i += blockDim.x) keys[i] = ~0ull;
  __syncthreads();
  bitonic_sort_u64(keys, N);
  const int keep = min(cnt, max_per_cell);
  for (int i = threadIdx.x; i < keep; i += blockDim.x) cell_out[(size_t)fc * out_stride + i] = keys[i];
  if (threadIdx.x == 0) cell_out_count[fc] = keep;
}

// -------------------------------------------------------------------------------------------------
// Per frame: aggregate the cells (feature_adjuster.cpp:259-282), removeDepthless (node.cpp:67-97), retainBest(K)
// (node.cpp:187-191), the extractor's border filter + octave sort (cv::ORB::compute), orientation, projectTo3D
// (node.cpp:900-965).  mode 0: stop after aggregation (== detector->detect output, with angles);
// mode 1: full Node constructor.
struct FrameKp {
  float x, y, resp;
  uint16_t lx, ly;  // level coordinates inside the cell pyramid
  uint8_t level, cell;
  uint16_t flag;
};
constexpr int kFrameCap = 4096;  // >= ncells * max_per_cell

// mode 0: detector output, cell-major, inside a cell |response| descending (canonical stand-in for the unspecified
//         nth_element order), ties by (level, y, x).
// mode 1: Node constructor: removeDepthless -> retainBest(K) by signed response (ties canonical, cut at K) ->
//         extractor border filter (31 px on cvRound'ed coordinates) -> stable sort by octave -> orientation ->
//         projectTo3D.  Final order = (octave, response descending, cell, level, y, x).
__global__ void __launch_bounds__(1024)
    k_frame_finalize(int mode, int max_keypoints, const unsigned long long* __restrict__ cell_out,
                     const int* __restrict__ cell_out_count, int out_stride, const uint8_t* __restrict__ cell_img,
                     const float* __restrict__ depth, float depth_scaling, float4 Kinv /* 1/fx, 1/fy, cx, cy */,
                     FrameKp* __restrict__ scratch /* nframes x 2 x kFrameCap */, rgbdslam_b200_keypoint* __restrict__ kp_out,
                     float4* __restrict__ xyz_out, int* __restrict__ n_out, int kp_stride) {
  __shared__ unsigned long long keys[kFrameCap];
  __shared__ int s_n, s_m;
  const int f = blockIdx.x;
  const int W = c_geom.W, H = c_geom.H;
  FrameKp* ka = scratch + (size_t)f * 2 * kFrameCap;  // gather order
  FrameKp* kc = ka + kFrameCap;                       // canonical order
  if (threadIdx.x == 0) { s_n = 0; s_m = 0; }
  __syncthreads();
  // A. gather, shift to image coordinates (pt *= scale; pt += cell origin: feature_adjuster.cpp:259-282), depth check
  for (int c = 0; c < c_geom.ncells; c++) {
    const int fc = f * c_geom.ncells + c;
    const int n = cell_out_count[fc];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned long long k = cell_out[(size_t)fc * out_stride + i];
      const uint32_t pos = (uint32_t)((k >> 9) & 0x7FFFFFu);
      const int level = pos >> 20, ly = (pos >> 10) & 1023, lx = pos & 1023;
      const uint32_t ord = ~(uint32_t)(k >> 32);
      float r = __uint_as_float(ord & 0x7FFFFFFFu);  // |resp| (ordered() of a non-negative float only sets bit 31)
      if (k & 1ull) r = -r;
      const float sc = c_geom.cell[c][level].scale;
      FrameKp q;
      q.x = __fadd_rn(__fmul_rn((float)lx, sc), (float)c_geom.cell_x0[c]);
      q.y = __fadd_rn(__fmul_rn((float)ly, sc), (float)c_geom.cell_y0[c]);
      q.resp = r;
      q.lx = (uint16_t)lx;
      q.ly = (uint16_t)ly;
      q.level = (uint8_t)level;
      q.cell = (uint8_t)c;
      q.flag = 0;
      bool ok = true;
      if (mode == 1) {  // removeDepthless (node.cpp:67-97)
        ok = !(q.x >= W || q.x < 0 || q.y >= H || q.y < 0);
        if (ok) {
          const int rx = (int)floorf(q.x + 0.5f), ry = (int)floorf(q.y + 0.5f);  // round(): half away from zero
          const size_t idx = (size_t)ry * W + rx;
          const float Z = idx < (size_t)W * H ? depth[(size_t)f * W * H + idx] : __int_as_float(0x7fc00000);
          ok = !(Z != Z);
        }
      }
      if (ok) {
        const int slot = atomicAdd(&s_n, 1);
        if (slot < kFrameCap) ka[slot] = q;
      }
    }
  }
  __syncthreads();
  const int cnt = min(s_n, kFrameCap);
  int N = 2;
  while (N < cnt) N <<= 1;
  // B. canonical order (cell, level, y, x)
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    unsigned long long k = ~0ull;
    if (i < cnt) {
      const FrameKp q = ka[i];
      const unsigned long long canon = ((unsigned long long)q.cell << 27) | ((unsigned long long)q.level << 24) |
                                       ((unsigned long long)q.ly << 12) | q.lx;
      k = (canon << 16) | (unsigned)i;
    }
    keys[i] = k;
  }
  __syncthreads();
  bitonic_sort_u64(keys, N);
  for (int r = threadIdx.x; r < cnt; r += blockDim.x) kc[r] = ka[keys[r] & 0xFFFFu];
  __threadfence_block();
  __syncthreads();
  int n_final = cnt;
  if (mode == 1) {
    // C. retainBest(max_keypoints) (node.cpp:187-191): signed response descending, ties canonical, cut at K
    for (int r = threadIdx.x; r < N; r += blockDim.x)
      keys[r] = r < cnt ? (((unsigned long long)(~f32_ordered(kc[r].resp)) << 32) | (unsigned)r) : ~0ull;
    __syncthreads();
    bitonic_sort_u64(keys, N);
    const int keepK = min(cnt, max_keypoints);
    // D. extractor->compute(): runByImageBorder(31) on cvRound'ed coordinates, then stable sort by octave
    unsigned long long mine[4];  // up to 4 keys per thread (kFrameCap / 1024)
    int nm = 0;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      unsigned long long k2 = ~0ull;
      if (j < keepK) {
        const int r = (int)(keys[j] & 0xFFFFu);
        const FrameKp q = kc[r];
        const int rx = __float2int_rn(q.x), ry = __float2int_rn(q.y);
        if (rx >= 31 && rx < W - 31 && ry >= 31 && ry < H - 31) {
          k2 = ((unsigned long long)q.level << 40) | ((unsigned long long)j << 16) | (unsigned)r;
          atomicAdd(&s_m, 1);
        }
      }
      mine[nm++] = k2;
    }
    __syncthreads();
    nm = 0;
    for (int j = threadIdx.x; j < N; j += blockDim.x) keys[j] = mine[nm++];
    __syncthreads();
    bitonic_sort_u64(keys, N);
    n_final = s_m;
  } else {
    for (int r = threadIdx.x; r < N; r += blockDim.x) {
      unsigned long long k = ~0ull;
      if (r < cnt) {
        const FrameKp q = kc[r];
        k = ((unsigned long long)q.cell << 56) | ((unsigned long long)(~f32_ordered(fabsf(q.resp))) << 16) | (unsigned)r;
      }
      keys[r] = k;
    }
    __syncthreads();
    bitonic_sort_u64(keys, N);
  }
  // E. emit in final order; orientation by one warp per keypoint on the detector's (cell) pyramid
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int t = warp; t < n_final; t += nwarps) {
    const FrameKp q = kc[keys[t] & 0xFFFFu];
    const OrbPlane& p = c_geom.cell[q.cell][q.level];
    const float ang = ic_angle_warp(cell_img + (size_t)f * c_geom.cell_bytes + p.off, p.w, q.lx, q.ly, lane);
    if (lane == 0) {
      rgbdslam_b200_keypoint o;
      o.x = q.x;
      o.y = q.y;
      o.size = __fmul_rn(31.f, p.scale);
      o.angle = ang;
      o.response = q.resp;
      o.octave = q.level;
      o.class_id = -1;
      kp_out[(size_t)f * kp_stride + t] = o;
      if (mode == 1) {  // projectTo3D (node.cpp:900-965) + backProject (misc2.h:49-65)
        const int rx = (int)floorf(q.x + 0.5f), ry = (int)floorf(q.y + 0.5f);
        const float Z = (float)((double)depth[(size_t)f * W * H + (size_t)ry * W + rx] * (double)depth_scaling);
        float4 v;
        v.x = __fmul_rn(__fmul_rn(__fsub_rn(q.x, Kinv.z), Z), Kinv.x);
        v.y = __fmul_rn(__fmul_rn(__fsub_rn(q.y, Kinv.w), Z), Kinv.y);
        v.z = Z;
        v.w = 1.f;
        xyz_out[(size_t)f * kp_stride + t] = v;
      }
    }
  }
  if (threadIdx.x == 0) n_out[f] = n_final;
}

// -------------------------------------------------------------------------------------------------
// GaussianBlur(level, 7x7, sigma 2, BORDER_REFLECT_101) as OpenCV evaluates it inside ORB: separable float filter,
// row pass accumulated left to right, column pass symmetric, round-half-even to uint8.
__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

__global__ void __launch_bounds__(256) k_blur(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int frame_stride,
                                              int level) {
  const OrbPlane& p = c_geom.full[level];
  const int f = blockIdx.z;
  __shared__ float rows[22][32];  // 16 output rows + 6 halo rows, 32 columns
  const uint8_t* im = src + (size_t)f * frame_stride + p.off;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int x = blockIdx.x * 32 + tx, y0 = blockIdx.y * 16;
  for (int r = ty; r < 22; r += 8) {
    const int yy = reflect101(y0 + r - 3, p.h);
    float acc = 0.f;
    if (x < p.w) {
#pragma unroll
      for (int j = 0; j < 7; j++) {
        const int xx = reflect101(x + j - 3, p.w);
        acc = __fadd_rn(acc, __fmul_rn(c_gauss[j], (float)im[yy * p.w + xx]));
      }
    }
    rows[r][tx] = acc;
  }
  __syncthreads();
  for (int r = ty; r < 16; r += 8) {
    const int y = y0 + r;
    if (x < p.w && y < p.h) {
      float c = __fmul_rn(c_gauss[3], rows[r + 3][tx]);
#pragma unroll
      for (int j = 1; j <= 3; j++) c = __fadd_rn(c, __fmul_rn(c_gauss[3 + j], __fadd_rn(rows[r + 3 + j][tx], rows[r + 3 - j][tx])));
      int v = __float2int_rn(c);
      v = min(max(v, 0), 255);
      dst[(size_t)f * frame_stride + p.off + (size_t)y * p.w + x] = (uint8_t)v;
    }
  }
}

// rBRIEF: one warp per keypoint, lane j produces descriptor byte j (8 tests).  Pixels outside the level are read
// from the UNBLURRED level with reflect-101 (OpenCV blurs only the level ROI of its bordered pyramid buffer).
__global__ void __launch_bounds__(256) k_describe(const uint8_t* __restrict__ pyr_raw, const uint8_t* __restrict__ pyr_blur,
                                                  int frame_stride, const rgbdslam_b200_keypoint* __restrict__ kps,
                                                  const int* __restrict__ n_kp, int kp_stride, uint8_t* __restrict__ desc) {
  const int f = blockIdx.y;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n_kp[f]) return;
  const rgbdslam_b200_keypoint kp = kps[(size_t)f * kp_stride + i];
  const OrbPlane& p = c_geom.full[kp.octave];
  const float sinv = __fdiv_rn(1.f, p.scale);
  const int cx = __float2int_rn(__fmul_rn(kp.x, sinv)), cy = __float2int_rn(__fmul_rn(kp.y, sinv));
  const float ang = __fmul_rn(kp.angle, 0.017453292519943295f);  // angle *= (float)(CV_PI/180.f)
  const float a = (float)cos((double)ang), b = (float)sin((double)ang);
  const uint8_t* raw = pyr_raw + (size_t)f * frame_stride + p.off;
  const uint8_t* blr = pyr_blur + (size_t)f * frame_stride + p.off;
  auto pix = [&](int k, int which) -> int {
    const float px = (float)c_pattern[k][2 * which], py = (float)c_pattern[k][2 * which + 1];
    const int ix = __float2int_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)));
    const int iy = __float2int_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)));
    const int xx = cx + ix, yy = cy + iy;
    if (xx >= 0 && yy >= 0 && xx < p.w && yy < p.h) return blr[yy * p.w + xx];
    return raw[reflect101(yy, p.h) * p.w + reflect101(xx, p.w)];
  };
  unsigned v = 0;
#pragma unroll
  for (int t = 0; t < 8; t++) {
    const int k = lane * 8 + t;
    v |= (unsigned)(pix(k, 0) < pix(k, 1)) << t;
  }
  desc[((size_t)f * kp_stride + i) * 32 + lane] = (uint8_t)v;
}

// ================================================================================================
// launch helpers (host)
static inline dim3 plane_grid(int w, int h, int z) { return dim3((w + 31) / 32, (h + 7) / 8, z); }

cudaError_t orb_run_detect(const OrbGeom& g, const OrbTables& tab, int nframes, const uint8_t* d_gray, const uint8_t* d_mask,
                           uint8_t* d_cell_img, uint8_t* d_cell_mask, uint8_t* d_score, OrbCand* d_cand, int* d_cand_count,
                           int* d_hist, cudaStream_t st, int* launches) {
  int maxw = 0, maxh = 0;
  for (int c = 0; c < g.ncells; c++) {
    maxw = g.cell[c][0].w > maxw ? g.cell[c][0].w : maxw;
    maxh = g.cell[c][0].h > maxh ? g.cell[c][0].h : maxh;
  }
  const int z = nframes * g.ncells;
  cudaMemsetAsync(d_cand_count, 0, sizeof(int) * z, st);
  cudaMemsetAsync(d_hist, 0, sizeof(int) * 256 * z, st);
  k_cell_extract<<<plane_grid(maxw, maxh, z), 256, 0, st>>>(d_gray, d_mask, d_cell_img, d_cell_mask);
  (*launches)++;
  for (int l = 0; l < kOrbLevels; l++) {
    int lw = 0, lh = 0;
    for (int c = 0; c < g.ncells; c++) {
      lw = g.cell[c][l].w > lw ? g.cell[c][l].w : lw;
      lh = g.cell[c][l].h > lh ? g.cell[c][l].h : lh;
    }
    if (l > 0) {
      k_resize<<<plane_grid(lw, lh, z), 256, 0, st>>>(d_cell_img, g.cell_bytes, 0, l, 0, tab);
      k_resize<<<plane_grid(lw, lh, z), 256, 0, st>>>(d_cell_mask, g.cell_bytes, 0, l, 1, tab);
      (*launches) += 2;
    }
    k_fast_score<<<plane_grid(lw, lh, z), 256, 0, st>>>(d_cell_img, d_score, l);
    k_nms_collect<<<plane_grid(lw, lh, z), 256, 0, st>>>(d_score, d_cell_mask, l, d_cand, d_cand_count, d_hist);
    (*launches) += 2;
  }
  return cudaGetLastError();
}

cudaError_t orb_run_select(const OrbGeom& g, int nframes, int mode, int max_per_cell, int max_keypoints,
                           const uint8_t* d_cell_img, const OrbCand* d_cand, const int* d_cand_count, const int* d_thr,
                           float* d_resp, unsigned long long* d_cell_out, int* d_cell_out_count, const float* d_depth,
                           float depth_scaling, float4 Kinv, void* d_scratch, rgbdslam_b200_keypoint* d_kp, float4* d_xyz, int* d_n,
                           int kp_stride, cudaStream_t st, int* launches) {
  const int z = nframes * g.ncells;
  k_harris<<<dim3((kOrbCandCap + 255) / 256, z), 256, 0, st>>>(d_cell_img, d_cand, d_cand_count, d_thr, d_resp);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_cell_select, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  k_cell_select<<<z, 1024, 16384 * 8, st>>>(d_cand, d_cand_count, d_resp, max_per_cell, d_cell_out, d_cell_out_count, max_per_cell);
  k_frame_finalize<<<nframes, 1024, 0, st>>>(mode, max_keypoints, d_cell_out, d_cell_out_count, max_per_cell, d_cell_img, d_depth,
                                             depth_scaling, Kinv, (FrameKp*)d_scratch, d_kp, d_xyz, d_n, kp_stride);
  (*launches) += 3;
  return cudaGetLastError();
}

cudaError_t orb_run_describe(const OrbGeom& g, const OrbTables& tab, int nframes, const uint8_t* d_gray, uint8_t* d_pyr_raw,
                             uint8_t* d_pyr_blur, const rgbdslam_b200_keypoint* d_kp, const int* d_n, int kp_stride, int max_kp,
                             uint8_t* d_desc, cudaStream_t st, int* launches) {
  // level 0 = the image itself
  cudaError_t e = cudaMemcpy2DAsync(d_pyr_raw, g.full_bytes, d_gray, (size_t)g.W * g.H, (size_t)g.W * g.H, nframes,
                                    cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return e;
  for (int l = 1; l < kOrbLevels; l++) {
    k_resize<<<plane_grid(g.full[l].w, g.full[l].h, nframes), 256, 0, st>>>(d_pyr_raw, g.full_bytes, 1, l, 0, tab);
    (*launches)++;
  }
  for (int l = 0; l < kOrbLevels; l++) {
    k_blur<<<dim3((g.full[l].w + 31) / 32, (g.full[l].h + 15) / 16, nframes), 256, 0, st>>>(d_pyr_raw, d_pyr_blur, g.full_bytes, l);
    (*launches)++;
  }
  k_describe<<<dim3((max_kp + 7) / 8, nframes), 256, 0, st>>>(d_pyr_raw, d_pyr_blur, g.full_bytes, d_kp, d_n, kp_stride, d_desc);
  (*launches)++;
  return cudaGetLastError();
}

}  // namespace rb200
