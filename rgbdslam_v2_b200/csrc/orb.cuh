// orb.cuh -- device-side geometry of the ORB detect / describe pipeline (orb.cu).
//
// OpenCV's ORB is an un-vendored dependency of the reference (feature_adjuster.cpp:94 builds
// cv::ORB::create(10000, 1.2, 8, 15, 0, 2, HARRIS_SCORE, 31, thresh) per grid cell; features.cpp:117-119 builds the
// default cv::ORB::create() extractor).  Its published algorithm is restated here; every arithmetic detail was
// pinned against cv2 4.13 (oracle/orb_oracle.py, tests/test_orb_oracle.py): chained INTER_LINEAR_EXACT pyramid,
// FAST-9/16 corner score + 3x3 NMS, Harris response, intensity-centroid angle (fastAtan2), 7-tap float Gaussian,
// rBRIEF with the learned bit_pattern_31_.
#pragma once
#include <stdint.h>

namespace rb200 {

constexpr int kOrbLevels = 8;
constexpr int kOrbMaxCells = 16;       // detector_grid_resolution <= 4
constexpr int kOrbCandCap = 12288;     // FAST/NMS candidates per (frame, cell), all levels
constexpr int kOrbHalfPatch = 15;

// One image plane of a pyramid: where it lives in the packed per-frame buffer and its resize tables.
struct OrbPlane {
  int32_t w, h;
  int32_t off;         // byte offset inside the per-frame packed buffer
  int32_t tx, ty;      // offsets (in entries) into the table arrays for the resize that PRODUCES this plane
  float scale;         // layerScale (level 0: 1.0)
};

// Geometry shared by all frames of one image size.
struct OrbGeom {
  int32_t W, H;                 // full image
  int32_t ncells, grid;         // grid x grid cells (1 = no grid)
  int32_t cell_x0[kOrbMaxCells], cell_y0[kOrbMaxCells];
  OrbPlane cell[kOrbMaxCells][kOrbLevels];   // detector pyramids (one per grid cell)
  int32_t cell_bytes;           // packed bytes per frame for all cell pyramids (image; mask uses the same layout)
  OrbPlane full[kOrbLevels];    // extractor pyramid of the whole image
  int32_t full_bytes;
  int32_t n_per_level[kOrbLevels];  // ORB feature quota per level for nfeatures = 10000
};

// resize tables: for every destination column/row the first source index and the weight of the second tap (x256)
struct OrbTables {
  const int16_t* ofs;
  const uint16_t* w1;
};

struct OrbCand {   // a FAST corner that survived NMS, mask and the 15 px border filter
  uint16_t x, y;   // level coordinates
  uint8_t level;
  uint8_t score;   // FAST corner score (largest threshold for which it is still a corner)
  uint16_t pad_;
};

}  // namespace rb200
