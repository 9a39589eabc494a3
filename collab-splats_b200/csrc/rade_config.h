// Conventions shared by every kernel.  Each constant mirrors the same-named constant in
// oracle/rade_oracle.py; SURVEY.md section 8c lists the open questions (Q1-Q7) they settle.
#pragma once

#define RS_TILE 16                       // tile edge in pixels (gsplat default tile_size=16)
#define RS_ALPHA_MIN (1.0f / 255.0f)     // skip pair if alpha < 1/255                 (A8)
#define RS_ALPHA_MAX 0.999f              // Q3                                          (A8)
#define RS_T_STOP 1e-4f                  // stop when T*(1-alpha) <= 1e-4               (A8)
#define RS_RADIUS_SIGMA 3.33f            // radius = ceil(3.33*sqrt(cov_ii))            (A4)
#define RS_DET_MIN 1e-10f                // det = max(det, 1e-10)                       (A4)
#define RS_FOV_PAD 0.3f                  // frustum clamp padding                       (A3)
#define RS_VBN_EPS 1e-7f                 // Q5                                          (A5)
#define RS_MEDIAN_INCLUSIVE 1            // Q2: T > 0.5 && T' <= 0.5
#define RS_NORMALIZE_EXPECTED_DEPTH 0    // Q1: raw sum(vis*t)/ln

// status codes returned by every extern "C" entry point
#define RS_OK 0
#define RS_ERR_BAD_ARG (-1)
#define RS_ERR_LAUNCH (-2)
#define RS_ERR_UNSUPPORTED (-3)
