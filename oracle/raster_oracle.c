/* CPU oracle of the per-tile compositing stage, in plain C.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates gsplat-rade `rasterize_to_pixels` forward and backward (SURVEY.md rows a10/a11, Appendix A8/A9;
 * reached from collab_splats/models/rade_gs_model.py:439-465 and rade_features_model.py:450-476) the way the
 * upstream kernels are organised: one independent loop per PIXEL over its tile's depth-sorted list.  It mirrors
 * oracle/rade_oracle.py:rasterize_to_pixels (the PyTorch restatement, whose backward is autograd) operation for
 * operation -- tests/test_c_oracle.py holds the two against each other in fp32 and fp64 -- but runs a full
 * 1920x1080 view of a million Gaussians in about a second on the host cores, so that
 *   (a) the parity tests can compare COMPLETE images and gradients at the BASELINE sizes, and
 *   (b) bench.py's `cpu_baseline` / `--impl reference` legs time the complete workload instead of a window.
 * PARITY UNPINNED (see rade_oracle.py): gsplat-rade is not under /root/reference; the conventions Q1-Q7 are the
 * named constants below and must equal rade_oracle.py's and csrc/rade_config.h's.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library; the product never does.
 * Build: oracle/build_oracle.py (gcc -O2 -shared -fPIC -pthread; no -ffast-math, no FMA contraction). */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define RO_TILE 16
#define RO_ALPHA_MIN (1.0 / 255.0)
#define RO_ALPHA_MAX 0.999
#define RO_T_STOP 1e-4
#define RO_MEDIAN_INCLUSIVE 1

#define REAL float
#define SUFFIX _f32
#define REAL_IS_DOUBLE 0
#define EXP expf
#define SQRT sqrtf
#define FABS fabsf
#include "raster_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef REAL_IS_DOUBLE
#undef EXP
#undef SQRT
#undef FABS

#define REAL double
#define SUFFIX _f64
#define REAL_IS_DOUBLE 1
#define EXP exp
#define SQRT sqrt
#define FABS fabs
#include "raster_oracle_impl.h"

int ro_version(void) { return 1; }
