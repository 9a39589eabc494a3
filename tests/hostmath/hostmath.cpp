// TEST-ONLY host build of the per-element math in collab-splats_b200/csrc/rade_math.cuh.
// Compiled by tests/conftest.py with g++ -ffp-contract=off and called through ctypes so the
// hand-derived VJPs and the exact-rounding contract can be checked against oracle autograd on a
// machine without a GPU.  Never linked into the product library.
#include <cstdint>
#include <cstring>
#include "rade_math.cuh"

using namespace rs;

template <typename T>
static void load_cam(const T* viewmat, const T* K, Cam<T>& cam) {
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) cam.W[i * 3 + j] = viewmat[i * 4 + j];
    cam.t[i] = viewmat[i * 4 + 3];
  }
  cam.fx = K[0]; cam.fy = K[4]; cam.cx = K[2]; cam.cy = K[5];
}

template <typename T>
static void project_fwd(const T* means, const T* quats, const T* scales, const T* viewmat, const T* K, int N,
                        int W, int H, T eps2d, T near_, T far_, T clip, int32_t* radii, T* m2, T* depths, T* conics,
                        T* comps, T* ray_ts, T* ray_planes, T* normals) {
  Cam<T> cam;
  load_cam(viewmat, K, cam);
  ProjParams<T> pp{(T)W, (T)H, eps2d, near_, far_, clip};
  for (int n = 0; n < N; ++n) {
    ProjOut<T> o;
    project_fwd_one(means + 3 * n, quats + 4 * n, scales + 3 * n, cam, pp, o);
    radii[2 * n] = o.rx; radii[2 * n + 1] = o.ry;
    m2[2 * n] = o.m2x; m2[2 * n + 1] = o.m2y;
    depths[n] = o.depth;
    conics[3 * n] = o.ca; conics[3 * n + 1] = o.cb; conics[3 * n + 2] = o.cc;
    comps[n] = o.comp;
    ray_ts[n] = o.ray_t;
    ray_planes[2 * n] = o.rp0; ray_planes[2 * n + 1] = o.rp1;
    normals[3 * n] = o.nx; normals[3 * n + 1] = o.ny; normals[3 * n + 2] = o.nz;
  }
}

template <typename T>
static void project_bwd(const T* means, const T* quats, const T* scales, const T* viewmat, const T* K, int N,
                        int W, int H, T eps2d, T near_, T far_, T clip, const T* v_m2, const T* v_depths,
                        const T* v_conics, const T* v_comps, const T* v_ray_ts, const T* v_ray_planes,
                        const T* v_normals, T* v_means, T* v_quats, T* v_scales, T* v_W, T* v_t) {
  Cam<T> cam;
  load_cam(viewmat, K, cam);
  ProjParams<T> pp{(T)W, (T)H, eps2d, near_, far_, clip};
  for (int n = 0; n < N; ++n) {
    ProjGradIn<T> g;
    g.v_m2x = v_m2[2 * n]; g.v_m2y = v_m2[2 * n + 1];
    g.v_depth = v_depths[n];
    g.v_ca = v_conics[3 * n]; g.v_cb = v_conics[3 * n + 1]; g.v_cc = v_conics[3 * n + 2];
    g.v_comp = v_comps[n];
    g.v_ray_t = v_ray_ts[n];
    g.v_rp0 = v_ray_planes[2 * n]; g.v_rp1 = v_ray_planes[2 * n + 1];
    g.v_nx = v_normals[3 * n]; g.v_ny = v_normals[3 * n + 1]; g.v_nz = v_normals[3 * n + 2];
    project_bwd_one(means + 3 * n, quats + 4 * n, scales + 3 * n, cam, pp, g, v_means + 3 * n, v_quats + 4 * n,
                    v_scales + 3 * n, v_W, v_t);
  }
}

template <typename T>
static void sh_fwd(int deg, int K, int N, const T* dirs, const T* coeffs, T* colors) {
  for (int n = 0; n < N; ++n) {
    T x = dirs[3 * n], y = dirs[3 * n + 1], z = dirs[3 * n + 2];
    T inv = T(1) / fmax(sqrt(x * x + y * y + z * z), T(1e-12));
    T b[16];
    sh_basis(deg, x * inv, y * inv, z * inv, b);
    int nb = (deg + 1) * (deg + 1);
    for (int c = 0; c < 3; ++c) {
      T acc = 0;
      for (int q = 0; q < nb; ++q) acc += b[q] * coeffs[(n * K + q) * 3 + c];
      colors[3 * n + c] = acc;
    }
  }
}

template <typename T>
static void sh_bwd(int deg, int K, int N, const T* dirs, const T* coeffs, const T* v_colors, T* v_coeffs,
                   T* v_dirs) {
  for (int n = 0; n < N; ++n) {
    T x = dirs[3 * n], y = dirs[3 * n + 1], z = dirs[3 * n + 2];
    T nrm = fmax(sqrt(x * x + y * y + z * z), T(1e-12));
    T inv = T(1) / nrm;
    T ux = x * inv, uy = y * inv, uz = z * inv;
    T b[16], g[16];
    sh_basis(deg, ux, uy, uz, b);
    int nb = (deg + 1) * (deg + 1);
    for (int q = 0; q < nb; ++q) {
      T acc = 0;
      for (int c = 0; c < 3; ++c) {
        v_coeffs[(n * K + q) * 3 + c] = b[q] * v_colors[3 * n + c];
        acc += v_colors[3 * n + c] * coeffs[(n * K + q) * 3 + c];
      }
      g[q] = acc;
    }
    T vx, vy, vz;
    sh_basis_vjp(deg, ux, uy, uz, g, vx, vy, vz);
    T dd = ux * vx + uy * vy + uz * vz;
    v_dirs[3 * n] = (vx - ux * dd) * inv;
    v_dirs[3 * n + 1] = (vy - uy * dd) * inv;
    v_dirs[3 * n + 2] = (vz - uz * dd) * inv;
  }
}

#define INST(SUF, T)                                                                                                 \
  extern "C" void hm_project_fwd_##SUF(const T* means, const T* quats, const T* scales, const T* viewmat,            \
                                       const T* K, int N, int W, int H, T eps2d, T near_, T far_, T clip,            \
                                       int32_t* radii, T* m2, T* depths, T* conics, T* comps, T* ray_ts,             \
                                       T* ray_planes, T* normals) {                                                  \
    project_fwd<T>(means, quats, scales, viewmat, K, N, W, H, eps2d, near_, far_, clip, radii, m2, depths, conics,   \
                   comps, ray_ts, ray_planes, normals);                                                              \
  }                                                                                                                  \
  extern "C" void hm_project_bwd_##SUF(const T* means, const T* quats, const T* scales, const T* viewmat,            \
                                       const T* K, int N, int W, int H, T eps2d, T near_, T far_, T clip,            \
                                       const T* v_m2, const T* v_depths, const T* v_conics, const T* v_comps,        \
                                       const T* v_ray_ts, const T* v_ray_planes, const T* v_normals, T* v_means,     \
                                       T* v_quats, T* v_scales, T* v_W, T* v_t) {                                    \
    project_bwd<T>(means, quats, scales, viewmat, K, N, W, H, eps2d, near_, far_, clip, v_m2, v_depths, v_conics,    \
                   v_comps, v_ray_ts, v_ray_planes, v_normals, v_means, v_quats, v_scales, v_W, v_t);                \
  }                                                                                                                  \
  extern "C" void hm_sh_fwd_##SUF(int deg, int K, int N, const T* dirs, const T* coeffs, T* colors) {                \
    sh_fwd<T>(deg, K, N, dirs, coeffs, colors);                                                                      \
  }                                                                                                                  \
  extern "C" void hm_sh_bwd_##SUF(int deg, int K, int N, const T* dirs, const T* coeffs, const T* v_colors,          \
                                  T* v_coeffs, T* v_dirs) {                                                          \
    sh_bwd<T>(deg, K, N, dirs, coeffs, v_colors, v_coeffs, v_dirs);                                                  \
  }

INST(f32, float)
INST(f64, double)
