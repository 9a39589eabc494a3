// Per-element math of the RaDe-GS hot path, written once for host and device.
//
// * project_fwd_one / project_bwd_one : EWA projection with ray-space depth/normal terms
//   (replaces gsplat-rade `fully_fused_projection` fwd/bwd; SURVEY.md a5/a12, Appendix A1-A5;
//   reference call sites collab_splats/models/rade_gs_model.py:373-389 and :439-465).
// * sh_fwd_one / sh_bwd_one : real spherical harmonics up to degree 3 (SURVEY a6;
//   reference call site collab_splats/models/rade_features_model.py:430-434).
//
// The functions are templated on the scalar type so tests/hostmath can run them in double
// on the CPU against oracle autograd; the kernels instantiate float only.
//
// Arithmetic-order contract (see oracle/rade_oracle.py header): everything that determines
// means2d, depths, the blurred 2-D covariance, conics and radii is a fixed sequence of
// individually rounded IEEE operations (struct X<T>: __fmul_rn/__fadd_rn/... on device, plain
// ops under -ffp-contract=off on host), so tile lists and sort keys are reproducible bit for
// bit against the CPU oracle.  The RaDe terms and the whole backward use ordinary arithmetic.
#pragma once
#include <math.h>
#include "rade_config.h"

#if defined(__CUDACC__)
#define RS_HD __host__ __device__ __forceinline__
#else
#define RS_HD inline
#endif

namespace rs {

template <typename T> struct X {
  static RS_HD T mul(T a, T b) { return a * b; }
  static RS_HD T add(T a, T b) { return a + b; }
  static RS_HD T sub(T a, T b) { return a - b; }
  static RS_HD T div(T a, T b) { return a / b; }
  static RS_HD T sqrt_(T a) { return sqrt(a); }
};
#if defined(__CUDA_ARCH__)
template <> struct X<float> {
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float sqrt_(float a) { return __fsqrt_rn(a); }
};
#endif

// Reciprocal / square root of the "ordinary arithmetic" sections (RaDe terms, the whole backward): exact on the host and
// in double, the SFU approximations (1-2 ulp) for float on the device.  An IEEE fp32 division is ~12 instructions with a
// slow-path call; the projection VJP had 38 of them per element.
template <typename T> RS_HD T rcp_(T x) { return T(1) / x; }
template <typename T> RS_HD T sqrt_fast(T x) { return sqrt(x); }
template <typename T> RS_HD T rsqrt_fast(T x) { return T(1) / sqrt(x); }
#if defined(__CUDA_ARCH__)
template <> __device__ __forceinline__ float rcp_<float>(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
template <> __device__ __forceinline__ float sqrt_fast<float>(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
template <> __device__ __forceinline__ float rsqrt_fast<float>(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
#endif

template <typename T> struct Cam {
  T W[9];  // rotation, row-major (world -> camera)
  T t[3];
  T fx, fy, cx, cy;
};

template <typename T> struct ProjParams {
  T width, height;  // image size as floats
  T eps2d, near_plane, far_plane, radius_clip;
};

template <typename T> struct ProjOut {
  int rx, ry;
  T m2x, m2y, depth, ca, cb, cc, comp, ray_t, rp0, rp1, nx, ny, nz;
};

// every intermediate the backward needs
template <typename T> struct ProjCtx {
  T qw, qx, qy, qz, qinv;            // normalised quaternion, 1/|q|
  T R[9], M[9];                      // rotation, R*diag(s)
  T S[6];                            // Sigma (00,01,02,11,12,22)
  T x, y, z;                         // camera-space mean
  T V[6];                            // Sigma_c
  T rz, rz2, u, v, tx, ty;           // perspective terms (u,v clamped)
  bool u_in, v_in;                   // clamp inactive
  T J00, J02, J11, J12;
  T B00, B01, B02, B10, B11, B12;    // J*V
  T c00, c01, c11;                   // un-blurred 2-D covariance
  T det0, c00b, c11b, detraw, det, ratio;
  // RaDe
  T aw[3], bl[3], cw[3], n[3], nn, h[3], d, vbn, w[3], pl0, pl1, l2, l, fac, g0, g1, cn[3], cnn;
  T is2[3], inn, ivbn, il, il2, icnn;   // reciprocals of scale^2, nn, vbn, l, l2, cnn (shared with the backward)
  bool rade_ok, valid;
};

template <typename T>
RS_HD void project_core(const T* mean, const T* quat, const T* scale, const Cam<T>& cam,
                        const ProjParams<T>& pp, ProjCtx<T>& k, ProjOut<T>& o) {
  typedef X<T> E;
  // ---- A1: quaternion -> rotation -> covariance (exact-order section)
  T w = quat[0], x = quat[1], y = quat[2], z = quat[3];
  T n2 = E::add(E::add(E::add(E::mul(w, w), E::mul(x, x)), E::mul(y, y)), E::mul(z, z));
  T inv = E::div(T(1), E::sqrt_(n2));
  w = E::mul(w, inv); x = E::mul(x, inv); y = E::mul(y, inv); z = E::mul(z, inv);
  k.qw = w; k.qx = x; k.qy = y; k.qz = z; k.qinv = inv;
  T x2 = E::mul(x, x), y2 = E::mul(y, y), z2 = E::mul(z, z);
  T xy = E::mul(x, y), xz = E::mul(x, z), yz = E::mul(y, z);
  T wx = E::mul(w, x), wy = E::mul(w, y), wz = E::mul(w, z);
  T* R = k.R;
  R[0] = E::sub(T(1), E::mul(T(2), E::add(y2, z2)));
  R[1] = E::mul(T(2), E::sub(xy, wz));
  R[2] = E::mul(T(2), E::add(xz, wy));
  R[3] = E::mul(T(2), E::add(xy, wz));
  R[4] = E::sub(T(1), E::mul(T(2), E::add(x2, z2)));
  R[5] = E::mul(T(2), E::sub(yz, wx));
  R[6] = E::mul(T(2), E::sub(xz, wy));
  R[7] = E::mul(T(2), E::add(yz, wx));
  R[8] = E::sub(T(1), E::mul(T(2), E::add(x2, y2)));
  T* M = k.M;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) M[i * 3 + j] = E::mul(R[i * 3 + j], scale[j]);
#define RS_DOT3(a0, b0, a1, b1, a2, b2) E::add(E::add(E::mul(a0, b0), E::mul(a1, b1)), E::mul(a2, b2))
  T* S = k.S;
  S[0] = RS_DOT3(M[0], M[0], M[1], M[1], M[2], M[2]);
  S[1] = RS_DOT3(M[0], M[3], M[1], M[4], M[2], M[5]);
  S[2] = RS_DOT3(M[0], M[6], M[1], M[7], M[2], M[8]);
  S[3] = RS_DOT3(M[3], M[3], M[4], M[4], M[5], M[5]);
  S[4] = RS_DOT3(M[3], M[6], M[4], M[7], M[5], M[8]);
  S[5] = RS_DOT3(M[6], M[6], M[7], M[7], M[8], M[8]);
  // ---- A2: world -> camera
  const T* Wm = cam.W;
  k.x = E::add(RS_DOT3(Wm[0], mean[0], Wm[1], mean[1], Wm[2], mean[2]), cam.t[0]);
  k.y = E::add(RS_DOT3(Wm[3], mean[0], Wm[4], mean[1], Wm[5], mean[2]), cam.t[1]);
  k.z = E::add(RS_DOT3(Wm[6], mean[0], Wm[7], mean[1], Wm[8], mean[2]), cam.t[2]);
  T A[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    A[i * 3 + 0] = RS_DOT3(Wm[i * 3], S[0], Wm[i * 3 + 1], S[1], Wm[i * 3 + 2], S[2]);
    A[i * 3 + 1] = RS_DOT3(Wm[i * 3], S[1], Wm[i * 3 + 1], S[3], Wm[i * 3 + 2], S[4]);
    A[i * 3 + 2] = RS_DOT3(Wm[i * 3], S[2], Wm[i * 3 + 1], S[4], Wm[i * 3 + 2], S[5]);
  }
  T* V = k.V;
  V[0] = RS_DOT3(A[0], Wm[0], A[1], Wm[1], A[2], Wm[2]);
  V[1] = RS_DOT3(A[0], Wm[3], A[1], Wm[4], A[2], Wm[5]);
  V[2] = RS_DOT3(A[0], Wm[6], A[1], Wm[7], A[2], Wm[8]);
  V[3] = RS_DOT3(A[3], Wm[3], A[4], Wm[4], A[5], Wm[5]);
  V[4] = RS_DOT3(A[3], Wm[6], A[4], Wm[7], A[5], Wm[8]);
  V[5] = RS_DOT3(A[6], Wm[6], A[7], Wm[7], A[8], Wm[8]);
#undef RS_DOT3
  // ---- A3: perspective EWA
  const T fx = cam.fx, fy = cam.fy, cx = cam.cx, cy = cam.cy;
  T tanx = E::div(E::mul(T(0.5), pp.width), fx);
  T tany = E::div(E::mul(T(0.5), pp.height), fy);
  T pad = T(RS_FOV_PAD);
  T lim_xp = E::add(E::div(E::sub(pp.width, cx), fx), E::mul(pad, tanx));
  T lim_xn = E::add(E::div(cx, fx), E::mul(pad, tanx));
  T lim_yp = E::add(E::div(E::sub(pp.height, cy), fy), E::mul(pad, tany));
  T lim_yn = E::add(E::div(cy, fy), E::mul(pad, tany));
  T rz = E::div(T(1), k.z);
  T u0 = E::mul(k.x, rz), v0 = E::mul(k.y, rz);
  T u = fmin(lim_xp, fmax(-lim_xn, u0));
  T v = fmin(lim_yp, fmax(-lim_yn, v0));
  k.u_in = (u0 >= -lim_xn) && (u0 <= lim_xp);
  k.v_in = (v0 >= -lim_yn) && (v0 <= lim_yp);
  k.rz = rz; k.u = u; k.v = v;
  k.tx = E::mul(k.z, u);
  k.ty = E::mul(k.z, v);
  k.rz2 = E::mul(rz, rz);
  k.J00 = E::mul(fx, rz);
  k.J02 = -E::mul(E::mul(fx, k.tx), k.rz2);
  k.J11 = E::mul(fy, rz);
  k.J12 = -E::mul(E::mul(fy, k.ty), k.rz2);
  k.B00 = E::add(E::mul(k.J00, V[0]), E::mul(k.J02, V[2]));
  k.B01 = E::add(E::mul(k.J00, V[1]), E::mul(k.J02, V[4]));
  k.B02 = E::add(E::mul(k.J00, V[2]), E::mul(k.J02, V[5]));
  k.B10 = E::add(E::mul(k.J11, V[1]), E::mul(k.J12, V[2]));
  k.B11 = E::add(E::mul(k.J11, V[3]), E::mul(k.J12, V[4]));
  k.B12 = E::add(E::mul(k.J11, V[4]), E::mul(k.J12, V[5]));
  k.c00 = E::add(E::mul(k.B00, k.J00), E::mul(k.B02, k.J02));
  k.c01 = E::add(E::mul(k.B01, k.J11), E::mul(k.B02, k.J12));
  k.c11 = E::add(E::mul(k.B11, k.J11), E::mul(k.B12, k.J12));
  o.m2x = E::add(E::mul(E::mul(fx, k.x), rz), cx);
  o.m2y = E::add(E::mul(E::mul(fy, k.y), rz), cy);
  o.depth = k.z;
  // ---- A4: blur, conic, radius, cull
  k.det0 = E::sub(E::mul(k.c00, k.c11), E::mul(k.c01, k.c01));
  k.c00b = E::add(k.c00, pp.eps2d);
  k.c11b = E::add(k.c11, pp.eps2d);
  k.detraw = E::sub(E::mul(k.c00b, k.c11b), E::mul(k.c01, k.c01));
  k.det = fmax(k.detraw, T(RS_DET_MIN));
  k.ratio = E::div(k.det0, k.det);
  o.comp = E::sqrt_(fmax(k.ratio, T(0)));
  o.ca = E::div(k.c11b, k.det);
  o.cb = E::div(-k.c01, k.det);
  o.cc = E::div(k.c00b, k.det);
  T rxf = ceil(E::mul(T(RS_RADIUS_SIGMA), E::sqrt_(k.c00b)));
  T ryf = ceil(E::mul(T(RS_RADIUS_SIGMA), E::sqrt_(k.c11b)));
  bool valid = (k.z > pp.near_plane) && (k.z < pp.far_plane);
  valid = valid && !((rxf <= pp.radius_clip) && (ryf <= pp.radius_clip));
  valid = valid && !((E::add(o.m2x, rxf) <= T(0)) || (E::sub(o.m2x, rxf) >= pp.width) ||
                     (E::add(o.m2y, ryf) <= T(0)) || (E::sub(o.m2y, ryf) >= pp.height));
  valid = valid && isfinite(rxf) && isfinite(ryf) && (rxf < T(1e9)) && (ryf < T(1e9));
  k.valid = valid;
  o.rx = valid ? (int)rxf : 0;
  o.ry = valid ? (int)ryf : 0;
  // ---- A5: RaDe ray-space plane and camera-space normal (ordinary arithmetic)
  const T* Rm = k.R;
#pragma unroll
  for (int j = 0; j < 3; ++j) k.aw[j] = Wm[j] * u + Wm[3 + j] * v + Wm[6 + j];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    k.is2[q] = rcp_(scale[q] * scale[q]);
    k.bl[q] = (Rm[q] * k.aw[0] + Rm[3 + q] * k.aw[1] + Rm[6 + q] * k.aw[2]) * k.is2[q];
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) k.cw[j] = Rm[j * 3] * k.bl[0] + Rm[j * 3 + 1] * k.bl[1] + Rm[j * 3 + 2] * k.bl[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) k.n[i] = Wm[i * 3] * k.cw[0] + Wm[i * 3 + 1] * k.cw[1] + Wm[i * 3 + 2] * k.cw[2];
  k.nn = sqrt_fast(k.n[0] * k.n[0] + k.n[1] * k.n[1] + k.n[2] * k.n[2]);
  bool ok = isfinite(k.nn) && (k.nn > T(0));
  T nns = ok ? k.nn : T(1);
  k.inn = rcp_(nns);
#pragma unroll
  for (int i = 0; i < 3; ++i) k.h[i] = k.n[i] * k.inn;
  k.d = k.h[0] * u + k.h[1] * v + k.h[2];
  k.vbn = fmax(k.d, T(RS_VBN_EPS));
  k.ivbn = rcp_(k.vbn);
#pragma unroll
  for (int i = 0; i < 3; ++i) k.w[i] = k.h[i] * k.ivbn;
  T uv = u * v;
  k.pl0 = (v * v + T(1)) * k.w[0] - uv * k.w[1] - u * k.w[2];
  k.pl1 = -uv * k.w[0] + (u * u + T(1)) * k.w[1] - v * k.w[2];
  k.l2 = u * u + v * v + T(1);
  k.l = sqrt_fast(k.tx * k.tx + k.ty * k.ty + k.z * k.z);
  k.il = rcp_(k.l);
  k.il2 = rcp_(k.l2);
  k.fac = k.l * k.il2;
  k.g0 = k.pl0 * k.fac;
  k.g1 = k.pl1 * k.fac;
  k.cn[0] = -k.g0 * rz - k.tx * k.il;
  k.cn[1] = -k.g1 * rz - k.ty * k.il;
  k.cn[2] = (k.g0 * k.tx + k.g1 * k.ty) * k.rz2 - k.z * k.il;
  k.cnn = sqrt_fast(k.cn[0] * k.cn[0] + k.cn[1] * k.cn[1] + k.cn[2] * k.cn[2]);
  ok = ok && isfinite(k.cnn) && (k.cnn > T(0));
  k.rade_ok = ok;
  if (ok) {
    k.icnn = rcp_(k.cnn);
    o.ray_t = k.l;
    o.rp0 = k.g0 * rcp_(fx);
    o.rp1 = k.g1 * rcp_(fy);
    o.nx = k.cn[0] * k.icnn; o.ny = k.cn[1] * k.icnn; o.nz = k.cn[2] * k.icnn;
  } else {
    k.icnn = T(0);
    o.ray_t = o.rp0 = o.rp1 = o.nx = o.ny = o.nz = T(0);
  }
  if (!valid) {
    o.m2x = o.m2y = o.depth = o.ca = o.cb = o.cc = o.comp = T(0);
    o.ray_t = o.rp0 = o.rp1 = o.nx = o.ny = o.nz = T(0);
  }
}

template <typename T>
RS_HD void project_fwd_one(const T* mean, const T* quat, const T* scale, const Cam<T>& cam,
                           const ProjParams<T>& pp, ProjOut<T>& o) {
  ProjCtx<T> k;
  project_core(mean, quat, scale, cam, pp, k, o);
}

// incoming gradients for one (camera, Gaussian)
template <typename T> struct ProjGradIn {
  T v_m2x, v_m2y, v_depth, v_ca, v_cb, v_cc, v_comp, v_ray_t, v_rp0, v_rp1, v_nx, v_ny, v_nz;
};

// VJP of project_core.  Accumulates (+=) into v_mean[3], v_quat[4], v_scale[3]; if v_W != nullptr
// also accumulates the camera gradient v_W[9] (rotation, row-major) and v_t[3].
template <typename T>
RS_HD void project_bwd_one(const T* mean, const T* quat, const T* scale, const Cam<T>& cam,
                           const ProjParams<T>& pp, const ProjGradIn<T>& g, T* v_mean, T* v_quat,
                           T* v_scale, T* v_W, T* v_t) {
  ProjCtx<T> k;
  ProjOut<T> o;
  project_core(mean, quat, scale, cam, pp, k, o);
  if (!k.valid) return;
  const T* Wm = cam.W;
  const T* R = k.R;
  const T fx = cam.fx, fy = cam.fy;
  const T u = k.u, v = k.v, rz = k.rz, rz2 = k.rz2;
  T vR[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  T vs[3] = {0, 0, 0};
  T vWl[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  T v_u = 0, v_v = 0, v_tx = 0, v_ty = 0, v_z = g.v_depth, v_x = 0, v_y = 0, v_rz = 0, v_rz2 = 0;

  // ---------------- RaDe stage
  if (k.rade_ok) {
    T N0 = k.cn[0] * k.icnn, N1 = k.cn[1] * k.icnn, N2 = k.cn[2] * k.icnn;
    T nd = N0 * g.v_nx + N1 * g.v_ny + N2 * g.v_nz;
    T vcn0 = (g.v_nx - N0 * nd) * k.icnn, vcn1 = (g.v_ny - N1 * nd) * k.icnn, vcn2 = (g.v_nz - N2 * nd) * k.icnn;
    T il = k.il;
    T v_g0 = g.v_rp0 * rcp_(fx) - vcn0 * rz + vcn2 * k.tx * rz2;
    T v_g1 = g.v_rp1 * rcp_(fy) - vcn1 * rz + vcn2 * k.ty * rz2;
    v_rz += -k.g0 * vcn0 - k.g1 * vcn1;
    v_rz2 += (k.g0 * k.tx + k.g1 * k.ty) * vcn2;
    v_tx += -vcn0 * il + k.g0 * rz2 * vcn2;
    v_ty += -vcn1 * il + k.g1 * rz2 * vcn2;
    v_z += -vcn2 * il;
    T v_fac = k.pl0 * v_g0 + k.pl1 * v_g1;
    T v_l = (k.tx * vcn0 + k.ty * vcn1 + k.z * vcn2) * il * il + g.v_ray_t + v_fac * k.il2;
    T v_pl0 = k.fac * v_g0, v_pl1 = k.fac * v_g1;
    T v_l2 = -k.fac * k.il2 * v_fac;
    v_tx += k.tx * il * v_l;
    v_ty += k.ty * il * v_l;
    v_z += k.z * il * v_l;
    v_u += T(2) * u * v_l2;
    v_v += T(2) * v * v_l2;
    T uv = u * v;
    T vw0 = (v * v + T(1)) * v_pl0 - uv * v_pl1;
    T vw1 = -uv * v_pl0 + (u * u + T(1)) * v_pl1;
    T vw2 = -u * v_pl0 - v * v_pl1;
    v_u += (-v * k.w[1] - k.w[2]) * v_pl0 + (-v * k.w[0] + T(2) * u * k.w[1]) * v_pl1;
    v_v += (T(2) * v * k.w[0] - u * k.w[1]) * v_pl0 + (-u * k.w[0] - k.w[2]) * v_pl1;
    T ivbn = k.ivbn;
    T vh0 = vw0 * ivbn, vh1 = vw1 * ivbn, vh2 = vw2 * ivbn;
    T v_vbn = -(k.w[0] * vw0 + k.w[1] * vw1 + k.w[2] * vw2) * ivbn;
    T v_d = (k.d >= T(RS_VBN_EPS)) ? v_vbn : T(0);
    vh0 += v_d * u; vh1 += v_d * v; vh2 += v_d;
    v_u += v_d * k.h[0];
    v_v += v_d * k.h[1];
    T hd = k.h[0] * vh0 + k.h[1] * vh1 + k.h[2] * vh2;
    T vn[3] = {(vh0 - k.h[0] * hd) * k.inn, (vh1 - k.h[1] * hd) * k.inn, (vh2 - k.h[2] * hd) * k.inn};
    T vcw[3], vbl[3], ve[3], vaw[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) vcw[j] = Wm[j] * vn[0] + Wm[3 + j] * vn[1] + Wm[6 + j] * vn[2];
#pragma unroll
    for (int q = 0; q < 3; ++q) vbl[q] = R[q] * vcw[0] + R[3 + q] * vcw[1] + R[6 + q] * vcw[2];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int q = 0; q < 3; ++q) vR[j * 3 + q] += vcw[j] * k.bl[q];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      ve[q] = vbl[q] * k.is2[q];
      vs[q] += -T(2) * k.bl[q] * (scale[q] * k.is2[q]) * vbl[q];
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      vaw[j] = R[j * 3] * ve[0] + R[j * 3 + 1] * ve[1] + R[j * 3 + 2] * ve[2];
#pragma unroll
      for (int q = 0; q < 3; ++q) vR[j * 3 + q] += k.aw[j] * ve[q];
    }
    v_u += Wm[0] * vaw[0] + Wm[1] * vaw[1] + Wm[2] * vaw[2];
    v_v += Wm[3] * vaw[0] + Wm[4] * vaw[1] + Wm[5] * vaw[2];
    if (v_W) {
      // n = W cw ; aw = W^T r
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) vWl[i * 3 + j] += vn[i] * k.cw[j];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        vWl[j] += u * vaw[j];
        vWl[3 + j] += v * vaw[j];
        vWl[6 + j] += vaw[j];
      }
    }
  }

  // ---------------- blur / conic stage
  T idet = rcp_(k.det);
  T v_c11b = g.v_ca * idet, v_c01 = -g.v_cb * idet, v_c00b = g.v_cc * idet;
  T ca = k.c11b * idet, cb = -k.c01 * idet, cc = k.c00b * idet;
  T v_det = -(g.v_ca * ca + g.v_cb * cb + g.v_cc * cc) * idet;
  T v_det0 = 0;
  if (k.ratio > T(0) && g.v_comp != T(0)) {
    T v_ratio = g.v_comp * T(0.5) * rsqrt_fast(k.ratio);
    v_det0 = v_ratio * idet;
    v_det += -k.ratio * idet * v_ratio;
  }
  if (!(k.detraw >= T(RS_DET_MIN))) v_det = 0;
  v_c00b += k.c11b * v_det;
  v_c11b += k.c00b * v_det;
  v_c01 += -T(2) * k.c01 * v_det;
  T v_c00 = v_c00b + k.c11 * v_det0;
  T v_c11 = v_c11b + k.c00 * v_det0;
  v_c01 += -T(2) * k.c01 * v_det0;

  // ---------------- perspective stage:  c = J V J^T
  T G00 = v_c00, G01 = T(0.5) * v_c01, G11 = v_c11;
  // FV = J^T G J (full symmetric), J = [[J00,0,J02],[0,J11,J12]]
  T GJ00 = G00 * k.J00, GJ01 = G01 * k.J11, GJ02 = G00 * k.J02 + G01 * k.J12;
  T GJ10 = G01 * k.J00, GJ11 = G11 * k.J11, GJ12 = G01 * k.J02 + G11 * k.J12;
  T FV[6];
  FV[0] = k.J00 * GJ00;
  FV[1] = k.J00 * GJ01;
  FV[2] = k.J00 * GJ02;
  FV[3] = k.J11 * GJ11;
  FV[4] = k.J11 * GJ12;
  FV[5] = k.J02 * GJ02 + k.J12 * GJ12;
  T vJ00 = T(2) * (G00 * k.B00 + G01 * k.B10);
  T vJ02 = T(2) * (G00 * k.B02 + G01 * k.B12);
  T vJ11 = T(2) * (G01 * k.B01 + G11 * k.B11);
  T vJ12 = T(2) * (G01 * k.B02 + G11 * k.B12);
  v_rz += fx * vJ00 + fy * vJ11;
  v_tx += -fx * rz2 * vJ02;
  v_ty += -fy * rz2 * vJ12;
  v_rz2 += -fx * k.tx * vJ02 - fy * k.ty * vJ12;
  v_x += fx * rz * g.v_m2x;
  v_y += fy * rz * g.v_m2y;
  v_rz += fx * k.x * g.v_m2x + fy * k.y * g.v_m2y;
  v_z += u * v_tx + v * v_ty;
  v_u += k.z * v_tx;
  v_v += k.z * v_ty;
  if (k.u_in) { v_x += rz * v_u; v_rz += k.x * v_u; }
  if (k.v_in) { v_y += rz * v_v; v_rz += k.y * v_v; }
  v_rz += T(2) * rz * v_rz2;
  v_z += -rz2 * v_rz;

  // ---------------- world -> camera stage
  v_mean[0] += Wm[0] * v_x + Wm[3] * v_y + Wm[6] * v_z;
  v_mean[1] += Wm[1] * v_x + Wm[4] * v_y + Wm[7] * v_z;
  v_mean[2] += Wm[2] * v_x + Wm[5] * v_y + Wm[8] * v_z;
  // FS = W^T FV W  (full symmetric)
  T FVf[9] = {FV[0], FV[1], FV[2], FV[1], FV[3], FV[4], FV[2], FV[4], FV[5]};
  T P[9];  // P = FV * W
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      P[i * 3 + j] = FVf[i * 3] * Wm[j] + FVf[i * 3 + 1] * Wm[3 + j] + FVf[i * 3 + 2] * Wm[6 + j];
  T FS[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) FS[i * 3 + j] = Wm[i] * P[j] + Wm[3 + i] * P[3 + j] + Wm[6 + i] * P[6 + j];
  if (v_W) {
    // p = W mean + t ;  V = W S W^T  ->  v_W += v_p mean^T + 2 FV W S
    T Sf[9] = {k.S[0], k.S[1], k.S[2], k.S[1], k.S[3], k.S[4], k.S[2], k.S[4], k.S[5]};
    T vp[3] = {v_x, v_y, v_z};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        T acc = vp[i] * mean[j];
#pragma unroll
        for (int q = 0; q < 3; ++q) acc += T(2) * P[i * 3 + q] * Sf[q * 3 + j];
        vWl[i * 3 + j] += acc;
      }
      v_t[i] += vp[i];
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) v_W[i] += vWl[i];
  }
  // ---------------- covariance stage: S = M M^T, M = R diag(s)
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      T vM = T(2) * (FS[i * 3] * k.M[j] + FS[i * 3 + 1] * k.M[3 + j] + FS[i * 3 + 2] * k.M[6 + j]);
      vR[i * 3 + j] += vM * scale[j];
      vs[j] += vM * R[i * 3 + j];
    }
  v_scale[0] += vs[0]; v_scale[1] += vs[1]; v_scale[2] += vs[2];
  // ---------------- quaternion stage
  T qw = k.qw, qx = k.qx, qy = k.qy, qz = k.qz;
  T vqw = T(2) * (-qz * vR[1] + qy * vR[2] + qz * vR[3] - qx * vR[5] - qy * vR[6] + qx * vR[7]);
  T vqx = T(2) * (qy * vR[1] + qz * vR[2] + qy * vR[3] - T(2) * qx * vR[4] - qw * vR[5] + qz * vR[6] + qw * vR[7] -
                  T(2) * qx * vR[8]);
  T vqy = T(2) * (-T(2) * qy * vR[0] + qx * vR[1] + qw * vR[2] + qx * vR[3] + qz * vR[5] - qw * vR[6] + qz * vR[7] -
                  T(2) * qy * vR[8]);
  T vqz = T(2) * (-T(2) * qz * vR[0] - qw * vR[1] + qx * vR[2] + qw * vR[3] - T(2) * qz * vR[4] + qy * vR[5] +
                  qx * vR[6] + qy * vR[7]);
  T qd = qw * vqw + qx * vqx + qy * vqy + qz * vqz;
  v_quat[0] += (vqw - qw * qd) * k.qinv;
  v_quat[1] += (vqx - qx * qd) * k.qinv;
  v_quat[2] += (vqy - qy * qd) * k.qinv;
  v_quat[3] += (vqz - qz * qd) * k.qinv;
}

// ------------------------------------------------------------------------------------------------ SH
#define RS_SH_C0 0.28209479177387814
#define RS_SH_C1 0.4886025119029199

// basis b[0..K) of the normalised direction (x,y,z), gsplat ordering
template <typename T> RS_HD void sh_basis(int deg, T x, T y, T z, T* b) {
  b[0] = T(RS_SH_C0);
  if (deg < 1) return;
  b[1] = -T(RS_SH_C1) * y; b[2] = T(RS_SH_C1) * z; b[3] = -T(RS_SH_C1) * x;
  if (deg < 2) return;
  T xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
  b[4] = T(1.0925484305920792) * xy;
  b[5] = T(-1.0925484305920792) * yz;
  b[6] = T(0.31539156525252005) * (T(2) * zz - xx - yy);
  b[7] = T(-1.0925484305920792) * xz;
  b[8] = T(0.5462742152960396) * (xx - yy);
  if (deg < 3) return;
  b[9] = T(-0.5900435899266435) * y * (T(3) * xx - yy);
  b[10] = T(2.890611442640554) * xy * z;
  b[11] = T(-0.4570457994644658) * y * (T(4) * zz - xx - yy);
  b[12] = T(0.3731763325901154) * z * (T(2) * zz - T(3) * xx - T(3) * yy);
  b[13] = T(-0.4570457994644658) * x * (T(4) * zz - xx - yy);
  b[14] = T(1.445305721320277) * z * (xx - yy);
  b[15] = T(-0.5900435899266435) * x * (xx - T(3) * yy);
}

// d(basis)/d(x,y,z) contracted with per-basis weights g[k] (= sum_c v_color[c]*coeff[k][c])
template <typename T> RS_HD void sh_basis_vjp(int deg, T x, T y, T z, const T* g, T& vx, T& vy, T& vz) {
  vx = vy = vz = T(0);
  if (deg < 1) return;
  vy += -T(RS_SH_C1) * g[1]; vz += T(RS_SH_C1) * g[2]; vx += -T(RS_SH_C1) * g[3];
  if (deg < 2) return;
  const T c4 = T(1.0925484305920792), c6 = T(0.31539156525252005), c8 = T(0.5462742152960396);
  vx += c4 * y * g[4];            vy += c4 * x * g[4];
  vy += -c4 * z * g[5];           vz += -c4 * y * g[5];
  vx += c6 * (-T(2) * x) * g[6];  vy += c6 * (-T(2) * y) * g[6];  vz += c6 * (T(4) * z) * g[6];
  vx += -c4 * z * g[7];           vz += -c4 * x * g[7];
  vx += c8 * T(2) * x * g[8];     vy += -c8 * T(2) * y * g[8];
  if (deg < 3) return;
  T xx = x * x, yy = y * y, zz = z * z, xy = x * y;
  const T d9 = T(-0.5900435899266435), d10 = T(2.890611442640554), d11 = T(-0.4570457994644658),
          d12 = T(0.3731763325901154), d14 = T(1.445305721320277);
  // b9 = d9*y*(3xx-yy)
  vx += d9 * y * T(6) * x * g[9];              vy += d9 * (T(3) * xx - T(3) * yy) * g[9];
  // b10 = d10*x*y*z
  vx += d10 * y * z * g[10];  vy += d10 * x * z * g[10];  vz += d10 * xy * g[10];
  // b11 = d11*y*(4zz-xx-yy)
  vx += d11 * y * (-T(2) * x) * g[11];  vy += d11 * (T(4) * zz - xx - T(3) * yy) * g[11];  vz += d11 * y * T(8) * z * g[11];
  // b12 = d12*z*(2zz-3xx-3yy)
  vx += d12 * z * (-T(6) * x) * g[12];  vy += d12 * z * (-T(6) * y) * g[12];  vz += d12 * (T(6) * zz - T(3) * xx - T(3) * yy) * g[12];
  // b13 = d11*x*(4zz-xx-yy)
  vx += d11 * (T(4) * zz - T(3) * xx - yy) * g[13];  vy += d11 * x * (-T(2) * y) * g[13];  vz += d11 * x * T(8) * z * g[13];
  // b14 = d14*z*(xx-yy)
  vx += d14 * z * T(2) * x * g[14];  vy += -d14 * z * T(2) * y * g[14];  vz += d14 * (xx - yy) * g[14];
  // b15 = d9*x*(xx-3yy)
  vx += d9 * (T(3) * xx - T(3) * yy) * g[15];  vy += d9 * x * (-T(6) * y) * g[15];
}

}  // namespace rs
