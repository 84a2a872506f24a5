// b2g_math.cuh — float32 2-D math for the device path.
//
// Restates the inlines of include/box2d/b2_math.h (b2Vec2 :117-129, b2Rot :313-318,
// b2Mul/b2MulT :597-640, b2Cross :416-433) with the SAME operation order, because the
// branchy thresholds of the narrowphase and the solver are only reproducible if every
// intermediate rounds the same way.  The library is compiled with -fmad=false so nvcc does
// not contract a*b+c into an FMA (the reference is built for baseline x86-64: no FMA).
#pragma once

// Programmatic dependent launch (sm_90+): every step kernel starts with this pair.  The kernels of
// a step are small and strictly ordered on one stream, so most of a step is launch latency and
// grid ramp-up/drain; letting grid N+1 be scheduled while grid N is still running hides that.
// launch_dependents lets the next grid's CTAs become resident; wait blocks them until every
// prerequisite grid has COMPLETED and its writes are visible, so nothing is read early.  Without
// the launch attribute (cudaLaunchAttributeProgrammaticStreamSerialization) both are no-ops.
#ifdef __CUDA_ARCH__
#define B2G_PDL_ENTER()                                     \
  do {                                                      \
    asm volatile("griddepcontrol.launch_dependents;");      \
    asm volatile("griddepcontrol.wait;" ::: "memory");      \
  } while (0)
#else
#define B2G_PDL_ENTER() do {} while (0)
#endif

#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

#define B2G_HD __host__ __device__ __forceinline__

// tuning constants, include/box2d/b2_common.h:110-182
#define B2G_EPSILON FLT_EPSILON
#define B2G_MAX_FLOAT FLT_MAX
#define B2G_PI 3.14159265359f
#define B2G_LINEAR_SLOP 0.005f
#define B2G_ANGULAR_SLOP (2.0f / 180.0f * B2G_PI)
#define B2G_POLYGON_RADIUS (2.0f * B2G_LINEAR_SLOP)
#define B2G_MAX_LINEAR_CORRECTION 0.2f
#define B2G_MAX_ANGULAR_CORRECTION (8.0f / 180.0f * B2G_PI)
#define B2G_MAX_TRANSLATION 2.0f
#define B2G_MAX_TRANSLATION_SQ (B2G_MAX_TRANSLATION * B2G_MAX_TRANSLATION)
#define B2G_MAX_ROTATION (0.5f * B2G_PI)
#define B2G_MAX_ROTATION_SQ (B2G_MAX_ROTATION * B2G_MAX_ROTATION)
#define B2G_BAUMGARTE 0.2f
#define B2G_TIME_TO_SLEEP 0.5f
#define B2G_LINEAR_SLEEP_TOL 0.01f
#define B2G_ANGULAR_SLEEP_TOL (2.0f / 180.0f * B2G_PI)
#define B2G_MAX_POLY_VERTS 8

struct Rot {
  float s, c;
};
struct Xf {
  float2 p;
  Rot q;
};

B2G_HD float2 v2(float x, float y) { return make_float2(x, y); }
B2G_HD float2 operator+(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
B2G_HD float2 operator-(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
B2G_HD float2 operator-(float2 a) { return make_float2(-a.x, -a.y); }
B2G_HD float2 operator*(float s, float2 a) { return make_float2(s * a.x, s * a.y); }
B2G_HD void operator+=(float2& a, float2 b) {
  a.x += b.x;
  a.y += b.y;
}
B2G_HD void operator-=(float2& a, float2 b) {
  a.x -= b.x;
  a.y -= b.y;
}
B2G_HD float dot2(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
B2G_HD float cross2(float2 a, float2 b) { return a.x * b.y - a.y * b.x; }
B2G_HD float2 cross_vs(float2 a, float s) { return make_float2(s * a.y, -s * a.x); }
B2G_HD float2 cross_sv(float s, float2 a) { return make_float2(-s * a.y, s * a.x); }
B2G_HD float dist_sq(float2 a, float2 b) {
  float2 c = a - b;
  return dot2(c, c);
}
B2G_HD float len2(float2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
// b2Vec2::Normalize: leaves the vector untouched when shorter than epsilon
B2G_HD float normalize2(float2& a) {
  float length = len2(a);
  if (length < B2G_EPSILON) return 0.0f;
  float inv = 1.0f / length;
  a.x *= inv;
  a.y *= inv;
  return length;
}
B2G_HD float2 rot_mul(Rot q, float2 v) { return make_float2(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
B2G_HD float2 rot_mulT(Rot q, float2 v) { return make_float2(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }
B2G_HD Rot rot_mulT(Rot q, Rot r) {
  Rot o;
  o.s = q.c * r.s - q.s * r.c;
  o.c = q.c * r.c + q.s * r.s;
  return o;
}
B2G_HD float2 xf_mul(Xf T, float2 v) {
  float x = (T.q.c * v.x - T.q.s * v.y) + T.p.x;
  float y = (T.q.s * v.x + T.q.c * v.y) + T.p.y;
  return make_float2(x, y);
}
B2G_HD float2 xf_mulT(Xf T, float2 v) {
  float px = v.x - T.p.x;
  float py = v.y - T.p.y;
  return make_float2(T.q.c * px + T.q.s * py, -T.q.s * px + T.q.c * py);
}
B2G_HD Xf xf_mulT(Xf A, Xf B) {
  Xf C;
  C.q = rot_mulT(A.q, B.q);
  C.p = rot_mulT(A.q, B.p - A.p);
  return C;
}
B2G_HD Xf xf_from4(float4 v) {
  Xf T;
  T.p = make_float2(v.x, v.y);
  T.q.s = v.z;
  T.q.c = v.w;
  return T;
}
B2G_HD float4 xf_to4(Xf T) { return make_float4(T.p.x, T.p.y, T.q.s, T.q.c); }
// b2Rot::Set (b2_math.h:313-318)
// The reference calls the host libm's sinf/cosf.  CUDA's float sinf/cosf differ from those by 1-2
// ulp, and even a correctly rounded result differs from glibc's in ~1 % of arguments; one ulp in
// a rotation is amplified by ill-conditioned resting contacts into 1e-4-level velocity differences
// and can flip the solver's branchy thresholds.  So the device evaluates the SAME function glibc
// (>= 2.28, x86-64 FMA variant; the published ARM optimized-routines sincosf) evaluates: argument
// reduction by pi/2 in double with a 2^24-prescaled quadrant, a degree-7 sine / degree-8 cosine
// minimax polynomial in double with fused multiply-adds, one rounding to float.  Checked here
// bit-for-bit against glibc 2.39 on 3e8 arguments (CPU restatement) and by tests/test_rotation_parity.py.
// |angle| >= 120 (glibc's large-argument path) falls back to double sincos rounded once.
B2G_HD Rot rot_set(float angle) {
  Rot q;
#ifdef __CUDA_ARCH__
  float ay = fabsf(angle);
  if (ay < 0x1p-12f) {
    q.s = angle;
    q.c = 1.0f;
    return q;
  }
  if (!(ay < 120.0f)) {
    double sd, cd;
    sincos((double)angle, &sd, &cd);
    q.s = (float)sd;
    q.c = (float)cd;
    return q;
  }
  double x = (double)angle;
  int n = 0;
  if (ay >= 0.75f) {  // glibc compares the top 12 bits with those of pi/4
    double r = __dmul_rn(x, 0x1.45F306DC9C883p+23);
    n = (__double2int_rz(r) + 0x800000) >> 24;
    x = __fma_rn(-(double)n, 0x1.921FB54442D18p0, x);
  }
  double cs = (n & 2) ? -1.0 : 1.0;  // second table: negated cosine polynomial
  double x2 = __dmul_rn(x, x);
  if (((n + 1) & 2)) x = -x;         // sign[n & 3] = {1, -1, -1, 1}
  double x4 = __dmul_rn(x2, x2), x3 = __dmul_rn(x2, x);
  double c2 = __fma_rn(x2, cs * 0x1.99343027bf8c3p-16, cs * -0x1.6c087e89a359dp-10);
  double s1 = __fma_rn(x2, -0x1.994eb3774cf24p-13, 0x1.1107605230bc4p-7);
  double c1 = __fma_rn(x2, cs * -0x1.ffffffd0c621cp-2, cs);
  double x5 = __dmul_rn(x3, x2), x6 = __dmul_rn(x4, x2);
  double s = __fma_rn(x3, -0x1.555545995a603p-3, x);
  double c = __fma_rn(x4, cs * 0x1.55553e1068f19p-5, c1);
  float sv = (float)__fma_rn(x5, s1, s), cv = (float)__fma_rn(x6, c2, c);
  q.s = (n & 1) ? cv : sv;
  q.c = (n & 1) ? sv : cv;
#else
  q.s = sinf(angle);
  q.c = cosf(angle);
#endif
  return q;
}
// transform of a body at (c, a) with local centre lc: b2Body::SynchronizeTransform (b2_body.h:957-961)
B2G_HD Xf xf_from_sweep(float2 c, float a, float2 lc) {
  Xf T;
  T.q = rot_set(a);
  T.p = c - rot_mul(T.q, lc);
  return T;
}
// b2Min / b2Max / b2Clamp (b2_math.h:646-670) as ternaries: identical NaN / signed-zero behaviour
B2G_HD float minf_(float a, float b) { return a < b ? a : b; }
B2G_HD float maxf_(float a, float b) { return a > b ? a : b; }
B2G_HD float clampf(float a, float lo, float hi) { return maxf_(lo, minf_(a, hi)); }
B2G_HD float absf_(float a) { return a > 0.0f ? a : -a; }
