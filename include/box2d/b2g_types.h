// -----------------------------------------------------------------------------------------------
// Third-party notice.  To stay source- and result-compatible with box2d-optimized, parts of this
// file restate declarations, inline math and creation-time algorithms of that library (itself a
// fork of Box2D).  Those parts are covered by the MIT License:
//   Copyright (c) 2019 Erin Catto, Copyright (c) 2020 Manolis Tsamis
// The full licence text and permission notice are in LICENSES/box2d-optimized-MIT.txt.
// -----------------------------------------------------------------------------------------------
// b2g_types.h — scalar types, tuning constants and 2-D math of the drop-in C++ API.
//
// API mirror of the reference's include/box2d/b2_types.h, b2_common.h:110-182, b2_settings.h:40-80
// and b2_math.h (b2Vec2 :41-135, b2Rot :287-345, b2Transform :347-380, b2Sweep :382-409,
// free functions :410-735).  Names, argument order and arithmetic order are the reference's so
// user code and the parity tests compile against either implementation unchanged.
#ifndef B2G_TYPES_H
#define B2G_TYPES_H

#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cassert>

typedef int8_t int8;
typedef int16_t int16;
typedef int32_t int32;
typedef int64_t int64;
typedef uint8_t uint8;
typedef uint16_t uint16;
typedef uint32_t uint32;
typedef float float32;
typedef double float64;

#define B2_API
#define B2_NOT_USED(x) ((void)(x))
#define b2Assert(A) assert(A)

#define b2_maxFloat FLT_MAX
#define b2_epsilon FLT_EPSILON
#define b2_pi 3.14159265359f

#define b2_lengthUnitsPerMeter 1.0f
#define b2_maxPolygonVertices 8
#define b2_maxManifoldPoints 2
#define b2_linearSlop (0.005f * b2_lengthUnitsPerMeter)
#define b2_angularSlop (2.0f / 180.0f * b2_pi)
#define b2_polygonRadius (2.0f * b2_linearSlop)
#define b2_maxLinearCorrection (0.2f * b2_lengthUnitsPerMeter)
#define b2_maxAngularCorrection (8.0f / 180.0f * b2_pi)
#define b2_maxTranslation (2.0f * b2_lengthUnitsPerMeter)
#define b2_maxTranslationSquared (b2_maxTranslation * b2_maxTranslation)
#define b2_maxRotation (0.5f * b2_pi)
#define b2_maxRotationSquared (b2_maxRotation * b2_maxRotation)
#define b2_baumgarte 0.2f
#define b2_timeToSleep 0.5f
#define b2_linearSleepTolerance (0.01f * b2_lengthUnitsPerMeter)
#define b2_angularSleepTolerance (2.0f / 180.0f * b2_pi)

struct b2BodyUserData {
  b2BodyUserData() : pointer(0) {}
  uintptr_t pointer;
};
struct b2FixtureUserData {
  b2FixtureUserData() : pointer(0) {}
  uintptr_t pointer;
};
struct b2JointUserData {
  b2JointUserData() : pointer(0) {}
  uintptr_t pointer;
};

inline bool b2IsValid(float x) { return std::isfinite(x); }
#define b2Sqrt(x) sqrtf(x)
#define b2Atan2(y, x) atan2f(y, x)

struct b2Vec2 {
  b2Vec2() {}
  b2Vec2(float xIn, float yIn) : x(xIn), y(yIn) {}
  void SetZero() { x = 0.0f; y = 0.0f; }
  void Set(float x_, float y_) { x = x_; y = y_; }
  b2Vec2 operator-() const { return b2Vec2(-x, -y); }
  float operator()(int32 i) const { return (&x)[i]; }
  float& operator()(int32 i) { return (&x)[i]; }
  void operator+=(const b2Vec2& v) { x += v.x; y += v.y; }
  void operator-=(const b2Vec2& v) { x -= v.x; y -= v.y; }
  void operator*=(float a) { x *= a; y *= a; }
  void operator/=(float a) { x /= a; y /= a; }
  float Length() const { return b2Sqrt(x * x + y * y); }
  float LengthSquared() const { return x * x + y * y; }
  float Normalize() {
    float length = Length();
    if (length < b2_epsilon) return 0.0f;
    float invLength = 1.0f / length;
    x *= invLength;
    y *= invLength;
    return length;
  }
  bool IsValid() const { return b2IsValid(x) && b2IsValid(y); }
  b2Vec2 Skew() const { return b2Vec2(-y, x); }
  float x, y;
};

struct b2Vec3 {
  b2Vec3() {}
  b2Vec3(float xIn, float yIn, float zIn) : x(xIn), y(yIn), z(zIn) {}
  void SetZero() { x = y = z = 0.0f; }
  void Set(float x_, float y_, float z_) { x = x_; y = y_; z = z_; }
  float x, y, z;
};

struct b2Mat22 {
  b2Mat22() {}
  b2Mat22(const b2Vec2& c1, const b2Vec2& c2) : ex(c1), ey(c2) {}
  b2Mat22(float a11, float a12, float a21, float a22) {
    ex.x = a11; ex.y = a21;
    ey.x = a12; ey.y = a22;
  }
  void Set(const b2Vec2& c1, const b2Vec2& c2) { ex = c1; ey = c2; }
  void SetIdentity() { ex.x = 1.0f; ey.x = 0.0f; ex.y = 0.0f; ey.y = 1.0f; }
  void SetZero() { ex.SetZero(); ey.SetZero(); }
  b2Mat22 GetInverse() const {
    float a = ex.x, b = ey.x, c = ex.y, d = ey.y;
    b2Mat22 B;
    float det = a * d - b * c;
    if (det != 0.0f) det = 1.0f / det;
    B.ex.x = det * d;  B.ey.x = -det * b;
    B.ex.y = -det * c; B.ey.y = det * a;
    return B;
  }
  b2Vec2 ex, ey;
};

struct b2Rot {
  b2Rot() {}
  explicit b2Rot(float angle) { s = sinf(angle); c = cosf(angle); }
  void Set(float angle) { s = sinf(angle); c = cosf(angle); }
  void SetIdentity() { s = 0.0f; c = 1.0f; }
  float GetAngle() const { return b2Atan2(s, c); }
  b2Vec2 GetXAxis() const { return b2Vec2(c, s); }
  b2Vec2 GetYAxis() const { return b2Vec2(-s, c); }
  float s, c;
};

struct b2Transform {
  b2Transform() {}
  b2Transform(const b2Vec2& position, const b2Rot& rotation) : p(position), q(rotation) {}
  void SetIdentity() { p.SetZero(); q.SetIdentity(); }
  void Set(const b2Vec2& position, float angle) { p = position; q.Set(angle); }
  b2Vec2 p;
  b2Rot q;
};

extern const b2Vec2 b2Vec2_zero;

inline float b2Dot(const b2Vec2& a, const b2Vec2& b) { return a.x * b.x + a.y * b.y; }
inline float b2Cross(const b2Vec2& a, const b2Vec2& b) { return a.x * b.y - a.y * b.x; }
inline b2Vec2 b2Cross(const b2Vec2& a, float s) { return b2Vec2(s * a.y, -s * a.x); }
inline b2Vec2 b2Cross(float s, const b2Vec2& a) { return b2Vec2(-s * a.y, s * a.x); }
inline b2Vec2 b2Mul(const b2Mat22& A, const b2Vec2& v) {
  return b2Vec2(A.ex.x * v.x + A.ey.x * v.y, A.ex.y * v.x + A.ey.y * v.y);
}
inline b2Vec2 b2MulT(const b2Mat22& A, const b2Vec2& v) { return b2Vec2(b2Dot(v, A.ex), b2Dot(v, A.ey)); }
inline b2Vec2 operator+(const b2Vec2& a, const b2Vec2& b) { return b2Vec2(a.x + b.x, a.y + b.y); }
inline b2Vec2 operator-(const b2Vec2& a, const b2Vec2& b) { return b2Vec2(a.x - b.x, a.y - b.y); }
inline b2Vec2 operator*(float s, const b2Vec2& a) { return b2Vec2(s * a.x, s * a.y); }
inline bool operator==(const b2Vec2& a, const b2Vec2& b) { return a.x == b.x && a.y == b.y; }
inline bool operator!=(const b2Vec2& a, const b2Vec2& b) { return a.x != b.x || a.y != b.y; }
inline float b2Distance(const b2Vec2& a, const b2Vec2& b) { b2Vec2 c = a - b; return c.Length(); }
inline float b2DistanceSquared(const b2Vec2& a, const b2Vec2& b) { b2Vec2 c = a - b; return b2Dot(c, c); }
inline b2Rot b2Mul(const b2Rot& q, const b2Rot& r) {
  b2Rot qr;
  qr.s = q.s * r.c + q.c * r.s;
  qr.c = q.c * r.c - q.s * r.s;
  return qr;
}
inline b2Rot b2MulT(const b2Rot& q, const b2Rot& r) {
  b2Rot qr;
  qr.s = q.c * r.s - q.s * r.c;
  qr.c = q.c * r.c + q.s * r.s;
  return qr;
}
inline b2Vec2 b2Mul(const b2Rot& q, const b2Vec2& v) { return b2Vec2(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
inline b2Vec2 b2MulT(const b2Rot& q, const b2Vec2& v) { return b2Vec2(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }
inline b2Vec2 b2Mul(const b2Transform& T, const b2Vec2& v) {
  float x = (T.q.c * v.x - T.q.s * v.y) + T.p.x;
  float y = (T.q.s * v.x + T.q.c * v.y) + T.p.y;
  return b2Vec2(x, y);
}
inline b2Vec2 b2MulT(const b2Transform& T, const b2Vec2& v) {
  float px = v.x - T.p.x;
  float py = v.y - T.p.y;
  return b2Vec2(T.q.c * px + T.q.s * py, -T.q.s * px + T.q.c * py);
}
inline b2Transform b2Mul(const b2Transform& A, const b2Transform& B) {
  b2Transform C;
  C.q = b2Mul(A.q, B.q);
  C.p = b2Mul(A.q, B.p) + A.p;
  return C;
}
inline b2Transform b2MulT(const b2Transform& A, const b2Transform& B) {
  b2Transform C;
  C.q = b2MulT(A.q, B.q);
  C.p = b2MulT(A.q, B.p - A.p);
  return C;
}
template <typename T> inline T b2Abs(T a) { return a > T(0) ? a : -a; }
inline b2Vec2 b2Abs(const b2Vec2& a) { return b2Vec2(b2Abs(a.x), b2Abs(a.y)); }
template <typename T> inline T b2Min(T a, T b) { return a < b ? a : b; }
inline b2Vec2 b2Min(const b2Vec2& a, const b2Vec2& b) { return b2Vec2(b2Min(a.x, b.x), b2Min(a.y, b.y)); }
template <typename T> inline T b2Max(T a, T b) { return a > b ? a : b; }
inline b2Vec2 b2Max(const b2Vec2& a, const b2Vec2& b) { return b2Vec2(b2Max(a.x, b.x), b2Max(a.y, b.y)); }
template <typename T> inline T b2Clamp(T a, T low, T high) { return b2Max(low, b2Min(a, high)); }
template <typename T> inline void b2Swap(T& a, T& b) { T tmp = a; a = b; b = tmp; }

/// Motion of a body's centre of mass over a step (b2_math.h:382-409, :697-733)
struct b2Sweep {
  b2Sweep() {}
  void GetTransform(b2Transform* xf, float beta) const {
    xf->p = (1.0f - beta) * c0 + beta * c;
    float angle = (1.0f - beta) * a0 + beta * a;
    xf->q.Set(angle);
    xf->p -= b2Mul(xf->q, localCenter);
  }
  void Advance(float alpha) {
    float beta = (alpha - alpha0) / (1.0f - alpha0);
    c0 += beta * (c - c0);
    a0 += beta * (a - a0);
    alpha0 = alpha;
  }
  void Normalize() {
    float twoPi = 2.0f * b2_pi;
    float d = twoPi * floorf(a0 / twoPi);
    a0 -= d;
    a -= d;
  }
  b2Vec2 localCenter;
  b2Vec2 c0, c;
  float a0, a;
  float alpha0;
};

#endif
