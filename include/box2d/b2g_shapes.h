// -----------------------------------------------------------------------------------------------
// Third-party notice.  To stay source- and result-compatible with box2d-optimized, parts of this
// file restate declarations, inline math and creation-time algorithms of that library (itself a
// fork of Box2D).  Those parts are covered by the MIT License:
//   Copyright (c) 2019 Erin Catto, Copyright (c) 2020 Manolis Tsamis
// The full licence text and permission notice are in LICENSES/box2d-optimized-MIT.txt.
// -----------------------------------------------------------------------------------------------
// b2g_shapes.h — shapes, AABBs and manifold types of the drop-in C++ API.
//
// API mirror of include/box2d/b2_shape.h:33-106, b2_circle_shape.h, b2_edge_shape.h,
// b2_polygon_shape.h and the data types of b2_collision.h:44-200 in the reference.  Shapes are
// host-side value objects: a fixture clones its shape and serialises it into the float4 shape
// pool consumed by the device (layout in include/b2cuda.h).  The reference's b2BlockAllocator
// parameter of Clone() is gone because fixtures own their shapes with ordinary new/delete.
#ifndef B2G_SHAPES_H
#define B2G_SHAPES_H

#include <climits>
#include "b2g_types.h"

struct b2MassData {
  float mass;
  b2Vec2 center;
  float I;
};

struct b2AABB {
  bool IsValid() const {
    return upperBound.x >= lowerBound.x && upperBound.y >= lowerBound.y && lowerBound.IsValid() && upperBound.IsValid();
  }
  b2Vec2 GetCenter() const { return 0.5f * (lowerBound + upperBound); }
  b2Vec2 GetExtents() const { return 0.5f * (upperBound - lowerBound); }
  float GetPerimeter() const {
    float wx = upperBound.x - lowerBound.x;
    float wy = upperBound.y - lowerBound.y;
    return 2.0f * (wx + wy);
  }
  void Combine(const b2AABB& aabb) {
    lowerBound = b2Min(lowerBound, aabb.lowerBound);
    upperBound = b2Max(upperBound, aabb.upperBound);
  }
  void Combine(const b2AABB& a, const b2AABB& b) {
    lowerBound = b2Min(a.lowerBound, b.lowerBound);
    upperBound = b2Max(a.upperBound, b.upperBound);
  }
  bool Contains(const b2AABB& aabb) const {
    return lowerBound.x <= aabb.lowerBound.x && lowerBound.y <= aabb.lowerBound.y &&
           aabb.upperBound.x <= upperBound.x && aabb.upperBound.y <= upperBound.y;
  }
  b2Vec2 lowerBound;
  b2Vec2 upperBound;
};

inline bool b2TestOverlap(const b2AABB& a, const b2AABB& b) {
  return (a.upperBound.x >= b.lowerBound.x) & (a.lowerBound.x <= b.upperBound.x) &
         (a.upperBound.y >= b.lowerBound.y) & (a.lowerBound.y <= b.upperBound.y);
}

/// b2_collision.h:147-160
struct b2RayCastInput {
  b2Vec2 p1, p2;
  float maxFraction;
};
struct b2RayCastOutput {
  b2Vec2 normal;
  float fraction;
};

class b2Shape {
 public:
  enum Type { e_circle = 0, e_edge = 1, e_polygon = 2, e_chain = 3, e_typeCount = 4 };
  virtual ~b2Shape() {}
  virtual b2Shape* Clone() const = 0;
  Type GetType() const { return m_type; }
  virtual bool TestPoint(const b2Transform& xf, const b2Vec2& p) const = 0;
  /// single-shape ray cast on the host (b2_shape.h:90-93); b2World::RayCast runs batched on the device
  virtual bool RayCast(b2RayCastOutput* output, const b2RayCastInput& input, const b2Transform& xf) const = 0;
  virtual void ComputeAABB(b2AABB* aabb, const b2Transform& xf) const = 0;
  virtual void ComputeMass(b2MassData* massData, float density) const = 0;
  /// number of float4 records this shape occupies in the device shape pool
  virtual int32 DeviceQuadCount() const = 0;
  /// serialise into the device shape pool layout (include/b2cuda.h)
  virtual void WriteDeviceQuads(float* quads) const = 0;
  Type m_type;
  float m_radius;
};

class b2CircleShape : public b2Shape {
 public:
  b2CircleShape() {
    m_type = e_circle;
    m_radius = 0.0f;
    m_p.SetZero();
  }
  b2Shape* Clone() const override { return new b2CircleShape(*this); }
  bool TestPoint(const b2Transform& xf, const b2Vec2& p) const override;
  bool RayCast(b2RayCastOutput* output, const b2RayCastInput& input, const b2Transform& xf) const override;
  void ComputeAABB(b2AABB* aabb, const b2Transform& xf) const override;
  void ComputeMass(b2MassData* massData, float density) const override;
  int32 DeviceQuadCount() const override { return 1; }
  void WriteDeviceQuads(float* quads) const override;
  b2Vec2 m_p;
};

class b2EdgeShape : public b2Shape {
 public:
  b2EdgeShape() {
    m_type = e_edge;
    m_radius = b2_polygonRadius;
    m_vertex0.SetZero();
    m_vertex1.SetZero();
    m_vertex2.SetZero();
    m_vertex3.SetZero();
    m_oneSided = false;
  }
  void SetOneSided(const b2Vec2& v0, const b2Vec2& v1, const b2Vec2& v2, const b2Vec2& v3);
  void SetTwoSided(const b2Vec2& v1, const b2Vec2& v2);
  b2Shape* Clone() const override { return new b2EdgeShape(*this); }
  bool TestPoint(const b2Transform& xf, const b2Vec2& p) const override;
  bool RayCast(b2RayCastOutput* output, const b2RayCastInput& input, const b2Transform& xf) const override;
  void ComputeAABB(b2AABB* aabb, const b2Transform& xf) const override;
  void ComputeMass(b2MassData* massData, float density) const override;
  int32 DeviceQuadCount() const override { return 3; }
  void WriteDeviceQuads(float* quads) const override;
  b2Vec2 m_vertex1, m_vertex2;
  b2Vec2 m_vertex0, m_vertex3;
  bool m_oneSided;
};

class b2PolygonShape : public b2Shape {
 public:
  b2PolygonShape() {
    m_type = e_polygon;
    m_radius = b2_polygonRadius;
    m_count = 0;
    m_centroid.SetZero();
  }
  b2Shape* Clone() const override { return new b2PolygonShape(*this); }
  void Set(const b2Vec2* points, int32 count);
  void SetAsBox(float hx, float hy);
  void SetAsBox(float hx, float hy, const b2Vec2& center, float angle);
  bool TestPoint(const b2Transform& xf, const b2Vec2& p) const override;
  bool RayCast(b2RayCastOutput* output, const b2RayCastInput& input, const b2Transform& xf) const override;
  void ComputeAABB(b2AABB* aabb, const b2Transform& xf) const override;
  void ComputeMass(b2MassData* massData, float density) const override;
  bool Validate() const;
  int32 DeviceQuadCount() const override { return 1 + m_count; }
  void WriteDeviceQuads(float* quads) const override;
  b2Vec2 m_centroid;
  b2Vec2 m_vertices[b2_maxPolygonVertices];
  b2Vec2 m_normals[b2_maxPolygonVertices];
  int32 m_count;
};

// ---- manifolds (b2_collision.h:44-200) ------------------------------------------------------
const uint8 b2_nullFeature = UCHAR_MAX;

struct b2ContactFeature {
  enum Type { e_vertex = 0, e_face = 1 };
  uint8 indexA;
  uint8 indexB;
  uint8 typeA;
  uint8 typeB;
};

union b2ContactID {
  b2ContactFeature cf;
  uint32 key;
};

struct b2ManifoldPoint {
  b2Vec2 localPoint;
  float normalImpulse;
  float tangentImpulse;
  b2ContactID id;
};

struct b2Manifold {
  enum Type { e_circles, e_faceA, e_faceB };
  b2ManifoldPoint points[b2_maxManifoldPoints];
  b2Vec2 localNormal;
  b2Vec2 localPoint;
  Type type;
  int32 pointCount;
};

struct b2WorldManifold {
  void Initialize(const b2Manifold* manifold, const b2Transform& xfA, float radiusA, const b2Transform& xfB,
                  float radiusB);
  b2Vec2 normal;
  b2Vec2 points[b2_maxManifoldPoints];
  float separations[b2_maxManifoldPoints];
};

enum b2PointState { b2_nullState, b2_addState, b2_persistState, b2_removeState };
void b2GetPointStates(b2PointState state1[b2_maxManifoldPoints], b2PointState state2[b2_maxManifoldPoints],
                      const b2Manifold* manifold1, const b2Manifold* manifold2);

#endif
