// -----------------------------------------------------------------------------------------------
// Third-party notice.  To stay source- and result-compatible with box2d-optimized, parts of this
// file restate declarations, inline math and creation-time algorithms of that library (itself a
// fork of Box2D).  Those parts are covered by the MIT License:
//   Copyright (c) 2019 Erin Catto, Copyright (c) 2020 Manolis Tsamis
// The full licence text and permission notice are in LICENSES/box2d-optimized-MIT.txt.
// -----------------------------------------------------------------------------------------------
// b2g_world.h — b2World / b2Body / b2Fixture / b2Contact / listeners of the drop-in C++ API.
//
// API mirror of the reference's include/box2d/b2_world.h:49-253, b2_body.h:45-449,
// b2_fixture.h:37-227, b2_contact.h:65-143, b2_world_callbacks.h:46-210, b2_joint.h,
// b2_revolute_joint.h and b2_time_step.h:29-36.  Same class names, method names, argument
// meaning and error behaviour (silent no-op / nullptr while the world is locked), but the
// objects are thin host handles: all simulation state lives in a device arena
// (include/b2cuda.h) and b2World::Step launches the CUDA step.  Host copies of body state
// are refreshed lazily after a step, on the first getter that needs them.
#ifndef B2G_WORLD_H
#define B2G_WORLD_H

#include <vector>
#include "b2g_shapes.h"

class b2World;
class b2Body;
class b2Fixture;
class b2Contact;
class b2Joint;
struct b2WorldImpl;
struct b2WorldBatchImpl;

enum b2BodyType { b2_staticBody = 0, b2_kinematicBody, b2_dynamicBody };

struct b2BodyDef {
  b2BodyDef() {
    position.Set(0.0f, 0.0f);
    angle = 0.0f;
    linearVelocity.Set(0.0f, 0.0f);
    angularVelocity = 0.0f;
    linearDamping = 0.0f;
    angularDamping = 0.0f;
    allowSleep = true;
    awake = true;
    fixedRotation = false;
    bullet = false;
    type = b2_staticBody;
    enabled = true;
    gravityScale = 1.0f;
  }
  b2BodyType type;
  b2Vec2 position;
  float angle;
  b2Vec2 linearVelocity;
  float angularVelocity;
  float linearDamping;
  float angularDamping;
  bool allowSleep;
  bool awake;
  bool fixedRotation;
  bool bullet;
  bool enabled;
  b2BodyUserData userData;
  float gravityScale;
};

struct b2Filter {
  b2Filter() {
    categoryBits = 0x0001;
    maskBits = 0xFFFF;
    groupIndex = 0;
  }
  uint16 categoryBits;
  uint16 maskBits;
  int16 groupIndex;
};

struct b2FixtureDef {
  b2FixtureDef() {
    shape = nullptr;
    friction = 0.2f;
    restitution = 0.0f;
    restitutionThreshold = 1.0f * b2_lengthUnitsPerMeter;
    density = 0.0f;
    isSensor = false;
  }
  const b2Shape* shape;
  b2FixtureUserData userData;
  float friction;
  float restitution;
  float restitutionThreshold;
  float density;
  bool isSensor;
  b2Filter filter;
};

class b2Fixture {
 public:
  b2Shape::Type GetType() const { return m_shape->GetType(); }
  b2Shape* GetShape() { return m_shape; }
  const b2Shape* GetShape() const { return m_shape; }
  void SetSensor(bool sensor);
  bool IsSensor() const { return m_isSensor; }
  void SetFilterData(const b2Filter& filter);
  const b2Filter& GetFilterData() const { return m_filter; }
  void Refilter();
  b2Body* GetBody() { return m_body; }
  const b2Body* GetBody() const { return m_body; }
  b2Fixture* GetNext() { return m_next; }
  const b2Fixture* GetNext() const { return m_next; }
  b2FixtureUserData& GetUserData() { return m_userData; }
  uint32 GetId() { return m_id; }
  bool TestPoint(const b2Vec2& p) const;
  bool RayCast(b2RayCastOutput* output, const b2RayCastInput& input) const;
  void GetMassData(b2MassData* massData) const { m_shape->ComputeMass(massData, m_density); }
  void SetDensity(float density) { m_density = density; }
  float GetDensity() const { return m_density; }
  float GetFriction() const { return m_friction; }
  void SetFriction(float friction);
  float GetRestitution() const { return m_restitution; }
  void SetRestitution(float restitution);
  float GetRestitutionThreshold() const { return m_restitutionThreshold; }
  void SetRestitutionThreshold(float threshold);
  void UpdateAABB();
  const b2AABB& GetAABB() const;

 private:
  friend class b2Body;
  friend class b2World;
  friend class b2Contact;
  friend struct b2WorldImpl;
  friend struct b2WorldBatchImpl;
  b2Fixture() {}
  float m_density;
  b2Fixture* m_next;
  b2Body* m_body;
  b2Shape* m_shape;
  float m_friction;
  float m_restitution;
  float m_restitutionThreshold;
  b2Filter m_filter;
  bool m_isSensor;
  b2FixtureUserData m_userData;
  mutable b2AABB m_aabb;
  uint32 m_id;
  int32 m_index;     // index in the device fixture arrays
  int32 m_shapeOff;  // offset in the device shape pool
};

struct b2JointEdge {
  b2Body* other;
  b2Joint* joint;
  b2JointEdge* prev;
  b2JointEdge* next;
};

class b2Body {
 public:
  b2Fixture* CreateFixture(const b2FixtureDef* def);
  b2Fixture* CreateFixture(const b2Shape* shape, float density);
  void DestroyFixture(b2Fixture* fixture);
  void SetTransform(const b2Vec2& position, float angle);
  const b2Transform& GetTransform() const;
  const b2Vec2& GetPosition() const;
  float GetAngle() const;
  const b2Vec2& GetWorldCenter() const;
  const b2Vec2& GetLocalCenter() const;
  void SetLinearVelocity(const b2Vec2& v);
  const b2Vec2& GetLinearVelocity() const;
  void SetAngularVelocity(float omega);
  float GetAngularVelocity() const;
  void ApplyForce(const b2Vec2& force, const b2Vec2& point) { ApplyForce(force, point, true); }
  void ApplyForceToCenter(const b2Vec2& force) { ApplyForceToCenter(force, true); }
  void ApplyTorque(float torque) { ApplyTorque(torque, true); }
  void ApplyLinearImpulse(const b2Vec2& impulse, const b2Vec2& point) { ApplyLinearImpulse(impulse, point, true); }
  void ApplyLinearImpulseToCenter(const b2Vec2& impulse) { ApplyLinearImpulseToCenter(impulse, true); }
  void ApplyAngularImpulse(float impulse) { ApplyAngularImpulse(impulse, true); }
  void ApplyForce(const b2Vec2& force, const b2Vec2& point, bool wake);
  void ApplyForceToCenter(const b2Vec2& force, bool wake);
  void ApplyTorque(float torque, bool wake);
  void ApplyLinearImpulse(const b2Vec2& impulse, const b2Vec2& point, bool wake);
  void ApplyLinearImpulseToCenter(const b2Vec2& impulse, bool wake);
  void ApplyAngularImpulse(float impulse, bool wake);
  float GetMass() const { return m_mass; }
  float GetInertia() const;
  void GetMassData(b2MassData* data) const;
  void SetMassData(const b2MassData* data);
  void ResetMassData();
  b2Vec2 GetWorldPoint(const b2Vec2& localPoint) const { return b2Mul(GetTransform(), localPoint); }
  b2Vec2 GetWorldVector(const b2Vec2& localVector) const { return b2Mul(GetTransform().q, localVector); }
  b2Vec2 GetLocalPoint(const b2Vec2& worldPoint) const { return b2MulT(GetTransform(), worldPoint); }
  b2Vec2 GetLocalVector(const b2Vec2& worldVector) const { return b2MulT(GetTransform().q, worldVector); }
  b2Vec2 GetLinearVelocityFromWorldPoint(const b2Vec2& worldPoint) const;
  b2Vec2 GetLinearVelocityFromLocalPoint(const b2Vec2& localPoint) const {
    return GetLinearVelocityFromWorldPoint(GetWorldPoint(localPoint));
  }
  float GetLinearDamping() const { return m_linearDamping; }
  void SetLinearDamping(float linearDamping);
  float GetAngularDamping() const { return m_angularDamping; }
  void SetAngularDamping(float angularDamping);
  float GetGravityScale() const { return m_gravityScale; }
  void SetGravityScale(float scale);
  void SetType(b2BodyType type);
  b2BodyType GetType() const { return m_type; }
  void SetBullet(bool flag);
  bool IsBullet() const;
  void SetSleepingAllowed(bool flag);
  bool IsSleepingAllowed() const;
  void SetAwake(bool flag);
  bool IsAwake() const;
  void SetEnabled(bool flag);
  bool IsEnabled() const;
  void SetFixedRotation(bool flag);
  bool IsFixedRotation() const;
  b2Fixture* GetFixtureList() { return m_fixtureList; }
  const b2Fixture* GetFixtureList() const { return m_fixtureList; }
  b2JointEdge* GetJointList() { return m_jointList; }
  const b2JointEdge* GetJointList() const { return m_jointList; }
  int32 GetContactCount();
  b2Contact* GetContact(int32 idx);
  b2Body* GetNext() { return m_next; }
  const b2Body* GetNext() const { return m_next; }
  b2BodyUserData& GetUserData() { return m_userData; }
  b2World* GetWorld() { return m_world; }
  const b2World* GetWorld() const { return m_world; }
  void UpdateAABBs();

 private:
  friend class b2World;
  friend class b2Fixture;
  friend class b2Contact;
  friend class b2Joint;
  friend class b2RevoluteJoint;
  friend struct b2WorldImpl;
  friend struct b2WorldBatchImpl;
  b2Body(const b2BodyDef* bd, b2World* world);
  ~b2Body() {}
  void SyncIn() const;  // make the host copy current
  void Touch();         // host copy changed: upload before the next step
  void SynchronizeTransform() {
    m_xf.q.Set(m_sweep.a);
    m_xf.p = m_sweep.c - b2Mul(m_xf.q, m_sweep.localCenter);
  }
  bool ShouldCollide(const b2Body* other) const;

  b2BodyType m_type;
  uint16 m_flags;
  int32 m_index;  // index in the device body arrays (creation order; indices of destroyed bodies are reused)
  bool m_onDevice;  // the device holds a copy of this body (false until its first upload)
  mutable b2Transform m_xf;
  mutable b2Sweep m_sweep;
  mutable b2Vec2 m_linearVelocity;
  mutable float m_angularVelocity;
  mutable b2Vec2 m_force;
  mutable float m_torque;
  b2World* m_world;
  b2Body* m_prev;
  b2Body* m_next;
  b2Fixture* m_fixtureList;
  int32 m_fixtureCount;
  b2JointEdge* m_jointList;
  std::vector<b2Contact*> m_contacts;
  float m_mass, m_invMass;
  float m_I, m_invI;
  float m_linearDamping;
  float m_angularDamping;
  float m_gravityScale;
  mutable float m_sleepTime;
  b2BodyUserData m_userData;

 public:
  enum {
    e_islandFlag = 0x0001,
    e_awakeFlag = 0x0002,
    e_autoSleepFlag = 0x0004,
    e_bulletFlag = 0x0008,
    e_fixedRotationFlag = 0x0010,
    e_enabledFlag = 0x0020,
    e_toiFlag = 0x0040
  };
};

// ---- contacts -------------------------------------------------------------------------------
inline float b2MixFriction(float friction1, float friction2) { return b2Sqrt(friction1 * friction2); }
inline float b2MixRestitution(float restitution1, float restitution2) {
  return restitution1 > restitution2 ? restitution1 : restitution2;
}
inline float b2MixRestitutionThreshold(float threshold1, float threshold2) {
  return threshold1 < threshold2 ? threshold1 : threshold2;
}

class b2Contact {
 public:
  b2Manifold* GetManifold() { return &m_manifold; }
  const b2Manifold* GetManifold() const { return &m_manifold; }
  void GetWorldManifold(b2WorldManifold* worldManifold) const;
  bool IsTouching() const { return (m_flags & e_touchingFlag) == e_touchingFlag; }
  void SetEnabled(bool flag);
  bool IsEnabled() const { return (m_flags & e_enabledFlag) == e_enabledFlag; }
  b2Contact* GetNext() { return m_next; }
  const b2Contact* GetNext() const { return m_next; }
  b2Fixture* GetFixtureA() { return m_fixtureA; }
  const b2Fixture* GetFixtureA() const { return m_fixtureA; }
  b2Fixture* GetFixtureB() { return m_fixtureB; }
  const b2Fixture* GetFixtureB() const { return m_fixtureB; }
  void SetFriction(float friction);
  float GetFriction() const { return m_friction; }
  void ResetFriction();
  /// b2_contact.h:246-249: offer this contact to the contact filter again at the next step
  void FlagForFiltering();
  void SetRestitution(float restitution);
  float GetRestitution() const { return m_restitution; }
  void ResetRestitution();
  void SetRestitutionThreshold(float threshold);
  float GetRestitutionThreshold() const { return m_restitutionThreshold; }
  void ResetRestitutionThreshold();
  void SetTangentSpeed(float speed);
  float GetTangentSpeed() const { return m_tangentSpeed; }

  enum {
    e_islandFlag = 0x0001,
    e_persistFlag = 0x0002,
    e_filterFlag = 0x0004,
    e_touchingFlag = 0x0008,
    e_enabledFlag = 0x0010,
    e_bulletHitFlag = 0x0020,
  };

 private:
  friend class b2World;
  friend class b2Body;
  friend struct b2WorldImpl;
  friend struct b2WorldBatchImpl;
  b2Contact() {}
  uint32 m_flags;
  b2Contact* m_prev;
  b2Contact* m_next;
  b2Fixture* m_fixtureA;
  b2Fixture* m_fixtureB;
  b2Manifold m_manifold;
  float m_friction;
  float m_restitution;
  float m_restitutionThreshold;
  float m_tangentSpeed;
  int32 m_deviceIndex;  // index in the device contact arrays after the last step
  bool m_overridden;    // a setter was called: push flags/material back before the solve
  b2World* m_world;
};

// ---- joints ---------------------------------------------------------------------------------
enum b2JointType {
  e_unknownJoint,
  e_revoluteJoint,
  e_prismaticJoint,
  e_distanceJoint,
  e_pulleyJoint,
  e_mouseJoint,
  e_gearJoint,
  e_wheelJoint,
  e_weldJoint,
  e_frictionJoint,
  e_ropeJoint,
  e_motorJoint
};

struct b2JointDef {
  b2JointDef() {
    type = e_unknownJoint;
    bodyA = nullptr;
    bodyB = nullptr;
    collideConnected = false;
  }
  b2JointType type;
  b2JointUserData userData;
  b2Body* bodyA;
  b2Body* bodyB;
  bool collideConnected;
};

struct b2RevoluteJointDef : public b2JointDef {
  b2RevoluteJointDef() {
    type = e_revoluteJoint;
    localAnchorA.Set(0.0f, 0.0f);
    localAnchorB.Set(0.0f, 0.0f);
    referenceAngle = 0.0f;
    lowerAngle = 0.0f;
    upperAngle = 0.0f;
    maxMotorTorque = 0.0f;
    motorSpeed = 0.0f;
    enableLimit = false;
    enableMotor = false;
  }
  void Initialize(b2Body* bodyA, b2Body* bodyB, const b2Vec2& anchor);
  b2Vec2 localAnchorA;
  b2Vec2 localAnchorB;
  float referenceAngle;
  bool enableLimit;
  float lowerAngle;
  float upperAngle;
  bool enableMotor;
  float motorSpeed;
  float maxMotorTorque;
};

class b2Joint {
 public:
  b2JointType GetType() const { return m_type; }
  b2Body* GetBodyA() { return m_bodyA; }
  b2Body* GetBodyB() { return m_bodyB; }
  virtual b2Vec2 GetAnchorA() const = 0;
  virtual b2Vec2 GetAnchorB() const = 0;
  virtual b2Vec2 GetReactionForce(float inv_dt) const = 0;
  virtual float GetReactionTorque(float inv_dt) const = 0;
  b2Joint* GetNext() { return m_next; }
  const b2Joint* GetNext() const { return m_next; }
  b2JointUserData& GetUserData() { return m_userData; }
  bool GetCollideConnected() const { return m_collideConnected; }
  virtual ~b2Joint() {}

 protected:
  friend class b2World;
  friend class b2Body;
  friend struct b2WorldImpl;
  friend struct b2WorldBatchImpl;
  b2Joint(const b2JointDef* def);
  /// device record of this joint (include/b2cuda.h b2gJointArrays): anchors[4], params[12], state[5]
  virtual void WriteDevice(float* anchors, float* params, float* state) const = 0;
  /// accumulated impulses read back from the device
  virtual void ReadDeviceState(const float* state) = 0;
  void Touch(bool wake = true);  // pull the accumulators, (wake both bodies,) mark the joint table dirty
  b2JointType m_type;
  b2Joint* m_prev;
  b2Joint* m_next;
  b2JointEdge m_edgeA;
  b2JointEdge m_edgeB;
  b2Body* m_bodyA;
  b2Body* m_bodyB;
  int32 m_index;
  bool m_collideConnected;
  b2JointUserData m_userData;
};

class b2RevoluteJoint : public b2Joint {
 public:
  b2Vec2 GetAnchorA() const override;
  b2Vec2 GetAnchorB() const override;
  const b2Vec2& GetLocalAnchorA() const { return m_localAnchorA; }
  const b2Vec2& GetLocalAnchorB() const { return m_localAnchorB; }
  float GetReferenceAngle() const { return m_referenceAngle; }
  float GetJointAngle() const;
  float GetJointSpeed() const;
  bool IsLimitEnabled() const { return m_enableLimit; }
  bool IsMotorEnabled() const { return m_enableMotor; }
  float GetMotorSpeed() const { return m_motorSpeed; }
  float GetMaxMotorTorque() const { return m_maxMotorTorque; }
  float GetLowerLimit() const { return m_lowerAngle; }
  float GetUpperLimit() const { return m_upperAngle; }
  // b2_revolute_joint.cpp:333-342, 364-447: reactions read the impulses the device accumulated in
  // the last step; setters wake both bodies and take effect at the next Step
  b2Vec2 GetReactionForce(float inv_dt) const override;
  float GetReactionTorque(float inv_dt) const override;
  float GetMotorTorque(float inv_dt) const;
  void EnableLimit(bool flag);
  void SetLimits(float lower, float upper);
  void EnableMotor(bool flag);
  void SetMotorSpeed(float speed);
  void SetMaxMotorTorque(float torque);

 protected:
  friend class b2World;
  friend struct b2WorldImpl;
  friend struct b2WorldBatchImpl;
  b2RevoluteJoint(const b2RevoluteJointDef* def);
  b2Vec2 m_localAnchorA;
  b2Vec2 m_localAnchorB;
  float m_referenceAngle;
  bool m_enableLimit;
  float m_lowerAngle;
  float m_upperAngle;
  bool m_enableMotor;
  float m_motorSpeed;
  float m_maxMotorTorque;
  // host copies of the warm-start accumulators, refreshed lazily from the device
  mutable b2Vec2 m_impulse;
  mutable float m_motorImpulse;
  mutable float m_lowerImpulse;
  mutable float m_upperImpulse;
  void WriteDevice(float* anchors, float* params, float* state) const override;
  void ReadDeviceState(const float* state) override;
};

/// b2_joint.h:76-84
void b2LinearStiffness(float& stiffness, float& damping, float frequencyHertz, float dampingRatio, const b2Body* bodyA,
                       const b2Body* bodyB);
void b2AngularStiffness(float& stiffness, float& damping, float frequencyHertz, float dampingRatio, const b2Body* bodyA,
                        const b2Body* bodyB);

/// b2_prismatic_joint.h:30-190: one translational degree of freedom along an axis fixed in bodyA, with an
/// optional translation limit and linear motor
struct b2PrismaticJointDef : public b2JointDef {
  b2PrismaticJointDef() {
    type = e_prismaticJoint;
    localAnchorA.Set(0.0f, 0.0f);
    localAnchorB.Set(0.0f, 0.0f);
    localAxisA.Set(1.0f, 0.0f);
    referenceAngle = 0.0f;
    enableLimit = false;
    lowerTranslation = 0.0f;
    upperTranslation = 0.0f;
    enableMotor = false;
    maxMotorForce = 0.0f;
    motorSpeed = 0.0f;
  }
  void Initialize(b2Body* bodyA, b2Body* bodyB, const b2Vec2& anchor, const b2Vec2& axis);
  b2Vec2 localAnchorA;
  b2Vec2 localAnchorB;
  b2Vec2 localAxisA;
  float referenceAngle;
  bool enableLimit;
  float lowerTranslation;
  float upperTranslation;
  bool enableMotor;
  float maxMotorForce;
  float motorSpeed;
};

class b2PrismaticJoint : public b2Joint {
 public:
  b2Vec2 GetAnchorA() const override;
  b2Vec2 GetAnchorB() const override;
  // b2_prismatic_joint.cpp:463-471: the reference uses the axis / perpendicular of the last
  // InitVelocityConstraints; here they are taken from bodyA's present transform
  b2Vec2 GetReactionForce(float inv_dt) const override;
  float GetReactionTorque(float inv_dt) const override;
  const b2Vec2& GetLocalAnchorA() const { return m_localAnchorA; }
  const b2Vec2& GetLocalAnchorB() const { return m_localAnchorB; }
  const b2Vec2& GetLocalAxisA() const { return m_localXAxisA; }
  float GetReferenceAngle() const { return m_referenceAngle; }
  float GetJointTranslation() const;
  float GetJointSpeed() const;
  bool IsLimitEnabled() const { return m_enableLimit; }
  void EnableLimit(bool flag);
  float GetLowerLimit() const { return m_lowerTranslation; }
  float GetUpperLimit() const { return m_upperTranslation; }
  void SetLimits(float lower, float upper);
  bool IsMotorEnabled() const { return m_enableMotor; }
  void EnableMotor(bool flag);
  void SetMotorSpeed(float speed);
  float GetMotorSpeed() const { return m_motorSpeed; }
  void SetMaxMotorForce(float force);
  float GetMaxMotorForce() const { return m_maxMotorForce; }
  float GetMotorForce(float inv_dt) const;

 protected:
  friend class b2World;
  friend struct b2WorldImpl;
  friend struct b2WorldBatchImpl;
  b2PrismaticJoint(const b2PrismaticJointDef* def);
  void WriteDevice(float* anchors, float* params, float* state) const override;
  void ReadDeviceState(const float* state) override;
  b2Vec2 m_localAnchorA;
  b2Vec2 m_localAnchorB;
  b2Vec2 m_localXAxisA;
  b2Vec2 m_localYAxisA;
  float m_referenceAngle;
  float m_lowerTranslation;
  float m_upperTranslation;
  float m_maxMotorForce;
  float m_motorSpeed;
  bool m_enableLimit;
  bool m_enableMotor;
  mutable b2Vec2 m_impulse;
  mutable float m_motorImpulse;
  mutable float m_lowerImpulse;
  mutable float m_upperImpulse;
};

/// b2_mouse_joint.h:30-130: drags a point of bodyB towards a world target with a soft, force-limited constraint
struct b2MouseJointDef : public b2JointDef {
  b2MouseJointDef() {
    type = e_mouseJoint;
    target.Set(0.0f, 0.0f);
    maxForce = 0.0f;
    stiffness = 0.0f;
    damping = 0.0f;
  }
  b2Vec2 target;
  float maxForce;
  float stiffness;
  float damping;
};

class b2MouseJoint : public b2Joint {
 public:
  b2Vec2 GetAnchorA() const override;
  b2Vec2 GetAnchorB() const override;
  b2Vec2 GetReactionForce(float inv_dt) const override;
  float GetReactionTorque(float inv_dt) const override;
  void SetTarget(const b2Vec2& target);
  const b2Vec2& GetTarget() const { return m_targetA; }
  void SetMaxForce(float force);
  float GetMaxForce() const { return m_maxForce; }
  void SetStiffness(float stiffness);
  float GetStiffness() const { return m_stiffness; }
  void SetDamping(float damping);
  float GetDamping() const { return m_damping; }
  const b2Vec2& GetLocalAnchorB() const { return m_localAnchorB; }  // (not in the reference; used by the scene shim)

 protected:
  friend class b2World;
  friend struct b2WorldImpl;
  friend struct b2WorldBatchImpl;
  b2MouseJoint(const b2MouseJointDef* def);
  void WriteDevice(float* anchors, float* params, float* state) const override;
  void ReadDeviceState(const float* state) override;
  b2Vec2 m_localAnchorB;
  b2Vec2 m_targetA;
  float m_maxForce;
  float m_stiffness;
  float m_damping;
  mutable b2Vec2 m_impulse;
};

/// b2_friction_joint.h:30-118: top-down friction between two bodies (clamped linear and angular drag)
struct b2FrictionJointDef : public b2JointDef {
  b2FrictionJointDef() {
    type = e_frictionJoint;
    localAnchorA.Set(0.0f, 0.0f);
    localAnchorB.Set(0.0f, 0.0f);
    maxForce = 0.0f;
    maxTorque = 0.0f;
  }
  void Initialize(b2Body* bodyA, b2Body* bodyB, const b2Vec2& anchor);
  b2Vec2 localAnchorA;
  b2Vec2 localAnchorB;
  float maxForce;
  float maxTorque;
};

class b2FrictionJoint : public b2Joint {
 public:
  b2Vec2 GetAnchorA() const override;
  b2Vec2 GetAnchorB() const override;
  b2Vec2 GetReactionForce(float inv_dt) const override;
  float GetReactionTorque(float inv_dt) const override;
  const b2Vec2& GetLocalAnchorA() const { return m_localAnchorA; }
  const b2Vec2& GetLocalAnchorB() const { return m_localAnchorB; }
  void SetMaxForce(float force);
  float GetMaxForce() const { return m_maxForce; }
  void SetMaxTorque(float torque);
  float GetMaxTorque() const { return m_maxTorque; }

 protected:
  friend class b2World;
  friend struct b2WorldImpl;
  friend struct b2WorldBatchImpl;
  b2FrictionJoint(const b2FrictionJointDef* def);
  void WriteDevice(float* anchors, float* params, float* state) const override;
  void ReadDeviceState(const float* state) override;
  b2Vec2 m_localAnchorA;
  b2Vec2 m_localAnchorB;
  float m_maxForce;
  float m_maxTorque;
  mutable b2Vec2 m_linearImpulse;
  mutable float m_angularImpulse;
};

/// b2_motor_joint.h:30-133: drives bodyB to a linear / angular offset from bodyA with bounded force
struct b2MotorJointDef : public b2JointDef {
  b2MotorJointDef() {
    type = e_motorJoint;
    linearOffset.Set(0.0f, 0.0f);
    angularOffset = 0.0f;
    maxForce = 1.0f;
    maxTorque = 1.0f;
    correctionFactor = 0.3f;
  }
  void Initialize(b2Body* bodyA, b2Body* bodyB);
  b2Vec2 linearOffset;
  float angularOffset;
  float maxForce;
  float maxTorque;
  float correctionFactor;
};

class b2MotorJoint : public b2Joint {
 public:
  b2Vec2 GetAnchorA() const override;
  b2Vec2 GetAnchorB() const override;
  b2Vec2 GetReactionForce(float inv_dt) const override;
  float GetReactionTorque(float inv_dt) const override;
  void SetLinearOffset(const b2Vec2& linearOffset);
  const b2Vec2& GetLinearOffset() const { return m_linearOffset; }
  void SetAngularOffset(float angularOffset);
  float GetAngularOffset() const { return m_angularOffset; }
  void SetMaxForce(float force);
  float GetMaxForce() const { return m_maxForce; }
  void SetMaxTorque(float torque);
  float GetMaxTorque() const { return m_maxTorque; }
  void SetCorrectionFactor(float factor);
  float GetCorrectionFactor() const { return m_correctionFactor; }

 protected:
  friend class b2World;
  friend struct b2WorldImpl;
  friend struct b2WorldBatchImpl;
  b2MotorJoint(const b2MotorJointDef* def);
  void WriteDevice(float* anchors, float* params, float* state) const override;
  void ReadDeviceState(const float* state) override;
  b2Vec2 m_linearOffset;
  float m_angularOffset;
  float m_maxForce;
  float m_maxTorque;
  float m_correctionFactor;
  mutable b2Vec2 m_linearImpulse;
  mutable float m_angularImpulse;
};

/// b2_wheel_joint.h:30-231: a point of bodyB rides on a line fixed in bodyA (suspension spring along the
/// line, optional translation limits) and rotates freely, optionally driven by a motor
struct b2WheelJointDef : public b2JointDef {
  b2WheelJointDef() {
    type = e_wheelJoint;
    localAnchorA.Set(0.0f, 0.0f);
    localAnchorB.Set(0.0f, 0.0f);
    localAxisA.Set(1.0f, 0.0f);
    enableLimit = false;
    lowerTranslation = 0.0f;
    upperTranslation = 0.0f;
    enableMotor = false;
    maxMotorTorque = 0.0f;
    motorSpeed = 0.0f;
    stiffness = 0.0f;
    damping = 0.0f;
  }
  void Initialize(b2Body* bodyA, b2Body* bodyB, const b2Vec2& anchor, const b2Vec2& axis);
  b2Vec2 localAnchorA;
  b2Vec2 localAnchorB;
  b2Vec2 localAxisA;
  bool enableLimit;
  float lowerTranslation;
  float upperTranslation;
  bool enableMotor;
  float maxMotorTorque;
  float motorSpeed;
  float stiffness;
  float damping;
};

class b2WheelJoint : public b2Joint {
 public:
  b2Vec2 GetAnchorA() const override;
  b2Vec2 GetAnchorB() const override;
  // b2_wheel_joint.cpp:459-467: the reference uses the axes of the last InitVelocityConstraints; here
  // they are taken from bodyA's present transform
  b2Vec2 GetReactionForce(float inv_dt) const override;
  float GetReactionTorque(float inv_dt) const override;
  const b2Vec2& GetLocalAnchorA() const { return m_localAnchorA; }
  const b2Vec2& GetLocalAnchorB() const { return m_localAnchorB; }
  const b2Vec2& GetLocalAxisA() const { return m_localXAxisA; }
  float GetJointTranslation() const;
  float GetJointLinearSpeed() const;
  float GetJointAngle() const;
  float GetJointAngularSpeed() const;
  bool IsLimitEnabled() const { return m_enableLimit; }
  void EnableLimit(bool flag);
  float GetLowerLimit() const { return m_lowerTranslation; }
  float GetUpperLimit() const { return m_upperTranslation; }
  void SetLimits(float lower, float upper);
  bool IsMotorEnabled() const { return m_enableMotor; }
  void EnableMotor(bool flag);
  void SetMotorSpeed(float speed);
  float GetMotorSpeed() const { return m_motorSpeed; }
  void SetMaxMotorTorque(float torque);
  float GetMaxMotorTorque() const { return m_maxMotorTorque; }
  float GetMotorTorque(float inv_dt) const;
  void SetStiffness(float stiffness);
  float GetStiffness() const { return m_stiffness; }
  void SetDamping(float damping);
  float GetDamping() const { return m_damping; }

 protected:
  friend class b2World;
  friend struct b2WorldImpl;
  friend struct b2WorldBatchImpl;
  b2WheelJoint(const b2WheelJointDef* def);
  void WriteDevice(float* anchors, float* params, float* state) const override;
  void ReadDeviceState(const float* state) override;
  b2Vec2 m_localAnchorA;
  b2Vec2 m_localAnchorB;
  b2Vec2 m_localXAxisA;
  b2Vec2 m_localYAxisA;
  float m_lowerTranslation;
  float m_upperTranslation;
  float m_maxMotorTorque;
  float m_motorSpeed;
  bool m_enableLimit;
  bool m_enableMotor;
  float m_stiffness;
  float m_damping;
  mutable float m_impulse;
  mutable float m_springImpulse;
  mutable float m_motorImpulse;
  mutable float m_lowerImpulse;
  mutable float m_upperImpulse;
};

/// b2_weld_joint.h:30-128: glues two bodies together (optionally with a rotational spring)
struct b2WeldJointDef : public b2JointDef {
  b2WeldJointDef() {
    type = e_weldJoint;
    localAnchorA.Set(0.0f, 0.0f);
    localAnchorB.Set(0.0f, 0.0f);
    referenceAngle = 0.0f;
    stiffness = 0.0f;
    damping = 0.0f;
  }
  void Initialize(b2Body* bodyA, b2Body* bodyB, const b2Vec2& anchor);
  b2Vec2 localAnchorA;
  b2Vec2 localAnchorB;
  float referenceAngle;
  float stiffness;
  float damping;
};

class b2WeldJoint : public b2Joint {
 public:
  b2Vec2 GetAnchorA() const override;
  b2Vec2 GetAnchorB() const override;
  b2Vec2 GetReactionForce(float inv_dt) const override;
  float GetReactionTorque(float inv_dt) const override;
  const b2Vec2& GetLocalAnchorA() const { return m_localAnchorA; }
  const b2Vec2& GetLocalAnchorB() const { return m_localAnchorB; }
  float GetReferenceAngle() const { return m_referenceAngle; }
  void SetStiffness(float stiffness);
  float GetStiffness() const { return m_stiffness; }
  void SetDamping(float damping);
  float GetDamping() const { return m_damping; }

 protected:
  friend class b2World;
  friend struct b2WorldImpl;
  friend struct b2WorldBatchImpl;
  b2WeldJoint(const b2WeldJointDef* def);
  void WriteDevice(float* anchors, float* params, float* state) const override;
  void ReadDeviceState(const float* state) override;
  b2Vec2 m_localAnchorA;
  b2Vec2 m_localAnchorB;
  float m_referenceAngle, m_stiffness, m_damping;
  mutable float m_impulse[3];
};

/// b2_distance_joint.h:30-170: rigid rod, spring (stiffness / damping) and min / max length limits
struct b2DistanceJointDef : public b2JointDef {
  b2DistanceJointDef() {
    type = e_distanceJoint;
    localAnchorA.Set(0.0f, 0.0f);
    localAnchorB.Set(0.0f, 0.0f);
    length = 1.0f;
    minLength = 0.0f;
    maxLength = FLT_MAX;
    stiffness = 0.0f;
    damping = 0.0f;
  }
  void Initialize(b2Body* bodyA, b2Body* bodyB, const b2Vec2& anchorA, const b2Vec2& anchorB);
  b2Vec2 localAnchorA;
  b2Vec2 localAnchorB;
  float length;
  float minLength;
  float maxLength;
  float stiffness;
  float damping;
};

class b2DistanceJoint : public b2Joint {
 public:
  b2Vec2 GetAnchorA() const override;
  b2Vec2 GetAnchorB() const override;
  /// along the current anchor-to-anchor direction (the reference uses the direction at the start of the last step)
  b2Vec2 GetReactionForce(float inv_dt) const override;
  float GetReactionTorque(float inv_dt) const override;
  const b2Vec2& GetLocalAnchorA() const { return m_localAnchorA; }
  const b2Vec2& GetLocalAnchorB() const { return m_localAnchorB; }
  float GetLength() const { return m_length; }
  float SetLength(float length);
  float GetMinLength() const { return m_minLength; }
  float SetMinLength(float minLength);
  float GetMaxLength() const { return m_maxLength; }
  float SetMaxLength(float maxLength);
  float GetCurrentLength() const;
  void SetStiffness(float stiffness);
  float GetStiffness() const { return m_stiffness; }
  void SetDamping(float damping);
  float GetDamping() const { return m_damping; }

 protected:
  friend class b2World;
  friend struct b2WorldImpl;
  friend struct b2WorldBatchImpl;
  b2DistanceJoint(const b2DistanceJointDef* def);
  void WriteDevice(float* anchors, float* params, float* state) const override;
  void ReadDeviceState(const float* state) override;
  b2Vec2 m_localAnchorA;
  b2Vec2 m_localAnchorB;
  float m_length, m_minLength, m_maxLength, m_stiffness, m_damping;
  mutable float m_impulse, m_lowerImpulse, m_upperImpulse;
};

// ---- callbacks (b2_world_callbacks.h) ---------------------------------------------------------
class b2DestructionListener {
 public:
  virtual ~b2DestructionListener() {}
  virtual void SayGoodbye(b2Joint* joint) = 0;
  virtual void SayGoodbye(b2Fixture* fixture) = 0;
};

class b2ContactFilter {
 public:
  virtual ~b2ContactFilter() {}
  virtual bool ShouldCollide(b2Fixture* fixtureA, b2Fixture* fixtureB);
};

struct b2ContactImpulse {
  float normalImpulses[b2_maxManifoldPoints];
  float tangentImpulses[b2_maxManifoldPoints];
  int32 count;
};

/// b2_world_callbacks.h:133-161
class b2QueryCallback {
 public:
  virtual ~b2QueryCallback() {}
  /// return false to terminate the query
  virtual bool ReportFixture(b2Fixture* fixture) = 0;
};
class b2RayCastCallback {
 public:
  virtual ~b2RayCastCallback() {}
  /// return -1 to ignore the fixture, 0 to terminate, a fraction to clip the ray, 1 to continue
  virtual float ReportFixture(b2Fixture* fixture, const b2Vec2& point, const b2Vec2& normal, float fraction) = 0;
};

class b2ContactListener {
 public:
  virtual ~b2ContactListener() {}
  virtual void BeginContact(b2Contact* contact) { B2_NOT_USED(contact); }
  virtual void EndContact(b2Contact* contact) { B2_NOT_USED(contact); }
  virtual void PreSolve(b2Contact* contact, const b2Manifold* oldManifold) {
    B2_NOT_USED(contact);
    B2_NOT_USED(oldManifold);
  }
  virtual void PostSolve(b2Contact* contact, const b2ContactImpulse* impulse) {
    B2_NOT_USED(contact);
    B2_NOT_USED(impulse);
  }
};

struct b2Profile {
  float step;
  float collide;
  float solve;
  float broadphase;
  float solveTOI;
};

// ---- the world --------------------------------------------------------------------------------
class b2World {
 public:
  b2World(const b2Vec2& gravity);
  ~b2World();
  void SetDestructionListener(b2DestructionListener* listener) { m_destructionListener = listener; }
  /// Category / mask / group filtering (the default b2ContactFilter) and joint collideConnected run on
  /// the device.  A user subclass is consulted on the host after every pair refresh for the pairs that
  /// passed the device rule (and again every step for rejected pairs that still overlap).
  void SetContactFilter(b2ContactFilter* filter);
  void SetContactListener(b2ContactListener* listener) { m_contactListener = listener; }
  b2Body* CreateBody(const b2BodyDef* def);
  void DestroyBody(b2Body* body);
  b2Joint* CreateJoint(const b2JointDef* def);
  void DestroyJoint(b2Joint* joint);
  /// b2_world.cpp:1193-1246, on the device's broadphase tree.  Fixtures are reported in ascending
  /// creation order (QueryAABB) / ascending fraction (RayCast), not in the reference's tree order.
  void QueryAABB(b2QueryCallback* callback, const b2AABB& aabb);
  void RayCast(b2RayCastCallback* callback, const b2Vec2& point1, const b2Vec2& point2);
  void Step(float timeStep, int32 velocityIterations, int32 positionIterations, int32 particleIterations);
  void Step(float timeStep, int32 velocityIterations, int32 positionIterations) {
    Step(timeStep, velocityIterations, positionIterations, 1);
  }
  void ClearForces();
  b2Body* GetBodyList() { return m_bodyListHead; }
  const b2Body* GetBodyList() const { return m_bodyListHead; }
  b2Joint* GetJointList() { return m_jointList; }
  const b2Joint* GetJointList() const { return m_jointList; }
  b2Contact* GetContactListStart();
  b2Contact* GetContactListEnd();
  void SetAllowSleeping(bool flag);
  bool GetAllowSleeping() const { return m_allowSleep; }
  void SetWarmStarting(bool flag) { m_warmStarting = flag; }
  bool GetWarmStarting() const { return m_warmStarting; }
  void SetContinuousPhysics(bool flag) { m_continuousPhysics = flag; }
  bool GetContinuousPhysics() const { return m_continuousPhysics; }
  void SetSubStepping(bool flag) { m_subStepping = flag; }
  bool GetSubStepping() const { return m_subStepping; }
  int32 GetProxyCount() const;
  int32 GetBodyCount() const { return m_bodyCount; }
  int32 GetJointCount() const { return m_jointCount; }
  int32 GetContactCount() const;
  int32 GetTreeHeight() const { return 0; }
  void SetGravity(const b2Vec2& gravity) { m_gravity = gravity; }
  /// b2_world.cpp:1287-1306: subtract newOrigin from every position (large worlds, floating origin)
  void ShiftOrigin(const b2Vec2& newOrigin);
  b2Vec2 GetGravity() const { return m_gravity; }
  bool IsLocked() const { return m_locked; }
  void SetAutoClearForces(bool flag) { m_clearForces = flag; }
  bool GetAutoClearForces() const { return m_clearForces; }
  const b2Profile& GetProfile() const { return m_profile; }

  // ---- extensions of the B200 build (not in the reference API) ----
  /// device capacities used by worlds constructed afterwards (grown automatically when exceeded)
  static void SetDefaultCapacity(int32 bodies, int32 fixtures, int32 contacts);
  /// CUDA device ordinal used by worlds constructed afterwards (default: env B2G_DEVICE or 0)
  static void SetDefaultDevice(int32 device);
  /// B2G_SOLVER_COLOURED (0, default) or B2G_SOLVER_SEQUENTIAL (1)
  void SetSolverMode(int32 mode) { m_solverMode = mode; }
  int32 GetSolverMode() const { return m_solverMode; }
  /// fill GetProfile() from CUDA events (adds a synchronisation per step)
  void SetProfiling(bool flag);
  /// rows of the device body table in use (live bodies + destroyed ones not yet reused): stays bounded in a world
  /// that keeps creating and destroying bodies
  int32 GetBodyIndexCount() const;
  b2WorldImpl* GetImpl() { return m_impl; }

 private:
  friend class b2Body;
  friend class b2Fixture;
  friend class b2Contact;
  friend class b2WorldBatch;
  friend struct b2WorldImpl;
  friend struct b2WorldBatchImpl;
  b2WorldImpl* m_impl;
  b2Body* m_bodyListHead;
  b2Body* m_bodyListTail;
  b2Joint* m_jointList;
  int32 m_bodyCount;
  int32 m_jointCount;
  b2Vec2 m_gravity;
  bool m_allowSleep;
  b2DestructionListener* m_destructionListener;
  b2ContactFilter* m_contactFilter;
  b2ContactListener* m_contactListener;
  bool m_newContacts;
  bool m_locked;
  bool m_clearForces;
  bool m_warmStarting;
  bool m_continuousPhysics;
  bool m_subStepping;
  int32 m_solverMode;
  b2Profile m_profile;
};

// ---- extension of the B200 build: many worlds, one device pass -----------------------------------
/// Thousands of small independent worlds (RL environments, parameter sweeps) cost one kernel-launch chain
/// EACH when every b2World steps on its own.  A b2WorldBatch puts its members into one device arena (bodies
/// of different worlds never meet) and steps all of them together: world k of a batch gets exactly the floats
/// it would get alone.  Members keep their public API (bodies, fixtures, joints, getters, setters, queries);
/// contact listeners and user contact filters are not called for batched worlds.  Gravity, the world flags
/// (sleeping, warm starting, ...) and the solver mode are taken from the first member.
struct b2WorldBatchImpl;
class b2WorldBatch {
 public:
  b2WorldBatch();
  ~b2WorldBatch();
  /// A world that has not been stepped yet (it may already hold bodies, fixtures and joints).  False otherwise.
  bool Add(b2World* world);
  int32 GetWorldCount() const;
  b2World* GetWorld(int32 index) const;
  /// b2World::Step of every member
  void Step(float timeStep, int32 velocityIterations, int32 positionIterations);
  void SetProfiling(bool flag);
  /// device time of the last Step (with SetProfiling(true))
  float GetLastStepMilliseconds() const;

 private:
  b2WorldBatchImpl* m_impl;
};

#endif
