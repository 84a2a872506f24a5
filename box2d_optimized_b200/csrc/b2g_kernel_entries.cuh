// b2g_kernel_entries.cuh — kernel-level C-ABI entry points (see the last section of
// include/b2cuda.h).  Each one stages explicit host arrays, runs the SAME device functions the
// step uses, and copies the result back, so parity tests can pin one kernel at a time against
// the oracle with the oracle's own ordered inputs.
#pragma once

__global__ void k_entry_aabbs(int n, const int* __restrict__ type, const int* __restrict__ off,
                              const float4* __restrict__ shapes, const float4* __restrict__ xf, float4* aabb) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  aabb[i] = shape_aabb(shapes, type[i], off[i], xf_from4(xf[i]));
}

extern "C" int b2g_compute_aabbs(int32_t device, int32_t n, const int32_t* type, const int32_t* shape_off,
                                 const float* shape_quads, int32_t num_quads, const float* xf, float* aabb) {
  if (n < 0 || !type || !shape_off || !shape_quads || !xf || !aabb) return B2G_ERR_INVALID;
  int rc = use_device(device);
  if (rc) return rc;
  if (n == 0) return B2G_OK;
  DevBuf dT, dO, dS, dX, dA;
  CK(dT.upload(type, (size_t)n * 4));
  CK(dO.upload(shape_off, (size_t)n * 4));
  CK(dS.upload(shape_quads, (size_t)num_quads * 16));
  CK(dX.upload(xf, (size_t)n * 16));
  CK(dA.alloc((size_t)n * 16));
  k_entry_aabbs<<<div_up(n, 256), 256>>>(n, dT.as<int>(), dO.as<int>(), dS.as<float4>(), dX.as<float4>(),
                                         dA.as<float4>());
  CK(cudaGetLastError());
  CK(cudaMemcpy(aabb, dA.p, (size_t)n * 16, cudaMemcpyDeviceToHost));
  return B2G_OK;
}

__global__ void k_entry_rotations(int n, const float* __restrict__ angle, float2* sc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Rot q = rot_set(angle[i]);
  sc[i] = make_float2(q.s, q.c);
}

extern "C" int b2g_rotations(int32_t device, int32_t n, const float* angle, float* sin_cos) {
  if (n < 0 || !angle || !sin_cos) return B2G_ERR_INVALID;
  int rc = use_device(device);
  if (rc) return rc;
  if (n == 0) return B2G_OK;
  DevBuf dA, dR;
  CK(dA.upload(angle, (size_t)n * 4));
  CK(dR.alloc((size_t)n * 8));
  k_entry_rotations<<<div_up(n, 256), 256>>>(n, dA.as<float>(), dR.as<float2>());
  CK(cudaGetLastError());
  CK(cudaMemcpy(sin_cos, dR.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
  return B2G_OK;
}

__global__ void __launch_bounds__(128)
k_entry_collide(int n, const int* __restrict__ typeA, const int* __restrict__ offA, const float4* __restrict__ xfA,
                const int* __restrict__ typeB, const int* __restrict__ offB, const float4* __restrict__ xfB,
                const float4* __restrict__ shapes, float4* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Manifold m;
  collide_dispatch(m, shapes, typeA[i], offA[i], xf_from4(xfA[i]), typeB[i], offB[i], xf_from4(xfB[i]));
  float4 q0, q1, q2, q3;
  manifold_pack(m, q0, q1, q2, q3);
  out[4 * i] = q0;
  out[4 * i + 1] = q1;
  out[4 * i + 2] = q2;
  out[4 * i + 3] = q3;
}

extern "C" int b2g_collide_pairs(int32_t device, int32_t n, const int32_t* type_a, const int32_t* shape_off_a,
                                 const float* xf_a, const int32_t* type_b, const int32_t* shape_off_b,
                                 const float* xf_b, const float* shape_quads, int32_t num_quads, float* manifold) {
  if (n < 0 || !type_a || !shape_off_a || !xf_a || !type_b || !shape_off_b || !xf_b || !shape_quads || !manifold)
    return B2G_ERR_INVALID;
  int rc = use_device(device);
  if (rc) return rc;
  if (n == 0) return B2G_OK;
  DevBuf dTA, dOA, dXA, dTB, dOB, dXB, dS, dM;
  CK(dTA.upload(type_a, (size_t)n * 4));
  CK(dOA.upload(shape_off_a, (size_t)n * 4));
  CK(dXA.upload(xf_a, (size_t)n * 16));
  CK(dTB.upload(type_b, (size_t)n * 4));
  CK(dOB.upload(shape_off_b, (size_t)n * 4));
  CK(dXB.upload(xf_b, (size_t)n * 16));
  CK(dS.upload(shape_quads, (size_t)num_quads * 16));
  CK(dM.alloc((size_t)n * 64));
  k_entry_collide<<<div_up(n, 128), 128>>>(n, dTA.as<int>(), dOA.as<int>(), dXA.as<float4>(), dTB.as<int>(),
                                           dOB.as<int>(), dXB.as<float4>(), dS.as<float4>(), dM.as<float4>());
  CK(cudaGetLastError());
  CK(cudaMemcpy(manifold, dM.p, (size_t)n * 64, cudaMemcpyDeviceToHost));
  return B2G_OK;
}

extern "C" int b2g_find_pairs(int32_t device, int32_t n, const float* aabb, const int32_t* body,
                              const int32_t* world, const uint8_t* dyn, int32_t* pairs, int32_t capacity,
                              int32_t* num_pairs) {
  if (n < 0 || !aabb || !body || !dyn || !pairs || !num_pairs || capacity <= 0) return B2G_ERR_INVALID;
  *num_pairs = 0;
  if (n == 0) return use_device(device);
  int maxBody = 0, maxWorld = 0;
  for (int i = 0; i < n; ++i) {
    if (body[i] < 0) return B2G_ERR_INVALID;
    if (body[i] > maxBody) maxBody = body[i];
    if (world && world[i] > maxWorld) maxWorld = world[i];
  }
  b2gArenaDef def;
  memset(&def, 0, sizeof(def));
  def.device = device;
  def.num_worlds = maxWorld + 1;
  def.max_bodies = maxBody + 1;
  def.max_fixtures = n;
  def.max_shape_quads = 1;
  def.max_contacts = capacity;
  b2gArena* A = nullptr;
  int rc = b2g_arena_create(&def, &A);
  if (rc) return rc;
  const int nb = maxBody + 1;
  std::vector<uint32_t> bflags(nb, B2G_BODY_ENABLED);
  std::vector<int32_t> bworld(nb, 0);
  std::vector<uint32_t> tflags(n, 0), filter((size_t)n * 2);
  for (int i = 0; i < n; ++i) {
    if (dyn[i]) bflags[body[i]] |= (2u << B2G_BODY_TYPE_SHIFT);
    if (world) bworld[body[i]] = world[i];
    filter[2 * i] = 0x0001u | (0xffffu << 16);
    filter[2 * i + 1] = 0;
  }
  b2gBodyArrays ba;
  memset(&ba, 0, sizeof(ba));
  ba.flags = bflags.data();
  ba.world = bworld.data();
  rc = b2g_upload_bodies(A, 0, nb, &ba);
  b2gFixtureArrays fa;
  memset(&fa, 0, sizeof(fa));
  fa.body = const_cast<int32_t*>(body);
  fa.type_flags = tflags.data();
  fa.filter = filter.data();
  if (!rc) rc = b2g_upload_fixtures(A, 0, n, &fa);
  if (!rc && cudaMemcpy(A->fAabb, aabb, (size_t)n * 16, cudaMemcpyHostToDevice) != cudaSuccess) rc = B2G_ERR_CUDA;
  if (!rc) {
    A->aabbAllDirty = 2;  // boxes are given: do not recompute them from shapes
    cudaMemsetAsync(A->dCounts, 0, sizeof(StepCounts), A->stream);
    rc = find_new_contacts(A, 0);
    *num_pairs = A->hCounts->numPairs;
  }
  if (!rc) {
    PackedContacts pc;
    rc = pack_contacts(A, pc);
    for (size_t k = 0; k < pc.order.size() && !rc; ++k) {
      unsigned long long key = pc.keys[pc.order[k]];
      pairs[2 * k] = (int)(key >> 32);
      pairs[2 * k + 1] = (int)(key & 0xffffffffull);
    }
  }
  b2g_arena_destroy(A);
  return rc;
}

__global__ void k_entry_prepare(int nc, SolverPlanes S, const int2* __restrict__ index, const float4* __restrict__ man,
                                const float4* __restrict__ material, const float2* __restrict__ radii,
                                const float4* __restrict__ pos, const float4* __restrict__ vel,
                                const float4* __restrict__ bodyMass, const float4* __restrict__ bodyCenter,
                                float dtRatio, int warmStarting) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nc) return;
  Manifold m;
  manifold_unpack(m, man[4 * s], man[4 * s + 1], man[4 * s + 2], man[4 * s + 3]);
  int2 ix = index[s];
  float2 r = radii[s];
  prepare_constraint(S, s, s, m, ix.x, ix.y, ix.x, ix.y, material[s], r.x, r.y,
                     GlobalBodies{const_cast<float4*>(pos)}, GlobalBodies{const_cast<float4*>(vel)}, bodyMass,
                     bodyCenter, dtRatio, warmStarting != 0);
}

__global__ void k_entry_split_mass(int nb, const float4* __restrict__ in, float4* mass, float4* center) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  float4 v = in[b];
  mass[b] = make_float4(v.x, v.y, 0.0f, 0.0f);
  center[b] = make_float4(v.z, v.w, 0.0f, 0.0f);
}

__global__ void k_entry_store(int nc, SolverPlanes S, float4* man) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nc) return;
  int4 ix = S.idx[s];
  float4 imp = S.imp[s];
  man[4 * s + 1].z = imp.x;
  man[4 * s + 1].w = imp.y;
  if (ix.z == 2) {
    man[4 * s + 2].z = imp.z;
    man[4 * s + 2].w = imp.w;
  }
}

__global__ void k_entry_integrate_positions(int nb, float4* pos, float4* vel, float h) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  float4 p4 = pos[b], v4 = vel[b];
  float2 v = make_float2(v4.x, v4.y);
  float w = v4.z;
  float2 translation = h * v;
  if (dot2(translation, translation) > B2G_MAX_TRANSLATION_SQ) {
    float ratio = B2G_MAX_TRANSLATION / len2(translation);
    v.x *= ratio;
    v.y *= ratio;
  }
  float rotation = h * w;
  if (rotation * rotation > B2G_MAX_ROTATION_SQ) {
    float ratio = B2G_MAX_ROTATION / absf_(rotation);
    w *= ratio;
  }
  p4.x += h * v.x;
  p4.y += h * v.y;
  p4.z += h * w;
  pos[b] = p4;
  vel[b] = make_float4(v.x, v.y, w, v4.w);
}

extern "C" int b2g_solve_sequential(int32_t device, int32_t nb, float* pos, float* vel, const float* mass, int32_t nc,
                                    const int32_t* index, float* manifold, const float* material,
                                    const float* radii, float dt, float dt_ratio, int32_t warm_starting,
                                    int32_t vel_iters, int32_t pos_iters, float* vel_iterates, float* pos_iterates,
                                    int32_t* pos_iters_done) {
  if (nb <= 0 || nc < 0 || !pos || !vel || !mass || (nc > 0 && (!index || !manifold || !material || !radii)) ||
      pos_iters > B2G_MAX_POS_ITERS)
    return B2G_ERR_INVALID;
  int rc = use_device(device);
  if (rc) return rc;
  DevBuf dPos, dVel, dMassIn, dMass, dCen, dIdx, dMan, dMat, dRad, dRoot, dPen;
  CK(dPos.upload(pos, (size_t)nb * 16));
  CK(dVel.upload(vel, (size_t)nb * 16));
  CK(dMassIn.upload(mass, (size_t)nb * 16));
  CK(dMass.alloc((size_t)nb * 16));
  CK(dCen.alloc((size_t)nb * 16));
  k_entry_split_mass<<<div_up(nb, 256), 256>>>(nb, dMassIn.as<float4>(), dMass.as<float4>(), dCen.as<float4>());
  const int ncap = nc > 0 ? nc : 1;
  CK(dIdx.upload(index ? (const void*)index : (const void*)pos, nc > 0 ? (size_t)nc * 8 : 8));
  CK(dMan.upload(manifold ? (const void*)manifold : (const void*)pos, nc > 0 ? (size_t)nc * 64 : 16));
  CK(dMat.upload(material ? (const void*)material : (const void*)pos, nc > 0 ? (size_t)nc * 16 : 16));
  CK(dRad.upload(radii ? (const void*)radii : (const void*)pos, nc > 0 ? (size_t)nc * 8 : 8));
  CK(dRoot.alloc((size_t)ncap * 4));
  CK(cudaMemset(dRoot.p, 0, (size_t)ncap * 4));
  CK(dPen.alloc(sizeof(uint32_t) * (B2G_MAX_POS_ITERS + 1)));
  CK(cudaMemset(dPen.p, 0, sizeof(uint32_t) * (B2G_MAX_POS_ITERS + 1)));
  DevBuf pl[13];
  for (int k = 0; k < 13; ++k) CK(pl[k].alloc((size_t)ncap * 16));
  SolverPlanes S;
  S.nf = pl[0].as<float4>();
  S.r1 = pl[1].as<float4>();
  S.r2 = pl[2].as<float4>();
  S.m1 = pl[3].as<float4>();
  S.m2 = pl[4].as<float4>();
  S.kk = pl[5].as<float4>();
  S.mass = pl[6].as<float4>();
  S.idx = pl[7].as<int4>();
  S.imp = pl[8].as<float4>();
  S.pn = pl[9].as<float4>();
  S.pp = pl[10].as<float4>();
  S.pc = pl[11].as<float4>();
  S.pr = pl[12].as<float4>();

  if (nc > 0) {
    k_entry_prepare<<<div_up(nc, 128), 128>>>(nc, S, dIdx.as<int2>(), dMan.as<float4>(), dMat.as<float4>(),
                                              dRad.as<float2>(), dPos.as<float4>(), dVel.as<float4>(),
                                              dMass.as<float4>(), dCen.as<float4>(), dt_ratio, warm_starting);
    if (warm_starting) k_warm_start_seq<<<1, 1>>>(0, nc, S, dVel.as<float4>());
  }
  for (int it = 0; it < vel_iters; ++it) {
    if (nc > 0) k_solve_velocity_seq<<<1, 1>>>(0, nc, S, dVel.as<float4>());
    if (vel_iterates)
      CK(cudaMemcpy(vel_iterates + (size_t)it * nb * 4, dVel.p, (size_t)nb * 16, cudaMemcpyDeviceToHost));
  }
  if (nc > 0) k_entry_store<<<div_up(nc, 256), 256>>>(nc, S, dMan.as<float4>());
  k_entry_integrate_positions<<<div_up(nb, 256), 256>>>(nb, dPos.as<float4>(), dVel.as<float4>(), dt);
  int done = 0;
  for (int it = 0; it < pos_iters; ++it) {
    if (nc > 0)
      k_solve_position_seq<<<1, 1>>>(0, nc, S, dPos.as<float4>(), dRoot.as<int>(), dPen.as<uint32_t>(), 1, it);
    ++done;
    if (pos_iterates)
      CK(cudaMemcpy(pos_iterates + (size_t)it * nb * 4, dPos.p, (size_t)nb * 16, cudaMemcpyDeviceToHost));
    uint32_t penBits = 0;
    CK(cudaMemcpy(&penBits, dPen.as<uint32_t>() + it, 4, cudaMemcpyDeviceToHost));
    float pen;
    memcpy(&pen, &penBits, 4);
    if (pen <= 3.0f * B2G_LINEAR_SLOP) break;  // contactsOkay -> early exit (b2_island.cpp:403-408)
  }
  CK(cudaGetLastError());
  CK(cudaMemcpy(pos, dPos.p, (size_t)nb * 16, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(vel, dVel.p, (size_t)nb * 16, cudaMemcpyDeviceToHost));
  if (nc > 0) CK(cudaMemcpy(manifold, dMan.p, (size_t)nc * 64, cudaMemcpyDeviceToHost));
  if (pos_iters_done) *pos_iters_done = done;
  return B2G_OK;
}
