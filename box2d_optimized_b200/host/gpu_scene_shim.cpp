// gpu_scene_shim.cpp — the shared scene shim expanded against THIS repo's drop-in C++ API.
// Same source as the oracle's expansion (oracle/ref_harness.cpp); only the headers differ.
#include "box2d/box2d.h"
#include "b2cuda.h"
#define SHIM(name) b2gpu_##name
#include "b2_scene_shim.h"

extern "C" {
// extensions that only exist on the CUDA build
void b2gpu_scene_set_solver_mode(void* h, int mode) { static_cast<Scene*>(h)->world->SetSolverMode(mode); }
void b2gpu_set_default_capacity(int bodies, int fixtures, int contacts) {
  b2World::SetDefaultCapacity(bodies, fixtures, contacts);
}
void b2gpu_set_default_device(int device) { b2World::SetDefaultDevice(device); }
void b2gpu_scene_set_profiling(void* h, int on) { static_cast<Scene*>(h)->world->SetProfiling(on != 0); }
// out[5] = step, collide, solve, broadphase, solveTOI milliseconds of the last step
void b2gpu_scene_get_profile(void* h, float* out) {
  const b2Profile& p = static_cast<Scene*>(h)->world->GetProfile();
  out[0] = p.step; out[1] = p.collide; out[2] = p.solve; out[3] = p.broadphase; out[4] = p.solveTOI;
}
}
