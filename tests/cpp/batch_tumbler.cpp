// BASELINE config 4 through the drop-in API: W copies of the reference's OWN tumbler benchmark (b3 of
// testbed/benchmarks/benchmarks.h, included unchanged: one box spawned per step into a motor-driven container),
// members of one b2WorldBatch.  Prints wall time per batch step in the spawn phase (the host creates W bodies per
// step) and in the full phase (no host edits).
// usage: batch_tumbler [worlds [size [steps_after_spawn]]]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "box2d/box2d.h"
#include "benchmarks.h"

int main(int argc, char** argv) {
  const int W = argc > 1 ? atoi(argv[1]) : 256, size = argc > 2 ? atoi(argv[2]) : 500, extra = argc > 3 ? atoi(argv[3]) : 100;
  std::vector<b3*> bm(W);
  std::vector<b2World*> worlds(W);
  b2WorldBatch batch;
  batch.SetProfiling(true);
  for (int k = 0; k < W; ++k) {
    bm[k] = new b3();
    worlds[k] = new b2World(bm[k]->gravity);
    worlds[k]->SetContinuousPhysics(false);
    bm[k]->InitWorld(worlds[k], size);
    if (!batch.Add(worlds[k])) return 2;
  }
  auto now = [] { return std::chrono::high_resolution_clock::now(); };
  double spawnMs = 0.0, fullMs = 0.0, fullDev = 0.0;
  for (int s = 0; s < size + extra; ++s) {
    auto t0 = now();
    batch.Step(bm[0]->timeStep, bm[0]->velocityIterations, bm[0]->positionIterations);
    for (int k = 0; k < W; ++k) bm[k]->AfterWorldStep(worlds[k]);
    volatile float sink = worlds[W - 1]->GetBodyList()->GetPosition().x;  // wait for the step like any consumer
    (void)sink;
    const double ms = std::chrono::duration<double, std::milli>(now() - t0).count();
    if (s < size) spawnMs += ms;
    else fullMs += ms, fullDev += batch.GetLastStepMilliseconds();
  }
  int bodies = 0, contacts = 0;
  for (int k = 0; k < W; ++k) bodies += worlds[k]->GetBodyCount();
  contacts = worlds[0]->GetContactCount();
  printf("BATCH tumbler worlds=%d size=%d bodies=%d contacts_world0=%d spawn_phase_ms_per_step=%.3f full_ms_per_step=%.3f "
         "full_device_ms_per_step=%.3f body_steps_per_s=%.1f\n",
         W, size, bodies, contacts, spawnMs / size, fullMs / extra, fullDev / extra, bodies / (fullMs / extra) * 1000.0);
  for (int k = 0; k < W; ++k) delete worlds[k];
  return 0;
}
