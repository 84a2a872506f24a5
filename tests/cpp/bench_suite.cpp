// Driver for the reference's OWN benchmark scenes (testbed/benchmarks/benchmarks.h, b1..b14), included
// unchanged from where it lies (-I<reference>/testbed/benchmarks).  The same source is compiled twice:
// against the reference's headers + its CPU library, and against this repo's drop-in headers + CUDA
// library.  Like the reference's testbed/benchmarks/single.cpp:46-63 it builds each world with continuous
// physics off, steps it `simulationSteps` times and prints the wall time; in addition it prints an
// end-state summary per scene so that the two builds can be compared (tests/test_reference_benchmarks.py).
//
// usage: bench_suite [first last [steps]]     (1-based inclusive benchmark numbers; default 1 14)
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "box2d/box2d.h"
#include "benchmarks.h"

int main(int argc, char** argv) {
  int first = argc > 1 ? atoi(argv[1]) : 1, last = argc > 2 ? atoi(argv[2]) : benchmark_count;
  int stepsOverride = argc > 3 ? atoi(argv[3]) : 0;
  for (int k = first - 1; k < last && k < benchmark_count; ++k) {
    b2Benchmark* bm = benchmarks[k];
    b2World world(bm->gravity);
    world.SetContinuousPhysics(false);
    bm->InitWorld(&world);
    const int steps = stepsOverride > 0 ? stepsOverride : bm->simulationSteps;
    double total = 0.0, worst = 0.0;
    for (int s = 0; s < steps; ++s) {
      auto t0 = std::chrono::high_resolution_clock::now();
      bm->StepWorld(&world);
      // the getter makes the GPU build wait for the step, as any consumer of the result would
      volatile float sink = world.GetBodyList() ? world.GetBodyList()->GetPosition().x : 0.0f;
      (void)sink;
      auto t1 = std::chrono::high_resolution_clock::now();
      const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
      total += ms;
      worst = std::max(worst, ms);
    }
    int bodies = 0, dynamic = 0, awake = 0;
    double sx = 0, sy = 0, ke = 0, miny = 1e30, maxy = -1e30, minx = 1e30, maxx = -1e30;
    std::vector<float> ys;
    for (b2Body* b = world.GetBodyList(); b; b = b->GetNext()) {
      ++bodies;
      if (b->GetType() != b2_dynamicBody) continue;
      ++dynamic;
      if (b->IsAwake()) ++awake;
      const b2Vec2 p = b->GetPosition(), v = b->GetLinearVelocity();
      sx += p.x;
      sy += p.y;
      ke += 0.5 * b->GetMass() * (v.x * v.x + v.y * v.y);
      miny = std::min(miny, (double)p.y);
      maxy = std::max(maxy, (double)p.y);
      minx = std::min(minx, (double)p.x);
      maxx = std::max(maxx, (double)p.x);
      ys.push_back(p.y);
    }
    std::sort(ys.begin(), ys.end());
    const double med = ys.empty() ? 0.0 : ys[ys.size() / 2];
    const double q10 = ys.empty() ? 0.0 : ys[ys.size() / 10], q90 = ys.empty() ? 0.0 : ys[ys.size() * 9 / 10];
    printf("BENCH b%d \"%s\" steps=%d total_ms=%.2f max_ms=%.3f bodies=%d dynamic=%d awake=%d contacts=%d "
           "mean_x=%.4f mean_y=%.4f min_y=%.4f max_y=%.4f min_x=%.4f max_x=%.4f med_y=%.4f q10_y=%.4f q90_y=%.4f ke=%.4f\n",
           k + 1, bm->name.c_str(), steps, total, worst, bodies, dynamic, awake, world.GetContactCount(),
           dynamic ? sx / dynamic : 0.0, dynamic ? sy / dynamic : 0.0, miny, maxy, minx, maxx, med, q10, q90, ke);
    fflush(stdout);
  }
  return 0;
}
