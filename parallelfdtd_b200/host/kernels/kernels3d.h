// kernels3d.h -- launchFDTD3d / launchFDTD3dDouble / launchFDTD3dStep with the reference's signatures
// (reference src/kernels/kernels3d.h:49-90) over the C ABI.  The per-step host work of the reference
// (one getSourceSample + one-element H2D copy per source, one D2D copy per receiver, three device
// synchronisations; kernels3d.cu:83-181) becomes: source table built once, uploaded once, steps
// enqueued in blocks of PROGRESS_MOD, responses fetched once.
#pragma once
#include <vector>
#include "../base/SimulationParameters.h"
#include "cudaMesh.h"

#define PROGRESS_MOD 100

namespace pfdtd_host {
inline void bind_sources_receivers(CudaMesh* d_mesh, SimulationParameters* sp, bool dbl, unsigned int steps) {
  const unsigned int ns = sp->getNumSources(), nr = sp->getNumReceivers();
  std::vector<int32_t> sxyz(3 * ns), stype(ns), rxyz(3 * nr);
  for (unsigned int i = 0; i < ns; i++) {
    nv::Vec3i p = sp->getSourceElementCoordinates(i);
    sxyz[3 * i] = p.x; sxyz[3 * i + 1] = p.y; sxyz[3 * i + 2] = p.z;
    stype[i] = (int32_t)sp->getSource(i).getSourceType();
  }
  for (unsigned int i = 0; i < nr; i++) {
    nv::Vec3i p = sp->getReceiverElementCoordinates(i);
    rxyz[3 * i] = p.x; rxyz[3 * i + 1] = p.y; rxyz[3 * i + 2] = p.z;
  }
  if (dbl) {
    std::vector<double> tab; sp->fillSourceTableDouble(tab, steps);
    pfdtd_safe(pfdtd_set_sources(d_mesh->handle(), ns, ns ? &sxyz[0] : 0, ns ? &stype[0] : 0, ns ? &tab[0] : 0, steps), "launchFDTD3d: sources");
  } else {
    std::vector<float> tab; sp->fillSourceTable(tab, steps);
    pfdtd_safe(pfdtd_set_sources(d_mesh->handle(), ns, ns ? &sxyz[0] : 0, ns ? &stype[0] : 0, ns ? &tab[0] : 0, steps), "launchFDTD3d: sources");
  }
  pfdtd_safe(pfdtd_set_receivers(d_mesh->handle(), nr, nr ? &rxyz[0] : 0), "launchFDTD3d: receivers");
}
struct CallbackBridge {   // bool(*)(void) -> int(*)(void) without a capturing lambda; per calling thread (pfdtd_run calls
                          // back on the thread that called it), so Apps running on different threads do not share it
  static bool (*&interrupt())(void) { static thread_local bool (*f)(void) = 0; return f; }
  static int call() { return interrupt()() ? 1 : 0; }
};
inline float run(CudaMesh* d_mesh, SimulationParameters* sp, void* h_return_ptr, bool dbl, bool (*interruptCallback)(void),
                 void (*progressCallback)(int, int, float)) {
  const unsigned int steps = sp->getNumSteps();
  bind_sources_receivers(d_mesh, sp, dbl, steps);
  d_mesh->setBound(sp, sp->generation());
  CallbackBridge::interrupt() = interruptCallback;
  float sps = 0.f;
  int rc = pfdtd_run(d_mesh->handle(), steps, h_return_ptr, interruptCallback ? &CallbackBridge::call : 0, progressCallback, &sps);
  if (rc != PFDTD_ERR_INTERRUPTED) pfdtd_safe(rc, "launchFDTD3d");
  return sps;
}
}  // namespace pfdtd_host

// h_return_ptr: caller-owned [numReceivers][numSteps]; returns seconds per step (reference kernels3d.cu:196-202)
inline float launchFDTD3d(CudaMesh* d_mesh, SimulationParameters* sp, float* h_return_ptr, bool (*interruptCallback)(void),
                          void (*progressCallback)(int, int, float)) {
  return pfdtd_host::run(d_mesh, sp, h_return_ptr, false, interruptCallback, progressCallback);
}
inline float launchFDTD3dDouble(CudaMesh* d_mesh, SimulationParameters* sp, double* h_return_ptr, bool (*interruptCallback)(void),
                                void (*progressCallback)(int, int, float)) {
  return pfdtd_host::run(d_mesh, sp, h_return_ptr, true, interruptCallback, progressCallback);
}
// one step; receivers of this step land in h_return_ptr[rec * numSteps + step] (reference kernels3d.cu:376-482)
inline void launchFDTD3dStep(CudaMesh* d_mesh, SimulationParameters* sp, float* h_return_ptr, unsigned int step, int step_direction,
                             void (*progressCallback)(int, int, float)) {
  // the tables inside the solver are rebuilt whenever the parameters object changed since they were uploaded (any
  // setter bumps its generation), another parameters object is passed, or the mesh was re-partitioned
  if (!d_mesh->boundTo(sp, sp->generation())) {
    pfdtd_host::bind_sources_receivers(d_mesh, sp, d_mesh->isDouble(), sp->getNumSteps());
    d_mesh->setBound(sp, sp->generation());
  }
  pfdtd_safe(pfdtd_step(d_mesh->handle(), step, step_direction, h_return_ptr, sp->getNumSteps()), "launchFDTD3dStep");
  if (progressCallback && step % PROGRESS_MOD == 0) progressCallback((int)step, (int)sp->getNumSteps(), 0.f);
}
