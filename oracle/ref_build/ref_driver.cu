// =============================================================================
// oracle/ref_build/ref_driver.cu -- driver that runs the UNMODIFIED reference hot path
// (CudaMesh::setupMesh -> makePartition -> launchFDTD3d[Double]) on a case file.
//
// TEST INFRASTRUCTURE ONLY. Compiled by oracle/Makefile together with the reference's
// own sources where they lie under /root/reference (never copied) into oracle/_ref/.
// Drives the reference the way its own tests do (tests/CudaMeshTest.cpp:220-258,
// :472-520): raw voxelizer-style `bid` + material volumes in place of the absent
// Voxelizer, then the reference's API only.
//
// usage: ref_fdtd <case.bin> <out.bin> [dump_nodes=0|1] [warmup_steps=0]
// warmup_steps > 0 (benchmark mode): launchFDTD3d[Double] is first run for that many untimed steps.
// Case/out formats: see oracle/casefile.py.
// =============================================================================
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <chrono>
#include <stdint.h>

#include "kernels/cudaUtils.h"
#include "kernels/cudaMesh.h"
#include "kernels/kernels3d.h"
#include "base/SimulationParameters.h"
#include "base/MaterialHandler.h"

static bool interruptNever(void) { return false; }
static void progressQuiet(int, int, float) {}

template <typename T> static bool rd(FILE* f, T* dst, size_t n = 1) { return fread(dst, sizeof(T), n, f) == n; }

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s case.bin out.bin [dump_nodes]\n", argv[0]); return 2; }
  int dump_nodes = argc > 3 ? atoi(argv[3]) : 0;
  int warmup_steps = argc > 4 ? atoi(argv[4]) : 0;
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror("case"); return 2; }
  char magic[8];
  uint32_t ver, vx, vy, vz, bx, by, bz, utype, is_double, steps, octave, n_parts, n_mat, n_src, n_rec, n_data;
  rd(f, magic, 8);
  if (memcmp(magic, "PFDTDCAS", 8)) { fprintf(stderr, "bad magic\n"); return 2; }
  rd(f, &ver); rd(f, &vx); rd(f, &vy); rd(f, &vz); rd(f, &bx); rd(f, &by); rd(f, &bz);
  rd(f, &utype); rd(f, &is_double); rd(f, &steps); rd(f, &octave);
  rd(f, &n_parts); rd(f, &n_mat); rd(f, &n_src); rd(f, &n_rec); rd(f, &n_data);
  std::vector<unsigned int> devices(n_parts);
  rd(f, devices.data(), n_parts);
  std::vector<float> coefs((size_t)n_mat * 20);
  rd(f, coefs.data(), coefs.size());
  std::vector<int32_t> src(n_src * 6), rec(n_rec * 3);
  rd(f, src.data(), src.size());
  rd(f, rec.data(), rec.size());
  std::vector<std::vector<double> > data(n_data);
  for (uint32_t i = 0; i < n_data; i++) { uint32_t len; rd(f, &len); data[i].resize(len); rd(f, data[i].data(), len); }
  size_t nvox = (size_t)vx * vy * vz;
  std::vector<unsigned char> bid(nvox), mat(nvox);
  if (!rd(f, bid.data(), nvox) || !rd(f, mat.data(), nvox)) { fprintf(stderr, "short case file\n"); return 2; }
  fclose(f);

  loggerInit();
  SimulationParameters sp;
  MaterialHandler mh;
  mh.addMaterials(coefs.data(), n_mat, 20);
  sp.setUpdateType((enum UpdateType)utype);
  sp.setOctave(octave);
  sp.setNumSteps(steps);
  sp.setSpatialFs(7000);
  sp.setAddPaddingToElementIdx(false);   // coordinates in the case file are final element indices
  float dx = sp.getC() / ((float)sp.getSpatialFs() * (float)sp.getLambda());
  for (uint32_t i = 0; i < n_data; i++) {
    std::vector<float> d32(data[i].begin(), data[i].end());
    sp.addInputData(d32);
    sp.addInputDataDouble(data[i]);
  }
  for (uint32_t i = 0; i < n_src; i++)
    sp.addSource(Source(src[6 * i] * dx, src[6 * i + 1] * dx, src[6 * i + 2] * dx, (enum SrcType)src[6 * i + 3],
                        (enum InputType)src[6 * i + 4], (unsigned)src[6 * i + 5]));
  for (uint32_t i = 0; i < n_rec; i++) sp.addReceiver(Receiver(rec[3 * i] * dx, rec[3 * i + 1] * dx, rec[3 * i + 2] * dx));
  for (uint32_t i = 0; i < n_src; i++) {
    nv::Vec3i p = sp.getSourceElementCoordinates(i);
    if (p.x != src[6 * i] || p.y != src[6 * i + 1] || p.z != src[6 * i + 2]) { fprintf(stderr, "source %u coordinate round-trip failed\n", i); return 3; }
  }
  for (uint32_t i = 0; i < n_rec; i++) {
    nv::Vec3i p = sp.getReceiverElementCoordinates(i);
    if (p.x != rec[3 * i] || p.y != rec[3 * i + 1] || p.z != rec[3 * i + 2]) { fprintf(stderr, "receiver %u coordinate round-trip failed\n", i); return 3; }
  }

  try {
    cudasafe(cudaSetDevice(0), "set device 0");
    cudasafe(cudaFree(0), "context");   // context creation is not part of any timed span
    auto t_e2e0 = std::chrono::steady_clock::now();
    unsigned char* d_pos = toDevice<unsigned char>((unsigned)nvox, bid.data(), 0);
    unsigned char* d_mat = toDevice<unsigned char>((unsigned)nvox, mat.data(), 0);
    CudaMesh mesh;
    mesh.setDouble(is_double != 0);
    uint3 dim = make_uint3(vx, vy, vz), block = make_uint3(bx, by, bz);
    if (is_double)
      mesh.setupMeshDouble(d_pos, d_mat, mh.getNumberOfUniqueMaterials(), mh.getMaterialCoefficientPtrDouble(),
                           sp.getParameterPtrDouble(), dim, block, utype);
    else
      mesh.setupMesh(d_pos, d_mat, mh.getNumberOfUniqueMaterials(), mh.getMaterialCoefficientPtr(),
                     sp.getParameterPtr(), dim, block, utype);
    mesh.makePartition(n_parts, devices);

    std::vector<double> resp((size_t)n_rec * steps, 0.0);
    double t_ret = 0;
    double setup_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_e2e0).count();
    if (warmup_steps > 0) {
      sp.setNumSteps(warmup_steps);
      if (is_double) { std::vector<double> w((size_t)n_rec * warmup_steps, 0.0); launchFDTD3dDouble(&mesh, &sp, w.data(), interruptNever, progressQuiet); }
      else { std::vector<float> w((size_t)n_rec * warmup_steps, 0.f); launchFDTD3d(&mesh, &sp, w.data(), interruptNever, progressQuiet); }
      sp.setNumSteps(steps);
    }
    auto t0 = std::chrono::steady_clock::now();
    if (is_double) {
      t_ret = launchFDTD3dDouble(&mesh, &sp, resp.data(), interruptNever, progressQuiet);
    } else {
      std::vector<float> r32((size_t)n_rec * steps, 0.f);
      t_ret = launchFDTD3d(&mesh, &sp, r32.data(), interruptNever, progressQuiet);
      for (size_t i = 0; i < r32.size(); i++) resp[i] = (double)r32[i];
    }
    double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    FILE* o = fopen(argv[2], "wb");
    if (!o) { perror("out"); return 2; }
    uint32_t hdr[9] = {mesh.getDimX(), mesh.getDimY(), mesh.getDimZ(), n_parts, n_rec, steps, is_double,
                       mesh.getNumberOfAirElements(), mesh.getNumberOfBoundaryElements()};
    fwrite("PFDTDOUT", 1, 8, o);
    fwrite(hdr, sizeof(uint32_t), 9, o);
    fwrite(&t_ret, sizeof(double), 1, o);
    fwrite(&wall, sizeof(double), 1, o);
    double wall_e2e = setup_s + wall;   // H2D of the volumes + setupMesh + makePartition + launchFDTD3d (incl. response D2H)
    fwrite(&wall_e2e, sizeof(double), 1, o);
    for (uint32_t k = 0; k < n_parts; k++) {
      uint32_t fs[2] = {mesh.getFirstSliceIdx(k), mesh.getPartitionSize(k)};
      fwrite(fs, sizeof(uint32_t), 2, o);
    }
    fwrite(resp.data(), sizeof(double), resp.size(), o);
    if (dump_nodes) {
      for (uint32_t k = 0; k < n_parts; k++) {
        unsigned n = mesh.getNumberOfElementsAt(k);
        unsigned char* hp = fromDevice<unsigned char>(n, mesh.getPositionIdxPtrAt(k), mesh.getDeviceAt(k));
        unsigned char* hm = fromDevice<unsigned char>(n, mesh.getMaterialIdxPtrAt(k), mesh.getDeviceAt(k));
        fwrite(hp, 1, n, o);
        fwrite(hm, 1, n, o);
        free(hp); free(hm);
      }
    }
    fclose(o);
    double mvox = (double)mesh.getNumberOfElements() * steps / wall / 1e6;
    printf("ref_fdtd: dim %u %u %u parts %u steps %u double %u wall %.6f s (with setup %.6f s) ret_per_step %.6g s  %.1f Mvox/s\n",
           hdr[0], hdr[1], hdr[2], n_parts, steps, is_double, wall, wall_e2e, t_ret, mvox);
  } catch (int e) {
    fprintf(stderr, "reference threw %d\n", e);
    return 4;
  }
  return 0;
}
