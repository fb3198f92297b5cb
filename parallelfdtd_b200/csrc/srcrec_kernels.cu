// srcrec_kernels.cu -- source injection and receiver capture on the device.
//
// Replaces the per-step, per-source/receiver cudaMemcpy traffic of the reference step loop
// (src/kernels/kernels3d.cu:93-104 sources via CudaMesh::setSample/addSample, cudaMesh.h:321-369;
// :164-173 receivers): source samples are uploaded once as a [n_src][n_steps] table, receiver
// samples accumulate in a [n_rec][n_steps] device buffer.  One tiny launch per step does, in order,
//   (1) record every receiver owned by this partition for step n-1 (the field after update n-1),
//   (2) inject every source that lies in this partition (halo slices included) for step n,
// which is exactly the reference's receiver(n-1) ; source(n) sequence at the step boundary.
// The step index lives in device memory so the same launch can be replayed from a CUDA graph.
#include "pfdtd_internal.h"
#include "tma_common.cuh"

namespace pfdtd {

// One process per GPU with peer-mapped halo stores: the neighbours' edge launches of the previous step write this
// slab's halo planes and then publish their step count (halo_publish).  Sources are injected into halo copies too and
// receivers are read from them, so the wait comes first.  Bounded: a neighbour that never arrives sets HALO_ERR (the
// host turns it into PFDTD_ERR_COMM at the next pfdtd_sync) instead of hanging the device.
__device__ __forceinline__ void halo_wait(int* __restrict__ flags, int from_slot, int want) {
  const long long t0 = clock64();
  int v;
  for (;;) {
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + from_slot) : "memory");
    if (v >= want) return;
    if (clock64() - t0 > (60LL << 30)) { flags[HALO_ERR] = 1; return; }   // ~30 s: ranks may start a run seconds apart
    __nanosleep(64);
  }
}

template <typename T>
__global__ void srcrec_kernel(T* __restrict__ P, int n_rec, const int64_t* __restrict__ rec_elem,
                              const int32_t* __restrict__ rec_slot, T* __restrict__ rec_out, int64_t rec_stride, int n_src,
                              const int64_t* __restrict__ src_elem, const int32_t* __restrict__ src_type,
                              const int32_t* __restrict__ src_slot, const T* __restrict__ src_samples, int64_t src_stride,
                              int* __restrict__ d_step, int do_record, int do_inject, int soft_accumulate, int advance,
                              int* __restrict__ halo_flags, int wait_lo, int wait_hi) {
  if (halo_flags != nullptr) {
    if (threadIdx.x == 0) {
      if (wait_lo) halo_wait(halo_flags, HALO_FROM_LO, halo_flags[HALO_SEQ + 0]);
      if (wait_hi) halo_wait(halo_flags, HALO_FROM_HI, halo_flags[HALO_SEQ + 1]);
    }
    __syncthreads();
  }
  const int n = d_step[0];
  const int first_recordable = d_step[1];   // receivers of steps before the current enqueue are already stored
  if (do_record && n >= 1 && n - 1 >= first_recordable && (int64_t)(n - 1) < rec_stride) {
    for (int r = threadIdx.x; r < n_rec; r += blockDim.x) rec_out[(int64_t)rec_slot[r] * rec_stride + (n - 1)] = P[rec_elem[r]];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (do_inject && (int64_t)n < src_stride) {
      // sequential, in source order, like the reference's host loop (a later source at the same voxel wins)
      for (int s = 0; s < n_src; s++) {
        T v = src_samples[(int64_t)src_slot[s] * src_stride + n];
        if (src_type[s] == PFDTD_SRC_HARD || !soft_accumulate) P[src_elem[s]] = v;
        else P[src_elem[s]] += v;
      }
    }
    if (advance) d_step[0] = n + 1;
  }
}

// The stores of the edge launches into the neighbours' halo planes are complete when those launches are (stream order);
// this one-thread launch right behind them publishes the step: own counter first, then the neighbour's flag word with
// system-scope release.
__global__ void halo_publish_kernel(int* __restrict__ local, int* __restrict__ remote_lo, int* __restrict__ remote_hi) {
  __threadfence_system();
  if (remote_lo != nullptr) {
    const int seq = local[HALO_SEQ + 0] + 1;
    local[HALO_SEQ + 0] = seq;
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(remote_lo), "r"(seq) : "memory");
  }
  if (remote_hi != nullptr) {
    const int seq = local[HALO_SEQ + 1] + 1;
    local[HALO_SEQ + 1] = seq;
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(remote_hi), "r"(seq) : "memory");
  }
}

int launch_halo_publish(int* local_flags, int* remote_lo, int* remote_hi, cudaStream_t stream) {
  halo_publish_kernel<<<1, 1, 0, stream>>>(local_flags, remote_lo, remote_hi);
  PF_CUDA(cudaGetLastError());
  return PFDTD_OK;
}

int launch_srcrec(const SrcRecArgs& a) {
  int threads = a.n_rec > 32 ? (a.n_rec > 256 ? 256 : ((a.n_rec + 31) / 32) * 32) : 32;
  if (a.dtype == PFDTD_F32)
    srcrec_kernel<float><<<1, threads, 0, a.stream>>>((float*)a.P, a.n_rec, a.rec_elem, a.rec_slot, (float*)a.rec_out,
                                                      a.rec_stride, a.n_src, a.src_elem, a.src_type, a.src_slot,
                                                      (const float*)a.src_samples, a.src_stride, a.d_step, a.do_record,
                                                      a.do_inject, a.soft_accumulate, a.advance, const_cast<int*>(a.halo_flags), a.wait_lo,
                                                      a.wait_hi);
  else
    srcrec_kernel<double><<<1, threads, 0, a.stream>>>((double*)a.P, a.n_rec, a.rec_elem, a.rec_slot, (double*)a.rec_out,
                                                       a.rec_stride, a.n_src, a.src_elem, a.src_type, a.src_slot,
                                                       (const double*)a.src_samples, a.src_stride, a.d_step, a.do_record,
                                                       a.do_inject, a.soft_accumulate, a.advance, const_cast<int*>(a.halo_flags), a.wait_lo,
                                                      a.wait_hi);
  PF_CUDA(cudaGetLastError());
  return PFDTD_OK;
}

}  // namespace pfdtd
