// update_host.cuh -- host-side construction of the per-launch constants shared by the update kernels
#pragma once
#include <cmath>
#include "pfdtd_internal.h"
#include "update_math.cuh"

namespace pfdtd {

template <typename T>
inline UpdConst<T> make_const(const UpdateArgs& a) {
  UpdConst<T> c;
  for (int i = 0; i < 4; i++) c.d[i] = (T)a.dcoef[i];
  c.lam = (T)a.params[0];
  c.lam2 = (T)a.params[1];
  c.octave = (T)a.params[3];
  // computed on the host with the same single-rounding fma the device would use
  if (a.scheme == SCH_CENTRED) c.a_air = (T)std::fma((T)a.params[1], (T)-6, (T)2);
  else c.a_air = (T)std::fma((T)6, -(T)a.params[1], (T)2);
  c.materials = (const T*)a.materials;
  c.n_coefs = a.n_coefs;
  c.matidx_as_written = a.matidx_as_written;
  c.dif_order = a.dif_order;
  return c;
}

template <typename T>
inline DifArgs<T> make_dif(const UpdateArgs& a) {
  DifArgs<T> d;
  d.state = (T*)a.dif_state;
  d.rowbase = (const uint2*)a.dif_rowbase;
  d.table = (const DifEntry<T>*)a.dif_table;
  d.nb = a.dif_nb;
  d.order = a.dif_order;
  d.dif_lo = a.dif_lo;
  d.n_dif = a.n_dif;
  d.segs = (a.X + 127) / 128;
  d.wide_mat = a.wide ? a.mat : nullptr;
  d.wide_table = (const DifEntry<T>*)a.wide_dif_table;
  d.n_mat = a.n_coefs / 20;
  return d;
}

template <typename T>
inline WideArgs<T> make_wide(const UpdateArgs& a) {
  WideArgs<T> w;
  w.mat = a.wide ? a.mat : nullptr;
  w.table = (const ClassEntry<T>*)a.wide_class_table;
  w.lo = a.dif_lo;
  w.n_mat = a.n_coefs / 20;
  return w;
}


}  // namespace pfdtd
