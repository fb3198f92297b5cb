// MaterialHandler.h -- per-surface material coefficients, de-duplicated into the unique-material table
// the kernels index with the material byte (reference src/base/MaterialHandler.h:29-161; table layout
// [unique][20], reference src/base/MaterialHandler.cpp:100-164).
#pragma once
#include <vector>

#define MATERIAL_COEF_NUM 20

class MaterialHandler {
 public:
  MaterialHandler();
  ~MaterialHandler() {}

  struct material_t { float coefs[MATERIAL_COEF_NUM]; };

  void addMaterials(float* material_ptr, unsigned int number_of_surfaces, unsigned int number_of_coefficients);
  unsigned int addSurfaceMaterial(std::vector<float> material_coefficients);
  void coefsAreAdmittances() { admitance_ = true; }
  void coefsAreReflectances() { admitance_ = false; }

  // Frequency-dependent boundaries (not in the reference, whose `dif_` members are a stub; PFDTD_OPT_DIF_ORDER in
  // include/pfdtd.h).  With a filter order N in 1..4 each surface row is the digital impedance filter of the surface,
  // [b0 .. bN, a1 .. aN] (a0 = 1) zero-extended to 20 values, and the unique-material table keeps exactly those
  // rows (no reflectance conversion, the octave index is not used).  Order 0 = the reference's scalar admittances.
  void setFilterOrder(unsigned int order);
  unsigned int getFilterOrder() const { return filter_order_; }
  // n_surfaces rows of 2*order+1 coefficients each
  void addFilterMaterials(const float* coefs, unsigned int number_of_surfaces, unsigned int order);
  // the same filter on every surface: b has order+1 taps, a has order taps (a1..)
  void setGlobalFilter(unsigned int number_of_surfaces, const std::vector<float>& b, const std::vector<float>& a);

  unsigned int getNumberOfCoefficients() { return number_of_coefficients_; }
  unsigned int getNumberOfSurfaces() { return number_of_surfaces_; }
  unsigned int getNumberOfUniqueMaterials() { return number_of_unique_materials_; }
  unsigned int getMaterialIdxAt(unsigned int idx) { return material_indices_.at(idx); }
  unsigned char* getMaterialIdxPtr();
  float* getMaterialCoefficientPtr();
  double* getMaterialCoefficientPtrDouble();
  float getUniqueCoefAt(unsigned int material, unsigned int coef_idx);
  float getSurfaceCoefAt(unsigned int surface, unsigned int coef_idx);
  float getMeanAbsorption(unsigned int octave);

  void setNumberOfCoefficients(unsigned int n) { number_of_coefficients_ = n; }
  void setGlobalMaterial(unsigned int number_of_surfaces, float coef);
  void setMaterialIndexAt(unsigned int surface_idx, unsigned char material_idx);

 private:
  unsigned int findOrAdd(const std::vector<float>& coefs);
  bool admitance_;
  unsigned int filter_order_;
  unsigned int number_of_coefficients_;
  unsigned int number_of_surfaces_;
  unsigned int number_of_unique_materials_;
  std::vector<unsigned char> material_indices_;
  std::vector<material_t> unique_coefficients_;
  std::vector<float> coefficient_vector_;
  std::vector<double> coefficient_vector_double_;
};
