// MaterialHandler.cpp -- see the header.
#include "MaterialHandler.h"

#include <stdexcept>

#include "../global_includes.h"

MaterialHandler::MaterialHandler()
    : admitance_(true), filter_order_(0), number_of_coefficients_(MATERIAL_COEF_NUM), number_of_surfaces_(0), number_of_unique_materials_(0) {}

void MaterialHandler::addMaterials(float* material_ptr, unsigned int number_of_surfaces, unsigned int number_of_coefficients) {
  // rows shorter than 20 coefficients are zero-extended (reference MaterialHandler.cpp:27-45)
  std::vector<float> row(MATERIAL_COEF_NUM, 0.f);
  for (unsigned int i = 0; i < number_of_surfaces; i++) {
    for (unsigned int j = 0; j < number_of_coefficients; j++) row.at(j) = material_ptr[(size_t)i * number_of_coefficients + j];
    addSurfaceMaterial(row);
  }
}

void MaterialHandler::setFilterOrder(unsigned int order) {
  if (order > 4) throw std::out_of_range("MaterialHandler::setFilterOrder: filter order > 4");
  if (order != filter_order_ && number_of_surfaces_ != 0)
    throw std::logic_error("MaterialHandler::setFilterOrder: surfaces were already added with another filter order");
  filter_order_ = order;
  if (order) admitance_ = true;                                  // filter rows are taken as they are
}

void MaterialHandler::addFilterMaterials(const float* coefs, unsigned int number_of_surfaces, unsigned int order) {
  if (order == 0) throw std::out_of_range("MaterialHandler::addFilterMaterials: filter order 0 (use addMaterials)");
  setFilterOrder(order);
  const unsigned int per_row = 2 * order + 1;
  std::vector<float> row(MATERIAL_COEF_NUM, 0.f);
  for (unsigned int i = 0; i < number_of_surfaces; i++) {
    for (unsigned int j = 0; j < per_row; j++) row[j] = coefs[(size_t)i * per_row + j];
    addSurfaceMaterial(row);
  }
}

void MaterialHandler::setGlobalFilter(unsigned int number_of_surfaces, const std::vector<float>& b, const std::vector<float>& a) {
  if (b.size() < 2 || a.size() + 1 != b.size()) throw std::out_of_range("MaterialHandler::setGlobalFilter: need order+1 b taps and order a taps");
  setFilterOrder((unsigned int)a.size());
  std::vector<float> row(MATERIAL_COEF_NUM, 0.f);
  for (size_t j = 0; j < b.size(); j++) row[j] = b[j];
  for (size_t j = 0; j < a.size(); j++) row[b.size() + j] = a[j];
  for (unsigned int i = 0; i < number_of_surfaces; i++) addSurfaceMaterial(row);
}

unsigned int MaterialHandler::addSurfaceMaterial(std::vector<float> material_coefficients) {
  if (material_coefficients.size() != number_of_coefficients_)
    log_msg<LOG_ERROR>(L"MaterialHandler::addSurfaceMaterial - %d coefficients given, %d expected") % material_coefficients.size() %
        number_of_coefficients_;
  material_coefficients.resize(material_coefficients.size() > MATERIAL_COEF_NUM ? material_coefficients.size() : MATERIAL_COEF_NUM, 0.f);
  const unsigned int idx = findOrAdd(material_coefficients);
  material_indices_.push_back((unsigned char)idx);
  number_of_surfaces_++;
  return idx;
}

unsigned int MaterialHandler::findOrAdd(const std::vector<float>& coefs) {
  for (unsigned int m = 0; m < unique_coefficients_.size(); m++) {
    bool same = true;
    for (unsigned int k = 0; k < MATERIAL_COEF_NUM && same; k++) same = coefs[k] == unique_coefficients_[m].coefs[k];
    if (same) return m;
  }
  material_t nm;
  for (unsigned int k = 0; k < MATERIAL_COEF_NUM; k++) nm.coefs[k] = coefs[k];
  unique_coefficients_.push_back(nm);
  number_of_unique_materials_++;
  return (unsigned int)unique_coefficients_.size() - 1;
}

float MaterialHandler::getUniqueCoefAt(unsigned int material, unsigned int coef_idx) {
  if (coef_idx >= (unsigned int)MATERIAL_COEF_NUM) throw std::out_of_range("Coef idx out of range");
  return unique_coefficients_.at(material).coefs[coef_idx];
}

float MaterialHandler::getSurfaceCoefAt(unsigned int surface, unsigned int coef_idx) {
  if (coef_idx >= (unsigned int)MATERIAL_COEF_NUM) throw std::out_of_range("Coef idx out of range");
  return getUniqueCoefAt(material_indices_.at(surface), coef_idx);
}

unsigned char* MaterialHandler::getMaterialIdxPtr() { return material_indices_.empty() ? (unsigned char*)0 : &material_indices_[0]; }

// flat [unique][number_of_coefficients_] table; reflectances are converted with reflection2Admitance in
// float, the double table is the widened float value (reference :100-164).  The kernels index rows with a
// stride of 20, so number_of_coefficients_ must stay 20 for the table to be usable (SURVEY C-9).
float* MaterialHandler::getMaterialCoefficientPtr() {
  const unsigned int coefs = getNumberOfCoefficients(), um = getNumberOfUniqueMaterials();
  coefficient_vector_.assign((size_t)coefs * um, 0.f);
  for (unsigned int i = 0; i < um; i++)
    for (unsigned int j = 0; j < coefs; j++) {
      float c = getUniqueCoefAt(i, j);
      coefficient_vector_[(size_t)i * coefs + j] = admitance_ ? c : reflection2Admitance(c);
    }
  return coefficient_vector_.empty() ? (float*)0 : &coefficient_vector_[0];
}

double* MaterialHandler::getMaterialCoefficientPtrDouble() {
  const unsigned int coefs = getNumberOfCoefficients(), um = getNumberOfUniqueMaterials();
  coefficient_vector_double_.assign((size_t)coefs * um, 0.0);
  for (unsigned int i = 0; i < um; i++)
    for (unsigned int j = 0; j < coefs; j++) {
      float c = getUniqueCoefAt(i, j);
      coefficient_vector_double_[(size_t)i * coefs + j] = admitance_ ? (double)c : (double)reflection2Admitance(c);
    }
  return coefficient_vector_double_.empty() ? (double*)0 : &coefficient_vector_double_[0];
}

float MaterialHandler::getMeanAbsorption(unsigned int octave) {
  float sum = 0.f;
  for (unsigned int i = 0; i < number_of_surfaces_; i++) {
    const float r = admitance2Reflection(getSurfaceCoefAt(i, octave));
    sum += 1 - r * r;
  }
  return sum / number_of_surfaces_;
}

void MaterialHandler::setGlobalMaterial(unsigned int number_of_surfaces, float coef) {
  std::vector<float> row(MATERIAL_COEF_NUM, coef);
  for (unsigned int i = 0; i < number_of_surfaces; i++) addSurfaceMaterial(row);
}

void MaterialHandler::setMaterialIndexAt(unsigned int surface_idx, unsigned char material_idx) {
  if (surface_idx >= number_of_surfaces_) return;   // out of bounds: ignored, like the reference (:218-226)
  material_indices_.at(surface_idx) = material_idx;
}
