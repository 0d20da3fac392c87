// Exercises include/basix_b200.hpp (the C++ host mirror of basix::FiniteElement) against SURVEY 8c anchors.
// Usage: host_mirror <element.txt> <host|gpu>.  element.txt is written by tests/test_cpp_mirror.py from a
// golden .npz (cell, polyset, degree, superdegree, map, value_shape, coeffs, entity dofs, transformations).
#include "basix_b200.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

using basix_b200::ElementData;
using basix_b200::FiniteElement;

static ElementData read_element(const char* path)
{
  std::ifstream in(path);
  if (!in)
    throw std::runtime_error(std::string("cannot open ") + path);
  ElementData d;
  std::size_t n;
  in >> d.cell_type >> d.polyset_type >> d.degree >> d.embedded_superdegree >> d.map_type;
  in >> n;
  d.value_shape.resize(n);
  for (auto& v : d.value_shape)
    in >> v;
  in >> n;
  d.coefficient_matrix.resize(n);
  for (auto& v : d.coefficient_matrix)
    in >> v;
  in >> n;
  d.dof_ordering.resize(n);
  for (auto& v : d.dof_ordering)
    in >> v;
  std::size_t ndim;
  in >> ndim;
  d.entity_dofs.resize(ndim);
  for (auto& row : d.entity_dofs)
  {
    in >> n;
    row.resize(n);
    for (auto& e : row)
    {
      std::size_t k;
      in >> k;
      e.resize(k);
      for (auto& v : e)
        in >> v;
    }
  }
  auto read_trans = [&](std::vector<double>& m, int& dim)
  {
    in >> dim >> n;
    m.resize(n);
    for (auto& v : m)
      in >> v;
  };
  read_trans(d.etrans_interval, d.etrans_interval_dim);
  read_trans(d.etrans_triangle, d.etrans_triangle_dim);
  read_trans(d.etrans_quadrilateral, d.etrans_quadrilateral_dim);
  if (!in)
    throw std::runtime_error("truncated element file");
  return d;
}

#define REQUIRE(cond)                                                                                      \
  do                                                                                                       \
  {                                                                                                        \
    if (!(cond))                                                                                           \
    {                                                                                                      \
      std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);                              \
      return 1;                                                                                            \
    }                                                                                                      \
  } while (0)

int main(int argc, char** argv)
{
  if (argc != 3)
    return 2;
  const std::string mode = argv[2];
  FiniteElement e(read_element(argv[1])); // P5 tetrahedron (gll_warped)
  REQUIRE(e.dim() == 56 and e.value_size() == 1);
  REQUIRE(e.dof_transformations_are_permutations() and !e.dof_transformations_are_identity());
  const auto shape = e.tabulate_shape(2, 7);
  REQUIRE(shape[0] == 10 and shape[1] == 7 and shape[2] == 56 and shape[3] == 1);
  REQUIRE(basix_b200::polyset::dim(BX_CELL_TETRAHEDRON, BX_POLYSET_STANDARD, 5) == 56);

  // native create_element: P5 tetrahedron (gll_warped = 2) reproduces the arrays the element was described with
  {
    FiniteElement n = FiniteElement::create(/*P*/ 1, BX_CELL_TETRAHEDRON, 5, /*gll_warped*/ 2);
    REQUIRE(n.dim() == 56 and n.dof_transformations_are_permutations());
    const auto a = n.coefficient_matrix(), b = e.coefficient_matrix();
    REQUIRE(a.size() == b.size());
    double err = 0, scale = 0;
    for (std::size_t i = 0; i < a.size(); ++i)
    {
      err = std::max(err, std::abs(a[i] - b[i]));
      scale = std::max(scale, std::abs(b[i]));
    }
    REQUIRE(err <= 1e-11 * scale);
    bool threw = false;
    try
    {
      FiniteElement::create(/*BDM*/ 4, BX_CELL_TETRAHEDRON, 2);
    }
    catch (const std::runtime_error&)
    {
      threw = true;
    }
    REQUIRE(threw);
  }

  // wrong point dimension: the reference's message (finite-element.cpp:1838-1843)
  const double x2[4] = {0.1, 0.2, 0.3, 0.4};
  try
  {
    e.tabulate(0, x2, 2, 2);
    REQUIRE(false);
  }
  catch (const std::runtime_error& err)
  {
    REQUIRE(std::string(err.what()) == "Point dim (2) does not match element dim (3).");
  }

  const double x[3] = {0.2, 0.3, 0.1};
  std::int32_t d[56];
  for (int i = 0; i < 56; ++i)
    d[i] = i;
  if (mode == "host")
  {
    // no device: every compute call must throw (there is no CPU fallback)
    int thrown = 0;
    try
    {
      e.tabulate(1, x, 1, 3);
    }
    catch (const std::runtime_error&)
    {
      ++thrown;
    }
    try
    {
      e.permute(d, 0x01000);
    }
    catch (const std::runtime_error&)
    {
      ++thrown;
    }
    REQUIRE(bx_device_count() > 0 or thrown == 2);
    std::puts("host mirror ok (host)");
    return 0;
  }

  // ---- GPU: SURVEY 8c anchors ----
  auto [basis, shp] = e.tabulate(0, x, 1, 3);
  double sum = 0;
  for (double v : basis)
    sum += v;
  REQUIRE(std::abs(sum - 1.0) < 1e-12); // Lagrange partition of unity
  e.permute(d, 0x01000);
  REQUIRE(d[4] == 7 and d[5] == 6 and d[6] == 5 and d[7] == 4 and d[8] == 8);
  e.permute_inv(d, 0x01000);
  for (int i = 0; i < 56; ++i)
    REQUIRE(d[i] == i);
  e.permute(d, 0x00001);
  const int want[6] = {28, 31, 33, 29, 32, 30};
  for (int i = 0; i < 6; ++i)
    REQUIRE(d[28 + i] == want[i]);
  double u[56];
  for (int i = 0; i < 56; ++i)
    u[i] = i;
  e.T_apply(u, 56, 1, 0x00001);
  for (int i = 0; i < 6; ++i)
    REQUIRE(u[28 + i] == double(want[i]));
  // sub-entity closure permutations (finite-element.h:860-1035): a triangular face of P5 carries 3 + 3 x 4 + 6 dofs
  REQUIRE(e.subentity_closure_size(BX_CELL_TRIANGLE) == 21 and e.subentity_closure_size(BX_CELL_INTERVAL) == 6);
  std::int32_t cl[21];
  for (int i = 0; i < 21; ++i)
    cl[i] = i;
  e.permute_subentity_closure(cl, 21, /*entity_info=*/3, BX_CELL_TRIANGLE);
  REQUIRE(cl[0] == 1 and cl[1] == 0 and cl[2] == 2); // rotate, then reflect: the vertex blocks
  e.permute_subentity_closure_inv(cl, 21, 3, BX_CELL_TRIANGLE);
  for (int i = 0; i < 21; ++i)
    REQUIRE(cl[i] == i);
  std::puts("host mirror ok (gpu)");
  return 0;
}
