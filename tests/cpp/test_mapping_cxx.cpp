// C++ host-side parity test: the scenarios and golden literals of the reference's own unit tests, written against the
// C++ mirror of its mapping interface (include/rbfmap_mapping.hpp) exactly the way the reference's tests drive
// precice::mapping::Mapping - setMeshes, computeMapping, map(Sample, values), clear - and linked against librbfmap.so
// only (no Python, no torch). Sources of the scenarios (relative to /root/reference/src/mapping/tests/):
//   PartitionOfUnityMappingTest.cpp:50-175   perform2DTestConsistentMapping
//   PartitionOfUnityMappingTest.cpp:380-476  perform2DTestConservativeMapping (golden :475)
//   PartitionOfUnityMappingTest.cpp:1102-1197 perform3DTestConsistentMappingVector (golden :1191)
//   RadialBasisFctHelper.hpp:50-140          global mappings reproduce linear functions
//   RadialBasisFctMappingTest.cpp:969-1034   conservative thin-plate-spline literals of the 4-rank test (union meshes)
//   PartitionOfUnityMappingTest.cpp:772-937, 1311-1495 just-in-time read / write mapping (goldens :1423-1424)
// Tolerance: relative 1e-9 like the reference (src/testing/main.cpp:84-87).
// Build + run: tests/test_cxx_host.py (g++ -std=c++17 -Iinclude ... -lrbfmap). Prints "CXX MAPPING OK".
#include <rbfmap_mapping.hpp>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <numeric>

using namespace rbfmap;
using mapping::Mapping;
using mapping::Polynomial;

static int g_failures = 0;

#define CHECK(cond)                                                          \
  do {                                                                       \
    if (!(cond)) {                                                           \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);          \
      ++g_failures;                                                          \
    }                                                                        \
  } while (0)

static bool close(double a, double b, double tol = 1e-9) { return std::abs(a - b) <= tol * std::max({std::abs(a), std::abs(b), 1e-300}) + 1e-13; }
#define CHECK_CLOSE(a, b) CHECK(close((a), (b)))

static mesh::PtrMesh squares2D()
{
  auto m = std::make_shared<mesh::Mesh>("InMesh", 2);
  for (double x0 : {0.0, 2.0, 4.0}) {
    m->createVertex({x0, 0.0});
    m->createVertex({x0 + 1.0, 0.0});
    m->createVertex({x0 + 1.0, 1.0});
    m->createVertex({x0, 1.0});
  }
  return m;
}

static mesh::PtrMesh cubes3D()
{
  auto m = std::make_shared<mesh::Mesh>("InMesh", 3);
  for (double x0 : {0.0, 2.0, 4.0})
    for (double z : {0.0, 1.0}) {
      m->createVertex({x0, 0.0, z});
      m->createVertex({x0 + 1.0, 0.0, z});
      m->createVertex({x0 + 1.0, 1.0, z});
      m->createVertex({x0, 1.0, z});
    }
  return m;
}

// PartitionOfUnityMappingTest.cpp:50-175
static void pumConsistent2D()
{
  mapping::PartitionOfUnityMapping mapping(Mapping::CONSISTENT, 2, mapping::CompactPolynomialC0(3.0), Polynomial::SEPARATE, 5, 0.4, false);
  auto inMesh = squares2D();
  const time::Sample inSample{1, {1, 2, 2, 1, 3, 4, 4, 3, 5, 6, 6, 5}};
  const std::pair<std::array<double, 2>, double> moving[] = {{{0.0, 0.0}, 1.0}, {{0.0, 0.5}, 1.0}, {{0.0, 1.0}, 1.0}, {{1.0, 0.0}, 2.0}, {{1.0, 0.5}, 2.0},
                                                             {{1.0, 1.0}, 2.0}, {{0.5, 0.0}, 1.5}, {{0.5, 0.5}, 1.5}, {{0.5, 1.0}, 1.5}};
  bool first = true;
  for (const auto &[p, expected] : moving) {
    auto outMesh = std::make_shared<mesh::Mesh>("OutMesh", 2);
    outMesh->createVertex({p[0], p[1]});
    outMesh->createVertex({3.5, 0.5});
    outMesh->createVertex({2.5, 0.5});
    outMesh->createVertex({0.0, 0.0});
    outMesh->createVertex({5.0, 1.0});
    mapping.setMeshes(inMesh, outMesh);
    CHECK(mapping.hasComputedMapping() == false);
    mapping.computeMapping();
    CHECK(mapping.hasComputedMapping() == true);
    std::vector<double> outValues;
    mapping.map(inSample, outValues);
    CHECK_CLOSE(outValues[0], expected);
    if (first) {
      CHECK_CLOSE(outValues[1], 4.5);
      CHECK_CLOSE(outValues[2], 3.5);
    }
    first = false;
    mapping.clear(); // PartitionOfUnityMapping.hpp:188: recomputing needs a cleared mapping
  }
}

// PartitionOfUnityMappingTest.cpp:380-476
static void pumConservative2D()
{
  mapping::PartitionOfUnityMapping mapping(Mapping::CONSERVATIVE, 2, mapping::CompactPolynomialC0(3.0), Polynomial::SEPARATE, 5, 0.4, false);
  auto inMesh = std::make_shared<mesh::Mesh>("InMesh", 2);
  inMesh->createVertex({1.0, 0.0});
  inMesh->createVertex({3.5, 0.0});
  inMesh->createVertex({2.5, 0.5});
  inMesh->createVertex({0.0, 0.0});
  inMesh->createVertex({5.0, 1.0});
  auto outMesh = squares2D();
  mapping.setMeshes(inMesh, outMesh);
  mapping.computeMapping();
  std::vector<double> out;

  mapping.map(time::Sample{1, {10, 0, 0, 10, 0}}, out);
  CHECK_CLOSE(out[0], 10.0);
  CHECK_CLOSE(out[1], 10.0);
  CHECK_CLOSE(out[2], 0.0);
  CHECK_CLOSE(out[3], 0.0);
  CHECK_CLOSE(std::accumulate(out.begin(), out.end(), 0.0), 20.0);

  mapping.map(time::Sample{1, {3, 4, 5, 7, 9}}, out);
  CHECK_CLOSE(out[5], 3.2931869752057001); // :475
  CHECK_CLOSE(std::accumulate(out.begin(), out.end(), 0.0), 28.0);
}

// PartitionOfUnityMappingTest.cpp:1102-1197
static void pumConsistentVector3D()
{
  mapping::PartitionOfUnityMapping mapping(Mapping::CONSISTENT, 3, mapping::CompactPolynomialC0(3.0), Polynomial::SEPARATE, 5, 0.265, false);
  auto inMesh  = cubes3D();
  auto outMesh = std::make_shared<mesh::Mesh>("OutMesh", 3);
  outMesh->createVertex({0.0, 0.0, 0.0});
  outMesh->createVertex({3.5, 0.5, 0.5});
  outMesh->createVertex({2.5, 0.5, 1.0});
  outMesh->createVertex({0.0, 0.0, 0.5});
  outMesh->createVertex({5.0, 1.0, 0.5});
  mapping.setMeshes(inMesh, outMesh);
  mapping.computeMapping();
  time::Sample constant{3, {}}, quadratic{3, {}};
  for (int i = 0; i < 24; ++i) {
    constant.values.insert(constant.values.end(), {7.0, 38.0, 19.0});
    const double q = std::pow(i * 3, 2);
    quadratic.values.insert(quadratic.values.end(), {q, 2 * q, 3 * q});
  }
  std::vector<double> out;
  mapping.map(constant, out);
  for (int v : {0, 3}) {
    CHECK_CLOSE(out[3 * v], 7.0);
    CHECK_CLOSE(out[3 * v + 1], 38.0);
    CHECK_CLOSE(out[3 * v + 2], 19.0);
  }
  mapping.map(quadratic, out);
  CHECK_CLOSE(out[12], 3673.7308684337013); // :1191
  CHECK_CLOSE(out[13], 2 * 3673.7308684337013);
  CHECK_CLOSE(out[14], 3 * 3673.7308684337013);
}

// error behaviour: the texts of the reference's PRECICE_CHECKs arrive as rbfmap::Error
static void errors()
{
  mapping::PartitionOfUnityMapping mapping(Mapping::CONSISTENT, 2, mapping::CompactPolynomialC0(3.0), Polynomial::SEPARATE, 5, 0.4, false);
  auto inMesh  = squares2D();
  auto outMesh = std::make_shared<mesh::Mesh>("OutMesh", 2);
  outMesh->createVertex({0.5, 0.5});
  outMesh->createVertex({50.0, 50.0}); // far away from every cluster (PartitionOfUnityMapping.hpp:314-318)
  mapping.setMeshes(inMesh, outMesh);
  std::vector<double> out;
  try {
    mapping.map(time::Sample{1, std::vector<double>(12, 1.0)}, out);
    CHECK(!"map before computeMapping must throw");
  } catch (const Error &e) {
    CHECK(e.status == RBFMAP_ERR_STATE);
  }
  try {
    mapping.computeMapping();
    CHECK(!"unassigned vertex must throw");
  } catch (const Error &e) {
    CHECK(e.status == RBFMAP_ERR_UNASSIGNED);
    CHECK(std::string(e.what()).find("could not be assigned to any cluster in the rbf-pum mapping") != std::string::npos);
  }
  try { // Integrated polynomial is not supported for partition of unity data mappings (PartitionOfUnityMapping.hpp:165)
    mapping::PartitionOfUnityMapping bad(Mapping::CONSISTENT, 2, mapping::CompactPolynomialC0(3.0), Polynomial::ON, 5, 0.4, false);
    CHECK(!"polynomial ON must be rejected");
  } catch (const Error &e) {
    CHECK(e.status == RBFMAP_ERR_INVALID);
  }
}

// RadialBasisFctHelper.hpp:50-140: a global mapping with separated polynomial reproduces linear functions
static void globalLinearReproduction()
{
  for (int iterative = 0; iterative < 2; ++iterative) {
    auto inMesh = std::make_shared<mesh::Mesh>("InMesh", 2), outMesh = std::make_shared<mesh::Mesh>("OutMesh", 2);
    time::Sample in{1, {}};
    for (int i = 0; i < 8; ++i)
      for (int j = 0; j < 8; ++j) {
        const double x = i / 7.0, y = j / 7.0;
        inMesh->createVertex({x, y});
        in.values.push_back(1.0 + 2.0 * x - 3.0 * y);
      }
    for (int i = 0; i < 5; ++i)
      for (int j = 0; j < 5; ++j)
        outMesh->createVertex({0.1 + 0.19 * i, 0.07 + 0.21 * j});
    std::unique_ptr<Mapping> mapping;
    if (iterative)
      mapping = std::make_unique<mapping::IterativeRadialBasisFctMapping>(Mapping::CONSISTENT, 2, mapping::CompactPolynomialC2(0.6), std::array<bool, 3>{false, false, false});
    else
      mapping = std::make_unique<mapping::RadialBasisFctMapping>(Mapping::CONSISTENT, 2, mapping::ThinPlateSplines(), std::array<bool, 3>{false, false, false},
                                                                 Polynomial::SEPARATE);
    mapping->setMeshes(inMesh, outMesh);
    mapping->computeMapping();
    std::vector<double> out, guess;
    if (mapping->requiresInitialGuess())
      mapping->map(in, out, guess); // Mapping::map(input, output, initialGuess)
    else
      mapping->map(in, out);
    for (std::size_t v = 0; v < outMesh->nVertices(); ++v) {
      const auto &c = outMesh->vertexCoords(v);
      CHECK(close(out[v], 1.0 + 2.0 * c[0] - 3.0 * c[1], iterative ? 1e-7 : 1e-9));
    }
    if (iterative) {
      CHECK(guess.size() == static_cast<std::size_t>(rbfmap_initial_guess_size(mapping->handle())));
      std::vector<double> again;
      mapping->map(in, again, guess); // warm start from the stored solution
      for (std::size_t v = 0; v < out.size(); ++v)
        CHECK(close(again[v], out[v], 1e-7));
    }
  }
}

// RadialBasisFctMappingTest.cpp:969-1034 (DistributedConservative2DV4): the literals of the reference's 4-rank gather-scatter
// test are those of the serial mapping between the union meshes
static void globalConservativeGolden()
{
  auto inMesh = std::make_shared<mesh::Mesh>("inMesh", 2), outMesh = std::make_shared<mesh::Mesh>("outMesh", 2);
  for (int x = 0; x < 4; ++x)
    for (int y = 0; y < 2; ++y)
      inMesh->createVertex({double(x), double(y)});
  outMesh->createVertex({0.0, 1.0});
  for (int x = 1; x < 4; ++x)
    for (int y = 0; y < 2; ++y)
      outMesh->createVertex({double(x), double(y)});
  mapping::RadialBasisFctMapping mapping(Mapping::CONSERVATIVE, 2, mapping::ThinPlateSplines(), {false, false, false}, Polynomial::SEPARATE);
  mapping.setMeshes(inMesh, outMesh);
  mapping.computeMapping();
  std::vector<double> out;
  mapping.map(time::Sample{1, {1, 2, 3, 4, 5, 6, 7, 8}}, out);
  const double expected[] = {3.1307987056295525, 4.0734031184906971, 3.0620533966233006, 4.4582564956060873,
                             5.8784343572772633, 7.4683403859032156, 7.9287135404698859}; // :1009-1029
  for (int i = 0; i < 7; ++i)
    CHECK_CLOSE(out[i], expected[i]);
  CHECK_CLOSE(std::accumulate(out.begin(), out.end(), 0.0), 36.0);
}

// Mapping.cpp:175-246: the scaled-consistent mapping preserves the surface integral of the data
static void scaledConsistentSurface()
{
  auto inMesh = std::make_shared<mesh::Mesh>("InMesh", 2), outMesh = std::make_shared<mesh::Mesh>("OutMesh", 2);
  time::Sample in{1, {}};
  const int nIn = 21, nOut = 14;
  for (int i = 0; i < nIn; ++i) {
    const double x = i / double(nIn - 1);
    inMesh->createVertex({x, 0.0});
    in.values.push_back(1.0 + x * x);
    if (i > 0)
      inMesh->createEdge(i - 1, i);
  }
  for (int i = 0; i < nOut; ++i) {
    outMesh->createVertex({i / double(nOut - 1), 0.0});
    if (i > 0)
      outMesh->createEdge(i - 1, i);
  }
  mapping::RadialBasisFctMapping mapping(Mapping::SCALED_CONSISTENT_SURFACE, 2, mapping::CompactPolynomialC2(0.5), {false, true, false}, Polynomial::SEPARATE);
  mapping.setMeshes(inMesh, outMesh);
  mapping.computeMapping();
  std::vector<double> out;
  mapping.map(in, out);
  auto integral = [](const mesh::Mesh &m, const std::vector<double> &v) {
    double s = 0.0;
    for (std::size_t e = 0; e + 1 < m.edges().size(); e += 2) {
      const int a = m.edges()[e], b = m.edges()[e + 1];
      s += 0.5 * (v[a] + v[b]) * std::abs(m.vertexCoords(b)[0] - m.vertexCoords(a)[0]);
    }
    return s;
  };
  CHECK_CLOSE(integral(*inMesh, in.values), integral(*outMesh, out));
}

// PartitionOfUnityMappingTest.cpp:772-937 (read) and :1311-1495 (write): just-in-time mapping
static void justInTime()
{
  const time::Sample values{1, {1, 2, 2, 1, 3, 6, 6, 3, 3, 4, 4, 3, 9, 12, 12, 9, 5, 6, 6, 5, 15, 18, 18, 15}};
  {
    mapping::PartitionOfUnityMapping mapping(Mapping::CONSISTENT, 3, mapping::CompactPolynomialC0(3.0), Polynomial::SEPARATE, 5, 0.265, false);
    mapping.setMeshes(cubes3D(), nullptr); // the output mesh is a just-in-time mesh
    mapping.computeMappingJustInTime();
    CHECK(mapping.hasComputedMapping());
    mapping::MappingDataCache cache(1);
    mapping.initializeMappingDataCache(cache);
    mapping.updateMappingDataCache(cache, values.values);
    std::vector<double> out;
    mapping.mapConsistentAt({{0.0, 0.0, 0.0}}, cache, out);
    CHECK_CLOSE(out[0], 1.0);
    mapping.mapConsistentAt({{3.5, 0.5, 0.5}, {2.5, 0.5, 1.0}}, cache, out);
    CHECK_CLOSE(out[0], 9.0);
    CHECK_CLOSE(out[1], 10.5);
    std::vector<double> doubled = values.values;
    for (double &v : doubled)
      v *= 2.0;
    mapping.updateMappingDataCache(cache, doubled);
    mapping.mapConsistentAt({{0.0, 0.5, 0.5}, {0.0, 1.0, 1.0}, {1.0, 0.0, 0.0}}, cache, out);
    CHECK_CLOSE(out[0], 4.0);
    CHECK_CLOSE(out[1], 6.0);
    CHECK_CLOSE(out[2], 4.0);
  }
  {
    mapping::PartitionOfUnityMapping mapping(Mapping::CONSERVATIVE, 3, mapping::CompactPolynomialC2(3.0), Polynomial::SEPARATE, 5, 0.265, false);
    auto outMesh = cubes3D();
    mapping.setMeshes(nullptr, outMesh); // the input mesh is a just-in-time mesh
    mapping.computeMappingJustInTime();
    mapping::MappingDataCache cache(1);
    mapping.initializeMappingDataCache(cache);
    std::vector<double> result(outMesh->nVertices(), 0.0);
    mapping.mapConservativeAt({{5.0, 0.0, 1.0}}, {4.3}, cache);
    mapping.mapConservativeAt({{3.5, 0.5, 0.5}, {4.5, 1.0, 1.0}, {0.1, 0.0, 1.0}}, {4.3, 17.3, 5.8}, cache);
    mapping.completeJustInTimeMapping(cache, result);
    CHECK_CLOSE(std::accumulate(result.begin(), result.end(), 0.0), 4.3 + 4.3 + 17.3 + 5.8);
    CHECK(close(result[6], 0.063700286667, 1e-7));  // :1423
    CHECK(close(result[23], 9.2744883594, 1e-7));   // :1424
  }
}

int main()
{
  if (rbfmap_device_count() < 1) {
    std::printf("no CUDA device: the C++ host test needs a B200 (there is no CPU fallback)\n");
    return 2;
  }
  try {
    pumConsistent2D();
    pumConservative2D();
    pumConsistentVector3D();
    errors();
    globalLinearReproduction();
    globalConservativeGolden();
    scaledConsistentSurface();
    justInTime();
  } catch (const Error &e) {
    std::printf("FAILED: unexpected rbfmap::Error %d: %s\n", e.status, e.what());
    return 1;
  }
  if (g_failures == 0)
    std::printf("CXX MAPPING OK\n");
  return g_failures == 0 ? 0 : 1;
}
