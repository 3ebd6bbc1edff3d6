// bulk_gridview_test.cc — the C++ bulk GridView (include/dune_grid_b200/bulkgrid.hh) exercised the way the reference's
// tests exercise a grid view: sizes (test/gridcheck.hh), index-set consecutiveness and sub-index consistency
// (test/checkindexset.hh:76-282), boundary-segment bijection (test/gridcheck.hh:844-925), intersection / neighbour symmetry
// (test/checkintersectionit.hh:275-400), geometry volume sums, and mass conservation of the FV functor.
// Usage: bulk_gridview_test [--device]   (without --device only the host-side descriptor queries run)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <sstream>
#include <numeric>
#include <set>

#include <dune_grid_b200/bulkgrid.hh>

using namespace DuneB200;

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } } while (0)

template<int dim>
void hostChecks() {
  std::array<double,dim> L; std::array<int,dim> s;
  for (int i = 0; i < dim; ++i) { L[i] = 1.0 + i; s[i] = 8 >> (i > 0); }          // 8x4(x4): test/yasp/test-yaspgrid.hh:48-54
  BulkYaspGrid<dim> grid(L, s, std::bitset<dim>(), 1);
  auto gv = grid.leafGridView();
  std::int64_t cells = 1, verts = 1;
  for (int i = 0; i < dim; ++i) { cells *= s[i]; verts *= s[i]+1; }
  CHECK(gv.size(0) == cells);
  CHECK(gv.size(dim) == verts);
  CHECK(gv.size(0, Interior_Partition) == cells);
  CHECK(gv.size(0, Ghost_Partition) == 0);
  CHECK(gv.overlapSize(0) == 1 && gv.ghostSize(0) == 0);
  CHECK(grid.maxLevel() == 0);
  std::int64_t nbs = 0;
  for (int k = 0; k < dim; ++k) { std::int64_t f = 2; for (int l = 0; l < dim; ++l) if (l != k) f *= s[l]; nbs += f; }
  CHECK(grid.numBoundarySegments() == nbs);
  grid.globalRefine(1);
  CHECK(grid.maxLevel() == 1);
  CHECK(grid.leafGridView().size(0) == cells << dim);
  CHECK(grid.levelGridView(0).size(0) == cells);
  bool threw = false;
  try { grid.levelGridView(5); } catch (const RangeError&) { threw = true; }       // yaspgrid.hh:1745
  CHECK(threw);
  BulkMapper<dim> all(gv, [] { std::array<std::uint32_t,4> l{}; for (int c = 0; c <= dim; ++c) l[c] = 1; return l; }());
  std::int64_t total = 0;
  for (int c = 0; c <= dim; ++c) total += gv.size(c);
  CHECK(all.size() == total);
  // two ranks: the default partitioner splits the longest direction (partitioning.hh:56-117)
  BulkYaspGrid<dim> r0(L, s, std::bitset<dim>(), 1, 0, 2), r1(L, s, std::bitset<dim>(), 1, 1, 2);
  CHECK(r0.torusDims()[0] == 2);
  CHECK(r0.leafGridView().size(0, Interior_Partition) + r1.leafGridView().size(0, Interior_Partition) == cells);
  CHECK(r0.leafGridView().size(0) == cells/2 + cells/s[0]);
  threw = false;
  try { BulkYaspGrid<dim> bad(L, s, std::bitset<dim>(), 3, 0, 4); } catch (const GridError&) { threw = true; }   // yaspgrid.hh:925-937
  CHECK(threw);
  // BackupRestoreFacility (yaspgrid/backuprestore.hh): backup -> restore -> same backup text, same level structure
  typedef BackupRestoreFacility<BulkYaspGrid<dim>> BRF;
  std::ostringstream os;
  BRF::backup(grid, os);
  CHECK(os.str().rfind("YaspGrid BackupRestore Format Version: 2", 0) == 0);
  std::istringstream is(os.str());
  std::unique_ptr<BulkYaspGrid<dim>> back(BRF::restore(is, DGB_COORD_EQUIDISTANT));
  CHECK(back && back->maxLevel() == grid.maxLevel());
  CHECK(back->leafGridView().size(0) == grid.leafGridView().size(0));
  CHECK(BRF::text(*back) == os.str());
}

template<int dim>
void deviceChecks() {
  std::array<double,dim> L; std::array<int,dim> s;
  for (int i = 0; i < dim; ++i) { L[i] = 1.0 + 0.5*i; s[i] = 12 - 2*i; }
  std::bitset<dim> periodic; periodic[0] = true;
  BulkYaspGrid<dim> grid(L, s, periodic, 1);
  auto gv = grid.leafGridView();
  // index set: All_Partition visits indices 0,1,2,... (checkindexset.hh), interior cells are Interior, the periodic layer Overlap
  auto el = gv.template elements<0>(All_Partition);
  auto idx = el.index.toHost(); auto pt = el.partitionType.toHost(); auto vol = el.volume.toHost();
  bool consecutive = true; std::int64_t nint = 0; double vsum = 0.0;
  for (std::size_t i = 0; i < idx.size(); ++i) { consecutive = consecutive && idx[i] == i; if (pt[i] == InteriorEntity) { ++nint; vsum += vol[i]; } }
  CHECK(consecutive);
  CHECK(nint == gv.size(0, Interior_Partition));
  double domain = 1.0; for (int i = 0; i < dim; ++i) domain *= L[i];
  CHECK(std::abs(vsum - domain) <= 1e-13*domain);
  // sub-indices: every vertex index of the view is hit by some cell (checkindexset.hh:140-200)
  BulkMapper<dim> vm(gv, mcmgVertexLayout<dim>());
  auto sub = vm.subIndex(dim).toHost();
  std::set<std::uint32_t> seen(sub.begin(), sub.end());
  CHECK(std::int64_t(seen.size()) == gv.size(dim) && *seen.rbegin() == std::uint32_t(gv.size(dim)-1));
  // intersections: outside(outside(e)) == e across every interior face; boundary segments form a bijection
  auto is = gv.intersections(All_Partition);
  auto nb = is.neighbor.toHost(); auto bd = is.boundary.toHost(); auto out = is.outside.toHost(); auto bsi = is.boundarySegmentIndex.toHost();
  const std::int64_t n = is.n;
  bool symmetric = true; std::set<int> segments; std::int64_t nbnd = 0;
  for (int k = 0; k < 2*dim; ++k)
    for (std::int64_t e = 0; e < n; ++e) {
      if (nb[k*n+e]) { const int o = out[k*n+e]; symmetric = symmetric && nb[(k^1)*n+o] && out[(k^1)*n+o] == e; }
      if (bd[k*n+e]) { ++nbnd; segments.insert(bsi[k*n+e]); } else symmetric = symmetric && bsi[k*n+e] == -1;
    }
  CHECK(symmetric);
  CHECK(std::int64_t(segments.size()) == grid.numBoundarySegments() && nbnd == grid.numBoundarySegments());
  CHECK(segments.empty() || (*segments.begin() == 0 && *segments.rbegin() == int(grid.numBoundarySegments())-1));
  // geometry at the cell centre: integrationElement == volume, J^-T diagonal = 1/h
  std::vector<std::array<double,dim>> qp(1); qp[0].fill(0.5);
  auto geo = gv.geometry(qp, All_Partition);
  auto detJ = geo.integrationElement.toHost(); auto glob = geo.global.toHost(); auto cen = el.center.toHost();
  bool same = true;
  for (std::int64_t e = 0; e < n; ++e) { same = same && detJ[e] == vol[e]; for (int d = 0; d < dim; ++d) same = same && glob[d*n+e] == cen[d*n+e]; }
  CHECK(same);
  // communicate: the periodic overlap cells receive the values of their interior images (InteriorBorder_All, forward, set)
  BulkMapper<dim> em(gv, mcmgElementLayout<dim>());
  std::vector<double> h(n);
  for (std::int64_t e = 0; e < n; ++e) h[e] = (pt[e] == InteriorEntity) ? cen[e] : -1.0;     // x coordinate of the centre
  DeviceArray<double> c(n); c.fromHost(h);
  BulkDataHandle<dim> dh{em, {c.data()}, {1}, CommOp::set};
  gv.communicate(dh, InteriorBorder_All_Interface, ForwardCommunication);
  auto after = c.toHost();
  bool filled = true;
  for (std::int64_t e = 0; e < n; ++e) filled = filled && std::abs(after[e] - cen[e]) <= 1e-14;   // wrapped centres equal the images' (up to rounding of i*h + L)
  CHECK(filled);
  // a data handle of variable size (fixedSize() == false): cell e carries (index mod 3) copies of its centre's x coordinate; every
  // overlap cell must receive exactly the items of its interior image
  {
    std::vector<std::int64_t> off(n+1, 0);
    std::vector<double> items;
    for (std::int64_t e = 0; e < n; ++e) { const int k = int(idx[e] % 3); off[e+1] = off[e] + k; for (int j = 0; j < k; ++j) items.push_back(cen[e] + j); }
    DeviceArray<std::int64_t> dOff(n+1); dOff.fromHost(off);
    DeviceArray<double> dItems(items.size()); dItems.fromHost(items);
    BulkVariableDataHandle<dim> vh{em, dOff.data(), dItems.data(), {}, {}, {}};
    gv.communicate(vh, InteriorBorder_All_Interface, ForwardCommunication);
    auto ri = vh.recvIndex.toHost(); auto ro = vh.recvOffsets.toHost(); auto rd = vh.recvData.toHost();
    bool okv = ro.size() == ri.size()+1 && (ro.empty() || std::size_t(ro.back()) == rd.size());
    std::int64_t received = 0;
    for (std::size_t k = 0; k < ri.size() && okv; ++k) {
      const std::int64_t e = ri[k];                               // element mapper: index == position in iteration order
      okv = okv && pt[e] != InteriorEntity;
      for (std::int64_t j = ro[k]; j < ro[k+1]; ++j) okv = okv && std::abs(rd[j] - (cen[e] + double(j-ro[k]))) <= 1e-14;
      ++received;
    }
    CHECK(okv && received == n - gv.size(0, Interior_Partition));
  }
  // GlobalIndexSet on one rank without periodic copies: a permutation of 0..size-1 (utility/globalindexset.hh)
  {
    BulkYaspGrid<dim> plain(L, s, std::bitset<dim>(), 1);
    auto pv = plain.leafGridView();
    for (int codim = 0; codim <= dim; ++codim) {
      GlobalIndexSet<dim> gis(pv, codim);
      auto gi = gis.index().toHost();
      std::set<int> distinct(gi.begin(), gi.end());
      CHECK(std::int64_t(gis.size(codim)) == pv.size(codim) && gis.size((codim+1)%(dim+1)) == 0);
      CHECK(std::int64_t(distinct.size()) == pv.size(codim) && *distinct.begin() == 0 && *distinct.rbegin() == int(pv.size(codim))-1);
    }
    // VTKWriter: cell data = x coordinate of the centre; the file names the interior cells and the field
    auto pel = pv.template elements<0>(All_Partition);
    VTKWriter<dim> vtk(pv);
    vtk.addCellData(pel.center.data(), "xcenter");
    const std::string file = vtk.write("bulk_gridview_test_out", "/tmp");
    std::ifstream in(file);
    std::string all((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    CHECK(file == "/tmp/bulk_gridview_test_out.vtu" && all.find("<CellData Scalars=\"xcenter\">") != std::string::npos);
    CHECK(all.find("NumberOfCells=\"" + std::to_string(pv.size(0)) + "\"") != std::string::npos);
    // write(name, VTK::appendedraw): the arrays move into the <AppendedData> section
    const std::string raw = vtk.write("bulk_gridview_test_raw", VTKWriter<dim>::appendedraw, "/tmp");
    std::ifstream inr(raw, std::ios::binary);
    std::string allr((std::istreambuf_iterator<char>(inr)), std::istreambuf_iterator<char>());
    CHECK(allr.find("Name=\"xcenter\" NumberOfComponents=\"1\" format=\"appended\" offset=\"0\" />") != std::string::npos);
    CHECK(allr.find("<AppendedData encoding=\"raw\">") != std::string::npos);
  }
  // the coordinate container is a template parameter, as in Dune::YaspGrid<dim,Coordinates>; functor sweeps; face normals
  {
    std::array<std::vector<double>,dim> tc;
    for (int i = 0; i < dim; ++i) for (int k = 0; k <= 6+i; ++k) tc[i].push_back(double(k*k)/double((6+i)*(6+i)));
    BulkYaspGrid<dim, TensorProductCoordinates<double,dim>> tg(tc, std::bitset<dim>(), 1);
    auto tv = tg.leafGridView();
    const double vol = tv.forEachElement(DGB_FUNCTOR_MEASURE, {1.0});                    // sum of the cell volumes = |domain| = 1
    CHECK(std::abs(vol - 1.0) <= 1e-14);
    const double flux = tv.forEachIntersection(DGB_FUNCTOR_BOUNDARY_FLUX_X);             // recipe-integration.cc:113-124: = dim*|domain|
    CHECK(std::abs(flux - double(dim)) <= 1e-13);
    const double mid = tv.forEachElement(DGB_FUNCTOR_MIDPOINT_EXP_NORM), gauss = tv.forEachElement(DGB_FUNCTOR_GAUSS5_EXP_NORM);
    CHECK(mid > 1.0 && std::abs(mid - gauss) <= 5e-2*gauss);
    auto nrm = tv.intersectionNormals(All_Partition);
    auto un = nrm.first.toHost();
    const std::int64_t nn = tv.size(0);
    bool okn = true;
    for (int k = 0; k < 2*dim; ++k) for (int d = 0; d < dim; ++d)
      okn = okn && un[std::size_t(k*dim+d)*nn] == ((d == k/2) ? ((k%2) ? 1.0 : -1.0) : 0.0);
    CHECK(okn);
    auto in1 = tv.geometryInInside(1), out1 = tv.geometryInOutside(1);
    CHECK(in1.first[0] == 1.0 && in1.second[0] == 1.0 && out1.first[0] == 0.0 && out1.second[0] == 0.0);
    std::array<double,dim> lowerleft; lowerleft.fill(-1.0);
    BulkYaspGrid<dim, EquidistantOffsetCoordinates<double,dim>> og(lowerleft, L, s, std::bitset<dim>(), 1);
    double expect = 1.0; for (int i = 0; i < dim; ++i) expect *= L[i]+1.0;
    CHECK(std::abs(og.leafGridView().forEachElement(DGB_FUNCTOR_MEASURE, {1.0}) - expect) <= 1e-12*expect);
  }
  // FV functor: three steps conserve mass while the shell is away from the boundary (b = 0), 0 <= c <= 1
  if (dim == 3) {
    std::array<double,dim> L1; L1.fill(1.0); std::array<int,dim> s1; s1.fill(48);
    BulkYaspGrid<dim> g(L1, s1, std::bitset<dim>(), 1);
    FiniteVolume<dim> fv(g, 0, 0, {1.0, 1.0, 1.0, 0.0, 0.5, 0.5, 0.5, 0.1, 0.25});
    const std::int64_t m = g.leafGridView().size(0);
    DeviceArray<double> a(m), b(m);
    fv.initial(a.data());
    auto h0 = a.toHost();
    for (int k = 0; k < 3; ++k) { fv.step(a.data(), b.data()); std::swap(a, b); }
    auto h1 = a.toHost();
    const double m0 = std::accumulate(h0.begin(), h0.end(), 0.0), m1 = std::accumulate(h1.begin(), h1.end(), 0.0);
    CHECK(m0 > 0 && std::abs(m1 - m0) <= 1e-12*m0);
    bool bounded = true; for (double x : h1) bounded = bounded && x >= 0.0 && x <= 1.0;
    CHECK(bounded);
    CHECK(std::abs(fv.lastDt() - 0.99/(3.0*48)) <= 1e-13*(0.99/(3.0*48)));       // h = 1/48 is not dyadic: 1/(x_{i+1}-x_i) varies in the last bits
  }
}

int main(int argc, char** argv) {
  const bool device = argc > 1 && !std::strcmp(argv[1], "--device");
  try {
    hostChecks<2>(); hostChecks<3>();
    if (device) { deviceChecks<2>(); deviceChecks<3>(); }
  } catch (const std::exception& e) { std::printf("exception: %s\n", e.what()); return 2; }
  std::printf("bulk_gridview_test (%s): %s\n", device ? "host+device" : "host", failures ? "FAILED" : "OK");
  return failures ? 1 : 0;
}
