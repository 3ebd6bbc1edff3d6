// dgb_vtk.cu — VTK output of device fields on a bulk YaspGrid view: what
// Dune::VTKWriter<GridView>(gridView, VTK::conforming) + addCellData / addVertexData + write(name, VTK::ascii) produces
// (dune/grid/io/file/vtk/vtkwriter.hh, vtuwriter.hh, pvtuwriter.hh, dataarraywriter.hh), restated:
//   * cells = the InteriorBorder_Partition cells (vtkwriter.hh:116) in iteration order; a vertex is numbered when the walk
//     over cells and their corners (Dune corner order) meets it for the first time (countEntities, :1243-1270)
//   * sections in the order PointData, CellData, Points, Cells (writeAllData :1205-1217); a data section is omitted when it has
//     no field (:1342, :1356); vector fields are padded to 3 components (:1313-1318, :1331-1334); coordinates too (:1379-1382)
//   * connectivity in VTK corner order (VTK::renumber, common.hh:186-199: quads 0,1,3,2; hexahedra 0,1,3,2,4,5,7,6), offsets,
//     types (quadrilateral 9, hexahedron 12; common.hh:151-169)
//   * ASCII arrays: 12 values per line, two-space indentation per level, setprecision(digits10 of the stored type)
//     (dataarraywriter.hh:109-190)
//   * file names: serial <name>.vtu (:943-950); parallel s####-p####-<name>.vtu and s####-<name>.pvtu (:858-909)
// The upstream writer cannot be built here (it needs dune-common's indent / path / iteratorfacades headers and
// dune-geometry), so there is no reference run to compare with: parity of this component is UNPINNED; tests check the
// format rules above and that the file describes the grid and the fields (tests/test_vtk.py).
// Host-side work by nature (text formatting); the device is only read: fields are copied D2H once.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <limits>
#include <sstream>
#include "dgb_internal.hh"

namespace dgb {

struct VtkIndent { int level = 0; };
static std::ostream& operator<<(std::ostream& s, const VtkIndent& i) { for (int k = 0; k < i.level; ++k) s << "  "; return s; }

// AsciiDataArrayWriter (dataarraywriter.hh:109-190)
class VtkArray {
  std::ostream& s; VtkIndent indent; int counter = 0; int prec;        // prec: 0 Float32, 1 Float64, 2 Int32, 3 UInt8
public:
  VtkArray(std::ostream& s_, const std::string& name, int ncomps, VtkIndent ind, int prec_) : s(s_), indent(ind), prec(prec_) {
    static const char* types[4] = { "Float32", "Float64", "Int32", "UInt8" };
    s << indent << "<DataArray type=\"" << types[prec] << "\" Name=\"" << name << "\" NumberOfComponents=\"" << ncomps << "\" format=\"ascii\">\n";
    ++indent.level;
  }
  ~VtkArray() { if (counter % 12 != 0) s << "\n"; --indent.level; s << indent << "</DataArray>\n"; }
  void lead() { if (counter % 12 == 0) s << indent; else s << " "; }
  void trail() { ++counter; if (counter % 12 == 0) s << "\n"; }
  void write(double x) {
    lead();
    if (prec == 0) { float f = float(x); if (std::fpclassify(f) == FP_SUBNORMAL) f = 0; s << std::setprecision(std::numeric_limits<float>::digits10) << f; }
    else if (prec == 1) { if (std::fpclassify(x) == FP_SUBNORMAL) x = 0; s << std::setprecision(std::numeric_limits<double>::digits10) << x; }
    else if (prec == 2) s << std::setprecision(std::numeric_limits<std::int32_t>::digits10) << std::int32_t(x);
    else s << std::setprecision(std::numeric_limits<unsigned>::digits10) << unsigned(std::uint8_t(x));
    trail();
  }
};

static std::string parallel_name(const std::string& name, int rank, int size) {      // getParallelPieceName, vtkwriter.hh:858-909
  std::ostringstream s;
  s << 's' << std::setw(4) << std::setfill('0') << size << '-';
  if (rank >= 0) s << 'p' << std::setw(4) << std::setfill('0') << rank << '-';
  s << name << "." << (rank < 0 ? "p" : "") << "vtu";
  return s.str();
}
static std::string with_path(const std::string& path, const std::string& file) {      // concatPaths for the simple cases
  if (path.empty()) return file;
  return path.back() == '/' ? path + file : path + "/" + file;
}
static void first_names(const dgb_vtk_field* f, int n, std::string& scalars, std::string& vectors) {   // getDataNames :1273-1296
  for (int i = 0; i < n; ++i) if (!f[i].is_vector) { scalars = f[i].name; break; }
  for (int i = 0; i < n; ++i) if (f[i].is_vector) { vectors = f[i].name; break; }
}

} // namespace dgb

using namespace dgb;

extern "C" int dgb_vtk_write(const dgb_grid* g, int level, const char* name, const char* path, const dgb_vtk_field* cell_fields,
                             int ncell_fields, const dgb_vtk_field* vertex_fields, int nvertex_fields, int coord_precision,
                             char* written, int cap, void* stream_) {
  return guarded([&]() -> int {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!g || !name) throw ArgError("null argument");
    if (level < 0 || level >= int(g->me.levels.size())) throw RangeError("level out of range");
    if (coord_precision < 0 || coord_precision > 1) throw ArgError("coordinate precision: 0 (Float32) or 1 (Float64)");
    const Level& L = g->me.levels[level];
    const LevelBoxes& b = L.b;
    const int dim = b.dim, ncorner = 1 << dim;
    const dgb_component& cellc = b.comp[b.comp_begin[0]];
    const dgb_component& vertc = b.comp[b.comp_begin[dim]];
    const dgb_box& in = cellc.box[DGB_BOX_INTERIORBORDER];            // cells of the InteriorBorder partition (= interior cells)
    const dgb_box& cof = cellc.box[DGB_BOX_OVERLAPFRONT];
    const dgb_box& vof = vertc.box[DGB_BOX_OVERLAPFRONT];
    long long ncells = 1, nallcells = 1, nallverts = 1;
    for (int i = 0; i < dim; ++i) { ncells *= in.size[i]; nallcells *= cof.size[i]; nallverts *= vof.size[i]; }
    if (box_empty(in, dim)) ncells = 0;
    // fields: device -> host (arrays are indexed by the element / vertex index of the level, ncomp values per entity)
    auto fetch = [&](const dgb_vtk_field* f, int n, long long entities, std::vector<std::vector<double>>& host) -> int {
      host.resize(n);
      for (int k = 0; k < n; ++k) {
        if (!f[k].name || !f[k].d_data || f[k].ncomp < 1) throw ArgError("vtk field: name, data and ncomp >= 1 are required");
        if (f[k].is_vector && f[k].ncomp > 3) throw ArgError("Cannot write VTK vectors with more than 3 components");   // :1315
        if (f[k].precision < 0 || f[k].precision > 1) throw ArgError("vtk field precision: 0 (Float32) or 1 (Float64)");
        host[k].resize(std::size_t(entities)*f[k].ncomp);
        DGB_CUDA(cudaMemcpyAsync(host[k].data(), f[k].d_data, host[k].size()*sizeof(double), cudaMemcpyDeviceToHost, stream));
      }
      return DGB_OK;
    };
    std::vector<std::vector<double>> hcell, hvert;
    if (ncell_fields || nvertex_fields) if (int rc = require_device()) return rc;
    if (int rc = fetch(cell_fields, ncell_fields, nallcells, hcell)) return rc;
    if (int rc = fetch(vertex_fields, nvertex_fields, nallverts, hvert)) return rc;
    if (ncell_fields || nvertex_fields) DGB_CUDA(cudaStreamSynchronize(stream));

    // countEntities (:1243-1270): walk cells and corners, number vertices at the first visit
    std::vector<int> number(std::size_t(nallverts), -1);
    std::vector<long long> visitVertex;                  // vertex index in visiting order
    std::vector<int> connectivity; connectivity.reserve(std::size_t(ncells)*ncorner);
    std::vector<long long> cellIndex; cellIndex.reserve(std::size_t(ncells));
    std::vector<double> points;                          // 3 per visited vertex
    static const int quadRenumbering[4] = {0,1,3,2}, cubeRenumbering[8] = {0,1,3,2,4,5,7,6};      // common.hh:186-199
    int c[3] = {0, 0, 0};
    for (long long t = 0; t < ncells; ++t) {
      long long q = t;
      for (int i = 0; i < dim; ++i) { c[i] = in.origin[i] + int(q % in.size[i]); q /= in.size[i]; }
      long long ci = 0, stride = 1;
      for (int i = 0; i < dim; ++i) { ci += (c[i]-cof.origin[i])*stride; stride *= cof.size[i]; }
      cellIndex.push_back(cellc.index_offset[DGB_BOX_OVERLAPFRONT] + ci);
      long long vid[8];
      for (int k = 0; k < ncorner; ++k) {
        long long vi = 0; stride = 1;
        for (int i = 0; i < dim; ++i) { vi += (c[i] + (k >> i & 1) - vof.origin[i])*stride; stride *= vof.size[i]; }
        vid[k] = vi;
        if (number[vi] < 0) {
          number[vi] = int(visitVertex.size());
          visitVertex.push_back(vi);
          for (int i = 0; i < 3; ++i) points.push_back(i < dim ? L.coords.coordinate(i, c[i] + (k >> i & 1)) : 0.0);   // :1379-1382
        }
      }
      for (int k = 0; k < ncorner; ++k) connectivity.push_back(number[vid[dim == 2 ? quadRenumbering[k] : cubeRenumbering[k]]]);
    }
    const long long nvertices = (long long)visitVertex.size();

    const int rank = g->me.rank, size = g->me.nranks;
    const std::string dir = path ? path : "";
    const std::string piece = with_path(dir, size > 1 ? parallel_name(name, rank, size) : std::string(name) + ".vtu");
    {
      std::vector<char> buffer(1 << 22);                 // declared before the stream: it must outlive the filebuf that uses it
      std::ofstream file;
      file.rdbuf()->pubsetbuf(buffer.data(), std::streamsize(buffer.size()));
      file.open(piece.c_str(), std::ios::binary);
      if (!file.is_open()) throw ArgError("Could not write to piece file " + piece);
      VtkIndent indent;
      file << indent << "<?xml version=\"1.0\"?>\n";
      file << indent << "<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n";
      ++indent.level;
      file << indent << "<UnstructuredGrid>\n"; ++indent.level;
      file << indent << "<Piece NumberOfCells=\"" << unsigned(ncells) << "\" NumberOfPoints=\"" << unsigned(nvertices) << "\">\n"; ++indent.level;
      auto write_fields = [&](const char* section, const dgb_vtk_field* f, int n, const std::vector<std::vector<double>>& host,
                              const std::vector<long long>& order, long long indexBase) {
        if (n == 0) return;
        std::string scalars, vectors; first_names(f, n, scalars, vectors);
        file << indent << "<" << section;
        if (scalars != "") file << " Scalars=\"" << scalars << "\"";
        if (vectors != "") file << " Vectors=\"" << vectors << "\"";
        file << ">\n"; ++indent.level;
        for (int k = 0; k < n; ++k) {
          const int writecomps = f[k].is_vector ? 3 : f[k].ncomp;
          VtkArray a(file, f[k].name, writecomps, indent, f[k].precision);
          for (long long e : order) {
            const double* v = host[k].data() + std::size_t(e - indexBase)*f[k].ncomp;
            for (int j = 0; j < f[k].ncomp; ++j) a.write(v[j]);
            for (int j = f[k].ncomp; j < writecomps; ++j) a.write(0.0);
          }
        }
        --indent.level; file << indent << "</" << section << ">\n";
      };
      write_fields("PointData", vertex_fields, nvertex_fields, hvert, visitVertex, 0);
      write_fields("CellData", cell_fields, ncell_fields, hcell, cellIndex, cellc.index_offset[DGB_BOX_OVERLAPFRONT]);
      file << indent << "<Points>\n"; ++indent.level;
      { VtkArray a(file, "Coordinates", 3, indent, coord_precision); for (double x : points) a.write(x); }
      --indent.level; file << indent << "</Points>\n";
      file << indent << "<Cells>\n"; ++indent.level;
      { VtkArray a(file, "connectivity", 1, indent, 2); for (int v : connectivity) a.write(v); }
      { VtkArray a(file, "offsets", 1, indent, 2); int off = 0; for (long long t = 0; t < ncells; ++t) { off += ncorner; a.write(off); } }
      { VtkArray a(file, "types", 1, indent, 3); for (long long t = 0; t < ncells; ++t) a.write(dim == 2 ? 9 : 12); }
      --indent.level; file << indent << "</Cells>\n";
      --indent.level; file << indent << "</Piece>\n";
      --indent.level; file << indent << "</UnstructuredGrid>\n";
      --indent.level; file << indent << "</VTKFile>\n" << std::flush;
      if (!file) throw ArgError("write failed on " + piece);
    }
    std::string result = piece;
    if (size > 1) {                                               // writeParallelHeader :1111-1174, by rank 0 (:1071-1085)
      const std::string header = with_path(dir, parallel_name(name, -1, size));
      result = header;
      if (rank == 0) {
        std::ofstream file(header.c_str(), std::ios::binary);
        if (!file.is_open()) throw ArgError("Could not write to parallel file " + header);
        VtkIndent indent;
        file << indent << "<?xml version=\"1.0\"?>\n";
        file << indent << "<VTKFile type=\"PUnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n"; ++indent.level;
        file << indent << "<PUnstructuredGrid GhostLevel=\"0\">\n"; ++indent.level;
        auto header_fields = [&](const char* section, const dgb_vtk_field* f, int n) {
          std::string scalars, vectors; first_names(f, n, scalars, vectors);
          file << indent << "<" << section;
          if (scalars != "") file << " Scalars=\"" << scalars << "\"";
          if (vectors != "") file << " Vectors=\"" << vectors << "\"";
          file << ">\n"; ++indent.level;
          for (int k = 0; k < n; ++k)
            file << indent << "<PDataArray type=\"" << (f[k].precision ? "Float64" : "Float32") << "\" Name=\"" << f[k].name
                 << "\" NumberOfComponents=\"" << (f[k].is_vector ? 3 : f[k].ncomp) << "\"/>\n";
          --indent.level; file << indent << "</" << section << ">\n";
        };
        header_fields("PPointData", vertex_fields, nvertex_fields);
        header_fields("PCellData", cell_fields, ncell_fields);
        file << indent << "<PPoints>\n"; ++indent.level;
        file << indent << "<PDataArray type=\"" << (coord_precision ? "Float64" : "Float32") << "\" Name=\"Coordinates\" NumberOfComponents=\"3\"/>\n";
        --indent.level; file << indent << "</PPoints>\n";
        for (int r = 0; r < size; ++r) file << indent << "<Piece  Source=\"" << parallel_name(name, r, size) << "\"/>\n";
        --indent.level; file << indent << "</PUnstructuredGrid>\n";
        --indent.level; file << indent << "</VTKFile>\n" << std::flush;
      }
    }
    if (written && cap > 0) { std::strncpy(written, result.c_str(), std::size_t(cap)-1); written[cap-1] = '\0'; }
    return DGB_OK;
  });
}
