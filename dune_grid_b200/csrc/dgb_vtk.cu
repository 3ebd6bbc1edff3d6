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

// ---- binary encodings (dataarraywriter.hh:196-420, streams.hh, b64enc.hh) -------------------------------------------------
// base64: every DataArray carries base64(uint32 byte count), flushed (8 characters), then base64(data), flushed. appendedraw /
// appendedbase64: the tag only names an offset into the <AppendedData> section, where the same header + data follow per array.
static const char kB64[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
struct B64Sink {                                       // Base64Stream: 3 bytes -> 4 characters, '=' padding on flush
  std::ostream& s; unsigned char t[3]; int n = 0;
  explicit B64Sink(std::ostream& s_) : s(s_) {}
  void put(unsigned char c) { t[n++] = c; if (n == 3) emit(); }
  void write(const void* p, std::size_t bytes) { const unsigned char* q = static_cast<const unsigned char*>(p); for (std::size_t i = 0; i < bytes; ++i) put(q[i]); }
  void emit() {
    const unsigned a = t[0], b = n > 1 ? t[1] : 0u, c = n > 2 ? t[2] : 0u;
    char o[4] = { kB64[a >> 2], kB64[(a & 3u) << 4 | b >> 4], n > 1 ? kB64[(b & 15u) << 2 | c >> 6] : '=', n > 2 ? kB64[c & 63u] : '=' };
    s.write(o, 4); n = 0;
  }
  void flush() { if (n > 0) emit(); }
};
static std::size_t vtk_type_size(int prec) { return prec == 0 ? 4 : prec == 1 ? 8 : prec == 2 ? 4 : 1; }
// one value in its file type (DataArrayWriter::write<T>: the static_cast of the virtual writeFloat32 / writeFloat64 / writeInt32 / writeUInt8)
template<class Sink> static void vtk_put_value(Sink& sink, int prec, double x) {
  if (prec == 0) { const float f = float(x); sink.write(&f, 4); }
  else if (prec == 1) sink.write(&x, 8);
  else if (prec == 2) { const std::int32_t i = std::int32_t(x); sink.write(&i, 4); }
  else { const std::uint8_t u = std::uint8_t(x); sink.write(&u, 1); }
}
struct RawSink { std::ostream& s; void write(const void* p, std::size_t bytes) { s.write(static_cast<const char*>(p), std::streamsize(bytes)); } };

// one array of the piece file, in file order
struct VtkArrayData { int section; std::string name; int ncomps; long long nitems; int prec; std::vector<double> v; };

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

extern "C" int dgb_vtk_write_ex(const dgb_grid* g, int level, const char* name, const char* path, const dgb_vtk_field* cell_fields,
                                int ncell_fields, const dgb_vtk_field* vertex_fields, int nvertex_fields, int coord_precision,
                                int outputtype, char* written, int cap, void* stream_);
extern "C" int dgb_vtk_write(const dgb_grid* g, int level, const char* name, const char* path, const dgb_vtk_field* cell_fields,
                             int ncell_fields, const dgb_vtk_field* vertex_fields, int nvertex_fields, int coord_precision,
                             char* written, int cap, void* stream_) {
  return dgb_vtk_write_ex(g, level, name, path, cell_fields, ncell_fields, vertex_fields, nvertex_fields, coord_precision, DGB_VTK_ASCII,
                          written, cap, stream_);
}
extern "C" int dgb_vtk_write_ex(const dgb_grid* g, int level, const char* name, const char* path, const dgb_vtk_field* cell_fields,
                                int ncell_fields, const dgb_vtk_field* vertex_fields, int nvertex_fields, int coord_precision,
                                int outputtype, char* written, int cap, void* stream_) {
  return guarded([&]() -> int {
    if (outputtype < DGB_VTK_ASCII || outputtype > DGB_VTK_APPENDEDBASE64) throw ArgError("unknown VTK output type");
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
      // the arrays of the piece in file order (VTKWriter::writeAllData :1205-1217): PointData, CellData, Points, Cells
      std::vector<VtkArrayData> arrays;
      auto add_fields = [&](int section, const dgb_vtk_field* f, int n, const std::vector<std::vector<double>>& host,
                            const std::vector<long long>& order, long long indexBase) {
        for (int k = 0; k < n; ++k) {
          const int writecomps = f[k].is_vector ? 3 : f[k].ncomp;
          VtkArrayData a{section, f[k].name, writecomps, (long long)order.size(), f[k].precision, {}};
          a.v.reserve(order.size()*writecomps);
          for (long long e : order) {
            const double* v = host[k].data() + std::size_t(e - indexBase)*f[k].ncomp;
            for (int j = 0; j < f[k].ncomp; ++j) a.v.push_back(v[j]);
            for (int j = f[k].ncomp; j < writecomps; ++j) a.v.push_back(0.0);
          }
          arrays.push_back(std::move(a));
        }
      };
      add_fields(0, vertex_fields, nvertex_fields, hvert, visitVertex, 0);
      add_fields(1, cell_fields, ncell_fields, hcell, cellIndex, cellc.index_offset[DGB_BOX_OVERLAPFRONT]);
      arrays.push_back(VtkArrayData{2, "Coordinates", 3, nvertices, coord_precision, points});
      { VtkArrayData a{3, "connectivity", 1, (long long)connectivity.size(), 2, {}}; a.v.assign(connectivity.begin(), connectivity.end()); arrays.push_back(std::move(a)); }
      { VtkArrayData a{3, "offsets", 1, ncells, 2, {}}; int off = 0; for (long long t = 0; t < ncells; ++t) { off += ncorner; a.v.push_back(off); } arrays.push_back(std::move(a)); }
      { VtkArrayData a{3, "types", 1, ncells, 3, {}}; a.v.assign(std::size_t(ncells), dim == 2 ? 9.0 : 12.0); arrays.push_back(std::move(a)); }

      static const char* types[4] = { "Float32", "Float64", "Int32", "UInt8" };
      VtkIndent indent;
      unsigned offset = 0;                                  // running offset of the appended output types
      // one DataArray tag (+ inline data) of the main section
      auto main_array = [&](const VtkArrayData& a) {
        if (outputtype == DGB_VTK_ASCII) { VtkArray w(file, a.name, a.ncomps, indent, a.prec); for (double x : a.v) w.write(x); return; }
        file << indent << "<DataArray type=\"" << types[a.prec] << "\" Name=\"" << a.name << "\" NumberOfComponents=\"" << a.ncomps << "\" ";
        const std::size_t bytes = std::size_t(a.ncomps)*std::size_t(a.nitems)*vtk_type_size(a.prec);
        if (outputtype == DGB_VTK_BASE64) {
          file << "format=\"binary\">\n";
          VtkIndent in1 = indent; ++in1.level;
          file << in1;
          B64Sink b(file);
          const std::uint32_t size = std::uint32_t(bytes);
          b.write(&size, 4); b.flush();
          for (double x : a.v) vtk_put_value(b, a.prec, x);
          b.flush();
          file << "\n" << indent << "</DataArray>\n";
        } else {
          file << "format=\"appended\" offset=\"" << offset << "\" />\n";
          if (outputtype == DGB_VTK_APPENDEDRAW) offset += 4 + unsigned(bytes);
          else { offset += 8 + unsigned(bytes/3*4); if (bytes % 3 != 0) offset += 4; }
        }
      };
      auto section_of = [&](int sec, const char* tag, const dgb_vtk_field* f, int n) {
        bool any = false;
        for (auto& a : arrays) any = any || a.section == sec;
        if (!any) return;
        std::string scalars, vectors; first_names(f, n, scalars, vectors);
        file << indent << "<" << tag;
        if (scalars != "") file << " Scalars=\"" << scalars << "\"";
        if (vectors != "") file << " Vectors=\"" << vectors << "\"";
        file << ">\n"; ++indent.level;
        for (auto& a : arrays) if (a.section == sec) main_array(a);
        --indent.level; file << indent << "</" << tag << ">\n";
      };
      file << indent << "<?xml version=\"1.0\"?>\n";
      file << indent << "<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n";
      ++indent.level;
      file << indent << "<UnstructuredGrid>\n"; ++indent.level;
      file << indent << "<Piece NumberOfCells=\"" << unsigned(ncells) << "\" NumberOfPoints=\"" << unsigned(nvertices) << "\">\n"; ++indent.level;
      section_of(0, "PointData", vertex_fields, nvertex_fields);
      section_of(1, "CellData", cell_fields, ncell_fields);
      file << indent << "<Points>\n"; ++indent.level;
      for (auto& a : arrays) if (a.section == 2) main_array(a);
      --indent.level; file << indent << "</Points>\n";
      file << indent << "<Cells>\n"; ++indent.level;
      for (auto& a : arrays) if (a.section == 3) main_array(a);
      --indent.level; file << indent << "</Cells>\n";
      --indent.level; file << indent << "</Piece>\n";
      --indent.level; file << indent << "</UnstructuredGrid>\n";
      if (outputtype == DGB_VTK_APPENDEDRAW || outputtype == DGB_VTK_APPENDEDBASE64) {          // vtuwriter.hh:345-365
        file << indent << "<AppendedData encoding=\"" << (outputtype == DGB_VTK_APPENDEDRAW ? "raw" : "base64") << "\">\n";
        ++indent.level;
        file << indent << "_";
        for (auto& a : arrays) {                                                               // Naked{Raw,Base64}DataArrayWriter
          const std::uint32_t size = std::uint32_t(std::size_t(a.ncomps)*std::size_t(a.nitems)*vtk_type_size(a.prec));
          if (outputtype == DGB_VTK_APPENDEDRAW) { RawSink r{file}; r.write(&size, 4); for (double x : a.v) vtk_put_value(r, a.prec, x); }
          else { B64Sink b(file); b.write(&size, 4); b.flush(); for (double x : a.v) vtk_put_value(b, a.prec, x); b.flush(); }
        }
        file << "\n";
        --indent.level;
        file << indent << "</AppendedData>\n";
      }
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
