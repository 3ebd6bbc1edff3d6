// dgb_capi.cu — grid life cycle, views, mapper set-up and host-side queries of the C ABI (include/dgb.h).
// Host arithmetic only (dgb_host.hh); the device is touched solely to upload tensor-product coordinates.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include "dgb_internal.hh"

namespace dgb {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
int fail(int code, const std::string& msg) { g_error = msg; return code; }

int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(DGB_ECUDA, std::string("no usable CUDA device (libdgb has no CPU fallback): ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  }
  return DGB_OK;
}

static bool have_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return n > 0;
}

static void free_dev_coords(dgb_grid* g) {
  for (auto& dc : g->devCoords)
    for (int i = 0; i < 3; ++i) if (dc.d[i]) { cudaFree(dc.d[i]); dc.d[i] = nullptr; }
  g->devCoords.clear();
}

// (re)upload tensor coordinates of all levels that do not have a device copy yet
static void sync_dev_coords(dgb_grid* g) {
  g->devCoords.resize(g->me.levels.size());
  if (g->me.spec.coord_mode != DGB_COORD_TENSORPRODUCT || !have_device()) return;
  for (std::size_t l = 0; l < g->me.levels.size(); ++l)
    for (int i = 0; i < g->me.spec.dim; ++i) {
      if (g->devCoords[l].d[i]) continue;
      const std::vector<double>& c = g->me.levels[l].coords.c[i];
      double* p = nullptr;
      if (cudaMalloc(&p, c.size()*sizeof(double)) != cudaSuccess) { cudaGetLastError(); continue; }
      cudaMemcpy(p, c.data(), c.size()*sizeof(double), cudaMemcpyHostToDevice);
      g->devCoords[l].d[i] = p;
    }
}

static void refresh_boundary_segments(dgb_grid* g) {
  // yaspgrid.hh:683-707 on the level-0 overlap box of codim 0
  const LevelBoxes& b = g->me.levels[0].b;
  const dgb_box& ov = b.comp[0].box[DGB_BOX_OVERLAP];
  int64_t n = 0;
  for (int k = 0; k < b.dim; ++k) {
    const int sides = (ov.origin[k] == 0) + (ov.origin[k]+ov.size[k] == g->me.N0[k]);
    int64_t off = 1;
    for (int l = 0; l < b.dim; ++l) if (l != k) off *= ov.size[l];
    n += sides*off;
  }
  g->nBoundarySegments = n;
}

int make_view(const dgb_grid* g, int level, dgb_view* v) {
  if (level < 0 || level >= int(g->me.levels.size())) return fail(DGB_ERANGE, "level out of range");
  const Level& L = g->me.levels[level];
  const LevelBoxes& b = L.b;
  std::memset(v, 0, sizeof(*v));
  v->dim = b.dim; v->level = level; v->rank = g->me.rank; v->nranks = g->me.nranks;
  v->periodic = g->me.spec.periodic; v->overlap = b.overlap; v->ncomp = b.ncomp;
  for (int i = 0; i < 3; ++i) { v->N[i] = b.N[i]; v->N0[i] = g->me.N0[i]; v->L[i] = g->me.L[i]; }
  for (int i = 0; i < 5; ++i) v->comp_begin[i] = b.comp_begin[i];
  for (int i = 0; i < 8; ++i) { v->shiftmap[i] = b.shiftmap[i]; v->comp[i] = b.comp[i]; }
  v->coord_mode = L.coords.mode;
  for (int i = 0; i < b.dim; ++i) {
    v->h[i] = L.coords.h[i]; v->origin[i] = L.coords.origin[i];
    if (L.coords.mode == DGB_COORD_TENSORPRODUCT) {
      v->tensor_host[i] = L.coords.c[i].data();
      v->tensor_offset[i] = L.coords.offset[i];
      v->tensor_size[i] = int(L.coords.c[i].size());
      v->tensor[i] = (level < int(g->devCoords.size())) ? g->devCoords[level].d[i] : nullptr;
    }
  }
  const dgb_box& ov0 = g->me.levels[0].b.comp[0].box[DGB_BOX_OVERLAP];
  for (int i = 0; i < b.dim; ++i) {
    v->bs_origin[i] = ov0.origin[i]; v->bs_size[i] = ov0.size[i];
    v->bs_sides[i] = (ov0.origin[i] == 0) + (ov0.origin[i]+ov0.size[i] == g->me.N0[i]);
  }
  for (int i = b.dim; i < 3; ++i) { v->bs_size[i] = 1; }
  return DGB_OK;
}

int64_t partition_size(const dgb_view& v, int codim, int pitype) {
  const int kind = box_kind_of(pitype);
  if (kind < 0) return 0;
  int64_t n = 0;
  for (int j = v.comp_begin[codim]; j < v.comp_begin[codim+1]; ++j) {
    int64_t t = 1;
    for (int i = 0; i < v.dim; ++i) t *= v.comp[j].box[kind].size[i];
    n += t;
  }
  return n;
}

} // namespace dgb

using namespace dgb;

extern "C" {

const char* dgb_last_error(void) { return g_error.c_str(); }
const char* dgb_version(void) { return "dune-grid-b200 0.1 (sm_100a)"; }

int dgb_device_count(int* n) {
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) { cudaGetLastError(); *n = 0; return fail(DGB_ECUDA, cudaGetErrorString(e)); }
  *n = c;
  return DGB_OK;
}

int dgb_partition(int dim, const int32_t* N, int P, int overlap, int partitioner, const int32_t* fixed, int32_t* dims_out) {
  return guarded([&] {
    if (dim < 1 || dim > 3) throw ArgError("dim must be 1..3");
    I3 n{{1,1,1}}, f{{1,1,1}};
    for (int i = 0; i < dim; ++i) { n[i] = N[i]; if (fixed) f[i] = fixed[i]; }
    I3 d = choose_dims(dim, n, P, overlap, partitioner, f);
    for (int i = 0; i < dim; ++i) dims_out[i] = d[i];
    return DGB_OK;
  });
}

int dgb_grid_create(const dgb_grid_spec* spec, int rank, int nranks, dgb_grid** out) {
  return guarded([&] {
    if (!spec || !out) throw ArgError("null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) throw ArgError("bad rank/nranks");
    std::unique_ptr<dgb_grid> g(new dgb_grid);
    g->me.construct(*spec, rank, nranks, nullptr, false);
    g->all.resize(nranks);
    for (int r = 0; r < nranks; ++r) g->all[r].construct(*spec, r, nranks, &g->me.torus, true);
    refresh_boundary_segments(g.get());
    sync_dev_coords(g.get());
    *out = g.release();
    return DGB_OK;
  });
}

// BackupRestoreFacility::backup (yaspgrid/backuprestore.hh:105-128, 222-246). std::ostream formatting = the reference's.
int dgb_grid_backup(const dgb_grid* g, char* text, int64_t cap, int64_t* needed) {
  return guarded([&] {
    if (!g || !needed) throw ArgError("null argument");
    const RankGrid& me = g->me;
    const int dim = me.spec.dim;
    std::ostringstream s;
    s << "YaspGrid BackupRestore Format Version: " << 2 << std::endl;
    s << "Torus structure: ";
    for (int i = 0; i < dim; ++i) s << me.torus.dims[i] << " ";
    s << std::endl << "Refinement level: " << int(me.levels.size())-1 << std::endl;
    s << "Periodicity: ";
    for (int i = 0; i < dim; ++i) s << ((me.spec.periodic >> i & 1u) ? "1 " : "0 ");
    s << std::endl << "Overlap: " << me.levels[0].b.overlap << std::endl;
    s << "KeepPhysicalOverlap: ";
    for (std::size_t l = 1; l < me.levels.size(); ++l) s << (me.levels[l].keep ? "1" : "0") << " ";
    s << std::endl;
    s << "Coarse Size: ";
    for (int i = 0; i < dim; ++i) s << me.N0[i] << " ";
    s << std::endl;
    const Coords& cc = me.levels[0].coords;
    if (cc.mode != DGB_COORD_TENSORPRODUCT) {
      s << "Meshsize: ";
      for (int i = 0; i < dim; ++i) s << cc.h[i] << " ";
      s << std::endl;
      if (cc.mode == DGB_COORD_EQUIDISTANT_OFFSET) {               // MaybeHaveOrigin::writeOrigin :51-57
        s << "Origin: ";
        for (int i = 0; i < dim; ++i) s << cc.origin[i] << " ";
        s << std::endl;
      }
    } else {                                                        // TensorProductCoordinates::print, coordinates.hh:347-356
      s << "Printing TensorProduct Coordinate information:" << std::endl;
      for (int i = 0; i < dim; ++i) {
        s << "Direction " << i << ": " << cc.c[i].size() << " coordinates" << std::endl;
        for (std::size_t j = 0; j < cc.c[i].size(); ++j) s << cc.c[i][j] << std::endl;
      }
    }
    const std::string out = s.str();
    *needed = int64_t(out.size()) + 1;
    if (text && cap >= *needed) std::memcpy(text, out.c_str(), out.size()+1);
    return DGB_OK;
  });
}

// BackupRestoreFacility::restore (yaspgrid/backuprestore.hh:146-218, 268-346): same token-skipping reads
int dgb_grid_restore(const char* text, int coord_mode, int rank, int nranks, dgb_grid** out) {
  return guarded([&] {
    if (!text || !out) throw ArgError("null argument");
    if (coord_mode < 0 || coord_mode > 2) throw ArgError("unknown coordinate mode");
    if (nranks < 1 || rank < 0 || rank >= nranks) throw ArgError("bad rank/nranks");
    std::istringstream stream(text);
    std::string input;
    int version = 0;
    stream >> input >> input >> input >> input;
    stream >> version;
    if (!stream || version != 2) throw ArgError("Your YaspGrid backup file is written in an outdated format!");
    // the reference is a template on dim; here the dimension is the number of integers on the torus line
    std::string line;
    std::getline(stream, line);                       // rest of line 1
    std::getline(stream, line);                       // "Torus structure: d0 d1 [d2] "
    std::vector<int> torus_dims;
    { std::istringstream ls(line); ls >> input >> input; int v; while (ls >> v) torus_dims.push_back(v); }
    const int dim = int(torus_dims.size());
    if (dim < 2 || dim > 3) throw ArgError("restore: dim must be 2 or 3");
    int refinement = 0;
    stream >> input >> input;
    stream >> refinement;
    unsigned periodic = 0;
    bool b = false;
    stream >> input;
    for (int i = 0; i < dim; ++i) { stream >> b; if (b) periodic |= 1u << i; }
    int overlap = 0;
    stream >> input;
    stream >> overlap;
    std::vector<bool> keep(refinement > 0 ? refinement : 0);
    stream >> input;
    for (int i = 0; i < refinement; ++i) { stream >> b; keep[i] = b; }
    dgb_grid_spec spec; std::memset(&spec, 0, sizeof(spec));
    spec.dim = dim; spec.coord_mode = coord_mode; spec.periodic = periodic; spec.overlap = overlap;
    spec.partitioner = DGB_PART_FIXED;
    for (int i = 0; i < 3; ++i) { spec.fixed_dims[i] = i < dim ? torus_dims[i] : 1; spec.N[i] = 1; }
    stream >> input >> input;
    for (int i = 0; i < dim; ++i) stream >> spec.N[i];
    std::vector<double> local[3];
    if (coord_mode != DGB_COORD_TENSORPRODUCT) {
      double h[3] = {0,0,0}, origin[3] = {0,0,0};
      stream >> input;
      for (int i = 0; i < dim; ++i) stream >> h[i];
      if (coord_mode == DGB_COORD_EQUIDISTANT_OFFSET) { stream >> input; for (int i = 0; i < dim; ++i) stream >> origin[i]; }
      if (!stream) throw ArgError("restore: truncated backup text");
      for (int i = 0; i < dim; ++i) {
        double length = h[i]; length *= spec.N[i];                   // :193-195
        if (coord_mode == DGB_COORD_EQUIDISTANT) spec.upper[i] = length;
        else { spec.lower[i] = origin[i]; double ur = origin[i]; ur += length; spec.upper[i] = ur; }     // :74-76
      }
    } else {
      stream >> input >> input >> input >> input;
      for (int d = 0; d < dim; ++d) {
        stream >> input >> input;
        int size = 0;
        stream >> size;
        stream >> input;
        if (!stream || size < 0) throw ArgError("restore: truncated backup text");
        local[d].resize(size);
        for (int i = 0; i < size; ++i) stream >> local[d][i];
      }
      if (!stream) throw ArgError("restore: truncated backup text");
    }
    std::unique_ptr<dgb_grid> g(new dgb_grid);
    g->me.construct(spec, rank, nranks, nullptr, false, coord_mode == DGB_COORD_TENSORPRODUCT ? local : nullptr);
    g->all.resize(nranks);
    for (int r = 0; r < nranks; ++r) g->all[r].construct(spec, r, nranks, &g->me.torus, true);
    refresh_boundary_segments(g.get());
    for (int i = 0; i < refinement; ++i) {                           // :206-210
      g->keepOverlap = keep[i];
      g->me.refine_once(g->keepOverlap, false);
      for (auto& o : g->all) o.refine_once(g->keepOverlap, true);
    }
    sync_dev_coords(g.get());
    *out = g.release();
    return DGB_OK;
  });
}

int dgb_grid_destroy(dgb_grid* g) {
  if (!g) return DGB_OK;
  free_dev_coords(g);
  delete g;
  return DGB_OK;
}

int dgb_grid_refine_options(dgb_grid* g, int keep) { g->keepOverlap = keep != 0; return DGB_OK; }

int dgb_grid_global_refine(dgb_grid* g, int refCount) {
  return guarded([&] {
    const int maxLevel = int(g->me.levels.size())-1;
    if (refCount < -maxLevel) throw GridError("Only " + std::to_string(maxLevel) + " levels left. Coarsening " + std::to_string(-refCount) + " levels requested!");
    std::lock_guard<std::mutex> lk(g->mtx);
    g->plans.clear();
    if (refCount < 0) {
      // only the device coordinate arrays of the popped levels are freed: views of the surviving levels (dgb_view::tensor)
      // held by callers keep pointing at live memory
      for (int k = refCount; k < 0; ++k) {
        g->me.levels.pop_back(); for (auto& o : g->all) o.levels.pop_back();
        if (g->devCoords.size() > g->me.levels.size()) {
          for (int i = 0; i < 3; ++i) if (g->devCoords.back().d[i]) cudaFree(g->devCoords.back().d[i]);
          g->devCoords.pop_back();
        }
      }
    }
    for (int k = 0; k < refCount; ++k) {
      if (g->me.levels.size() >= 32) throw GridError("at most 32 levels (ReservedVector<YGridLevel,32>)");
      g->me.refine_once(g->keepOverlap, false);
      for (auto& o : g->all) o.refine_once(g->keepOverlap, true);
    }
    sync_dev_coords(g);
    return DGB_OK;
  });
}

int dgb_grid_max_level(const dgb_grid* g, int* maxLevel) { *maxLevel = int(g->me.levels.size())-1; return DGB_OK; }
int dgb_grid_torus_dims(const dgb_grid* g, int32_t* dims) { for (int i = 0; i < g->me.spec.dim; ++i) dims[i] = g->me.torus.dims[i]; return DGB_OK; }
int dgb_grid_num_boundary_segments(const dgb_grid* g, int64_t* n) { *n = g->nBoundarySegments; return DGB_OK; }
int dgb_grid_domain_size(const dgb_grid* g, double* L) { for (int i = 0; i < g->me.spec.dim; ++i) L[i] = g->me.L[i]; return DGB_OK; }

int dgb_view_leaf(const dgb_grid* g, dgb_view* v) { return make_view(g, int(g->me.levels.size())-1, v); }
int dgb_view_level(const dgb_grid* g, int level, dgb_view* v) { return make_view(g, level, v); }

int dgb_view_size(const dgb_view* v, int codim, int64_t* n) {
  if (codim < 0 || codim > v->dim) return fail(DGB_EARG, "bad codim");
  *n = partition_size(*v, codim, DGB_All_Partition);
  return DGB_OK;
}
int dgb_view_partition_size(const dgb_view* v, int codim, int pitype, int64_t* n) {
  if (codim < 0 || codim > v->dim) return fail(DGB_EARG, "bad codim");
  if (box_kind_of(pitype) == -2) return fail(DGB_EGRID, "YaspLevelIterator with this codim or partition type not implemented");
  *n = partition_size(*v, codim, pitype);
  return DGB_OK;
}
int dgb_view_overlap_size(const dgb_view* v, int, int* n) { *n = v->overlap; return DGB_OK; }
int dgb_view_ghost_size(const dgb_view*, int, int* n) { *n = 0; return DGB_OK; }

int dgb_view_comm_list(const dgb_grid* g, int level, int codim, int which, int32_t* out, int cap, int* count) {
  return guarded([&] {
    if (level < 0 || level >= int(g->me.levels.size())) throw RangeError("level out of range");
    if (codim < 0 || codim > g->me.spec.dim || which < 0 || which > 7) throw ArgError("bad codim / list id");
    std::vector<ListEntry> l;
    comm_list(g->me, g->all, level, codim, which, l);
    *count = int(l.size());
    for (int k = 0; k < int(l.size()) && k < cap; ++k) {
      int32_t* o = out + 10*k;
      o[0] = l[k].rank; o[1] = l[k].distance; o[2] = int32_t(l[k].shift); o[3] = l[k].index_offset;
      for (int i = 0; i < 3; ++i) { o[4+i] = l[k].box.origin[i]; o[7+i] = l[k].box.size[i]; }
    }
    return DGB_OK;
  });
}

// MultipleCodimMultipleGeomTypeMapper::update_ (mcmgmapper.hh:373-402): offsets accumulate over codim 0..dim for
// codims with a non-zero block; n is a 32-bit unsigned int (wraps like the reference's `unsigned int n`).
int dgb_mapper_init(const dgb_view* v, const uint32_t* block, dgb_mapper* m) {
  std::memset(m, 0, sizeof(*m));
  m->dim = v->dim;
  uint32_t n = 0;
  for (int codim = 0; codim <= v->dim; ++codim) {
    m->block[codim] = block[codim];
    if (block[codim]) { m->offset[codim] = n; n += uint32_t(partition_size(*v, codim, DGB_All_Partition))*block[codim]; }
    else m->offset[codim] = DGB_INVALID_OFFSET;
  }
  for (int c = v->dim+1; c < 4; ++c) m->offset[c] = DGB_INVALID_OFFSET;
  m->size = n;
  return DGB_OK;
}
int dgb_mapper_size(const dgb_mapper* m, int64_t* n) { *n = m->size; return DGB_OK; }

// ---- field checkpoint ----------------------------------------------------------------------------------------------
static std::string field_header(const dgb_grid* g, int level, const dgb_mapper* m, int item_size, const std::string& gridText) {
  std::ostringstream h;
  h << "dune-grid-b200 field checkpoint v1 rank " << g->me.rank << " nranks " << g->me.nranks << " level " << level
    << " mapper " << m->block[0] << " " << m->block[1] << " " << m->block[2] << " " << m->block[3] << " size " << m->size
    << " item " << item_size << " gridtext " << gridText.size() << "\n";
  return h.str();
}
static int grid_text(const dgb_grid* g, std::string& t) {
  int64_t need = 0;
  if (int rc = dgb_grid_backup(g, nullptr, 0, &need)) return rc;
  t.assign(std::size_t(need), '\0');
  if (int rc = dgb_grid_backup(g, &t[0], need, &need)) return rc;
  t.resize(std::size_t(need)-1);
  return DGB_OK;
}

int dgb_field_backup(const dgb_grid* g, int level, const dgb_mapper* m, const double* d_field, int item_size, const char* filename, void* stream) {
  return guarded([&]() -> int {
    if (!g || !m || !d_field || !filename || item_size < 1) throw ArgError("bad argument");
    if (level < 0 || level >= int(g->me.levels.size())) throw RangeError("level out of range");
    if (int rc = require_device()) return rc;
    std::string gt; if (int rc = grid_text(g, gt)) return rc;
    const std::size_t n = std::size_t(m->size)*item_size;
    std::vector<double> host(n);
    DGB_CUDA(cudaMemcpyAsync(host.data(), d_field, n*sizeof(double), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    DGB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    const std::string path = std::string(filename) + "." + std::to_string(g->me.rank);
    std::ofstream f(path, std::ios::binary);
    if (!f) throw ArgError("field backup: cannot open " + path);
    const std::string h = field_header(g, level, m, item_size, gt);
    f.write(h.data(), std::streamsize(h.size()));
    f.write(gt.data(), std::streamsize(gt.size()));
    f.write(reinterpret_cast<const char*>(host.data()), std::streamsize(n*sizeof(double)));
    if (!f) throw ArgError("field backup: write failed on " + path);
    return DGB_OK;
  });
}

int dgb_field_restore(const dgb_grid* g, int level, const dgb_mapper* m, double* d_field, int item_size, const char* filename, void* stream) {
  return guarded([&]() -> int {
    if (!g || !m || !d_field || !filename || item_size < 1) throw ArgError("bad argument");
    if (level < 0 || level >= int(g->me.levels.size())) throw RangeError("level out of range");
    if (int rc = require_device()) return rc;
    std::string gt; if (int rc = grid_text(g, gt)) return rc;
    const std::string path = std::string(filename) + "." + std::to_string(g->me.rank);
    std::ifstream f(path, std::ios::binary);
    if (!f) throw ArgError("field restore: cannot open " + path);
    std::string h; std::getline(f, h); h += "\n";
    if (h != field_header(g, level, m, item_size, gt)) throw ArgError("field restore: " + path + " was written for another rank, level, mapper or item size");
    std::string stored(gt.size(), '\0');
    f.read(&stored[0], std::streamsize(stored.size()));
    if (!f || stored != gt) throw ArgError("field restore: " + path + " belongs to a different grid");
    const std::size_t n = std::size_t(m->size)*item_size;
    std::vector<double> host(n);
    f.read(reinterpret_cast<char*>(host.data()), std::streamsize(n*sizeof(double)));
    if (!f) throw ArgError("field restore: " + path + " is truncated");
    DGB_CUDA(cudaMemcpyAsync(d_field, host.data(), n*sizeof(double), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    DGB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return DGB_OK;
  });
}

int dgb_comm_last_bytes(const dgb_grid* g, int64_t* bytes) { *bytes = g->lastBytesSent; return DGB_OK; }

} // extern "C"
