// dgb_capi.cu — grid life cycle, views, mapper set-up and host-side queries of the C ABI (include/dgb.h).
// Host arithmetic only (dgb_host.hh); the device is touched solely to upload tensor-product coordinates.
#include <cstring>
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
      free_dev_coords(g);
      for (int k = refCount; k < 0; ++k) { g->me.levels.pop_back(); for (auto& o : g->all) o.levels.pop_back(); }
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

int dgb_comm_last_bytes(const dgb_grid* g, int64_t* bytes) { *bytes = g->lastBytesSent; return DGB_OK; }

} // extern "C"
