// dgb_host.hh — host-side descriptor arithmetic of libdgb: processor grid, per-rank nested partition
// boxes, coordinate containers, refinement, and the send/recv box lists of communicate.
//
// Design (not a port): the reference builds these tables inside every rank by exchanging box
// descriptors with its 3^d-1 torus neighbours (two message rounds per component and interface,
// dune/grid/yaspgrid.hh:561-671). Here every rank evaluates the closed-form box rules for ANY rank,
// so a neighbour's boxes are computed locally and no message is ever needed to set up a grid.
// All results must equal the reference's bit for bit; each rule cites the lines it restates.
#ifndef DGB_HOST_HH
#define DGB_HOST_HH
#include <array>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/dgb.h"

namespace dgb {

struct GridError : std::runtime_error { using std::runtime_error::runtime_error; };
struct RangeError : std::runtime_error { using std::runtime_error::runtime_error; };
struct ArgError : std::runtime_error { using std::runtime_error::runtime_error; };

typedef std::array<int,3> I3;

inline int popcount3(unsigned x) { return int((x&1u)+((x>>1)&1u)+((x>>2)&1u)); }

// ---- a1: processor grid (partitioning.hh:56-117,122-139,145-165) ---------------------------------------
struct DimsSearch {
  int d; I3 N; int overlap; I3 trial{{-1,-1,-1}}, best{{-1,-1,-1}}; double bestCost = 1e100;
  void descend(int axis, int procs) {
    if (axis > 0) {
      for (int k = 1; k <= procs; ++k)
        if (procs % k == 0 && (k == 1 || N[axis]/k >= 2*overlap)) { trial[axis] = k; descend(axis-1, procs/k); }
      return;
    }
    if (!(procs == 1 || N[0]/procs >= 2*overlap)) return;
    trial[0] = procs;
    double cost = -1.0;
    for (int k = 0; k < d; ++k) {
      double per = double(N[k])/double(trial[k]);
      if (std::fmod(double(N[k]), double(trial[k])) > 0.0001) per *= 3;
      if (per > cost) cost = per;
    }
    if (cost < bestCost) { bestCost = cost; best = trial; }       // strict: first minimum wins
  }
};

inline I3 choose_dims(int dim, const I3& N, int P, int overlap, int partitioner, const I3& fixed) {
  I3 dims{{1,1,1}};
  if (partitioner == DGB_PART_DEFAULT) {
    DimsSearch s; s.d = dim; s.N = N; s.overlap = overlap;
    s.descend(dim-1, P);
    if (s.best[0] == -1) throw GridError("Failed to find a suitable partition");
    for (int i = 0; i < dim; ++i) dims[i] = s.best[i];
  } else if (partitioner == DGB_PART_POWER_D) {
    int root = -1;
    for (int i = 1; i <= P; ++i) { long p = 1; for (int k = 0; k < dim; ++k) p *= i; if (p == P) { root = i; break; } }
    if (root < 0) throw GridError("Power partitioning failed: your number of processes needs to be a d-th power.");
    for (int i = 0; i < dim; ++i) dims[i] = root;
  } else if (partitioner == DGB_PART_FIXED) {
    int prod = 1; for (int i = 0; i < dim; ++i) prod *= fixed[i];
    if (prod != P) throw ArgError("Your processor number doesn't match your partitioning information");
    for (int i = 0; i < dim; ++i) dims[i] = fixed[i];
  } else throw ArgError("unknown partitioner");
  return dims;
}

// ---- a5: sub-entity numbering of the reference cube (yaspgridentity.hh:85-89,154-233) -------------------
// number of codim-c sub-entities of a d-cube: C(d,c) * 2^c
inline int sub_entities(int d, int c) {
  if (d < c || c < 0) return 0;
  long b = 1;
  for (int k = 0; k < c; ++k) b = b*(d-k)/(k+1);
  return int(b << c);
}
// axes along which sub-entity `index` of codim cc extends (bit set = extends)
inline unsigned sub_shift(int dim, int index, int cc) {
  unsigned bits = 0;
  for (int d = dim; d > 0; --d) {
    if (cc == d) return bits;
    const int below = sub_entities(d-1, cc);
    if (index < below) bits |= 1u << (d-1);
    else { index = (index-below) % sub_entities(d-1, cc-1); --cc; }
  }
  return bits;
}
// axes along which the cell carrying sub-entity `index` is the +1 neighbour (bit set = move by one)
inline unsigned sub_move(int dim, int index, int cc) {
  unsigned bits = 0;
  for (int d = dim; d > 0; --d) {
    const unsigned top = 1u << (d-1);
    if (d == cc) { bits = (bits & ~top) | (unsigned(index) & top); index &= ~int(top); }
    const int below = sub_entities(d-1, cc);
    if (index >= below) {
      const int side = sub_entities(d-1, cc-1);
      if ((index-below)/side == 1) bits |= top;
      index = (index-below) % side;
      --cc;
    }
  }
  return bits;
}

// ---- torus (torus.hh:144-163,239-268,472-522) ----------------------------------------------------------
struct Torus {
  int dim = 0, P = 1; I3 dims{{1,1,1}}, inc{{1,1,1}};
  void init(int d, const I3& dm) {
    dim = d; dims = dm; int run = 1;
    for (int i = 0; i < d; ++i) { inc[i] = run; run *= dims[i]; }
    P = run;
  }
  I3 coord_of(int rank) const {
    I3 c{{0,0,0}}; rank %= P;
    for (int i = dim-1; i >= 0; --i) { c[i] = rank/inc[i]; rank %= inc[i]; }
    return c;
  }
  int rank_of(I3 c) const {
    int r = 0;
    for (int i = 0; i < dim; ++i) r += (c[i] % dims[i])*inc[i];
    return r;
  }
  // block split of N cells along each axis: the last N%dims blocks are one cell larger
  void block(const I3& c, const I3& N, I3& o, I3& s) const {
    for (int i = 0; i < dim; ++i) {
      const int m = N[i]/dims[i], r = N[i]%dims[i], small = dims[i]-r;
      if (c[i] < small) { o[i] = c[i]*m; s[i] = m; }
      else { o[i] = small*m + (c[i]-small)*(m+1); s[i] = m+1; }
    }
  }
  int neighbors() const { int n = 1; for (int i = 0; i < dim; ++i) n *= 3; return n-1; }
  // k-th non-zero delta in ascending order, delta[0] fastest, starting at (-1,..,-1)
  I3 delta(int k) const {
    int total = neighbors()+1, centre = total/2;
    int lin = k < centre ? k : k+1;
    I3 d{{0,0,0}};
    for (int i = 0; i < dim; ++i) { d[i] = lin%3 - 1; lin /= 3; }
    return d;
  }
};

// ---- a3: coordinate containers (coordinates.hh:27-116,129-230,243-361) ---------------------------------
struct Coords {
  int mode = 0, dim = 0;
  double h[3] = {0,0,0}, origin[3] = {0,0,0};
  int s[3] = {0,0,0};                 // cells held
  std::vector<double> c[3];           // tensor
  int offset[3] = {0,0,0};
  int size(int d) const { return mode == DGB_COORD_TENSORPRODUCT ? int(c[d].size())-1 : s[d]; }
  double coordinate(int d, int i) const {
    if (mode == DGB_COORD_EQUIDISTANT) return i*h[d];
    if (mode == DGB_COORD_EQUIDISTANT_OFFSET) return origin[d] + i*h[d];
    return c[d][i-offset[d]];
  }
  double meshsize(int d, int i) const {
    if (mode == DGB_COORD_TENSORPRODUCT) return c[d][i+1-offset[d]] - c[d][i-offset[d]];
    return h[d];
  }
  // one uniform refinement (coordinates.hh:84-104,196-216,297-344)
  Coords refine(const bool* low, const bool* up, int overlap, bool keep) const {
    Coords r; r.mode = mode; r.dim = dim;
    for (int i = 0; i < dim; ++i) {
      if (mode != DGB_COORD_TENSORPRODUCT) {
        int news = 2*s[i];
        if (!keep) { if (low[i]) news -= overlap; if (up[i]) news -= overlap; }
        r.s[i] = news;
        r.origin[i] = origin[i];
        if (mode == DGB_COORD_EQUIDISTANT) {
          const double ur = (h[i]/2.0)*news;
          r.h[i] = ur/news;
        } else {
          const double ur = origin[i] + (h[i]/2.0)*news;
          r.h[i] = (ur - origin[i])/news;
        }
      } else {
        const std::vector<double>& old = c[i];
        int newoffset = 2*offset[i];
        std::size_t first = 0, last = old.size()-1;          // index range of old coordinates that survive
        std::vector<double>& nc = r.c[i];
        if (!keep) {
          if (low[i]) {
            newoffset += overlap;
            first += overlap/2;
            if (overlap%2) { nc.push_back((old[first] + old[first+1])/2.0); ++first; }
          }
          if (up[i]) last -= overlap/2;
        }
        for (std::size_t j = first; j < last; ++j) { nc.push_back(old[j]); nc.push_back((old[j] + old[j+1])/2.0); }
        int newsize = 2*int(old.size()) - 1;
        if (!keep) { if (low[i]) newsize -= overlap; if (up[i]) newsize -= overlap; }
        if (int(nc.size()) < newsize) nc.push_back(old[last]);
        r.offset[i] = newoffset;
      }
    }
    return r;
  }
};

// ---- a2: integer structure of one level on one rank (yaspgrid.hh:327-529) -------------------------------
struct LevelBoxes {
  int dim = 0, level = 0, overlap = 0;
  I3 N{{1,1,1}};                         // cells of the whole level
  I3 oI{{0,0,0}}, sStored{{1,1,1}};      // interior origin, stored (coordinate container) size
  int ncomp = 0;
  int comp_begin[5] = {0,0,0,0,0};
  int shiftmap[8] = {0,0,0,0,0,0,0,0};
  dgb_component comp[8];
  bool low[3] = {false,false,false}, up[3] = {false,false,false};
};

inline int box_total(const dgb_box& b, int dim) { int t = 1; for (int i = 0; i < dim; ++i) t *= b.size[i]; return t; }
inline bool box_empty(const dgb_box& b, int dim) { for (int i = 0; i < dim; ++i) if (b.size[i] == 0) return true; return false; }

inline void build_boxes(LevelBoxes& g, unsigned periodic) {
  const int dim = g.dim, ov = g.overlap;
  I3 oO{{0,0,0}};
  for (int i = 0; i < dim; ++i) {
    g.low[i] = g.up[i] = false;
    if (periodic & (1u<<i)) { oO[i] = g.oI[i]-ov; g.low[i] = g.up[i] = true; }
    else {
      if (g.oI[i]-ov < 0) oO[i] = 0; else { oO[i] = g.oI[i]-ov; g.low[i] = true; }
      if (oO[i] + g.sStored[i] < g.N[i]) g.up[i] = true;
    }
  }
  int n = 0;
  for (int codim = 0; codim <= dim; ++codim) {
    g.comp_begin[codim] = n;
    int k = 0;
    for (unsigned shift = 0; shift < (1u<<dim); ++shift) {
      if (popcount3(shift) != dim-codim) continue;
      dgb_component& c = g.comp[n];
      c.shift = shift; c.codim = codim;
      dgb_box b; for (int i = 0; i < 3; ++i) { b.origin[i] = i < dim ? oO[i] : 0; b.size[i] = i < dim ? g.sStored[i] : 1; }
      for (int i = 0; i < dim; ++i) if (!(shift>>i & 1u)) b.size[i]++;          // entities without extent along i: one more layer
      c.box[DGB_BOX_OVERLAPFRONT] = b;
      for (int i = 0; i < dim; ++i) if (!(shift>>i & 1u)) {                      // overlap: drop the front layers
        if (g.low[i]) { b.origin[i]++; b.size[i]--; }
        if (g.up[i]) b.size[i]--;
      }
      c.box[DGB_BOX_OVERLAP] = b;
      for (int i = 0; i < dim; ++i) {                                           // interiorborder
        const bool flat = !(shift>>i & 1u);
        if (g.low[i]) { b.origin[i] += ov; b.size[i] -= ov; if (flat) { b.origin[i]--; b.size[i]++; } }
        if (g.up[i])  { b.size[i] -= ov; if (flat) b.size[i]++; }
      }
      c.box[DGB_BOX_INTERIORBORDER] = b;
      for (int i = 0; i < dim; ++i) if (!(shift>>i & 1u)) {                      // interior: drop the border layers
        if (g.low[i]) { b.origin[i]++; b.size[i]--; }
        if (g.up[i]) b.size[i]--;
      }
      c.box[DGB_BOX_INTERIOR] = b;
      g.shiftmap[shift] = k++;
      ++n;
    }
    // YGrid::finalize (ygrid.hh:771-791): the running offset uses each YGrid's OWN component sizes, so for
    // the sub-partitions of codim>0 it differs from the overlapfront offset (reference behaviour, reproduced).
    for (int q = 0; q < 4; ++q) {
      int run = 0;
      for (int j = g.comp_begin[codim]; j < n; ++j) { g.comp[j].index_offset[q] = run; run += box_total(g.comp[j].box[q], dim); }
    }
  }
  g.comp_begin[dim+1] = n;
  for (int c = dim+2; c < 5; ++c) g.comp_begin[c] = n;
  g.ncomp = n;
}

// one entry of a send/recv list (YGridList::Intersection, ygrid.hh:828-838 + finalize :962-1012)
// `move`: the partner's box was shifted by this vector into this rank's frame (periodic wrap: +-N, else 0), i.e. a cell at
// coordinate x here is the cell x - move in the partner's own frame
struct ListEntry { int rank, distance; unsigned shift; int index_offset; dgb_box box; int move[3]; int delta[3]; };   // delta: torus offset of the partner

struct Level {
  LevelBoxes b;
  Coords coords;
  bool keep = true;
};

// Everything a rank knows about the grid. `ints_only` skips coordinates (used for neighbour ranks).
struct RankGrid {
  dgb_grid_spec spec;
  int rank = 0, nranks = 1;
  Torus torus;
  I3 N0{{1,1,1}};
  double L[3] = {0,0,0};
  std::vector<Level> levels;
  std::vector<std::vector<double>> tensor_in;     // host copy of the user's tensor coordinates

  // `local` (tensor mode, BackupRestoreFacility::restore only): the rank's own level-0 coordinates incl. overlap, as the
  // private constructor yaspgrid.hh:1156-1197 takes them; the global coordinate arrays are then not known.
  void construct(const dgb_grid_spec& s, int rk, int np, const Torus* shared, bool ints_only,
                 const std::vector<double>* local = nullptr) {
    spec = s; rank = rk; nranks = np;
    const int dim = s.dim;
    if (dim < 2 || dim > 3) throw ArgError("dim must be 2 or 3");
    if (s.overlap < 0) throw ArgError("negative overlap");
    for (int i = 0; i < 3; ++i) N0[i] = i < dim ? s.N[i] : 1;
    for (int i = 0; i < dim; ++i) if (N0[i] < 1) throw ArgError("grid size must be positive");
    if (s.coord_mode == DGB_COORD_TENSORPRODUCT && !ints_only && local) {
      for (int i = 0; i < dim; ++i) {
        if (local[i].size() <= 1) throw GridError("Setup of a tensorproduct grid requires monotonous sequences of coordinates.");
        for (std::size_t j = 1; j < local[i].size(); ++j)
          if (local[i][j] < local[i][j-1]) throw GridError("Setup of a tensorproduct grid requires monotonous sequences of coordinates.");
      }
    } else if (s.coord_mode == DGB_COORD_TENSORPRODUCT && !ints_only) {
      tensor_in.resize(dim);
      for (int i = 0; i < dim; ++i) {
        if (!s.tensor[i]) throw ArgError("tensor coordinates missing");
        tensor_in[i].assign(s.tensor[i], s.tensor[i]+N0[i]+1);
        // Yasp::checkIfMonotonous (coordinates.hh:371-383)
        if (tensor_in[i].size() <= 1) throw GridError("Setup of a tensorproduct grid requires monotonous sequences of coordinates.");
        for (std::size_t j = 1; j < tensor_in[i].size(); ++j)
          if (tensor_in[i][j] < tensor_in[i][j-1]) throw GridError("Setup of a tensorproduct grid requires monotonous sequences of coordinates.");
      }
    }
    if (shared) torus = *shared;
    else {
      I3 fixed{{s.fixed_dims[0], s.fixed_dims[1], s.fixed_dims[2]}};
      I3 dims = choose_dims(dim, N0, np, s.overlap, s.partitioner, fixed);
      torus.init(dim, dims);
      if (torus.P != np) throw ArgError("Communicator size and result of the given load balancer do not match!");
    }
    const int ov = s.overlap;
    I3 me = torus.coord_of(rank), oI{{0,0,0}}, sI{{1,1,1}};
    torus.block(me, N0, oI, sI);
    if (!shared) {
      // "grid too small" test of the constructors (yaspgrid.hh:925-937, 997-1009, 1072-1084): a collective
      // logical-or over ranks in the reference; evaluated here for every rank arithmetically.
      for (int r = 0; r < np; ++r) {
        I3 c = torus.coord_of(r), o{{0,0,0}}, sz{{1,1,1}};
        torus.block(c, N0, o, sz);
        for (int i = 0; i < dim; ++i) {
          const bool per = s.periodic>>i & 1u;
          if (sz[i] < 2*ov && (per || sz[i] != N0[i]))
            throw GridError("YaspGrid does not support degrees of freedom shared by more than immediately neighboring subdomains."
                            " Note that this also holds for DOFs on subdomain boundaries."
                            " Increase grid elements or decrease overlap accordingly.");
        }
      }
    }
    Level lv; lv.keep = true;
    LevelBoxes& b = lv.b;
    b.dim = dim; b.level = 0; b.overlap = ov; b.N = N0; b.oI = oI;
    Coords& cc = lv.coords; cc.mode = s.coord_mode; cc.dim = dim;
    for (int i = 0; i < dim; ++i) {
      const bool per = s.periodic>>i & 1u;
      if (s.coord_mode != DGB_COORD_TENSORPRODUCT) {
        int so = sI[i];                                          // yaspgrid.hh:940-947 / 1012-1019
        if (oI[i]-ov > 0 || per) so += ov;
        if (oI[i]+sI[i]+ov <= N0[i] || per) so += ov;
        b.sStored[i] = so; cc.s[i] = so;
        if (s.coord_mode == DGB_COORD_EQUIDISTANT) {
          L[i] = s.upper[i];
          const double ur = (s.upper[i]/N0[i])*so;               // :952
          cc.h[i] = ur/so;                                       // coordinates.hh:49
        } else {
          L[i] = s.upper[i]-s.lower[i];
          cc.origin[i] = s.lower[i];
          const double ur = s.lower[i] + so*(s.upper[i]-s.lower[i])/N0[i];    // :1023-1024
          cc.h[i] = (ur - s.lower[i])/so;                        // coordinates.hh:153
        }
      } else {
        // yaspgrid.hh:1088-1131: copy the local coordinate range, extend periodically at physical boundaries
        int begin = oI[i], end = oI[i]+sI[i]+1, off = oI[i];
        if (oI[i]-ov > 0) { begin -= ov; off -= ov; }
        if (oI[i]+sI[i]+ov < N0[i]) end += ov;
        int cnt = end-begin;
        const bool addHigh = per && (oI[i]+sI[i]+ov >= N0[i]);
        const bool addLow = per && (oI[i]-ov <= 0);
        if (addHigh) cnt += ov;
        if (addLow) { cnt += ov; off -= ov; }
        b.sStored[i] = cnt-1;
        cc.offset[i] = off;
        if (!ints_only && local) {
          // yaspgrid.hh:1172-1190: _L from the LOCAL coordinates, offset = o_interior (- overlap if periodic or o_interior > 0)
          if (int(local[i].size()) != cnt) throw GridError("restore: the number of coordinates does not match the partition of this rank");
          L[i] = local[i].back() - local[i].front();
          cc.offset[i] = oI[i] - ((per || oI[i] > 0) ? ov : 0);
          cc.c[i] = local[i];
        } else if (!ints_only) {
          const std::vector<double>& in = tensor_in[i];
          L[i] = in[N0[i]] - in[0];
          std::vector<double> nc(in.begin()+begin, in.begin()+end);
          if (addHigh) for (int j = 0; j < ov; ++j) nc.push_back(nc.back() - in[j] + in[j+1]);
          if (addLow) { int rc = N0[i]; for (int j = 0; j < ov; ++j, --rc) nc.insert(nc.begin(), nc.front() - in[rc] + in[rc-1]); }
          cc.c[i] = nc;
        }
      }
    }
    build_boxes(b, s.periodic);
    levels.clear();
    levels.push_back(lv);
  }

  // yaspgrid.hh:1234-1263
  void refine_once(bool keep, bool ints_only) {
    const Level& cg = levels.back();
    const int dim = spec.dim;
    bool low[3], up[3];
    const dgb_box& ovb = cg.b.comp[0].box[DGB_BOX_OVERLAP];
    for (int i = 0; i < dim; ++i) {
      const bool per = spec.periodic>>i & 1u;
      low[i] = ovb.origin[i] > 0 || per;
      up[i] = ovb.origin[i]+ovb.size[i]-1 + 1 < cg.b.N[i] || per;
    }
    Level lv; lv.keep = keep;
    LevelBoxes& b = lv.b;
    b.dim = dim; b.level = cg.b.level+1;
    b.overlap = keep ? 2*cg.b.overlap : cg.b.overlap;
    for (int i = 0; i < 3; ++i) b.N[i] = i < dim ? N0[i]*(1<<b.level) : 1;
    for (int i = 0; i < dim; ++i) b.oI[i] = 2*cg.b.comp[0].box[DGB_BOX_INTERIOR].origin[i];
    if (ints_only) {
      lv.coords.mode = spec.coord_mode; lv.coords.dim = dim;
      for (int i = 0; i < dim; ++i) {
        int news = 2*cg.b.sStored[i];
        if (!keep) { if (low[i]) news -= cg.b.overlap; if (up[i]) news -= cg.b.overlap; }
        b.sStored[i] = news;
      }
    } else {
      lv.coords = cg.coords.refine(low, up, cg.b.overlap, keep);
      for (int i = 0; i < dim; ++i) b.sStored[i] = lv.coords.size(i);
    }
    build_boxes(b, spec.periodic);
    levels.push_back(lv);
  }
};

// ---- send/recv lists of one component pair (yaspgrid.hh:561-671) ----------------------------------------
inline dgb_box box_intersect(const dgb_box& a, const dgb_box& b, int dim, bool& empty) {
  dgb_box r; for (int i = 0; i < 3; ++i) { r.origin[i] = 0; r.size[i] = i < dim ? 0 : 1; }
  empty = true;
  if (box_empty(a, dim) || box_empty(b, dim)) return r;
  for (int i = 0; i < dim; ++i) {
    const int lo = std::max(a.origin[i], b.origin[i]);
    const int hi = std::min(a.origin[i]+a.size[i]-1, b.origin[i]+b.size[i]-1);
    r.origin[i] = lo; r.size[i] = hi-lo+1;
  }
  empty = box_empty(r, dim);
  return r;
}

// which: see dgb_view_comm_list. `others[r]` must hold RankGrid descriptors (ints only) of all ranks.
inline void comm_list(const RankGrid& me, const std::vector<RankGrid>& others, int level, int codim, int which,
                      std::vector<ListEntry>& out) {
  static const int sendKind[4] = { DGB_BOX_OVERLAPFRONT, DGB_BOX_OVERLAP, DGB_BOX_INTERIORBORDER, DGB_BOX_INTERIORBORDER };
  static const int recvKind[4] = { DGB_BOX_OVERLAPFRONT, DGB_BOX_OVERLAPFRONT, DGB_BOX_INTERIORBORDER, DGB_BOX_OVERLAPFRONT };
  const int pair = which/2; const bool wantSend = (which%2 == 0);
  const int dim = me.spec.dim;
  const LevelBoxes& mb = me.levels[level].b;
  const Torus& T = me.torus;
  const I3 mc = T.coord_of(me.rank);
  out.clear();
  int running = 0;                                                     // YGridList::finalize offset (supersize products)
  for (int j = mb.comp_begin[codim]; j < mb.comp_begin[codim+1]; ++j) {
    const dgb_component& mine = mb.comp[j];
    std::vector<ListEntry> tmp;
    const int nn = T.neighbors();
    for (int k = 0; k < nn; ++k) {                                      // recv-list order: ascending delta
      const I3 dl = T.delta(k);
      I3 nbc{{0,0,0}};
      for (int i = 0; i < dim; ++i) nbc[i] = (mc[i] + T.dims[i] + dl[i]) % T.dims[i];
      const int nbrank = T.rank_of(nbc);
      // the neighbour prepared its grids for delta' = -delta: moved by v unless it falls off a non-periodic end
      bool skip = false; I3 v{{0,0,0}};
      for (int i = 0; i < dim; ++i) {
        const int tgt = nbc[i] - dl[i];
        const bool per = me.spec.periodic>>i & 1u;
        if (tgt < 0) { if (per) v[i] += mb.N[i]; else skip = true; }
        if (tgt >= T.dims[i]) { if (per) v[i] -= mb.N[i]; else skip = true; }
      }
      if (skip) continue;
      const dgb_component& theirs = others[nbrank].levels[level].b.comp[j];
      dgb_box other = wantSend ? theirs.box[recvKind[pair]] : theirs.box[sendKind[pair]];
      for (int i = 0; i < dim; ++i) other.origin[i] += v[i];
      const dgb_box& local = wantSend ? mine.box[sendKind[pair]] : mine.box[recvKind[pair]];
      bool empty;
      dgb_box is = box_intersect(local, other, dim, empty);
      if (empty) continue;
      int dist = 0; for (int i = 0; i < dim; ++i) dist += std::abs(dl[i]);
      ListEntry e{nbrank, dist, mine.shift, running, is, {v[0], v[1], v[2]}, {dl[0], dl[1], dl[2]}};
      tmp.push_back(e);
    }
    if (wantSend) out.insert(out.end(), tmp.rbegin(), tmp.rend());      // push_front: descending delta
    else out.insert(out.end(), tmp.begin(), tmp.end());
    running += box_total(mine.box[DGB_BOX_OVERLAPFRONT], dim);
  }
}

} // namespace dgb
#endif
