// port_oracle.cc — oracle/liboracle_port.so: a self-contained CPU restatement of the reference's YaspGrid leaf-sweep
// path, exporting oracle/oracle_api.h. TEST INFRASTRUCTURE ONLY (never linked into the product, never on a timed
// product path). It needs nothing from /root/reference at build or run time, so it travels in git; it is pinned to the
// reference by the golden vectors under tests/golden/ (generated from oracle/_ref = the reference's own headers) and by
// the reference's known-answer tests (tests/test_reference_kats.py).
//
// Structure follows the reference, not the product: every rank of a "world" builds its level the way
// YaspGrid::makelevel does (dune/grid/yaspgrid.hh:327-529), and the send/recv lists are obtained like
// YaspGrid::intersections (:561-671) by letting every rank post its moved boxes for each torus neighbour into an
// in-process mailbox and intersecting what it "receives" — no closed-form neighbour arithmetic. Sweeps walk boxes with
// a lexicographic iterator like YGridComponent::Iterator (yaspgrid/ygrid.hh:317-400). Paths below are relative to the
// dune-grid source tree. Geometry arithmetic (AxisAlignedCubeGeometry, MultiLinearGeometry) lives in dune-geometry,
// which the reference does not vendor: restated from its published semantics — unpinned.
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <limits>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../oracle_api.h"
#include "../common/fv_problem.hh"
#include "../common/multilinear.hh"

namespace {

thread_local std::string g_err;
struct GridError : std::runtime_error { using std::runtime_error::runtime_error; };

typedef std::array<int,3> I3;

// ---- a5: sub-entity numbering of the reference cube -----------------------------------------------------------
// Built bottom-up from the prism construction of the reference element instead of the reference's top-down bit
// recursion (yaspgridentity.hh:154-233): a d-cube is the prism over a (d-1)-cube; its codim-cc sub-entities are
// first the base's codim-cc ones extruded along axis d-1, then the base's codim-(cc-1) ones on the bottom, then on the
// top. shift bit j = the sub-entity extends along axis j; move bit j = it lies on the upper side of axis j.
struct SubEntity { unsigned shift, move; };
std::vector<SubEntity> subEntityTable(int d, int cc) {
  if (cc < 0 || cc > d) return {};
  if (d == 0) return { SubEntity{0u, 0u} };
  std::vector<SubEntity> t;
  for (SubEntity e : subEntityTable(d-1, cc)) t.push_back({ e.shift | (1u << (d-1)), e.move });
  const std::vector<SubEntity> faces = subEntityTable(d-1, cc-1);
  for (SubEntity e : faces) t.push_back({ e.shift, e.move });
  for (SubEntity e : faces) t.push_back({ e.shift, e.move | (1u << (d-1)) });
  return t;
}
int binomial(int n, int k) { if (k < 0 || k > n) return 0; long b = 1; for (int i = 0; i < k; ++i) b = b*(n-i)/(i+1); return int(b); }
int subEnt(int d, int c) { return d < c ? 0 : binomial(d, c) << c; }            // yaspgridentity.hh:85-89

// ---- a1: processor grid (yaspgrid/partitioning.hh:45-165) -------------------------------------------------------
void optimizeDims(int dim, int i, const I3& size, int P, I3& dims, I3& trydims, double& opt, int overlap) {
  if (i > 0) {
    for (int k = 1; k <= P; ++k)
      if (P % k == 0 && (k == 1 || size[i]/k >= 2*overlap)) { trydims[i] = k; optimizeDims(dim, i-1, size, P/k, dims, trydims, opt, overlap); }
  } else {
    if (P == 1 || size[0]/P >= 2*overlap) trydims[0] = P; else return;
    double m = -1.0;
    for (int k = 0; k < dim; ++k) {
      double mm = double(size[k])/double(trydims[k]);
      if (std::fmod(double(size[k]), double(trydims[k])) > 0.0001) mm *= 3;
      if (mm > m) m = mm;
    }
    if (m < opt) { opt = m; dims = trydims; }
  }
}
I3 partitionDims(int dim, const I3& size, int P, int overlap, int partitioner, const I3& fixed) {
  I3 dims{{-1,-1,-1}};
  if (partitioner == 0) {
    I3 trydims{{-1,-1,-1}}; double opt = 1e100;
    optimizeDims(dim, dim-1, size, P, dims, trydims, opt, overlap);
    if (dims[0] == -1) throw GridError("Failed to find a suitable partition");
  } else if (partitioner == 1) {
    bool ok = false;
    for (int i = 1; i <= P && !ok; ++i) { long p = 1; for (int k = 0; k < dim; ++k) p *= i; if (p == P) { for (int k = 0; k < dim; ++k) dims[k] = i; ok = true; } }
    if (!ok) throw GridError("Power partitioning failed: your number of processes needs to be a d-th power.");
  } else {
    int prod = 1; for (int i = 0; i < dim; ++i) prod *= fixed[i];
    if (prod != P) throw std::runtime_error("Your processor number doesn't match your partitioning information");
    for (int i = 0; i < dim; ++i) dims[i] = fixed[i];
  }
  return dims;
}

// ---- torus (yaspgrid/torus.hh) --------------------------------------------------------------------------------
struct CommPartner { int rank; I3 delta; int index; int distance() const { return std::abs(delta[0])+std::abs(delta[1])+std::abs(delta[2]); } };
struct Torus {
  int d = 0, P = 1, me = 0; I3 dims{{1,1,1}}, inc{{1,1,1}};
  std::deque<CommPartner> sendlist, recvlist;
  I3 rankToCoord(int rank) const { I3 c{{0,0,0}}; rank %= P; for (int i = d-1; i >= 0; --i) { c[i] = rank/inc[i]; rank %= inc[i]; } return c; }    // :144-154
  int coordToRank(I3 c) const { int r = 0; for (int i = 0; i < d; ++i) r += (c[i] % dims[i])*inc[i]; return r; }                             // :157-163
  int neighbors() const { int n = 1; for (int i = 0; i < d; ++i) n *= 3; return n-1; }
  void init(int dim, const I3& dm, int rank) {
    d = dim; dims = dm; me = rank; int run = 1;
    for (int i = 0; i < d; ++i) { inc[i] = run; run *= dims[i]; }
    P = run;
    // proclists (:472-522): recv list ascending in delta (delta[0] fastest), send list its reverse
    I3 delta{{0,0,0}}; for (int i = 0; i < d; ++i) delta[i] = -1;
    const I3 mec = rankToCoord(me);
    int index = 0; const int last = neighbors()-1;
    bool ready = false;
    while (!ready) {
      I3 nb{{0,0,0}};
      for (int i = 0; i < d; ++i) nb[i] = (mec[i]+dims[i]+delta[i]) % dims[i];
      const int nbrank = coordToRank(nb);
      bool nonzero = false; for (int i = 0; i < d; ++i) if (delta[i] != 0) nonzero = true;
      if (nonzero) { recvlist.push_back({nbrank, delta, index}); sendlist.push_front({nbrank, delta, last-index}); ++index; }
      ready = true;
      for (int i = 0; i < d; ++i) { if (delta[i] < 1) { ++delta[i]; ready = false; break; } else delta[i] = -1; }
    }
  }
  void partition(const I3& sizeIn, I3& o, I3& s) const {                                                                                    // :239-268
    const I3 c = rankToCoord(me);
    for (int i = 0; i < d; ++i) {
      const int m = sizeIn[i]/dims[i], r = sizeIn[i] % dims[i];
      if (c[i] < dims[i]-r) { o[i] = c[i]*m; s[i] = m; }
      else { o[i] = (dims[i]-r)*m + (c[i]-(dims[i]-r))*(m+1); s[i] = m+1; }
    }
  }
};

// ---- a3: coordinate containers (yaspgrid/coordinates.hh) ----------------------------------------------------------
struct Coordinates {
  int mode = 0, dim = 0;
  double h[3] = {0,0,0}, origin[3] = {0,0,0}; int s[3] = {0,0,0};
  std::vector<double> c[3]; int offset[3] = {0,0,0};
  double coordinate(int d, int i) const {
    if (mode == 0) return i*h[d];                                   // :65-68
    if (mode == 1) return origin[d] + i*h[d];                       // :169-172
    return c[d][i-offset[d]];                                       // :278-281
  }
  int size(int d) const { return mode == 2 ? int(c[d].size())-1 : s[d]; }
  void setEquidistant(int d, double upperRight, int size) { s[d] = size; h[d] = upperRight/size; }                        // :45-50
  void setOffset(int d, double lowerLeft, double upperRight, int size) { origin[d] = lowerLeft; s[d] = size; h[d] = (upperRight-lowerLeft)/size; }   // :148-154
  Coordinates refine(const bool* low, const bool* up, int overlap, bool keep) const {                                      // :84-104,196-216,297-344
    Coordinates r; r.mode = mode; r.dim = dim;
    for (int i = 0; i < dim; ++i) {
      if (mode != 2) {
        int news = 2*s[i];
        if (!keep) { if (low[i]) news -= overlap; if (up[i]) news -= overlap; }
        if (mode == 0) r.setEquidistant(i, (h[i]/2.0)*news, news);
        else r.setOffset(i, origin[i], origin[i] + (h[i]/2.0)*news, news);
      } else {
        int newoffset = 2*offset[i];
        int newsize = 2*int(c[i].size()) - 1;
        std::size_t it = 0, end = c[i].size()-1;
        std::vector<double> nc;
        if (!keep) {
          if (low[i]) { newoffset += overlap; newsize -= overlap; }
          if (up[i]) newsize -= overlap;
        }
        nc.reserve(newsize);
        if (!keep) {
          if (low[i]) { it += overlap/2; if (overlap % 2) { nc.push_back((c[i][it] + c[i][it+1])/2.0); ++it; } }
          if (up[i]) end -= overlap/2;
        }
        while (it != end) { nc.push_back(c[i][it]); nc.push_back((c[i][it] + c[i][it+1])/2.0); ++it; }
        if (int(nc.size()) != newsize) nc.push_back(c[i][it]);
        r.c[i] = nc; r.offset[i] = newoffset;
      }
    }
    return r;
  }
};

// ---- boxes (YGridComponent, yaspgrid/ygrid.hh:99-289) -------------------------------------------------------------
struct Box {
  I3 origin{{0,0,0}}, size{{0,0,0}};
  bool empty(int d) const { for (int i = 0; i < d; ++i) if (size[i] == 0) return true; return false; }
  int total(int d) const { int t = 1; for (int i = 0; i < d; ++i) t *= size[i]; return t; }
  int max(int i) const { return origin[i]+size[i]-1; }
  bool inside(int d, const I3& c) const { for (int i = 0; i < d; ++i) if (c[i] < origin[i] || c[i] >= origin[i]+size[i]) return false; return true; }
  Box moved(int d, const I3& v) const { Box b = *this; for (int i = 0; i < d; ++i) b.origin[i] += v[i]; return b; }
  Box intersection(int d, const Box& r) const {                                                   // :271-289
    Box e;
    if (empty(d) || r.empty(d)) return e;
    for (int i = 0; i < d; ++i) { e.origin[i] = std::max(origin[i], r.origin[i]); e.size[i] = std::min(max(i), r.max(i)) - e.origin[i] + 1; }   // may be negative: not "empty" upstream either (overlap 0)
    return e;
  }
};
// lexicographic walk over a box, axis 0 fastest (YGridComponent::Iterator, :317-400)
template<class F> void forEachCoord(int d, const Box& b, F&& f) {
  for (int i = 0; i < d; ++i) if (b.size[i] <= 0) return;          // negative sizes (overlap-0 partitions) are undefined upstream: walk nothing
  I3 c = b.origin;
  for (;;) {
    f(c);
    int i = 0;
    for (; i < d; ++i) { if (++c[i] <= b.max(i)) break; c[i] = b.origin[i]; }
    if (i == d) break;
  }
}

enum { OF = 0, OV = 1, IB = 2, IN = 3 };
struct Component { unsigned shift = 0; Box box[4]; int offset[4] = {0,0,0,0}; };            // offset: YGrid::finalize per partition kind
struct ListItem { int rank = 0, distance = 0; unsigned shift = 0; int offset = 0; Box grid; };
struct Level {
  int level = 0, overlap = 0; I3 N{{1,1,1}};
  Coordinates coords;
  std::vector<Component> comp[4];                          // per codim, ascending shift
  std::vector<ListItem> lists[8][4];                       // which x codim, see oracle_api.h
  int shiftmap[4][8];
  bool ovlpLow[3] = {false,false,false}, ovlpUp[3] = {false,false,false};
  int superindex(int codim, int which, const I3& coord, int dim, int kind = OF) const {     // ygrid.hh:114-120,764-767
    const Box& of = comp[codim][which].box[OF];
    int idx = 0, inc = 1;
    for (int i = 0; i < dim; ++i) { idx += (coord[i]-of.origin[i])*inc; inc *= of.size[i]; }
    return comp[codim][which].offset[kind] + idx;
  }
};

struct Spec { int dim, mode; double lower[3], upper[3]; I3 N; std::vector<double> tensor[3]; unsigned periodic; int overlap, partitioner; I3 fixed; int refine; bool keep; };

struct RankGrid {
  Torus torus; I3 oInterior{{0,0,0}}, sInterior{{1,1,1}};
  std::vector<Level> levels;
  int nBSegments = 0;
};

struct World {
  Spec sp; int P = 1, dim = 0; double L[3] = {0,0,0};
  std::vector<RankGrid> ranks;

  bool per(int i) const { return sp.periodic >> i & 1u; }

  // makelevel part 1 (yaspgrid.hh:327-493): the four nested boxes of every component
  void makeBoxes(RankGrid& rg, Level& g, const I3& oInt) {
    I3 oOv{{0,0,0}}, sOv{{1,1,1}};
    for (int i = 0; i < dim; ++i) {
      sOv[i] = g.coords.size(i);
      g.ovlpLow[i] = g.ovlpUp[i] = false;
      if (per(i)) { oOv[i] = oInt[i]-g.overlap; g.ovlpLow[i] = g.ovlpUp[i] = true; }
      else {
        if (oInt[i]-g.overlap < 0) oOv[i] = 0; else { oOv[i] = oInt[i]-g.overlap; g.ovlpLow[i] = true; }
        if (oOv[i] + g.coords.size(i) < g.N[i]) g.ovlpUp[i] = true;
      }
    }
    for (int codim = 0; codim <= dim; ++codim) {
      g.comp[codim].clear();
      for (unsigned index = 0; index < (1u << dim); ++index) {
        int cnt = 0; for (int i = 0; i < dim; ++i) cnt += index >> i & 1u;
        if (cnt != dim-codim) continue;
        Component c; c.shift = index;
        I3 origin = oOv, size = sOv;
        for (int i = 0; i < dim; ++i) if (!(index >> i & 1u)) size[i]++;
        c.box[OF].origin = origin; c.box[OF].size = size;
        for (int i = 0; i < dim; ++i) if (!(index >> i & 1u)) { if (g.ovlpLow[i]) { origin[i]++; size[i]--; } if (g.ovlpUp[i]) size[i]--; }
        c.box[OV].origin = origin; c.box[OV].size = size;
        for (int i = 0; i < dim; ++i) {
          const bool r = index >> i & 1u;
          if (g.ovlpLow[i]) { origin[i] += g.overlap; size[i] -= g.overlap; if (!r) { origin[i]--; size[i]++; } }
          if (g.ovlpUp[i]) { size[i] -= g.overlap; if (!r) size[i]++; }
        }
        c.box[IB].origin = origin; c.box[IB].size = size;
        for (int i = 0; i < dim; ++i) if (!(index >> i & 1u)) { if (g.ovlpLow[i]) { origin[i]++; size[i]--; } if (g.ovlpUp[i]) size[i]--; }
        c.box[IN].origin = origin; c.box[IN].size = size;
        g.shiftmap[codim][index] = int(g.comp[codim].size());
        g.comp[codim].push_back(c);
      }
      for (int kind = 0; kind < 4; ++kind) {                 // YGrid::finalize (ygrid.hh:771-791): offsets from the grid's own sizes
        int run = 0;
        for (Component& c : g.comp[codim]) { c.offset[kind] = run; run += c.box[kind].total(dim); }
      }
    }
    (void)rg;
  }

  // makelevel part 2 = YaspGrid::intersections (yaspgrid.hh:561-671) for all ranks of one level at once: every rank
  // "sends" its moved send- and recv-box to each torus neighbour; the boxes a rank receives from the neighbour at delta
  // are those that neighbour posted for -delta (k-th send to a rank pairs with the k-th receive from it, torus.hh:387-440).
  void makeLists(int lvl) {
    static const int sendKind[4] = { OF, OV, IB, IB }, recvKind[4] = { OF, OF, IB, OF };
    for (int p = 0; p < P; ++p) for (int w = 0; w < 8; ++w) for (int c = 0; c < 4; ++c) ranks[p].levels[lvl].lists[w][c].clear();
    for (int codim = 0; codim <= dim; ++codim) {
      const int ncomp = int(ranks[0].levels[lvl].comp[codim].size());
      for (int k = 0; k < ncomp; ++k)
        for (int pair = 0; pair < 4; ++pair) {
          // mailbox[p][index of p's send list entry] = (moved sendgrid, moved recvgrid, valid)
          struct Posted { Box sendgrid, recvgrid; };
          std::vector<std::vector<Posted>> mailbox(P);
          for (int p = 0; p < P; ++p) {
            const RankGrid& rg = ranks[p]; const Level& g = rg.levels[lvl];
            const Box& sendgrid = g.comp[codim][k].box[sendKind[pair]];
            const Box& recvgrid = g.comp[codim][k].box[recvKind[pair]];
            mailbox[p].assign(rg.torus.neighbors(), Posted());
            const I3 coord = rg.torus.rankToCoord(p);
            for (const CommPartner& cp : rg.torus.sendlist) {
              bool skip = false; I3 v{{0,0,0}};
              for (int i = 0; i < dim; ++i) {
                const int nb = coord[i] + cp.delta[i];
                if (nb < 0) { if (per(i)) v[i] += g.N[i]; else skip = true; }
                if (nb >= rg.torus.dims[i]) { if (per(i)) v[i] -= g.N[i]; else skip = true; }
              }
              if (!skip) { mailbox[p][cp.index].sendgrid = sendgrid.moved(dim, v); mailbox[p][cp.index].recvgrid = recvgrid.moved(dim, v); }
            }
          }
          for (int p = 0; p < P; ++p) {
            RankGrid& rg = ranks[p]; Level& g = rg.levels[lvl];
            const Box& sendgrid = g.comp[codim][k].box[sendKind[pair]];
            const Box& recvgrid = g.comp[codim][k].box[recvKind[pair]];
            std::deque<ListItem> sendlist, recvlist;
            for (const CommPartner& cp : rg.torus.recvlist) {
              // the partner's send-list entry for the opposite delta: its index is last - (recv index of -delta) and the recv
              // index of -delta is last - cp.index, hence cp.index itself
              const Posted& got = mailbox[cp.rank][cp.index];
              ListItem s; s.rank = cp.rank; s.distance = cp.distance(); s.shift = g.comp[codim][k].shift;
              s.grid = sendgrid.intersection(dim, got.recvgrid);
              if (!s.grid.empty(dim)) sendlist.push_front(s);
              ListItem r; r.rank = cp.rank; r.distance = cp.distance(); r.shift = g.comp[codim][k].shift;
              r.grid = recvgrid.intersection(dim, got.sendgrid);
              if (!r.grid.empty(dim)) recvlist.push_back(r);
            }
            // YGridList::finalize (ygrid.hh:962-1012): artificial offset = sum of supersize products of earlier components
            int off = 0; for (int j = 0; j < k; ++j) off += g.comp[codim][j].box[OF].total(dim);
            for (ListItem& it : sendlist) { it.offset = off; g.lists[2*pair][codim].push_back(it); }
            for (ListItem& it : recvlist) { it.offset = off; g.lists[2*pair+1][codim].push_back(it); }
          }
        }
    }
  }

  // `local` (tensor-product restore, yaspgrid.hh:1156-1197): local[p][i] = rank p's own level-0 coordinates incl. overlap
  typedef std::vector<std::array<std::vector<double>,3>> LocalCoords;
  World(const orc_spec& s, int nranks, const LocalCoords* local = nullptr) {
    P = nranks; dim = s.dim;
    if (dim != 2 && dim != 3) throw GridError("oracle supports dim 2 and 3");
    sp.dim = dim; sp.mode = s.coord_mode; sp.periodic = s.periodic; sp.overlap = s.overlap; sp.partitioner = s.partitioner;
    sp.refine = s.refine; sp.keep = s.keep_overlap != 0;
    for (int i = 0; i < 3; ++i) { sp.lower[i] = s.lower[i]; sp.upper[i] = s.upper[i]; sp.N[i] = i < dim ? s.N[i] : 1; sp.fixed[i] = s.fixed_dims[i]; }
    if (sp.mode == 2 && local) {
      for (const auto& rk : *local) for (int i = 0; i < dim; ++i) {
        if (rk[i].size() <= 1) throw GridError("Setup of a tensorproduct grid requires monotonous sequences of coordinates.");
        for (std::size_t j = 1; j < rk[i].size(); ++j)
          if (rk[i][j] < rk[i][j-1]) throw GridError("Setup of a tensorproduct grid requires monotonous sequences of coordinates.");
      }
    } else if (sp.mode == 2)
      for (int i = 0; i < dim; ++i) {
        sp.tensor[i].assign(s.tensor[i], s.tensor[i]+sp.N[i]+1);
        if (sp.tensor[i].size() <= 1) throw GridError("Setup of a tensorproduct grid requires monotonous sequences of coordinates.");
        for (std::size_t j = 1; j < sp.tensor[i].size(); ++j)                                 // coordinates.hh:371-383
          if (sp.tensor[i][j] < sp.tensor[i][j-1]) throw GridError("Setup of a tensorproduct grid requires monotonous sequences of coordinates.");
      }
    const I3 dims = partitionDims(dim, sp.N, P, sp.overlap, sp.partitioner, sp.fixed);        // Torus ctor, torus.hh:71-91
    { int prod = 1; for (int i = 0; i < dim; ++i) prod *= dims[i];
      if (prod != P) throw std::runtime_error("Communicator size and result of the given load balancer do not match!"); }
    ranks.resize(P);
    bool tooSmall = false;
    for (int p = 0; p < P; ++p) {
      RankGrid& rg = ranks[p];
      rg.torus.init(dim, dims, p);
      rg.torus.partition(sp.N, rg.oInterior, rg.sInterior);
      for (int i = 0; i < dim; ++i)                                                           // yaspgrid.hh:925-937
        if (rg.sInterior[i] < 2*sp.overlap && (per(i) || rg.sInterior[i] != sp.N[i])) tooSmall = true;
    }
    if (tooSmall) throw GridError("YaspGrid does not support degrees of freedom shared by more than immediately neighboring subdomains.");
    for (int i = 0; i < dim; ++i)
      L[i] = sp.mode == 0 ? sp.upper[i] : sp.mode == 1 ? sp.upper[i]-sp.lower[i]
           : local ? (*local)[0][i].back() - (*local)[0][i].front()            // _L from the LOCAL coordinates (:1172-1173)
           : sp.tensor[i][sp.N[i]] - sp.tensor[i][0];
    // level 0: constructors yaspgrid.hh:904-961 (equidistant), :974-1032 (offset), :1043-1138 (tensor product)
    for (int p = 0; p < P; ++p) {
      RankGrid& rg = ranks[p];
      Level g; g.level = 0; g.overlap = sp.overlap; g.N = sp.N; g.coords.mode = sp.mode; g.coords.dim = dim;
      const int ov = sp.overlap;
      for (int i = 0; i < dim; ++i) {
        const int oI = rg.oInterior[i], sI = rg.sInterior[i];
        if (sp.mode != 2) {
          int sOv = sI;
          if (oI - ov > 0 || per(i)) sOv += ov;
          if (oI + sI + ov <= sp.N[i] || per(i)) sOv += ov;
          if (sp.mode == 0) g.coords.setEquidistant(i, (sp.upper[i]/sp.N[i])*sOv, sOv);
          else g.coords.setOffset(i, sp.lower[i], sp.lower[i] + sOv*(sp.upper[i]-sp.lower[i])/sp.N[i], sOv);
        } else if (local) {                                                   // :1183-1190
          g.coords.c[i] = (*local)[p][i];
          g.coords.offset[i] = oI - ((per(i) || oI > 0) ? ov : 0);
        } else {
          const std::vector<double>& in = sp.tensor[i];
          int begin = oI, end = oI + sI + 1, offset = oI;
          if (oI - ov > 0) { begin -= ov; offset -= ov; }
          if (oI + sI + ov < sp.N[i]) end += ov;
          std::vector<double> nc(in.begin()+begin, in.begin()+end);
          if (per(i) && oI + sI + ov >= sp.N[i]) for (int j = 0; j < ov; ++j) nc.push_back(nc.back() - in[j] + in[j+1]);
          if (per(i) && oI - ov <= 0) { offset -= ov; int it = sp.N[i]; for (int j = 0; j < ov; ++j, --it) nc.insert(nc.begin(), nc.front() - in[it] + in[it-1]); }
          g.coords.c[i] = nc; g.coords.offset[i] = offset;
        }
      }
      makeBoxes(rg, g, rg.oInterior);
      rg.levels.push_back(g);
    }
    makeLists(0);
    for (int p = 0; p < P; ++p) {                                                             // boundarysegmentssize, yaspgrid.hh:683-707
      RankGrid& rg = ranks[p]; const Box& ov = rg.levels[0].comp[0][0].box[OV];
      rg.nBSegments = 0;
      for (int k = 0; k < dim; ++k) {
        const int sides = (ov.origin[k] == 0) + (ov.origin[k]+ov.size[k] == sp.N[k]);
        int offset = 1; for (int l = 0; l < dim; ++l) if (l != k) offset *= ov.size[l];
        rg.nBSegments += sides*offset;
      }
    }
    for (int r = 0; r < sp.refine; ++r) {                                                     // globalRefine, yaspgrid.hh:1234-1263
      const int lvl = int(ranks[0].levels.size());
      for (int p = 0; p < P; ++p) {
        RankGrid& rg = ranks[p]; const Level& cg = rg.levels.back();
        bool low[3] = {false,false,false}, up[3] = {false,false,false};
        const Box& ovb = cg.comp[0][0].box[OV];
        for (int i = 0; i < dim; ++i) { low[i] = ovb.origin[i] > 0 || per(i); up[i] = ovb.max(i)+1 < cg.N[i] || per(i); }
        Level g; g.level = lvl; g.overlap = sp.keep ? 2*cg.overlap : cg.overlap;
        for (int i = 0; i < 3; ++i) g.N[i] = i < dim ? sp.N[i] << lvl : 1;
        g.coords = cg.coords.refine(low, up, cg.overlap, sp.keep);
        I3 oInt{{0,0,0}}; for (int i = 0; i < dim; ++i) oInt[i] = 2*cg.comp[0][0].box[IN].origin[i];
        makeBoxes(rg, g, oInt);
        rg.levels.push_back(g);
      }
      makeLists(lvl);
    }
  }

  // ---- partitions and entities -------------------------------------------------------------------------------------
  static int kindOf(int pit) { switch (pit) { case 0: return IN; case 1: return IB; case 2: return OV; case 3: case 4: return OF; default: return -1; } }   // yaspgrid.hh:1819-1830
  const Level& lev(int rank, int level) const {
    if (level < 0 || level >= int(ranks[rank].levels.size())) throw std::range_error("level out of range");
    return ranks[rank].levels[level];
  }
  // vertex coordinate range of an entity (ygrid.hh:403-433) with the periodic wrap of cells and faces (yaspgridentity.hh:497-510)
  void bounds(const Level& g, const I3& coord, unsigned shift, bool wrap, double* ll, double* ur) const {
    for (int i = 0; i < dim; ++i) {
      ll[i] = g.coords.coordinate(i, coord[i]);
      ur[i] = g.coords.coordinate(i, coord[i] + int(shift >> i & 1u));
      if (wrap && per(i)) {
        if (coord[i] < 0) { ll[i] += L[i]; ur[i] += L[i]; }
        else if (coord[i]+1 > g.N[i]) { ll[i] -= L[i]; ur[i] -= L[i]; }
      }
    }
  }
  int partitionType(const Level& g, int codim, int which, const I3& c) const {               // yaspgridentity.hh:294-305,480-488
    const Component& cm = g.comp[codim][which];
    if (cm.box[IN].inside(dim, c)) return 0;
    if (codim == 0) return 2;
    if (cm.box[IB].inside(dim, c)) return 1;
    if (cm.box[OV].inside(dim, c)) return 2;
    if (cm.box[OF].inside(dim, c)) return 3;
    return 4;
  }
  template<class F> int64_t forEachEntity(int rank, int level, int codim, int pit, F&& f) const {
    const Level& g = lev(rank, level);
    if (codim < 0 || codim > dim) throw std::range_error("codim out of range");
    const int kind = kindOf(pit);
    int64_t n = 0;
    if (kind < 0) return 0;
    for (int k = 0; k < int(g.comp[codim].size()); ++k)
      forEachCoord(dim, g.comp[codim][k].box[kind], [&](const I3& c) { f(g, k, c, g.superindex(codim, k, c, dim, kind), n); ++n; });
    return n;
  }
  int subIndex(const Level& g, const I3& cell, int i, int cc) const {                          // yaspgridentity.hh:764-776
    const std::vector<SubEntity> t = subEntityTable(dim, cc);
    I3 c = cell; for (int j = 0; j < dim; ++j) c[j] += int(t[i].move >> j & 1u);
    return g.superindex(cc, g.shiftmap[cc][t[i].shift], c, dim);
  }
  // MultipleCodimMultipleGeomTypeMapper::update (common/mcmgmapper.hh:373-407)
  struct Mapper { uint32_t block[4], offset[4], n; };
  Mapper mapper(int rank, int level, const uint32_t* block) const {
    const Level& g = lev(rank, level);
    Mapper m; m.n = 0;
    for (int codim = 0; codim <= dim; ++codim) {
      m.block[codim] = block[codim];
      if (block[codim]) { m.offset[codim] = m.n; uint32_t sz = 0; for (const Component& c : g.comp[codim]) sz += uint32_t(c.box[OF].total(dim)); m.n += sz*block[codim]; }
      else m.offset[codim] = 0xffffffffu;
    }
    return m;
  }
  // intersection k of a cell (yaspgridintersection.hh:40-288)
  bool neighbor(const Level& g, const I3& c, int dir, int side) const {
    const Box& ov = g.comp[0][0].box[OV]; const int q = c[dir]+side;
    return q > ov.origin[dir] && q <= ov.max(dir);
  }
  bool boundary(const Level& g, const I3& c, int dir, int side) const {
    if (per(dir)) return false;
    const int q = c[dir]+side;
    return q == 0 || q == g.N[dir];
  }
  int boundarySegmentIndex(int rank, const Level& g, const I3& c, int dir, int side) const {   // :105-162
    const Box& b0 = ranks[rank].levels[0].comp[0][0].box[OV];
    int sides[3], pos[3], fsize[3], vol = 1;
    for (int k = 0; k < dim; ++k) { sides[k] = (b0.origin[k] == 0) + (b0.origin[k]+b0.size[k] == sp.N[k]); vol *= b0.size[k]; }
    for (int k = 0; k < dim; ++k) { pos[k] = c[k]/(1 << g.level) - b0.origin[k]; fsize[k] = vol/b0.size[k]; }
    int index = 0, localoffset = 1;
    for (int k = dim-1; k >= 0; --k) { if (k == dir) continue; index += pos[k]*localoffset; localoffset *= b0.size[k]; }
    for (int k = 0; k < dir; ++k) index += sides[k]*fsize[k];
    index += side*(sides[dir] > 1)*fsize[dir];
    return index;
  }
  void faceBounds(const Level& g, const I3& c, int dir, int side, double* ll, double* ur) const {   // :232-269
    for (int i = 0; i < dim; ++i) {
      int coord = c[i];
      if (i == dir && side) coord++;
      ll[i] = g.coords.coordinate(i, coord);
      if (i != dir) coord++;
      ur[i] = g.coords.coordinate(i, coord);
      if (per(i)) {
        if (c[i] < 0) { ll[i] += L[i]; ur[i] += L[i]; }
        else if (c[i]+1 > g.N[i]) { ll[i] -= L[i]; ur[i] -= L[i]; }
      }
    }
  }

  // ---- a10: communicate (yaspgrid.hh:1470-1730, codims dim..0 per YaspCommunicateMeta :122-143) -------------------
  int64_t communicate(int level, const uint32_t* block, int iftype, int dir, int op, int narrays, double** arrays, const int* items) {
    int sw, rw;
    switch (iftype) { case 0: sw = 4; rw = 5; break; case 1: sw = 6; rw = 7; break; case 2: case 3: sw = 2; rw = 3; break; case 4: sw = 0; rw = 1; break;
                      default: throw std::runtime_error("unknown interface"); }
    if (dir == 1) std::swap(sw, rw);
    int64_t sent = 0;
    for (int codim = dim; codim >= 0; --codim) {
      if (!block[codim]) continue;
      // all ranks gather, then all ranks scatter (Torus::exchange): message k from p to q pairs with q's k-th receive from p
      std::vector<std::vector<std::vector<double>>> msgs(P);
      for (int p = 0; p < P; ++p) {
        const Level& g = lev(p, level); const Mapper m = mapper(p, level, block);
        for (const ListItem& it : g.lists[sw][codim]) {
          std::vector<double> buf;
          forEachCoord(dim, it.grid, [&](const I3& c) {
            const Box& of = g.comp[codim][g.shiftmap[codim][it.shift]].box[OF];
            int idx = 0, inc = 1; for (int i = 0; i < dim; ++i) { idx += (c[i]-of.origin[i])*inc; inc *= of.size[i]; }
            const uint32_t base = uint32_t(it.offset + idx)*m.block[codim] + m.offset[codim];
            for (int a = 0; a < narrays; ++a)
              for (uint32_t b = 0; b < m.block[codim]; ++b)
                for (int i = 0; i < items[a]; ++i) buf.push_back(arrays[p*narrays+a][std::size_t(base+b)*items[a]+i]);
          });
          sent += int64_t(buf.size());
          msgs[p].push_back(std::move(buf));
        }
      }
      for (int p = 0; p < P; ++p) {
        const Level& g = lev(p, level); const Mapper m = mapper(p, level, block);
        std::vector<int> taken(P, 0);                               // number of messages already received from each rank
        for (const ListItem& it : g.lists[rw][codim]) {
          // the k-th message rank it.rank sent to p
          const Level& og = lev(it.rank, level);
          int seen = 0; const std::vector<double>* buf = nullptr;
          for (std::size_t j = 0; j < og.lists[sw][codim].size(); ++j)
            if (og.lists[sw][codim][j].rank == p) { if (seen == taken[it.rank]) { buf = &msgs[it.rank][j]; break; } ++seen; }
          if (!buf) throw std::runtime_error("local sends/receives do not match in exchange");
          ++taken[it.rank];
          std::size_t pos = 0;
          forEachCoord(dim, it.grid, [&](const I3& c) {
            const Box& of = g.comp[codim][g.shiftmap[codim][it.shift]].box[OF];
            int idx = 0, inc = 1; for (int i = 0; i < dim; ++i) { idx += (c[i]-of.origin[i])*inc; inc *= of.size[i]; }
            const uint32_t base = uint32_t(it.offset + idx)*m.block[codim] + m.offset[codim];
            for (int a = 0; a < narrays; ++a)
              for (uint32_t b = 0; b < m.block[codim]; ++b)
                for (int i = 0; i < items[a]; ++i) {
                  double& local = arrays[p*narrays+a][std::size_t(base+b)*items[a]+i];
                  const double remote = (*buf)[pos++];
                  if (op == 2) { if (remote >= 0) local = std::min(remote, local); }     // MinimumExchange, utility/globalindexset.hh:153-160
                  else if (op == 3) { if (remote >= 0) local = remote; }                   // IndexExchange, :262-289
                  else local = op == 1 ? remote+local : remote;
                }
          });
        }
      }
    }
    return sent;
  }

  // ---- a10, variable-size data handle (yaspgrid.hh:1562-1634: the sender asks data.size(e) per entity and sends the sizes ahead
  // of the items, :1566-1612; the receiver scatters entity by entity with the sender's n, :1712-1722) ----------------------------
  int64_t communicateVariable(int level, const uint32_t* block, int iftype, int dir, const int64_t* const* offsets, const double* const* data,
                              int64_t* nCalls, int64_t* nItems, uint32_t** rIndex, int64_t** rN, double** rData) {
    int sw, rw;
    switch (iftype) { case 0: sw = 4; rw = 5; break; case 1: sw = 6; rw = 7; break; case 2: case 3: sw = 2; rw = 3; break; case 4: sw = 0; rw = 1; break;
                      default: throw std::runtime_error("unknown interface"); }
    if (dir == 1) std::swap(sw, rw);
    for (int c = 0; c <= dim; ++c) if (block[c] > 1) throw std::runtime_error("variable-size handle: block must be 0 or 1");
    int64_t sent = 0;
    for (int p = 0; p < P; ++p) { nCalls[p] = 0; nItems[p] = 0; }
    for (int codim = dim; codim >= 0; --codim) {
      if (!block[codim]) continue;
      std::vector<std::vector<std::vector<double>>> msgs(P);           // items
      std::vector<std::vector<std::vector<std::size_t>>> sizes(P);     // the size message that travels first
      for (int p = 0; p < P; ++p) {
        const Level& g = lev(p, level); const Mapper m = mapper(p, level, block);
        for (const ListItem& it : g.lists[sw][codim]) {
          std::vector<double> buf; std::vector<std::size_t> sz;
          forEachCoord(dim, it.grid, [&](const I3& c) {
            const Box& of = g.comp[codim][g.shiftmap[codim][it.shift]].box[OF];
            int idx = 0, inc = 1; for (int i = 0; i < dim; ++i) { idx += (c[i]-of.origin[i])*inc; inc *= of.size[i]; }
            const uint32_t e = uint32_t(it.offset + idx)*m.block[codim] + m.offset[codim];
            sz.push_back(std::size_t(offsets[p][e+1] - offsets[p][e]));
            for (int64_t k = offsets[p][e]; k < offsets[p][e+1]; ++k) buf.push_back(data[p][k]);
          });
          sent += int64_t(buf.size());
          msgs[p].push_back(std::move(buf)); sizes[p].push_back(std::move(sz));
        }
      }
      for (int p = 0; p < P; ++p) {
        const Level& g = lev(p, level); const Mapper m = mapper(p, level, block);
        std::vector<int> taken(P, 0);
        for (const ListItem& it : g.lists[rw][codim]) {
          const Level& og = lev(it.rank, level);
          int seen = 0; const std::vector<double>* buf = nullptr; const std::vector<std::size_t>* sz = nullptr;
          for (std::size_t j = 0; j < og.lists[sw][codim].size(); ++j)
            if (og.lists[sw][codim][j].rank == p) { if (seen == taken[it.rank]) { buf = &msgs[it.rank][j]; sz = &sizes[it.rank][j]; break; } ++seen; }
          if (!buf) throw std::runtime_error("local sends/receives do not match in exchange");
          ++taken[it.rank];
          std::size_t pos = 0, ent = 0;
          forEachCoord(dim, it.grid, [&](const I3& c) {
            const Box& of = g.comp[codim][g.shiftmap[codim][it.shift]].box[OF];
            int idx = 0, inc = 1; for (int i = 0; i < dim; ++i) { idx += (c[i]-of.origin[i])*inc; inc *= of.size[i]; }
            const uint32_t e = uint32_t(it.offset + idx)*m.block[codim] + m.offset[codim];
            const std::size_t n = (*sz)[ent++];
            if (rIndex) rIndex[p][nCalls[p]] = e;
            if (rN) rN[p][nCalls[p]] = int64_t(n);
            for (std::size_t k = 0; k < n; ++k) { if (rData) rData[p][nItems[p]] = (*buf)[pos]; ++pos; ++nItems[p]; }
            ++nCalls[p];
          });
        }
      }
    }
    return sent;
  }

  // ---- (f2) GlobalIndexSet (dune/grid/utility/globalindexset.hh:318-440), all ranks at once ---------------------------
  int64_t globalIndex(int level, int codim, int32_t** out) {
    if (codim < 0 || codim > dim) throw std::range_error("codim out of range");
    uint32_t block[4] = {0, 0, 0, 0}; block[codim] = 1;
    const int items[1] = {1};
    const std::vector<SubEntity> tab = subEntityTable(dim, codim);
    std::vector<std::vector<double>> assignment(P), gidx(P);
    std::vector<double*> ptr(P);
    // UniqueEntityPartition (:176-203): own rank where the sub-entity is interior or border, else -1; minimum exchange
    for (int p = 0; p < P; ++p) {
      const Level& g = lev(p, level);
      std::size_t n = 0; for (const Component& c : g.comp[codim]) n += std::size_t(c.box[OF].total(dim));
      assignment[p].assign(n, -1.0); gidx[p].assign(n, -1.0);
      if (codim != 0)
        forEachEntity(p, level, 0, 4, [&](const Level&, int, const I3& cell, int, int64_t) {
          for (int i = 0; i < int(tab.size()); ++i) {
            I3 c = cell; for (int j = 0; j < dim; ++j) c[j] += int(tab[i].move >> j & 1u);
            const int pt = partitionType(g, codim, g.shiftmap[codim][tab[i].shift], c);
            assignment[p][subIndex(g, cell, i, codim)] = (pt == 0 || pt == 1) ? double(p) : -1.0;
          }
        });
      ptr[p] = assignment[p].data();
    }
    if (codim != 0) communicate(level, block, 4, 0, 2, 1, ptr.data(), items);
    // counts and offsets (:338-362)
    std::vector<int64_t> nLocal(P, 0);
    for (int p = 0; p < P; ++p) {
      if (codim == 0) nLocal[p] = forEachEntity(p, level, 0, 0, [](const Level&, int, const I3&, int, int64_t) {});
      else for (double a : assignment[p]) if (a == double(p)) ++nLocal[p];
    }
    int64_t total = 0; for (int64_t v : nLocal) total += v;
    // numbering in the order of the first visit (:384-428)
    for (int p = 0; p < P; ++p) {
      const Level& g = lev(p, level);
      int64_t myoffset = 0; for (int q = 0; q < p; ++q) myoffset += nLocal[q];
      int64_t contrib = 0;
      std::vector<bool> firstTime(gidx[p].size(), true);
      forEachEntity(p, level, 0, 4, [&](const Level&, int, const I3& cell, int cellIndex, int64_t) {
        if (codim == 0) {
          if (partitionType(g, 0, 0, cell) == 0) gidx[p][cellIndex] = double(myoffset + contrib++);
          return;
        }
        for (int i = 0; i < int(tab.size()); ++i) {
          const int idx = subIndex(g, cell, i, codim);
          if (!firstTime[idx]) continue;
          firstTime[idx] = false;
          if (assignment[p][idx] == double(p)) gidx[p][idx] = double(myoffset + contrib++);
        }
      });
      ptr[p] = gidx[p].data();
    }
    communicate(level, block, 4, 0, 3, 1, ptr.data(), items);             // IndexExchange (:430-431)
    for (int p = 0; p < P; ++p) for (std::size_t e = 0; e < gidx[p].size(); ++e) out[p][e] = int32_t(gidx[p][e]);
    return total;
  }

  // ---- a11: FV transport step on every rank, then communicate(InteriorBorder_All, Forward) ----------
  double fvStep(int level, int problem, const double* params, double** c, double** upd) {
    std::vector<std::vector<double>> update(P);
    double dt = std::numeric_limits<double>::infinity();
    for (int p = 0; p < P; ++p) {
      const Level& g = lev(p, level);
      update[p].assign(g.comp[0][0].box[OF].total(dim), 0.0);
      forEachCoord(dim, g.comp[0][0].box[IN], [&](const I3& cell) {
        const int i = g.superindex(0, 0, cell, dim);
        double ll[3], ur[3]; bounds(g, cell, (1u << dim)-1, true, ll, ur);
        double volume = 1.0; for (int k = 0; k < dim; ++k) volume *= ur[k]-ll[k];
        double sumfactor = 0.0, u = 0.0;
        for (int k = 0; k < 2*dim; ++k) {
          const int dir = k/2, side = k % 2;
          double fl[3], fu[3]; faceBounds(g, cell, dir, side, fl, fu);
          double fc[3], area = 1.0, vel[3];
          for (int q = 0; q < dim; ++q) { if (q == dir) fu[q] = fl[q]; fc[q] = 0.5*(fl[q]+fu[q]); if (q != dir) area *= fu[q]-fl[q]; }
          orc::fv_velocity(problem, params, dim, fc, vel);
          double vn = 0.0; for (int q = 0; q < dim; ++q) vn += vel[q]*(q == dir ? (side ? 1.0 : -1.0) : 0.0);
          const double factor = vn*area/volume;
          if (factor >= 0) { sumfactor += factor; u -= c[p][i]*factor; }
          else if (neighbor(g, cell, dir, side)) { I3 o = cell; o[dir] += side ? 1 : -1; u -= c[p][g.superindex(0, 0, o, dim)]*factor; }
          else if (boundary(g, cell, dir, side)) u -= orc::fv_boundary(problem, params, dim, fc)*factor;
        }
        update[p][i] = u;
        dt = std::min(dt, 1.0/sumfactor);
      });
    }
    dt *= 0.99;
    for (int p = 0; p < P; ++p) {
      const Level& g = lev(p, level);
      forEachCoord(dim, g.comp[0][0].box[IN], [&](const I3& cell) { const int i = g.superindex(0, 0, cell, dim); c[p][i] += dt*update[p][i]; });
      if (upd && upd[p]) std::memcpy(upd[p], update[p].data(), update[p].size()*sizeof(double));
    }
    const uint32_t block[4] = {1,0,0,0}; const int items[1] = {1};
    communicate(level, block, 1, 0, 0, 1, c, items);
    return dt;
  }
};

template<class F, class R> R guarded(R onerr, F&& f) {
  try { g_err.clear(); return f(); }
  catch (const std::exception& e) { g_err = e.what(); if (g_err.empty()) g_err = "exception"; }
  catch (...) { g_err = "unknown exception"; }
  return onerr;
}
inline World* W(void* w) { return static_cast<World*>(w); }

} // namespace

extern "C" {

const char* orc_kind(void) { return "port"; }
const char* orc_last_error(void) { return g_err.c_str(); }

int orc_partition(int dim, const int* N, int P, int overlap, int partitioner, const int* fixed, int* dims_out) {
  return guarded(-1, [&] {
    if (dim < 1 || dim > 3) throw std::runtime_error("dim 1..3");
    I3 size{{1,1,1}}, fx{{1,1,1}};
    for (int i = 0; i < dim; ++i) { size[i] = N[i]; fx[i] = fixed ? fixed[i] : 1; }
    const I3 d = partitionDims(dim, size, P, overlap, partitioner, fx);
    for (int i = 0; i < dim; ++i) dims_out[i] = d[i];
    return 0;
  });
}
int orc_entity_shift(int dim, int i, int cc) { const auto t = subEntityTable(dim, cc); return i >= 0 && i < int(t.size()) ? int(t[i].shift) : -1; }
int orc_entity_move(int dim, int i, int cc) { const auto t = subEntityTable(dim, cc); return i >= 0 && i < int(t.size()) ? int(t[i].move) : -1; }
int orc_sub_entities(int /*dim*/, int d, int c) { return subEnt(d, c); }

void* orc_create(const orc_spec* spec, int nranks) { return guarded((void*)nullptr, [&]() -> void* { return new World(*spec, nranks); }); }
void orc_destroy(void* w) { delete W(w); }
int orc_max_level(void* w) { return int(W(w)->ranks[0].levels.size())-1; }
void orc_torus_dims(void* w, int* dims) { for (int i = 0; i < W(w)->dim; ++i) dims[i] = W(w)->ranks[0].torus.dims[i]; }
void orc_domain_size(void* w, double* L) { for (int i = 0; i < W(w)->dim; ++i) L[i] = W(w)->L[i]; }
int64_t orc_size(void* w, int rank, int level, int codim) {
  return guarded(int64_t(-1), [&]() -> int64_t { int64_t n = 0; for (const Component& c : W(w)->lev(rank, level).comp[codim]) n += c.box[OF].total(W(w)->dim); return n; });
}
int orc_num_boundary_segments(void* w, int rank) { return W(w)->ranks[rank].nBSegments; }
int orc_overlap_size(void* w, int level) { return guarded(-1, [&] { return W(w)->lev(0, level).overlap; }); }

int orc_boxes(void* w, int rank, int level, int codim, int* out) {
  return guarded(-1, [&] {
    const World& wd = *W(w); const Level& g = wd.lev(rank, level);
    int k = 0;
    for (const Component& c : g.comp[codim]) {
      int* o = out + 26*k++;
      o[0] = int(c.shift); o[1] = c.offset[OF];
      for (int q = 0; q < 4; ++q) for (int i = 0; i < 3; ++i) { o[2+6*q+i] = i < wd.dim ? c.box[q].origin[i] : 0; o[2+6*q+3+i] = i < wd.dim ? c.box[q].size[i] : 1; }
    }
    return k;
  });
}
int orc_list(void* w, int rank, int level, int codim, int which, int* out, int cap) {
  return guarded(-1, [&] {
    const World& wd = *W(w); const Level& g = wd.lev(rank, level);
    if (which < 0 || which > 7) return -1;
    int n = 0;
    for (const ListItem& it : g.lists[which][codim]) {
      if (n < cap) { int* o = out + 10*n; o[0] = it.rank; o[1] = it.distance; o[2] = int(it.shift); o[3] = it.offset;
                     for (int i = 0; i < 3; ++i) { o[4+i] = i < wd.dim ? it.grid.origin[i] : 0; o[7+i] = i < wd.dim ? it.grid.size[i] : 1; } }
      ++n;
    }
    return n;
  });
}
int orc_coordinates(void* w, int rank, int level, int axis, double* out, int* first) {
  return guarded(-1, [&] {
    const World& wd = *W(w); const Level& g = wd.lev(rank, level);
    const Box& b = g.comp[wd.dim][0].box[OF];
    *first = b.origin[axis];
    for (int i = 0; i < b.size[axis]; ++i) out[i] = g.coords.coordinate(axis, b.origin[axis]+i);
    return b.size[axis];
  });
}
int64_t orc_entities(void* w, int rank, int level, int codim, int pitype, uint32_t* index, uint8_t* ptype, int32_t* coord,
                     double* lower, double* upper, double* center, double* volume) {
  return guarded(int64_t(-1), [&]() -> int64_t {
    const World& wd = *W(w); const int dim = wd.dim;
    return wd.forEachEntity(rank, level, codim, pitype, [&](const Level& g, int k, const I3& c, int idx, int64_t n) {
      if (index) index[n] = uint32_t(idx);
      if (ptype) ptype[n] = uint8_t(wd.partitionType(g, codim, k, c));
      if (coord) for (int i = 0; i < dim; ++i) coord[n*dim+i] = c[i];
      if (lower || upper || center || volume) {
        const unsigned shift = g.comp[codim][k].shift;
        double ll[3] = {0,0,0}, ur[3] = {0,0,0}; wd.bounds(g, c, shift, codim == 0, ll, ur);                       // only cells are wrapped (yaspgridentity.hh:270-274 vs 491-514)
        double vol = 1.0;
        for (int i = 0; i < dim; ++i) {
          if (!(shift >> i & 1u)) ur[i] = ll[i];
          if (lower) lower[n*dim+i] = ll[i];
          if (upper) upper[n*dim+i] = ur[i];
          if (center) center[n*dim+i] = codim == dim ? ll[i] : 0.5*(ll[i]+ur[i]);
          if (shift >> i & 1u) vol *= ur[i]-ll[i];
        }
        if (volume) volume[n] = vol;
      }
    });
  });
}
int64_t orc_subindices(void* w, int rank, int level, int pitype, int cc, uint32_t* out) {
  return guarded(int64_t(-1), [&]() -> int64_t {
    const World& wd = *W(w); const int ns = subEnt(wd.dim, cc);
    return wd.forEachEntity(rank, level, 0, pitype, [&](const Level& g, int, const I3& c, int, int64_t n) {
      for (int i = 0; i < ns; ++i) out[n*ns+i] = uint32_t(wd.subIndex(g, c, i, cc));
    });
  });
}
int64_t orc_mapper_size(void* w, int rank, int level, const uint32_t* block) { return guarded(int64_t(-1), [&]() -> int64_t { return W(w)->mapper(rank, level, block).n; }); }
int64_t orc_mapper_index(void* w, int rank, int level, const uint32_t* block, int codim, int pitype, uint32_t* out) {
  return guarded(int64_t(-1), [&]() -> int64_t {
    const World& wd = *W(w); const World::Mapper m = wd.mapper(rank, level, block);
    if (!m.block[codim]) throw std::runtime_error("mapper layout does not contain this codimension");
    return wd.forEachEntity(rank, level, codim, pitype, [&](const Level&, int, const I3&, int idx, int64_t n) { out[n] = uint32_t(idx)*m.block[codim] + m.offset[codim]; });
  });
}
int64_t orc_mapper_subindex(void* w, int rank, int level, const uint32_t* block, int pitype, int cc, uint32_t* out) {
  return guarded(int64_t(-1), [&]() -> int64_t {
    const World& wd = *W(w); const World::Mapper m = wd.mapper(rank, level, block); const int ns = subEnt(wd.dim, cc);
    if (!m.block[cc]) throw std::runtime_error("mapper layout does not contain this codimension");
    return wd.forEachEntity(rank, level, 0, pitype, [&](const Level& g, int, const I3& c, int, int64_t n) {
      for (int i = 0; i < ns; ++i) out[n*ns+i] = uint32_t(wd.subIndex(g, c, i, cc))*m.block[cc] + m.offset[cc];
    });
  });
}
int64_t orc_intersections(void* w, int rank, int level, int pitype, uint8_t* neighbor, uint8_t* boundary, int32_t* outside_index, int32_t* bsi,
                          double* normal, double* face_center, double* face_volume) {
  return guarded(int64_t(-1), [&]() -> int64_t {
    const World& wd = *W(w); const int dim = wd.dim;
    return wd.forEachEntity(rank, level, 0, pitype, [&](const Level& g, int, const I3& c, int, int64_t n) {
      for (int f = 0; f < 2*dim; ++f) {
        const int dir = f/2, side = f % 2; const int64_t k = n*2*dim + f;
        const bool nb = wd.neighbor(g, c, dir, side), bd = wd.boundary(g, c, dir, side);
        if (neighbor) neighbor[k] = nb;
        if (boundary) boundary[k] = bd;
        if (outside_index) { I3 o = c; o[dir] += side ? 1 : -1; outside_index[k] = nb ? g.superindex(0, 0, o, dim) : -1; }
        if (bsi) bsi[k] = bd ? wd.boundarySegmentIndex(rank, g, c, dir, side) : -1;
        if (normal) for (int i = 0; i < dim; ++i) normal[k*dim+i] = i == dir ? (side ? 1.0 : -1.0) : 0.0;
        if (face_center || face_volume) {
          double ll[3] = {0,0,0}, ur[3] = {0,0,0}; wd.faceBounds(g, c, dir, side, ll, ur);
          double vol = 1.0;
          for (int i = 0; i < dim; ++i) { if (i == dir) ur[i] = ll[i]; else vol *= ur[i]-ll[i]; if (face_center) face_center[k*dim+i] = 0.5*(ll[i]+ur[i]); }
          if (face_volume) face_volume[k] = vol;
        }
      }
    });
  });
}
// AxisAlignedCubeGeometry<mydim,dim> of a codim entity: the axes of its shift span it, the others are flat at the lower corner
int64_t orc_geometry_eval_codim(void* w, int rank, int level, int codim, int pitype, int nqp, const double* qp, double* global, double* jt, double* jit, double* detJ) {
  return guarded(int64_t(-1), [&]() -> int64_t {
    const World& wd = *W(w); const int dim = wd.dim, mydim = dim-codim;
    return wd.forEachEntity(rank, level, codim, pitype, [&](const Level& g, int kcomp, const I3& c, int, int64_t n) {
      const unsigned shift = g.comp[codim][kcomp].shift;
      double ll[3] = {0,0,0}, ur[3] = {0,0,0}; wd.bounds(g, c, shift, codim == 0, ll, ur);
      double vol = 1.0;
      for (int i = 0; i < dim; ++i) { if (!(shift >> i & 1u)) ur[i] = ll[i]; else vol *= ur[i]-ll[i]; }
      for (int q = 0; q < nqp; ++q) {
        const int64_t k = n*nqp + q;
        int lc = 0;
        for (int i = 0; i < dim; ++i) {
          const bool sp = shift >> i & 1u;
          if (global) global[k*dim+i] = sp ? ll[i] + qp[q*mydim+lc]*(ur[i]-ll[i]) : ll[i];
          if (jt) for (int r = 0; r < mydim; ++r) jt[(k*mydim+r)*dim+i] = (sp && r == lc) ? ur[i]-ll[i] : 0.0;
          if (jit) for (int r = 0; r < mydim; ++r) jit[(k*dim+i)*mydim+r] = (sp && r == lc) ? 1.0/(ur[i]-ll[i]) : 0.0;
          if (sp) ++lc;
        }
        if (detJ) detJ[k] = vol;
      }
    });
  });
}
int64_t orc_geometry_eval(void* w, int rank, int level, int pitype, int nqp, const double* qp, double* global, double* jit, double* detJ) {
  return guarded(int64_t(-1), [&]() -> int64_t {
    const World& wd = *W(w); const int dim = wd.dim;
    return wd.forEachEntity(rank, level, 0, pitype, [&](const Level& g, int, const I3& c, int, int64_t n) {
      double ll[3], ur[3]; wd.bounds(g, c, (1u << dim)-1, true, ll, ur);
      double vol = 1.0; for (int i = 0; i < dim; ++i) vol *= ur[i]-ll[i];
      for (int q = 0; q < nqp; ++q) {
        const int64_t k = n*nqp + q;
        if (global) for (int i = 0; i < dim; ++i) global[k*dim+i] = ll[i] + qp[q*dim+i]*(ur[i]-ll[i]);
        if (jit) for (int i = 0; i < dim; ++i) for (int j = 0; j < dim; ++j) jit[(k*dim+i)*dim+j] = i == j ? 1.0/(ur[i]-ll[i]) : 0.0;
        if (detJ) detJ[k] = vol;
      }
    });
  });
}
} // extern "C"
template<int dim>
int64_t geogridEval(const World& wd, int rank, int level, int pitype, int def, const double* params, int nqp, const double* qp,
                    double* corners, double* global, double* jt, double* jit, double* detJ) {
  constexpr int nc = 1 << dim;
  return wd.forEachEntity(rank, level, 0, pitype, [&](const Level& g, int, const I3& cell, int, int64_t n) {
    double ll[3], ur[3]; wd.bounds(g, cell, (1u << dim)-1, true, ll, ur);
    double c[nc][dim];
    for (int k = 0; k < nc; ++k) {                                                             // cornerstorage.hh:54-60: F(hostCorner_k)
      double x[dim]; for (int i = 0; i < dim; ++i) x[i] = (k >> i & 1) ? ur[i] : ll[i];
      orc::deformation(def, params, dim, x, c[k]);
      if (corners) for (int i = 0; i < dim; ++i) corners[(n*nc+k)*dim+i] = c[k][i];
    }
    for (int q = 0; q < nqp; ++q) {
      const int64_t k = n*nqp + q; const double* x = qp + q*dim;
      double J[dim][dim], Ji[dim][dim], y[dim];
      if (global) { orc::multilinear_global<dim>(c, x, y); for (int i = 0; i < dim; ++i) global[k*dim+i] = y[i]; }
      orc::multilinear_jt<dim>(c, x, J);
      if (jt) for (int i = 0; i < dim; ++i) for (int j = 0; j < dim; ++j) jt[(k*dim+i)*dim+j] = J[i][j];
      if (jit) { orc::right_inv<dim>(J, Ji); for (int i = 0; i < dim; ++i) for (int j = 0; j < dim; ++j) jit[(k*dim+i)*dim+j] = Ji[i][j]; }
      if (detJ) detJ[k] = orc::sqrt_det_aat<dim>(J);
    }
  });
}
// GeoGrid::Intersection over the Yasp host (geometrygrid/intersection.hh:103-172): face corners = insideGeo.global(
// hostLocalGeometry.corner(i)) (cornerstorage.hh:148-154, local geometry of yaspgridintersection.hh:195-209), face geometry =
// multilinear map over them, normals = J^-T(x) n_ref (|det J(x)|) at x = geometryInInside().global(local)
template<int dim>
int64_t geogridIntersections(const World& wd, int rank, int level, int pitype, int def, const double* params, int nqp, const double* qp,
                             double* corners, double* center, double* volume, double* ion, double* on, double* uon, double* cuon) {
  constexpr int nc = 1 << dim, nfc = 1 << (dim-1);
  return wd.forEachEntity(rank, level, 0, pitype, [&](const Level& g, int, const I3& cell, int, int64_t n) {
    double ll[3], ur[3]; wd.bounds(g, cell, (1u << dim)-1, true, ll, ur);
    double c[nc][dim];
    for (int k = 0; k < nc; ++k) {
      double x[dim]; for (int i = 0; i < dim; ++i) x[i] = (k >> i & 1) ? ur[i] : ll[i];
      orc::deformation(def, params, dim, x, c[k]);
    }
    auto normals = [&](int dir, int side, const double* xl, double* vI, double* vO, double* vU) {
      double x[dim]; int b = 0;
      for (int i = 0; i < dim; ++i) x[i] = (i == dir) ? double(side) : 0.0 + xl[b++]*(1.0 - 0.0);      // AxisAlignedCubeGeometry::global
      double J[dim][dim], Ji[dim][dim], refn[dim], nrm[dim];
      orc::multilinear_jt<dim>(c, x, J);
      const double detInv = orc::right_inv<dim>(J, Ji);
      for (int i = 0; i < dim; ++i) refn[i] = 0.0;
      refn[dir] = side ? 1.0 : -1.0;
      for (int i = 0; i < dim; ++i) { nrm[i] = 0; for (int j = 0; j < dim; ++j) nrm[i] += Ji[i][j]*refn[j]; }      // jit.mv
      if (vO) for (int i = 0; i < dim; ++i) vO[i] = nrm[i];
      if (vI) for (int i = 0; i < dim; ++i) vI[i] = nrm[i]*detInv;
      if (vU) { double s = 0; for (int i = 0; i < dim; ++i) s += nrm[i]*nrm[i]; const double f = 1.0/std::sqrt(s); for (int i = 0; i < dim; ++i) vU[i] = nrm[i]*f; }
    };
    for (int f = 0; f < 2*dim; ++f) {
      const int dir = f/2, side = f % 2; const int64_t k = n*2*dim + f;
      double fc[nfc][dim];
      for (int i = 0; i < nfc; ++i) {
        double xl[dim]; int b = 0;
        for (int ax = 0; ax < dim; ++ax) xl[ax] = (ax == dir) ? double(side) : double((i >> b++) & 1);
        orc::multilinear_global<dim>(c, xl, fc[i]);
        if (corners) for (int d = 0; d < dim; ++d) corners[(k*nfc+i)*dim+d] = fc[i][d];
      }
      double half[dim > 1 ? dim-1 : 1]; for (int i = 0; i < dim-1; ++i) half[i] = 0.5;
      if (center) { double y[dim]; const double (*cit)[dim] = fc; orc::ml_global<dim>(dim-1, cit, half, 1.0, false, y); for (int d = 0; d < dim; ++d) center[k*dim+d] = y[d]; }
      if (volume) { double jt[dim > 1 ? dim-1 : 1][dim]; const double (*cit)[dim] = fc; orc::ml_jt<dim>(dim-1, cit, half, 1.0, false, jt); volume[k] = orc::sqrt_det_aat_rect<dim-1,dim>(jt); }
      for (int q = 0; q < nqp; ++q)
        normals(dir, side, qp + q*(dim-1), ion ? ion + (k*nqp+q)*dim : nullptr, on ? on + (k*nqp+q)*dim : nullptr, uon ? uon + (k*nqp+q)*dim : nullptr);
      if (cuon) normals(dir, side, half, nullptr, nullptr, cuon + k*dim);
    }
  });
}
extern "C" {
int64_t orc_geogrid_intersections(void* w, int rank, int level, int pitype, int deformation, const double* params, int nqp, const double* qp,
                                  double* corners, double* center, double* volume, double* ion, double* on, double* uon, double* cuon) {
  return guarded(int64_t(-1), [&]() -> int64_t {
    const World& wd = *W(w);
    return wd.dim == 2 ? geogridIntersections<2>(wd, rank, level, pitype, deformation, params, nqp, qp, corners, center, volume, ion, on, uon, cuon)
                       : geogridIntersections<3>(wd, rank, level, pitype, deformation, params, nqp, qp, corners, center, volume, ion, on, uon, cuon);
  });
}
int64_t orc_geogrid_eval(void* w, int rank, int level, int pitype, int deformation, const double* params, int nqp, const double* qp,
                         double* corners, double* global, double* jt, double* jit, double* detJ) {
  return guarded(int64_t(-1), [&]() -> int64_t {
    const World& wd = *W(w);
    return wd.dim == 2 ? geogridEval<2>(wd, rank, level, pitype, deformation, params, nqp, qp, corners, global, jt, jit, detJ)
                       : geogridEval<3>(wd, rank, level, pitype, deformation, params, nqp, qp, corners, global, jt, jit, detJ);
  });
}
int64_t orc_communicate(void* w, int level, const uint32_t* block, int iftype, int dir, int op, int narrays, double** arrays, const int* item_sizes) {
  return guarded(int64_t(-1), [&]() -> int64_t { return W(w)->communicate(level, block, iftype, dir, op, narrays, arrays, item_sizes); });
}
int64_t orc_communicate_variable(void* w, int level, const uint32_t* block, int iftype, int dir, const int64_t* const* offsets,
                                 const double* const* data, int64_t* n_calls, int64_t* n_items, uint32_t** recv_index, int64_t** recv_n, double** recv_data) {
  return guarded(int64_t(-1), [&]() -> int64_t { return W(w)->communicateVariable(level, block, iftype, dir, offsets, data, n_calls, n_items, recv_index, recv_n, recv_data); });
}
// (f3) BackupRestoreFacility::backup, yaspgrid/backuprestore.hh:105-128 (equidistant, offset), :222-246 (tensor product:
// TensorProductCoordinates::print of the rank's level-0 container, coordinates.hh:347-356); default ostream formatting
int64_t orc_global_index(void* w, int level, int codim, int32_t** out) {
  return guarded(int64_t(-1), [&] { return W(w)->globalIndex(level, codim, out); });
}
int orc_backup(void* w, int rank, char* out, int cap) {
  return guarded(-1, [&] {
    World& W_ = *W(w); const RankGrid& rg = W_.ranks[rank]; const int dim = W_.dim;
    std::ostringstream s;
    s << "YaspGrid BackupRestore Format Version: " << 2 << std::endl;
    s << "Torus structure: ";
    for (int i = 0; i < dim; ++i) s << rg.torus.dims[i] << " ";
    s << std::endl << "Refinement level: " << int(rg.levels.size())-1 << std::endl;
    s << "Periodicity: ";
    for (int i = 0; i < dim; ++i) s << (W_.per(i) ? "1 " : "0 ");
    s << std::endl << "Overlap: " << rg.levels[0].overlap << std::endl;
    s << "KeepPhysicalOverlap: ";
    for (std::size_t l = 1; l < rg.levels.size(); ++l) s << (W_.sp.keep ? "1" : "0") << " ";
    s << std::endl;
    s << "Coarse Size: ";
    for (int i = 0; i < dim; ++i) s << W_.sp.N[i] << " ";
    s << std::endl;
    const Coordinates& cc = rg.levels[0].coords;
    if (cc.mode != 2) {
      s << "Meshsize: ";
      for (int i = 0; i < dim; ++i) s << cc.h[i] << " ";
      s << std::endl;
      if (cc.mode == 1) { s << "Origin: "; for (int i = 0; i < dim; ++i) s << cc.origin[i] << " "; s << std::endl; }
    } else {
      s << "Printing TensorProduct Coordinate information:" << std::endl;
      for (int i = 0; i < dim; ++i) {
        s << "Direction " << i << ": " << cc.c[i].size() << " coordinates" << std::endl;
        for (std::size_t j = 0; j < cc.c[i].size(); ++j) s << cc.c[i][j] << std::endl;
      }
    }
    const std::string t = s.str();
    if (out && cap > int(t.size())) std::memcpy(out, t.c_str(), t.size()+1);
    return int(t.size());
  });
}
// BackupRestoreFacility::restore, :146-218 and :268-346 (tensor product: rank-local coordinates through the private
// constructor yaspgrid.hh:1156-1197). One keepOverlap flag for all levels (what orc_spec can express).
void* orc_restore(int dim, int coord_mode, int nranks, const char* const* texts) {
  return guarded((void*)nullptr, [&]() -> void* {
    std::istringstream stream(texts[0]);
    std::string input; int version = 0;
    stream >> input >> input >> input >> input; stream >> version;
    if (version != 2) throw std::runtime_error("Your YaspGrid backup file is written in an outdated format!");
    orc_spec sp; std::memset(&sp, 0, sizeof(sp));
    sp.dim = dim; sp.coord_mode = coord_mode; sp.partitioner = 2;
    stream >> input >> input;
    for (int i = 0; i < 3; ++i) sp.fixed_dims[i] = 1;
    for (int i = 0; i < dim; ++i) stream >> sp.fixed_dims[i];
    stream >> input >> input; stream >> sp.refine;
    bool b; stream >> input;
    for (int i = 0; i < dim; ++i) { stream >> b; if (b) sp.periodic |= 1u << i; }
    stream >> input; stream >> sp.overlap;
    stream >> input; sp.keep_overlap = 1;
    for (int i = 0; i < sp.refine; ++i) {
      stream >> b;
      if (i == 0) sp.keep_overlap = b; else if (int(b) != sp.keep_overlap) throw std::runtime_error("orc_restore: mixed keepOverlap flags are not restated in the port oracle");
    }
    stream >> input >> input;
    for (int i = 0; i < 3; ++i) sp.N[i] = 1;
    for (int i = 0; i < dim; ++i) stream >> sp.N[i];
    if (coord_mode == 2) {                                                     // :318-334, one text per rank
      World::LocalCoords local(nranks);
      for (int p = 0; p < nranks; ++p) {
        std::istringstream rs(texts[p]);
        std::string tok;
        while (rs >> tok && tok != "information:") {}                            // "... Printing TensorProduct Coordinate information:"
        for (int d = 0; d < dim; ++d) {
          int size = 0;
          rs >> tok >> tok; rs >> size; rs >> tok;
          local[p][d].resize(size);
          for (int i = 0; i < size; ++i) rs >> local[p][d][i];
        }
      }
      return new World(sp, nranks, &local);
    }
    double h[3] = {0,0,0}, origin[3] = {0,0,0};
    stream >> input; for (int i = 0; i < dim; ++i) stream >> h[i];
    if (coord_mode == 1) { stream >> input; for (int i = 0; i < dim; ++i) stream >> origin[i]; }
    for (int i = 0; i < dim; ++i) {
      double length = h[i]; length *= sp.N[i];
      if (coord_mode == 0) sp.upper[i] = length; else { sp.lower[i] = origin[i]; double ur = origin[i]; ur += length; sp.upper[i] = ur; }
    }
    return new World(sp, nranks);
  });
}

// the VTK file syntax is checked against the reference's own writer classes only (oracle/_ref); the port has no restatement of it
int64_t orc_vtu_piece(int, unsigned, unsigned, int, const int*, const char* const*, const int*, const int*, const int*, const double* const*,
                      const char*, const char*, const char*, const char*, char*, int64_t) {
  g_err = "orc_vtu_piece: reference library only";
  return -1;
}
int orc_fv_init(void* w, int rank, int level, int problem, const double* params, double* c) {
  return guarded(-1, [&] {
    const World& wd = *W(w); const int dim = wd.dim;
    wd.forEachEntity(rank, level, 0, 4, [&](const Level& g, int, const I3& cell, int idx, int64_t) {
      double ll[3], ur[3], x[3]; wd.bounds(g, cell, (1u << dim)-1, true, ll, ur);
      for (int i = 0; i < dim; ++i) x[i] = 0.5*(ll[i]+ur[i]);
      c[idx] = orc::fv_initial(problem, params, dim, x);
    });
    return 0;
  });
}
double orc_fv_step(void* w, int level, int problem, const double* params, double** c, double** upd) {
  return guarded(-1.0, [&] { return W(w)->fvStep(level, problem, params, c, upd); });
}
double orc_fv_run(void* w, int level, int problem, const double* params, double** c, int steps, double* dt_last) {
  return guarded(-1.0, [&] {
    const auto t0 = std::chrono::steady_clock::now();
    double dt = 0.0;
    for (int s = 0; s < steps; ++s) dt = W(w)->fvStep(level, problem, params, c, nullptr);
    if (dt_last) *dt_last = dt;
    return std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();
  });
}

} // extern "C"
