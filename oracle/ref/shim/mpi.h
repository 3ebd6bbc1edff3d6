/* In-process stand-in for <mpi.h> (no MPI in this image). Each "rank" is a thread of the oracle
 * process; point-to-point messages are buffered FIFO queues per (source,destination) pair, which
 * preserves MPI's non-overtaking order that Torus::exchange relies on (reference
 * dune/grid/yaspgrid/torus.hh:270-275,407-439). Test infrastructure only. */
#ifndef SHIM_VMPI_H
#define SHIM_VMPI_H
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <vector>

struct vmpi_world {
  int size;
  std::mutex mtx;
  std::condition_variable cv;
  std::vector<std::deque<std::vector<char>>> box;   // box[src*size+dst]
  // collective scratch
  int arrived = 0; long generation = 0;
  std::vector<std::vector<char>> contrib, snapshot_;
  explicit vmpi_world(int n) : size(n), box(size_t(n)*n), contrib(n) {}
};
struct vmpi_rank { vmpi_world* world; int rank; };
typedef vmpi_rank* MPI_Comm;

typedef int MPI_Datatype;
enum { MPI_BYTE = 1, MPI_INT = 4, MPI_DOUBLE = 8 };
typedef int MPI_Op;
enum { MPI_LOR = 1, MPI_MAX = 2, MPI_MIN = 3, MPI_SUM = 4 };
struct MPI_Request { int kind; void* buf; int bytes; int peer; MPI_Comm comm; };
struct MPI_Status { int dummy; };
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_SUCCESS 0

inline int vmpi_typesize(MPI_Datatype t) { return t; }

inline int MPI_Isend(const void* buf, int count, MPI_Datatype t, int dest, int /*tag*/, MPI_Comm c, MPI_Request* r) {
  vmpi_world* w = c->world;
  std::vector<char> msg((const char*)buf, (const char*)buf + size_t(count)*vmpi_typesize(t));
  {
    std::lock_guard<std::mutex> g(w->mtx);
    w->box[size_t(c->rank)*w->size + dest].push_back(std::move(msg));
  }
  w->cv.notify_all();
  r->kind = 0; r->buf = 0; r->bytes = 0; r->peer = dest; r->comm = c;
  return MPI_SUCCESS;
}
inline int MPI_Irecv(void* buf, int count, MPI_Datatype t, int src, int /*tag*/, MPI_Comm c, MPI_Request* r) {
  r->kind = 1; r->buf = buf; r->bytes = count*vmpi_typesize(t); r->peer = src; r->comm = c;
  return MPI_SUCCESS;
}
inline int MPI_Waitall(int n, MPI_Request* reqs, MPI_Status*) {
  for (int i = 0; i < n; ++i) {
    MPI_Request& r = reqs[i];
    if (r.kind != 1) continue;
    vmpi_world* w = r.comm->world;
    std::unique_lock<std::mutex> lk(w->mtx);
    auto& q = w->box[size_t(r.peer)*w->size + r.comm->rank];
    w->cv.wait(lk, [&]{ return !q.empty(); });
    std::vector<char> msg = std::move(q.front());
    q.pop_front();
    lk.unlock();
    std::memcpy(r.buf, msg.data(), std::min<size_t>(msg.size(), size_t(r.bytes)));
  }
  return MPI_SUCCESS;
}
/* all ranks deposit `bytes` bytes; returns every rank's contribution (rank order) in `all` */
inline void vmpi_allgather(MPI_Comm c, const void* in, int bytes, std::vector<std::vector<char>>& all) {
  vmpi_world* w = c->world;
  std::unique_lock<std::mutex> lk(w->mtx);
  long gen = w->generation;
  w->contrib[c->rank].assign((const char*)in, (const char*)in + bytes);
  if (++w->arrived == w->size) {
    all = w->contrib;           // last arriver: snapshot, then release the others
    w->snapshot_ = all;
    w->arrived = 0; ++w->generation;
    lk.unlock(); w->cv.notify_all();
  } else {
    w->cv.wait(lk, [&]{ return w->generation != gen; });
    all = w->snapshot_;
  }
}
template<class T, class F>
inline void vmpi_reduce(MPI_Comm c, const T* in, T* out, int n, F f) {
  std::vector<std::vector<char>> all;
  vmpi_allgather(c, in, int(sizeof(T))*n, all);
  for (int i = 0; i < n; ++i) {
    T acc; std::memcpy(&acc, all[0].data()+sizeof(T)*i, sizeof(T));
    for (size_t r = 1; r < all.size(); ++r) { T v; std::memcpy(&v, all[r].data()+sizeof(T)*i, sizeof(T)); acc = f(acc, v); }
    out[i] = acc;
  }
}
inline int MPI_Allreduce(const void* in, void* out, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  if (t == MPI_INT) {
    if (op == MPI_LOR) vmpi_reduce<int>(c, (const int*)in, (int*)out, n, [](int a, int b){ return int(a||b); });
    else if (op == MPI_MAX) vmpi_reduce<int>(c, (const int*)in, (int*)out, n, [](int a, int b){ return a>b?a:b; });
    else if (op == MPI_MIN) vmpi_reduce<int>(c, (const int*)in, (int*)out, n, [](int a, int b){ return a<b?a:b; });
    else vmpi_reduce<int>(c, (const int*)in, (int*)out, n, [](int a, int b){ return a+b; });
  } else {
    if (op == MPI_MAX) vmpi_reduce<double>(c, (const double*)in, (double*)out, n, [](double a, double b){ return a>b?a:b; });
    else if (op == MPI_MIN) vmpi_reduce<double>(c, (const double*)in, (double*)out, n, [](double a, double b){ return a<b?a:b; });
    else vmpi_reduce<double>(c, (const double*)in, (double*)out, n, [](double a, double b){ return a+b; });
  }
  return MPI_SUCCESS;
}
inline int MPI_Barrier(MPI_Comm c) { int x = 0, y; return MPI_Allreduce(&x, &y, 1, MPI_INT, MPI_SUM, c); }
#endif
