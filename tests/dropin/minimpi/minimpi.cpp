// tests/dropin/minimpi/minimpi.cpp -- TEST INFRASTRUCTURE (see mpi.h in this directory).
//
// A single-box MPI subset over one POSIX shared-memory file.  Ranks are ordinary processes
// started by minimpirun.py with MINIMPI_SHM / MINIMPI_RANK / MINIMPI_SIZE in the environment
// (without them MPI_Init makes a one-rank world).  All shared state is valid when zero-filled,
// so no rank has to initialise it first:
//
//   [ header | inbox control blocks | one message arena per rank | window heap ]
//
//   * point-to-point: the sender copies the message into the RECEIVER's arena (a byte ring of
//     64-byte-aligned records under a spin lock) -- every send is eager/buffered, so Isend
//     completes at once.  Receives scan the ring oldest-first for (context, source, tag), which
//     gives MPI's non-overtaking order; consumed records are reclaimed from the head.
//   * collectives on any communicator: built from point-to-point on a private context
//     (MPI_COMM_WORLD's barrier is a sense-reversing counter in the header).
//   * RMA: MPI_Win_allocate carves the window from the shared heap, so MPI_Put is a memcpy and
//     MPI_Fetch_and_op a hardware atomic; lock/unlock are full fences.
#include "mpi.h"

#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

#include <map>
#include <vector>

namespace {

constexpr size_t kRec = 64;  // record granularity and header size
constexpr int kCollCtx = 0x40000000;

struct Header {
  int bar_count;
  int bar_sense;
  int abort_flag;
  int pad0;
  uint64_t heap_top;
};

struct Inbox {
  int lock;
  int pad0;
  uint64_t head;  // monotonic byte counters, guarded by lock
  uint64_t tail;
  char pad[kRec - 24];
};
static_assert(sizeof(Inbox) == kRec, "inbox control block is one record");

struct MsgHdr {
  uint64_t total;  // bytes this record occupies in the ring, header included
  uint64_t bytes;  // payload
  int src;         // sender's rank in the communicator
  int tag;
  int ctx;
  int consumed;
  char pad[kRec - 32];
};
static_assert(sizeof(MsgHdr) == kRec, "message header is one record");

struct CommInfo {
  int ctx;
  std::vector<int> members;  // world ranks, indexed by rank in this communicator
  int me;
  int coll_seq;
  int win_seq;
};

struct WinMember {
  int64_t off, size, disp_unit;
};
struct WinInfo {
  int comm;
  std::vector<WinMember> m;
  bool live;
};

struct DerivedType {
  size_t extent, align;
};

int g_rank = 0, g_size = 1, g_local_sense = 0;
bool g_init = false;
char *g_base = nullptr;
size_t g_map_bytes = 0, g_arena_bytes = 0, g_heap_bytes = 0;
Header *g_hdr = nullptr;
Inbox *g_inbox = nullptr;
char *g_arenas = nullptr, *g_heap = nullptr;
double g_timeout_s = 600.0;
std::vector<CommInfo> g_comms;
std::vector<std::vector<int>> g_groups;
std::vector<WinInfo> g_wins;
std::vector<DerivedType> g_types;
std::map<std::vector<int>, int> g_ctx_seq;

double now_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

[[noreturn]] void die(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "minimpi[%d/%d]: ", g_rank, g_size);
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
  fflush(stderr);
  if (g_hdr) __atomic_store_n(&g_hdr->abort_flag, 1, __ATOMIC_SEQ_CST);
  _exit(87);
}

struct Waiter {  // one blocking wait: yields, notices a peer's abort, gives up after the timeout
  double t0 = -1.0;
  unsigned spins = 0;
  void relax(const char *what) {
    if (__atomic_load_n(&g_hdr->abort_flag, __ATOMIC_RELAXED)) _exit(86);
    if (++spins < 64) return;
    sched_yield();
    if ((spins & 1023) == 0) {
      double t = now_s();
      if (t0 < 0) t0 = t;
      if (t - t0 > g_timeout_s) die("timed out after %.0f s in %s", g_timeout_s, what);
      if (spins > (1u << 16)) usleep(50);
    }
  }
};

size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

size_t basic_size(MPI_Datatype t) {
  switch (t) {
    case MPI_CHAR: case MPI_BYTE: return 1;
    case MPI_INT: case MPI_FLOAT: case MPI_UNSIGNED: return 4;
    case MPI_DOUBLE: case MPI_UNSIGNED_LONG_LONG: case MPI_UINT64_T: case MPI_UNSIGNED_LONG:
    case MPI_LONG_LONG: return 8;
    default: return 0;
  }
}
size_t type_extent(MPI_Datatype t) {
  if (t >= 64 && (size_t)(t - 64) < g_types.size()) return g_types[t - 64].extent;
  size_t s = basic_size(t);
  if (!s) die("unknown datatype %d", t);
  return s;
}
size_t type_align(MPI_Datatype t) {
  if (t >= 64 && (size_t)(t - 64) < g_types.size()) return g_types[t - 64].align;
  return basic_size(t);
}

CommInfo &comm_of(MPI_Comm c) {
  if (!g_init) die("MPI call before MPI_Init");
  if (c < 0 || (size_t)c >= g_comms.size()) die("invalid communicator %d", c);
  return g_comms[c];
}

void lock(Inbox &b) {
  Waiter w;
  int expect = 0;
  while (!__atomic_compare_exchange_n(&b.lock, &expect, 1, false, __ATOMIC_ACQUIRE,
                                      __ATOMIC_RELAXED)) {
    expect = 0;
    w.relax("inbox lock");
  }
}
void unlock(Inbox &b) { __atomic_store_n(&b.lock, 0, __ATOMIC_RELEASE); }

// copy one message into world rank `dest`'s arena
void post(int dest, int ctx, int src_in_comm, int tag, const void *buf, size_t bytes) {
  if (dest < 0 || dest >= g_size) die("send to invalid rank %d", dest);
  const size_t need = kRec + round_up(bytes, kRec);
  if (need + kRec > g_arena_bytes)
    die("message of %zu bytes exceeds the arena (%zu); raise MINIMPI_ARENA_MB", bytes,
        g_arena_bytes);
  Inbox &b = g_inbox[dest];
  char *arena = g_arenas + (size_t)dest * g_arena_bytes;
  Waiter w;
  for (;;) {
    lock(b);
    size_t pos = b.tail % g_arena_bytes;
    size_t pad = (pos + need > g_arena_bytes) ? g_arena_bytes - pos : 0;
    size_t used = b.tail - b.head;
    if (used + pad + need <= g_arena_bytes) {
      if (pad) {  // a consumed filler record up to the end of the ring
        MsgHdr *f = (MsgHdr *)(arena + pos);
        f->total = pad; f->bytes = 0; f->src = -1; f->tag = 0; f->ctx = -1; f->consumed = 1;
        b.tail += pad;
        pos = 0;
      }
      MsgHdr *h = (MsgHdr *)(arena + pos);
      h->total = need; h->bytes = bytes; h->src = src_in_comm; h->tag = tag; h->ctx = ctx;
      h->consumed = 0;
      if (bytes) memcpy(arena + pos + kRec, buf, bytes);
      b.tail += need;
      unlock(b);
      return;
    }
    unlock(b);
    w.relax("send (receiver's arena full)");
  }
}

// find the oldest matching message in my arena; with `take` copy it out and consume it
bool match(int ctx, int src, int tag, bool take, void *buf, size_t cap, MPI_Status *st) {
  Inbox &b = g_inbox[g_rank];
  char *arena = g_arenas + (size_t)g_rank * g_arena_bytes;
  bool found = false;
  lock(b);
  for (uint64_t p = b.head; p < b.tail;) {
    MsgHdr *h = (MsgHdr *)(arena + p % g_arena_bytes);
    if (!h->consumed && h->ctx == ctx && (src == MPI_ANY_SOURCE || src == h->src) &&
        (tag == MPI_ANY_TAG || tag == h->tag)) {
      if (st) {
        st->MPI_SOURCE = h->src; st->MPI_TAG = h->tag; st->MPI_ERROR = MPI_SUCCESS;
        st->_bytes = (long)h->bytes;
      }
      if (take) {
        if (h->bytes > cap) {
          unlock(b);
          die("message truncated: %zu bytes into a %zu-byte receive (src %d tag %d)",
              (size_t)h->bytes, cap, h->src, h->tag);
        }
        if (h->bytes) memcpy(buf, (char *)h + kRec, h->bytes);
        h->consumed = 1;
        while (b.head < b.tail) {
          MsgHdr *o = (MsgHdr *)(arena + b.head % g_arena_bytes);
          if (!o->consumed) break;
          b.head += o->total;
        }
        // an empty ring restarts at offset 0, so a quiet exchange keeps touching the same few
        // pages of the (sparse) shared-memory file instead of walking through all of it
        if (b.head == b.tail) b.head = b.tail = 0;
      }
      found = true;
      break;
    }
    p += h->total;
  }
  unlock(b);
  return found;
}

void recv_blocking(int ctx, int src, int tag, void *buf, size_t cap, MPI_Status *st) {
  Waiter w;
  while (!match(ctx, src, tag, true, buf, cap, st)) w.relax("receive");
}

// collectives over point-to-point, private context, one sequence number per call
struct Coll {
  CommInfo &c;
  int ctx, tag;
  explicit Coll(MPI_Comm comm) : c(comm_of(comm)), ctx(c.ctx | kCollCtx), tag(c.coll_seq++) {}
  int n() const { return (int)c.members.size(); }
  void send(int to, const void *buf, size_t bytes) { post(c.members[to], ctx, c.me, tag, buf, bytes); }
  void recv(int from, void *buf, size_t bytes) { recv_blocking(ctx, from, tag, buf, bytes, nullptr); }
};

template <typename T> void reduce_into(T *acc, const T *in, int n, MPI_Op op) {
  for (int i = 0; i < n; ++i) {
    if (op == MPI_SUM) acc[i] += in[i];
    else if (op == MPI_MAX) acc[i] = in[i] > acc[i] ? in[i] : acc[i];
    else if (op == MPI_MIN) acc[i] = in[i] < acc[i] ? in[i] : acc[i];
    else die("unsupported reduction op %d", op);
  }
}

void world_barrier() {
  if (g_size == 1) return;
  g_local_sense ^= 1;
  if (__atomic_add_fetch(&g_hdr->bar_count, 1, __ATOMIC_ACQ_REL) == g_size) {
    __atomic_store_n(&g_hdr->bar_count, 0, __ATOMIC_RELAXED);
    __atomic_store_n(&g_hdr->bar_sense, g_local_sense, __ATOMIC_RELEASE);
  } else {
    Waiter w;
    while (__atomic_load_n(&g_hdr->bar_sense, __ATOMIC_ACQUIRE) != g_local_sense) w.relax("barrier");
  }
}

char *win_target(MPI_Win win, int target_rank, MPI_Aint disp, size_t bytes) {
  if (win < 0 || (size_t)win >= g_wins.size() || !g_wins[win].live) die("invalid window %d", win);
  WinInfo &w = g_wins[win];
  if (target_rank < 0 || (size_t)target_rank >= w.m.size()) die("window target rank %d out of range", target_rank);
  const WinMember &t = w.m[target_rank];
  size_t start = (size_t)disp * (size_t)t.disp_unit;
  if (start + bytes > (size_t)t.size) die("window access [%zu, %zu) beyond the target's %ld bytes", start, start + bytes, (long)t.size);
  return g_heap + t.off + start;
}
template <typename T> void fetch_op(T *target, const void *origin, void *result, MPI_Op op) {
  T old;
  T v = origin ? *(const T *)origin : 0;
  switch (op) {
    case MPI_NO_OP: old = __atomic_load_n(target, __ATOMIC_SEQ_CST); break;
    case MPI_SUM: old = __atomic_fetch_add(target, v, __ATOMIC_SEQ_CST); break;
    case MPI_BOR: old = __atomic_fetch_or(target, v, __ATOMIC_SEQ_CST); break;
    case MPI_BAND: old = __atomic_fetch_and(target, v, __ATOMIC_SEQ_CST); break;
    case MPI_REPLACE: old = __atomic_exchange_n(target, v, __ATOMIC_SEQ_CST); break;
    default: die("unsupported fetch-and-op %d", op);
  }
  *(T *)result = old;
}
}  // namespace

extern "C" {

int MPI_Init(int *, char ***) {
  if (g_init) die("MPI_Init called twice");
  const char *shm = getenv("MINIMPI_SHM");
  if (shm) {
    const char *r = getenv("MINIMPI_RANK"), *s = getenv("MINIMPI_SIZE");
    if (!r || !s) die("MINIMPI_SHM set without MINIMPI_RANK / MINIMPI_SIZE");
    g_rank = atoi(r);
    g_size = atoi(s);
    if (g_size < 1 || g_rank < 0 || g_rank >= g_size) die("bad rank/size %s/%s", r, s);
  }
  const char *e;
  size_t arena_mb = (e = getenv("MINIMPI_ARENA_MB")) ? (size_t)atol(e) : 64;
  size_t heap_mb = (e = getenv("MINIMPI_HEAP_MB")) ? (size_t)atol(e) : 256;
  if ((e = getenv("MINIMPI_TIMEOUT_S"))) g_timeout_s = atof(e);
  g_arena_bytes = arena_mb << 20;
  g_heap_bytes = heap_mb << 20;
  const size_t ctl = round_up(sizeof(Inbox) * (size_t)g_size, 4096);
  g_map_bytes = 4096 + ctl + g_arena_bytes * (size_t)g_size + g_heap_bytes;
  void *p;
  if (shm) {
    int fd = open(shm, O_RDWR | O_CREAT, 0600);
    if (fd < 0) die("open %s: %s", shm, strerror(errno));
    if (ftruncate(fd, (off_t)g_map_bytes) != 0) die("ftruncate %s: %s", shm, strerror(errno));
    p = mmap(nullptr, g_map_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
  } else {
    p = mmap(nullptr, g_map_bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  }
  if (p == MAP_FAILED) die("mmap of %zu bytes: %s", g_map_bytes, strerror(errno));
  g_base = (char *)p;
  g_hdr = (Header *)g_base;
  g_inbox = (Inbox *)(g_base + 4096);
  g_arenas = g_base + 4096 + ctl;
  g_heap = g_arenas + g_arena_bytes * (size_t)g_size;
  CommInfo world;
  world.ctx = 0;
  world.me = g_rank;
  world.coll_seq = world.win_seq = 0;
  for (int i = 0; i < g_size; ++i) world.members.push_back(i);
  g_comms.push_back(world);
  g_init = true;
  world_barrier();
  return MPI_SUCCESS;
}

int MPI_Finalize(void) {
  world_barrier();
  munmap(g_base, g_map_bytes);
  g_init = false;
  return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm, int code) {
  fflush(stdout);
  fflush(stderr);
  if (g_hdr) __atomic_store_n(&g_hdr->abort_flag, 1, __ATOMIC_SEQ_CST);
  _exit(code ? code : 1);
}

double MPI_Wtime(void) { return now_s(); }
double MPI_Wtick(void) { return 1e-9; }

int MPI_Comm_rank(MPI_Comm comm, int *rank) { *rank = comm_of(comm).me; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int *size) {
  *size = (int)comm_of(comm).members.size();
  return MPI_SUCCESS;
}
int MPI_Get_processor_name(char *name, int *len) {
  if (gethostname(name, MPI_MAX_PROCESSOR_NAME) != 0) strcpy(name, "localhost");
  name[MPI_MAX_PROCESSOR_NAME - 1] = 0;
  *len = (int)strlen(name);
  return MPI_SUCCESS;
}

int MPI_Type_create_struct(int n, const int *blocklengths, const MPI_Aint *offsets,
                           const MPI_Datatype *types, MPI_Datatype *newtype) {
  size_t ub = 0, align = 1;
  for (int i = 0; i < n; ++i) {
    if (offsets[i] < 0) die("negative struct offset");
    size_t end = (size_t)offsets[i] + (size_t)blocklengths[i] * type_extent(types[i]);
    if (end > ub) ub = end;
    if (type_align(types[i]) > align) align = type_align(types[i]);
  }
  g_types.push_back({round_up(ub, align), align});
  *newtype = 64 + (int)g_types.size() - 1;
  return MPI_SUCCESS;
}
int MPI_Type_commit(MPI_Datatype *) { return MPI_SUCCESS; }
int MPI_Type_free(MPI_Datatype *t) { *t = MPI_DATATYPE_NULL; return MPI_SUCCESS; }
int MPI_Type_size(MPI_Datatype t, int *size) { *size = (int)type_extent(t); return MPI_SUCCESS; }
int MPI_Get_count(const MPI_Status *status, MPI_Datatype t, int *count) {
  *count = (int)((size_t)status->_bytes / type_extent(t));
  return MPI_SUCCESS;
}

int MPI_Send(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm comm) {
  CommInfo &c = comm_of(comm);
  if (dest == MPI_PROC_NULL) return MPI_SUCCESS;
  if (dest < 0 || dest >= (int)c.members.size()) die("send to rank %d of a %zu-rank communicator", dest, c.members.size());
  post(c.members[dest], c.ctx, c.me, tag, buf, (size_t)count * type_extent(t));
  return MPI_SUCCESS;
}
int MPI_Isend(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm comm,
              MPI_Request *req) {
  MPI_Send(buf, count, t, dest, tag, comm);  // eager: the payload is already in the receiver's arena
  *req = 1;
  return MPI_SUCCESS;
}
int MPI_Test(MPI_Request *req, int *flag, MPI_Status *) {
  *flag = 1;
  *req = MPI_REQUEST_NULL;
  return MPI_SUCCESS;
}
int MPI_Wait(MPI_Request *req, MPI_Status *) { *req = MPI_REQUEST_NULL; return MPI_SUCCESS; }

int MPI_Recv(void *buf, int count, MPI_Datatype t, int source, int tag, MPI_Comm comm,
             MPI_Status *status) {
  CommInfo &c = comm_of(comm);
  if (source == MPI_PROC_NULL) {
    if (status) { status->MPI_SOURCE = MPI_PROC_NULL; status->MPI_TAG = MPI_ANY_TAG; status->_bytes = 0; }
    return MPI_SUCCESS;
  }
  recv_blocking(c.ctx, source, tag, buf, (size_t)count * type_extent(t), status);
  return MPI_SUCCESS;
}
int MPI_Sendrecv(const void *sbuf, int scount, MPI_Datatype st, int dest, int stag, void *rbuf,
                 int rcount, MPI_Datatype rt, int source, int rtag, MPI_Comm comm,
                 MPI_Status *status) {
  MPI_Send(sbuf, scount, st, dest, stag, comm);
  return MPI_Recv(rbuf, rcount, rt, source, rtag, comm, status);
}
int MPI_Iprobe(int source, int tag, MPI_Comm comm, int *flag, MPI_Status *status) {
  CommInfo &c = comm_of(comm);
  if (__atomic_load_n(&g_hdr->abort_flag, __ATOMIC_RELAXED)) _exit(86);
  *flag = match(c.ctx, source, tag, false, nullptr, 0, status) ? 1 : 0;
  return MPI_SUCCESS;
}
int MPI_Probe(int source, int tag, MPI_Comm comm, MPI_Status *status) {
  CommInfo &c = comm_of(comm);
  Waiter w;
  while (!match(c.ctx, source, tag, false, nullptr, 0, status)) w.relax("probe");
  return MPI_SUCCESS;
}

int MPI_Barrier(MPI_Comm comm) {
  if (comm == MPI_COMM_WORLD) {
    comm_of(comm);
    world_barrier();
    return MPI_SUCCESS;
  }
  Coll k(comm);
  char z = 0;
  if (k.c.me == 0) {
    for (int r = 1; r < k.n(); ++r) k.recv(r, &z, 1);
    for (int r = 1; r < k.n(); ++r) k.send(r, &z, 1);
  } else {
    k.send(0, &z, 1);
    k.recv(0, &z, 1);
  }
  return MPI_SUCCESS;
}

int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm comm) {
  Coll k(comm);
  size_t bytes = (size_t)count * type_extent(t);
  if (k.c.me == root) {
    for (int r = 0; r < k.n(); ++r)
      if (r != root) k.send(r, buf, bytes);
  } else {
    k.recv(root, buf, bytes);
  }
  return MPI_SUCCESS;
}

int MPI_Allreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype t, MPI_Op op,
                  MPI_Comm comm) {
  Coll k(comm);
  size_t bytes = (size_t)count * type_extent(t);
  if (k.c.me == 0) {
    memcpy(rbuf, sbuf, bytes);
    std::vector<char> tmp(bytes ? bytes : 1);
    for (int r = 1; r < k.n(); ++r) {
      k.recv(r, tmp.data(), bytes);
      switch (t) {
        case MPI_INT: reduce_into((int *)rbuf, (const int *)tmp.data(), count, op); break;
        case MPI_UNSIGNED: reduce_into((unsigned *)rbuf, (const unsigned *)tmp.data(), count, op); break;
        case MPI_FLOAT: reduce_into((float *)rbuf, (const float *)tmp.data(), count, op); break;
        case MPI_DOUBLE: reduce_into((double *)rbuf, (const double *)tmp.data(), count, op); break;
        case MPI_LONG_LONG: reduce_into((long long *)rbuf, (const long long *)tmp.data(), count, op); break;
        case MPI_UNSIGNED_LONG_LONG: case MPI_UINT64_T: case MPI_UNSIGNED_LONG:
          reduce_into((uint64_t *)rbuf, (const uint64_t *)tmp.data(), count, op); break;
        default: die("allreduce on unsupported datatype %d", t);
      }
    }
    for (int r = 1; r < k.n(); ++r) k.send(r, rbuf, bytes);
  } else {
    k.send(0, sbuf, bytes);
    k.recv(0, rbuf, bytes);
  }
  return MPI_SUCCESS;
}

int MPI_Gatherv(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, const int *rcounts,
                const int *displs, MPI_Datatype rt, int root, MPI_Comm comm) {
  Coll k(comm);
  if (k.c.me != root) {
    k.send(root, sbuf, (size_t)scount * type_extent(st));
    return MPI_SUCCESS;
  }
  const size_t e = type_extent(rt);
  for (int r = 0; r < k.n(); ++r) {
    char *dst = (char *)rbuf + (size_t)displs[r] * e;
    if (r == root) {
      if ((size_t)scount * type_extent(st) > (size_t)rcounts[r] * e) die("gatherv: own block too large");
      memcpy(dst, sbuf, (size_t)scount * type_extent(st));
    } else {
      k.recv(r, dst, (size_t)rcounts[r] * e);
    }
  }
  return MPI_SUCCESS;
}

int MPI_Gather(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, int rcount,
               MPI_Datatype rt, int root, MPI_Comm comm) {
  CommInfo &c = comm_of(comm);
  std::vector<int> counts, displs;
  if (c.me == root)
    for (size_t r = 0; r < c.members.size(); ++r) {
      counts.push_back(rcount);
      displs.push_back((int)r * rcount);
    }
  return MPI_Gatherv(sbuf, scount, st, rbuf, counts.data(), displs.data(), rt, root, comm);
}

int MPI_Allgather(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, int rcount,
                  MPI_Datatype rt, MPI_Comm comm) {
  MPI_Gather(sbuf, scount, st, rbuf, rcount, rt, 0, comm);
  int n = (int)comm_of(comm).members.size();
  return MPI_Bcast(rbuf, rcount * n, rt, 0, comm);
}

int MPI_Comm_group(MPI_Comm comm, MPI_Group *group) {
  g_groups.push_back(comm_of(comm).members);
  *group = (int)g_groups.size() - 1;
  return MPI_SUCCESS;
}
int MPI_Group_incl(MPI_Group group, int n, const int *ranks, MPI_Group *newgroup) {
  if (group < 0 || (size_t)group >= g_groups.size()) die("invalid group %d", group);
  std::vector<int> sel;
  for (int i = 0; i < n; ++i) {
    if (ranks[i] < 0 || (size_t)ranks[i] >= g_groups[group].size()) die("group_incl: rank %d out of range", ranks[i]);
    sel.push_back(g_groups[group][ranks[i]]);
  }
  g_groups.push_back(sel);
  *newgroup = (int)g_groups.size() - 1;
  return MPI_SUCCESS;
}
int MPI_Group_free(MPI_Group *group) { *group = MPI_GROUP_NULL; return MPI_SUCCESS; }

int MPI_Comm_create_group(MPI_Comm comm, MPI_Group group, int tag, MPI_Comm *newcomm) {
  comm_of(comm);
  if (group < 0 || (size_t)group >= g_groups.size()) die("invalid group %d", group);
  const std::vector<int> &mem = g_groups[group];
  int me = -1;
  for (size_t i = 0; i < mem.size(); ++i)
    if (mem[i] == g_rank) me = (int)i;
  if (me < 0) { *newcomm = MPI_COMM_NULL; return MPI_SUCCESS; }
  // every member derives the same context id locally: hash of the ordered member list, the tag
  // and how many communicators this list has produced so far (creation is collective, so the
  // count agrees on all members)
  std::vector<int> key(mem);
  key.push_back(tag);
  int seq = g_ctx_seq[key]++;
  uint32_t h = 2166136261u;
  for (int v : key) { h ^= (uint32_t)v; h *= 16777619u; }
  h ^= (uint32_t)seq; h *= 16777619u;
  CommInfo c;
  c.ctx = (int)((h & 0x3fffff00u) | 0x80u) ;
  c.members = mem;
  c.me = me;
  c.coll_seq = c.win_seq = 0;
  g_comms.push_back(c);
  *newcomm = (int)g_comms.size() - 1;
  return MPI_SUCCESS;
}
int MPI_Comm_free(MPI_Comm *comm) { *comm = MPI_COMM_NULL; return MPI_SUCCESS; }

int MPI_Win_allocate(MPI_Aint size, int disp_unit, MPI_Info, MPI_Comm comm, void *baseptr,
                     MPI_Win *win) {
  CommInfo &c = comm_of(comm);
  WinMember mine = {0, (int64_t)size, (int64_t)disp_unit};
  if (size > 0) {
    size_t need = round_up((size_t)size, kRec);
    uint64_t off = __atomic_fetch_add(&g_hdr->heap_top, (uint64_t)need, __ATOMIC_ACQ_REL);
    if (off + need > g_heap_bytes) die("window heap exhausted (%zu bytes); raise MINIMPI_HEAP_MB", g_heap_bytes);
    mine.off = (int64_t)off;
  }
  WinInfo w;
  w.comm = comm;
  w.live = true;
  w.m.resize(c.members.size());
  const int ctx = c.ctx | kCollCtx, tag = 0x57000000 + c.win_seq++;
  for (size_t r = 0; r < c.members.size(); ++r)
    if ((int)r != c.me) post(c.members[r], ctx, c.me, tag, &mine, sizeof mine);
  for (size_t r = 0; r < c.members.size(); ++r) {
    if ((int)r == c.me) w.m[r] = mine;
    else recv_blocking(ctx, (int)r, tag, &w.m[r], sizeof(WinMember), nullptr);
  }
  *(void **)baseptr = size > 0 ? (void *)(g_heap + mine.off) : nullptr;
  g_wins.push_back(w);
  *win = (int)g_wins.size() - 1;
  return MPI_SUCCESS;
}
int MPI_Win_free(MPI_Win *win) {
  if (*win >= 0 && (size_t)*win < g_wins.size()) g_wins[*win].live = false;
  *win = -1;
  return MPI_SUCCESS;
}
int MPI_Win_lock(int, int, int, MPI_Win) { __atomic_thread_fence(__ATOMIC_SEQ_CST); return MPI_SUCCESS; }
int MPI_Win_unlock(int, MPI_Win) { __atomic_thread_fence(__ATOMIC_SEQ_CST); return MPI_SUCCESS; }

int MPI_Put(const void *origin, int ocount, MPI_Datatype ot, int target_rank, MPI_Aint target_disp,
            int, MPI_Datatype, MPI_Win win) {
  size_t bytes = (size_t)ocount * type_extent(ot);
  memcpy(win_target(win, target_rank, target_disp, bytes), origin, bytes);
  return MPI_SUCCESS;
}
int MPI_Get(void *origin, int ocount, MPI_Datatype ot, int target_rank, MPI_Aint target_disp, int,
            MPI_Datatype, MPI_Win win) {
  size_t bytes = (size_t)ocount * type_extent(ot);
  memcpy(origin, win_target(win, target_rank, target_disp, bytes), bytes);
  return MPI_SUCCESS;
}

int MPI_Fetch_and_op(const void *origin, void *result, MPI_Datatype t, int target_rank,
                     MPI_Aint target_disp, MPI_Op op, MPI_Win win) {
  size_t e = type_extent(t);
  char *p = win_target(win, target_rank, target_disp, e);
  if (e == 8) fetch_op((uint64_t *)p, origin, result, op);
  else if (e == 4) fetch_op((uint32_t *)p, origin, result, op);
  else die("fetch-and-op on a %zu-byte type", e);
  return MPI_SUCCESS;
}

}  // extern "C"
