// Multi-GPU exchange step of the path: particles are sharded by index slice, every
// rank deposits its slice, and the node charge density is summed with ONE NCCL all-reduce per
// step over NVLink/NVSwitch before the replicated field solve.  The reference has no
// communication layer at all (single process, single thread), so nothing is replaced here.
//
// NCCL is bound at run time with dlopen("libnccl.so.2") -- in a torch process that resolves to
// the copy torch already loaded -- so the library has no link-time NCCL dependency and
// single-GPU use never touches it.
#include <dlfcn.h>

#include "common.cuh"

namespace {

typedef struct { char internal[128]; } nccl_uid;       // ncclUniqueId, nccl.h:37-38
typedef int (*fn_get_uid)(nccl_uid *);
typedef int (*fn_init_rank)(void **, int, nccl_uid, int);
typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_destroy)(void *);
typedef const char *(*fn_errstr)(int);

struct Nccl {
  void *h = nullptr;
  fn_get_uid get_uid = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_destroy destroy = nullptr;
  fn_errstr errstr = nullptr;
} g_nccl;

int32_t load_nccl() {
  if (g_nccl.h) return ISKB_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  if (!g_nccl.h) return iskb_fail(ISKB_E_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
  g_nccl.get_uid = (fn_get_uid)dlsym(g_nccl.h, "ncclGetUniqueId");
  g_nccl.init_rank = (fn_init_rank)dlsym(g_nccl.h, "ncclCommInitRank");
  g_nccl.allreduce = (fn_allreduce)dlsym(g_nccl.h, "ncclAllReduce");
  g_nccl.destroy = (fn_destroy)dlsym(g_nccl.h, "ncclCommDestroy");
  g_nccl.errstr = (fn_errstr)dlsym(g_nccl.h, "ncclGetErrorString");
  if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.allreduce || !g_nccl.destroy)
    return iskb_fail(ISKB_E_NCCL, "libnccl is missing required symbols");
  return ISKB_OK;
}

int32_t nccl_fail(int rc, const char *what) {
  return iskb_fail(ISKB_E_NCCL, "%s failed: %s", what, g_nccl.errstr ? g_nccl.errstr(rc) : "?");
}

}  // namespace

extern "C" int32_t iskb_comm_unique_id(void *id128) {
  if (!id128) return iskb_fail(ISKB_E_INVALID, "null id");
  ISKB_TRY(load_nccl());
  const int rc = g_nccl.get_uid((nccl_uid *)id128);
  return rc ? nccl_fail(rc, "ncclGetUniqueId") : ISKB_OK;
}

extern "C" int32_t iskb_comm_init(iskb_ctx *c, int32_t n_ranks, int32_t rank, const void *id128) {
  if (!c || n_ranks < 1 || rank < 0 || rank >= n_ranks) return iskb_fail(ISKB_E_INVALID, "bad rank / n_ranks");
  c->n_ranks = n_ranks;
  c->rank = rank;
  if (n_ranks == 1) return ISKB_OK;
  if (!id128) return iskb_fail(ISKB_E_INVALID, "null id");
  ISKB_TRY(load_nccl());
  CU_TRY(cudaSetDevice(c->device));
  nccl_uid id;
  memcpy(&id, id128, sizeof(id));
  const int rc = g_nccl.init_rank(&c->nccl_comm, n_ranks, id, rank);
  return rc ? nccl_fail(rc, "ncclCommInitRank") : ISKB_OK;
}

int32_t comm_allreduce_sum(iskb_ctx *c, double *d_buf, int64_t n) {
  if (c->n_ranks == 1) return ISKB_OK;
  if (!c->nccl_comm) return iskb_fail(ISKB_E_NCCL, "iskb_comm_init was not called");
  const int rc = g_nccl.allreduce(d_buf, d_buf, (size_t)n, /*ncclDouble*/ 8, /*ncclSum*/ 0, c->nccl_comm, c->stream);
  return rc ? nccl_fail(rc, "ncclAllReduce") : ISKB_OK;
}

int32_t comm_allreduce_sum_i64(iskb_ctx *c, long long *d_buf, int64_t n) {
  if (c->n_ranks == 1) return ISKB_OK;
  if (!c->nccl_comm) return iskb_fail(ISKB_E_NCCL, "iskb_comm_init was not called");
  const int rc = g_nccl.allreduce(d_buf, d_buf, (size_t)n, /*ncclInt64*/ 4, /*ncclSum*/ 0, c->nccl_comm, c->stream);
  return rc ? nccl_fail(rc, "ncclAllReduce") : ISKB_OK;
}

int32_t comm_destroy(iskb_ctx *c) {
  if (c->nccl_comm && g_nccl.destroy) g_nccl.destroy(c->nccl_comm);
  c->nccl_comm = nullptr;
  return ISKB_OK;
}
