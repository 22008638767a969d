// comm.cu -- z-slab communication: one-plane (or two-plane) halo exchange and scalar
// all-reduces over NCCL (NVLink 5 / NVSwitch).  One process per GPU; the communicator is
// bootstrapped from a 128-byte unique id handed over by the launcher (torchrun, MPI, a file).
// NCCL is resolved at run time (dlopen) so that the single-GPU library has no link-time
// dependency on it; whichever libnccl.so.2 the process already has loaded (e.g. torch's) wins.
#include <dlfcn.h>
#include <time.h>
#include <cstdlib>
#include <cstring>

#include "adp_internal.cuh"

namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt64 = 4, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2, ncclMin = 3 };

struct Nccl {
    void *h = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
} g_nccl;

bool load_nccl(std::string &err)
{
    if (g_nccl.ok) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
    for (const char *n : names) {
        g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.h) break;
    }
    if (!g_nccl.h) { err = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return false; }
#define SYM(field, name)                                                    \
    *(void **)(&g_nccl.field) = dlsym(g_nccl.h, name);                      \
    if (!g_nccl.field) { err = std::string("NCCL symbol missing: ") + name; return false; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.ok = true;
    return true;
}
}  // namespace

struct adp_comm {
    ncclComm_t comm = nullptr;
};

#define NCCL_TRY(c, call)                                                                       \
    do {                                                                                        \
        int r__ = (call);                                                                       \
        if (r__ != ncclSuccess) {                                                               \
            (c)->err = std::string(#call) + ": " + g_nccl.GetErrorString(r__);                  \
            return ADP_ERR_NCCL;                                                                \
        }                                                                                       \
    } while (0)

extern "C" int adp_comm_unique_id(void *uid128)
{
    std::string err;
    if (!uid128 || !load_nccl(err)) return ADP_ERR_NCCL;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return ADP_ERR_NCCL;
    memcpy(uid128, &id, sizeof(id));
    return ADP_OK;
}

extern "C" int adp_comm_init(adp_ctx *c, int nranks, int rank, const void *uid128)
{
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, !c->geometry_set, "adp_comm_init must precede adp_set_geometry");
    ADP_REQUIRE(c, nranks >= 1 && rank >= 0 && rank < nranks, "adp_comm_init: bad rank / nranks");
    if (nranks == 1) { c->nranks = 1; c->rank = 0; return ADP_OK; }
    ADP_REQUIRE(c, uid128 != nullptr, "adp_comm_init: unique id missing");
    if (!load_nccl(c->err)) return ADP_ERR_NCCL;
    CUDA_TRY(c, cudaSetDevice(c->device));
    ncclUniqueId id;
    memcpy(&id, uid128, sizeof(id));
    c->comm = new adp_comm();
    NCCL_TRY(c, g_nccl.CommInitRank(&c->comm->comm, nranks, id, rank));
    c->nranks = nranks; c->rank = rank;
    return ADP_OK;
}

// Launcher-free bootstrap for the Fortran driver started N times (mpirun / a shell loop / torchrun):
// ADP_NRANKS (or WORLD_SIZE), ADP_RANK (or RANK), ADP_UID_FILE (default /tmp/adpres_b200.uid).
// Rank 0 writes the NCCL unique id to the file (atomically, via rename), the others wait for it.
extern "C" int adp_comm_init_env(adp_ctx *c)
{
    if (!c) return ADP_ERR_USAGE;
    auto env_int = [](const char *a, const char *b, int dflt) {
        const char *v = getenv(a);
        if (!v) v = getenv(b);
        return v ? atoi(v) : dflt;
    };
    const int nranks = env_int("ADP_NRANKS", "WORLD_SIZE", 1), rank = env_int("ADP_RANK", "RANK", 0);
    if (nranks <= 1) return adp_comm_init(c, 1, 0, nullptr);
    const char *path = getenv("ADP_UID_FILE");
    std::string file = path ? path : "/tmp/adpres_b200.uid";
    char uid[128];
    if (rank == 0) {
        int rc = adp_comm_unique_id(uid);
        if (rc) { c->err = "adp_comm_init_env: cannot create the NCCL unique id"; return rc; }
        std::string tmp = file + ".tmp";
        FILE *f = fopen(tmp.c_str(), "wb");
        if (!f || fwrite(uid, 1, 128, f) != 128) { c->err = "adp_comm_init_env: cannot write " + tmp; if (f) fclose(f); return ADP_ERR_USAGE; }
        fclose(f);
        if (rename(tmp.c_str(), file.c_str()) != 0) { c->err = "adp_comm_init_env: rename failed"; return ADP_ERR_USAGE; }
    } else {
        bool ok = false;
        for (int tries = 0; tries < 6000 && !ok; ++tries) {
            FILE *f = fopen(file.c_str(), "rb");
            if (f) { ok = fread(uid, 1, 128, f) == 128; fclose(f); }
            if (!ok) { struct timespec ts = {0, 10000000}; nanosleep(&ts, nullptr); }
        }
        if (!ok) { c->err = "adp_comm_init_env: timed out waiting for " + file; return ADP_ERR_NCCL; }
    }
    int rc = adp_comm_init(c, nranks, rank, uid);
    if (rank == 0 && rc == ADP_OK) remove(file.c_str());   // every rank has joined once CommInitRank returns
    return rc;
}

void adp_comm_destroy(adp_ctx *c)
{
    if (c->comm) {
        if (c->comm->comm && g_nccl.ok) g_nccl.CommDestroy(c->comm->comm);
        delete c->comm;
        c->comm = nullptr;
    }
}

// ghost planes of one node vector: send the top / bottom `n` owned planes to the upper / lower
// neighbour, receive theirs into the ghost planes.  n*np doubles per direction (193 KB at the
// 1 cm IAEA-3D mesh): latency, not bandwidth, is what this costs.
int adp_comm_halo(adp_ctx *c, double *v, int n)
{
    if (c->nranks == 1) return ADP_OK;
    const size_t np = c->np, cnt = n * np;
    double *own_lo = v + (size_t)ADP_GH * np, *own_hi = v + (size_t)(ADP_GH + c->nzl - n) * np;
    double *gh_lo = v + (size_t)(ADP_GH - n) * np, *gh_hi = v + (size_t)(ADP_GH + c->nzl) * np;
    ncclComm_t comm = c->comm->comm;
    NCCL_TRY(c, g_nccl.GroupStart());
    if (c->rank + 1 < c->nranks) {
        NCCL_TRY(c, g_nccl.Send(own_hi, cnt, ncclFloat64, c->rank + 1, comm, c->stream));
        NCCL_TRY(c, g_nccl.Recv(gh_hi, cnt, ncclFloat64, c->rank + 1, comm, c->stream));
    }
    if (c->rank > 0) {
        NCCL_TRY(c, g_nccl.Send(own_lo, cnt, ncclFloat64, c->rank - 1, comm, c->stream));
        NCCL_TRY(c, g_nccl.Recv(gh_lo, cnt, ncclFloat64, c->rank - 1, comm, c->stream));
    }
    NCCL_TRY(c, g_nccl.GroupEnd());
    return ADP_OK;
}

int adp_comm_allreduce_sum(adp_ctx *c, double *d, int count)
{
    if (c->nranks == 1) return ADP_OK;
    NCCL_TRY(c, g_nccl.AllReduce(d, d, count, ncclFloat64, ncclSum, c->comm->comm, c->stream));
    return ADP_OK;
}
int adp_comm_allreduce_max(adp_ctx *c, double *d, int count)
{
    if (c->nranks == 1) return ADP_OK;
    NCCL_TRY(c, g_nccl.AllReduce(d, d, count, ncclFloat64, ncclMax, c->comm->comm, c->stream));
    return ADP_OK;
}
int adp_comm_allreduce_min_ll(adp_ctx *c, long long *d, int count)
{
    if (c->nranks == 1) return ADP_OK;
    NCCL_TRY(c, g_nccl.AllReduce(d, d, count, ncclInt64, ncclMin, c->comm->comm, c->stream));
    return ADP_OK;
}
