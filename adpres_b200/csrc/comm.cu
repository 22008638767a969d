// comm.cu -- z-slab communication: one-plane (or two-plane) halo exchange and scalar
// all-reduces over NCCL (NVLink 5 / NVSwitch).  One process per GPU; the communicator is
// bootstrapped from a 128-byte unique id handed over by the launcher (torchrun, MPI, a file).
// NCCL is resolved at run time (dlopen) so that the single-GPU library has no link-time
// dependency on it; whichever libnccl.so.2 the process already has loaded (e.g. torch's) wins.
#include <dlfcn.h>
#include <time.h>
#include <sys/stat.h>
#include <cstdlib>
#include <cstring>

#include "adp_internal.cuh"
#include "mail.cuh"

namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclInt64 = 4, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2, ncclMin = 3 };

struct Nccl {
    void *h = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
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
    SYM(AllReduce, "ncclAllReduce") SYM(AllGather, "ncclAllGather") SYM(Broadcast, "ncclBroadcast") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(GroupStart, "ncclGroupStart")
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
    // The id file must belong to THIS job: ADP_UID_FILE names it, otherwise the name is derived from the
    // launcher's job key (MASTER_PORT / ADP_JOB_ID) -- a fixed name in /tmp would pick up the stale file of a
    // crashed or concurrent job.  Rank 0 removes any old file before it creates the id; the file carries a
    // 16-byte header (magic + job key) that the readers check.
    const char *path = getenv("ADP_UID_FILE");
    const char *job = getenv("ADP_JOB_ID");
    if (!job) job = getenv("MASTER_PORT");
    if (!path && !job) {
        c->err = "adp_comm_init_env: set ADP_UID_FILE (a path private to this job) or ADP_JOB_ID / MASTER_PORT";
        return ADP_ERR_USAGE;
    }
    std::string file = path ? path : std::string("/tmp/adpres_b200.") + job + ".uid";
    char hdr[16] = "ADPUID1";
    strncpy(hdr + 8, job ? job : "", 7);
    char uid[128];
    if (rank == 0) {
        remove(file.c_str());
        int rc = adp_comm_unique_id(uid);
        if (rc) { c->err = "adp_comm_init_env: cannot create the NCCL unique id"; return rc; }
        std::string tmp = file + ".tmp";
        FILE *f = fopen(tmp.c_str(), "wb");
        if (!f || fwrite(hdr, 1, 16, f) != 16 || fwrite(uid, 1, 128, f) != 128) { c->err = "adp_comm_init_env: cannot write " + tmp; if (f) fclose(f); return ADP_ERR_USAGE; }
        fclose(f);
        if (rename(tmp.c_str(), file.c_str()) != 0) { c->err = "adp_comm_init_env: rename failed"; return ADP_ERR_USAGE; }
    } else {
        bool ok = false;
        const time_t started = time(nullptr);
        for (int tries = 0; tries < 6000 && !ok; ++tries) {
            struct stat sb;
            // a file much older than this process is a leftover of another run: rank 0 is about to replace it
            if (stat(file.c_str(), &sb) != 0 || sb.st_mtime + 60 < started) { struct timespec ts = {0, 10000000}; nanosleep(&ts, nullptr); continue; }
            FILE *f = fopen(file.c_str(), "rb");
            char h2[16];
            if (f) { ok = fread(h2, 1, 16, f) == 16 && memcmp(h2, hdr, 16) == 0 && fread(uid, 1, 128, f) == 128; fclose(f); }
            if (!ok) { struct timespec ts = {0, 10000000}; nanosleep(&ts, nullptr); }
        }
        if (!ok) { c->err = "adp_comm_init_env: timed out waiting for " + file; return ADP_ERR_NCCL; }
    }
    int rc = adp_comm_init(c, nranks, rank, uid);
    if (rank == 0 && rc == ADP_OK) remove(file.c_str());   // every rank has joined once CommInitRank returns
    return rc;
}

// Map the z-neighbours' BiCGSTAB vectors and flux buffers into this process (CUDA IPC over
// NVLink) so that kernels can store their boundary plane straight into the neighbour's ghost
// plane.  All ranks must agree, so the outcome is min-reduced; on any failure the NCCL
// send/recv halo path stays in use.
int adp_comm_map_peers(adp_ctx *c)
{
    c->peer_ok = false;
    if (c->nranks == 1) return ADP_OK;
    if (getenv("ADP_NO_PEER")) return ADP_OK;
    double *bufs[PB_COUNT] = {c->d_rs, c->d_v, c->d_v2, c->d_r, c->d_f0[0], c->d_f0[1]};
    const size_t HB = sizeof(cudaIpcMemHandle_t), per = PB_COUNT * HB;
    std::vector<char> mine(per), all(per * c->nranks);
    int ok = 1;
    for (int b = 0; b < PB_COUNT; ++b)
        if (cudaIpcGetMemHandle((cudaIpcMemHandle_t *)(mine.data() + b * HB), bufs[b]) != cudaSuccess) { ok = 0; cudaGetLastError(); }
    char *d_send = nullptr, *d_recv = nullptr;
    CUDA_TRY(c, cudaMalloc((void **)&d_send, per));
    CUDA_TRY(c, cudaMalloc((void **)&d_recv, per * c->nranks));
    CUDA_TRY(c, cudaMemcpyAsync(d_send, mine.data(), per, cudaMemcpyHostToDevice, c->stream));
    NCCL_TRY(c, g_nccl.AllGather(d_send, d_recv, per, ncclInt8, c->comm->comm, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(all.data(), d_recv, per * c->nranks, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    cudaFree(d_send); cudaFree(d_recv);
    auto open_rank = [&](int peer, double **dst) {
        for (int b = 0; b < PB_COUNT; ++b) {
            void *ptr = nullptr;
            cudaIpcMemHandle_t h;
            memcpy(&h, all.data() + (size_t)peer * per + b * HB, HB);
            if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); ptr = nullptr; }
            dst[b] = (double *)ptr;
        }
    };
    for (int b = 0; b < PB_COUNT; ++b) c->peer_lo[b] = c->peer_hi[b] = nullptr;
    if (ok && c->rank > 0) open_rank(c->rank - 1, c->peer_lo);
    if (ok && c->rank + 1 < c->nranks) open_rank(c->rank + 1, c->peer_hi);
    // neighbours' slab sizes (same partition formula as adp_set_geometry)
    auto planes = [&](int r) { const int base = c->nzz / c->nranks, rem = c->nzz % c->nranks; return base + (r < rem ? 1 : 0); };
    c->nzl_lo = c->rank > 0 ? planes(c->rank - 1) : 0;
    c->nzl_hi = c->rank + 1 < c->nranks ? planes(c->rank + 1) : 0;
    c->NV_lo = (long long)c->np * (c->nzl_lo + 2 * ADP_GH);
    c->NV_hi = (long long)c->np * (c->nzl_hi + 2 * ADP_GH);
    // agreement
    double flag = ok;
    CUDA_TRY(c, cudaMemcpyAsync(c->d_scal + S_TMP1, &flag, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    NCCL_TRY(c, g_nccl.AllReduce(c->d_scal + S_TMP1, c->d_scal + S_TMP1, 1, ncclFloat64, ncclMin, c->comm->comm, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(&flag, c->d_scal + S_TMP1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->peer_ok = flag > 0.5;
    if (!c->peer_ok) { adp_comm_unmap_peers(c); return ADP_OK; }
    // ---- mailboxes of the in-kernel all-reduce: every rank maps every other rank's
    c->peer_ar = false;
    if (c->nranks > ADP_MAX_RANKS || getenv("ADP_NO_PEER_AR")) return ADP_OK;
    const size_t mail_bytes = (size_t)ADP_MAIL_SLOTS * c->nranks * ADP_MAIL_WORDS * sizeof(double);
    if (!c->d_mail) {
        CUDA_TRY(c, cudaMalloc((void **)&c->d_mail, mail_bytes));
        CUDA_TRY(c, cudaMalloc((void **)&c->d_arseq, sizeof(unsigned long long)));
    }
    CUDA_TRY(c, cudaMemsetAsync(c->d_mail, 0, mail_bytes, c->stream));
    CUDA_TRY(c, cudaMemsetAsync(c->d_arseq, 0, sizeof(unsigned long long), c->stream));
    int ok2 = 1;
    std::vector<char> hm(HB), hall(HB * c->nranks);
    if (cudaIpcGetMemHandle((cudaIpcMemHandle_t *)hm.data(), c->d_mail) != cudaSuccess) { ok2 = 0; cudaGetLastError(); }
    CUDA_TRY(c, cudaMalloc((void **)&d_send, HB));
    CUDA_TRY(c, cudaMalloc((void **)&d_recv, HB * c->nranks));
    CUDA_TRY(c, cudaMemcpyAsync(d_send, hm.data(), HB, cudaMemcpyHostToDevice, c->stream));
    NCCL_TRY(c, g_nccl.AllGather(d_send, d_recv, HB, ncclInt8, c->comm->comm, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(hall.data(), d_recv, HB * c->nranks, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    cudaFree(d_send); cudaFree(d_recv);
    for (int q = 0; q < c->nranks; ++q) {
        if (q == c->rank) { c->mail_peer[q] = c->d_mail; continue; }
        void *ptr = nullptr;
        cudaIpcMemHandle_t h;
        memcpy(&h, hall.data() + (size_t)q * HB, HB);
        if (!ok2 || cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok2 = 0; cudaGetLastError(); ptr = nullptr; }
        c->mail_peer[q] = (double *)ptr;
    }
    if (!c->d_mail_table) CUDA_TRY(c, cudaMalloc((void **)&c->d_mail_table, ADP_MAX_RANKS * sizeof(double *)));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_mail_table, c->mail_peer, ADP_MAX_RANKS * sizeof(double *), cudaMemcpyHostToDevice, c->stream));
    flag = ok2;
    CUDA_TRY(c, cudaMemcpyAsync(c->d_scal + S_TMP1, &flag, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    NCCL_TRY(c, g_nccl.AllReduce(c->d_scal + S_TMP1, c->d_scal + S_TMP1, 1, ncclFloat64, ncclMin, c->comm->comm, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(&flag, c->d_scal + S_TMP1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->peer_ar = flag > 0.5;      // the min-all-reduce also guarantees every mailbox is zeroed before first use
    return ADP_OK;
}

void adp_comm_unmap_peers(adp_ctx *c)
{
    for (int b = 0; b < PB_COUNT; ++b) {
        if (c->peer_lo[b]) cudaIpcCloseMemHandle(c->peer_lo[b]);
        if (c->peer_hi[b]) cudaIpcCloseMemHandle(c->peer_hi[b]);
        c->peer_lo[b] = c->peer_hi[b] = nullptr;
    }
    for (int q = 0; q < ADP_MAX_RANKS; ++q) {
        if (c->mail_peer[q] && c->mail_peer[q] != c->d_mail) cudaIpcCloseMemHandle(c->mail_peer[q]);
        c->mail_peer[q] = nullptr;
    }
    c->peer_ar = false;
    c->peer_ok = false;
}

void adp_comm_destroy(adp_ctx *c)
{
    adp_comm_unmap_peers(c);
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

// Scalar all-reduce over peer memory as a kernel of its own (the reductions of the outer-iteration tail, the
// transient glue, ...; inside BiCGSTAB the two halves ride in the producing / consuming kernels, mail.cuh).
// ONE warp posts this rank's values into row `rank` of slot (seq % ADP_MAIL_SLOTS) of every rank's mailbox
// (values, release at system scope with the sequence number) and waits until every row of its own mailbox carries
// that sequence number; rows are combined in rank order (deterministic).  Like ncclAllReduce it is a barrier: it
// completes only after every rank's preceding kernel (and its halo pushes) has.
// A rank that waits longer than the timeout never combines the stale row: the result becomes NaN and the sticky
// S_FAULT slot is raised, which every host synchronisation turns into ADP_ERR_NCCL (ADP_CHECK_FAULT).
template <bool MAX>
__global__ void k_mail_allreduce(double *vals, int count, Mail m)
{
    // lane q talks to rank q: posts this rank's values into rank q's mailbox and polls
    // the row rank q writes into this rank's mailbox -- all ranks in parallel
    const int q = threadIdx.x;
    const unsigned long long seq = *m.seq + 1ull;
    const size_t slot = (size_t)(seq % ADP_MAIL_SLOTS) * m.nranks;
    double mine[4] = {0.0, 0.0, 0.0, 0.0}, got[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = 0; i < count; ++i) mine[i] = vals[i];
    if (q < m.nranks) {
        volatile double *dst = m.box[q] + (slot + m.rank) * ADP_MAIL_WORDS;
        for (int i = 0; i < count; ++i) dst[i] = mine[i];
        mail_st_release((unsigned long long *)(dst + 7), seq);
        const volatile double *src = m.mine + (slot + q) * ADP_MAIL_WORDS;
        const unsigned long long *flag = (const unsigned long long *)(m.mine + (slot + q) * ADP_MAIL_WORDS + 7);
        const long long t0 = clock64();
        bool ok = true;
        while (mail_ld_acquire(flag) != seq)
            if (clock64() - t0 > m.timeout) { ok = false; break; }
        if (ok) {
            for (int i = 0; i < count; ++i) got[i] = src[i];
        } else {
            for (int i = 0; i < count; ++i) got[i] = __longlong_as_double(0x7ff8000000000000LL);
            *(volatile double *)m.fault = 1.0;
        }
    }
    __syncwarp();
    // combine in rank order on lane 0 (deterministic)
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int r = 0; r < m.nranks; ++r)
        for (int i = 0; i < 4; ++i) {
            const double w = __shfl_sync(0xffffffffu, got[i], r);
            acc[i] = MAX ? ((w != w) ? w : fmax(acc[i], w)) : acc[i] + w;     // fmax would swallow the NaN of a timed-out row
        }
    if (q == 0) {
        for (int i = 0; i < count; ++i) vals[i] = acc[i];
        *m.seq = seq;
    }
}

// the wait half alone, for a reduction that was posted by a kernel but whose consumer is not a fused one
__global__ void k_mail_drain(MailWait w, double *scal)
{
    __shared__ double out[2];
    mail_wait(w.m, w.n, out, scal, w.slot[0], w.slot[1], true);
}

Mail adp_comm_mail(const adp_ctx *c)
{
    Mail m;
    m.box = c->d_mail_table; m.mine = c->d_mail; m.seq = c->d_arseq; m.fault = c->d_scal + S_FAULT;
    m.nranks = c->nranks; m.rank = c->rank; m.ll = c->mail_ll ? 1 : 0;
    m.timeout = (long long)(c->mail_timeout_s * 1.9e9);      // clock64 runs at the SM clock (<= 1.965 GHz)
    return m;
}

int adp_comm_drain(adp_ctx *c, int n, int slot0, int slot1)
{
    MailWait w;
    w.m = adp_comm_mail(c); w.n = n; w.slot[0] = slot0; w.slot[1] = slot1;
    k_mail_drain<<<1, 32, 0, c->stream>>>(w, c->d_scal);
    c->launches++;
    if (cudaPeekAtLastError() != cudaSuccess) { c->err = "k_mail_drain launch failed"; return ADP_ERR_CUDA; }
    return ADP_OK;
}

static int mail_allreduce(adp_ctx *c, double *d, int count, bool is_max)
{
    const Mail m = adp_comm_mail(c);
    if (count > 4) { c->err = "mail_allreduce: at most 4 values"; return ADP_ERR_USAGE; }
    if (is_max) k_mail_allreduce<true><<<1, 32, 0, c->stream>>>(d, count, m);
    else k_mail_allreduce<false><<<1, 32, 0, c->stream>>>(d, count, m);
    c->launches++;
    if (cudaPeekAtLastError() != cudaSuccess) { c->err = "k_mail_allreduce launch failed"; return ADP_ERR_CUDA; }
    return ADP_OK;
}

static bool debug_skip_ar()
{
    static int v = -1;
    if (v < 0) v = getenv("ADP_DEBUG_SKIP_AR") ? 1 : 0;   // timing experiments only: results are wrong
    return v == 1;
}

int adp_comm_allreduce_sum(adp_ctx *c, double *d, int count)
{
    if (c->nranks == 1 || debug_skip_ar()) return ADP_OK;
    if (c->peer_ar) return mail_allreduce(c, d, count, false);
    NCCL_TRY(c, g_nccl.AllReduce(d, d, count, ncclFloat64, ncclSum, c->comm->comm, c->stream));
    return ADP_OK;
}
int adp_comm_allreduce_max(adp_ctx *c, double *d, int count)
{
    if (c->nranks == 1 || debug_skip_ar()) return ADP_OK;
    if (c->peer_ar) return mail_allreduce(c, d, count, true);
    NCCL_TRY(c, g_nccl.AllReduce(d, d, count, ncclFloat64, ncclMax, c->comm->comm, c->stream));
    return ADP_OK;
}
// explicit NCCL all-reduce (results that do NOT come out of grid_reduce, e.g. the ndmax arg-max)
int adp_comm_allreduce_max_nccl(adp_ctx *c, double *d, int count)
{
    if (c->nranks == 1) return ADP_OK;
    NCCL_TRY(c, g_nccl.AllReduce(d, d, count, ncclFloat64, ncclMax, c->comm->comm, c->stream));
    return ADP_OK;
}
// Results back to the host on several ranks: the unchanged Fortran drivers (printers, th_upd's axial march,
// reactivity) read WHOLE sdata arrays, so by default every rank receives every slab -- rank q broadcasts its owned
// planes of one node column, the others stage them and copy them to the rows of rank q in their global host
// column.  Rare (once per outer*() call), so plain NCCL; d_owned = this rank's owned planes of the column.
int adp_comm_gather_column(adp_ctx *c, double *h_col, const double *d_owned)
{
    if (c->nranks == 1) return ADP_OK;
    const int base = c->nzz / c->nranks, rem = c->nzz % c->nranks;
    const size_t maxcnt = (size_t)(base + (rem ? 1 : 0)) * c->np;
    if (!c->d_gather) CUDA_TRY(c, cudaMalloc((void **)&c->d_gather, maxcnt * sizeof(double)));
    for (int q = 0; q < c->nranks; ++q) {
        const int k0q = q * base + (q < rem ? q : rem), nz = base + (q < rem ? 1 : 0);
        const size_t cnt = (size_t)nz * c->np;
        if (q == c->rank) {
            NCCL_TRY(c, g_nccl.Broadcast(d_owned, const_cast<double *>(d_owned), cnt, ncclFloat64, q, c->comm->comm, c->stream));
        } else {
            NCCL_TRY(c, g_nccl.Broadcast(c->d_gather, c->d_gather, cnt, ncclFloat64, q, c->comm->comm, c->stream));
            CUDA_TRY(c, cudaMemcpyAsync(h_col + (size_t)k0q * c->np, c->d_gather, cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        }
    }
    return ADP_OK;
}

int adp_comm_chain_recv(adp_ctx *c, double *d, int count)
{
    if (c->nranks == 1 || c->rank == 0) return ADP_OK;
    NCCL_TRY(c, g_nccl.Recv(d, count, ncclFloat64, c->rank - 1, c->comm->comm, c->stream));
    return ADP_OK;
}
int adp_comm_chain_send(adp_ctx *c, const double *d, int count)
{
    if (c->nranks == 1 || c->rank == c->nranks - 1) return ADP_OK;
    NCCL_TRY(c, g_nccl.Send(d, count, ncclFloat64, c->rank + 1, c->comm->comm, c->stream));
    return ADP_OK;
}
int adp_comm_allreduce_sum_nccl(adp_ctx *c, double *d, int count)
{
    if (c->nranks == 1) return ADP_OK;
    NCCL_TRY(c, g_nccl.AllReduce(d, d, count, ncclFloat64, ncclSum, c->comm->comm, c->stream));
    return ADP_OK;
}
int adp_comm_allreduce_min_ll(adp_ctx *c, long long *d, int count)
{
    if (c->nranks == 1) return ADP_OK;
    NCCL_TRY(c, g_nccl.AllReduce(d, d, count, ncclInt64, ncclMin, c->comm->comm, c->stream));
    return ADP_OK;
}
