// adp_internal.cuh -- device data model shared by the kernels and the C ABI.
//
// Layout in HBM (all fp64 unless noted).  The core is numbered k-major with a constant
// `np` nodes per z-plane (mod_io.f90:1319-1330, mod_cmfd.f90:174,208), so a z-slab is a
// contiguous node range.  Every node array on the device covers the planes
//      [-GH, nzl+GH)      GH = 2 ghost planes below and above the nzl owned planes
// and is indexed  idx = (kl + GH) * np + r   (kl local plane, r position inside the plane).
// Ghost planes hold the neighbouring slab's data (or zeros outside the core) so that the
// z neighbours idx -+ np are always addressable: the stencil is branch free, a missing
// neighbour simply has a zero coefficient.  x neighbours are idx -+ 1, y neighbours
// idx -+ ypm[r] / ypp[r] (plane-invariant tables, the jagged core outline).
//   vectors      f0[2][G] (ping-pong old/new), fs[2], r, rs, p, v, s, t           [NV]
//   matrix       a[G][7][NV]  diagonals in set_ind order z-,y-,x-,diag,x+,y+,z+
//   coupling     df[G][6][NV], dn[G][6][NV]   (reference AoS nod(n,g)%df(6),dn(6))
//   XS           D,sigr,nuf,sigf,exsrc [G][NV]; sigs[G(to h)][G(from g)][NV]; dc[6][G][NV]
//   nodal        S[3][G][NV]
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <map>

#include "../../include/adpres_b200.h"

#define ADP_GH 2          // ghost planes on each side
#define ADP_MAXG 16       // groups supported by the pointer tables in kernel params
#define ADP_TILE 256      // threads per block = rows of one plane per tile
#define ADP_NF 6          // delayed-neutron families (mod_data.f90:120)
#define ADP_MAXPART 4096  // max blocks contributing partial sums

// ---- device scalars (one array of doubles in HBM; slots) ---------------------------------
enum {
    S_KE = 0,     // k-eff
    S_F,          // f  = Integrate(fs0)   (new)
    S_FC,         // fc = previous f
    S_E1,         // l2 norm of previous fission-source difference
    S_EXC,        // extrapolation coefficient domiR / (1 - domiR)
    // --- reduction results; contiguous groups are all-reduced together across ranks
    S_RSV,        // (rs, v)
    S_TT,         // (t, t)
    S_TS,         // (t, s)
    S_RHO0,       // rho ping
    S_RHO1,       // rho pong
    S_E2SQ,       // sum errn^2
    S_FINT,       // sum vdel * fs
    S_SER,        // max rel. fission source change
    S_FER,        // max rel. flux change
    S_NDMAX,      // max |delta dn|
    S_POW,        // total power
    S_TMP0, S_TMP1,
    S_FAULT,      // sticky: set to 1 when a peer-memory all-reduce timed out (read back with the scalars at every host sync)
    S_COUNT
};

struct Geo {
    int np;        // nodes per plane
    int nzl;       // owned planes
    int nzz;       // global planes
    int k0;        // global index (0-based) of owned plane 0
    int tpp;       // tiles per plane = ceil(np / ADP_TILE)
    int ntiles;    // tpp * nzl
    long long NV;  // (nzl + 2 GH) * np
    int bc[6];     // xeast, xwest, ynorth, ysouth, zbott, ztop
    const int *ypm, *ypp;          // [np] offsets to the y-/y+ neighbour (0 = none)
    const unsigned char *flag;     // [np] bit0 x- missing, bit1 x+ missing, bit2 y- missing, bit3 y+ missing
    const double *hx, *hy;         // [np] node size in x, y at plane position r
    const double *hz;              // [nzz + 2] node size in z, hz[1 + kg]; ends padded
    const double *area;            // [np] xdel*ydel
    const int *ixr, *iyr;          // [np] 1-based i, j of plane position r
};
// radial map for kernels that walk (i, j) instead of plane positions (kept OUT of Geo: Geo is the first parameter of
// every streaming kernel and its layout is part of their tuned code generation)
struct GeoXY {
    int nxx = 0, nyy = 0;          // radial mesh
    const int *nodp = nullptr;     // [nyy][nxx] 1-based plane position of node (i, j), 0 outside the core outline
};

// Destination of a boundary-plane "push": the ghost planes of the z-neighbours' copy of the same
// vector, mapped into this process over NVLink peer memory (CUDA IPC).  lo = the lower
// neighbour's upper ghost plane, hi = the upper neighbour's lower ghost plane; nullptr = none.
struct Push {
    double *lo = nullptr, *hi = nullptr;
};
enum { PB_RS = 0, PB_V0, PB_V1, PB_R, PB_F0A, PB_F0B, PB_COUNT };

// In-kernel all-reduce over peer memory.  Every rank owns a mailbox
//     mail[ADP_MAIL_SLOTS][nranks][ADP_MAIL_WORDS]   (doubles; word 7 = sequence number)
// mapped into all other ranks.  The last CTA of a reduction posts its local results into slot
// (seq % ADP_MAIL_SLOTS), row `rank`, of EVERY rank's mailbox (values, system fence, sequence
// number) and then waits until all rows of its own mailbox carry that sequence number; the
// results are combined in rank order (deterministic).  Two consecutive reductions use different
// slots and a rank cannot run two reductions ahead of another, so slots are never overwritten early.
#define ADP_MAIL_SLOTS 4
#define ADP_MAIL_WORDS 16   // doubles per (slot, rank) row: [0..6] values + [7] sequence number (release/acquire form); [8..15] tagged words (LL form)
#define ADP_MAX_RANKS 8
struct Mail {
    double *const *box = nullptr;     // device table: box[q] = rank q's mailbox (box[rank] is local).  A table in
                                      // global memory, NOT an array inside the kernel parameters: indexing a
                                      // parameter array with a runtime q makes every thread copy it to a stack frame
    double *mine = nullptr;           // this rank's own mailbox (== box[rank])
    unsigned long long *seq = nullptr;   // device counter of reductions posted by this rank
    double *fault = nullptr;          // -> scal[S_FAULT]: sticky timeout flag (never shared with the STOP error flag)
    long long timeout = 0;            // clock64 ticks a rank waits for its peers before it gives up
    int nranks = 1, rank = 0;
    int ll = 1;                       // fused form: 1 = self-validating 8-byte words (no fence on the critical path), 0 = values + released flag
};
// what a kernel waits for in its prologue: the `n` sums the preceding kernel posted (n = 0: nothing); they are combined
// in rank order by every CTA, CTA 0 also stores them to scal[slot[i]] for the kernels that follow
struct MailWait {
    Mail m;
    int n = 0;
    int slot[2] = {0, 0};
};

#define FLAG_XM 1
#define FLAG_XP 2
#define FLAG_YM 4
#define FLAG_YP 8

struct adp_comm;  // comm.cu

struct adp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    // sizes
    int nxx = 0, nyy = 0, nzz = 0, nnod = 0, ng = 0, nmat = 0, np = 0;
    int k0 = 0, k1 = 0, nzl = 0;
    long long NV = 0, NL = 0;
    Geo geo{};
    GeoXY geoxy{};
    int bc[6] = {0, 0, 0, 0, 0, 0};
    // host copies of small geometry
    std::vector<int> h_ix, h_iy, h_iz;
    std::vector<int> h_nodp;               // (nxx,nyy) -> 1-based plane position, 0 outside the core outline
    std::vector<double> h_xdel, h_ydel, h_zdel;
    // control
    int nout = 500, nin = 2, nac = 5, nupd = 1000000, kern = ADP_KERN_SANM;
    double serc = 1e-5, ferc = 1e-5;
    double last_ser = 1.0, last_fer = 1.0;   // source / flux error of the last outer iteration
    double ndmax = 0.0;  // persists across outer*() calls, starts at 0 (mod_data.f90:199)
    int im = 0, jm = 0, km = 0;
    bool coup_first = true, have_flux = false, geometry_set = false, xs_set = false, matrix_ready = false;
    bool outer_first = true, outer_ad_first = true;
    // device arrays
    int *d_ypm = nullptr, *d_ypp = nullptr, *d_ixr = nullptr, *d_iyr = nullptr, *d_mat = nullptr, *d_nodp = nullptr;
    unsigned char *d_flag = nullptr;
    double *d_hx = nullptr, *d_hy = nullptr, *d_hz = nullptr, *d_area = nullptr;
    double *d_f0[2] = {nullptr, nullptr};  // [G][NV] each
    double *d_fs[2] = {nullptr, nullptr};
    int cur[ADP_MAXG] = {0};               // which of d_f0[.] holds the current flux of group g
    int fcur = 0;
    double *d_r = nullptr, *d_rs = nullptr, *d_p = nullptr, *d_v = nullptr, *d_s = nullptr, *d_t = nullptr;
    double *d_s0 = nullptr;                // scattering source of the last group swept (s0 quirk)
    int s0_group = 0;                      // 1-based group whose column of s0 is non-zero (0: none yet)
    double *d_a = nullptr;                 // [G][7][NV]
    double *d_df = nullptr, *d_dn = nullptr;  // [G][6][NV]
    double *d_D = nullptr, *d_sigr = nullptr, *d_nuf = nullptr, *d_sigf = nullptr, *d_exsrc = nullptr;
    double *d_sigs = nullptr;              // [h][g][NV]  = sigs(n,g,h)
    double *d_dc = nullptr;                // [f][g][NV]
    double *d_chi = nullptr;               // [g][nmat]
    double *d_S = nullptr;                 // [3][G][NV]
    double *d_nd = nullptr;                // [3][G*G+3G][NV] node-direction store of the nodal update (Bc, a2, a4, L1)
    double *d_abefgh = nullptr;            // [3][6][G][NV] SANM constants, valid until D / sigr change
    bool abefgh_valid = false;
    // transient
    double *d_af = nullptr;                // adjoint flux kept for reactivity() [G][NV]
    // material tables and control-rod data for the device-side XS update (adp_xs_update)
    double *d_xtab = nullptr, *d_dtab = nullptr;   // [xsigtr|xsiga|xnuf|xsigf (nmat,ng)] + xsigs (nmat,ng,ng); same for the rod increments
    int *d_fb = nullptr;                   // [np] control-rod bank of each plane position (0: none)
    double *d_bpos = nullptr, *d_dumtop = nullptr;
    int nb = 0;
    double coreh = 0.0, pos0 = 0.0, ssize = 0.0;
    double *d_c0 = nullptr, *d_ft = nullptr, *d_fst = nullptr, *d_omeg = nullptr, *d_sigrp = nullptr,
           *d_L = nullptr, *d_dfis = nullptr, *d_tbeta = nullptr, *d_velo = nullptr;
    double ibeta[ADP_NF] = {0}, lamb[ADP_NF] = {0};
    double sth = 1.0, bth = 0.0;
    bool kinetics_set = false;
    bool kin_xtab = false;                 // kinetics data per material (adp_set_kinetics_xtab, %XTAB decks)
    double *d_mkin = nullptr;              // [nmat][6] lamb | [nmat][6] iBeta | [nmat][ng] velo
    // thermal-hydraulic channel solve (th.cu): parameters of %THER and the state the reference keeps in sdata
    struct ThPar {
        double pi = 0, rf = 0, rg = 0, rc = 0, dia = 0, dh = 0, farea = 0, cflow = 0, cf = 0, tin = 0, enti = 0;
        double rpos[12] = {0}, rdel[12] = {0};
        int ntem = 0;
    } th;
    bool th_set = false, th_state_set = false, th_pline_set = false;
    double *d_ftab[4] = {nullptr, nullptr, nullptr, nullptr};   // feedback tables: bcon, ftem, mtem, cden (layout of d_xtab)
    double fref[4] = {0, 0, 0, 0}, bcon = 0.0;
    bool xs_feedback = false;              // the current adp_xs_update applies the feedback tables
    // %XTAB branch tables (adp_set_xtab): per-material dims / offsets, branch parameters, (un)rodded value blocks
    int *d_brmeta = nullptr;               // [nmat][6] nd, nb, nf, nm, trod, offset into d_brpar
    long long *d_brtoff = nullptr;         // [nmat] offset of the material's block in d_brtab / d_brrtab
    double *d_brpar = nullptr, *d_brtab = nullptr, *d_brrtab = nullptr;
    double *d_stab = nullptr;              // (ntem, 6) column-major
    double *d_tfm = nullptr;               // [nt+1][NV] radial pin temperatures
    double *d_heatf = nullptr, *d_ent = nullptr, *d_ftem = nullptr, *d_mtem = nullptr, *d_cden = nullptr, *d_frate = nullptr,
           *d_pline = nullptr, *d_nodenf = nullptr /*[np]*/, *d_chain = nullptr /*[2][np] entm, bfrate*/;
    // result reductions (results.cu): column / plane sums
    double *d_res = nullptr, *h_res = nullptr;
    size_t res_elems = 0;
    // reductions
    double *d_scal = nullptr;              // [S_COUNT]
    double *d_part = nullptr;              // [4][ADP_MAXPART]
    unsigned int *d_ticket = nullptr;
    long long *d_argidx = nullptr;         // location of ndmax
    int *d_errflag = nullptr;              // LU diagonal abort etc.
    double *h_scal = nullptr;              // pinned mirror
    int *h_flags = nullptr;                // pinned
    double *d_stage = nullptr;             // staging for host<->device copies of node arrays
    double *h_stage = nullptr;             // pinned staging [max(NL*?)]
    size_t stage_elems = 0;
    int grid_blocks = 0;                   // persistent grid size override (option "grid_blocks")
    bool grid_override = false;
    bool balance_rounds = false;   // measured: no effect on these bandwidth-bound kernels (A/B, tools/kbench.py)
    int sm_count = 0;
    // multi-rank
    adp_comm *comm = nullptr;
    int nranks = 1, rank = 0;
    double *d_v2 = nullptr;                // second v buffer (iteration parity; see bicg_core)
    double *d_gather = nullptr;            // staging for one slab of another rank (adp_comm_gather_column)
    bool gather_results = true;            // node arrays returned to the host are completed with the other ranks' slabs
    bool peer_ok = false;                  // neighbours' vectors are mapped: halos are pushed by the kernels
    double *peer_lo[PB_COUNT] = {nullptr}, *peer_hi[PB_COUNT] = {nullptr};
    int nzl_lo = 0, nzl_hi = 0;            // planes owned by the lower / upper neighbour
    long long NV_lo = 0, NV_hi = 0;
    bool xghost_valid[2][ADP_MAXG] = {{false}};   // ghost planes of f0[which][g] are current
    bool peer_ar = false;                  // reductions are all-reduced over peer-memory mailboxes
    bool fuse_mail = true;                 // BiCGSTAB: post in the producer's last CTA, wait in the consumer's prologue (no extra kernel)
    double mail_timeout_s = 30.0;          // ADP_MAIL_TIMEOUT_S
    bool mail_ll = true;                   // fused all-reduce posts self-validating words (mail.cuh)
    double *d_mail = nullptr;              // this rank's mailbox
    double *mail_peer[ADP_MAX_RANKS] = {nullptr};
    double **d_mail_table = nullptr;       // device copy of mail_peer[]
    unsigned long long *d_arseq = nullptr;
    // CUDA graphs of one outer iteration, keyed by (mode, parity pattern, extrapolate)
    std::map<unsigned long long, cudaGraphExec_t> graphs;
    std::map<unsigned long long, long long> graph_launches;   // kernels inside each graph
    bool use_graphs = true;
    int bench_warmup = 3;
    int nodal_coop = -1;                   // nodal kernel form: -1 automatic (quad kernels from G = 5), 0 one thread per item,
                                           // 1 sixteen lanes per surface (round 1's form for G >= 7), 2 quad kernels
    int nodal_fused = 0;                   // experiment, G <= 2: per-direction kernels that carry the node-direction record in
                                           // registers instead of storing it (bit-identical, 1.7x less DRAM traffic, but 2.7x
                                           // SLOWER: 8 warps per SM cannot hide the fp64 division chains -- DESIGN.md)
    bool fuse_st = true;                   // C kernel: s on the fly inside t = A s (false: k_s then k_t)
    int spmv_var = 4;                      // formulation of the B kernel (k_spmv_dot), single rank: 0..6 (4: loads grouped, 48 registers)
    int st_var = 6;                        // formulation of the C kernel (k_st), single rank: 0..7 (6: loads grouped, 64 registers)
    int st_m_var = 6;                      // the same for the multi-rank C kernel (k_st_m)
    bool st_tma = false;                   // C kernel staged with cp.async.bulk + mbarrier (experiment, single rank, np even)
    // option "lazy_adf": adp_set_xs only remembers the host pointers of dc and sigf -- the two arrays the CMFD iteration never
    // reads (ADFs: nodal update; sigf: PowDis) -- adp_outer_begin enqueues their upload on a second stream, and the first
    // consumer makes the main stream wait for it (adp_lazy_sync): 0.5 of the 1.0 GB an outer() call uploads then travels
    // while the outer iterations run.  The caller keeps both host arrays unchanged until a consumer has returned.
    bool lazy_adf = false, lazy_pending = false;
    const double *lazy_dc = nullptr, *lazy_sigf = nullptr;
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_lazy = nullptr;
    // per-launch profile (option "profile"): an event after every kernel launch of the CMFD path, tagged with the source line
    bool prof = false;
    cudaEvent_t prof_start = nullptr;
    std::vector<std::pair<int, cudaEvent_t>> prof_ev;
    // bookkeeping
    long long launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    adp_trace_fn trace = nullptr;
    void *trace_user = nullptr;
};

#define CUDA_TRY(ctx, call)                                                                     \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + \
                         ":" + std::to_string(__LINE__) + ")";                                  \
            return ADP_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

// after every host synchronisation that brought the device scalars to h_scal: a peer-memory all-reduce that timed
// out poisons its result with NaN and raises the sticky S_FAULT slot; no call may return ADP_OK after that
#define ADP_CHECK_FAULT(ctx)                                                                              \
    do {                                                                                                  \
        if ((ctx)->h_scal[S_FAULT] != 0.0) {                                                              \
            (ctx)->err = "peer-memory all-reduce timed out: a rank fell more than ADP_MAIL_TIMEOUT_S behind (results are poisoned with NaN)"; \
            return ADP_ERR_NCCL;                                                                          \
        }                                                                                                 \
    } while (0)

#define ADP_REQUIRE(ctx, cond, msg)                 \
    do {                                            \
        if (!(cond)) {                              \
            (ctx)->err = (msg);                     \
            return ADP_ERR_USAGE;                   \
        }                                           \
    } while (0)

// Persistent grids: one wave of exactly (SMs x resident CTAs per SM of THIS kernel) CTAs, each
// looping over tiles.  Sizing every kernel with one fixed multiple of the SM count leaves a
// ragged second wave whenever a kernel's registers allow fewer CTAs per SM (ncu: k_st at 40
// registers ran at 54 % achieved occupancy with a 8-per-SM grid).
template <typename K>
static inline int adp_grid(adp_ctx *c, K kernel, int ntiles)
{
    static std::map<const void *, int> cache;
    const void *key = (const void *)kernel;
    auto it = cache.find(key);
    int per_sm;
    if (it == cache.end()) {
        per_sm = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, ADP_TILE, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
        cache[key] = per_sm;
    } else per_sm = it->second;
    long long g = (long long)c->sm_count * per_sm;
    if (c->grid_blocks > 0 && c->grid_override) g = c->grid_blocks;
    if (g > ADP_MAXPART) g = ADP_MAXPART;
    if (ntiles < g) g = ntiles;
    if (g < 1) g = 1;
    // balance the rounds of the tile loop: with g CTAs the loop takes ceil(ntiles/g) rounds and the
    // last one is generally almost empty (18 050 tiles on 1 184 CTAs: 16 rounds, the last 24 % full);
    // the smallest grid that needs the same number of rounds fills all of them
    if (!c->grid_override && c->balance_rounds) {
        const long long rounds = (ntiles + g - 1) / g;
        g = (ntiles + rounds - 1) / rounds;
    }
    return (int)g;
}

int adp_lazy_enqueue(adp_ctx *c);   // capi.cu: start the deferred uploads (no-op without any)
int adp_lazy_sync(adp_ctx *c);      // capi.cu: the main stream waits for them; call before any kernel / copy that reads d_dc, d_sigf

static inline void adp_prof_mark(adp_ctx *c, int line)
{
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, c->stream);
    c->prof_ev.emplace_back(line, e);
}

// ---- launch wrappers implemented in the kernel files ---------------------------------------
// cmfd_kernels.cu
int adp_k_coup_coef(adp_ctx *c);
int adp_k_matrix_setup(adp_ctx *c);
int adp_k_init_flux(adp_ctx *c, int adjoint);
int adp_k_outer_begin(adp_ctx *c, int mode);
int adp_k_bicg_group(adp_ctx *c, int mode, int g /*0-based*/, int nin, bool write_s0);
int adp_k_bicg_raw(adp_ctx *c, int g, int imax, const double *d_b, double *d_x);
int adp_k_spmv(adp_ctx *c, int g, const double *d_x, double *d_v);
int adp_k_outer_tail(adp_ctx *c, int mode, bool extrapolate);
int adp_k_powdis(adp_ctx *c, double *d_pow);
int adp_k_scale_by_slot(adp_ctx *c, double *d_vec, int slot);
int adp_k_get_exsrc(adp_ctx *c, double ht);
int adp_k_integrate(adp_ctx *c, const double *d_vec, int slot);
int adp_k_xs_update(adp_ctx *c);
int adp_k_xs_update_xtab(adp_ctx *c);
int adp_th_alloc(adp_ctx *c);            // th.cu: node arrays of the thermal-hydraulic state
int adp_k_ipden(adp_ctx *c);
int adp_k_upden(adp_ctx *c, double ht);
int adp_k_begin_step(adp_ctx *c, double ht);
int adp_k_omeg(adp_ctx *c, double ht, int bextr);
int adp_k_reactivity(adp_ctx *c, const double *d_af, const double *d_sigr_for_rem);
// capi.cu: host (nnod, ncol) column-major, global  <->  device [col][NV], own planes of this rank
int adp_upload_nodes(adp_ctx *c, double *d, const double *h, int ncol);
int adp_download_nodes(adp_ctx *c, double *h, const double *d, int ncol);
// comm.cu: pass a per-channel vector up the z-slab chain (thermal-hydraulic march)
int adp_comm_chain_recv(adp_ctx *c, double *d_buf, int count);   // from rank - 1 (no-op on rank 0)
int adp_comm_chain_send(adp_ctx *c, const double *d_buf, int count);   // to rank + 1 (no-op on the last rank)
// nodal_kernels.cu
int adp_k_nodal_source(adp_ctx *c, int cmode);
int adp_k_nodal_update(adp_ctx *c, int cmode);
int adp_k_lxyz_total(adp_ctx *c, double *d_L);
// comm.cu
int adp_comm_halo(adp_ctx *c, double *d_vec, int nplanes);            // exchange ghost planes of one vector
int adp_comm_gather_column(adp_ctx *c, double *h_col, const double *d_owned);   // other ranks' slabs of one host column
int adp_comm_allreduce_sum(adp_ctx *c, double *d_scal, int count);
int adp_comm_allreduce_max(adp_ctx *c, double *d_scal, int count);
int adp_comm_allreduce_min_ll(adp_ctx *c, long long *d_val, int count);
int adp_comm_allreduce_max_nccl(adp_ctx *c, double *d_scal, int count);
int adp_comm_allreduce_sum_nccl(adp_ctx *c, double *d_vec, int count);   // any length (not a grid_reduce result)
void adp_comm_destroy(adp_ctx *c);
int adp_comm_map_peers(adp_ctx *c);                                     // after the vectors are allocated
Mail adp_comm_mail(const adp_ctx *c);                                   // mailbox descriptor for kernel parameters
int adp_comm_drain(adp_ctx *c, int n, int slot0, int slot1);            // wait for a posted reduction outside a fused consumer
void adp_comm_unmap_peers(adp_ctx *c);
static inline Push adp_push(const adp_ctx *c, int buf, long long goff_lo = 0, long long goff_hi = 0)
{
    Push ps;
    if (!c->peer_ok) return ps;
    if (c->peer_lo[buf]) ps.lo = c->peer_lo[buf] + goff_lo + (long long)(ADP_GH + c->nzl_lo) * c->np;
    if (c->peer_hi[buf]) ps.hi = c->peer_hi[buf] + goff_hi + (long long)(ADP_GH - 1) * c->np;
    return ps;
}
