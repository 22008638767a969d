// mail.cuh -- scalar all-reduce over NVLink peer memory, split into its two halves so that it can ride
// inside the BiCGSTAB kernels instead of costing a launch of its own:
//
//   mail_post   (producer, the warp that finishes grid_reduce in the LAST CTA of the kernel)
//               stores this rank's partial sums into row `rank` of slot (seq % ADP_MAIL_SLOTS) of EVERY
//               rank's mailbox and releases them with the sequence number (st.release.sys): lane q
//               talks to rank q, all ranks in parallel.
//   mail_wait   (consumer, warp 0 of EVERY CTA in the kernel prologue)
//               polls the rows of its OWN mailbox (local L2; the peers wrote them over NVLink) with
//               ld.acquire.sys until all carry the sequence number, and combines them in rank order
//               (deterministic, identical on every rank and in every CTA).  CTA 0 also stores the result to
//               the scalar array for the kernels that follow.
//
// Like the separate k_mail_allreduce kernel this is a barrier across the ranks: a CTA leaves the
// prologue only after every rank has finished the preceding kernel -- including the halo planes that
// kernel pushed into this rank's ghost planes (each pushing thread fences at system scope before its
// CTA takes the reduction ticket, and the last CTA posts only after all tickets are in).
// Slot re-use is safe with ADP_MAIL_SLOTS >= 2: a rank posts reduction n+1 only after it has seen every
// rank's post n, i.e. after every rank has finished reading slot n-1.
//
// A rank that waits longer than Mail::timeout (ADP_MAIL_TIMEOUT_S, default 30 s) gives up: the result is
// poisoned with NaN and the sticky fault slot is raised; every host synchronisation checks it
// (ADP_CHECK_FAULT) -- a late peer can never turn into a silently wrong sum.
#pragma once
#include "adp_internal.cuh"

__device__ __forceinline__ unsigned long long mail_ld_acquire(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mail_st_release(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Two wire formats.
//   ll = 0  values, then the sequence number with st.release.sys; the reader acquires the sequence number and then
//           reads the values.  The release makes the posting warp wait for the NVLink acknowledgement of its value
//           stores before the flag may go out: one more round trip in the tail of every producing kernel.
//   ll = 1  every 8-byte word carries its own validity tag: word 2i = (seq32 << 32) | low half of value i, word 2i+1 =
//           (seq32 << 32) | high half.  8-byte stores are single transactions on NVLink, so a reader that sees the tag
//           sees the payload; the words need no ordering among themselves and the post issues them without any fence
//           (the scheme of NCCL's LL protocol).  What still orders the HALO planes before the post is unchanged: every
//           pushing thread fences at system scope before its CTA takes the reduction ticket, and the post is issued by
//           the CTA that saw all tickets.
__device__ __forceinline__ size_t mail_row(const Mail &m, unsigned long long seq, int r)
{
    return ((size_t)(seq % ADP_MAIL_SLOTS) * m.nranks + r) * ADP_MAIL_WORDS;
}

// warp-collective (all 32 lanes of ONE warp); v0, v1 valid in every lane
static __device__ __forceinline__ void mail_post(const Mail m, int count, double v0, double v1)
{
    const int q = threadIdx.x & 31;
    const unsigned long long seq = *(volatile unsigned long long *)m.seq + 1ull;
    if (q < m.nranks) {
        double *dst = m.box[q] + mail_row(m, seq, m.rank);
        if (m.ll) {
            volatile unsigned long long *w = (volatile unsigned long long *)dst + 8;
            const unsigned long long tag = (seq & 0xffffffffull) << 32;
            const unsigned long long b0 = (unsigned long long)__double_as_longlong(v0), b1 = (unsigned long long)__double_as_longlong(v1);
            w[0] = tag | (b0 & 0xffffffffull);
            w[1] = tag | (b0 >> 32);
            if (count > 1) {
                w[2] = tag | (b1 & 0xffffffffull);
                w[3] = tag | (b1 >> 32);
            }
        } else {
            ((volatile double *)dst)[0] = v0;
            if (count > 1) ((volatile double *)dst)[1] = v1;
            mail_st_release((unsigned long long *)(dst + 7), seq);
        }
    }
    __syncwarp();
    if (q == 0) *(volatile unsigned long long *)m.seq = seq;
}

// warp-collective (all 32 lanes of ONE warp); lane 0 writes out[0..1] (shared memory of the CTA)
static __device__ __forceinline__ void mail_wait(const Mail m, int count, double *out, double *scal, int s0, int s1, bool store)
{
    const int q = threadIdx.x & 31;
    const unsigned long long seq = *(volatile unsigned long long *)m.seq;
    double g0 = 0.0, g1 = 0.0;
    if (q < m.nranks) {
        const double *src = m.mine + mail_row(m, seq, q);
        const long long t0 = clock64();
        bool ok = true;
        if (m.ll) {
            const volatile unsigned long long *w = (const volatile unsigned long long *)src + 8;
            const unsigned long long tag = seq & 0xffffffffull;
            unsigned long long w0, w1, w2 = tag << 32, w3 = tag << 32;
            for (;;) {
                w0 = w[0]; w1 = w[1];
                if (count > 1) { w2 = w[2]; w3 = w[3]; }
                if ((w0 >> 32) == tag && (w1 >> 32) == tag && (w2 >> 32) == tag && (w3 >> 32) == tag) break;
                if (clock64() - t0 > m.timeout) { ok = false; break; }
            }
            // acquire side: the ghost planes this CTA reads next were written before the tags.  One acquire LOAD of a word
            // that is already valid (LDG.STRONG.SYS + L1 invalidate) instead of a fence: __threadfence_system() here is a
            // MEMBAR.SC.SYS in every CTA of every consumer kernel and cost 0.35 ms per step at two ranks (round 2, A/B)
            (void)mail_ld_acquire((const unsigned long long *)src + 8);
            g0 = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
            g1 = __longlong_as_double((long long)((w2 & 0xffffffffull) | (w3 << 32)));
        } else {
            const unsigned long long *flag = (const unsigned long long *)(src + 7);
            while (mail_ld_acquire(flag) != seq)
                if (clock64() - t0 > m.timeout) { ok = false; break; }
            if (ok) {
                g0 = ((const volatile double *)src)[0];
                if (count > 1) g1 = ((const volatile double *)src)[1];
            }
        }
        if (!ok) {
            g0 = g1 = __longlong_as_double(0x7ff8000000000000LL);   // never combine a stale row
            *(volatile double *)m.fault = 1.0;
        }
    }
    __syncwarp();
    double a0 = 0.0, a1 = 0.0;
    for (int r = 0; r < m.nranks; ++r) {
        a0 = a0 + __shfl_sync(0xffffffffu, g0, r);
        a1 = a1 + __shfl_sync(0xffffffffu, g1, r);
    }
    if (q == 0) {
        out[0] = a0;
        out[1] = a1;
        if (store) {
            scal[s0] = a0;
            if (count > 1) scal[s1] = a1;
        }
    }
}

// kernel prologue: every thread of the CTA calls it; val[i] = the all-reduced sums (untouched when w.n == 0)
__device__ __forceinline__ void mail_prologue(const MailWait &w, double *scal, double (&val)[2])
{
    __shared__ double mw[2];
    if (w.n > 0) {   // uniform over the grid
        if (threadIdx.x < 32) mail_wait(w.m, w.n, mw, scal, w.slot[0], w.slot[1], blockIdx.x == 0);
        __syncthreads();
        val[0] = mw[0];
        val[1] = mw[1];
    }
}
