// CUDA kernels of the B200 SparseBasisSampler (sm_100a).
//
// Work decomposition: a GROUP of W lanes (W = 8, 16 or 32) owns one Markov chain, so a warp advances
// 32/W independent chains in lock step.  One Metropolis step has only a handful of rows, a few dozen
// coupled columns and several strictly sequential floating-point chains (the reference's summation
// orders are kept), so a full warp per chain leaves most lanes idle in most instructions; narrow groups
// share every issued instruction between several chains.  All control flow is warp-uniform (loop bounds
// are maxima over the warp's groups); per-group work is predicated.
//
// Reference semantics are cited per function (paths relative to /root/reference/src).  Floating point
// is IEEE double with the reference's evaluation order; the file is compiled with -fmad=false so no
// multiply-add is contracted.  The only arithmetic that is not bit-identical to the host reference is
// exp() (CUDA libdevice vs glibc), see DESIGN.md "Bit-exactness".
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "abm_device.h"

namespace abm {

constexpr int kWarpsPerCta = 4;
#ifndef ABM_IMP_EARLY
#define ABM_IMP_EARLY 0
#endif
constexpr bool kImpEarly = ABM_IMP_EARLY != 0;
// ABM_PHILOX_BLOCK: the lanes of a group draw the Philox blocks of the chain's next W iterations at once (one iteration per
// lane, handed over by shuffles) instead of every lane repeating the block of the current one; same counters, same uniforms.
#ifndef ABM_PHILOX_BLOCK
#define ABM_PHILOX_BLOCK 1
#endif
// ABM_SMEM_SCRATCH: blocked tree: the {DE', leaf} pairs a step stages for its commit stay in shared memory for as many
// leading coupled columns as 7 resident CTAs per SM leave room for (abm_problem_create sizes it); the rest goes through
// S.scratch as before.
// ABM_T2_SMEM: blocked tree: the sums of 1024 leaves stay in shared memory during a launch when they fit next to the other
// per-chain arrays at 7 CTAs per SM (32x32x16: 161 doubles per chain): one L2 round trip less per column draw, and their
// updates are shared-memory adds instead of L2 reductions.
#ifndef ABM_T2_SMEM
#define ABM_T2_SMEM 0
#endif
// ABM_END_DEFERRED: sample launches: the end-state effect list of an accepted proposal is requested when the proposal is
// accepted and applied after the commit (its round trip overlaps the commit) instead of through an out-of-line call.
#ifndef ABM_END_DEFERRED
#define ABM_END_DEFERRED 1
#endif
// ---- measurement switches of the record path (A/B runs of round 2, 32x32x16, 8192 chains, one B200; profiles/r2_ab_*.log)
// ABM_PIN_IDS: lane / chain ids made opaque to the compiler so that it keeps them instead of re-deriving them from the
// special registers inside the loop (fewer instructions, but the extra live registers spill: 473 M vs 489 M steps/s)
#ifndef ABM_PIN_IDS
#define ABM_PIN_IDS 0
#endif
// ABM_REC_BULK: the record window is filled by ONE bulk asynchronous copy (cp.async.bulk, completion on an mbarrier:
// SASS UBLKCP / SYNCS) instead of two LDG.128 + STS.128 per lane.  Measured slower here (465 M vs 489 M steps/s): the copy
// engine's latency sits on the dependent chain of a step (column draw -> record -> chain state) and nothing can overlap it.
#ifndef ABM_REC_BULK
#define ABM_REC_BULK 0
#endif
// ABM_REC_EVICT_LAST (with ABM_REC_BULK): the bulk copy carries an L2 evict_last hint; ABM_STATE_EVICT_FIRST: coupled-column
// state loads carry an evict_first hint.  Neither changed the L2 hit rate enough to matter (465 / 457 M).
#ifndef ABM_REC_EVICT_LAST
#define ABM_REC_EVICT_LAST 0
#endif
#ifndef ABM_STATE_EVICT_FIRST
#define ABM_STATE_EVICT_FIRST 0
#endif
// ABM_PF_CHUNKS: with the whole coupled-column list in the shared-memory window, the chain state of EVERY later chunk is
// prefetched (1: into L2, 2: into L1) as soon as the record has arrived; the chunk loop then finds it on chip instead of
// exposing a DRAM round trip per chunk.  Costs no registers.  481 -> 489 M steps/s.
#ifndef ABM_PF_CHUNKS
#define ABM_PF_CHUNKS 1
#endif
// ABM_PF_IMP: the trajectory words the importance change will read (own state and the four neighbours of each stale factor)
// are prefetched into L2 when the record has arrived; the reads themselves follow the coupled columns (489 -> 491 M)
#ifndef ABM_PF_IMP
#define ABM_PF_IMP 1
#endif
// ABM_DRAW_RECIDX: the leaf level of the column draw also fetches the record index entries of its 32 candidate columns (and
// keeps their signed leaves), so the drawn column's record address and delta_j are in registers when the draw completes: one
// L2 round trip (the index) and one L1 hit (the leaf's sign) leave the dependent chain of a step (16x16x8 stationary: 478 ->
// 485 M steps/s, profiles/r2_ab_p_hops.log)
#ifndef ABM_DRAW_RECIDX
#define ABM_DRAW_RECIDX 1
#endif
// ABM_RATIO_EARLY: measurement switch: root / (root + changeInSum) issued before the importance change instead of after it
// (16x16x8 stationary: 467 vs 478 M steps/s; off)
#ifndef ABM_RATIO_EARLY
#define ABM_RATIO_EARLY 0
#endif
// ABM_POOL: measurement switch: the two chains of a warp share their lanes over the coupled-column chunks and over the commit
// rounds: while both have chunks left each half-warp serves its own chain, afterwards both serve the longer list, two
// adjacent chunks per round, so a round is as long as HALF the two lists together instead of the longer one (3.9 -> 2.9
// rounds per warp-step at the stationary state of 32x32x16).  Every chunk is still handled by 16 consecutive lanes and the
// sums receive their deltas in ascending chunk order, so the arithmetic is bit for bit that of the unpooled loops (the GPU
// suite passes with it).  Measured SLOWER (32x32x16: 436 vs 487 M steps/s; 16x16x8 stationary: 429 vs 478,
// profiles/r2_ab_m_pool.log): serving the partner's chain puts two more dependent shared-memory reads (its header fields,
// then the slot word) in front of every chunk's chain-state load, and that costs more than the saved rounds.  Off.
#ifndef ABM_POOL
#define ABM_POOL 0
#endif
// ABM_SYN_LAZY: sample launches: the synopsis sums are advanced only when the synopsis changes (value x samples since the last
// change, like the end-state accumulators) instead of once per feasible sample.  Alone it changes nothing measurable (35.4 ->
// 35.5 ms per 8192 x 2000 steps, profiles/r2g_e2e_ab.log); it is what lets ABM_SMP_SMEM == 2 keep the synopsis out of the registers.
#ifndef ABM_SYN_LAZY
#define ABM_SYN_LAZY 1
#endif
#if ABM_SYN_LAZY && !ABM_END_DEFERRED
#error "ABM_SYN_LAZY needs ABM_END_DEFERRED"
#endif
// ABM_SMP_SMEM: sample launches keep their rarely used state out of the step loop's registers: whether (and with which sign) the
// accepted proposal moves the end state is one shared-memory word per chain, and the effect list is fetched AFTER the commit by
// the out-of-line endStateEffects from the record address kept in shared memory (instead of three registers holding the
// prefetched list across the commit); the series countdown lives in shared memory and is touched only when a series is
// recorded.  The sampling kernels carry 4 registers less through the loop, which is what made the compiler re-derive lane ids
// and pointers inside it (131 against 65 instructions per step in the Metropolis kernel; profiles/r2h_step_vs_sample_*).
// Sampling launch of 8192 x 2000 steps at 32x32x16: 35.2 ms (0) -> 34.8 ms (1) -> 34.7 ms (2); a Metropolis launch: 32.3 ms
// (profiles/r2i_e2e_ab.log).
#ifndef ABM_SMP_SMEM
#define ABM_SMP_SMEM 2
#endif
// ABM_SMP_SMEM == 2 (needs ABM_SYN_LAZY): the sample count and the synopsis of the current state live in shared memory as well;
// a sampling launch then carries NO register through the step loop that a Metropolis launch does not.
#if ABM_SMP_SMEM == 2 && !ABM_SYN_LAZY
#error "ABM_SMP_SMEM == 2 needs ABM_SYN_LAZY"
#endif
#if ABM_SMP_SMEM && ABM_POOL
#error "ABM_SMP_SMEM uses the shared-memory words of ABM_POOL"
#endif
#ifndef ABM_UNIT_ONLY
#define ABM_UNIT_ONLY 0
#endif
#ifndef ABM_SMEM_SCRATCH
#define ABM_SMEM_SCRATCH 0
#endif
constexpr unsigned kFull = 0xffffffffu;

enum { MODE_MCMC = 0, MODE_INIT = 1, MODE_SAMPLE = 2, MODE_SAMPLE_FIXED = 3 };
// SAMPLE = MCMC + per-feasible-sample statistics until a per-chain sample quota; SAMPLE_FIXED = the same for exactly
// nSteps steps (no quota), which keeps every chain active in every iteration like MODE_MCMC

struct RunParams {
    long long nSteps;            // Metropolis steps (MCMC) / iteration budget of this launch (INIT)
    const double *injU;          // injected uniforms [nChains][uStride] or null (Philox)
    long long uStride;
    long long uOffset;           // first uniform to consume in this launch
    const int32_t *injJ;         // injected proposals [nChains][nSteps] or null
    int32_t *outJ;               // traces [nChains][nSteps] or null
    uint8_t *outAcc;
    int32_t *outInf;
    unsigned long long seed, firstChain;
    long long *itersOut;         // INIT: iterations done, accumulated per chain
    long long nSamples;          // MODE_SAMPLE: feasible samples wanted per chain in this launch
    int32_t *series;             // MODE_SAMPLE: thinned synopsis series [nSeriesChains][seriesCap][nSynopsis] or null
    int nSeriesChains, seriesCap, seriesThin;
};

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 and libstdc++'s generate_canonical<double,53> (bits/random.tcc:3349-3381): the same two
// constructions as oracle/sbs_oracle.c, so GPU and oracle consume identical uniform streams.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll 1
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ double canonical_from_words(uint32_t w0, uint32_t w1) {
    double sum = __dadd_rn((double)w0, __dmul_rn((double)w1, 4294967296.0));
    double r = __dmul_rn(sum, 5.421010862427522170037264004349708557128906250e-20);  // exact 2^-64
    return r >= 1.0 ? 0.99999999999999988897769753748434595763683319091796875 : r;
}

// ------------------------------------------------------------------------------------------------
// energy(i,x) and infeasibility(i,x), SparseBasisSampler.h:341-345 and :267-271.  The factor closure is a
// host-tabulated table indexed by the clamped argument; f - kappa*0 == f exactly, so one expression
// serves the three branches.
__device__ __forceinline__ double energyOf(int x, int L, int U, int tabBase, const double *__restrict__ facVal,
                                           double kappa) {
    int c = max(L, min(x, U));
    int dist = x < L ? L - x : (x > U ? x - U : 0);
    double f = __ldg(facVal + tabBase + c);
    return __dsub_rn(f, __dmul_rn(kappa, (double)dist));
}
__device__ __forceinline__ int infeasOf(int x, int L, int U) { return x < L ? L - x : (x > U ? x - U : 0); }

__device__ __forceinline__ void prefetchL2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetchL1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// ------------------------------------------------------------------------------------------------
// Record fetch: one bulk asynchronous copy per proposal (cp.async.bulk: global -> shared memory by the copy engine, no
// registers, no per-lane address arithmetic) whose completion is counted in bytes on an mbarrier in shared memory.
constexpr int kRecWinWords = 128, kRecWinBytes = kRecWinWords * 4;   // the head of a record that is staged in shared memory
// ABM_FIX_WIN_WORDS: the window of the compile-time layout (stepKernel<..., FIX>).  128: as above, every record word is read
// through `index < window ? shared : global`.  288 (1152 bytes, the longest record of the paper's problems is 1104): the host
// selects the layout only when every record fits, and the FIX kernels read the window unconditionally (about 90 instructions
// per step less; 3.6 instead of 3.0 KB of shared memory per chain, still 7 CTAs per SM).  Measured (32x32x16, 8192 chains,
// profiles/r2l_pp32_ab.log): 511 -> 543 M chain-steps/s straight after the search, sampling launches 34.7 -> 33.5 ms.
#ifndef ABM_FIX_WIN_WORDS
#define ABM_FIX_WIN_WORDS 288
#endif
constexpr int kFixWinWords = ABM_FIX_WIN_WORDS, kFixWinBytes = kFixWinWords * 4;
constexpr bool kFixWinWhole = kFixWinWords > kRecWinWords;   // the FIX layout holds whole records
__device__ __forceinline__ uint32_t smemAddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void barInit(unsigned long long *bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void bulkLoad(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    const uint32_t b = smemAddr(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dst)), "l"(src),
                 "r"(bytes), "r"(b)
                 : "memory");
}
// the same copy with an L2 eviction-priority hint (createpolicy): the record table is what every chain re-reads
__device__ __forceinline__ unsigned long long policyEvictLast() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulkLoadHint(void *dst, const void *src, unsigned bytes, unsigned long long *bar, unsigned long long pol) {
    const uint32_t b = smemAddr(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smemAddr(dst)),
                 "l"(src), "r"(bytes), "r"(b), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ unsigned long long policyEvictFirst() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ double2 ldHint(const double2 *p, unsigned long long pol) {
    double2 v;
    asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void barWait(unsigned long long *bar, unsigned phase) {
    const uint32_t b = smemAddr(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "ABM_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra ABM_WAIT_%=;\n"
        "}\n" ::"r"(b),
        "r"(phase)
        : "memory");
}

// One out-of-line copy of exp(): the step kernel is bounded by instruction fetch as much as by anything else.
__device__ __noinline__ double expShared(double x) { return exp(x); }
// SparseBasisSampler.h:82
__device__ __forceinline__ double DEtoBasisProb(double DE) { return DE < 0.0 ? expShared(DE) : 1.0; }

// ------------------------------------------------------------------------------------------------
// One chain's HBM state.
struct Chain {
    int8_t *Xb;
    double2 *colRec;
    int8_t *delta;
    double *tier1, *tier2;
    double *top;      // shared-memory copy of the <= 32 highest nodes (index multiple of 2^topShift) while a step kernel runs
    int topShift;     // construction kernels hold no copy: top = tier2, topShift = 31 (only node 0 qualifies)
    __device__ __forceinline__ double *node(int i) const {
        if (i & 31) return &colRec[i].y;
        if (i & 1023) return &tier1[i >> 5];
        if (i & ((1 << topShift) - 1)) return &tier2[i >> 10];
        return &top[i >> topShift];
    }
};

// ------------------------------------------------------------------------------------------------
// PredPrey trajectory importance (TrajectoryImportance.h:90-113 on PredPreyAgent.h:124-182).
// An agent state's 6 act counts are one aligned 64-bit word of Xb.  posMask has bit 8a+7 set iff
// X[event(t,psi,a)] > 0; its popcount is State::fermionicBoundedForwardOccupation (State.h:66-74).
__device__ __forceinline__ unsigned long long posMask(unsigned long long w) {
    const unsigned long long m7 = 0x00007f7f7f7f7f7full, m8 = 0x0000808080808080ull;
    return (((w & m7) + m7) & ~w) & m8;
}
// newLogP of refresh() (TrajectoryImportance.h:96-108): lgamma(occ+1) if occ>1, then += log(pi/pibar) per act with X>0, in act
// order.  It depends only on which of the 6 acts are present, the agent type and the neighbour flag, so the host tabulates
// it with exactly that summation (abm_problem_create): impTab[31 + ((type*2 + neighbour) << 6) + presentActs].
__device__ __forceinline__ int presentActs(unsigned long long pm) {   // bit a = (X[act a] > 0), from posMask's bits 8a+7
    const unsigned lo = (unsigned)(pm >> 7), hi = (unsigned)(pm >> 39);
    // bits 0,8,16,24 of lo times (1 + 2^7 + 2^14 + 2^21) land, without carries, on bits 21..24
    return (int)(((lo * 0x00204081u) >> 21) & 15u) | (int)((hi & 1u) << 4) | (int)((hi >> 3) & 32u);
}
__device__ __forceinline__ double factorLogP(unsigned long long own, bool neighbour, int type,
                                             const double *__restrict__ impTab) {
    return __ldg(impTab + 31 + ((type * 2 + (neighbour ? 1 : 0)) << 6) + presentActs(posMask(own)));
}

// Class and influencer state ids of a state id from the agent table (DevProblem::agentTab) -- grids that are not a power of two
// and agents described by tables; cold path, the paper's grids use shifts.  An unused influencer slot reads the always-zero
// state word that follows the trajectory (index nStates).
struct StateInfo { int cls, n0, n1, n2, n3; };
__device__ __noinline__ StateInfo stateFromTable(int sid, int S, int nStates, const int32_t *__restrict__ tab) {
    const int t = sid / S, psi = sid - t * S, base = sid - psi;
    const int32_t *e = tab + 5 * psi;
    auto at = [&](int k) { const int v = __ldg(e + 1 + k); return v >= 0 ? base + v : nStates; };
    StateInfo r;
    r.cls = __ldg(e);
    r.n0 = at(0); r.n1 = at(1); r.n2 = at(2); r.n3 = at(3);
    return r;
}

__device__ __forceinline__ int warpSumInt(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

template <int W>
__device__ __forceinline__ int groupSumInt(int v) {
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
template <int W>
__device__ __forceinline__ int groupInclusiveScan(int v, int gl) {
#pragma unroll
    for (int o = 1; o < W; o <<= 1) {
        const int t = __shfl_up_sync(kFull, v, o, W);
        if (gl >= o) v += t;
    }
    return v;
}
__device__ __forceinline__ int warpMax(int v) { return __reduce_max_sync(kFull, v); }

// Blocked sum tree: one 32-ary level.  32 entries are spread over the W lanes of a group (lane gl holds entries gl*NG ..
// gl*NG+NG-1); they are consumed from the highest index down: entry e owns [sum_{e'>e} v, sum_{e'>=e} v).
// suffixScan: sfx = sum of the entries held by lanes >= gl (the group's total in lane 0).
template <int W>
__device__ __forceinline__ double suffixScan(const double (&v)[32 / W], int gl) {
    constexpr int NG = 32 / W;
    double own = 0.0;
#pragma unroll
    for (int k = NG - 1; k >= 0; --k) own = __dadd_rn(own, v[k]);
    double sfx = own;   // inclusive suffix sum over lanes >= gl
#pragma unroll
    for (int o = 1; o < W; o <<= 1) {
        const double t = __shfl_down_sync(kFull, sfx, o, W);
        if (gl + o < W) sfx = __dadd_rn(sfx, t);
    }
    return sfx;
}
// Returns the entry that owns `target` and rebases target to the start of that entry's interval.
template <int W>
__device__ __forceinline__ int selectFromScan(const double (&v)[32 / W], double sfx, double &target, int gl, int gBase) {
    constexpr int NG = 32 / W;
    const double aboveRaw = __shfl_down_sync(kFull, sfx, 1, W);
    const double above = gl + 1 < W ? aboveRaw : 0.0;             // entries held by higher lanes
    const unsigned hits = (__ballot_sync(kFull, target < sfx) >> gBase) & (W == 32 ? 0xffffffffu : ((1u << W) - 1u));   // gBase: the group's first lane
    const int winner = hits ? 31 - __clz(hits) : 0;               // highest lane whose suffix exceeds the target
    int e = gl * NG;                                               // rounding fallback: the lane's lowest entry
    double start = above;
    bool found = false;
#pragma unroll
    for (int k = NG - 1; k >= 0; --k) {
        if (!found) {
            const double next = __dadd_rn(start, v[k]);
            if (target < next || k == 0) { e = gl * NG + k; found = true; }
            else start = next;
        }
    }
    e = __shfl_sync(kFull, e, winner, W);
    start = __shfl_sync(kFull, start, winner, W);
    target = hits ? fmax(__dsub_rn(target, start), 0.0) : 0.0;
    return e;
}
template <int W>
__device__ __forceinline__ int selectDescending(const double (&v)[32 / W], double &target, int gl, int gBase) {
    return selectFromScan<W>(v, suffixScan<W>(v, gl), target, gl, gBase);
}

// Bytes of shared memory one chain (group) needs; must match the carve-up in stepKernel.
// nodeCap: 32 doubles of descent / summation staging (reference tree; 0 for the blocked tree); t2Cap: the blocked tree's sums of
// 1024 leaves when they are kept in shared memory for the launch (0 otherwise)
// suCap: the coupled-column ids of a step (reference tree: sCap; the blocked tree reads them from the record window: 0)
constexpr int kCtxBytes = ABM_SMP_SMEM == 2 ? 64 : 32;   // sCtx: 8 ints (+ 8 with the sampling state in shared memory)
__host__ __device__ inline int groupSmemBytes(int rowsCap, int sCap, int tCap, int scrCap, int maxAbs, int nodeCap, int t2Cap, int suCap,
                                              int winBytes = kRecWinBytes) {
    return (winBytes + 16 + kCtxBytes + scrCap * 16 + sCap * 8 + rowsCap * 16 * maxAbs + nodeCap * 8 + t2Cap * 8 + 32 * 8 + 16 * 8 + 8 + 8 * 4 + suCap * 4 + rowsCap * 4 + ((tCap * 2 + rowsCap * 2 + 7) & ~7) + 15) & ~15;
}

// MODE_SAMPLE: an accepted proposal (dj != 0) with events in the last timestep moves the end state ModelState(traj,T,T).
// The column's effect is static up to the sign delta_j; fx = its list in the record: {nEnd, 8 x int8 synopsis change per
// unit of delta_j, nEnd x (state | entry << 24)}.  The samples from index nSamp on see the new occupancy: their share
// occ_end * N - sum(change * nSamp) is settled by sampleEndKernel, so this only issues reductions and waits for nothing.
// Out of line: it runs in one step out of T and must not cost the step loop registers.  Returns the lane's synopsis change.
__device__ __noinline__ int endStateEffects(const int32_t *__restrict__ fx, int dj, int nSynopsis, EndRec *endRec, int nSamp, int gl,
                                            int W) {
    if (dj == 0) return 0;
    const int nEnd = __ldg(fx);
    for (int k = gl; k < nEnd; k += W) {
        const int w = __ldg(fx + 3 + k);
        atomicAdd(reinterpret_cast<unsigned long long *>(&endRec[w & 0xffffff].pend),
                  (unsigned long long)((long long)(dj * (w >> 24)) * (long long)nSamp));
    }
    return gl < nSynopsis ? dj * (int)(int8_t)(__ldg(fx + 1 + (gl >> 2)) >> (8 * (gl & 3))) : 0;
}

// every seriesThin-th feasible sample of a chain goes to its series (slot = sample index / seriesThin); the step loop counts
// the samples down to the next one, so this runs once per seriesThin samples
__device__ __noinline__ void recordSeries(int32_t *series, int seriesThin, int seriesCap, int nSynopsis, int chain, long long tot,
                                          int gl, int syn) {
    const long long slot = tot / seriesThin;
    if (slot < seriesCap) series[((size_t)chain * seriesCap + slot) * nSynopsis + gl] = syn;
}

// ------------------------------------------------------------------------------------------------
// MODE_MCMC: the loop body of nextSampleWithInfeasibleImportance (SparseBasisSampler.h:221-246).
// MODE_SAMPLE: the same, plus what the reference's statistics pipeline does with every feasible sample
// (FiguresForPaper.h:93-104): end-state occupancy and synopsis sums, maintained incrementally.
// MODE_INIT: findInitialFeasibleSolution (:254-263): same proposal, always applied, no importance, no
// statistics, until the chain is feasible.
//
// BLK selects the categorical's storage.  false: the reference's binary sum tree, node for node and add for add (above).
// true: a 32-ary blocked sum tree -- leaf p_u next to currentDE[u], sums of 32 leaves (tier1), of 1024 (tier2), of 32768
// (shared-memory top) -- with the reference's cdf order (the descent consumes the HIGH indices first, i.e. column j owns
// [sum_{i>j} p_i, sum_{i>=j} p_i)).  The same uniform then selects the same column except when it falls within rounding
// distance (~1e-13 relative) of a boundary; get(u) is the exact leaf instead of the reference's drifting difference of
// sums.  An update touches 3 nodes instead of 1 + popcount(u), which removes most of the commit work and traffic.
// FIX: the per-chain shared-memory layout is fixed at compile time (kFixRowsCap rows, kFixSCap coupled columns, unit tableau
// entries, blocked tree): every array is then a constant offset from one base register and the per-row tables have constant
// strides.  With the layout taken from the problem at run time the compiler re-derives the eleven array pointers inside the
// step loop (12 % of all instructions, profiles/r1d_*).  abm_problem_create selects it whenever the problem fits, which the
// paper's PredPrey problems do; anything else runs the same code with the run-time layout.
constexpr int kFixRowsCap = 24, kFixSCap = 176;
// ABM_FIX_SCR: measurement switch: coupled columns whose staged {DE', leaf} stay in shared memory between the proposal and its
// commit (16 = the first chunk: the commit's first round starts from shared memory instead of an L2 round trip; 16x16x8
// stationary: 472 vs 478 M steps/s; off)
#ifndef ABM_FIX_SCR
#define ABM_FIX_SCR 0
#endif
constexpr int kFixScrCap = ABM_POOL ? 0 : ABM_FIX_SCR;
template <int MODE, int W, bool BLK, bool FIX = false>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 7) stepKernel(DevProblem P, DevChains S, RunParams R) {
    static_assert(!FIX || BLK, "the fixed layout exists for the blocked tree only");
    extern __shared__ __align__(16) unsigned char smemRaw[];
    constexpr int NG = 32 / W;
    constexpr bool kMC = MODE != MODE_INIT;   // Metropolis test, importance, counters
    constexpr bool kSample = MODE == MODE_SAMPLE || MODE == MODE_SAMPLE_FIXED;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int gl = lane & (W - 1);
    const int g = lane / W;
    int gBase = lane & ~(W - 1);        // first lane of the group
    const int chain = (blockIdx.x * kWarpsPerCta + warp) * NG + g;
    int validBit = chain < S.nChains ? 1 : 0;
    int chainC = validBit ? chain : 0;   // pointer arithmetic only; every access is predicated
#if ABM_PIN_IDS == 1
    // With 72 registers per thread the compiler otherwise re-derives the lane, the chain and everything that hangs off them
    // from the special registers wherever they are used in the step loop (7 % of the executed instructions,
    // profiles/r2_stationary_*): an empty asm makes the values opaque, so they are kept (or, at worst, reloaded).
    asm volatile("" : "+r"(gl));
    asm volatile("" : "+r"(gBase));
    asm volatile("" : "+r"(validBit));
    asm volatile("" : "+r"(chainC));
#elif ABM_PIN_IDS == 2
    asm volatile("" : "+r"(validBit));
#elif ABM_PIN_IDS == 3
    asm volatile("" : "+r"(validBit));
    asm volatile("" : "+r"(chainC));
#endif
    const bool chainValid = validBit != 0;

    // ---- per-chain shared memory
    const int cRowsCap = FIX ? kFixRowsCap : P.rowsCap, cSCap = FIX ? kFixSCap : P.sCap, cTCap = FIX ? 0 : P.tCap;
    const int cScrCap = FIX ? kFixScrCap : P.scrCap, cNodeCap = FIX ? 0 : P.nodeCap, cT2Cap = FIX ? 0 : P.t2Cap, cMaxAbs = FIX ? 1 : P.maxAbs;
    const int cSUCap = BLK ? 0 : cSCap;
    constexpr int kWinWords = FIX ? kFixWinWords : kRecWinWords, kWinBytes = kWinWords * 4;
    constexpr bool kWinWhole = FIX && kFixWinWhole;   // every record lies inside the window: no bounds check on its words
    const int gBytes = groupSmemBytes(cRowsCap, cSCap, cTCap, cScrCap, cMaxAbs, cNodeCap, cT2Cap, cSUCap, kWinBytes);
    unsigned char *gs = smemRaw + (size_t)(warp * NG + g) * gBytes;
#if ABM_PIN_IDS == 1
    asm volatile("" : "+l"(gs));
#endif
    uint32_t *sRec = reinterpret_cast<uint32_t *>(gs);                   // [kWinWords] the proposed column's record (FIX), or its head
    [[maybe_unused]] unsigned long long *sBar = reinterpret_cast<unsigned long long *>(gs + kWinBytes);   // mbarrier of the record copy (ABM_REC_BULK)
    int *sCtx = reinterpret_cast<int *>(gs + kWinBytes + 16);          // [8] this step's record header fields, for the partner
                                                                         //   half-warp when it serves this chain (ABM_POOL)
    double2 *sScr = reinterpret_cast<double2 *>(gs + kWinBytes + 16 + kCtxBytes);  // [scrCap] blocked tree: {DE', signed leaf} of the first
                                                                         //   coupled columns of a step (the rest: S.scratch)
    double *sDl = reinterpret_cast<double *>(sScr + cScrCap);            // [sCap] delta of set() per coupled column
    double *sG = sDl + cSCap;                                            // [rowsCap][2 maxAbs] per row: DE' term for a coupled
                                                                         //   column whose flip moves the row by -maxAbs..-1, 1..maxAbs
    double *sNode = sG + 2 * cMaxAbs * cRowsCap;                         // [32] tree nodes of one descent round
    double *sT2 = sNode + cNodeCap;                                      // [t2Cap] blocked tree: sums of 1024 leaves, when they fit
    double *sTop = sT2 + cT2Cap;                                         // [32] the chain's highest tree nodes (root = sTop[0])
    long long *sSyn = reinterpret_cast<long long *>(sTop + 32);          // [2][8] MODE_SAMPLE: synopsis sum / sum of squares
    unsigned long long *sIter0 = reinterpret_cast<unsigned long long *>(sSyn + 16);   // the chain's Philox counter when the launch began
    int *sCnt = reinterpret_cast<int *>(sIter0 + 1);                     // [8] this launch's MCMCStatistics counts (lane 0 of the group)
    int *sU = sCnt + 8;                                                  // [sCap] coupled columns, ascending
    int *sXoff = sU + cSUCap;                                            // [rowsCap] Xb offsets of the rows
    uint16_t *sTask = reinterpret_cast<uint16_t *>(sXoff + cRowsCap);    // [tCap] commit tasks
    uint16_t *sXX = sTask + cTCap;                                       // [rowsCap] x | xn << 8 (int8 each)

    Chain ch;
    ch.Xb = S.Xb + (size_t)chainC * P.xStride;
    ch.colRec = S.colRec + (size_t)chainC * P.nCols;
    ch.delta = S.delta + (size_t)chainC * P.nCols;
    ch.tier1 = S.tier1 + (size_t)chainC * P.t1Len;
    ch.tier2 = S.tier2 + (size_t)chainC * P.t2Len;
    ch.top = sTop;
    ch.topShift = P.topShift;
    // the top of the sum tree lives in shared memory for the whole launch: the root, the first descent round and the
    // longest ordered sums of every commit never touch global memory
    double *tier3 = S.tier3 + (size_t)chainC * 32;   // BLK: sums of 32768 leaves (global home of the shared-memory top)
    const int nTop = BLK ? ((P.nCols - 1) >> 15) + 1 : P.nTop;
    for (int k = gl; k < 32; k += W) {
        double v = 0.0;
        if (chainValid && k < nTop) v = BLK ? tier3[k] : ch.tier2[(k << P.topShift) >> 10];
        sTop[k] = v;
    }
    const bool t2s = BLK && !FIX && P.t2Cap > 0;   // the second-highest level of the blocked tree lives in shared memory, too
    if (t2s)
        for (int k = gl; k < P.t2Len; k += W) sT2[k] = chainValid ? ch.tier2[k] : 0.0;
    __syncwarp();
    double2 *scratch = S.scratch + (size_t)chainC * P.sCap;
    const int scrChunks = !BLK ? 0 : (FIX ? kFixScrCap / W : P.scrCap / W);   // chunks of coupled columns whose staged values stay in shared memory

    double logImp = chainValid ? S.logImp[chain] : 0.0;
    int infeas = chainValid ? S.infeas[chain] : 0;
    // Registers are what bounds the residency of this kernel (72 per thread for 28 warps per SM), so what a step does not need in
    // a register lives in shared memory: the launch's counters, and the Philox counter base (iteration = base + itersDone)
    if (gl == 0) sIter0[0] = chainValid ? S.iter[chain] : 0ull;
    if (gl < 8) sCnt[gl] = 0;
#if ABM_REC_BULK
    if (gl == 0) {
        barInit(sBar);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    unsigned barPhase = 0u;             // parity of the record barrier's current phase (one phase per proposal)
#endif
#if ABM_REC_EVICT_LAST
    const unsigned long long recPolicy = policyEvictLast();
#endif
    __syncwarp();
    int itersDone = 0;                  // iterations (init + MH) this chain made in this launch (a launch is at most 2^30 steps)
    constexpr bool kPool = FIX && W == 16 && ABM_POOL != 0;
    // a pointer into this group's shared memory, moved to the same array of group `grp` of the warp
    auto ofGroup = [&](auto *ptr, int grp) { return reinterpret_cast<decltype(ptr)>(reinterpret_cast<unsigned char *>(ptr) + (grp - g) * gBytes); };
    const int chainBase0 = (blockIdx.x * kWarpsPerCta + warp) * NG;   // the chain of the warp's group 0
    const int nCols = P.nCols;
    const int kTwoM = FIX ? 2 : P.twoM, kMaxAbs = FIX ? 1 : P.maxAbs;   // FIX: unit tableau entries (constants after inlining)
    // ABM_PHILOX_BLOCK: lane gl of a group draws the Philox block of the chain's iteration iter + gl, i.e. W iterations at once,
    // and every step reads its two uniforms from one lane by shuffle; same counters, same uniforms as one block per step
    constexpr bool kPhBlk = ABM_PHILOX_BLOCK != 0;
    constexpr int PB = W;
    int phIdx = PB;                     // next unread iteration of the drawn Philox block (PB: none left)
    double pu1 = 0.5, pu2 = 0.0;        // this lane's drawn uniforms

    // ---- MODE_SAMPLE state: current end-state occupancy ModelState(traj,T,T) (ModelState.h:27-36 via
    // State::backwardOccupation, State.h:57-63, and PredPreyAgent::consequences, PredPreyAgent.h:98-115), kept
    // incrementally, with per-state accumulators advanced lazily (occupancy x samples since the last change)
#if ABM_SMP_SMEM == 2
    int &nSamp = sCtx[0];               // feasible samples taken in this launch (written by lane 0 of the chain)
    int &syn = sCtx[8 + (gl & 7)];      // lane k < nSynopsis: agents inside the k-th nested square (FiguresForPaper.h:245-263);
                                        // only lanes below nSynopsis (<= 8) write it
    if (gl == 0) sCtx[0] = 0;
    if (gl < 8) sCtx[8 + gl] = 0;
#else
    int nSamp = 0;                      // feasible samples taken in this launch
    int fxN = 0, fxK = 0, fxS = 0;      // pending end-state effect of an accepted proposal (ABM_END_DEFERRED)
    int syn = 0;                        // lane k < nSynopsis: agents inside the k-th nested square (FiguresForPaper.h:245-263)
#endif
    int serLeft = 0x7fffffff;           // feasible samples until the next one that goes to the thinned series
    if (kSample) {   // end-state occupancy and synopsis were prepared by sampleBeginKernel
        if (R.series && chainValid && chainC < R.nSeriesChains) {
            const long long r = S.nSamples[chainC] % R.seriesThin;
            serLeft = r == 0 ? 0 : (int)(R.seriesThin - r);
        }
        if (ABM_SMP_SMEM && gl == 0) { sCtx[1] = serLeft; sCtx[3] = 0; }
        if (chainValid && gl < P.nSynopsis) syn = S.synNow[(size_t)chain * P.nSynopsis + gl];
        if (gl < 16) sSyn[gl] = 0;
        if (ABM_SYN_LAZY && gl == 0) sCtx[7] = 0;   // samples of this launch already folded into the synopsis sums
        __syncwarp();
    }
    // ABM_SYN_LAZY: before the synopsis changes, the samples since the last change are folded into the sums with the old value
    // (called where the two chains of a warp may have diverged: the barriers name the chain's own lanes)
    auto foldSynopsis = [&](int nSampNow) {
        const unsigned own = (W == 32 ? 0xffffffffu : ((1u << W) - 1u)) << gBase;
        const int cnt = nSampNow - sCtx[7];
        __syncwarp(own);
        if (gl < P.nSynopsis) {
            sSyn[gl] += (long long)syn * cnt;
            sSyn[8 + gl] += (long long)syn * syn * cnt;
        }
        if (gl == 0) sCtx[7] = nSampNow;
        __syncwarp(own);
    };

    const int nStepsLaunch = (int)R.nSteps;
#pragma unroll 1
    for (int s = 0; s < nStepsLaunch; ++s) {
        const bool act = chainValid && (MODE == MODE_MCMC || MODE == MODE_SAMPLE_FIXED || (MODE == MODE_INIT && infeas > 0) ||
                                        (MODE == MODE_SAMPLE && (long long)nSamp < R.nSamples));
        if ((MODE == MODE_INIT || MODE == MODE_SAMPLE) && !__any_sync(kFull, act)) break;

        // ---- uniforms: one Philox block per iteration, or the injected stream
        double u1 = 0.5, u2 = 0.0;
        if (act) {
            if (R.injU) {
                const double *up = R.injU + (size_t)chainC * R.uStride + R.uOffset;
                if (MODE == MODE_INIT) u1 = up[s];
                else { u1 = up[2 * s]; u2 = up[2 * s + 1]; }
            } else if constexpr (!kPhBlk) {
                uint32_t w[4];
                const unsigned long long iter = sIter0[0] + (unsigned long long)itersDone, chainId = R.firstChain + (unsigned long long)chainC;
                philox4x32_10((uint32_t)iter, (uint32_t)(iter >> 32), (uint32_t)chainId, (uint32_t)(chainId >> 32),
                              (uint32_t)R.seed, (uint32_t)(R.seed >> 32), w);
                u1 = canonical_from_words(w[0], w[1]);
                u2 = canonical_from_words(w[2], w[3]);
            }
        }
        if constexpr (kPhBlk) {
            if (!R.injU) {
                const bool need = act && phIdx >= PB;
                if (__any_sync(kFull, need)) {
                    if (need) {   // lane gl: the block of iteration iter + gl
                        const unsigned long long it = sIter0[0] + (unsigned long long)(itersDone + gl), chainId = R.firstChain + (unsigned long long)chainC;
                        uint32_t w[4];
                        philox4x32_10((uint32_t)it, (uint32_t)(it >> 32), (uint32_t)chainId, (uint32_t)(chainId >> 32),
                                      (uint32_t)R.seed, (uint32_t)(R.seed >> 32), w);
                        pu1 = canonical_from_words(w[0], w[1]);
                        pu2 = canonical_from_words(w[2], w[3]);
                        phIdx = 0;
                    }
                }
                const double s1 = __shfl_sync(kFull, pu1, phIdx & (PB - 1), W), s2 = __shfl_sync(kFull, pu2, phIdx & (PB - 1), W);
                if (act) { u1 = s1; u2 = s2; ++phIdx; }
            }
        }

        // ---- column draw: MutableCategoricalArray::operator()(RNG&), MutableCategoricalArray.h:119-132.
        // LW tree levels per memory round trip: lane r fetches the node reached by branch pattern r, then the
        // reference's strict comparisons / subtractions are resolved with shuffles.  A child index >= nCols
        // holds 0.0: "0.0 > target" is false and "target - 0.0" is a no-op, which is the reference's skip.
        double root = 1.0;
        int j = 0;
        uint2 drawnRi[NG], drawnRiSel = make_uint2(0u, 0u);   // ABM_DRAW_RECIDX: index entries of the leaf level's candidates
        int drawnLeafSign[NG], drawnSign = 0;
        bool drawnOk = false;
#pragma unroll
        for (int k = 0; k < NG; ++k) { drawnRi[k] = make_uint2(0u, 0u); drawnLeafSign[k] = 0; }
        if constexpr (BLK) {
            // top level (sums of 32768 leaves, shared memory): one scan over the group's lanes gives the total and the
            // selection; absent entries are 0.0 (sTop is zero beyond nTop)
            double vt[NG];
#pragma unroll
            for (int k = 0; k < NG; ++k) vt[k] = sTop[gl * NG + k];
            const double sfxT = suffixScan<W>(vt, gl);
            const double tot = __shfl_sync(kFull, sfxT, 0, W);
            if (act) root = tot;
            double target = __dmul_rn(u1, root);
            int sel = selectFromScan<W>(vt, sfxT, target, gl, gBase);
            // three 32-ary levels: tier2 (1024 leaves each), tier1 (32 leaves), leaves
#pragma unroll
            for (int lvl = 2; lvl >= 0; --lvl) {
                const int first = sel * 32;
                const int count = lvl == 2 ? P.t2Len : (lvl == 1 ? P.t1Len : nCols);
                double v[NG];
#pragma unroll
                for (int k = 0; k < NG; ++k) {
                    const int e = first + gl * NG + k;
                    v[k] = 0.0;
                    if (act && e < count) {   // leaf sign = delta_e; tiers in global memory are updated by L2 reductions
                        if (lvl == 0) {
                            const double leaf = ch.colRec[e].y;
                            v[k] = fabs(leaf);
                            if constexpr (ABM_DRAW_RECIDX != 0) {
                                drawnLeafSign[k] = __double2hiint(leaf);
                                drawnRi[k] = __ldg(P.recIdx + e);
                            }
                        } else {
                            v[k] = lvl == 2 ? (t2s ? sT2[e] : __ldcg(ch.tier2 + e)) : __ldcg(ch.tier1 + e);
                        }
                    }
                }
                const int w = selectDescending<W>(v, target, gl, gBase);
                if constexpr (ABM_DRAW_RECIDX != 0) {
                    if (lvl == 0) {   // the winner's index entry and leaf sign, from the lane that loaded them
                        unsigned rx = drawnRi[0].x, ry = drawnRi[0].y;
                        int sg = drawnLeafSign[0];
#pragma unroll
                        for (int k = 1; k < NG; ++k)
                            if ((w % NG) == k) { rx = drawnRi[k].x; ry = drawnRi[k].y; sg = drawnLeafSign[k]; }
                        drawnRiSel.x = __shfl_sync(kFull, rx, w / NG, W);
                        drawnRiSel.y = __shfl_sync(kFull, ry, w / NG, W);
                        drawnSign = __shfl_sync(kFull, sg, w / NG, W);
                        drawnOk = first + w < nCols;
                    }
                }
                sel = first + w;
            }
            j = min(sel, nCols - 1);
        } else {
            if (act) root = sTop[0];
        {
            double target = __dmul_rn(u1, root);
            #pragma unroll 1
            // rounds are aligned with the storage tiers (levels [0,5), [5,10), [10,15), ...): the 31 candidates of a round
            // are then contiguous in memory, and the first (partial) round reads the shared-memory top
            for (int hi = P.treeDepth - 1; hi >= 0;) {
                const int lo = (hi / 5) * 5;
                const int nl = hi - lo + 1;
                // all candidates of a round live in one storage tier (their lowest set bit is in [lo, lo + 5))
                const double *base = &ch.colRec[0].y;
                int sh = 0, mul = 2;
                if (lo >= P.topShift) { base = sTop; sh = P.topShift; mul = 1; }
                else if (lo >= 10) { base = ch.tier2; sh = 10; mul = 1; }
                else if (lo >= 5) { base = ch.tier1; sh = 5; mul = 1; }
#pragma unroll
                for (int k = 0; k < NG; ++k) {   // 31 candidate nodes of this round, 32/W per lane
                    const int r = gl + W * k;
                    const int c = j + (r << lo);
                    double v = 0.0;
                    if (act && r > 0 && r < (1 << nl) && c < nCols) v = base[(c >> sh) * mul];
                    sNode[r] = v;
                }
                __syncwarp();
                int pref = 0;
                #pragma unroll 1
                for (int b = nl - 1; b >= 0; --b) {
                    const int rc = pref | (1 << b);
                    const double vv = sNode[rc];
                    if (vv > target) pref = rc; else target = __dsub_rn(target, vv);
                }
                j += pref << lo;
                hi = lo - 1;
                __syncwarp();
            }
        }
        }
        if (act && R.injJ) j = R.injJ[(size_t)chainC * R.nSteps + s];

        // ---- Proposal::Proposal (:288-315): record of column j.  The head of the record -- rows, stale factors, coupled columns --
        // is staged in the chain's shared-memory window by two coalesced 16-byte loads per lane (or, ABM_REC_BULK, by one bulk
        // asynchronous copy whose completion is counted on the chain's mbarrier); lists longer than the window (a few per cent
        // of the proposals) are read from the table itself.
        uint2 ri = make_uint2(0u, 0u);
        const bool haveDrawn = BLK && ABM_DRAW_RECIDX != 0 && drawnOk && !R.injJ;
        if (act) ri = haveDrawn ? drawnRiSel : __ldg(P.recIdx + j);   // {offset in 16-byte units, bytes}
        const uint32_t *__restrict__ recG = P.rec + (size_t)ri.x * 4;
#if !ABM_REC_BULK
        if (act) {
            const unsigned nv = min(ri.y, (unsigned)kWinBytes) >> 4;
            for (unsigned q = gl; q < nv; q += W) reinterpret_cast<uint4 *>(sRec)[q] = __ldg(reinterpret_cast<const uint4 *>(recG) + q);
        }
#elif ABM_REC_EVICT_LAST
        if (act && gl == 0) bulkLoadHint(sRec, recG, min(ri.y, (unsigned)kWinBytes), sBar, recPolicy);
#else
        if (act && gl == 0) bulkLoad(sRec, recG, min(ri.y, (unsigned)kWinBytes), sBar);
#endif
        // BLK: delta_u is the sign bit of the stored leaf p_u (p_u >= 0), so a coupled column costs one 16-byte load
        int dj = 1;
        if (act) dj = BLK ? ((haveDrawn ? drawnSign : __double2hiint(ch.colRec[j].y)) < 0 ? -1 : 1) : (int)ch.delta[j];
#if !ABM_REC_BULK
        __syncwarp();
#else
        if (act) { barWait(sBar, barPhase); barPhase ^= 1u; }
#endif
        auto recWord = [&](int i) -> uint32_t { return (kWinWhole || i < kWinWords) ? sRec[i] : __ldg(recG + i); };
        const uint32_t h0 = act ? sRec[0] : 0u, h1 = act ? sRec[1] : 0u;
        const int n = h0 & 0xff, nSF = (h0 >> 8) & 0xff, C = h0 >> 16;
        const int nEv = h1 & 0xff, nXt = (h1 >> 8) & 0x3fff;
        const int wide = (h1 >> 30) & 1;                                    // slots of two words: {u, pairs}                 // nEv: leading rows that are events (row id < nEvents)
        const bool colTouchesEnd = (int)h1 < 0;                             // the column has an event row in the last timestep
        const int uBase = act ? (int)sRec[2] : 0;
        const int staleOff = 4 + 2 * n, slotOff = staleOff + nSF, xtraOff = slotOff + (C << wide);
        [[maybe_unused]] const int fxOff = xtraOff + nXt;   // end-state effects (read here only without ABM_SMP_SMEM)
        // slot l of the record: its pair word, and the column it names
        auto slotWord = [&](int l) -> uint32_t { return recWord(slotOff + (l << wide) + wide); };
        auto slotCol = [&](int l, uint32_t w) -> int { return wide ? (int)recWord(slotOff + 2 * l) : uBase + (int)(w & 0xffffu); };

        // rows of the proposed column, and -- requested in the same memory round trip -- the chain state of the first chunk of
        // coupled columns
        const int nW = warpMax(n);
        int dinf = 0;
        uint32_t rw0 = 0u, rw1 = 0u;
        int x0 = 0;
        if (gl < n) {
            rw0 = sRec[4 + 2 * gl];
            rw1 = sRec[5 + 2 * gl];
            x0 = ch.Xb[rw0 & 0xffffff];
        }
        uint32_t swN = 0xffff0000u;   // slot word of the coming chunk (0xffff....: no pair)
        double2 crN = make_double2(0.0, 0.0);
        int duN = 0;
        bool realN = gl < C;
        int uN = 0;
        if (realN) {
            swN = slotWord(gl);
            uN = slotCol(gl, swN);
            crN = ABM_STATE_EVICT_FIRST ? ldHint(ch.colRec + uN, policyEvictFirst()) : ch.colRec[uN];
            if constexpr (!BLK) duN = ch.delta[uN];
        }
#if ABM_PF_IMP
        if (kMC && gl < nSF && P.gShift >= 0) {
            const int sidP = (int)recWord(staleOff + gl);
            const int G = P.G, GG = G * G;
            const int psi = sidP & (P.S - 1), cell = psi & (GG - 1), type = psi >> (2 * P.gShift);
            const int cy = cell >> P.gShift, cx = cell & (G - 1), ob = (sidP - psi) + (1 - type) * GG;
            const unsigned long long *X8 = reinterpret_cast<const unsigned long long *>(ch.Xb);
            prefetchL2(X8 + sidP);
            prefetchL2(X8 + ob + ((cx - 1) & (G - 1)) + G * cy);   // left / right share a line with high probability
            prefetchL2(X8 + ob + cx + G * ((cy + 1) & (G - 1)));
            prefetchL2(X8 + ob + cx + G * ((cy - 1) & (G - 1)));
        }
#endif
#if ABM_PF_CHUNKS
        {
            const int cW = warpMax(C);
            #pragma unroll 1
            for (int l = W + gl; l < cW; l += W)
                if (l < C) {
                    const double2 *pa = ch.colRec + slotCol(l, wide ? 0u : slotWord(l));
                    if (ABM_PF_CHUNKS == 2) prefetchL1(pa); else prefetchL2(pa);
                }
        }
#endif
        #pragma unroll 1
        for (int rb = 0; rb < nW; rb += W) {
            const int k = rb + gl;
            if (k < n) {
                uint32_t w0 = rw0, w1 = rw1;
                int x = x0;
                if (rb != 0) { w0 = sRec[4 + 2 * k]; w1 = sRec[5 + 2 * k]; x = ch.Xb[w0 & 0xffffff]; }
                const int xoff = w0 & 0xffffff, m = (int)w0 >> 24;
                const int L = (int)(int8_t)(w1 >> 16), U = (int)w1 >> 24;
                const double *__restrict__ eRow = P.eTab + (w1 & 0xffffu);                   // energy(i, x) = eRow[x]
                const int xn = x + m * dj;                                                  // :299-300
                sXoff[k] = xoff;
                sXX[k] = (uint16_t)((x & 0xff) | ((xn & 0xff) << 8));
                // dE = energy(x') - energy(x) (:302), and what this row adds to DE' of a coupled column u whose flip would move
                // the row by s = M_iu delta_u (:309-310): (energy(x'+s) - energy(x+s)) - dE, for s = -maxAbs..-1, 1..maxAbs
                // (PredPrey tableaux on the paper's problems have unit entries: two values per row)
                const double dE = __dsub_rn(__ldg(eRow + xn), __ldg(eRow + x));
                if constexpr (FIX) {
                    sG[2 * k] = __dsub_rn(__dsub_rn(__ldg(eRow + xn - 1), __ldg(eRow + x - 1)), dE);
                    sG[2 * k + 1] = __dsub_rn(__dsub_rn(__ldg(eRow + xn + 1), __ldg(eRow + x + 1)), dE);
                } else {
#pragma unroll 1
                    for (int d = 1; d <= kTwoM; ++d) {   // the shifts in ascending order
                        const int sh = 2 * d <= kTwoM ? d - kMaxAbs - 1 : d - kMaxAbs;
                        sG[kTwoM * k + d - 1] = __dsub_rn(__dsub_rn(__ldg(eRow + xn + sh), __ldg(eRow + x + sh)), dE);
                    }
                }
                dinf += infeasOf(xn, L, U) - infeasOf(x, L, U);                              // :303
            }
        }
        const int infN = infeas + groupSumInt<W>(dinf);
        if constexpr (kPool) {
            if (gl == 0) {
                sCtx[0] = C; sCtx[1] = slotOff; sCtx[2] = uBase; sCtx[3] = wide; sCtx[4] = xtraOff; sCtx[5] = j; sCtx[6] = (int)ri.x;
            }
        }
        __syncwarp();

        // coupled columns, one per lane, ascending.  A column's DE' is currentDE[u] plus one term per shared
        // row in row order -- the order in which the reference's map accumulation visits them (:305-312).
        double changeInSum = 0.0;
        const int chunksW = warpMax((C + W - 1) / W);
        // The importance change (perturb + refresh) can be evaluated between issuing the first chunk's loads and consuming
        // them (ABM_IMP_EARLY), so that its own loads of the trajectory overlap with them, or after the chunks.
        double logImpN = logImp;
        auto importance = [&]() {
            // ---- importance: perturb + refresh (TrajectoryImportance.h:36-54, 90-113)
            const int sfW = warpMax(nSF);
            const int evW = warpMax(nEv);
            if (sfW > 0) {
                const int G = P.G, GG = G * G;
                const unsigned long long *X8 = reinterpret_cast<const unsigned long long *>(ch.Xb);
                #pragma unroll 1
                for (int fb = 0; fb < sfW; fb += W) {
                    const int f = fb + gl;
                    const bool fa = f < nSF;
                    int sid = -1, type = 0, ns[4] = {-1, -1, -1, -1};
                    unsigned long long own = 0ull, nb[4] = {0ull, 0ull, 0ull, 0ull};
                    if (fa) {
                        sid = (int)recWord(staleOff + f);
                        if (P.gShift >= 0) {   // power-of-two grids: psi = x + G y + G^2 type (PredPreyAgent.h:59) by shifts
                            const int psi = sid & (P.S - 1), cell = psi & (GG - 1);
                            type = psi >> (2 * P.gShift);
                            const int cy = cell >> P.gShift, cx = cell & (G - 1);
                            const int ob = (sid - psi) + (1 - type) * GG;
                            // PredPreyAgent.h:199-207: left, right, up, down, opposite species, torus
                            const int xl = cx == 0 ? G - 1 : cx - 1, xr = cx == G - 1 ? 0 : cx + 1;
                            const int yu = cy == G - 1 ? 0 : cy + 1, yd = cy == 0 ? G - 1 : cy - 1;
                            ns[0] = ob + xl + G * cy;
                            ns[1] = ob + xr + G * cy;
                            ns[2] = ob + cx + G * yu;
                            ns[3] = ob + cx + G * yd;
                        } else {               // any other grid, any other agent: the states timestep() reads, from the table
                            const StateInfo si = stateFromTable(sid, P.S, P.nStates, P.agentTab);
                            type = si.cls;
                            ns[0] = si.n0; ns[1] = si.n1; ns[2] = si.n2; ns[3] = si.n3;
                        }
                        own = X8[sid];
#pragma unroll
                        for (int q = 0; q < 4; ++q) nb[q] = X8[ns[q]];
                    }
                    unsigned long long ownN = own, nbN[4] = {nb[0], nb[1], nb[2], nb[3]};
                    // apply the proposal's event changes (the X swap at :225-228) to the loaded words
                    #pragma unroll 1
                    for (int r = 0; r < evW; ++r) {   // event rows come first in a column (ascending row ids)
                        if (fa && r < nEv) {
                            const int xo = sXoff[r];
                            const int rs = xo >> 3, sh = (xo & 7) * 8;
                            const unsigned long long ins = (unsigned long long)(sXX[r] >> 8) << sh, msk = ~(0xffull << sh);
                            if (rs == sid) ownN = (ownN & msk) | ins;
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (rs == ns[q]) nbN[q] = (nbN[q] & msk) | ins;
                        }
                    }
                    double diff = 0.0;
                    if (fa) {
                        const bool fl = (posMask(nb[0]) | posMask(nb[1]) | posMask(nb[2]) | posMask(nb[3])) != 0ull;
                        const bool flN = (posMask(nbN[0]) | posMask(nbN[1]) | posMask(nbN[2]) | posMask(nbN[3])) != 0ull;
                        diff = __dsub_rn(factorLogP(ownN, flN, type, P.impTab), factorLogP(own, fl, type, P.impTab));  // :109
                    }
                    // totalLogProb += newLogP - oldLogP in stale-factor insertion order (:109).  Most factors do not change:
                    // their difference is exactly +0.0 and x + 0.0 == x bit for bit (the total is never -0.0), so only the
                    // non-zero differences are added, in the same order.
                    unsigned nz = (__ballot_sync(kFull, fa && diff != 0.0) >> gBase) & (W == 32 ? 0xffffffffu : ((1u << W) - 1u));
#pragma unroll 1
                    while (__any_sync(kFull, nz != 0u)) {
                        const double dq = __shfl_sync(kFull, diff, nz ? __ffs(nz) - 1 : 0, W);
                        if (nz) logImpN = __dadd_rn(logImpN, dq);
                        nz &= nz - 1u;
                    }
                }
            }
        };
        if (kMC && kImpEarly) importance();
        double ratioEarly = 1.0;
        double csLane = 0.0;   // BLK: this lane's share of changeInSum
        // chunks of 16 coupled columns: mine, the partner group's, and how the two half-warps share them (ABM_POOL)
        const int nChMine = (C + W - 1) / W;
        int nChOth = 0, nChBoth = 0, poolRounds = 0;
        bool poolMineLong = false;
        if constexpr (kPool) {
            nChOth = __shfl_xor_sync(kFull, nChMine, W);
            nChBoth = min(nChMine, nChOth);
            poolMineLong = nChMine > nChOth;
            poolRounds = nChBoth + ((max(nChMine, nChOth) - nChBoth + 1) >> 1);
        }
        // round r: (chain group, chunk) this half-warp serves, or nothing
        auto poolAssign = [&](int r, int nMine_, int nOth_, int nBoth_, bool mineLong_, int &grp, int &c) -> bool {
            if (r < nBoth_) { grp = g; c = r; return true; }
            const int k = r - nBoth_;
            if (mineLong_) { grp = g; c = nBoth_ + 2 * k; return c < nMine_; }
            grp = 1 - g; c = nBoth_ + 2 * k + 1;
            return c < nOth_;
        };
        if constexpr (kPool) {
            const double2 *colRec0 = S.colRec + (size_t)chainBase0 * nCols;
            double2 *scratch0 = S.scratch + (size_t)chainBase0 * P.sCap;
            // the slot a lane handles in round r: its pair word, column and chain state (requested one round ahead)
            auto fetch = [&](int r, uint32_t &sw_, int &u_, double2 &cr_, bool &real_) {
                sw_ = 0xffff0000u; u_ = 0; cr_ = make_double2(0.0, 0.0); real_ = false;
                int grp, c;
                if (r < poolRounds && poolAssign(r, nChMine, nChOth, nChBoth, poolMineLong, grp, c)) {
                    const int *cx = ofGroup(sCtx, grp);
                    const int slot = c * W + gl;
                    if (slot < cx[0]) {
                        const uint32_t *recX = ofGroup(sRec, grp);
                        const uint32_t *recGX = P.rec + (size_t)(unsigned)cx[6] * 4;
                        const int so = cx[1], wd = cx[3];
                        auto rw = [&](int i) -> uint32_t { return (kWinWhole || i < kWinWords) ? recX[i] : __ldg(recGX + i); };
                        sw_ = rw(so + (slot << wd) + wd);
                        u_ = wd ? (int)rw(so + 2 * slot) : cx[2] + (int)(sw_ & 0xffffu);
                        cr_ = (colRec0 + (size_t)grp * nCols)[u_];
                        real_ = true;
                    }
                }
            };
            // round 0 is this half-warp's own first chunk, requested before the rows were evaluated (swN, uN, crN, realN) --
            // unless its chain proposes nothing in this step and it serves the partner's from the start
            if (nChMine == 0) fetch(0, swN, uN, crN, realN);
            #pragma unroll 1
            for (int r = 0; r < poolRounds; ++r) {
                const uint32_t sw = swN;
                const double2 cr = crN;
                const bool real = realN;
                const int u = uN;
                fetch(r + 1, swN, uN, crN, realN);
                int grp = g, c = 0;
                poolAssign(r, nChMine, nChOth, nChBoth, poolMineLong, grp, c);
                const int slot = c * W + gl;
                const int *cx = ofGroup(sCtx, grp);
                const double *sGX = ofGroup(sG, grp);
                const int du = __double2hiint(cr.y) < 0 ? -1 : 1;
                const bool self = real && u == cx[5];
                const double nodeOld = fabs(cr.y);
                double acc = cr.x;                                                           // try_emplace(u, currentDE[u])
                const unsigned p0 = (sw >> 16) & 0xffu, p1 = sw >> 24;
                auto rowTerm = [&](unsigned pb) -> double {   // unit entries: the row moves by M_iu delta_u = +-1
                    const int k = pb & 31, code = pb >> 5;
                    return sGX[2 * k + (((code >> 2) & 1) == (du < 0 ? 1 : 0) ? 1 : 0)];
                };
                const bool ext = real && p0 == (unsigned)kPairExt;
                if (real && p0 < (unsigned)kPairExt) {
                    acc = __dadd_rn(acc, rowTerm(p0));                                        // :308-310
                    if (p1 != (unsigned)kPairNone) acc = __dadd_rn(acc, rowTerm(p1));
                }
                if (__any_sync(kFull, ext)) {   // pair lists of columns sharing more than two rows with the proposed one
                    const uint32_t *recX = ofGroup(sRec, grp);
                    const uint32_t *recGX = P.rec + (size_t)(unsigned)cx[6] * 4;
                    auto rw = [&](int i) -> uint32_t { return (kWinWhole || i < kWinWords) ? recX[i] : __ldg(recGX + i); };
                    const int lw = cx[4] + (int)p1;
                    unsigned long long lst = 0ull;
                    if (ext) lst = (unsigned long long)rw(lw) | ((unsigned long long)rw(lw + 1) << 32);
                    const int cnt = (int)(lst & 0xffull);
                    const int tW = warpMax(cnt);
#pragma unroll 1
                    for (int t = 0; t < tW; ++t) {
                        if (t < cnt) {
                            const int bi = 1 + t;
                            const unsigned pb = bi < 8 ? (unsigned)(lst >> (8 * bi)) & 0xffu : (rw(lw + (bi >> 2)) >> (8 * (bi & 3))) & 0xffu;
                            acc = __dadd_rn(acc, rowTerm(pb));
                        }
                    }
                }
                if (self) acc = -cr.x;                                                       // :294
                const double newP = real ? DEtoBasisProb(acc) : 0.0;
                if (real) {
                    ofGroup(sDl, grp)[slot] = __dsub_rn(newP, nodeOld);   // get(u) is the stored leaf: the change of the leaf
                    // the stored leaf carries delta_u in its sign; the proposed column's flips on acceptance (:334)
                    (scratch0 + (size_t)grp * P.sCap)[slot] = make_double2(acc, ((du < 0) != self) ? -newP : newP);
                }
            }
            __syncwarp();
            // changeInSum: this chain's leaf changes, summed exactly as the unpooled loop sums them (lane, then chunk order)
            if (kMC) {
#pragma unroll 1
                for (int c = 0; c < nChMine; ++c) {
                    const int slot = c * W + gl;
                    csLane = __dadd_rn(csLane, slot < C ? sDl[slot] : 0.0);
                }
            }
        } else {
        #pragma unroll 1
        for (int c = 0; c < chunksW; ++c) {
            const int slot = c * W + gl;
            const uint32_t sw = swN;
            const double2 cr = crN;
            const bool real = realN;
            const int u = uN;
            const int du = BLK ? (__double2hiint(cr.y) < 0 ? -1 : 1) : duN;
            // the coming chunk's slot words are in shared memory already; its chain state is requested now
            swN = 0xffff0000u;
            crN = make_double2(0.0, 0.0);
            duN = 0;
            realN = slot + W < C;
            if (realN) {
                swN = slotWord(slot + W);
                uN = slotCol(slot + W, swN);
                crN = ABM_STATE_EVICT_FIRST ? ldHint(ch.colRec + uN, policyEvictFirst()) : ch.colRec[uN];
                if constexpr (!BLK) duN = ch.delta[uN];
            }
            const bool self = real && u == j;
            // (reference tree) the nodes get(u) / set(u,.) will read: u + 1, 2, 4, 8, 16 sit in the same 32-record block
            const int nDesc = (!BLK && real) ? (u ? (__ffs(u) - 1) : P.treeDepth) : 0;
            double dn[5];
            double nodeOld = BLK ? fabs(cr.y) : cr.y;
            if constexpr (!BLK) {
                const double *nearBase = &ch.colRec[real ? u : 0].y;
#pragma unroll
                for (int i = 0; i < 5; ++i) dn[i] = (i < nDesc && u + (1 << i) < nCols) ? nearBase[2 << i] : 0.0;
                if (real && (u & 31) == 0) nodeOld = *ch.node(u);
            }
            double acc = cr.x;                                                               // try_emplace(u, currentDE[u])
            // the rows u shares with the proposed column, in row order: pair = row k of the column | code of M_iu << 5
            const unsigned p0 = (sw >> 16) & 0xffu, p1 = sw >> 24;
            auto rowTerm = [&](unsigned pb) -> double {
                const int k = pb & 31, code = pb >> 5;
                if constexpr (FIX) {   // unit entries: the row moves by M_iu delta_u = +-1
                    const int up = ((code >> 2) & 1) == (du < 0 ? 1 : 0) ? 1 : 0;
                    return sG[2 * k + up];
                } else {
                    const int mu = (code & 4) ? -((code & 3) + 1) : (code & 3) + 1;
                    const int sv = mu * du;                                                  // the row's move if u flipped
                    return sG[kTwoM * k + kMaxAbs + sv - (sv > 0 ? 1 : 0)];
                }
            };
            const bool ext = real && p0 == (unsigned)kPairExt;
            if (real && p0 < (unsigned)kPairExt) {
                acc = __dadd_rn(acc, rowTerm(p0));                                                // :308-310
                if (p1 != (unsigned)kPairNone) acc = __dadd_rn(acc, rowTerm(p1));
            }
            if (__any_sync(kFull, ext)) {   // columns sharing more than two rows with the proposed one: their pair lists
                // {count, pair, pair, ...} one byte each; the first two words (seven pairs) are fetched at once and the pairs
                // picked out of registers, longer lists (rare) continue from the record
                const int lw = xtraOff + (int)p1;
                unsigned long long lst = 0ull;
                if (ext) lst = (unsigned long long)recWord(lw) | ((unsigned long long)recWord(lw + 1) << 32);
                const int cnt = (int)(lst & 0xffull);
                const int tW = warpMax(cnt);
#pragma unroll 1
                for (int t = 0; t < tW; ++t) {
                    if (t < cnt) {
                        const int bi = 1 + t;
                        const unsigned pb = bi < 8 ? (unsigned)(lst >> (8 * bi)) & 0xffu : (recWord(lw + (bi >> 2)) >> (8 * (bi & 3))) & 0xffu;
                        acc = __dadd_rn(acc, rowTerm(pb));
                    }
                }
            }
            if (self) acc = -cr.x;                                                           // :294
            // new probability, old leaf value get(u) = tree[u] - descendantSum(u) (MutableCategoricalArray.h:78,
            // .cpp:30-39), and the sum / delta that set(u, p') will compute (.cpp:10-17).  Absent descendants were
            // loaded as 0.0 and x + 0.0 == x, so the unconditional adds reproduce the reference's skips.
            const double newP = real ? DEtoBasisProb(acc) : 0.0;
            double term, dlt, base;
            if constexpr (BLK) {
                term = __dsub_rn(newP, nodeOld);   // the leaf itself is stored: get(u) = p_u
                dlt = term;
                base = newP;
            } else {
                double dsum = 0.0, ssum = newP;
    #pragma unroll
                for (int i = 0; i < 5; ++i) {
                    dsum = __dadd_rn(dsum, dn[i]);
                    ssum = __dadd_rn(ssum, dn[i]);
                }
                const int descW = warpMax(nDesc);
                if (descW > 5) {
                    // columns with (u & 31) == 0 (one in 32).  Levels 5..9: u + 2^i has its low 5 bits clear and bit i set, so it
                    // is tier1[(u >> 5) + 2^(i-5)]; levels >= 10 are tier2 / shared-memory top nodes.
                    const double *t1 = ch.tier1 + (real ? (u >> 5) : 0);
    #pragma unroll
                    for (int i = 0; i < 5; ++i) dn[i] = (5 + i < nDesc && u + (32 << i) < nCols) ? t1[1 << i] : 0.0;
    #pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        dsum = __dadd_rn(dsum, dn[i]);
                        ssum = __dadd_rn(ssum, dn[i]);
                    }
    #pragma unroll 1
                    for (int i = 10; i < descW; ++i) {   // one column in 1024
                        const double t = (i < nDesc && u + (1 << i) < nCols) ? *ch.node(u + (1 << i)) : 0.0;
                        dsum = __dadd_rn(dsum, t);
                        ssum = __dadd_rn(ssum, t);
                    }
                }
                term = __dsub_rn(newP, __dsub_rn(nodeOld, dsum));                   // :279-280
                dlt = __dsub_rn(ssum, nodeOld);                                              // delta of set(), .cpp:17
                base = ssum;
            }
            if (real) {
                if constexpr (!BLK) sU[slot] = u;
                sDl[slot] = dlt;
                // BLK: the stored leaf carries delta_u in its sign; the proposed column's flips on acceptance (:334)
                if constexpr (BLK) base = ((du < 0) != self) ? -base : base;
                if (c < scrChunks) sScr[slot] = make_double2(acc, base); else scratch[slot] = make_double2(acc, base);
            }
            if (kMC) {
                if constexpr (BLK) {
                    // the blocked tree's get(u) is the exact leaf, so changeInSum already differs from the reference's
                    // drifting differences of sums in the last bits: a fixed reduction tree replaces the serial order
                    csLane = __dadd_rn(csLane, real ? term : 0.0);
                } else {
                    // changeInSum += ... in ascending column order (:278-281), staged through shared memory; padding
                    // slots carry term == +0.0 and x + 0.0 == x
                    sNode[gl] = real ? term : 0.0;
                    __syncwarp();
#pragma unroll 4
                    for (int q = 0; q < W; ++q) changeInSum = __dadd_rn(changeInSum, sNode[q]);
                    __syncwarp();
                }
            }
        }
        }
        if constexpr (BLK && kMC) {
#pragma unroll
            for (int o = W / 2; o > 0; o >>= 1) csLane = __dadd_rn(csLane, __shfl_xor_sync(kFull, csLane, o));
            changeInSum = csLane;
        }
        __syncwarp();

        bool accept = act;
        if (kMC) {
            if constexpr (ABM_RATIO_EARLY != 0) ratioEarly = root / __dadd_rn(root, changeInSum);
            if (!kImpEarly) importance();
            // ---- calcRatioOfSums (:276-284) and the Metropolis test (:232-234)
            const double ratio = ABM_RATIO_EARLY != 0 ? ratioEarly : root / __dadd_rn(root, changeInSum);
            const double acceptance = __dmul_rn(expShared(__dsub_rn(logImpN, logImp)), ratio);
            accept = act && (u2 < acceptance);
            const int ci = (infeas == 0 ? 2 : 0) + (infN == 0 ? 1 : 0);                      // MCMCStatistics.cpp:12-17
            if (act && gl == 0) {
                sCnt[ci] += 1;
                if (accept) sCnt[4 + ci] += 1;
            }
        }
        if (R.outJ && act && gl == 0) {
            const size_t o = (size_t)chainC * R.nSteps + s;
            R.outJ[o] = j;
            if (R.outAcc) R.outAcc[o] = accept ? 1 : 0;
            if (R.outInf) R.outInf[o] = infN;
        }

        if (kSample) {
            // events of the last timestep changed by an accepted proposal move the end state (evaluated here, while the
            // record pointers of the proposal are still live)
#if ABM_SMP_SMEM
            // the sign with which the accepted proposal moves the end state (0: it does not) and its record, for after the commit
            if (gl == 0) { sCtx[2] = (int)ri.x; sCtx[3] = (accept && colTouchesEnd) ? dj : 0; }
#elif ABM_END_DEFERRED
            // the effect list is requested now and applied after the commit, so its round trip overlaps the commit
            if (accept && colTouchesEnd) {
                const int32_t *__restrict__ fx = reinterpret_cast<const int32_t *>(recG) + fxOff;
                fxN = __ldg(fx) * dj;                    // nEnd, signed by delta_j (nEnd >= 1 for such a column)
                fxK = __ldg(fx + 3 + gl);                // speculative beyond nEnd: the record table is padded
                fxS = __ldg(fx + 1 + (gl >> 2));
                if (abs(fxN) > W) {
                    fxN = 0;
                    if (ABM_SYN_LAZY) foldSynopsis(nSamp);
                    syn += endStateEffects(fx, dj, P.nSynopsis, S.endRec + (size_t)chainC * P.S, nSamp, gl, W);
                }
            }
#else
            if (__any_sync(kFull, accept && colTouchesEnd))
                syn += endStateEffects(reinterpret_cast<const int32_t *>(recG) + fxOff, (accept && colTouchesEnd) ? dj : 0, P.nSynopsis, S.endRec + (size_t)chainC * P.S, nSamp,
                                       gl, W);
#endif
        }

        // ---- applyProposal (:322-336)
        if (accept) {
            logImp = logImpN;
            infeas = infN;
            if (!BLK && gl == 0) ch.delta[j] = (int8_t)(-dj);
        }
        #pragma unroll 1
        for (int rb = 0; rb < nW; rb += W) {
            const int k = rb + gl;
            if (accept && k < n) ch.Xb[sXoff[k]] = (int8_t)(sXX[k] >> 8);
        }
        // For every changed column u (ascending): currentDE[u] = DE'; tree.set(u, p').  Every set() delta
        // depends only on pre-commit node values (an update never touches nodes at or above its own index
        // before it runs), so they were computed above.  What remains is, per distinct tree node, the
        // ORDERED sum of the deltas of the changed columns below it (MutableCategoricalArray.cpp:18-27):
        // a "task" = (first contributing entry l, tree level b); entries of one node are contiguous in the
        // ascending list, namely those sharing u >> b.
        const int Ca = accept ? C : 0;
        const int CW = warpMax(Ca);
        if constexpr (kPool) {
            // The blocked tree's commit (below) with the two half-warps sharing the rounds of the accepted proposals.  While both
            // chains have chunks left each half-warp commits its own; then both commit the longer list, chunks 2k and 2k+1 in
            // one round, and add their deltas to the sums one after the other (lower chunk first), so that every sum receives
            // its deltas in ascending chunk order exactly as in the unpooled loop.
            constexpr unsigned gMask = (1u << W) - 1u;
            const int nAMine = (Ca + W - 1) / W, nAOth = __shfl_xor_sync(kFull, nAMine, W), CaOth = __shfl_xor_sync(kFull, Ca, W);
            const int nABoth = min(nAMine, nAOth);
            const bool aMineLong = nAMine > nAOth;
            const int roundsA = nABoth + ((max(nAMine, nAOth) - nABoth + 1) >> 1);
            double2 *colRec0 = S.colRec + (size_t)chainBase0 * nCols;
            const double2 *scratch0 = S.scratch + (size_t)chainBase0 * P.sCap;
            double *tier10 = S.tier1 + (size_t)chainBase0 * P.t1Len, *tier20 = S.tier2 + (size_t)chainBase0 * P.t2Len;
            double2 scN = make_double2(0.0, 0.0);   // the staged {DE', signed leaf} of the coming round, one round ahead
            auto fetchStaged = [&](int r) {
                int grp, c;
                scN = make_double2(0.0, 0.0);
                if (r < roundsA && poolAssign(r, nAMine, nAOth, nABoth, aMineLong, grp, c)) {
                    const int l = c * W + gl;
                    if (l < (grp == g ? Ca : CaOth)) scN = (scratch0 + (size_t)grp * P.sCap)[l];
                }
            };
            fetchStaged(0);
            #pragma unroll 1
            for (int r = 0; r < roundsA; ++r) {
                int grp = g, c = 0;
                const bool has = poolAssign(r, nAMine, nAOth, nABoth, aMineLong, grp, c);
                const int l = c * W + gl;
                const bool valid = has && l < (grp == g ? Ca : CaOth);
                const int *cx = ofGroup(sCtx, grp);
                unsigned u = 0x7fffffffu;
                if (valid) {
                    const uint32_t *recX = ofGroup(sRec, grp);
                    const uint32_t *recGX = P.rec + (size_t)(unsigned)cx[6] * 4;
                    const int so = cx[1], wd = cx[3];
                    auto rw = [&](int i) -> uint32_t { return (kWinWhole || i < kWinWords) ? recX[i] : __ldg(recGX + i); };
                    u = wd ? rw(so + 2 * l) : (unsigned)cx[2] + (rw(so + l) & 0xffffu);
                }
                const double dl = valid ? ofGroup(sDl, grp)[l] : 0.0;
                const double2 sc = scN;
                fetchStaged(r + 1);
                if (valid) (colRec0 + (size_t)grp * nCols)[u] = sc;
                double pin = dl;
#pragma unroll
                for (int o = 1; o < W; o <<= 1) {
                    const double t = __shfl_up_sync(kFull, pin, o, W);
                    if (gl >= o) pin = __dadd_rn(pin, t);
                }
                const double pexRaw = __shfl_up_sync(kFull, pin, 1, W);
                const double pex = gl ? pexRaw : 0.0;
                const unsigned prevU = __shfl_up_sync(kFull, u, 1, W), nextU = __shfl_down_sync(kFull, u, 1, W);
                double seg[3];
                bool emit[3];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int lv = 5 + 5 * q;
                    const bool head = gl == 0 || (prevU >> lv) != (u >> lv);
                    const bool tail = gl == W - 1 || (nextU >> lv) != (u >> lv);
                    const unsigned hm = (__ballot_sync(kFull, head) >> gBase) & gMask & ((2u << gl) - 1u);
                    const int start = 31 - __clz(hm);                          // first slot of this slot's run (bit 0 is always set)
                    seg[q] = __dsub_rn(pin, __shfl_sync(kFull, pex, start, W));
                    emit[q] = valid && tail;
                }
                // both half-warps on one chain (r >= nABoth): the lower chunk (its owner's) adds first
                const bool shared = r >= nABoth;
                double *sTopX = ofGroup(sTop, grp);
#pragma unroll 1
                for (int ph = 0; ph < (shared ? 2 : 1); ++ph) {
                    if (!shared || ph == (grp == g ? 0 : 1)) {
                        if (emit[0]) atomicAdd(tier10 + (size_t)grp * P.t1Len + (u >> 5), seg[0]);
                        if (emit[1]) atomicAdd(tier20 + (size_t)grp * P.t2Len + (u >> 10), seg[1]);
                        if (emit[2]) sTopX[u >> 15] = __dadd_rn(sTopX[u >> 15], seg[2]);
                    }
                    __syncwarp();   // a run may continue in the next chunk: its two updates of one sum stay ordered
                }
            }
        } else if constexpr (BLK) {
            // Blocked tree: currentDE[u] and the signed leaf in one 16-byte store; then every sum of 32 / 1024 / 32768 leaves
            // that covers changed columns receives the sum of their leaf changes.  The columns are ascending, so the entries
            // under one sum are consecutive slots: one inclusive prefix sum over the W slots of a round gives every such run
            // as a difference of two prefix values.  The run's last slot adds it to the sum (fire-and-forget reductions for
            // the two global tiers: nothing waits on the old value; the top tier is in shared memory).
            constexpr unsigned gMask = W == 32 ? 0xffffffffu : ((1u << W) - 1u);
            const int gShiftLane = gBase;
            double2 scN = make_double2(0.0, 0.0);   // the staged {DE', signed leaf} of the coming round, one round ahead
            if (gl < Ca) scN = 0 < scrChunks ? sScr[gl] : scratch[gl];
            #pragma unroll 1
            for (int lb = 0; lb < CW; lb += W) {
                const int l = lb + gl;
                const bool valid = l < Ca;
                const unsigned u = valid ? (unsigned)slotCol(l, wide ? 0u : slotWord(l)) : 0x7fffffffu;
                const double dl = valid ? sDl[l] : 0.0;
                const double2 sc = scN;
                if (l + W < Ca) scN = lb + W < scrChunks * W ? sScr[l + W] : scratch[l + W];
                if (valid) ch.colRec[u] = sc;
                double pin = dl;
#pragma unroll
                for (int o = 1; o < W; o <<= 1) {
                    const double t = __shfl_up_sync(kFull, pin, o, W);
                    if (gl >= o) pin = __dadd_rn(pin, t);
                }
                const double pexRaw = __shfl_up_sync(kFull, pin, 1, W);
                const double pex = gl ? pexRaw : 0.0;
                const unsigned prevU = __shfl_up_sync(kFull, u, 1, W), nextU = __shfl_down_sync(kFull, u, 1, W);
#pragma unroll
                for (int lv = 5; lv <= 15; lv += 5) {
                    const bool head = gl == 0 || (prevU >> lv) != (u >> lv);
                    const bool tail = gl == W - 1 || (nextU >> lv) != (u >> lv);
                    const unsigned hm = (__ballot_sync(kFull, head) >> gShiftLane) & gMask & ((2u << gl) - 1u);
                    const int start = 31 - __clz(hm);                          // first slot of this slot's run (bit 0 is always set)
                    const double basev = __shfl_sync(kFull, pex, start, W);
                    if (valid && tail) {
                        const double seg = __dsub_rn(pin, basev);
                        if (lv == 5) atomicAdd(&ch.tier1[u >> 5], seg);
                        else if (lv == 10) {
                            if (t2s) sT2[u >> 10] = __dadd_rn(sT2[u >> 10], seg); else atomicAdd(&ch.tier2[u >> 10], seg);
                        }
                        else sTop[u >> 15] = __dadd_rn(sTop[u >> 15], seg);
                    }
                }
                __syncwarp();   // a run may continue in the next round: its two updates of one sum stay ordered
            }
        } else {
            // Tasks are binned by tree level -- [10, root] (tier2 nodes), [5, 10) (tier1), [0, 5) (in-record nodes) -- because
            // a node's run length grows with its level: rounds then hold tasks of similar length.  The three per-lane
            // counts share one prefix scan (10 bits each).
            int nTA = 0, nTB = 0, nTC = 0;
            const int baseB = P.tCapA, baseC = P.tCapA + P.tCapB;
            #pragma unroll 1
            for (int lb = 0; lb < CW; lb += W) {
                const int l = lb + gl;
                unsigned taskMask = 0;
                if (l < Ca) {
                    const int u = sU[l];
                    const unsigned prev = l ? (unsigned)sU[l - 1] : 0u;
                    ch.colRec[u].x = scratch[l].x;
                    const int hbit = l ? (31 - __clz(prev ^ (unsigned)u)) : 31;
                    const unsigned low = hbit >= 31 ? 0xffffffffu : ((2u << hbit) - 1u);
                    const unsigned lowbit = (unsigned)u & (0u - (unsigned)u);
                    taskMask = (unsigned)u & low & ~(lowbit | (lowbit - 1u));                    // ancestors first reached here
                    if (l == 0 && u != 0) taskMask |= 0x80000000u;                               // the root
                    taskMask |= u ? lowbit : 0x80000000u;                                        // the column's own node
                }
                const int cnt = __popc(taskMask & 0xfffffc00u) | (__popc(taskMask & 0x000003e0u) << 10) | (__popc(taskMask & 0x0000001fu) << 20);
                const int incl = groupInclusiveScan<W>(cnt, gl);
                const int excl = incl - cnt, tot = __shfl_sync(kFull, incl, W - 1, W);
                int posA = nTA + (excl & 1023), posB = baseB + nTB + ((excl >> 10) & 1023), posC = baseC + nTC + (excl >> 20);
    #pragma unroll 1
                while (taskMask) {
                    const int b = __ffs(taskMask) - 1;
                    const int pos = b >= 10 ? posA++ : (b >= 5 ? posB++ : posC++);
                    sTask[pos] = (uint16_t)((l << 5) | b);
                    taskMask &= taskMask - 1;
                }
                nTA += tot & 1023;
                nTB += (tot >> 10) & 1023;
                nTC += tot >> 20;
            }
            const int nTasks = nTA + nTB + nTC;
            __syncwarp();
            {
                // rounds of W tasks per chain: fetch, ordered sum over the run, store
                const int roundsW = warpMax((nTasks + W - 1) / W);
                // task of the coming round, fetched one round ahead so its node load overlaps the current ordered sum
                int qN = 0, endN = 0;
                bool haveN = false;
                double vN = 0.0;
                double *npN = nullptr;
                auto fetch = [&](int t) {
                    haveN = t < nTasks;
                    qN = endN = 0;
                    if (haveN) {
                        const int tk = sTask[t < nTA ? t : (t < nTA + nTB ? baseB + t - nTA : baseC + t - nTA - nTB)];
                        const int l0 = tk >> 5, b = tk & 31;
                        const unsigned u0 = (unsigned)sU[l0];
                        const unsigned pre = u0 >> b;
                        const unsigned a = b == 31 ? 0u : (pre << b);
                        npN = ch.node((int)a);
                        const bool own = a == u0;
                        vN = own ? scratch[l0].y : *npN;
                        qN = own ? l0 + 1 : l0;
                        // the run ends at the first entry outside the node's index range [a, a + 2^b): binary search in the
                        // ascending list (unsigned compare makes the root's key 2^31 exceed every column)
                        const unsigned key = (pre + 1u) << b;
                        int lo = qN, hi = Ca;
                        #pragma unroll 1
                        while (lo < hi) {
                            const int mid = (lo + hi) >> 1;
                            if ((unsigned)sU[mid] < key) lo = mid + 1; else hi = mid;
                        }
                        endN = lo;
                    }
                };
                fetch(gl);
                #pragma unroll 1
                for (int rd = 0; rd < roundsW; ++rd) {
                    int q = qN;
                    const int end = endN;
                    const bool have = haveN;
                    double v = vN;
                    double *np = npN;
                    fetch((rd + 1) * W + gl);
                    #pragma unroll 1
                    while (__any_sync(kFull, q < end)) {
                        if (q < end) {
                            v = __dadd_rn(v, sDl[q]);
                            ++q;
                        }
                    }
                    if (have) *np = v;
                }
            }
        }
        __syncwarp();
#if ABM_SMP_SMEM
        if (kSample) {
            const int djE = sCtx[3];
            if (djE != 0) {   // one accepted proposal in ~T: the header in the record window still is this step's
                const uint32_t hb0 = sRec[0], hb1 = sRec[1];
                const int off = 4 + 2 * (int)(hb0 & 0xff) + (int)((hb0 >> 8) & 0xff) + (int)((hb0 >> 16) << ((hb1 >> 30) & 1)) + (int)((hb1 >> 8) & 0x3fff);
                const int32_t *fx = reinterpret_cast<const int32_t *>(P.rec + (size_t)(unsigned)sCtx[2] * 4) + off;
                if (ABM_SYN_LAZY) foldSynopsis(nSamp);
                const int dSyn = endStateEffects(fx, djE, P.nSynopsis, S.endRec + (size_t)chainC * P.S, nSamp, gl, W);
                if (gl < P.nSynopsis) syn += dSyn;
            }
        }
#elif ABM_END_DEFERRED
        if (kSample && fxN != 0) {
            const int sgn = fxN < 0 ? -1 : 1;
            if (gl < abs(fxN))
                atomicAdd(reinterpret_cast<unsigned long long *>(&S.endRec[(size_t)chainC * P.S + (fxK & 0xffffff)].pend),
                          (unsigned long long)((long long)(sgn * (fxK >> 24)) * (long long)nSamp));
            if (ABM_SYN_LAZY) foldSynopsis(nSamp);
            if (gl < P.nSynopsis) syn += sgn * (int)(int8_t)(fxS >> (8 * (gl & 3)));
            fxN = 0;
        }
#endif
        if (kSample) {
            // the do-while at SparseBasisSampler.h:247 hands out the state whenever it is feasible after a step
            if (act && infeas == 0) {
                if (gl < P.nSynopsis) {   // MeanAndVariance::operator(), diagnostics/MeanAndVariance.h:35-45 (exact in int64)
                    if (!ABM_SYN_LAZY) {
                        sSyn[gl] += syn;
                        sSyn[8 + gl] += (long long)syn * syn;
                    }
#if !ABM_SMP_SMEM
                    if (serLeft == 0) recordSeries(R.series, R.seriesThin, R.seriesCap, P.nSynopsis, chainC, S.nSamples[chainC] + nSamp, gl, syn);
#endif
                }
#if ABM_SMP_SMEM
                if (R.series) {   // the countdown to the next sample of the thinned series
                    const unsigned own = (W == 32 ? 0xffffffffu : ((1u << W) - 1u)) << gBase;
                    const int left = sCtx[1];
                    if (left == 0 && gl < P.nSynopsis)
                        recordSeries(R.series, R.seriesThin, R.seriesCap, P.nSynopsis, chainC, S.nSamples[chainC] + nSamp, gl, syn);
                    __syncwarp(own);
                    if (gl == 0) sCtx[1] = left == 0 ? R.seriesThin - 1 : left - 1;
                }
#else
                serLeft = serLeft == 0 ? R.seriesThin - 1 : serLeft - 1;
#endif
#if ABM_SMP_SMEM == 2
                {
                    const unsigned own = (W == 32 ? 0xffffffffu : ((1u << W) - 1u)) << gBase;
                    __syncwarp(own);
                    if (gl == 0) sCtx[0] += 1;
                    __syncwarp(own);
                }
#else
                ++nSamp;
#endif
            }
        }
        if (act) ++itersDone;
    }

    __syncwarp();
    for (int k = gl; k < nTop; k += W)
        if (chainValid) {
            if (BLK) tier3[k] = sTop[k]; else ch.tier2[(k << P.topShift) >> 10] = sTop[k];
        }
    if (t2s && chainValid)
        for (int k = gl; k < P.t2Len; k += W) ch.tier2[k] = sT2[k];
    if (kSample && ABM_SYN_LAZY) foldSynopsis(nSamp);
    if (kSample && chainValid) {   // the lazy accumulators are settled by sampleEndKernel
        if (gl < P.nSynopsis) {
            S.synSum[(size_t)chain * P.nSynopsis + gl] += sSyn[gl];
            S.synSumSq[(size_t)chain * P.nSynopsis + gl] += sSyn[8 + gl];
            S.synNow[(size_t)chain * P.nSynopsis + gl] = syn;
        }
        if (gl == 0) S.launchSamples[chain] = nSamp;
    }
    if (chainValid && gl == 0) {
        S.logImp[chain] = logImp;
        S.infeas[chain] = infeas;
        S.iter[chain] = sIter0[0] + (unsigned long long)itersDone;
        if (kMC) {
#pragma unroll
            for (int q = 0; q < 8; ++q) S.counters[(size_t)chain * 8 + q] += sCnt[q];
        } else if (R.itersOut) {
            R.itersOut[chain] += itersDone;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Construction kernels: SparseBasisSampler.h:144-178.

// delta_j = -1 iff the initial trajectory has the column's variable set (:147)
__global__ void initDeltaKernel(DevProblem P, DevChains S, const int8_t *__restrict__ initEvents, int nInit) {
    const int chain = blockIdx.y;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < P.nCols; j += gridDim.x * blockDim.x) {
        const int ev = P.colEvent[j];
        const bool on = nInit > ev && initEvents[(size_t)chain * nInit + ev] > 0;
        S.delta[(size_t)chain * P.nCols + j] = on ? -1 : 1;
    }
}

// X_i = F_i + sum_{j: delta_j == -1} M_ij, and the total infeasibility (:161-169)
__global__ void initRowsKernel(DevProblem P, DevChains S) {
    const int chain = blockIdx.y;
    const int8_t *delta = S.delta + (size_t)chain * P.nCols;
    int8_t *Xb = S.Xb + (size_t)chain * P.xStride;
    int inf = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.nRows; i += gridDim.x * blockDim.x) {
        int x = P.rowF[i];
        for (int k = P.rowPtr[i]; k < P.rowPtr[i + 1]; ++k)
            if (delta[P.rowCol[k]] == -1) x += P.rowVal[k];
        const int4 rp = P.rowParam[i];
        Xb[rp.x] = (int8_t)x;
        inf += infeasOf(x, rp.y, rp.z);
    }
    inf = warpSumInt(inf);
    if ((threadIdx.x & 31) == 0 && inf) atomicAdd(&S.infeas[chain], inf);
}

// currentDE_j (:172-178), and the sum tree after `set(j, 1.0)` for every j (:154): node i holds the
// number of leaves below it, an exact integer.
__global__ void initColsKernel(DevProblem P, DevChains S) {
    const int chain = blockIdx.y;
    const int8_t *delta = S.delta + (size_t)chain * P.nCols;
    const int8_t *Xb = S.Xb + (size_t)chain * P.xStride;
    Chain ch;
    ch.colRec = S.colRec + (size_t)chain * P.nCols;
    ch.tier1 = S.tier1 + (size_t)chain * P.t1Len;
    ch.tier2 = S.tier2 + (size_t)chain * P.t2Len;
    ch.top = ch.tier2;
    ch.topShift = 31;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < P.nCols; j += gridDim.x * blockDim.x) {
        const int dj = delta[j];
        double de = 0.0;
        for (int k = P.colPtr[j]; k < P.colPtr[j + 1]; ++k) {
            const int4 rp = P.rowParam[P.colRow[k]];
            const int x = Xb[rp.x];
            de = __dadd_rn(de, __dsub_rn(energyOf(x + dj * P.colVal[k], rp.y, rp.z, rp.w, P.facVal, P.kappa),
                                         energyOf(x, rp.y, rp.z, rp.w, P.facVal, P.kappa)));
        }
        ch.colRec[j].x = de;
        if (P.blockedTree) {   // every leaf 1.0; block sums = number of columns in the block
            ch.colRec[j].y = dj < 0 ? -1.0 : 1.0;   // leaf 1.0 with delta_j in the sign bit
            if ((j & 31) == 0) ch.tier1[j >> 5] = (double)(min(j + 32, P.nCols) - j);
            if ((j & 1023) == 0) ch.tier2[j >> 10] = (double)(min(j + 1024, P.nCols) - j);
            if ((j & 32767) == 0) S.tier3[(size_t)chain * 32 + (j >> 15)] = (double)(min(j + 32768, P.nCols) - j);
            continue;
        }
        const int span = j ? (j & -j) : P.nCols;
        *ch.node(j) = (double)(min(j + span, P.nCols) - j);
    }
}

// Per-chain initial trajectories drawn from the prior, on the device: Forecast::nextSample(startState, T, isFermionic =
// false) (Forecast.h:102-150) with the Bernoulli start state (BernoulliStartState.h:36-47): every agent state is occupied
// with its start probability (or, startKind 1, by a Poisson number of agents: PoissonStartState.h:38-44); then, timestep by timestep, every agent draws an act from PredPreyAgent::timestep given
// the opposite-species occupation of its four neighbours (PredPreyAgent.h:124-157) and its consequences populate the next
// state (:98-115).  One CTA per chain; uniforms from Philox4x32-10 keyed by (seed, chain) with counter (state, time, agent).
// Only "event count > 0" matters to the sampler's constructor (SparseBasisSampler.h:147), so events are written as 0/1.
__global__ void priorSampleKernel(DevProblem P, int nChains, unsigned long long seed, unsigned long long firstChain,
                                  const double *__restrict__ actPmf /* [2][2][6] */, double pPredator, double pPrey,
                                  int startKind, int8_t *__restrict__ events /* [nChains][nEvents] */) {
    extern __shared__ int priorSmem[];
    int *cur = priorSmem, *nxt = priorSmem + P.S;
    const int chain = blockIdx.x;
    if (chain >= nChains) return;
    const unsigned long long chainId = firstChain + (unsigned long long)chain;
    const int G = P.G, GG = G * G;
    int8_t *ev = events + (size_t)chain * P.nEvents;
    for (int e = threadIdx.x; e < P.nEvents; e += blockDim.x) ev[e] = 0;
    for (int psi = threadIdx.x; psi < P.S; psi += blockDim.x) {
        uint32_t w[4];
        philox4x32_10((uint32_t)psi, 0xffffffffu, (uint32_t)chainId, (uint32_t)(chainId >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), w);
        const double u0 = canonical_from_words(w[0], w[1]), rate = psi < GG ? pPredator : pPrey;   // type 0 = PREDATOR (PredPreyAgent.h:27-30)
        int n0 = u0 < rate ? 1 : 0;
        if (startKind == 1) {   // PoissonStartState::nextSample (PoissonStartState.h:38-44): Poisson(rate) by inversion
            double pk = exp(-rate), cum = pk;
            n0 = 0;
            while (u0 >= cum && n0 < 64) { ++n0; pk *= rate / n0; cum += pk; }
        }
        cur[psi] = n0;
        nxt[psi] = 0;
    }
    __syncthreads();
    for (int t = 0; t < P.T; ++t) {
        for (int psi = threadIdx.x; psi < P.S; psi += blockDim.x) {
            const int nAgents = cur[psi];
            if (nAgents == 0) continue;
            const int type = psi / GG, cell = psi - type * GG, cy = cell / G, cx = cell - cy * G, tb = type * GG, ob = (1 - type) * GG;
            const int xl = cx == 0 ? G - 1 : cx - 1, xr = cx == G - 1 ? 0 : cx + 1, yu = cy == G - 1 ? 0 : cy + 1, yd = cy == 0 ? G - 1 : cy - 1;
            const bool nb = cur[ob + xr + G * cy] + cur[ob + xl + G * cy] + cur[ob + cx + G * yu] + cur[ob + cx + G * yd] >= 1;
            const double *pmf = actPmf + (type * 2 + (nb ? 1 : 0)) * kActs;
            for (int a = 0; a < nAgents; ++a) {
                uint32_t w[4];
                philox4x32_10((uint32_t)psi, (uint32_t)(t * 4096 + a), (uint32_t)chainId, (uint32_t)(chainId >> 32), (uint32_t)seed,
                              (uint32_t)(seed >> 32), w);
                const double u = canonical_from_words(w[0], w[1]);
                int act = kActs - 1;
                double cum = 0.0;
                for (int k = 0; k < kActs; ++k) {
                    cum += pmf[k];
                    if (u < cum) { act = k; break; }
                }
                ev[((size_t)t * P.S + psi) * kActs + act] = 1;
                // consequences (PredPreyAgent.h:98-115): DIE -> none, moves -> the neighbour cell, GIVEBIRTH -> self + right
                if (act == 1) atomicAdd(&nxt[tb + xl + G * cy], 1);
                else if (act == 2) atomicAdd(&nxt[tb + xr + G * cy], 1);
                else if (act == 3) atomicAdd(&nxt[tb + cx + G * yu], 1);
                else if (act == 4) atomicAdd(&nxt[tb + cx + G * yd], 1);
                else if (act == 5) { atomicAdd(&nxt[psi], 1); atomicAdd(&nxt[tb + xr + G * cy], 1); }
            }
        }
        __syncthreads();
        for (int psi = threadIdx.x; psi < P.S; psi += blockDim.x) { cur[psi] = nxt[psi]; nxt[psi] = 0; }
        __syncthreads();
    }
}

// importanceFunc->setState(X); currentLogImportance = getLogValue(X) (:183-184; TrajectoryImportance.h:116-140):
// factors are added in (t, psi) order, one warp per chain, 32 states per round.
__global__ void setStateKernel(DevProblem P, DevChains S) {
    const int lane = threadIdx.x & 31;
    const int chain = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (chain >= S.nChains) return;
    double total = 0.0;
    if (P.nStates > 0) {
        const unsigned long long *X8 = reinterpret_cast<const unsigned long long *>(S.Xb + (size_t)chain * P.xStride);
        const int G = P.G, GG = G * G;
        for (int sb = 0; sb < P.nStates; sb += 32) {
            const int sid = sb + lane;
            double v = 0.0;
            bool nz = false;
            if (sid < P.nStates) {
                const unsigned long long own = X8[sid];
                if (posMask(own)) {
                    if (P.agentTab) {
                        const StateInfo si = stateFromTable(sid, P.S, P.nStates, P.agentTab);
                        const unsigned long long nbm = posMask(X8[si.n0]) | posMask(X8[si.n1]) | posMask(X8[si.n2]) | posMask(X8[si.n3]);
                        v = factorLogP(own, nbm != 0ull, si.cls, P.impTab);
                    } else {
                        const int t = sid / P.S, psi = sid - t * P.S;
                        const int type = psi / GG, cell = psi - type * GG;
                        const int cy = cell / G, cx = cell - cy * G;
                        const int ob = t * P.S + (1 - type) * GG;
                        const unsigned long long nbm = posMask(X8[ob + (cx + G - 1) % G + G * cy]) | posMask(X8[ob + (cx + 1) % G + G * cy]) |
                                                       posMask(X8[ob + cx + G * ((cy + 1) % G)]) | posMask(X8[ob + cx + G * ((cy + G - 1) % G)]);
                        v = factorLogP(own, nbm != 0ull, type, P.impTab);
                    }
                    nz = true;
                }
            }
            unsigned bal = __ballot_sync(kFull, nz);
            while (bal) {
                const int src = __ffs(bal) - 1;
                total = __dadd_rn(total, __shfl_sync(kFull, v, src));
                bal &= bal - 1;
            }
        }
    }
    if (lane == 0) S.logImp[chain] = total;
}

// End-state occupancy ModelState(traj,T,T) (ModelState.h:27-36 via State::backwardOccupation, State.h:57-63, and
// PredPreyAgent::consequences, PredPreyAgent.h:98-115) of one agent state from the events of the last timestep (Xl): MOVELEFT
// from the right neighbour, MOVERIGHT from the left one, MOVEUP from below, MOVEDOWN from above, GIVEBIRTH from the state
// itself (the parent stays) and from the left neighbour (the child)
__device__ __forceinline__ int endOccupancy(const int8_t *__restrict__ Xl, int psi, int G, int gShift = -1) {
    const int GG = G * G;
    int type, cx, cy;
    if (gShift >= 0) { type = psi >> (2 * gShift); cy = (psi & (GG - 1)) >> gShift; cx = psi & (G - 1); }
    else { type = psi / GG; const int cell = psi - type * GG; cy = cell / G; cx = cell - cy * G; }
    const int tb = type * GG;
    const int left = tb + (cx == 0 ? G - 1 : cx - 1) + G * cy, right = tb + (cx == G - 1 ? 0 : cx + 1) + G * cy;
    const int up = tb + cx + G * (cy == G - 1 ? 0 : cy + 1), down = tb + cx + G * (cy == 0 ? G - 1 : cy - 1);
    // one aligned 64-bit word per agent state: act a is byte a
    const unsigned long long *X8 = reinterpret_cast<const unsigned long long *>(Xl);
    const unsigned long long wl = X8[left];
    auto act = [](unsigned long long w, int a) { return (int)(int8_t)(w >> (8 * a)); };
    return act(X8[right], 1) + act(wl, 2) + act(X8[down], 3) + act(X8[up], 4) + act(X8[psi], 5) + act(wl, 5);
}

// The same for an agent described by tables: the events of the last timestep listed for the state (DevProblem::endIn)
__device__ __forceinline__ int endOccupancyTab(const int8_t *__restrict__ Xl, int psi, const int32_t *__restrict__ ptr,
                                               const int32_t *__restrict__ list) {
    int occ = 0;
    for (int q = __ldg(ptr + psi); q < __ldg(ptr + psi + 1); ++q) occ += Xl[__ldg(list + q)];
    return occ;
}

// Before a sampling launch: the synopsis (FiguresForPaper.h:245-263) of every chain's current end state, from which
// stepKernel<MODE_SAMPLE> proceeds incrementally.  One warp per chain.
__global__ void sampleBeginKernel(DevProblem P, DevChains S) {
    const int lane = threadIdx.x & 31;
    const int chain = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (chain >= S.nChains) return;
    const int G = P.G, GG = G * G;
    const int8_t *Xl = S.Xb + (size_t)chain * P.xStride + (size_t)(P.T - 1) * P.S * kStateBytes;
    for (int k = 0; k < P.nSynopsis; ++k) {
        const int sz = G >> (k + 1), org = G - 2 * sz;   // sizes G/2, G/4, ...; origins 0, G/2, 3G/4, ...
        int part = 0;
        for (int c2 = lane; c2 < 2 * sz * sz; c2 += 32) {
            const int type = c2 / (sz * sz), cc = c2 - type * sz * sz;
            part += endOccupancy(Xl, type * GG + (org + cc % sz) + G * (org + cc / sz), G);
        }
        part = warpSumInt(part);
        if (lane == 0) S.synNow[(size_t)chain * P.nSynopsis + k] = part;
    }
}

// After a sampling launch of N samples: a state's occupancy summed over them is occ_end * N - sum over its changes of
// (change * index of the first sample that saw it); the second term was accumulated by the step kernel.
__global__ void sampleEndKernel(DevProblem P, DevChains S) {   // one CTA per chain
    const int chain = blockIdx.x;
    const long long nSamp = S.launchSamples[chain];
    const int8_t *Xl = S.Xb + (size_t)chain * P.xStride + (size_t)(P.T - 1) * P.S * kStateBytes;
    longlong2 *rec = reinterpret_cast<longlong2 *>(S.endRec + (size_t)chain * P.S);   // {acc, pend}
    for (int psi = threadIdx.x; psi < P.S; psi += blockDim.x) {
        longlong2 e = rec[psi];
        const int occ = P.endInPtr ? endOccupancyTab(Xl, psi, P.endInPtr, P.endIn) : endOccupancy(Xl, psi, P.G, P.gShift);
        e.x += (long long)occ * nSamp - e.y;
        e.y = 0;
        rec[psi] = e;
    }
    if (threadIdx.x == 0) S.nSamples[chain] += nSamp;
}

// Blocked sum tree, before every launch: the two upper levels are recomputed from the level below (sums of 32 entries,
// highest index first -- the order the selection consumes them in), so incremental rounding cannot accumulate there across
// launches.  tier1 (sums of 32 leaves) stays incremental: rebuilding it would read every leaf.  One warp per chain.
__global__ void refreshBlockedTopKernel(DevProblem P, DevChains S) {
    const int lane = threadIdx.x & 31;
    const int chain = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (chain >= S.nChains) return;
    const double *t1 = S.tier1 + (size_t)chain * P.t1Len;
    double *t2 = S.tier2 + (size_t)chain * P.t2Len, *t3 = S.tier3 + (size_t)chain * 32;
    for (int s2 = lane; s2 < P.t2Len; s2 += 32) {
        double v = 0.0;
        for (int k = min(32 * s2 + 31, P.t1Len - 1); k >= 32 * s2; --k) v = __dadd_rn(v, t1[k]);
        t2[s2] = v;
    }
    __syncwarp();
    const int nTop = ((P.nCols - 1) >> 15) + 1;
    if (lane < nTop) {
        double v = 0.0;
        for (int k = min(32 * lane + 31, P.t2Len - 1); k >= 32 * lane; --k) v = __dadd_rn(v, t2[k]);
        t3[lane] = v;
    }
}

// Blocked sum tree: the lowest sums (of 32 leaves) recomputed from the leaves, highest index first.  The step kernel keeps them
// incrementally (one rounded add per update); this exact rebuild runs every kTier1RebuildEvery launches and on request
// (abm_chains_rebuild_tree), so their rounding cannot accumulate over the life of a chain.  One thread per sum.
__global__ void rebuildTier1Kernel(DevProblem P, DevChains S) {
    const int chain = blockIdx.y;
    const double2 *cr = S.colRec + (size_t)chain * P.nCols;
    double *t1 = S.tier1 + (size_t)chain * P.t1Len;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < P.t1Len; b += gridDim.x * blockDim.x) {
        double v = 0.0;
        for (int k = min(32 * b + 31, P.nCols - 1); k >= 32 * b; --k) v = __dadd_rn(v, fabs(cr[k].y));
        t1[b] = v;
    }
}

// ensemble end-state sum: out[psi] += sum over a slice of chains of endRec[chain][psi].acc (out zeroed by the caller)
constexpr int kReduceSlice = 64;
__global__ void reduceEndStateKernel(DevProblem P, DevChains S) {
    const int psi = blockIdx.x * blockDim.x + threadIdx.x;
    if (psi >= P.S) return;
    const int c0 = blockIdx.y * kReduceSlice, c1 = min(c0 + kReduceSlice, S.nChains);
    long long acc = 0;
    for (int c = c0; c < c1; ++c) acc += S.endRec[(size_t)c * P.S + psi].acc;
    atomicAdd(reinterpret_cast<unsigned long long *>(S.endStateSum) + psi, (unsigned long long)acc);
}

// Variogram of the thinned synopsis series over all chains (diagnostics/ChainStats.h:66-79, V_t = mean_i (x_i - x_{i+t})^2 at
// strided lags), as exact integer sums: out[j][k] += sum over chains with a full window and positions i < nPoints - j of
// (x_i - x_{i+j})^2 for synopsis statistic k; the synopsis values are integers, so the sums are exact, order-independent and
// additive across GPUs.  out[nLags][8] int64, zeroed by the caller; chainsOut counts the contributing chains.
// grid (nLags, chain slices), one lag per block column.
constexpr int kVarioSlice = 32;
__global__ void __launch_bounds__(128) variogramKernel(DevChains S, const int32_t *__restrict__ series, int seriesCap, int nSeriesChains,
                                                       int nSyn, int nPoints, long long minSamples, long long *__restrict__ out,
                                                       long long *__restrict__ chainsOut) {
    const int j = blockIdx.x;
    const int c0 = blockIdx.y * kVarioSlice, c1 = min(c0 + kVarioSlice, nSeriesChains);
    long long acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int nc = 0;
    for (int c = c0; c < c1; ++c) {
        if (S.nSamples[c] < minSamples) continue;
        ++nc;
        const int32_t *x = series + (size_t)c * seriesCap * nSyn;
        for (int i = threadIdx.x; i + j < nPoints; i += blockDim.x) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < nSyn) {
                    const long long d = (long long)x[(size_t)i * nSyn + k] - (long long)x[(size_t)(i + j) * nSyn + k];
                    acc[k] += d * d;
                }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        long long v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        if ((threadIdx.x & 31) == 0 && k < nSyn && v) atomicAdd(reinterpret_cast<unsigned long long *>(out) + (size_t)j * 8 + k, (unsigned long long)v);
    }
    if (j == 0 && threadIdx.x == 0 && nc) atomicAdd(reinterpret_cast<unsigned long long *>(chainsOut), (unsigned long long)nc);
}

// Per-GPU additive moments of the ensemble, reduced on the device (the reference sums its
// threads' results in a serial loop over std::futures, FiguresForPaper.h:62-65).  Everything a multi-GPU SUM all-reduce needs
// to rebuild the marginals and W, B, var+, R-hat (diagnostics/MultiChainStats.h:141-181):
//   exact[0] chains with at least two samples, [1] feasible samples, [2..10) MCMCStatistics counters,
//        [10..18) sum over chains of the synopsis sums, [18..26) of the sums of squares, [26..26+S) end-state occupancy sums
//   real [0..8) sum over chains of the chain mean (MeanAndVariance::mean, MeanAndVariance.h:52-54), [8..16) of its square,
//        [16..24) of the chain variance (sampleVariance, :57-59)
// One CTA; every partial sum is combined in a fixed order (shuffle tree, then warps in order), so the doubles are reproducible.
constexpr int kMomExactHead = 26, kMomReal = 24;
__global__ void __launch_bounds__(256) momentsKernel(DevProblem P, DevChains S, long long *__restrict__ exact, double *__restrict__ real) {
    __shared__ long long shI[8][kMomExactHead];
    __shared__ double shD[8][kMomReal];
    long long vi[kMomExactHead];
    double vd[kMomReal];
#pragma unroll
    for (int k = 0; k < kMomExactHead; ++k) vi[k] = 0;
#pragma unroll
    for (int k = 0; k < kMomReal; ++k) vd[k] = 0.0;
    const int d = P.nSynopsis;
    for (int c = threadIdx.x; c < S.nChains; c += blockDim.x) {
        const long long n = S.nSamples[c];
        vi[1] += n;
#pragma unroll
        for (int q = 0; q < 8; ++q) vi[2 + q] += S.counters[(size_t)c * 8 + q];
        if (n > 1) vi[0] += 1;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (k < d) {
                const long long s = S.synSum[(size_t)c * d + k], q = S.synSumSq[(size_t)c * d + k];
                vi[10 + k] += s;
                vi[18 + k] += q;
                if (n > 1) {
                    const double sd = (double)s, qd = (double)q, inv = 1.0 / (double)n;
                    const double mean = sd * inv;
                    const double var = (qd - sd * sd * inv) * (1.0 / (double)(n - 1));
                    vd[k] += mean;
                    vd[8 + k] += mean * mean;
                    vd[16 + k] += var;
                }
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < kMomExactHead; ++k) vi[k] += __shfl_xor_sync(kFull, vi[k], o);
#pragma unroll
        for (int k = 0; k < kMomReal; ++k) vd[k] += __shfl_xor_sync(kFull, vd[k], o);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kMomExactHead; ++k) shI[warp][k] = vi[k];
#pragma unroll
        for (int k = 0; k < kMomReal; ++k) shD[warp][k] = vd[k];
    }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    if (threadIdx.x < kMomExactHead) {
        long long a = 0;
        for (int w = 0; w < nw; ++w) a += shI[w][threadIdx.x];
        exact[threadIdx.x] = a;
    }
    if (threadIdx.x >= 32 && threadIdx.x < 32 + kMomReal) {
        double a = 0.0;
        for (int w = 0; w < nw; ++w) a += shD[w][threadIdx.x - 32];
        real[threadIdx.x - 32] = a;
    }
}

}  // namespace abm
