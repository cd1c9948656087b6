// CUDA kernels of the B200 SparseBasisSampler (sm_100a).  One warp owns one Markov chain.
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
constexpr int kSCap = 64;    // distinct-column entries kept in shared memory per warp (rest: global scratch)
constexpr int kTCap = 128;   // commit tasks kept in shared memory per warp (rest: global scratch)
constexpr unsigned kFull = 0xffffffffu;

enum { MODE_MCMC = 0, MODE_INIT = 1 };

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
    long long nSamples;          // sample mode: feasible samples wanted per chain (0 = plain stepping)
};

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 and libstdc++'s generate_canonical<double,53> (bits/random.tcc:3349-3381): the same two
// constructions as oracle/sbs_oracle.c, so GPU and oracle consume identical uniform streams.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
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

// SparseBasisSampler.h:82
__device__ __forceinline__ double DEtoBasisProb(double DE) { return DE < 0.0 ? exp(DE) : 1.0; }

// ------------------------------------------------------------------------------------------------
// Per-warp view of one chain's HBM state.
struct Chain {
    int8_t *Xb;
    double2 *colRec;
    int8_t *delta;
    double *tier1, *tier2;
    __device__ __forceinline__ double *node(int i) const {
        if (i & 31) return &colRec[i].y;
        if (i & 1023) return &tier1[i >> 5];
        return &tier2[i >> 10];
    }
};

// Per-warp commit scratch: the first kSCap / kTCap entries live in shared memory, the rest in a
// per-chain global overflow area (only columns coupled to > kSCap others ever reach it).
struct Scratch {
    int *sU;
    double *sD;   // [3][kSCap]: DEnew, base, dlt
    uint16_t *sT;
    int *gU;
    double *gD;
    uint16_t *gT;
    __device__ __forceinline__ int &U(int l) const { return l < kSCap ? sU[l] : gU[l - kSCap]; }
    __device__ __forceinline__ double &DEn(int l) const { return l < kSCap ? sD[l] : gD[3 * (l - kSCap)]; }
    __device__ __forceinline__ double &Base(int l) const { return l < kSCap ? sD[kSCap + l] : gD[3 * (l - kSCap) + 1]; }
    __device__ __forceinline__ double &Dl(int l) const { return l < kSCap ? sD[2 * kSCap + l] : gD[3 * (l - kSCap) + 2]; }
    __device__ __forceinline__ uint16_t &T(int t) const { return t < kTCap ? sT[t] : gT[t - kTCap]; }
};

// MutableCategoricalArray::operator()(RNG&), MutableCategoricalArray.h:119-132, evaluated five tree
// levels per memory round trip: lane r fetches the node reached by the branch pattern r, then the same
// strict comparisons / subtractions as the reference are resolved with shuffles.  Every lane returns j.
__device__ __forceinline__ int treeDescend(const Chain &ch, int nCols, int depth, double target, int lane) {
    int idx = 0;
    int hi = depth - 1;
    while (hi >= 0) {
        int tierBase = hi >= 10 ? 10 : (hi >= 5 ? 5 : 0);
        int lo = max(tierBase, hi - 4);
        int nl = hi - lo + 1;
        int c = idx + (lane << lo);
        double v = 0.0;
        if (lane > 0 && lane < (1 << nl) && c < nCols) v = *ch.node(c);
        int pref = 0;
        for (int b = nl - 1; b >= 0; --b) {
            int rc = pref | (1 << b);
            double vv = __shfl_sync(kFull, v, rc);
            if (idx + (rc << lo) < nCols) {
                if (vv > target) pref = rc; else target = __dsub_rn(target, vv);
            }
        }
        idx += pref << lo;
        hi = lo - 1;
    }
    return idx;
}

// ------------------------------------------------------------------------------------------------
// PredPrey trajectory importance (TrajectoryImportance.h:90-113 on PredPreyAgent.h:124-182).
// An agent state's 6 act counts are one aligned 64-bit word of Xb.  posMask has bit 8a+7 set iff
// X[event(t,psi,a)] > 0; its popcount is State::fermionicBoundedForwardOccupation (State.h:66-74).
__device__ __forceinline__ unsigned long long posMask(unsigned long long w) {
    const unsigned long long m7 = 0x00007f7f7f7f7f7full, m8 = 0x0000808080808080ull;
    return (((w & m7) + m7) & ~w) & m8;
}
// newLogP of refresh(): lgamma(occ+1) if occ>1, then += log(pi/pibar) per act with X>0, act order
__device__ __forceinline__ double factorLogP(unsigned long long own, bool neighbour, int type,
                                             const double *__restrict__ impTab) {
    unsigned long long pm = posMask(own);
    int occ = __popcll(pm);
    double p = occ > 1 ? __ldg(impTab + occ) : 0.0;
    if (occ > 0) {
        const double *ratio = impTab + 7 + (type * 2 + (neighbour ? 1 : 0)) * kActs;
#pragma unroll
        for (int a = 0; a < kActs; ++a)
            if ((pm >> (8 * a + 7)) & 1ull) p = __dadd_rn(p, __ldg(ratio + a));
    }
    return p;
}

__device__ __forceinline__ int warpSumInt(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------
// One chain per warp.  MODE_MCMC: the loop body of nextSampleWithInfeasibleImportance
// (SparseBasisSampler.h:221-246).  MODE_INIT: findInitialFeasibleSolution (:254-263): same proposal,
// always applied, no importance, no statistics, until the chain is feasible.
template <int MODE>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 8) stepKernel(DevProblem P, DevChains S, RunParams R) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chain = blockIdx.x * kWarpsPerCta + warp;
    if (chain >= S.nChains) return;

    constexpr int kWarpSmem = kSCap * 4 + 3 * kSCap * 8 + kTCap * 2;
    unsigned char *ws = smemRaw + warp * kWarpSmem;
    Scratch sc;
    sc.sD = reinterpret_cast<double *>(ws);
    sc.sU = reinterpret_cast<int *>(ws + 3 * kSCap * 8);
    sc.sT = reinterpret_cast<uint16_t *>(ws + 3 * kSCap * 8 + kSCap * 4);
    const int scrCap = max(P.maxCoupled - kSCap, 0);
    sc.gU = S.scrU + (size_t)chain * scrCap;
    sc.gD = S.scrD + (size_t)chain * scrCap * 3;
    sc.gT = S.scrTask + (size_t)chain * 32 * 32;

    Chain ch;
    ch.Xb = S.Xb + (size_t)chain * P.xStride;
    ch.colRec = S.colRec + (size_t)chain * P.nCols;
    ch.delta = S.delta + (size_t)chain * P.nCols;
    ch.tier1 = S.tier1 + (size_t)chain * P.t1Len;
    ch.tier2 = S.tier2 + (size_t)chain * P.t2Len;

    double logImp = S.logImp[chain];
    int infeas = S.infeas[chain];
    unsigned long long iter = S.iter[chain];
    int cProp[4] = {0, 0, 0, 0}, cAcc[4] = {0, 0, 0, 0};
    const unsigned long long chainId = R.firstChain + (unsigned long long)chain;
    const double kappa = P.kappa;
    const double *__restrict__ facVal = P.facVal;

    long long s = 0;
    for (; s < R.nSteps; ++s) {
        if (MODE == MODE_INIT && infeas == 0) break;

        // ---- uniforms: one Philox block per iteration, or the injected stream
        double u1, u2 = 0.0;
        if (R.injU) {
            const double *up = R.injU + (size_t)chain * R.uStride + R.uOffset;
            if (MODE == MODE_INIT) u1 = up[s];
            else { u1 = up[2 * s]; u2 = up[2 * s + 1]; }
        } else {
            uint32_t w[4];
            philox4x32_10((uint32_t)iter, (uint32_t)(iter >> 32), (uint32_t)chainId, (uint32_t)(chainId >> 32),
                          (uint32_t)R.seed, (uint32_t)(R.seed >> 32), w);
            u1 = canonical_from_words(w[0], w[1]);
            u2 = canonical_from_words(w[2], w[3]);
        }

        // ---- Proposal::Proposal (:288-315)
        const double root = ch.tier2[0];
        int j = treeDescend(ch, P.nCols, P.treeDepth, __dmul_rn(u1, root), lane);   // :289
        if (R.injJ) j = R.injJ[(size_t)chain * R.nSteps + s];

        const int32_t *__restrict__ rec = P.rec + (size_t)P.recOff[j] * 4;
        const int4 hdr = __ldg(reinterpret_cast<const int4 *>(rec));
        const int n = hdr.x, nPairs = hdr.y, C = hdr.z, nSF = hdr.w;
        const int4 *__restrict__ recRows = reinterpret_cast<const int4 *>(rec + 4);
        const int32_t *__restrict__ pairU = rec + 4 + 4 * n;
        const int16_t *__restrict__ pairKM = reinterpret_cast<const int16_t *>(pairU + nPairs);
        const int32_t *__restrict__ staleF = pairU + nPairs + ((nPairs + 1) >> 1);
        const int dj = ch.delta[j];

        // rows of the proposed column: lane k holds row k
        int xoff = 0, L = 0, U = 0, tb = 0, x = 0, xn = 0, dinf = 0;
        double dE = 0.0;
        if (lane < n) {
            int4 r = __ldg(recRows + lane);
            xoff = r.x & 0xffffff;
            int m = r.x >> 24;
            L = r.y; U = r.z; tb = r.w;
            x = ch.Xb[xoff];
            xn = x + m * dj;                                                           // :299-300
            dE = __dsub_rn(energyOf(xn, L, U, tb, facVal, kappa), energyOf(x, L, U, tb, facVal, kappa));  // :302
            dinf = infeasOf(xn, L, U) - infeasOf(x, L, U);                              // :303
        }
        const int infN = infeas + warpSumInt(dinf);

        // coupled columns: lane p holds pair p = (u, k); runs of equal u are contiguous (sorted) and the
        // contributions of one u are added in row order, as the map accumulation at :305-312 does
        double changeInSum = 0.0;
        int nDistinct = 0;
        for (int base = 0; base < nPairs; base += 32) {
            const int p = base + lane;
            int u = -1, km = 0;
            if (p < nPairs) { u = __ldg(pairU + p); km = __ldg(pairKM + p); }
            const bool real = u >= 0;
            const int k = km & 0xff;
            const bool self = real && k == kSelfMarker;
            const int mu = km >> 8;
            int du = 0;
            if (real && !self) du = ch.delta[u];
            const int uprev = __shfl_up_sync(kFull, u, 1);
            const bool head = real && (lane == 0 || uprev != u);
            double2 cr = make_double2(0.0, 0.0);
            if (head) cr = ch.colRec[u];
            // row data of row k
            const int ks = k & 31;
            const int xk = __shfl_sync(kFull, x, ks), xnk = __shfl_sync(kFull, xn, ks);
            const int Lk = __shfl_sync(kFull, L, ks), Uk = __shfl_sync(kFull, U, ks), tbk = __shfl_sync(kFull, tb, ks);
            const double dEk = __shfl_sync(kFull, dE, ks);
            double cval = 0.0;
            if (real && !self) {
                const int sft = mu * du;                                                 // :308
                cval = __dsub_rn(__dsub_rn(energyOf(xnk + sft, Lk, Uk, tbk, facVal, kappa),
                                           energyOf(xk + sft, Lk, Uk, tbk, facVal, kappa)), dEk);   // :309-310
            }
            double acc = self ? -cr.x : __dadd_rn(cr.x, cval);                          // :294 / try_emplace + first +=
            for (int r = 1; r < 32; ++r) {
                const int tu = __shfl_down_sync(kFull, u, r);
                const double tv = __shfl_down_sync(kFull, cval, r);
                const bool more = head && (lane + r < 32) && tu == u;
                if (more) acc = __dadd_rn(acc, tv);
                if (!__any_sync(kFull, more)) break;
            }
            // per distinct column: new probability, old leaf value (get(u) = tree[u] - descendantSum(u),
            // MutableCategoricalArray.h:78, .cpp:30-39) and what set(u, p) will need (.cpp:10-17)
            double term = 0.0, newP = 0.0, ssum = 0.0, nodeOld = 0.0;
            if (head) {
                newP = DEtoBasisProb(acc);
                nodeOld = (u & 31) ? cr.y : *ch.node(u);
                double dsum = 0.0;
                ssum = newP;
                for (int off = 1; (off & u) == 0 && off < P.nCols; off <<= 1) {
                    const int d = u + off;
                    if (d < P.nCols) {
                        const double t = *ch.node(d);
                        dsum = __dadd_rn(dsum, t);
                        ssum = __dadd_rn(ssum, t);
                    }
                }
                term = __dsub_rn(newP, __dsub_rn(nodeOld, dsum));                        // :279-280
            }
            const unsigned hb = __ballot_sync(kFull, head);
            if (MODE == MODE_MCMC) {
                // changeInSum += ... in ascending column order (:278-281)
                unsigned rest = hb;
                while (rest) {
                    const int src = __ffs(rest) - 1;
                    changeInSum = __dadd_rn(changeInSum, __shfl_sync(kFull, term, src));
                    rest &= rest - 1;
                }
            }
            if (head) {
                const int l = nDistinct + __popc(hb & ((1u << lane) - 1u));
                sc.U(l) = u;
                sc.DEn(l) = acc;
                sc.Base(l) = ssum;
                sc.Dl(l) = __dsub_rn(ssum, nodeOld);                                     // delta of set(), .cpp:17
            }
            nDistinct += __popc(hb);
        }

        bool accept = true;
        double logImpN = logImp;
        if (MODE == MODE_MCMC) {
            // ---- importance: perturb + refresh (TrajectoryImportance.h:36-54, 90-113)
            if (P.G > 0 && nSF > 0) {
                const int G = P.G, GG = G * G;
                const unsigned long long *X8 = reinterpret_cast<const unsigned long long *>(ch.Xb);
                for (int fb = 0; fb < nSF; fb += 32) {
                    const int f = fb + lane;
                    const bool act = f < nSF;
                    int sid = 0, type = 0, ns[4] = {0, 0, 0, 0};
                    unsigned long long own = 0ull, nb[4] = {0ull, 0ull, 0ull, 0ull};
                    if (act) {
                        sid = __ldg(staleF + f);
                        const int t = sid / P.S, psi = sid - t * P.S;
                        type = psi / GG;
                        const int cell = psi - type * GG;
                        const int cy = cell / G, cx = cell - cy * G;
                        const int ob = t * P.S + (1 - type) * GG;   // PredPreyAgent.h:199-207: left, right, up, down
                        ns[0] = ob + (cx + G - 1) % G + G * cy;
                        ns[1] = ob + (cx + 1) % G + G * cy;
                        ns[2] = ob + cx + G * ((cy + 1) % G);
                        ns[3] = ob + cx + G * ((cy + G - 1) % G);
                        own = X8[sid];
#pragma unroll
                        for (int q = 0; q < 4; ++q) nb[q] = X8[ns[q]];
                    }
                    unsigned long long ownN = own, nbN[4] = {nb[0], nb[1], nb[2], nb[3]};
                    // apply the proposal's event changes (the X swap at :225-228) to the loaded words
                    for (int r = 0; r < n; ++r) {
                        const int xo = __shfl_sync(kFull, xoff, r);
                        const int xv = __shfl_sync(kFull, xn, r);
                        if (act && xo < P.evBytes) {
                            const int rs = xo >> 3, sh = (xo & 7) * 8;
                            const unsigned long long ins = (unsigned long long)(xv & 0xff) << sh, msk = ~(0xffull << sh);
                            if (rs == sid) ownN = (ownN & msk) | ins;
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (rs == ns[q]) nbN[q] = (nbN[q] & msk) | ins;
                        }
                    }
                    double diff = 0.0;
                    if (act) {
                        const bool fl = (posMask(nb[0]) | posMask(nb[1]) | posMask(nb[2]) | posMask(nb[3])) != 0ull;
                        const bool flN = (posMask(nbN[0]) | posMask(nbN[1]) | posMask(nbN[2]) | posMask(nbN[3])) != 0ull;
                        diff = __dsub_rn(factorLogP(ownN, flN, type, P.impTab), factorLogP(own, fl, type, P.impTab));  // :109
                    }
                    const int cnt = min(32, nSF - fb);
                    for (int q = 0; q < cnt; ++q) logImpN = __dadd_rn(logImpN, __shfl_sync(kFull, diff, q));
                }
            }
            // ---- calcRatioOfSums (:276-284) and the Metropolis test (:232-234)
            const double ratio = root / __dadd_rn(root, changeInSum);
            const double acceptance = __dmul_rn(exp(__dsub_rn(logImpN, logImp)), ratio);
            accept = u2 < acceptance;
            const int ci = (infeas == 0 ? 2 : 0) + (infN == 0 ? 1 : 0);                 // MCMCStatistics.cpp:12-17
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                cProp[q] += (q == ci);
                cAcc[q] += (q == ci && accept);
            }
        }
        if (R.outJ) {
            if (lane == 0) {
                const size_t o = (size_t)chain * R.nSteps + s;
                R.outJ[o] = j;
                if (R.outAcc) R.outAcc[o] = accept ? 1 : 0;
                if (R.outInf) R.outInf[o] = infN;
            }
        }

        if (accept) {
            // ---- applyProposal (:322-336)
            logImp = logImpN;
            infeas = infN;
            if (lane < n) ch.Xb[xoff] = (int8_t)xn;
            if (lane == 0) ch.delta[j] = (int8_t)(-dj);
            __syncwarp();
            // For every changed column u (ascending): currentDE[u] = DE'; tree.set(u, p').  All set() deltas
            // depend only on pre-commit node values (an update never touches nodes >= its own index before
            // it runs), so they were computed above; what remains is, per distinct node, the ordered sum of
            // the deltas of the changed columns below it (MutableCategoricalArray.cpp:18-27).
            for (int lb = 0; lb < C; lb += 32) {
                const int l = lb + lane;
                int u = 0;
                unsigned ancMask = 0;
                if (l < C) {
                    u = sc.U(l);
                    ch.colRec[u].x = sc.DEn(l);
                    // own node: base + deltas of later changed columns inside its subtree
                    const int b = u ? (__ffs(u) - 1) : 31;
                    double v = sc.Base(l);
                    for (int q = l + 1; q < C && (sc.U(q) >> b) == (u >> b); ++q) v = __dadd_rn(v, sc.Dl(q));
                    *ch.node(u) = v;
                    // ancestors this entry is the first contributor of
                    const unsigned prev = l ? (unsigned)sc.U(l - 1) : 0u;
                    const int hbit = l ? (31 - __clz(prev ^ (unsigned)u)) : 31;
                    const unsigned low = hbit >= 31 ? 0xffffffffu : ((2u << hbit) - 1u);
                    const unsigned lowbit = (unsigned)u & (0u - (unsigned)u);
                    ancMask = (unsigned)u & low & ~(lowbit | (lowbit - 1u));
                    if (l == 0 && u != 0) ancMask |= 0x80000000u;                        // the root
                }
                // compact (l, bit) tasks
                const int cnt = __popc(ancMask);
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(kFull, incl, o);
                    if (lane >= o) incl += t;
                }
                const int nTasks = __shfl_sync(kFull, incl, 31);
                int pos = incl - cnt;
                unsigned mm = ancMask;
                while (mm) {
                    const int b = __ffs(mm) - 1;
                    sc.T(pos++) = (uint16_t)((lane << 5) | b);
                    mm &= mm - 1;
                }
                __syncwarp();
                for (int t = lane; t < nTasks; t += 32) {
                    const int tk = sc.T(t);
                    const int l0 = lb + (tk >> 5), b = tk & 31;
                    const int u0 = sc.U(l0);
                    const int a = b == 31 ? 0 : ((u0 >> b) << b);
                    double *np = ch.node(a);
                    double v = *np;
                    for (int q = l0; q < C && (sc.U(q) >> b) == (u0 >> b); ++q) v = __dadd_rn(v, sc.Dl(q));
                    *np = v;
                }
                __syncwarp();
            }
        }
        ++iter;
    }

    if (lane == 0) {
        S.logImp[chain] = logImp;
        S.infeas[chain] = infeas;
        S.iter[chain] = iter;
        if (MODE == MODE_MCMC) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                S.counters[(size_t)chain * 8 + q] += cProp[q];
                S.counters[(size_t)chain * 8 + 4 + q] += cAcc[q];
            }
        } else if (R.itersOut) {
            R.itersOut[chain] += s;
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
        const int span = j ? (j & -j) : P.nCols;
        *ch.node(j) = (double)(min(j + span, P.nCols) - j);
    }
}

// importanceFunc->setState(X); currentLogImportance = getLogValue(X) (:183-184; TrajectoryImportance.h:116-140):
// factors are added in (t, psi) order, one warp per chain, 32 states per round.
__global__ void setStateKernel(DevProblem P, DevChains S) {
    const int lane = threadIdx.x & 31;
    const int chain = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (chain >= S.nChains) return;
    double total = 0.0;
    if (P.G > 0) {
        const unsigned long long *X8 = reinterpret_cast<const unsigned long long *>(S.Xb + (size_t)chain * P.xStride);
        const int G = P.G, GG = G * G;
        for (int sb = 0; sb < P.nStates; sb += 32) {
            const int sid = sb + lane;
            double v = 0.0;
            bool nz = false;
            if (sid < P.nStates) {
                const unsigned long long own = X8[sid];
                if (posMask(own)) {
                    const int t = sid / P.S, psi = sid - t * P.S;
                    const int type = psi / GG, cell = psi - type * GG;
                    const int cy = cell / G, cx = cell - cy * G;
                    const int ob = t * P.S + (1 - type) * GG;
                    const unsigned long long nbm = posMask(X8[ob + (cx + G - 1) % G + G * cy]) | posMask(X8[ob + (cx + 1) % G + G * cy]) |
                                                   posMask(X8[ob + cx + G * ((cy + 1) % G)]) | posMask(X8[ob + cx + G * ((cy + G - 1) % G)]);
                    v = factorLogP(own, nbm != 0ull, type, P.impTab);
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

}  // namespace abm
