// Device-side data layout of the B200 sampler (shared between the host builder and the kernels).
//
// Shared, read-only, replicated per GPU ("problem"):
//   * one packed RECORD per basis column j: the CSR column with each row's bounds and factor table, and the
//     pre-joined list of the columns u coupled to j through a shared row together with the rows they share.
//     Proposal::Proposal discovers that list at run time by walking rows[i] for every i in cols[j]
//     (SparseBasisSampler.h:297-313) and accumulates into a std::map; the join is static, so the host does it
//     once and the kernel streams one record with coalesced loads instead of chasing
//     colPtr -> rowIdx -> rowPtr -> colIdx and inserting into a tree.
// Per chain, mutable, chain-contiguous in HBM:
//   * Xb     int8   [xStride]   X (SparseBasisSampler.h:44); event rows padded to 8 bytes per agent
//                               state so one aligned 64-bit load yields all 6 act counts of a state
//   * colRec double2[nCols]     .x = currentDE[u] (:47), .y = sum-tree node u when (u & 31) != 0
//   * delta  int8   [nCols]     (:45)
//   * tier1  double [nCols/32]  sum-tree nodes u with (u & 31) == 0, (u & 1023) != 0
//   * tier2  double [nCols/1024] sum-tree nodes u with (u & 1023) == 0 (root = tier2[0])
//   The node VALUES and their update arithmetic are exactly MutableCategoricalArray's
//   (MutableCategoricalArray.h:9-24); only their addresses are permuted so that the 5 lowest levels of
//   any subtree, the next 5, and the top levels are each contiguous.
#pragma once
#include <cstdint>

namespace abm {

constexpr int kActs = 6;            // PredPreyAgentBase::actDomainSize(), PredPreyAgent.h:46
constexpr int kStateBytes = 8;      // bytes reserved per agent state in Xb (6 acts + 2 pad)

// Record of column j: 32-bit words, 16-byte aligned, contiguous, independent of the lanes per chain.  recIdx[j] = {offset in
// 16-byte units, bytes}.  The step kernel copies the head of a record into shared memory with ONE bulk asynchronous copy
// (cp.async.bulk, completion on an mbarrier) and parses it there.
//   w0: n | nSF << 8 | C << 16      rows in the column; stale factors; coupled columns incl. j itself
//   w1: nEv | nXtra << 8 | wide << 30 | touchesEnd << 31    leading rows that are events; words of extra pair lists;
//       wide: slots are two words {u, 0 | pair0 << 16 | pair1 << 24} (coupled ids spanning more than 16 bits: 64x64 grids)
//   w2: uBase                       lowest coupled column id;  w3: reserved
//   rows:   n x {xoff | M_ij << 24,  eBase | L8 << 16 | U8 << 24}     energy(i, x) = eTab[eBase + x]; bounds clamped to int8
//   staleF: nSF x state id, TrajectoryImportance::perturb insertion order, de-duplicated
//   slots:  C x {u - uBase | pair0 << 16 | pair1 << 24}, u ascending; pair = k | code(M_iu) << 5 (row k of the column),
//           0xff = none; pair0 == 0xfe: the column shares more than two rows, pair1 = word index of its list in xtra
//   xtra:   lists {count, pair codes ...}, one byte each, padded to words
//   end-state effects (only if touchesEnd): nEnd, 8 x int8 synopsis change, nEnd x (state | entry << 24)
constexpr int kPairNone = 0xff, kPairExt = 0xfe;
struct DevProblem {
    int nRows, nCols, nEvents, xStride;
    int treeDepth;          // D = bit length of nCols-1
    int t1Len, t2Len;
    int topShift, nTop;     // nodes whose index is a multiple of 2^topShift (<= 32 of them) are held in shared memory by the step kernel
    int evBytes;            // bytes of the padded event region of Xb = nStates * kStateBytes
    int G, S, T, nStates;   // PredPrey grid (0: none), agent states per timestep, timesteps, T * S (0: no importance function)
    int gShift;             // log2(G) when G is a power of two, else -1
    int blockedTree;        // 1: 32-ary blocked sum tree (colRec.y = leaf p_u), 0: the reference's binary sum tree
    int groupWidth;         // W
    int rowsCap, sCap, tCap;   // per-chain shared-memory capacities: rows, coupled columns, commit tasks
    int maxAbs, twoM;          // largest |M_ij| of the basis (1 on the paper's PredPrey problems), and twice that: the row
                               // moves tabulated per row of a proposal, -maxAbs..-1, 1..maxAbs
    int nodeCap, t2Cap;        // shared-memory doubles per chain: descent staging (reference tree: 32) / blocked tree's sums of 1024
                               // leaves when 7 resident CTAs per SM leave room for them (t2Len, else 0)
    int scrCap;                // blocked tree: leading coupled columns whose staged {DE', leaf} stay in shared memory (multiple of W)
    int tCapA, tCapB;          // task-list regions: levels >= 10, levels 5..9 (the rest of tCap: levels < 5)
    int nSynopsis;
    int fixLayout;             // 1: the compile-time shared-memory layout of stepKernel<..., FIX> applies (rowsCap / sCap are its constants)
    double kappa;
    const uint2 *recIdx;    // [nCols] {record offset in 16-byte units, record bytes}
    const uint32_t *rec;
    const double *eTab;     // energy tables of the row classes (step kernel)
    const double *facVal;   // de-duplicated factor tables (construction kernels)
    const double *impTab;   // [7] lgamma table, [2][2][6] log ratios (PredPrey), then at 31 + ((class * 2 + flag) << 6 | present acts) the
                            // factor log-probability of TrajectoryImportance::refresh
    const int32_t *agentTab;   // [S][5] {class, 4 influencer agent states (-1: unused)}: grids that are not a power of two and agents
                               // described by tables (abmcmc_b200.h); null on power-of-two PredPrey grids (geometry by shifts)
    const int32_t *endInPtr, *endIn;   // table agents: per agent state, the byte offsets (state * 8 + act) of the events of the previous
                                       // timestep that feed it (State::backwardOccupation, State.h:57-63); null for PredPrey
    // construction only
    const int32_t *colPtr, *colRow, *colVal;   // CSR by column (sampler row ids)
    const int32_t *rowPtr, *rowCol, *rowVal;   // CSR by row
    const int32_t *rowF;                       // F per sampler row
    const int4 *rowParam;                      // {xoff, L, U, tabBase} per sampler row
    const int32_t *colEvent;                   // tableau column of each basis column
};

// MODE_SAMPLE, per chain and agent state
struct EndRec {
    long long acc;           // sum over this chain's samples of the end-state occupancy (up to the last finished launch)
    long long pend;          // running launch: sum over the occupancy changes of (change * index of the first sample seeing it)
};

struct DevChains {
    int nChains;
    int8_t *Xb;
    double2 *colRec;
    int8_t *delta;
    double *tier1, *tier2;
    double *tier3;           // [nChains][32] blocked tree: sums of 32768 leaves
    double *logImp;          // currentLogImportance (:50)
    int32_t *infeas;         // currentInfeasibility (:49)
    unsigned long long *iter;   // Philox counter: iterations (init + MH) done so far
    long long *counters;     // [nChains][8]
    double2 *scratch;        // [nChains][sCap] per step: {DE', new own-node base} of every coupled column
    // statistics
    long long *endStateSum;  // [S]   ensemble sum
    long long *synSum, *synSumSq; // [nChains][nSynopsis]
    long long *nSamples;     // [nChains]
    EndRec *endRec;          // [nChains][S] end-state occupancy accumulators
    int32_t *synNow;         // [nChains][nSynopsis] synopsis of the current state (between sampling launches)
    long long *launchSamples; // [nChains] samples taken by the last sampling launch
};

}  // namespace abm
