// Device-side data layout of the B200 sampler (shared between the host builder and the kernels).
//
// Shared, read-only, replicated per GPU ("problem"):
//   * one packed RECORD per basis column j (the CSR column, the rows' bounds and factor tables, and the
//     pre-joined list of (coupled column u, row k) pairs that Proposal::Proposal discovers by walking
//     rows[i] for every i in cols[j], SparseBasisSampler.h:297-313).  The join is static, so it is done
//     once on the host; the kernel streams a record with coalesced loads instead of chasing
//     colPtr -> rowIdx -> rowPtr -> colIdx.
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
constexpr int kSelfMarker = 0xff;   // pair entry that stands for the proposed column itself

// record header (4 int32 words)
//   [0] n      rows in the column (<= 32)
//   [1] nPairs padded pair count (runs of equal u never straddle a 32-entry chunk; pad u = -1)
//   [2] C      distinct columns in the pair list, the proposed column included
//   [3] nSF    stale factors (TrajectoryImportance::perturb insertion order, de-duplicated)
// then n x {xoff | m<<24, L, U, tabBase}, nPairs x u, ceil(nPairs/2) words of int16 {k | m<<8}, nSF x sid
struct DevProblem {
    int nRows, nCols, nEvents, xStride;
    int treeDepth;          // D = bit length of nCols-1
    int t1Len, t2Len;
    int evBytes;            // bytes of the padded event region of Xb = nStates * kStateBytes
    int G, S, T, nStates;   // PredPrey geometry (G == 0: no importance)
    int maxCoupled;         // max C over columns
    int nSynopsis;
    double kappa;
    const uint32_t *recOff; // [nCols+1] record offsets in units of 4 words
    const int32_t *rec;
    const double *facVal;   // de-duplicated factor tables
    const double *impTab;   // [7] lgamma table, then [2][2][6] log ratios
    // construction only
    const int32_t *colPtr, *colRow, *colVal;   // CSR by column (sampler row ids)
    const int32_t *rowPtr, *rowCol, *rowVal;   // CSR by row
    const int32_t *rowF;                       // F per sampler row
    const int4 *rowParam;                      // {xoff, L, U, tabBase} per sampler row
    const int32_t *colEvent;                   // tableau column of each basis column
};

struct DevChains {
    int nChains;
    int8_t *Xb;
    double2 *colRec;
    int8_t *delta;
    double *tier1, *tier2;
    double *logImp;          // currentLogImportance (:50)
    int32_t *infeas;         // currentInfeasibility (:49)
    unsigned long long *iter;   // Philox counter: iterations (init + MH) done so far
    long long *counters;     // [nChains][8]
    // commit scratch overflow (entries beyond the shared-memory capacity), per chain
    int32_t *scrU;
    double *scrD;            // 3 doubles per entry: DEnew, base, dlt
    uint16_t *scrTask;
    // statistics
    long long *endStateSum;  // [S]   ensemble sum
    long long *synSum, *synSumSq; // [nChains][nSynopsis]
    long long *nSamples;     // [nChains]
    int32_t *endOcc;         // [nChains][S] current end-state occupancy
    long long *endAcc;       // [nChains][S] per-chain accumulated end-state occupancy
    long long *endStamp;     // [nChains][S] sample index of the last flush
};

}  // namespace abm
