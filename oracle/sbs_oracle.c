/* sbs_oracle.c -- CPU restatement (plain C) of the reference SparseBasisSampler hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product path (agentbasedmcmc_b200) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file bit-for-bit (integers,
 * decisions AND doubles) against golden step traces produced by the unmodified reference compiled
 * from /root/reference/src (oracle/ref/abm_ref.cpp, fixtures in tests/golden/).
 *
 * Every function cites the reference code it restates (paths relative to /root/reference/src).
 * Floating point evaluation order is the reference's; compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------- */
/* Random streams                                                                              */

/* std::mt19937 (Random.h:15).  Standard MT19937; seeding as std::mersenne_twister_engine::seed. */
typedef struct {
    uint32_t mt[624];
    int idx;
} Mt19937;

static void mt_seed(Mt19937 *g, uint32_t seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}

static uint32_t mt_next(Mt19937 *g) {
    if (g->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* libstdc++ generate_canonical<double,53> on a 32-bit URBG (bits/random.tcc:3349-3381): two words,
 * low word first, sum rounded once to double, divided by 2^64, clamped below 1.
 * Used by MutableCategoricalArray.cpp:8 / .h:122 and Random.h:21-23. */
static double canonical_from_words(uint32_t w0, uint32_t w1) {
    double sum = (double)w0 + (double)w1 * 4294967296.0;
    double r = sum / 18446744073709551616.0;
    if (r >= 1.0) r = nextafter(1.0, 0.0);
    return r;
}

/* Philox4x32-10 (Salmon et al. 2011), the counter-based stream of the GPU sampler. */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum { RNG_MT = 0, RNG_PHILOX = 1, RNG_INJECT = 2 };

/* ------------------------------------------------------------------------------------------- */
/* Problem: what SparseBasisSampler's constructor copies out of the tableau                     */

typedef struct {
    int nRows, nCols, nEvents;
    int *colPtr, *colIdx, *colVal; /* cols[j]: rows ascending  (SparseBasisSampler.h:148-153)   */
    int *rowPtr, *rowIdx, *rowVal; /* rows[i]: columns ascending                                 */
    int *L, *U, *F;                /* per sampler row (SparseBasisSampler.h:128-140)             */
    int *colEvent;                 /* tableau column (= event id) of each basis column           */
    int *facLo, *facOff;           /* per row: table domain start, offset into facVal            */
    double *facVal;
    double kappa;
    /* importance TrajectoryImportance<AGENT>; S==0 -> no importance.  The agent -- PredPreyAgent<G> (G>0, tables filled from
     * its geometry below) or any other agent type (orc_problem_set_agent) -- is four tables of its own functions:
     * class of a state, the states its timestep() reads (influencers), neighbours(), consequences(act)            */
    int G, T, S, A;
    double lgammaTab[8];
    int nCls, *agCls, *agInfl, *agNbPtr, *agNb, *agCqPtr, *agCq;
    double *ratio;                 /* [nCls][2][A] log(pmf/marginal) by class, "any influencer occupied", act       */
} OrcProblem;

static void agentTablesAlloc(OrcProblem *p, int nNb, int nCq) {
    int S = p->S, A = p->A;
    p->agCls = (int *)calloc((size_t)S, sizeof(int));
    p->agInfl = (int *)malloc(sizeof(int) * 4 * (size_t)S);
    p->agNbPtr = (int *)calloc((size_t)S + 1, sizeof(int));
    p->agNb = (int *)malloc(sizeof(int) * (size_t)(nNb + 1));
    p->agCqPtr = (int *)calloc((size_t)S * A + 1, sizeof(int));
    p->agCq = (int *)malloc(sizeof(int) * (size_t)(nCq + 1));
    p->ratio = (double *)calloc((size_t)p->nCls * 2 * A, sizeof(double));
    for (int k = 0; k < 4 * S; ++k) p->agInfl[k] = -1;
}

ORC_API OrcProblem *orc_problem_create(int nTabRows, int nNonBasic, const int32_t *tabL, const int32_t *tabU,
                                       const int32_t *tabF, const int32_t *nbCol, const int32_t *nbPtr,
                                       const int32_t *nbRow, const int32_t *nbVal, const int32_t *facId,
                                       const int32_t *utLo, const int32_t *utOff, const double *utVal, double kappa,
                                       int G, int T, const double *lgammaTab, const double *logRatio) {
    OrcProblem *p = (OrcProblem *)calloc(1, sizeof(OrcProblem));
    int *map = (int *)malloc(sizeof(int) * (size_t)nTabRows);
    int nRows = 0;
    for (int i = 0; i < nTabRows; ++i) map[i] = (tabL[i] == tabU[i]) ? -1 : nRows++; /* :129-139 */
    p->nRows = nRows;
    p->nCols = nNonBasic;
    p->kappa = kappa;
    p->L = (int *)malloc(sizeof(int) * nRows);
    p->U = (int *)malloc(sizeof(int) * nRows);
    p->F = (int *)malloc(sizeof(int) * nRows);
    p->facLo = (int *)malloc(sizeof(int) * nRows);
    p->facOff = (int *)malloc(sizeof(int) * nRows);
    for (int i = 0; i < nTabRows; ++i)
        if (map[i] >= 0) {
            int b = map[i];
            p->L[b] = tabL[i];
            p->U[b] = tabU[i];
            p->F[b] = tabF[i];
            p->facLo[b] = utLo[facId[b]];
            p->facOff[b] = utOff[facId[b]];
        }
    int nUT = 0;
    for (int b = 0; b < nRows; ++b)
        if (facId[b] + 1 > nUT) nUT = facId[b] + 1;
    int nVal = utOff[nUT];
    p->facVal = (double *)malloc(sizeof(double) * (size_t)nVal);
    memcpy(p->facVal, utVal, sizeof(double) * (size_t)nVal);
    int nnz = nbPtr[nNonBasic];
    p->colPtr = (int *)malloc(sizeof(int) * (nNonBasic + 1));
    p->colIdx = (int *)malloc(sizeof(int) * nnz);
    p->colVal = (int *)malloc(sizeof(int) * nnz);
    p->rowPtr = (int *)calloc((size_t)nRows + 1, sizeof(int));
    p->rowIdx = (int *)malloc(sizeof(int) * nnz);
    p->rowVal = (int *)malloc(sizeof(int) * nnz);
    p->colEvent = (int *)malloc(sizeof(int) * nNonBasic);
    for (int j = 0; j <= nNonBasic; ++j) p->colPtr[j] = nbPtr[j];
    for (int k = 0; k < nnz; ++k) {
        p->colIdx[k] = map[nbRow[k]];
        p->colVal[k] = nbVal[k];
        p->rowPtr[map[nbRow[k]] + 1]++;
    }
    for (int i = 0; i < nRows; ++i) p->rowPtr[i + 1] += p->rowPtr[i];
    int *fill = (int *)malloc(sizeof(int) * nRows);
    memcpy(fill, p->rowPtr, sizeof(int) * nRows);
    for (int j = 0; j < nNonBasic; ++j) { /* insert(basisi,basisj,Mij) in column order, :87-90 */
        p->colEvent[j] = nbCol[j];
        for (int k = p->colPtr[j]; k < p->colPtr[j + 1]; ++k) {
            int i = p->colIdx[k];
            p->rowIdx[fill[i]] = j;
            p->rowVal[fill[i]] = p->colVal[k];
            fill[i]++;
        }
    }
    free(fill);
    free(map);
    p->G = G;
    p->T = T;
    p->A = 6;
    p->S = 2 * G * G;
    p->nEvents = T * p->S * p->A;
    if (G > 0) {
        int GG = G * G, A = p->A;
        memcpy(p->lgammaTab, lgammaTab, sizeof(double) * 7);
        p->nCls = 2; /* PREDATOR, PREY */
        agentTablesAlloc(p, 4 * p->S, 7 * p->S);
        memcpy(p->ratio, logRatio, sizeof(double) * 24);
        int nNb = 0, nCq = 0;
        for (int psi = 0; psi < p->S; ++psi) {
            int x = psi % G, y = (psi / G) % G, type = psi / GG, other = (1 - type) * GG, own = type * GG;
            int xl = (x + G - 1) % G, xr = (x + 1) % G, yu = (y + 1) % G, yd = (y + G - 1) % G;
            p->agCls[psi] = type;
            /* PredPreyAgent.h:199-207 neighbours(): left, right, up, down of the other species; the same four cells are what
             * timestep() reads through surroundingCountOf (:86-91 with xLeft..yDown :65-68) */
            int nb4[4] = {other + xl + G * y, other + xr + G * y, other + x + G * yu, other + x + G * yd};
            for (int k = 0; k < 4; ++k) { p->agInfl[4 * psi + k] = nb4[k]; p->agNb[nNb++] = nb4[k]; }
            p->agNbPtr[psi + 1] = nNb;
            /* PredPreyAgent.h:98-115 consequences(): DIE nothing, the four moves, GIVEBIRTH parent stays + child to the right */
            for (int a = 0; a < A; ++a) {
                if (a == 1) p->agCq[nCq++] = own + xl + G * y;
                else if (a == 2) p->agCq[nCq++] = own + xr + G * y;
                else if (a == 3) p->agCq[nCq++] = own + x + G * yu;
                else if (a == 4) p->agCq[nCq++] = own + x + G * yd;
                else if (a == 5) { p->agCq[nCq++] = own + x + G * y; p->agCq[nCq++] = own + xr + G * y; }
                p->agCqPtr[psi * A + a + 1] = nCq;
            }
        }
    }
    return p;
}

/* An agent type other than PredPrey, by tables (include/abmcmc_b200/reference_flatten.hpp flattenAgentImportance fills them
 * from the agent's own timestep / marginalTimestep / neighbours / consequences); call on a problem created with G == 0. */
ORC_API void orc_problem_set_agent(OrcProblem *p, int S, int A, int T, int nCls, const int32_t *cls, const int32_t *infl,
                                   const int32_t *nbPtr, const int32_t *nb, const int32_t *cqPtr, const int32_t *cq,
                                   const double *lgammaTab, const double *logRatio) {
    p->G = 0; p->S = S; p->A = A; p->T = T; p->nCls = nCls;
    p->nEvents = T * S * A;
    memcpy(p->lgammaTab, lgammaTab, sizeof(double) * 7);
    agentTablesAlloc(p, nbPtr[S], cqPtr[S * A]);
    for (int k = 0; k < S; ++k) p->agCls[k] = cls[k];
    for (int k = 0; k < 4 * S; ++k) p->agInfl[k] = infl[k];
    for (int k = 0; k <= S; ++k) p->agNbPtr[k] = nbPtr[k];
    for (int k = 0; k < nbPtr[S]; ++k) p->agNb[k] = nb[k];
    for (int k = 0; k <= S * A; ++k) p->agCqPtr[k] = cqPtr[k];
    for (int k = 0; k < cqPtr[S * A]; ++k) p->agCq[k] = cq[k];
    memcpy(p->ratio, logRatio, sizeof(double) * (size_t)nCls * 2 * A);
}

ORC_API void orc_problem_destroy(OrcProblem *p) {
    if (!p) return;
    free(p->colPtr); free(p->colIdx); free(p->colVal); free(p->rowPtr); free(p->rowIdx); free(p->rowVal);
    free(p->L); free(p->U); free(p->F); free(p->colEvent); free(p->facLo); free(p->facOff); free(p->facVal);
    free(p->agCls); free(p->agInfl); free(p->agNbPtr); free(p->agNb); free(p->agCqPtr); free(p->agCq); free(p->ratio);
    free(p);
}

ORC_API int orc_problem_nrows(const OrcProblem *p) { return p->nRows; }
ORC_API int orc_problem_ncols(const OrcProblem *p) { return p->nCols; }
ORC_API int orc_problem_nnz(const OrcProblem *p) { return p->colPtr[p->nCols]; }

/* ------------------------------------------------------------------------------------------- */
/* Chain state                                                                                 */

typedef struct {
    const OrcProblem *p;
    int *X;           /* SparseBasisSampler.h:44 */
    int *delta;       /* :45 */
    double *tree;     /* MutableCategoricalArray.h:38 */
    int treeHighBit;  /* MutableCategoricalArray.h:39 */
    double *DE;       /* :47 currentDE */
    double *E;        /* :48 currentE */
    int infeas;       /* :49 */
    double logImp;    /* :50 */
    /* importance (TrajectoryImportance.h:17-26) */
    int *occ;         /* stateTrajectory[t][psi], flat by state id */
    double *facLogP;  /* stateLogProbs */
    double totalLogP;
    int *staleStates, nStaleStates, *staleFactors, nStaleFactors;
    unsigned char *isStaleState, *isStaleFactor;
    int *occUndo;
    double *facUndo, totalUndo;
    /* proposal scratch (Proposal, :56-67): ordered map emulated by slot[] + touched list */
    int *slot, *touched, nTouched;
    double *newDE;
    int *newX;
    /* counters (MCMCStatistics.h:14-15) widened to 64 bit */
    int64_t nProposals[2][2], nAccepted[2][2];
    /* rng */
    int rngKind;
    Mt19937 mt;
    uint32_t phKey[2];
    uint64_t phChain, phIter;
    int phDraw;
    double phU[2];
    const double *inj;
    long injPos, injLen;
    int lastJ; /* optional injected proposal */
    const int32_t *injJ;
    long injJPos;
} OrcChain;

static double rng_next(OrcChain *c) {
    if (c->rngKind == RNG_MT) {
        uint32_t w0 = mt_next(&c->mt), w1 = mt_next(&c->mt);
        return canonical_from_words(w0, w1);
    } else if (c->rngKind == RNG_PHILOX) {
        if (c->phDraw == 0) {
            uint32_t ctr[4] = {(uint32_t)c->phIter, (uint32_t)(c->phIter >> 32), (uint32_t)c->phChain,
                               (uint32_t)(c->phChain >> 32)};
            uint32_t out[4];
            philox4x32_10(ctr, c->phKey, out);
            c->phU[0] = canonical_from_words(out[0], out[1]);
            c->phU[1] = canonical_from_words(out[2], out[3]);
        }
        return c->phU[c->phDraw++];
    } else {
        if (c->injPos >= c->injLen) return 0.5;
        return c->inj[c->injPos++];
    }
}
static void rng_end_iteration(OrcChain *c) { /* Philox: one counter value per iteration (init or MH step) */
    if (c->rngKind == RNG_PHILOX) {
        c->phIter++;
        c->phDraw = 0;
    }
}

/* SparseBasisSampler.h:267-271 */
static int infeasibility(const OrcProblem *p, int i, int x) {
    if (x < p->L[i]) return p->L[i] - x;
    if (x > p->U[i]) return x - p->U[i];
    return 0;
}
static double factor(const OrcProblem *p, int i, int x) { return p->facVal[p->facOff[i] + (x - p->facLo[i])]; }
/* SparseBasisSampler.h:341-345 */
static double energy(const OrcProblem *p, int i, int x) {
    if (x < p->L[i]) return factor(p, i, p->L[i]) - p->kappa * (p->L[i] - x);
    if (x > p->U[i]) return factor(p, i, p->U[i]) - p->kappa * (x - p->U[i]);
    return factor(p, i, x);
}
/* SparseBasisSampler.h:82 */
static double DEtoBasisProb(double DE) { return DE < 0.0 ? exp(DE) : 1.0; }

/* MutableCategoricalArray.cpp:41-48 */
static int highestOneBit(int i) {
    i |= i >> 1; i |= i >> 2; i |= i >> 4; i |= i >> 8; i |= i >> 16;
    return i - (int)((unsigned)i >> 1);
}
/* MutableCategoricalArray.cpp:30-39 */
static double tree_descendantSum(const OrcChain *c, int index) {
    int n = c->p->nCols, off = 1;
    double sum = 0.0;
    while ((off & index) == 0 && off < n) {
        int d = index + off;
        if (d < n) sum += c->tree[d];
        off <<= 1;
    }
    return sum;
}
/* MutableCategoricalArray.h:78 */
static double tree_get(const OrcChain *c, int index) { return c->tree[index] - tree_descendantSum(c, index); }
/* MutableCategoricalArray.cpp:10-28 */
static void tree_set(OrcChain *c, int index, double probability) {
    int n = c->p->nCols, off = 1;
    double sum = probability;
    while ((off & index) == 0 && off < n) {
        int d = index + off;
        if (d < n) sum += c->tree[d];
        off <<= 1;
    }
    double delta = sum - c->tree[index];
    int anc = index;
    c->tree[index] = sum;
    while (off < n) {
        anc ^= off;
        c->tree[anc] += delta;
        do { off <<= 1; } while ((anc & off) == 0 && off < n);
    }
}
/* MutableCategoricalArray.h:119-132 */
static int tree_sample(OrcChain *c, double u) {
    int n = c->p->nCols, index = 0;
    double target = u * c->tree[0];
    int off = c->treeHighBit;
    while (off != 0) {
        int child = index + off;
        if (child < n) {
            if (c->tree[child] > target) index += off; else target -= c->tree[child];
        }
        off >>= 1;
    }
    return index;
}

/* --- trajectory importance ------------------------------------------------------------------ */
/* what timestep() of (t, psi) reads of the other agents: is any of its influencer states occupied -- for PredPrey the
 * opposite-species occupation of the 4 neighbours, PredPreyAgent.h:86-91 with xLeft..yDown (:65-68); for CatMouse the cat at
 * the mouse's position, CatMouseAgent.cpp:14 */
static int anyInfluencerOccupied(const OrcChain *c, int t, int psi) {
    const OrcProblem *p = c->p;
    int cnt = 0;
    for (int k = 0; k < 4; ++k) {
        int q = p->agInfl[4 * psi + k];
        if (q >= 0) cnt += c->occ[t * p->S + q];
    }
    return cnt >= 1;
}
/* the newLogP computation shared by refresh (:96-108) and setState (:125-134) */
static double factorLogP(const OrcChain *c, int sid) {
    const OrcProblem *p = c->p;
    int t = sid / p->S, psi = sid % p->S;
    int occ = c->occ[sid];
    double logP = (occ > 1) ? p->lgammaTab[occ] : 0.0;
    if (occ > 0) {
        int flag = anyInfluencerOccupied(c, t, psi);
        const double *row = p->ratio + (size_t)(p->agCls[psi] * 2 + flag) * p->A;
        for (int a = 0; a < p->A; ++a)
            if (c->X[sid * p->A + a] > 0) logP += row[a]; /* 0.0 where pmf[a]==0 */
    }
    return logP;
}
/* State.h:66-74 */
static int fermionicOcc(const OrcChain *c, int sid) {
    int o = 0;
    for (int a = 0; a < c->p->A; ++a) o += c->X[sid * c->p->A + a] > 0 ? 1 : 0;
    return o;
}
static void imp_clearUndo(OrcChain *c) { /* TrajectoryImportance.h:144-149, IndexSet.h:37-40 */
    for (int k = 0; k < c->nStaleStates; ++k) c->isStaleState[c->staleStates[k]] = 0;
    for (int k = 0; k < c->nStaleFactors; ++k) c->isStaleFactor[c->staleFactors[k]] = 0;
    c->nStaleStates = c->nStaleFactors = 0;
}
static void imp_insertFactor(OrcChain *c, int sid) { /* IndexSet.h:28-35 + :46,49 */
    if (!c->isStaleFactor[sid]) {
        c->isStaleFactor[sid] = 1;
        c->facUndo[c->nStaleFactors] = c->facLogP[sid];
        c->staleFactors[c->nStaleFactors++] = sid;
    }
}
/* TrajectoryImportance.h:36-54 (the second "populate caches" loop :56-64 only appends unused
 * duplicates and is not restated) */
static void imp_perturb(OrcChain *c, const int *idx, int n) {
    const OrcProblem *p = c->p;
    if (p->S == 0) return;
    imp_clearUndo(c);
    c->totalUndo = c->totalLogP;
    for (int k = 0; k < n; ++k) {
        int i = idx[k];
        if (i < p->nEvents) {
            int sid = i / p->A; /* Event.h:15-16, State.h:35-37 */
            if (!c->isStaleState[sid]) {
                c->isStaleState[sid] = 1;
                c->occUndo[c->nStaleStates] = c->occ[sid];
                c->staleStates[c->nStaleStates++] = sid;
            }
            imp_insertFactor(c, sid);
            int t = sid / p->S, psi = sid % p->S;
            for (int q = p->agNbPtr[psi]; q < p->agNbPtr[psi + 1]; ++q) /* agent().neighbours(), :48-51 */
                imp_insertFactor(c, t * p->S + p->agNb[q]);
        }
    }
}
/* TrajectoryImportance.h:90-113 */
static void imp_refresh(OrcChain *c) {
    for (int k = 0; k < c->nStaleStates; ++k) c->occ[c->staleStates[k]] = fermionicOcc(c, c->staleStates[k]);
    for (int k = 0; k < c->nStaleFactors; ++k) {
        int sid = c->staleFactors[k];
        double newLogP = factorLogP(c, sid);
        c->totalLogP += newLogP - c->facLogP[sid];
        c->facLogP[sid] = newLogP;
    }
}
/* TrajectoryImportance.h:69-79 */
static void imp_undo(OrcChain *c) {
    if (c->p->S == 0) return;
    for (int k = 0; k < c->nStaleStates; ++k) c->occ[c->staleStates[k]] = c->occUndo[k];
    for (int k = 0; k < c->nStaleFactors; ++k) c->facLogP[c->staleFactors[k]] = c->facUndo[k];
    c->totalLogP = c->totalUndo;
    imp_clearUndo(c);
}
/* TrajectoryImportance.h:116-140 */
static void imp_setState(OrcChain *c) {
    const OrcProblem *p = c->p;
    if (p->S == 0) { c->totalLogP = 0.0; return; }
    c->totalLogP = 0.0;
    for (int t = 0; t < p->T; ++t) {
        for (int a = 0; a < p->S; ++a) c->occ[t * p->S + a] = fermionicOcc(c, t * p->S + a);
        for (int a = 0; a < p->S; ++a) {
            int sid = t * p->S + a;
            if (c->occ[sid] > 0) {
                c->facLogP[sid] = factorLogP(c, sid);
                c->totalLogP += c->facLogP[sid];
            }
        }
    }
    imp_clearUndo(c);
}
static double imp_getLogValue(OrcChain *c) { /* :82-85 */
    if (c->p->S == 0) return 0.0;
    imp_refresh(c);
    return c->totalLogP;
}

/* --- construction: SparseBasisSampler.h:107-178 (everything before findInitialFeasibleSolution) */
ORC_API OrcChain *orc_chain_create(const OrcProblem *p, const int8_t *initState, int nInit) {
    OrcChain *c = (OrcChain *)calloc(1, sizeof(OrcChain));
    int nR = p->nRows, nC = p->nCols;
    c->p = p;
    c->X = (int *)malloc(sizeof(int) * nR);
    c->E = (double *)malloc(sizeof(double) * nR);
    c->delta = (int *)malloc(sizeof(int) * nC);
    c->DE = (double *)calloc(nC, sizeof(double));
    c->tree = (double *)calloc(nC, sizeof(double));
    c->treeHighBit = highestOneBit(nC - 1);
    c->slot = (int *)malloc(sizeof(int) * nC);
    c->touched = (int *)malloc(sizeof(int) * nC);
    c->newDE = (double *)malloc(sizeof(double) * nC);
    c->newX = (int *)malloc(sizeof(int) * nR);
    for (int j = 0; j < nC; ++j) c->slot[j] = -1;
    int nStates = p->S > 0 ? p->T * p->S : 1;
    c->occ = (int *)calloc(nStates, sizeof(int));
    c->facLogP = (double *)calloc(nStates, sizeof(double));
    c->staleStates = (int *)malloc(sizeof(int) * nStates);
    c->staleFactors = (int *)malloc(sizeof(int) * nStates);
    c->isStaleState = (unsigned char *)calloc(nStates, 1);
    c->isStaleFactor = (unsigned char *)calloc(nStates, 1);
    c->occUndo = (int *)malloc(sizeof(int) * nStates);
    c->facUndo = (double *)malloc(sizeof(double) * nStates);
    for (int i = 0; i < nR; ++i) c->X[i] = p->F[i]; /* :138 */
    for (int j = 0; j < nC; ++j) {
        int ev = p->colEvent[j];
        c->delta[j] = (nInit > ev && initState[ev] > 0) ? -1 : 1; /* :147 */
        tree_set(c, j, DEtoBasisProb(c->DE[j]));                /* :154, DE still 0.0 */
    }
    c->infeas = 0;
    for (int i = 0; i < nR; ++i) { /* :161-169 */
        for (int k = p->rowPtr[i]; k < p->rowPtr[i + 1]; ++k)
            if (c->delta[p->rowIdx[k]] == -1) c->X[i] += p->rowVal[k];
        c->E[i] = energy(p, i, c->X[i]);
        c->infeas += infeasibility(p, i, c->X[i]);
    }
    for (int j = 0; j < nC; ++j) /* :172-178 */
        for (int k = p->colPtr[j]; k < p->colPtr[j + 1]; ++k) {
            int i = p->colIdx[k];
            c->DE[j] += energy(p, i, c->X[i] + c->delta[j] * p->colVal[k]) - c->E[i];
        }
    c->rngKind = RNG_MT;
    mt_seed(&c->mt, 5489u);
    return c;
}

ORC_API void orc_chain_destroy(OrcChain *c) {
    if (!c) return;
    free(c->X); free(c->E); free(c->delta); free(c->DE); free(c->tree); free(c->slot); free(c->touched);
    free(c->newDE); free(c->newX); free(c->occ); free(c->facLogP); free(c->staleStates); free(c->staleFactors);
    free(c->isStaleState); free(c->isStaleFactor); free(c->occUndo); free(c->facUndo);
    free(c);
}

ORC_API void orc_chain_seed_mt(OrcChain *c, uint32_t seed) { c->rngKind = RNG_MT; mt_seed(&c->mt, seed); }
ORC_API void orc_chain_seed_philox(OrcChain *c, uint64_t seed, uint64_t chainId, uint64_t iter) {
    c->rngKind = RNG_PHILOX;
    c->phKey[0] = (uint32_t)seed;
    c->phKey[1] = (uint32_t)(seed >> 32);
    c->phChain = chainId;
    c->phIter = iter;
    c->phDraw = 0;
}
ORC_API void orc_chain_set_injected(OrcChain *c, const double *u, long n) {
    c->rngKind = RNG_INJECT;
    c->inj = u;
    c->injLen = n;
    c->injPos = 0;
}
ORC_API void orc_chain_set_injected_proposals(OrcChain *c, const int32_t *j) { c->injJ = j; c->injJPos = 0; }
ORC_API uint64_t orc_chain_philox_iter(const OrcChain *c) { return c->phIter; }

/* --- Proposal::Proposal, SparseBasisSampler.h:288-315 --------------------------------------- */
typedef struct { int j, infeas; } Proposal;

static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }

static Proposal make_proposal(OrcChain *c) {
    const OrcProblem *p = c->p;
    Proposal pr;
    double u = rng_next(c);
    pr.j = tree_sample(c, u);                      /* :289 */
    if (c->injJ) pr.j = c->injJ[c->injJPos++];     /* replay mode: proposal stream injected */
    pr.infeas = c->infeas;
    int j = pr.j;
    c->nTouched = 0;
    c->slot[j] = c->nTouched; c->touched[c->nTouched] = j; c->newDE[c->nTouched++] = -c->DE[j]; /* :294 */
    int dj = c->delta[j];
    for (int k = p->colPtr[j]; k < p->colPtr[j + 1]; ++k) {
        int i = p->colIdx[k];
        int newXi = c->X[i] + p->colVal[k] * dj;
        c->newX[k - p->colPtr[j]] = newXi;
        double deltaEi = energy(p, i, newXi) - c->E[i];
        pr.infeas += infeasibility(p, i, newXi) - infeasibility(p, i, c->X[i]);
        for (int r = p->rowPtr[i]; r < p->rowPtr[i + 1]; ++r) {
            int uj = p->rowIdx[r];
            if (uj != j) {
                int s = c->slot[uj];
                if (s < 0) { /* try_emplace(updatej, currentDE[updatej]) */
                    s = c->slot[uj] = c->nTouched;
                    c->touched[c->nTouched] = uj;
                    c->newDE[c->nTouched++] = c->DE[uj];
                }
                int dx = p->rowVal[r] * c->delta[uj];
                c->newDE[s] += energy(p, i, newXi + dx) - energy(p, i, c->X[i] + dx) - deltaEi;
            }
        }
    }
    return pr;
}

/* ascending-column iteration order of std::map<int,double> changedDE */
static void sort_touched(OrcChain *c, int *order) {
    for (int k = 0; k < c->nTouched; ++k) order[k] = c->touched[k];
    qsort(order, c->nTouched, sizeof(int), cmp_int);
}
static void clear_touched(OrcChain *c) {
    for (int k = 0; k < c->nTouched; ++k) c->slot[c->touched[k]] = -1;
    c->nTouched = 0;
}

/* SparseBasisSampler.h:276-284 */
static double calcRatioOfSums(OrcChain *c, const int *order) {
    double changeInSum = 0.0;
    for (int k = 0; k < c->nTouched; ++k) {
        int uj = order[k];
        double oldP = tree_get(c, uj);
        changeInSum += DEtoBasisProb(c->newDE[c->slot[uj]]) - oldP;
    }
    return c->tree[0] / (c->tree[0] + changeInSum);
}

/* SparseBasisSampler.h:322-336 */
static void applyProposal(OrcChain *c, const Proposal *pr, const int *order, int updateX) {
    const OrcProblem *p = c->p;
    c->infeas = pr->infeas;
    int j = pr->j;
    for (int k = p->colPtr[j]; k < p->colPtr[j + 1]; ++k) {
        int i = p->colIdx[k];
        if (updateX) c->X[i] = c->newX[k - p->colPtr[j]];
        c->E[i] = energy(p, i, c->X[i]);
    }
    for (int k = 0; k < c->nTouched; ++k) {
        int uj = order[k];
        double nd = c->newDE[c->slot[uj]];
        c->DE[uj] = nd;
        tree_set(c, uj, DEtoBasisProb(nd));
    }
    c->delta[j] *= -1;
}

/* findInitialFeasibleSolution (:254-263) followed by :183-184 */
ORC_API long orc_chain_find_feasible(OrcChain *c, long maxIter) {
    long it = 0;
    int *order = (int *)malloc(sizeof(int) * c->p->nCols);
    while (c->infeas > 0 && (maxIter <= 0 || it < maxIter)) {
        Proposal pr = make_proposal(c);
        sort_touched(c, order);
        applyProposal(c, &pr, order, 1);
        clear_touched(c);
        rng_end_iteration(c);
        ++it;
    }
    free(order);
    imp_setState(c);
    c->logImp = imp_getLogValue(c);
    return it;
}

/* nextSampleWithInfeasibleImportance loop body, SparseBasisSampler.h:221-246.
 * Runs exactly nSteps Metropolis steps (feasible or not).  Optional per-step outputs. */
ORC_API void orc_chain_step(OrcChain *c, long nSteps, int32_t *outJ, uint8_t *outAcc, int32_t *outInf) {
    const OrcProblem *p = c->p;
    int *order = (int *)malloc(sizeof(int) * p->nCols);
    for (long s = 0; s < nSteps; ++s) {
        Proposal pr = make_proposal(c);
        int j = pr.j, n = p->colPtr[j + 1] - p->colPtr[j];
        const int *idx = p->colIdx + p->colPtr[j];
        imp_perturb(c, idx, n);                                            /* :224 */
        for (int k = 0; k < n; ++k) { int t = c->X[idx[k]]; c->X[idx[k]] = c->newX[k]; c->newX[k] = t; } /* :225-228 */
        double proposedLogImp = imp_getLogValue(c);                        /* :230 */
        sort_touched(c, order);
        double ratio = calcRatioOfSums(c, order);                          /* :231 */
        double acceptance = exp(proposedLogImp - c->logImp) * ratio;       /* :232 */
        int curF = c->infeas == 0, propF = pr.infeas == 0;
        int acc = rng_next(c) < acceptance;                                /* :234 */
        c->nProposals[curF][propF]++;
        if (acc) {
            c->nAccepted[curF][propF]++;
            c->logImp = proposedLogImp;
            applyProposal(c, &pr, order, 0);
        } else {
            for (int k = 0; k < n; ++k) c->X[idx[k]] = c->newX[k];
            imp_undo(c);
        }
        clear_touched(c);
        rng_end_iteration(c);
        if (outJ) outJ[s] = j;
        if (outAcc) outAcc[s] = (uint8_t)acc;
        if (outInf) outInf[s] = pr.infeas;
    }
    free(order);
}

/* operator() (:212-214): steps until the state is feasible; returns the number of steps taken */
ORC_API long orc_chain_next_sample(OrcChain *c) {
    long n = 0;
    do { orc_chain_step(c, 1, 0, 0, 0); ++n; } while (c->infeas > 0);
    return n;
}

/* --- getters -------------------------------------------------------------------------------- */
ORC_API void orc_chain_get_X(const OrcChain *c, int32_t *out) { memcpy(out, c->X, sizeof(int) * c->p->nRows); }
ORC_API void orc_chain_get_delta(const OrcChain *c, int8_t *out) { for (int j = 0; j < c->p->nCols; ++j) out[j] = (int8_t)c->delta[j]; }
ORC_API void orc_chain_get_DE(const OrcChain *c, double *out) { memcpy(out, c->DE, sizeof(double) * c->p->nCols); }
ORC_API void orc_chain_get_tree(const OrcChain *c, double *out) { memcpy(out, c->tree, sizeof(double) * c->p->nCols); }
ORC_API void orc_chain_get_leaf(const OrcChain *c, double *out) { for (int j = 0; j < c->p->nCols; ++j) out[j] = tree_get(c, j); }
ORC_API double orc_chain_sum(const OrcChain *c) { return c->tree[0]; }
ORC_API double orc_chain_logimp(const OrcChain *c) { return c->logImp; }
ORC_API int orc_chain_infeasibility(const OrcChain *c) { return c->infeas; }
ORC_API void orc_chain_get_counters(const OrcChain *c, int64_t *out) {
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) out[a * 2 + b] = c->nProposals[a][b];
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) out[4 + a * 2 + b] = c->nAccepted[a][b];
}
static uint64_t mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
ORC_API uint64_t orc_chain_hash_X(const OrcChain *c) {
    uint64_t h = 0;
    for (int i = 0; i < c->p->nRows; ++i) h += (uint64_t)(int64_t)c->X[i] * mix64((uint64_t)i);
    return h;
}
/* End-state occupancy ModelState(traj,T,T) (ModelState.h:27-36; State.h:57-63 over the agent's consequences(), e.g.
 * PredPreyAgent.h:98-115): out[psi] = sum of incoming events at t = T-1. */
ORC_API void orc_chain_end_state(const OrcChain *c, int32_t *out) {
    const OrcProblem *p = c->p;
    int S = p->S, A = p->A;
    for (int k = 0; k < S; ++k) out[k] = 0;
    const int *Xt = c->X + (p->T - 1) * S * A;
    for (int e = 0; e < S * A; ++e)
        for (int q = p->agCqPtr[e]; q < p->agCqPtr[e + 1]; ++q) out[p->agCq[q]] += Xt[e];
}
/* The reference's per-sample statistics (FiguresForPaper.h:93-104): for every feasible sample, the end state
 * ModelState(traj,T,T) is summed (Dataflow Sum) and the synopsis (FiguresForPaper.h:245-263: agents inside nested
 * squares of side G/2, G/4, ... > 1 along the diagonal) goes through MeanAndVariance (sum, sum of squares).
 * Steps until nSamples samples were taken or maxSteps steps were made; returns the samples taken.
 * series (may be NULL): every thin-th sample's synopsis, [cap][nSyn], indexed by global sample number / thin. */
ORC_API long orc_chain_sample_stats(OrcChain *c, long nSamples, long maxSteps, int64_t *endSum, int64_t *synSum,
                                    int64_t *synSumSq, long sampleBase, int32_t *series, long cap, long thin) {
    const OrcProblem *p = c->p;
    int G = p->G, GG = G * G, S = p->S;
    int32_t *end = (int32_t *)malloc(sizeof(int32_t) * S);
    long taken = 0;
    for (long st = 0; st < maxSteps && taken < nSamples; ++st) {
        orc_chain_step(c, 1, 0, 0, 0);
        if (c->infeas != 0) continue;
        orc_chain_end_state(c, end);
        for (int k = 0; k < S; ++k) endSum[k] += end[k];
        int varid = 0, origin = 0;
        for (int part = G / 2; part > 1; part /= 2) {
            int64_t occ = 0;
            for (int x = 0; x < part; ++x)
                for (int y = 0; y < part; ++y)
                    occ += end[(origin + x) + G * (origin + y)] + end[(origin + x) + G * (origin + y) + GG];
            synSum[varid] += occ;
            synSumSq[varid] += occ * occ;
            if (series) {
                long tot = sampleBase + taken;
                if (tot % thin == 0 && tot / thin < cap) series[(tot / thin) * 8 + varid] = (int32_t)occ;
            }
            ++varid;
            origin += part;
        }
        ++taken;
    }
    free(end);
    return taken;
}

/* bulk uniform generators, exposed so tests can feed the GPU the oracle's exact streams */
ORC_API void orc_mt_uniforms(uint32_t seed, long n, double *out) {
    Mt19937 g;
    mt_seed(&g, seed);
    for (long k = 0; k < n; ++k) { uint32_t w0 = mt_next(&g), w1 = mt_next(&g); out[k] = canonical_from_words(w0, w1); }
}
ORC_API void orc_philox_uniforms(uint64_t seed, uint64_t chainId, uint64_t iter0, long nIter, double *out) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (long k = 0; k < nIter; ++k) {
        uint64_t it = iter0 + (uint64_t)k;
        uint32_t ctr[4] = {(uint32_t)it, (uint32_t)(it >> 32), (uint32_t)chainId, (uint32_t)(chainId >> 32)}, o[4];
        philox4x32_10(ctr, key, o);
        out[2 * k] = canonical_from_words(o[0], o[1]);
        out[2 * k + 1] = canonical_from_words(o[2], o[3]);
    }
}
