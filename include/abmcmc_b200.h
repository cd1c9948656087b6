/* abmcmc_b200.h -- C ABI of the B200-native SparseBasisSampler.
 *
 * The reference (danftang/AgentBasedMCMC) has no FFI: its seam is the C++ class
 * SparseBasisSampler<T> (src/SparseBasisSampler.h:31-101).  This header is the C-ABI a drop-in for that
 * class binds to; every entry point cites the reference interface it replaces.  A header-only C++
 * adaptor with the reference's constructor signature lives in include/abmcmc_b200/gpu_sampler.hpp;
 * INTEGRATION.md shows the binding.
 *
 * Conventions: opaque handles; every function returns 0 on success or a negative abm_status and sets
 * a thread-local message readable through abm_last_error(); no exception crosses the boundary; host
 * buffers are caller-owned and copied before the call returns; a handle is bound to one CUDA device
 * and is not thread-safe (the reference is "one sampler per thread" too, FiguresForPaper.h:54-57).
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with ABM_ERR_CUDA.
 */
#ifndef ABMCMC_B200_H
#define ABMCMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABM_API __attribute__((visibility("default")))

typedef enum {
    ABM_OK = 0,
    ABM_ERR_INVALID = -1,     /* bad argument / unsupported problem shape */
    ABM_ERR_CUDA = -2,        /* CUDA runtime error or no device */
    ABM_ERR_STATE = -3,       /* call sequence error (e.g. step before init) */
    ABM_ERR_NOT_FEASIBLE = -4 /* feasibility search hit its iteration limit */
} abm_status;

typedef struct abm_problem abm_problem; /* shared read-only structure, replicated per device */
typedef struct abm_chains abm_chains;   /* a batch of independent Markov chains on one device */

/* Flat description of what SparseBasisSampler's constructor reads from its arguments
 * (SparseBasisSampler.h:107-158): the tableau restricted to non-basic columns, bounds and constants
 * per tableau row, the factor closures tabulated on the host, kappa, and -- for the importance
 * function TrajectoryImportance<PredPreyAgent<G>> (TrajectoryImportance.h, agents/PredPreyAgent.h) --
 * the two small tables it can ever add.  Array meanings equal the ABMP container
 * (include/abmcmc_b200/abmp_io.hpp). */
typedef struct {
    int32_t n_tab_rows;      /* tableau.rows.size(): one row per constraint, equalities included   */
    int32_t n_nonbasic;      /* tableau.nNonBasic(): the sampler's columns                        */
    const int32_t *tab_L;    /* [n_tab_rows] tableau.L (INT32_MIN = unbounded)                     */
    const int32_t *tab_U;    /* [n_tab_rows] tableau.U (INT32_MAX = unbounded); L==U <=> equality  */
    const int32_t *tab_F;    /* [n_tab_rows] tableau.F                                            */
    const int32_t *nb_col;   /* [n_nonbasic] tableau column (= variable / event id) of each column */
    const int32_t *nb_ptr;   /* [n_nonbasic+1] CSR pointer into nb_row / nb_val                    */
    const int32_t *nb_row;   /* [nnz] tableau row index, ascending inside a column                */
    const int32_t *nb_val;   /* [nnz] integer entry M_ij                                          */
    const int32_t *fac_id;   /* [n_rows] factor table of each sampler row (non-equality rows, in order) */
    int32_t n_tables;
    const int32_t *ut_lo;    /* [n_tables] first tabulated argument                               */
    const int32_t *ut_off;   /* [n_tables+1] offsets into ut_val                                  */
    const double *ut_val;    /* factor values f(x), x = lo .. lo+len-1 (clamp of reachable range) */
    double kappa;            /* SparseBasisSampler::kappa (:33)                                   */
    int32_t grid_size;       /* PredPrey GRIDSIZE; 0 = another agent (agent_* below) or, without those, no importance function (log W == 0) */
    int32_t n_timesteps;
    const double *imp_lgamma;   /* [7]  lgamma(occ+1) for occ>1 else 0 (TrajectoryImportance.h:96) */
    const double *imp_logratio; /* [2][2][6] log(pi/pibar) by type, has-neighbour flag, act (:105) */
    int32_t sum_tree;        /* storage of basisDistribution (MutableCategoricalArray): ABM_TREE_BLOCKED (0, default) or
                              * ABM_TREE_REFERENCE; the environment variable ABM_TREE=blocked|reference overrides it */
    /* TrajectoryImportance<AGENT> for an agent type OTHER than PredPreyAgent (grid_size == 0, n_agent_states > 0; all zero /
     * null otherwise), described by tables of the agent's own functions (reference_flatten.hpp fills them by calling
     * AGENT::timestep / marginalTimestep / neighbours / consequences on the host).  The form covered is the one every agent
     * of the reference has: timestep() of an agent state reads the other states only through "is any of up to four given
     * states (its influencers) occupied" (PredPreyAgent.h:124-157 -- the four adjacent cells of the other species;
     * CatMouseAgent.cpp:8-24 -- the cat at the mouse's position; BinomialAgent.h:33 -- none).  Event ids are
     * (t * n_agent_states + state) * n_acts + act (Event.h:18-19), state ids t * n_agent_states + state (State.h:40-42). */
    int32_t n_agent_states;            /* AGENT::domainSize()                                                  */
    int32_t n_acts;                    /* AGENT::actDomainSize(), at most 6                                    */
    int32_t n_agent_classes;           /* agent states with the same log-ratio rows share a class              */
    const int32_t *agent_class;        /* [n_agent_states]                                                     */
    const int32_t *agent_influencers;  /* [n_agent_states][4] agent states, -1 = unused                        */
    const int32_t *agent_nb_ptr;       /* [n_agent_states+1] AGENT::neighbours() in its own order: the states whose  */
    const int32_t *agent_nb;           /*   factor TrajectoryImportance::perturb marks stale (:48-51)           */
    const int32_t *agent_cq_ptr;       /* [n_agent_states*n_acts+1] AGENT::consequences(act): the states an event feeds */
    const int32_t *agent_cq;           /*   at the next timestep (ModelState.h:27-36 end-state statistics)      */
    /* with agent tables: imp_lgamma is [7] as above and imp_logratio is [n_agent_classes][2][n_acts] (class, any influencer
     * occupied, act) */
} abm_problem_desc;

/* ABM_TREE_REFERENCE keeps the reference's binary sum tree node for node and add for add (every double of the tree
 * reproduces MutableCategoricalArray's, rounding drift included).  ABM_TREE_BLOCKED keeps the same categorical -- same
 * leaves, same cdf order, so the same uniform selects the same column except within rounding distance (~1e-13) of an
 * interval boundary -- in a 32-ary blocked tree whose updates touch 3 nodes instead of 1 + popcount(index); get(u) is
 * then the exact leaf instead of a difference of drifting sums.  Integer state and accept decisions are identical in
 * both (tests/test_gpu_parity.py runs the golden reference traces through both). */
enum { ABM_TREE_BLOCKED = 0, ABM_TREE_REFERENCE = 1 };

typedef struct {
    int32_t n_rows, n_cols, nnz, n_events;
    int32_t max_col_nnz, max_coupled; /* longest column; most columns sharing a row with one column */
    int32_t tree_depth;               /* bit length of n_cols-1 (MutableCategoricalArray.cpp:41-48) */
    int32_t n_states;                 /* T * agent states per timestep: 2 * G * G for PredPrey (0 without importance) */
    int64_t chain_state_bytes;        /* HBM bytes of one chain (everything abm_chains_create allocates per chain) */
    int64_t shared_bytes;             /* HBM bytes of the shared structure                        */
} abm_problem_info;

/* Per-chain Metropolis counters: MCMCStatistics::nProposals / nAccepted (MCMCStatistics.h:14-15),
 * index [startFeasible][proposalFeasible], widened to 64 bit. */
typedef struct {
    int64_t n_proposals[2][2];
    int64_t n_accepted[2][2];
} abm_mcmc_counters;

ABM_API const char *abm_last_error(void);
ABM_API int abm_version(void);   /* 2 since abm_problem_desc carries the agent_* tables: zero-initialise the struct (`abm_problem_desc d = {0}`) */
ABM_API int abm_device_count(void); /* 0 when no CUDA device is usable */

/* ---- problem: replaces the tableau-copy half of SparseBasisSampler::SparseBasisSampler (:121-158) */
ABM_API int abm_problem_create(const abm_problem_desc *desc, int device, abm_problem **out);
ABM_API int abm_problem_destroy(abm_problem *p);
/* The host half of abm_problem_create alone -- every check (entry magnitudes, int8 row ranges, factor-table coverage, record
 * limits) and the sizes it would allocate -- without touching a device: a host program can reject a problem, or size its
 * chain count for the memory of a GPU (info->chain_state_bytes), before it owns one. */
ABM_API int abm_problem_validate(const abm_problem_desc *desc, abm_problem_info *info);
ABM_API int abm_problem_get_info(const abm_problem *p, abm_problem_info *info);

/* ---- chains: n_chains independent samplers; chain c draws from the Philox4x32-10 stream keyed by
 * (seed, first_chain_id + c).  stream = a cudaStream_t (NULL = default stream). */
ABM_API int abm_chains_create(const abm_problem *p, int32_t n_chains, uint64_t seed, uint64_t first_chain_id,
                              void *stream, abm_chains **out);
ABM_API int abm_chains_destroy(abm_chains *c);

/* The state-initialising half of the constructor (:144-178): delta from the initial trajectory
 * (init_events[c][e] > 0  =>  delta = -1), X = F + M.[delta==-1], currentDE, basisDistribution == 1.
 * init_events: host, [n_chains][n_init] int8 (n_init may be 0: all delta = +1). */
ABM_API int abm_chains_init(abm_chains *c, const int8_t *init_events, int32_t n_init);

/* Same, with one initial trajectory per chain drawn on the device from the prior: Forecast::nextSample(startState, T,
 * isFermionic = false) with the Bernoulli start state (Forecast.h:102-150, BernoulliStartState.h:36-47), which is what
 * FiguresForPaper.h:56,162 hands to every chain.  act_pmf = PredPreyAgent::timestep by (type, has opposite-species
 * neighbour, act) ([2][2][6], ABMP array pp_actpmf); Philox streams keyed by (seed, chain).  events_out (host,
 * [n_chains][n_events], may be NULL) receives the drawn trajectories as 0/1.  start_kind selects the start state
 * distribution: ABM_START_BERNOULLI (the one PredPreyProblem.h:64-66 uses) or ABM_START_POISSON (its commented-out
 * alternative, PredPreyProblem.h:60-62 / PoissonStartState.h:38-44: a Poisson(p) number of agents per state). */
#define ABM_START_BERNOULLI 0
#define ABM_START_POISSON 1
ABM_API int abm_chains_init_from_prior(abm_chains *c, const double *act_pmf, double p_predator, double p_prey, uint64_t seed,
                                       int8_t *events_out, int32_t start_kind);

/* findInitialFeasibleSolution (:254-263) + importanceFunc->setState / getLogValue (:183-184).
 * Uniforms: Philox unless `uniforms` (host, [n_chains][n_uniforms], one per iteration) is given.
 * iterations_out (host, [n_chains], may be NULL) receives the per-chain iteration count. */
ABM_API int abm_chains_find_feasible(abm_chains *c, int64_t max_iterations, const double *uniforms,
                                     int64_t n_uniforms, int64_t *iterations_out);

/* n_steps iterations of the do-loop of nextSampleWithInfeasibleImportance (:221-246) on every chain
 * (feasible or not: one step == one MCMCStatistics::addSample).  Asynchronous on the chains' stream. */
ABM_API int abm_chains_step(abm_chains *c, int64_t n_steps);

/* Same, with the random inputs injected and the decisions returned, for parity testing:
 *   uniforms  host [n_chains][2*n_steps]  (u1 = column draw, u2 = accept draw), or NULL for Philox
 *   proposals host [n_chains][n_steps] column indices replacing the sum-tree draw, or NULL
 *   out_j / out_accepted / out_infeasibility  host [n_chains][n_steps], each may be NULL
 * (out_infeasibility is Proposal::infeasibility, :303).  Synchronous. */
ABM_API int abm_chains_step_traced(abm_chains *c, int64_t n_steps, const double *uniforms, const int32_t *proposals,
                                   int32_t *out_j, uint8_t *out_accepted, int32_t *out_infeasibility);

/* operator() (:212-214) for every chain: step until n_samples further feasible states were visited
 * (a feasible state is a sample each time the loop condition :247 sees infeasibility == 0), accumulating
 * the ensemble statistics below after each one.  max_steps bounds the work per chain (0 = unbounded); n_samples == 0 with
 * max_steps > 0 means "sample for exactly max_steps steps". */
/* One call is one launch: at most 2^30 steps (max_steps == 0 stands for that bound) and, when the quota is the binding
 * limit, at most 2^30 samples; abm_chains_step splits longer requests into several launches itself. */
ABM_API int abm_chains_sample(abm_chains *c, int64_t n_samples, int64_t max_steps);

ABM_API int abm_chains_synchronize(abm_chains *c);

/* ---- state read-back (public members X, delta, currentDE, basisDistribution, currentInfeasibility,
 * currentLogImportance, stats; SparseBasisSampler.h:44-54).  Any pointer may be NULL.
 * tree = the raw sum-tree nodes in the reference's index order (MutableCategoricalArray.h:38). */
ABM_API int abm_chains_get_state(abm_chains *c, int32_t chain, int32_t *X, int8_t *delta, double *currentDE,
                                 double *tree, double *log_importance, int32_t *infeasibility);
ABM_API int abm_chains_get_counters(abm_chains *c, abm_mcmc_counters *per_chain /* [n_chains] */);
ABM_API int abm_chains_get_infeasibility(abm_chains *c, int32_t *per_chain /* [n_chains] */);
/* Event trajectories (the first n_events entries of X, Trajectory.h:27-29) of all chains, int8,
 * host [n_chains][n_events]. */
ABM_API int abm_chains_get_trajectories(abm_chains *c, int8_t *events);

/* ---- ensemble statistics (FiguresForPaper.h:93-104,245-263; diagnostics/MeanAndVariance.h:35-45)
 * accumulated on the device by abm_chains_sample:
 *   end_state_sum [n_agent_states]  sum over chains and samples of ModelState(traj,T,T) occupancies
 *   syn_sum / syn_sumsq [n_chains][n_synopsis]  per-chain sums of the synopsis statistics
 *   n_samples [n_chains]
 * All are int64 (exact).  Pointers are HOST buffers unless device_ptrs != 0 (then device buffers on the
 * chains' device, so a caller can hand them to an NCCL all-reduce without a host round trip). */
ABM_API int abm_chains_get_statistics(abm_chains *c, int64_t *end_state_sum, int64_t *syn_sum, int64_t *syn_sumsq,
                                      int64_t *n_samples, int device_ptrs);
/* The same read-back without stalling the stream (FiguresForPaper.h:93-104 drains its statistics while the sampler thread
 * keeps running): enqueue orders the ensemble reduction and the copies (into pinned staging owned by the handle) after
 * everything launched so far and returns; launches issued afterwards run on.  wait blocks until those copies have landed and
 * fills the caller's HOST buffers (any may be NULL); counters as abm_chains_get_counters.  One enqueue may be pending. */
ABM_API int abm_chains_enqueue_statistics(abm_chains *c);
ABM_API int abm_chains_wait_statistics(abm_chains *c, int64_t *end_state_sum, int64_t *syn_sum, int64_t *syn_sumsq,
                                       int64_t *n_samples, abm_mcmc_counters *per_chain_counters);
ABM_API int abm_chains_reset_statistics(abm_chains *c);
/* Optional thinned time series of the synopsis statistics (what `save(firstSynopsisSamples)` keeps, FiguresForPaper.h:96),
 * for the first n_chains chains: every `thin`-th feasible sample, at most `capacity` per chain.  The series feeds the
 * reference's variogram / effective-sample-size formulas (diagnostics/ChainStats.h:66-79, MultiChainStats.h:81-102).
 * abm_chains_get_series: host int32 [n_chains][capacity][n_synopsis]. */
ABM_API int abm_chains_record_series(abm_chains *c, int32_t n_chains, int32_t thin, int32_t capacity);
ABM_API int abm_chains_get_series(abm_chains *c, int32_t *out);
ABM_API int abm_problem_n_synopsis(const abm_problem *p); /* floor(log2(G)) - 1 (FiguresForPaper.h:246) */

/* The random inputs of a sampling launch injected, for parity testing of the statistics: exactly n_steps steps per chain with
 * uniforms host [n_chains][2*n_steps] (u1 = column draw, u2 = accept draw), statistics accumulated as by abm_chains_sample
 * after every step that ends feasible.  Synchronous. */
ABM_API int abm_chains_sample_injected(abm_chains *c, int64_t n_steps, const double *uniforms);

/* ---- ensemble moments and their multi-GPU reduction.  The reference adds up its threads'
 * results in a serial loop over std::futures (FiguresForPaper.h:62-65) and computes W, B, var+, R-hat from per-sequence
 * means and variances (diagnostics/MultiChainStats.h:141-181).  Here each GPU reduces its chains on the device to one short
 * ADDITIVE vector, and the vectors of all GPUs are summed by one NCCL all-reduce:
 *   exact[ABM_MOMENTS_EXACT_HEAD + n_agent_states] int64:
 *     [0] chains with >= 2 samples  [1] feasible samples  [2..10) MCMCStatistics counters (n_proposals[2][2], n_accepted[2][2])
 *     [10..18) sum over chains of the synopsis sums  [18..26) of the synopsis sums of squares (first n_synopsis entries used)
 *     [26 ...) end-state occupancy sums (ModelState(traj,T,T) summed over chains and samples)
 *   real[ABM_MOMENTS_REAL] double: [0..8) sum over chains of the chain mean, [8..16) of its square, [16..24) of the chain
 *     variance (MeanAndVariance::mean / sampleVariance, diagnostics/MeanAndVariance.h:52-59)
 * abm_chains_get_moments: this GPU's vectors into host buffers.  abm_chains_allreduce_moments: the same summed over all ranks
 * of `comm` (NULL = this GPU only).  enqueue / wait: the same without stalling the stream (one may be pending), like
 * abm_chains_enqueue_statistics; the read-back is ~17 KB instead of the per-chain tables. */
#define ABM_MOMENTS_EXACT_HEAD 26
#define ABM_MOMENTS_REAL 24
typedef struct abm_comm abm_comm;
/* NCCL plumbing without an NCCL type in the signature: rank 0 makes an id (ncclGetUniqueId, 128 bytes), the host program
 * distributes it by whatever it has (MPI, a file, torch.distributed), every rank then joins (ncclCommInitRank on the device
 * of its chains).  abm_comm_from_nccl borrows an existing ncclComm_t instead (not destroyed by abm_comm_destroy).
 * libnccl.so.2 is loaded on first use (environment ABM_NCCL_LIB overrides the name); without it these calls fail with
 * ABM_ERR_CUDA and single-GPU use is unaffected. */
#define ABM_COMM_ID_BYTES 128
ABM_API int abm_comm_unique_id(uint8_t id[ABM_COMM_ID_BYTES]);
ABM_API int abm_comm_create(int32_t n_ranks, int32_t rank, const uint8_t id[ABM_COMM_ID_BYTES], int device, abm_comm **out);
ABM_API int abm_comm_from_nccl(void *nccl_comm, abm_comm **out);
ABM_API int abm_comm_destroy(abm_comm *comm);
ABM_API int abm_chains_get_moments(abm_chains *c, int64_t *exact, double *real);
ABM_API int abm_chains_allreduce_moments(abm_chains *c, abm_comm *comm, int64_t *exact, double *real);
ABM_API int abm_chains_enqueue_moments(abm_chains *c, abm_comm *comm);
ABM_API int abm_chains_wait_moments(abm_chains *c, int64_t *exact, double *real);

/* Variogram of the recorded synopsis series over ALL chains (ChainStats::variogram, diagnostics/ChainStats.h:66-79, at the lags
 * j * thin, j < n_lags, of the series' first n_points points), reduced on the device and summed over the ranks of `comm` (NULL:
 * this GPU only): sums[j][k] = sum over the chains with at least min_samples feasible samples, and over the positions
 * i < n_points - j, of (x_i - x_{i+j})^2 for synopsis statistic k -- exact integers; V = sums / (n_chains * (n_points - j)).
 * Together with the moments it is everything MultiChainStats::effectiveSamples (MultiChainStats.h:81-102) needs. */
ABM_API int abm_chains_series_variogram(abm_chains *c, abm_comm *comm, int32_t n_points, int32_t n_lags, int64_t min_samples,
                                        int64_t *sums /* [n_lags][n_synopsis] */, int64_t *n_chains_out);

/* Blocked sum tree (ABM_TREE_BLOCKED): its sums are kept incrementally by the step kernel -- the two upper levels are recomputed
 * from the level below before every launch, the sums of 32 leaves from the leaves every 256th launch.  abm_chains_rebuild_tree
 * forces that exact rebuild now; abm_chains_tree_drift measures one chain: out = {root, max |sum - exact sum of its leaves| over
 * the sums of 32, of 1024, of 32768 leaves} (exact sums in extended precision on the host).  Measurement / test aid. */
ABM_API int abm_chains_rebuild_tree(abm_chains *c);
ABM_API int abm_chains_tree_drift(abm_chains *c, int32_t chain, double out[4]);

/* Device time of the last abm_chains_step / _step_traced / _sample launch (CUDA events on the chains' stream around the
 * step kernel; waits for it to finish).  For measurements: the traced variant's host copies are outside. */
ABM_API int abm_chains_last_launch_ms(abm_chains *c, float *ms);

/* ---- basis builder (host code, no device needed): the step BEFORE the sampler, SURVEY.md section 8(f)2.
 * abm_basis_minimise = TableauNormMinimiser<T>::TableauNormMinimiser(const ConvexPolyhedron<T>&) followed by
 * findMinimalBasis() (TableauNormMinimiser.h:201-256): constraints lower[i] <= sum_k val[k] X[idx[k]] <= upper[i]
 * (CSR, con_ptr[n_constraints + 1]; equality where lower == upper) become a tableau with one auxiliary variable per
 * inequality, and every equality row is eliminated by the reference's Markowitz pivoting (:258-335) -- the same pivots in
 * the same order, hence the same basis entry for entry; only the search skips ineligible rows/columns without walking
 * them (the reference's quadratic cost).  Outputs, in the layout reference_flatten.hpp writes to ABMP files:
 *   L, U, F, basis [n_rows]; col_is_basic [n_cols]; the non-basic columns as CSR by column over the NON-REDUCED rows
 *   (nb_col [n_nonbasic], nb_ptr [n_nonbasic + 1], nb_row / nb_val [nnz_nonbasic], rows ascending) -- what
 *   SparseBasisSampler's constructor reads (:121-158); abm_basis_get_rows gives every tableau row in full
 *   (row_ptr [n_rows + 1], ascending columns), from which a TableauNormMinimiser object can be re-assembled. */
typedef struct abm_basis abm_basis;
ABM_API int abm_basis_minimise(int32_t n_constraints, const int32_t *con_ptr, const int32_t *con_idx, const int32_t *con_val,
                               const int32_t *lower, const int32_t *upper, abm_basis **out);
ABM_API int abm_basis_destroy(abm_basis *b);
ABM_API int abm_basis_get_dims(const abm_basis *b, int32_t *n_rows, int32_t *n_cols, int32_t *n_aux, int32_t *n_nonbasic,
                               int64_t *nnz_nonbasic, int64_t *nnz_rows);
ABM_API int abm_basis_get(const abm_basis *b, int32_t *L, int32_t *U, int32_t *F, int32_t *basis, uint8_t *col_is_basic,
                          int32_t *nb_col, int32_t *nb_ptr, int32_t *nb_row, int32_t *nb_val);
ABM_API int abm_basis_get_rows(const abm_basis *b, int64_t *row_ptr, int32_t *col_idx, int32_t *val);

#ifdef __cplusplus
}
#endif
#endif /* ABMCMC_B200_H */
