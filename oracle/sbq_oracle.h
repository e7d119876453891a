/* TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the quantification hot path of ruolin/strawberry v1.1.2
 * (citations are file:line under the reference checkout). It exists to CHECK the CUDA product
 * (strawberry_b200/csrc -> libsbq.so). Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product never links, imports or calls it.
 *
 * Parity status: PINNED. Every function here is checked (tests/test_oracle_*.py) against the
 * unmodified reference compiled from its own sources into oracle/_ref/libsbref.so, and against
 * committed golden vectors that were generated from that library (tests/golden/, generator
 * tests/golden/make_golden.py). Exception: the bias mode (orc_em_bias_csr) has no reference
 * implementation to pin against (src/bias.cpp is fully commented out) - parity unpinned.
 */
#ifndef SBQ_ORACLE_H_
#define SBQ_ORACLE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* locus status, same numbering as include/sbq.h */
enum { ORC_OK = 0, ORC_ITER_CAP = 1, ORC_ZERO_DENOM = 2, ORC_NO_ROWS = 3 };

typedef struct {
   int max_iter;      /* EmSolver::_max_iter_num = 1000          include/estimate.hpp:236 */
   double theta_tol;  /* EmSolver::_theta_change_limit = 1e-2    include/estimate.hpp:240 */
   double row_eps;    /* row filter threshold 1e-5               src/estimate.cpp:381     */
} orc_em_params;

/* EmSolver::init + EmSolver::run on a dense row-major R x T model (src/estimate.cpp:366-488). */
int orc_em_dense(int T, int R, const int32_t* count, const double* alpha, const orc_em_params* p,
                 double* theta, int32_t* iters);

/* Same semantics on CSR rows (explicit entries only; implicit entries are exact zeros). */
int orc_em_csr(int T, int R, const int64_t* row_ptr, const int32_t* col, const double* alpha,
               const int32_t* count, const orc_em_params* p, double* theta, int32_t* iters);

/* FPKM / frac / low-fraction filter of LocusContext::estimate_abundances (src/estimate.cpp:310-356).
 * keep[j] = 0 when the isoform is erased (frac < min_iso_frac), na[j] = 1 for the
 * effective_len_norm "NA" case. Returns the sum of FPKM over the isoforms kept. */
double orc_epilogue(int T, const double* theta, const int32_t* iso_len, int64_t total_mapped_reads,
                    double min_iso_frac, int effective_len_norm, double insert_mean,
                    double* fpkm, double* frac, int32_t* keep, int32_t* na);

/* Batched driver over flat loci (the shapes sbq_submit_flat takes), n_threads pthreads.
 * status 3 (NO_ROWS) loci get keep = 0 for every isoform (the reference emits nothing for them,
 * src/alignments.cpp:1526-1529). tpm follows src/alignments.cpp:1821-1829. Returns wall seconds. */
double orc_quantify_batch(int64_t n_loci, const int64_t* loc_row_off, const int64_t* loc_iso_off,
                          const int64_t* row_ptr, const int32_t* col, const double* alpha,
                          const int32_t* count, const int32_t* iso_len, int64_t total_mapped_reads,
                          const orc_em_params* p, double min_iso_frac, int effective_len_norm,
                          double insert_mean, int n_threads,
                          double* theta, double* fpkm, double* frac, double* tpm, int32_t* keep,
                          int32_t* iters, int32_t* status);

/* Bias mode (bias_mode = 1) - OUR definition; the reference has none (src/bias.cpp is commented
 * out), so this restatement is the only oracle of that mode: PARITY UNPINNED against the reference.
 *   row weight w_i = exp(clamp(beta . x_i, +-30)), biased model F_ij = alpha_ij w_i, s_j = sum_kept alpha_ij w_i
 *   outer loop (<= max_out_it): theta-EM to ||theta' - theta|| < theta_tol (<= max_theta_it, theta advanced),
 *   then <= max_bias_it Newton steps of the Poisson log-linear fit n_i ~ w_i * d_i (d_i = sum_j alpha_ij theta_j / s_j
 *   held fixed), stop when ||delta beta|| < bias_tol; outer stop when beta moved less than bias_tol.
 * x is row-major R x n_cov. Returns the locus status; iters = total theta-EM iterations. */
typedef struct {
   int max_out_it;      /* EmSolver::_max_out_it_num   = 100   include/estimate.hpp:239 (declared, never read) */
   int max_theta_it;    /* EmSolver::_max_theta_it_num = 5000  include/estimate.hpp:238 */
   int max_bias_it;     /* EmSolver::_max_bias_it_num  = 10    include/estimate.hpp:237 */
   double theta_tol;    /* 1e-2 */
   double bias_tol;     /* EmSolver::_bias_change_limit = 1e-2 include/estimate.hpp:241 */
   double row_eps;      /* 1e-5 */
} orc_bias_params;
int orc_em_bias_csr(int T, int R, const int64_t* row_ptr, const int32_t* col, const double* alpha, const int32_t* count,
                    const double* x, int n_cov, const orc_bias_params* p, double* theta, double* beta, int32_t* iters,
                    int32_t* outer_iters);

#ifdef __cplusplus
}
#endif
#endif
