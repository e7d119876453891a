/* TEST INFRASTRUCTURE ONLY - see sbq_oracle.h. Plain C11, no fast-math, no FMA contraction.
 * Citations are file:line under the ruolin/strawberry v1.1.2 checkout. */
#include "sbq_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------ */
/* EmSolver::init (src/estimate.cpp:366-409) + EmSolver::run (src/estimate.cpp:411-488), dense. */
int orc_em_dense(int T, int R, const int32_t* count, const double* alpha, const orc_em_params* p,
                 double* theta, int32_t* iters) {
   /* :374-375  theta_0 = (sum over ALL rows of n_i) / T */
   double total = 0.0;
   for (int i = 0; i < R; ++i) total += count[i];
   for (int j = 0; j < T; ++j) theta[j] = total / T;
   *iters = 0;

   /* :377-391  drop rows whose every entry is <= 1e-5 */
   int Rk = 0;
   int* rows = (int*)malloc(sizeof(int) * (size_t)(R > 0 ? R : 1));
   for (int i = 0; i < R; ++i) {
      int remove = 1;
      for (int j = 0; j < T; ++j)
         if (alpha[(size_t)i * T + j] > p->row_eps) remove = 0;
      if (!remove) rows[Rk++] = i;
   }
   if (Rk == 0) { free(rows); return ORC_NO_ROWS; }   /* :391 */

   double* F = (double*)malloc(sizeof(double) * (size_t)Rk * T);
   double* U = (double*)malloc(sizeof(double) * (size_t)Rk * T);
   double* u = (double*)malloc(sizeof(double) * (size_t)Rk);
   double* cur = (double*)malloc(sizeof(double) * (size_t)T);
   double* next = (double*)malloc(sizeof(double) * (size_t)T);
   for (int r = 0; r < Rk; ++r) {
      u[r] = (double)count[rows[r]];
      memcpy(F + (size_t)r * T, alpha + (size_t)rows[r] * T, sizeof(double) * (size_t)T);
   }
   memcpy(cur, theta, sizeof(double) * (size_t)T);

   int status = ORC_ITER_CAP;
   for (int it = 0; it < p->max_iter; ++it) {          /* :444 */
      *iters = it + 1;
      for (int r = 0; r < Rk; ++r) {                    /* E-step :449-458 */
         double denom = 0.0;
         for (int j = 0; j < T; ++j) denom += F[(size_t)r * T + j] * cur[j];
         if (denom == 0) {                              /* :451-453 run() returns false, _theta untouched */
            status = ORC_ZERO_DENOM;
            goto done_no_update;
         }
         for (int j = 0; j < T; ++j) {
            double num = u[r] * F[(size_t)r * T + j] * cur[j];
            U[(size_t)r * T + j] = num / denom;
         }
      }
      for (int j = 0; j < T; ++j) {                     /* M-step :462-464 */
         double s = 0.0;
         for (int r = 0; r < Rk; ++r) s += U[(size_t)r * T + j];
         next[j] = s;
      }
      for (int j = 0; j < T; ++j) {                     /* column normalisation :466-478 */
         double s = 0.0;
         for (int r = 0; r < Rk; ++r) s += F[(size_t)r * T + j];
         if (s != 0)                                    /* :469-471: the s==0 branch is a no-op '==' */
            for (int r = 0; r < Rk; ++r) F[(size_t)r * T + j] /= s;
      }
      double d2 = 0.0;                                  /* :479-480 */
      for (int j = 0; j < T; ++j) d2 += (next[j] - cur[j]) * (next[j] - cur[j]);
      if (sqrt(d2) < p->theta_tol) { status = ORC_OK; break; }   /* theta NOT advanced */
      memcpy(cur, next, sizeof(double) * (size_t)T);    /* :481 */
   }
   memcpy(theta, cur, sizeof(double) * (size_t)T);      /* :484-486 */
done_no_update:
   free(rows); free(F); free(U); free(u); free(cur); free(next);
   return status;
}

/* ------------------------------------------------------------------------------------------ */
/* Same algorithm on CSR rows. Skipping implicit zeros leaves every sum unchanged (x + 0.0 = x). */
int orc_em_csr(int T, int R, const int64_t* row_ptr, const int32_t* col, const double* alpha,
               const int32_t* count, const orc_em_params* p, double* theta, int32_t* iters) {
   double total = 0.0;
   for (int i = 0; i < R; ++i) total += count[i];
   for (int j = 0; j < T; ++j) theta[j] = total / T;
   *iters = 0;

   const int64_t base = R > 0 ? row_ptr[0] : 0;
   const int64_t nnz = R > 0 ? row_ptr[R] - base : 0;
   char* keep = (char*)calloc((size_t)(R > 0 ? R : 1), 1);
   int Rk = 0;
   for (int i = 0; i < R; ++i) {
      for (int64_t k = row_ptr[i]; k < row_ptr[i + 1]; ++k)
         if (alpha[k] > p->row_eps) keep[i] = 1;
      Rk += keep[i];
   }
   if (Rk == 0) { free(keep); return ORC_NO_ROWS; }

   double* F = (double*)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
   double* cur = (double*)malloc(sizeof(double) * (size_t)T);
   double* next = (double*)malloc(sizeof(double) * (size_t)T);
   double* s = (double*)malloc(sizeof(double) * (size_t)T);
   memcpy(F, alpha + base, sizeof(double) * (size_t)nnz);
   memcpy(cur, theta, sizeof(double) * (size_t)T);

   int status = ORC_ITER_CAP;
   for (int it = 0; it < p->max_iter; ++it) {
      *iters = it + 1;
      for (int j = 0; j < T; ++j) { next[j] = 0.0; s[j] = 0.0; }
      for (int i = 0; i < R; ++i) {
         if (!keep[i]) continue;
         double denom = 0.0;
         for (int64_t k = row_ptr[i]; k < row_ptr[i + 1]; ++k) denom += F[k - base] * cur[col[k]];
         if (denom == 0) { status = ORC_ZERO_DENOM; goto done_no_update; }
         for (int64_t k = row_ptr[i]; k < row_ptr[i + 1]; ++k) {
            double num = (double)count[i] * F[k - base] * cur[col[k]];
            next[col[k]] += num / denom;       /* rows visited in order => same order as U.col(j).sum() */
            s[col[k]] += F[k - base];
         }
      }
      for (int i = 0; i < R; ++i) {
         if (!keep[i]) continue;
         for (int64_t k = row_ptr[i]; k < row_ptr[i + 1]; ++k)
            if (s[col[k]] != 0) F[k - base] /= s[col[k]];
      }
      double d2 = 0.0;
      for (int j = 0; j < T; ++j) d2 += (next[j] - cur[j]) * (next[j] - cur[j]);
      if (sqrt(d2) < p->theta_tol) { status = ORC_OK; break; }
      memcpy(cur, next, sizeof(double) * (size_t)T);
   }
   memcpy(theta, cur, sizeof(double) * (size_t)T);
done_no_update:
   free(keep); free(F); free(cur); free(next); free(s);
   return status;
}

/* ------------------------------------------------------------------------------------------ */
/* src/estimate.cpp:310-356 */
double orc_epilogue(int T, const double* theta, const int32_t* iso_len, int64_t total_mapped_reads,
                    double min_iso_frac, int effective_len_norm, double insert_mean,
                    double* fpkm, double* frac, int32_t* keep, int32_t* na) {
   double sum_fpkm = 0.0;
   for (int j = 0; j < T; ++j) {
      double kb;
      na[j] = 0;
      fpkm[j] = 0.0;                                   /* Isoform::_FPKM default, include/isoform.h:53 */
      frac[j] = 0.0;                                   /* Isoform::_frac default, include/isoform.h:52 */
      if (effective_len_norm) {                        /* :317-324 */
         kb = iso_len[j] - insert_mean;
         if (kb < 0) { na[j] = 1; continue; }
         kb = 1e3 / kb;
      } else {
         kb = 1e3 / iso_len[j];                        /* :326 */
      }
      double rpm = 1e6 / (double)(int)total_mapped_reads;   /* :328, total_mapped_reads() is int */
      fpkm[j] = theta[j] * rpm * kb;                   /* :329 */
      sum_fpkm += fpkm[j];
   }
   double kept_sum = 0.0;
   for (int j = 0; j < T; ++j) {
      if (!na[j]) frac[j] = fpkm[j] / sum_fpkm;        /* :342 */
      keep[j] = !(frac[j] < min_iso_frac);             /* :346-355 (filter_by_expression is always true) */
      if (keep[j]) kept_sum += fpkm[j];
   }
   return kept_sum;
}

/* ------------------------------------------------------------------------------------------ */
typedef struct {
   int64_t n_loci;
   const int64_t *loc_row_off, *loc_iso_off, *row_ptr;
   const int32_t *col, *count, *iso_len;
   const double* alpha;
   int64_t total_mapped_reads;
   const orc_em_params* p;
   double min_iso_frac, insert_mean;
   int effective_len_norm;
   double *theta, *fpkm, *frac;
   int32_t *keep, *iters, *status;
   int64_t* next;
   pthread_mutex_t* mu;
} batch_job;

static void* batch_worker(void* arg) {
   batch_job* b = (batch_job*)arg;
   for (;;) {
      pthread_mutex_lock(b->mu);
      int64_t l0 = *b->next;
      *b->next = l0 + 16;
      pthread_mutex_unlock(b->mu);
      if (l0 >= b->n_loci) break;
      int64_t l1 = l0 + 16 < b->n_loci ? l0 + 16 : b->n_loci;
      for (int64_t l = l0; l < l1; ++l) {
         int64_t r0 = b->loc_row_off[l], t0 = b->loc_iso_off[l];
         int R = (int)(b->loc_row_off[l + 1] - r0), T = (int)(b->loc_iso_off[l + 1] - t0);
         b->status[l] = orc_em_csr(T, R, b->row_ptr + r0, b->col, b->alpha, b->count + r0, b->p,
                                   b->theta + t0, b->iters + l);
         int32_t* na = (int32_t*)malloc(sizeof(int32_t) * (size_t)(T > 0 ? T : 1));
         orc_epilogue(T, b->theta + t0, b->iso_len + t0, b->total_mapped_reads, b->min_iso_frac,
                      b->effective_len_norm, b->insert_mean, b->fpkm + t0, b->frac + t0, b->keep + t0, na);
         free(na);
         if (b->status[l] == ORC_NO_ROWS)
            for (int j = 0; j < T; ++j) b->keep[t0 + j] = 0;
      }
   }
   return NULL;
}

double orc_quantify_batch(int64_t n_loci, const int64_t* loc_row_off, const int64_t* loc_iso_off,
                          const int64_t* row_ptr, const int32_t* col, const double* alpha,
                          const int32_t* count, const int32_t* iso_len, int64_t total_mapped_reads,
                          const orc_em_params* p, double min_iso_frac, int effective_len_norm,
                          double insert_mean, int n_threads,
                          double* theta, double* fpkm, double* frac, double* tpm, int32_t* keep,
                          int32_t* iters, int32_t* status) {
   struct timespec t0, t1;
   clock_gettime(CLOCK_MONOTONIC, &t0);
   int64_t next = 0;
   pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
   batch_job b = {n_loci, loc_row_off, loc_iso_off, row_ptr, col, count, iso_len, alpha,
                  total_mapped_reads, p, min_iso_frac, insert_mean, effective_len_norm,
                  theta, fpkm, frac, keep, iters, status, &next, &mu};
   if (n_threads <= 1) {
      batch_worker(&b);
   } else {
      pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
      for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, batch_worker, &b);
      for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
      free(th);
   }
   /* src/alignments.cpp:1821-1829: TPM over the isoforms that survived, in locus order */
   int64_t n_iso = loc_iso_off[n_loci];
   double total_fpkm = 0.0;
   for (int64_t j = 0; j < n_iso; ++j)
      if (keep[j]) total_fpkm += fpkm[j];
   for (int64_t j = 0; j < n_iso; ++j) tpm[j] = 1e6 * fpkm[j] / total_fpkm;
   clock_gettime(CLOCK_MONOTONIC, &t1);
   return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------------------------------ */
/* Bias mode: OUR definition (see sbq_oracle.h). No reference behaviour exists - parity unpinned. */
static int solve_spd(int K, double* H, double* g) {
   /* Gaussian elimination with partial pivoting on the K x K system H d = g (H row-major); d returned in g */
   for (int c = 0; c < K; ++c) {
      int piv = c;
      for (int r = c + 1; r < K; ++r)
         if (fabs(H[r * K + c]) > fabs(H[piv * K + c])) piv = r;
      if (H[piv * K + c] == 0.0) return 0;
      if (piv != c) {
         for (int k = 0; k < K; ++k) { double t = H[c * K + k]; H[c * K + k] = H[piv * K + k]; H[piv * K + k] = t; }
         double t = g[c]; g[c] = g[piv]; g[piv] = t;
      }
      for (int r = c + 1; r < K; ++r) {
         const double f = H[r * K + c] / H[c * K + c];
         for (int k = c; k < K; ++k) H[r * K + k] -= f * H[c * K + k];
         g[r] -= f * g[c];
      }
   }
   for (int c = K - 1; c >= 0; --c) {
      double v = g[c];
      for (int k = c + 1; k < K; ++k) v -= H[c * K + k] * g[k];
      g[c] = v / H[c * K + c];
   }
   return 1;
}

int orc_em_bias_csr(int T, int R, const int64_t* row_ptr, const int32_t* col, const double* alpha, const int32_t* count,
                    const double* x, int K, const orc_bias_params* p, double* theta, double* beta, int32_t* iters,
                    int32_t* outer_iters) {
   double total = 0.0;
   for (int i = 0; i < R; ++i) total += count[i];
   for (int j = 0; j < T; ++j) theta[j] = total / T;
   for (int k = 0; k < K; ++k) beta[k] = 0.0;
   *iters = 0;
   *outer_iters = 0;
   char* keep = (char*)calloc((size_t)(R > 0 ? R : 1), 1);
   int Rk = 0;
   for (int i = 0; i < R; ++i) {
      for (int64_t k = row_ptr[i]; k < row_ptr[i + 1]; ++k)
         if (alpha[k] > p->row_eps) keep[i] = 1;
      Rk += keep[i];
   }
   if (Rk == 0) { free(keep); return ORC_NO_ROWS; }
   double* w = (double*)malloc(sizeof(double) * (size_t)R);
   double* d = (double*)malloc(sizeof(double) * (size_t)R);
   double* s = (double*)malloc(sizeof(double) * (size_t)T);
   double* th = (double*)malloc(sizeof(double) * (size_t)T);
   double* next = (double*)malloc(sizeof(double) * (size_t)T);
   double* cur = (double*)malloc(sizeof(double) * (size_t)T);
   double H[64], g[8], bprev[8];
   for (int i = 0; i < R; ++i) w[i] = 1.0;
   memcpy(cur, theta, sizeof(double) * (size_t)T);
   int status = ORC_ITER_CAP;
   for (int out = 0; out < p->max_out_it; ++out) {
      *outer_iters = out + 1;
      for (int j = 0; j < T; ++j) s[j] = 0.0;
      for (int i = 0; i < R; ++i)
         if (keep[i])
            for (int64_t k = row_ptr[i]; k < row_ptr[i + 1]; ++k) s[col[k]] += alpha[k] * w[i];
      /* theta-EM with the current bias */
      int conv = 0;
      for (int it = 0; it < p->max_theta_it; ++it) {
         *iters += 1;
         for (int j = 0; j < T; ++j) { th[j] = s[j] != 0 ? cur[j] / s[j] : 0.0; next[j] = 0.0; }
         for (int i = 0; i < R; ++i) {
            if (!keep[i]) continue;
            double dd = 0.0;
            for (int64_t k = row_ptr[i]; k < row_ptr[i + 1]; ++k) dd += alpha[k] * th[col[k]];
            if (dd == 0) { status = ORC_ZERO_DENOM; goto done; }
            const double r = (double)count[i] / dd;
            for (int64_t k = row_ptr[i]; k < row_ptr[i + 1]; ++k) next[col[k]] += alpha[k] * th[col[k]] * r;
         }
         double d2 = 0.0;
         for (int j = 0; j < T; ++j) d2 += (next[j] - cur[j]) * (next[j] - cur[j]);
         memcpy(cur, next, sizeof(double) * (size_t)T);
         if (sqrt(d2) < p->theta_tol) { conv = 1; break; }
      }
      (void)conv;
      if (K == 0) { status = ORC_OK; break; }
      /* bias update: Newton steps of the Poisson log-linear fit with d_i fixed */
      for (int j = 0; j < T; ++j) th[j] = s[j] != 0 ? cur[j] / s[j] : 0.0;
      for (int i = 0; i < R; ++i) {
         d[i] = 0.0;
         if (keep[i])
            for (int64_t k = row_ptr[i]; k < row_ptr[i + 1]; ++k) d[i] += alpha[k] * th[col[k]];
      }
      memcpy(bprev, beta, sizeof(double) * (size_t)K);
      for (int nb = 0; nb < p->max_bias_it; ++nb) {
         for (int a = 0; a < K * K; ++a) H[a] = 0.0;
         for (int a = 0; a < K; ++a) g[a] = 0.0;
         double tr = 0.0;
         for (int i = 0; i < R; ++i) {
            if (!keep[i]) continue;
            const double mu = w[i] * d[i];
            const double* xi = x + (size_t)i * K;
            for (int a = 0; a < K; ++a) {
               g[a] += ((double)count[i] - mu) * xi[a];
               for (int b = 0; b < K; ++b) H[a * K + b] += mu * xi[a] * xi[b];
            }
         }
         for (int a = 0; a < K; ++a) tr += H[a * K + a];
         for (int a = 0; a < K; ++a) H[a * K + a] += 1e-9 * tr + 1e-12;   /* ridge */
         if (!solve_spd(K, H, g)) break;
         double n2 = 0.0;
         for (int a = 0; a < K; ++a) { beta[a] += g[a]; n2 += g[a] * g[a]; }
         for (int i = 0; i < R; ++i) {
            double e = 0.0;
            for (int a = 0; a < K; ++a) e += beta[a] * x[(size_t)i * K + a];
            e = e > 30.0 ? 30.0 : (e < -30.0 ? -30.0 : e);
            w[i] = exp(e);
         }
         if (sqrt(n2) < p->bias_tol) break;
      }
      double m2 = 0.0;
      for (int a = 0; a < K; ++a) m2 += (beta[a] - bprev[a]) * (beta[a] - bprev[a]);
      if (sqrt(m2) < p->bias_tol) { status = ORC_OK; break; }
   }
   memcpy(theta, cur, sizeof(double) * (size_t)T);
done:
   free(keep); free(w); free(d); free(s); free(th); free(next); free(cur);
   return status;
}
