// sbq_kernels.cuh - sm_100a EM kernels for the per-locus Latent-Class-Model quantification.
//
// What is computed (semantics of EmSolver::init + run, reference src/estimate.cpp:366-488, and of the
// FPKM/frac/filter tail of LocusContext::estimate_abundances, src/estimate.cpp:310-356):
//
//   total   = sum_i n_i over ALL rows;            theta_j = total / T                      (:374-375)
//   keep_i  = any_j alpha_ij > row_eps;           no kept row -> NO_ROWS                   (:377-391)
//   it = 0..max_iter-1:                                                                    (:444)
//       d_i      = sum_j F_ij theta_j  (kept rows);   any d_i == 0 -> ZERO_DENOM, theta := total/T (:449-453)
//       theta'_j = sum_i n_i F_ij theta_j / d_i                                           (:455-464)
//       F_ij    <- F_ij / s_j,  s_j = sum_i F_ij  (columns with s_j == 0 untouched)        (:466-478)
//       ||theta' - theta||_2 < tol -> stop, theta is NOT advanced                          (:479-480)
//       theta <- theta'                                                                    (:481)
//
// Design notes shared by the three tiers:
//  * The matrix is never rewritten. Column normalisation is idempotent after the first pass, so
//    F^(it>=1)_ij theta_j == alpha_ij * (theta_j / s_j): the kernels keep a scaled copy
//    th_j = theta_j / s_j in shared memory (th_j = theta_j during iteration 0) and stream raw alpha.
//  * One division per ROW (r_i = n_i / d_i), then u_ij = alpha_ij * th_j * r_i.
//  * No floating-point atomics anywhere. theta' is accumulated into accumulators that are PRIVATE to
//    the lane group that owns the row (a row's columns are distinct, so the lanes of one group never
//    collide) and then reduced over groups / CTAs in a fixed order -> results are bit-reproducible
//    run to run for a given plan.
//  * fp64 throughout.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sbq {
namespace cg = cooperative_groups;

enum : int { LOCUS_OK = 0, LOCUS_ITER_CAP = 1, LOCUS_ZERO_DENOM = 2, LOCUS_NO_ROWS = 3 };

struct DevParams {
   // flat batch (see include/sbq.h, sbq_submit_flat)
   const int64_t* loc_row_off;
   const int64_t* loc_iso_off;
   const int64_t* row_ptr;
   const int32_t* col;
   const double* alpha;
   const int32_t* count;
   const int32_t* iso_len;
   // scratch: per-row effective count, -1 = row dropped by the row filter
   int32_t* neff;
   // scratch: per-non-zero transposed index (CSR position | local row << 16) for cluster-tier slices whose CSR
   // fits shared memory but whose CSC index does not
   unsigned* csc;
   // per-isoform outputs
   double* theta;
   double* fpkm;
   double* frac;
   int32_t* keep;
   // per-locus outputs
   int32_t* iters;
   int32_t* status;
   double* locus_fpkm;   // sum of FPKM over the isoforms the locus keeps
   // configuration
   int max_iter;
   double tol;
   double row_eps;
   double min_frac;
   int eff_len_norm;
   double insert_mean;
   double rpm;           // 1e6 / total_mapped_reads   (src/estimate.cpp:328)
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
   return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
   return v;
}
template <int W>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
   for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
   return v;
}

// FPKM of one isoform (src/estimate.cpp:314-331). na = the effective_len_norm "NA" case.
__device__ __forceinline__ double iso_fpkm(const DevParams& p, double theta, int len, bool& na) {
   double kb;
   na = false;
   if (p.eff_len_norm) {
      kb = (double)len - p.insert_mean;
      if (kb < 0) { na = true; return 0.0; }
      kb = 1e3 / kb;
   } else {
      kb = 1e3 / (double)len;
   }
   return theta * p.rpm * kb;
}

// --------------------------------------------------------------------------------------------
// Tier 1: one warp per small locus (T <= 32). Lane i owns rows i, i+32, ...; lane j owns theta_j.
// Accumulators acc[j][lane] are lane-private (stride 33 doubles per column: conflict-free both for
// the lane-private updates and for the per-column fixed-order sum over lanes).
// --------------------------------------------------------------------------------------------
constexpr int WT_WARPS = 8;
constexpr int WT_MAX_ISO = 32;
constexpr int WT_STRIDE = 33;

__host__ __device__ inline size_t warp_tier_smem_bytes(int max_iso) {
   return (size_t)WT_WARPS * ((size_t)max_iso * WT_STRIDE + WT_MAX_ISO) * sizeof(double);
}

constexpr int WT_REG = 4;   // non-zeros per row a lane keeps in registers on the single-row fast path

// Persistent warps: every warp pulls the next locus from a global queue (list is sorted by descending
// non-zeros), so a long-running locus never holds back the other warps of its CTA.
__global__ void __launch_bounds__(WT_WARPS * 32)
em_warp_kernel(DevParams p, const int32_t* __restrict__ list, int n_list, int max_iso, int* queue) {
   extern __shared__ double smem[];
   const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
   double* acc = smem + (size_t)warp * ((size_t)max_iso * WT_STRIDE + WT_MAX_ISO);
   double* th = acc + (size_t)max_iso * WT_STRIDE;
   const int32_t* __restrict__ col = p.col;
   const double* __restrict__ al = p.alpha;
   const double tol2 = p.tol * p.tol;

   for (;;) {
      int w = 0;
      if (lane == 0) w = atomicAdd(queue, 1);
      w = __shfl_sync(0xffffffffu, w, 0);
      if (w >= n_list) break;
      const int l = list[w];
      const int64_t r0 = p.loc_row_off[l];
      const int R = (int)(p.loc_row_off[l + 1] - r0);
      const int64_t t0 = p.loc_iso_off[l];
      const int T = (int)(p.loc_iso_off[l + 1] - t0);
      const int64_t* __restrict__ rp = p.row_ptr + r0;
      int32_t* neff = p.neff + r0;

      for (int x = lane; x < T * WT_STRIDE; x += 32) acc[x] = 0.0;
      __syncwarp();

      // ---- setup: total, row filter, column sums over kept rows
      long long tot = 0;
      int kept = 0;
      int64_t my_k0 = 0;
      int my_cnt = 0, my_ne = -1;
      for (int i = lane; i < R; i += 32) {
         const int64_t k0 = rp[i], k1 = rp[i + 1];
         const int n = p.count[r0 + i];
         tot += n;
         bool keep = false;
         for (int64_t k = k0; k < k1; ++k) keep |= al[k] > p.row_eps;
         neff[i] = keep ? n : -1;
         if (keep) {
            ++kept;
            for (int64_t k = k0; k < k1; ++k) acc[col[k] * WT_STRIDE + lane] += al[k];
         }
         if (i == lane) { my_k0 = k0; my_cnt = (int)(k1 - k0); my_ne = keep ? n : -1; }
      }
      tot = warp_sum_ll(tot);
      kept = (int)warp_sum_ll(kept);
      __syncwarp();
      const int nl = R < 32 ? R : 32;   // lanes that own at least one row
      double s = 0.0;
      if (lane < T) {
         for (int x = 0; x < nl; ++x) { s += acc[lane * WT_STRIDE + x]; acc[lane * WT_STRIDE + x] = 0.0; }
      }
      // single-row fast path: one row per lane, at most WT_REG non-zeros each, held in registers
      const bool fast = R <= 32 && __all_sync(0xffffffffu, my_cnt <= WT_REG);
      double ra[WT_REG];
      int rc[WT_REG];
#pragma unroll
      for (int e = 0; e < WT_REG; ++e) {
         const bool v = fast && my_ne >= 0 && e < my_cnt;
         ra[e] = v ? al[my_k0 + e] : 0.0;
         rc[e] = v ? col[my_k0 + e] : 0;
      }
      const double theta0 = (double)tot / (double)T;
      double cur = theta0;
      int status = LOCUS_ITER_CAP, iters = 0;
      if (kept == 0) {
         status = LOCUS_NO_ROWS;
      } else {
         if (lane < T) th[lane] = cur;   // iteration 0 runs on raw alpha
         __syncwarp();
         for (int it = 0; it < p.max_iter; ++it) {
            iters = it + 1;
            bool zero = false;
            if (fast) {
               double t[WT_REG], d = 0.0;
#pragma unroll
               for (int e = 0; e < WT_REG; ++e) { t[e] = th[rc[e]]; d += ra[e] * t[e]; }
               if (my_ne >= 0) {
                  if (d == 0) {
                     zero = true;
                  } else {
                     const double r = (double)my_ne / d;
#pragma unroll
                     for (int e = 0; e < WT_REG; ++e)
                        if (ra[e] != 0.0) acc[rc[e] * WT_STRIDE + lane] += ra[e] * t[e] * r;
                  }
               }
            } else {
               for (int i = lane; i < R; i += 32) {
                  const int ne = neff[i];
                  if (ne < 0) continue;
                  const int64_t k0 = rp[i], k1 = rp[i + 1];
                  double d = 0.0;
                  for (int64_t k = k0; k < k1; ++k) d += al[k] * th[col[k]];
                  if (d == 0) {
                     zero = true;
                  } else {
                     const double r = (double)ne / d;
                     for (int64_t k = k0; k < k1; ++k) {
                        const int c = col[k];
                        acc[c * WT_STRIDE + lane] += al[k] * th[c] * r;
                     }
                  }
               }
            }
            if (__any_sync(0xffffffffu, zero)) { status = LOCUS_ZERO_DENOM; break; }
            __syncwarp();
            double nw = 0.0;
            if (lane < T) {
               double n0 = 0.0, n1 = 0.0;   // fixed two-way split of the sum over lanes (order does not depend on data)
               int x = 0;
               for (; x + 1 < nl; x += 2) {
                  n0 += acc[lane * WT_STRIDE + x];
                  n1 += acc[lane * WT_STRIDE + x + 1];
                  acc[lane * WT_STRIDE + x] = 0.0;
                  acc[lane * WT_STRIDE + x + 1] = 0.0;
               }
               if (x < nl) { n0 += acc[lane * WT_STRIDE + x]; acc[lane * WT_STRIDE + x] = 0.0; }
               nw = n0 + n1;
            }
            const double diff = lane < T ? nw - cur : 0.0;
            const double d2 = warp_sum(diff * diff);
            if (d2 < tol2) { status = LOCUS_OK; break; }   // ||theta' - theta||_2 < tol  (src/estimate.cpp:479-480)
            cur = nw;
            __syncwarp();
            if (lane < T) th[lane] = (s != 0) ? cur / s : 0.0;
            __syncwarp();
         }
      }

      // ---- outputs + epilogue (src/estimate.cpp:310-356)
      const double theta_out = (status == LOCUS_ZERO_DENOM || status == LOCUS_NO_ROWS) ? theta0 : cur;
      bool na = false;
      double f = 0.0;
      if (lane < T && status != LOCUS_NO_ROWS) f = iso_fpkm(p, theta_out, p.iso_len[t0 + lane], na);
      const double sum = warp_sum(f);
      double fr = 0.0;
      int kp = 0;
      if (lane < T && status != LOCUS_NO_ROWS) {
         if (!na) fr = f / sum;
         kp = !(fr < p.min_frac) ? (na ? -1 : 1) : 0;
      }
      const double kept_sum = warp_sum(kp != 0 ? f : 0.0);
      if (lane < T) {
         p.theta[t0 + lane] = theta_out;
         p.fpkm[t0 + lane] = f;
         p.frac[t0 + lane] = fr;
         p.keep[t0 + lane] = kp;
      }
      if (lane == 0) {
         p.iters[l] = iters;
         p.status[l] = status;
         p.locus_fpkm[l] = kept_sum;
      }
      __syncwarp();
   }
}

// --------------------------------------------------------------------------------------------
// Tier 2: one thread-block CLUSTER per locus (cluster size 1, 2, 4, 8 or 16 chosen by nnz).
// Rows are split over the CTAs of the cluster by non-zero count; inside a CTA, groups of LPR lanes
// own one row at a time and a private accumulator row acc[g][T] in shared memory. Per iteration:
// row pass -> fixed-order sum over groups -> partial theta' exchanged through distributed shared
// memory (one cluster barrier, double-buffered) -> every CTA forms the same theta' and norm.
// --------------------------------------------------------------------------------------------
constexpr int CL_NT = 512;           // threads per CTA of the cluster tier (small single-CTA loci use CL_NT_SMALL)
constexpr int CL_NT_SMALL = 128;
constexpr int CL_LPR_STREAM = 32;    // lanes per row of the streaming fallback

__host__ __device__ inline size_t cluster_fixed_doubles(int T) { return (size_t)6 * T + 8; }

// groups of the streaming fallback that fit next to the fixed arrays
__host__ __device__ inline int cluster_stream_groups(int T, size_t smem_bytes, int nt = CL_NT) {
   const long long budget = (long long)(smem_bytes / sizeof(double)) - (long long)cluster_fixed_doubles(T);
   long long G = budget / (T > 0 ? T : 1);
   if (G > nt / CL_LPR_STREAM) G = nt / CL_LPR_STREAM;
   return (int)G;   // 0 => does not fit
}
// bytes of one CTA's resident slice: CSR part (alpha, col, row pointers, counts, r) and the CSC index (pos, row)
__host__ __device__ inline size_t cluster_resident_csr_bytes(size_t nnz_c, size_t nrows, int T) {
   return nnz_c * 8 + nrows * 8 + (nrows + 1) * 4 + nrows * 4 + ((size_t)T + 1) * 4 + nnz_c * 2 + 64;
}
__host__ __device__ inline size_t cluster_resident_bytes(size_t nnz_c, size_t nrows, int T) {
   return cluster_resident_csr_bytes(nnz_c, nrows, T) + nnz_c * 4 + 4;
}

template <int NT>
__device__ __forceinline__ double block_sum(double v, double* red) {
   // fixed shape: warp butterfly, then every thread adds the warp totals in warp order
   v = warp_sum(v);
   if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
   __syncthreads();
   double t = 0.0;
#pragma unroll
   for (int w = 0; w < NT / 32; ++w) t += red[w];
   __syncthreads();
   return t;
}

// Row accessor of one CTA's slice streamed from global/L2 (fallback when the slice does not fit shared memory).
struct GlobalRows {
   const int64_t* rp;      // row_ptr of the slice's first row
   const double* al;       // alpha + k_a
   const int32_t* col;     // col + k_a
   int32_t* ne;            // neff of the slice's first row
   int64_t k_a;
   __device__ __forceinline__ unsigned start(int i) const { return (unsigned)(rp[i] - k_a); }
   __device__ __forceinline__ double a(unsigned k) const { return al[k]; }
   __device__ __forceinline__ int c(unsigned k) const { return col[k]; }
   __device__ __forceinline__ int get_ne(int i) const { return ne[i]; }
   __device__ __forceinline__ void set_ne(int i, int v) const { ne[i] = v; }
};

// streaming fallback, setup pass: total count, row filter (-> ne), column sums of the kept rows
template <int LPR, typename Rows>
__device__ __forceinline__ void cluster_setup_pass(const Rows& rows, const int32_t* __restrict__ cnt, int nrows, int G, int g, int lg,
                                                   double row_eps, double* my_acc, long long& tot, int& kept) {
   for (int base = 0; base < nrows; base += G) {
      const int i = base + g;
      const bool valid = i < nrows;
      unsigned k0 = 0, k1 = 0;
      int n = 0;
      if (valid) { k0 = rows.start(i); k1 = rows.start(i + 1); n = cnt[i]; }
      bool keep = false;
      for (unsigned k = k0 + lg; k < k1; k += LPR) keep |= rows.a(k) > row_eps;
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) keep |= (bool)__shfl_xor_sync(0xffffffffu, (int)keep, o);
      if (valid && lg == 0) { rows.set_ne(i, keep ? n : -1); tot += n; kept += keep; }
      if (keep)
         for (unsigned k = k0 + lg; k < k1; k += LPR) my_acc[rows.c(k)] += rows.a(k);
      __syncwarp();   // the next row of this group may touch the same accumulator column from another lane
   }
}

// streaming fallback, one E/M pass with the scaled theta in th[]; two rows per group in flight
template <int LPR, typename Rows>
__device__ __forceinline__ void cluster_em_pass(const Rows& rows, int nrows, int G, int g, int lg, const double* th, double* my_acc, int& zero) {
   for (int base = 0; base < nrows; base += 2 * G) {
      const int i0 = base + g, i1 = base + G + g;
      int ne0 = -1, ne1 = -1;
      unsigned a0 = 0, b0 = 0, a1 = 0, b1 = 0;
      if (i0 < nrows) { ne0 = rows.get_ne(i0); if (ne0 >= 0) { a0 = rows.start(i0); b0 = rows.start(i0 + 1); } }
      if (i1 < nrows) { ne1 = rows.get_ne(i1); if (ne1 >= 0) { a1 = rows.start(i1); b1 = rows.start(i1 + 1); } }
      double d0 = 0.0, d1 = 0.0;
      for (unsigned k = a0 + lg; k < b0; k += LPR) d0 += rows.a(k) * th[rows.c(k)];
      for (unsigned k = a1 + lg; k < b1; k += LPR) d1 += rows.a(k) * th[rows.c(k)];
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) {
         d0 += __shfl_xor_sync(0xffffffffu, d0, o);
         d1 += __shfl_xor_sync(0xffffffffu, d1, o);
      }
      if (ne0 >= 0) {
         if (d0 == 0) zero = 1;
         else {
            const double r = (double)ne0 / d0;
            for (unsigned k = a0 + lg; k < b0; k += LPR) { const int c = rows.c(k); my_acc[c] += rows.a(k) * th[c] * r; }
         }
      }
      __syncwarp();   // row i1 of the same group may share columns with row i0
      if (ne1 >= 0) {
         if (d1 == 0) zero = 1;
         else {
            const double r = (double)ne1 / d1;
            for (unsigned k = a1 + lg; k < b1; k += LPR) { const int c = rows.c(k); my_acc[c] += rows.a(k) * th[c] * r; }
         }
      }
      __syncwarp();   // the next rows of this group may touch the same accumulator columns from other lanes
   }
}

// Resident slice of one CTA: CSR arrays plus a per-CTA transposed (CSC) index, all in shared memory.
struct ResidentSlice {
   double* al;            // [nnz]   alpha, CSR order
   double* r;             // [nrows] per-row r_i = n_i / d_i of the current iteration (0 for dropped rows)
   unsigned* rp;          // [nrows+1]
   int* ne;               // [nrows]
   unsigned* cp;          // [T+1]   column pointers into pos/row
   unsigned short* col;   // [nnz]   CSR order (ascending within a row)
   unsigned* ent;         // [nnz]   CSC order, rows ascending within a column: CSR position | local row << 16
                          //         (shared memory, or global scratch when only the CSR part fits)
   int nrows, T;
   unsigned nnz;
};

// lanes per row / column: the largest power of two <= 32 that still gives every item its own lane group
#ifndef SBQ_MIN_LANES
#define SBQ_MIN_LANES 1
#endif
__device__ __forceinline__ int lanes_for(int items, int nt) {
   int l = 32;
   while (l > SBQ_MIN_LANES && (long long)items * l > nt) l >>= 1;
   return l;
}

// E-step on the resident slice: d_i = sum_k alpha_k th[col_k]  ->  r_i = n_i / d_i   (one division per row)
template <int NT>
__device__ __forceinline__ void resident_e_pass(const ResidentSlice& S, const double* th, int lpr, int& zero) {
   const int tid = threadIdx.x, g = tid / lpr, lg = tid % lpr, par = NT / lpr;
   for (int base = 0; base < S.nrows; base += par) {
      const int i = base + g;
      int ne = -1;
      unsigned k0 = 0, k1 = 0;
      if (i < S.nrows) {
         ne = S.ne[i];
         if (ne >= 0) { k0 = S.rp[i]; k1 = S.rp[i + 1]; }
      }
      double d = 0.0;
      for (unsigned k = k0 + lg; k < k1; k += 4 * lpr) {   // four independent gathers in flight per lane
         double a[4];
         int c[4];
#pragma unroll
         for (int u = 0; u < 4; ++u) {
            const bool v = k + u * lpr < k1;
            a[u] = v ? S.al[k + u * lpr] : 0.0;
            c[u] = v ? S.col[k + u * lpr] : 0;
         }
#pragma unroll
         for (int u = 0; u < 4; ++u) d += a[u] * th[c[u]];
      }
      for (int o = lpr >> 1; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (i < S.nrows && lg == 0) {
         double r = 0.0;
         if (ne >= 0) {
            if (d == 0) zero = 1; else r = (double)ne / d;
         }
         S.r[i] = r;
      }
   }
}

// M-step on the resident slice through the CSC index: out_j = scale_j * sum_{k in column j} alpha_k r_row(k).
// Fixed lane assignment and reduction shape -> deterministic, no accumulators, no atomics. Eight index entries per
// lane are fetched before they are used so that a CSC index living in global scratch (L2) is latency-tolerant.
template <int NT>
__device__ __forceinline__ void resident_col_pass(const ResidentSlice& S, const double* scale, double* out, int lpc) {
   const int tid = threadIdx.x, g = tid / lpc, lg = tid % lpc, par = NT / lpc;
   const unsigned* __restrict__ ent = S.ent;
   for (int base = 0; base < S.T; base += par) {
      const int j = base + g;
      double sum = 0.0;
      if (j < S.T) {
         const unsigned x1 = S.cp[j + 1];
         for (unsigned x = S.cp[j] + lg; x < x1; x += 8 * lpc) {
            unsigned e[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) e[u] = (x + u * lpc < x1) ? ent[x + u * lpc] : 0xffffffffu;
#pragma unroll
            for (int u = 0; u < 8; ++u)
               if (e[u] != 0xffffffffu) sum += S.al[e[u] & 0xffffu] * S.r[e[u] >> 16];
         }
      }
      for (int o = lpc >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (j < S.T && lg == 0) out[j] = scale ? scale[j] * sum : sum;
   }
}

template <int NT>
__global__ void __launch_bounds__(NT)
em_cluster_kernel(DevParams p, const int32_t* __restrict__ list, int n_list, unsigned smem_bytes) {
   cg::cluster_group cluster = cg::this_cluster();
   const unsigned CS = cluster.num_blocks();
   const unsigned rank = cluster.block_rank();
   const int item = blockIdx.x / CS;
   const int l = list[item];
   const int tid = threadIdx.x, lane = tid & 31;

   const int64_t r0 = p.loc_row_off[l];
   const int R = (int)(p.loc_row_off[l + 1] - r0);
   const int64_t t0 = p.loc_iso_off[l];
   const int T = (int)(p.loc_iso_off[l + 1] - t0);

   extern __shared__ double smem[];
   double* th = smem;                  // [T] scaled theta used by the passes
   double* bufA = th + T;              // [T] theta (current / next, swapped by pointer)
   double* bufB = bufA + T;            // [T]
   double* sdiv = bufB + T;            // [T] column sums s_j of the kept rows
   double* part = sdiv + T;            // [2][T+4] exchange buffers: partial theta', flag, total, kept
   double* dyn = part + 2 * (T + 4);   // resident slice, or the streaming accumulators
   __shared__ double red[NT / 32];
   __shared__ int s_rows[2];
   double* cur = bufA;
   double* nxt = bufB;

   const int64_t* __restrict__ rp = p.row_ptr + r0;

   // rows of this CTA: split by non-zeros (lower_bound on row_ptr)
   if (tid < 2) {
      const int64_t base = rp[0], nnz = rp[R] - base;
      const int64_t target = base + (nnz * (int64_t)(rank + tid)) / CS;
      int lo = 0, hi = R;
      if (rank + tid >= CS) lo = R;
      else if (rank + tid == 0) hi = 0;
      while (lo < hi) {
         const int mid = (lo + hi) >> 1;
         if (rp[mid] < target) lo = mid + 1; else hi = mid;
      }
      s_rows[tid] = lo;
   }
   __syncthreads();
   const int ra = s_rows[0], rb = s_rows[1], nrows = rb - ra;
   const int64_t k_a = rp[ra];
   const unsigned nnz_c = (unsigned)(rp[rb] - k_a);

   const size_t used = (size_t)((char*)dyn - (char*)smem);
   const bool small_idx = nnz_c <= 65535u && nrows <= 65535;
   const bool csc_in_smem = small_idx && used + cluster_resident_bytes(nnz_c, (size_t)nrows, T) <= smem_bytes;
   const bool resident = small_idx && used + cluster_resident_csr_bytes(nnz_c, (size_t)nrows, T) <= smem_bytes;
   ResidentSlice S;
   S.al = dyn;
   S.r = S.al + nnz_c;
   S.rp = (unsigned*)(S.r + nrows);
   S.ne = (int*)(S.rp + nrows + 1);
   S.cp = (unsigned*)(S.ne + nrows);
   S.col = (unsigned short*)(S.cp + T + 1);
   if (csc_in_smem) S.ent = (unsigned*)(S.col + nnz_c + (nnz_c & 1u));                       // 4-byte aligned
   else S.ent = p.csc + (size_t)k_a;   // CSC index of this slice in global scratch (L2); the CSR part stays in shared memory
   S.nrows = nrows; S.T = T; S.nnz = nnz_c;
   GlobalRows grows{rp + ra, p.alpha + k_a, p.col + k_a, p.neff + r0 + ra, k_a};
   const int G = resident ? 0 : cluster_stream_groups(T, smem_bytes, NT);
   double* acc = dyn;                                           // [G][T] (streaming only)
   const int g32 = tid / CL_LPR_STREAM, lg32 = tid % CL_LPR_STREAM;
   double* my_acc = acc + (size_t)(g32 < G ? g32 : 0) * T;
   const int lpr = lanes_for(nrows, NT), lpc = lanes_for(T, NT);

   long long tot = 0;
   int kept = 0;
   if (resident) {
      // ---- load the slice, row filter, CSC index
      const int32_t* __restrict__ cnt = p.count + r0 + ra;
      for (unsigned k = tid; k < nnz_c; k += NT) { S.al[k] = grows.al[k]; S.col[k] = (unsigned short)grows.col[k]; }
      for (int i = tid; i <= nrows; i += NT) S.rp[i] = (unsigned)(rp[ra + i] - k_a);
      for (int j = tid; j <= T; j += NT) S.cp[j] = 0;
      __syncthreads();
      {
         const int g = tid / lpr, lg = tid % lpr, par = NT / lpr;
         for (int base = 0; base < nrows; base += par) {
            const int i = base + g;
            unsigned k0 = 0, k1 = 0;
            if (i < nrows) { k0 = S.rp[i]; k1 = S.rp[i + 1]; }
            int keep = 0;
            for (unsigned k = k0 + lg; k < k1; k += lpr) keep |= S.al[k] > p.row_eps;
            for (int o = lpr >> 1; o > 0; o >>= 1) keep |= __shfl_xor_sync(0xffffffffu, keep, o);
            if (i < nrows && lg == 0) {
               const int n = cnt[i];
               S.ne[i] = keep ? n : -1;
               S.r[i] = keep ? 1.0 : 0.0;
               tot += n;
               kept += keep;
            }
         }
      }
      for (unsigned k = tid; k < nnz_c; k += NT) atomicAdd(&S.cp[S.col[k] + 1], 1u);       // column counts (integers: exact)
      __syncthreads();
      if (tid < 32) {                                                                        // exclusive scan over T (warp 0)
         const int chunk = (T + 31) / 32, j0 = lane * chunk, j1 = min(T, j0 + chunk);
         unsigned sum = 0;
         for (int j = j0; j < j1; ++j) sum += S.cp[j + 1];
         unsigned incl = sum;
         for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
         }
         unsigned run = incl - sum;
         for (int j = j0; j < j1; ++j) { const unsigned c = S.cp[j + 1]; S.cp[j + 1] = run + c; run += c; }
      }
      __syncthreads();
      // ordered fill, one warp per column: rows are visited in order and a row holds column j at most once
      // (binary search, columns ascend within a row), so every column lists its entries by ascending row.
      for (int j = tid >> 5; j < T; j += NT / 32) {
         unsigned w = S.cp[j];
         const unsigned w_end = S.cp[j + 1];
         for (int base = 0; base < nrows && w < w_end; base += 32) {
            const int i = base + lane;
            unsigned found = 0xffffffffu;
            if (i < nrows) {
               unsigned lo = S.rp[i], hi = S.rp[i + 1];
               while (lo < hi) {
                  const unsigned mid = (lo + hi) >> 1;
                  if (S.col[mid] < (unsigned)j) lo = mid + 1; else hi = mid;
               }
               if (lo < S.rp[i + 1] && S.col[lo] == (unsigned)j) found = lo;
            }
            const unsigned m = __ballot_sync(0xffffffffu, found != 0xffffffffu);
            if (found != 0xffffffffu) S.ent[w + __popc(m & ((1u << lane) - 1u))] = found | ((unsigned)i << 16);
            w += __popc(m);
         }
      }
      __syncthreads();
      resident_col_pass<NT>(S, nullptr, part, lpc);                                         // s_j partial (r = keep flag)
   } else {
      for (int x = tid; x < G * T; x += NT) acc[x] = 0.0;
      __syncthreads();
      if (g32 < G) cluster_setup_pass<CL_LPR_STREAM>(grows, p.count + r0 + ra, nrows, G, g32, lg32, p.row_eps, my_acc, tot, kept);
      __syncthreads();
      for (int j = tid; j < T; j += NT) {
         double sj = 0.0;
         for (int gg = 0; gg < G; ++gg) { sj += acc[(size_t)gg * T + j]; acc[(size_t)gg * T + j] = 0.0; }
         part[j] = sj;
      }
   }
   {
      const double tot_b = block_sum<NT>((double)tot, red);
      const double kept_b = block_sum<NT>((double)kept, red);
      if (tid == 0) { part[T + 1] = tot_b; part[T + 2] = kept_b; }
   }
   cluster.sync();
   double total = 0.0, kept_all = 0.0;
   for (unsigned r = 0; r < CS; ++r) {
      const double* rp_part = CS > 1 ? cluster.map_shared_rank(part, r) : part;
      total += rp_part[T + 1];
      kept_all += rp_part[T + 2];
   }
   for (int j = tid; j < T; j += NT) {
      double sj = 0.0;
      for (unsigned r = 0; r < CS; ++r) {
         const double* rp_part = CS > 1 ? cluster.map_shared_rank(part, r) : part;
         sj += rp_part[j];
      }
      sdiv[j] = sj;
   }
   const double theta0 = total / (double)T;
   for (int j = tid; j < T; j += NT) { cur[j] = theta0; th[j] = theta0; }
   cluster.sync();   // everyone has read part[0] before it is reused; th/cur visible CTA-wide

   const double tol2 = p.tol * p.tol;
   int status = LOCUS_ITER_CAP, iters = 0;
#ifdef SBQ_PHASE_TIMING
   long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define SBQ_TICK(k) { const long long t_ = clock64(); ph[k] += t_ - t_prev; t_prev = t_; }
   long long t_prev = clock64();
#else
#define SBQ_TICK(k)
#endif
   if (kept_all == 0) {
      status = LOCUS_NO_ROWS;
   } else {
      for (int it = 0; it < p.max_iter; ++it) {
         iters = it + 1;
         double* pb = part + (size_t)(it & 1) * (T + 4);
         int zero = 0;
         if (resident) {
            resident_e_pass<NT>(S, th, lpr, zero);
            SBQ_TICK(0)
            zero = __syncthreads_or(zero);
            SBQ_TICK(1)
            resident_col_pass<NT>(S, th, pb, lpc);
            SBQ_TICK(2)
         } else {
            if (g32 < G) cluster_em_pass<CL_LPR_STREAM>(grows, nrows, G, g32, lg32, th, my_acc, zero);
            zero = __syncthreads_or(zero);
            for (int j = tid; j < T; j += NT) {
               double sj = 0.0;
               for (int gg = 0; gg < G; ++gg) { sj += acc[(size_t)gg * T + j]; acc[(size_t)gg * T + j] = 0.0; }
               pb[j] = sj;
            }
         }
         if (tid == 0) pb[T] = (double)zero;
         cluster.sync();
         SBQ_TICK(3)
         double zf = 0.0;
         for (unsigned r = 0; r < CS; ++r) {
            const double* rpb = CS > 1 ? cluster.map_shared_rank(pb, r) : pb;
            zf += rpb[T];
         }
         for (int j = tid; j < T; j += NT) {
            double nj = 0.0;
            for (unsigned r = 0; r < CS; ++r) {
               const double* rpb = CS > 1 ? cluster.map_shared_rank(pb, r) : pb;
               nj += rpb[j];
            }
            nxt[j] = nj;
         }
         SBQ_TICK(4)
         __syncthreads();
         SBQ_TICK(5)
         // every warp forms the same ||theta' - theta||^2 (same order in every warp and every CTA)
         double d2 = 0.0;
         for (int j = lane; j < T; j += 32) { const double diff = nxt[j] - cur[j]; d2 += diff * diff; }
         d2 = warp_sum(d2);
         if (zf != 0.0) { status = LOCUS_ZERO_DENOM; break; }
         if (d2 < tol2) { status = LOCUS_OK; break; }           // ||theta' - theta||_2 < tol; theta is NOT advanced
         for (int j = tid; j < T; j += NT) {
            const double sj = sdiv[j];
            th[j] = (sj != 0) ? nxt[j] / sj : 0.0;
         }
         { double* t_ = cur; cur = nxt; nxt = t_; }
         SBQ_TICK(6)
         __syncthreads();
         SBQ_TICK(7)
      }
   }
#ifdef SBQ_PHASE_TIMING
   if (rank == 0 && (tid == 0 || tid == NT - 32))
      printf("locus %d tid %d iters %d resident %d csc_smem %d lpr %d lpc %d nrows %d nnz %u | E %lld or %lld col %lld csync %lld comb %lld sync %lld upd %lld sync %lld (cycles/iter)\n", l, tid, iters,
             (int)resident, (int)csc_in_smem, lpr, lpc, nrows, nnz_c, ph[0] / iters, ph[1] / iters, ph[2] / iters, ph[3] / iters, ph[4] / iters, ph[5] / iters, ph[6] / iters, ph[7] / iters);
#endif
   cluster.sync();   // no CTA may exit while a peer can still read its shared memory

   if (rank != 0) return;
   // ---- outputs + epilogue (src/estimate.cpp:310-356), CTA 0 of the cluster
   const bool uniform = status == LOCUS_ZERO_DENOM || status == LOCUS_NO_ROWS;
   double fsum = 0.0;
   for (int j = tid; j < T; j += NT) {
      const double tj = uniform ? theta0 : cur[j];
      bool na = false;
      double f = 0.0;
      if (status != LOCUS_NO_ROWS) f = iso_fpkm(p, tj, p.iso_len[t0 + j], na);
      p.theta[t0 + j] = tj;
      p.fpkm[t0 + j] = f;
      th[j] = na ? -1.0 : 0.0;   // remember NA
      fsum += f;
   }
   fsum = block_sum<NT>(fsum, red);
   double ksum = 0.0;
   for (int j = tid; j < T; j += NT) {
      const bool na = th[j] < 0;
      const double f = p.fpkm[t0 + j];
      double fr = 0.0;
      int kp = 0;
      if (status != LOCUS_NO_ROWS) {
         if (!na) fr = f / fsum;
         kp = !(fr < p.min_frac) ? (na ? -1 : 1) : 0;
      }
      p.frac[t0 + j] = fr;
      p.keep[t0 + j] = kp;
      if (kp != 0) ksum += f;
   }
   ksum = block_sum<NT>(ksum, red);
   if (tid == 0) {
      p.iters[l] = iters;
      p.status[l] = status;
      p.locus_fpkm[l] = ksum;
   }
}

// --------------------------------------------------------------------------------------------
// TPM denominator and TPM (src/alignments.cpp:1821-1829)
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) fpkm_sum_kernel(const double* __restrict__ locus_fpkm, int64_t n, double* out) {
   __shared__ double red[32];
   double v = 0.0;
   for (int64_t i = threadIdx.x; i < n; i += 1024) v += locus_fpkm[i];
   v = block_sum<1024>(v, red);
   if (threadIdx.x == 0) *out = v;
}

__global__ void tpm_kernel(const double* __restrict__ fpkm, double* __restrict__ tpm, int64_t n, double total) {
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) tpm[i] = 1e6 * fpkm[i] / total;
}

}  // namespace sbq
