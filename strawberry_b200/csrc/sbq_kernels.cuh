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

__global__ void __launch_bounds__(WT_WARPS * 32)
em_warp_kernel(DevParams p, const int32_t* __restrict__ list, int n_list, int max_iso) {
   extern __shared__ double smem[];
   const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const int w = blockIdx.x * WT_WARPS + warp;
   if (w >= n_list) return;   // whole warps leave; the kernel has no CTA-wide barrier
   const int l = list[w];
   const int64_t r0 = p.loc_row_off[l];
   const int R = (int)(p.loc_row_off[l + 1] - r0);
   const int64_t t0 = p.loc_iso_off[l];
   const int T = (int)(p.loc_iso_off[l + 1] - t0);
   double* acc = smem + (size_t)warp * ((size_t)max_iso * WT_STRIDE + WT_MAX_ISO);
   double* th = acc + (size_t)max_iso * WT_STRIDE;

   const int64_t* __restrict__ rp = p.row_ptr + r0;
   const int32_t* __restrict__ col = p.col;
   const double* __restrict__ al = p.alpha;
   int32_t* neff = p.neff + r0;

   for (int x = lane; x < T * WT_STRIDE; x += 32) acc[x] = 0.0;
   __syncwarp();

   // ---- setup: total, row filter, column sums over kept rows
   long long tot = 0;
   int kept = 0;
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
   }
   tot = warp_sum_ll(tot);
   kept = (int)warp_sum_ll(kept);
   __syncwarp();
   const int nl = R < 32 ? R : 32;   // lanes that own at least one row
   double s = 0.0;
   if (lane < T) {
      for (int x = 0; x < nl; ++x) { s += acc[lane * WT_STRIDE + x]; acc[lane * WT_STRIDE + x] = 0.0; }
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
         if (__any_sync(0xffffffffu, zero)) { status = LOCUS_ZERO_DENOM; break; }
         __syncwarp();
         double nw = 0.0;
         if (lane < T) {
            for (int x = 0; x < nl; ++x) { nw += acc[lane * WT_STRIDE + x]; acc[lane * WT_STRIDE + x] = 0.0; }
         }
         const double diff = lane < T ? nw - cur : 0.0;
         const double d2 = warp_sum(diff * diff);
         if (sqrt(d2) < p.tol) { status = LOCUS_OK; break; }
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
}

// --------------------------------------------------------------------------------------------
// Tier 2: one thread-block CLUSTER per locus (cluster size 1, 2, 4, 8 or 16 chosen by nnz).
// Rows are split over the CTAs of the cluster by non-zero count; inside a CTA, groups of LPR lanes
// own one row at a time and a private accumulator row acc[g][T] in shared memory. Per iteration:
// row pass -> fixed-order sum over groups -> partial theta' exchanged through distributed shared
// memory (one cluster barrier, double-buffered) -> every CTA forms the same theta' and norm.
// --------------------------------------------------------------------------------------------
template <int LPR, int NT>
__host__ __device__ inline int cluster_groups_for(int T, size_t smem_bytes) {
   // accumulators may take at most half of the dynamic shared memory; the rest is for the resident CSR slice
   const long long budget = (long long)(smem_bytes / 2 / sizeof(double)) - 5LL * T - 8;
   long long G = budget / (T > 0 ? T : 1);
   const int per_warp = 32 / LPR;
   if (G > NT / LPR) G = NT / LPR;
   G = (G / per_warp) * per_warp;
   if (G < per_warp) {   // large T: let the accumulators use all of it (streaming mode)
      const long long full = (long long)(smem_bytes / sizeof(double)) - 5LL * T - 8;
      G = full / (T > 0 ? T : 1);
      if (G > NT / LPR) G = NT / LPR;
      G = (G / per_warp) * per_warp;
   }
   return (int)G;   // 0 => does not fit
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

// Row accessors of one CTA's slice: streamed from global/L2, or resident in shared memory.
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
struct SmemRows {
   const unsigned* rp;
   const double* al;
   const unsigned short* col;
   int* ne;
   __device__ __forceinline__ unsigned start(int i) const { return rp[i]; }
   __device__ __forceinline__ double a(unsigned k) const { return al[k]; }
   __device__ __forceinline__ int c(unsigned k) const { return col[k]; }
   __device__ __forceinline__ int get_ne(int i) const { return ne[i]; }
   __device__ __forceinline__ void set_ne(int i, int v) const { ne[i] = v; }
};

// Setup pass over the CTA's rows: total count, row filter (-> ne), column sums of the kept rows.
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
   }
}

// One E/M pass over the CTA's rows with the scaled theta in th[]; two rows per group in flight.
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
      if (ne1 >= 0) {
         if (d1 == 0) zero = 1;
         else {
            const double r = (double)ne1 / d1;
            for (unsigned k = a1 + lg; k < b1; k += LPR) { const int c = rows.c(k); my_acc[c] += rows.a(k) * th[c] * r; }
         }
      }
   }
}

template <int LPR, int NT>
__global__ void __launch_bounds__(NT)
em_cluster_kernel(DevParams p, const int32_t* __restrict__ list, int n_list, unsigned smem_bytes) {
   cg::cluster_group cluster = cg::this_cluster();
   const unsigned CS = cluster.num_blocks();
   const unsigned rank = cluster.block_rank();
   const int item = blockIdx.x / CS;
   const int l = list[item];
   const int tid = threadIdx.x;

   const int64_t r0 = p.loc_row_off[l];
   const int R = (int)(p.loc_row_off[l + 1] - r0);
   const int64_t t0 = p.loc_iso_off[l];
   const int T = (int)(p.loc_iso_off[l + 1] - t0);
   const int G = cluster_groups_for<LPR, NT>(T, smem_bytes);

   extern __shared__ double smem[];
   double* th = smem;                  // [T] scaled theta used by the row pass
   double* cur = th + T;               // [T] theta
   double* sdiv = cur + T;             // [T] column sums s_j of the kept rows
   double* part = sdiv + T;            // [2][T+4] exchange buffers: partial theta', flag, total, kept
   double* acc = part + 2 * (T + 4);   // [G][T] group-private accumulators
   double* res = acc + (size_t)G * T;  // resident CSR slice (if it fits)
   __shared__ double red[NT / 32];
   __shared__ int s_rows[2];

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
   for (int x = tid; x < G * T; x += NT) acc[x] = 0.0;
   __syncthreads();
   const int ra = s_rows[0], rb = s_rows[1], nrows = rb - ra;
   const int64_t k_a = rp[ra];
   const unsigned nnz_c = (unsigned)(rp[rb] - k_a);
   const int g = tid / LPR, lg = tid % LPR;
   const bool active = g < G;
   double* my_acc = acc + (size_t)(active ? g : 0) * T;

   // resident slice: alpha f64 | row starts u32 | ne i32 | col u16
   const size_t used = (size_t)((char*)res - (char*)smem);
   const size_t need = (size_t)nnz_c * 8 + ((size_t)nrows + 1) * 4 + (size_t)nrows * 4 + (size_t)nnz_c * 2 + 32;
   const bool resident = used + need <= smem_bytes;
   double* s_al = res;
   unsigned* s_rp = (unsigned*)(s_al + nnz_c);
   int* s_ne = (int*)(s_rp + nrows + 1);
   unsigned short* s_col = (unsigned short*)(s_ne + nrows);
   GlobalRows grows{rp + ra, p.alpha + k_a, p.col + k_a, p.neff + r0 + ra, k_a};
   SmemRows srows{s_rp, s_al, s_col, s_ne};
   if (resident) {
      for (unsigned k = tid; k < nnz_c; k += NT) { s_al[k] = grows.al[k]; s_col[k] = (unsigned short)grows.col[k]; }
      for (int i = tid; i <= nrows; i += NT) s_rp[i] = (unsigned)(rp[ra + i] - k_a);
      __syncthreads();
   }

   // ---- setup pass: total, row filter, column sums
   long long tot = 0;
   int kept = 0;
   if (active) {
      if (resident) cluster_setup_pass<LPR>(srows, p.count + r0 + ra, nrows, G, g, lg, p.row_eps, my_acc, tot, kept);
      else cluster_setup_pass<LPR>(grows, p.count + r0 + ra, nrows, G, g, lg, p.row_eps, my_acc, tot, kept);
   }
   __syncthreads();
   {
      const double tot_b = block_sum<NT>((double)tot, red);
      const double kept_b = block_sum<NT>((double)kept, red);
      for (int j = tid; j < T; j += NT) {
         double sj = 0.0;
         for (int gg = 0; gg < G; ++gg) { sj += acc[(size_t)gg * T + j]; acc[(size_t)gg * T + j] = 0.0; }
         part[j] = sj;
      }
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

   int status = LOCUS_ITER_CAP, iters = 0;
   if (kept_all == 0) {
      status = LOCUS_NO_ROWS;
   } else {
      for (int it = 0; it < p.max_iter; ++it) {
         iters = it + 1;
         double* pb = part + (size_t)(it & 1) * (T + 4);
         int zero = 0;
         if (active) {
            if (resident) cluster_em_pass<LPR>(srows, nrows, G, g, lg, th, my_acc, zero);
            else cluster_em_pass<LPR>(grows, nrows, G, g, lg, th, my_acc, zero);
         }
         zero = __syncthreads_or(zero);
         for (int j = tid; j < T; j += NT) {
            double sj = 0.0;
            for (int gg = 0; gg < G; ++gg) { sj += acc[(size_t)gg * T + j]; acc[(size_t)gg * T + j] = 0.0; }
            pb[j] = sj;
         }
         if (tid == 0) pb[T] = (double)zero;
         cluster.sync();
         double zf = 0.0, d2 = 0.0;
         for (unsigned r = 0; r < CS; ++r) {
            const double* rpb = CS > 1 ? cluster.map_shared_rank(pb, r) : pb;
            zf += rpb[T];
         }
         // theta'_j, kept in th[] until we know whether to advance (th is dead after the row pass)
         for (int j = tid; j < T; j += NT) {
            double nj = 0.0;
            for (unsigned r = 0; r < CS; ++r) {
               const double* rpb = CS > 1 ? cluster.map_shared_rank(pb, r) : pb;
               nj += rpb[j];
            }
            const double diff = nj - cur[j];
            d2 += diff * diff;
            th[j] = nj;
         }
         d2 = block_sum<NT>(d2, red);
         if (zf != 0.0) { status = LOCUS_ZERO_DENOM; break; }
         if (sqrt(d2) < p.tol) { status = LOCUS_OK; break; }
         for (int j = tid; j < T; j += NT) {
            const double nj = th[j];
            cur[j] = nj;
            const double sj = sdiv[j];
            th[j] = (sj != 0) ? nj / sj : 0.0;
         }
         __syncthreads();
      }
   }
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
