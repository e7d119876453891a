// sbq_grid_tma.cuh - Tier 3, main path: multi-CTA streaming EM for giant loci with a TMA bulk-copy pipeline.
//
// Same algorithm and reduction structure as em_grid_kernel (sbq_grid.cuh), different data movement:
//  * one persistent CTA per SM = 16 consumer warps + 1 producer warp;
//  * the CTA's contiguous row range is cut into chunks of G4_ROWS rows; the producer lane streams every chunk
//    into a G4_NS-deep shared-memory ring with 1-D bulk copies (cp.async.bulk ... mbarrier::complete_tx::bytes,
//    SASS UBLKCP): alpha slab (8 B/nz), 16-bit column slab (2 B/nz), row-pointer slab, count slab. The rows of a
//    CTA are contiguous in CSR, so each slab is one contiguous, 16-byte aligned piece of HBM;
//  * full/empty mbarriers per stage; consumers wait (try_wait.parity), take two rows of the chunk per warp,
//    and release the stage. HBM latency is covered by the ring depth instead of by registers;
//  * columns are pre-narrowed to u16 by cols_to_u16_kernel at upload (10 B of HBM traffic per non-zero and
//    iteration instead of 12; the algorithmic-byte accounting of SURVEY 8d stays 12).
// Reductions (warp-private accumulators -> per-CTA partial -> column owners -> theta') are unchanged and
// deterministic. Chunks whose non-zeros exceed the stage capacity are processed straight from global memory
// by the consumers (both sides skip the ring for them).
#pragma once
#include "sbq_grid.cuh"

namespace sbq {

constexpr int G4_ROWS = 32;                  // rows per chunk (two per consumer warp)
constexpr int G4_CAP = 2048;                 // non-zeros a stage can hold
constexpr int G4_NS = 4;                     // ring depth
constexpr int G4_CONSUMERS = 16;             // consumer warps, one warp-private accumulator row each
constexpr int G4_ACC = G4_CONSUMERS;         // accumulator rows
constexpr int G4_EPL = 2;                    // elements per lane per row on the register path (rows up to 64 non-zeros)
constexpr int G4_NT = (G4_CONSUMERS + 1) * 32;
constexpr int G4_MAX_CHUNKS = 2048;          // chunk table entries per CTA
// stage layout (bytes): alpha | col16 | row pointers | counts
constexpr int G4_A_BYTES = (G4_CAP + 2) * 8;             // 16400 -> 16-byte multiple
constexpr int G4_C_OFF = G4_A_BYTES;
constexpr int G4_C_BYTES = (G4_CAP + 8) * 2 + 16;        // 4128
constexpr int G4_R_OFF = G4_C_OFF + G4_C_BYTES;
constexpr int G4_R_BYTES = 36 * 8;                       // up to 35 row pointers
constexpr int G4_N_OFF = G4_R_OFF + G4_R_BYTES;
constexpr int G4_N_BYTES = 40 * 4;                       // up to 38 counts
constexpr int G4_STAGE_BYTES = ((G4_N_OFF + G4_N_BYTES + 127) / 128) * 128;
constexpr int G4_CHUNK_BYTES = (((G4_MAX_CHUNKS + 1) * 4 + 127) / 128) * 128;
static_assert(G4_A_BYTES % 16 == 0 && G4_C_OFF % 16 == 0 && G4_R_OFF % 16 == 0 && G4_N_OFF % 16 == 0, "bulk copies need 16-byte alignment");

__host__ __device__ inline size_t grid_tma_smem_bytes(int T) {
   // th[T] | acc[16][T] | chunk table | ring
   return ((size_t)T * (1 + G4_ACC)) * sizeof(double) + (size_t)G4_CHUNK_BYTES + (size_t)G4_NS * G4_STAGE_BYTES;
}
inline bool grid_tma_supports(int T, long long rows, int n_cta) {
   return grid_tma_smem_bytes(T) <= 225 * 1024 && T <= 65535 && rows / n_cta + G4_ROWS < (long long)G4_MAX_CHUNKS * G4_ROWS;
}

__global__ void cols_to_u16_kernel(const int32_t* __restrict__ col, unsigned short* __restrict__ out, int64_t n) {
   for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (unsigned short)col[i];
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
   asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
   asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                "r"(bytes), "r"(smem_u32(bar))
                : "memory");
}

struct G4Ring {
   char* stage;          // G4_NS * G4_STAGE_BYTES
   uint64_t* full;       // [G4_NS]
   uint64_t* empty;      // [G4_NS]
};

// Producer: stream the chunks [0, n_chunk) of this CTA's rows into the ring. `use` counts ring uses across
// passes so that the mbarrier parities stay in step with the consumers.
__device__ __forceinline__ void g4_produce(const DevParams& p, const unsigned short* __restrict__ col16, const int32_t* __restrict__ cnt_or_neff,
                                           int64_t row_base /* absolute row of local row 0 */, int64_t kb, const unsigned* s_chunk, int ra, int rb,
                                           int n_chunk, const G4Ring& ring, unsigned& use) {
   asm volatile("fence.proxy.async;" ::: "memory");   // neff written with ordinary stores in the setup pass is read by bulk copies
   for (int c = 0; c < n_chunk; ++c) {
      const unsigned ck0 = s_chunk[c], ck1 = s_chunk[c + 1];
      if (ck1 - ck0 > (unsigned)G4_CAP) continue;          // oversize chunk: consumers read it from global memory
      const int s = use % G4_NS;
      const unsigned n_use = use / G4_NS;
      mbar_wait(&ring.empty[s], (n_use & 1u) ^ 1u);
      char* st = ring.stage + (size_t)s * G4_STAGE_BYTES;
      const int i0 = ra + c * G4_ROWS, i1 = min(i0 + G4_ROWS, rb);
      const int64_t k0 = kb + ck0, k1 = kb + ck1;
      const int64_t ka = k0 & ~(int64_t)1, kc = k0 & ~(int64_t)7;
      const unsigned a_bytes = (unsigned)(((k1 - ka + 1) & ~(int64_t)1) * 8);
      const unsigned c_bytes = (unsigned)(((k1 - kc + 7) & ~(int64_t)7) * 2);
      const int64_t g0 = row_base + i0, g1 = row_base + i1;   // absolute rows; need row_ptr[g0 .. g1]
      const int64_t gr = g0 & ~(int64_t)1, gn = g0 & ~(int64_t)3;
      const unsigned r_bytes = (unsigned)(((g1 + 1 - gr + 1) & ~(int64_t)1) * 8);
      const unsigned n_bytes = (unsigned)(((g1 - gn + 3) & ~(int64_t)3) * 4);
      mbar_expect_tx(&ring.full[s], a_bytes + c_bytes + r_bytes + n_bytes);
      if (a_bytes) bulk_g2s(st, p.alpha + ka, a_bytes, &ring.full[s]);
      if (c_bytes) bulk_g2s(st + G4_C_OFF, col16 + kc, c_bytes, &ring.full[s]);
      bulk_g2s(st + G4_R_OFF, p.row_ptr + gr, r_bytes, &ring.full[s]);
      if (n_bytes) bulk_g2s(st + G4_N_OFF, cnt_or_neff + gn, n_bytes, &ring.full[s]);
      ++use;
   }
}

// Consumer side of one pass over the CTA's rows. SETUP: row filter + column sums (+ total / kept counts, neff
// written to global). !SETUP: one E/M pass with the scaled theta in th[].
template <bool SETUP>
__device__ __forceinline__ void g4_consume(const DevParams& p, const unsigned short* __restrict__ col16, int32_t* neff_glob /* local row 0 */,
                                           const int32_t* __restrict__ cnt_glob, const int64_t* __restrict__ rp_loc, int64_t row_base, int64_t kb,
                                           const unsigned* s_chunk, int ra, int rb, int n_chunk, const G4Ring& ring, unsigned& use,
                                           const double* th, double* my_half /* this half-warp's accumulator row */, int acc_stride,
                                           long long& tot, long long& kept, int& zero) {
   const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
   for (int c = 0; c < n_chunk; ++c) {
      const unsigned ck0 = s_chunk[c], ck1 = s_chunk[c + 1];
      const int i0 = ra + c * G4_ROWS, i1 = min(i0 + G4_ROWS, rb);
      if (ck1 - ck0 > (unsigned)G4_CAP) {
         // oversize chunk, generic path straight from global memory (whole warp per row, half-0 accumulator row)
         double* my = my_half;
         for (int q = warp; i0 + q < i1; q += G4_CONSUMERS) {
            const int i = i0 + q;
            const int64_t a = rp_loc[i], b = rp_loc[i + 1];
            if (SETUP) {
               const int n = cnt_glob[i];
               bool keep = false;
               for (int64_t k = a + lane; k < b; k += 32) keep |= p.alpha[k] > p.row_eps;
               keep = __any_sync(0xffffffffu, keep);
               if (lane == 0) { neff_glob[i] = keep ? n : -1; tot += n; kept += keep; }
               if (keep)
                  for (int64_t k = a + lane; k < b; k += 32) my[col16[k]] += p.alpha[k];
            } else {
               const int ne = neff_glob[i];
               if (ne < 0) continue;
               double d = 0.0;
               for (int64_t k = a + lane; k < b; k += 32) d += p.alpha[k] * th[col16[k]];
               d = warp_sum(d);
               if (d == 0) { zero = 1; continue; }
               const double rr = (double)ne / d;
               for (int64_t k = a + lane; k < b; k += 32) { const int cc = col16[k]; my[cc] += p.alpha[k] * th[cc] * rr; }
            }
         }
         continue;
      }
      double* my = my_half;
      const int s = use % G4_NS;
      const unsigned n_use = use / G4_NS;
      mbar_wait(&ring.full[s], n_use & 1u);
      const char* st = ring.stage + (size_t)s * G4_STAGE_BYTES;
      const int64_t k0 = kb + ck0;
      const double* a_s = (const double*)st + (k0 & 1);
      const unsigned short* c_s = (const unsigned short*)(st + G4_C_OFF) + (k0 & 7);
      const int64_t g0 = row_base + i0;
      const int64_t* r_s = (const int64_t*)(st + G4_R_OFF) + (g0 & 1);
      const int* n_s = (const int*)(st + G4_N_OFF) + (g0 & 3);
      const int nrow = i1 - i0;
      // Two rows per warp (rows warp and warp + 16 of the chunk), each spread over all 32 lanes with up to G4_EPL
      // elements per lane held in registers between the normaliser and the accumulation. A row's columns are
      // distinct, so the 32 lanes never collide in the warp-private accumulator row.
      double a[2][G4_EPL];
      int cc[2][G4_EPL], ne[2];
      unsigned rs[2], re[2];
      bool longrow = false;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
         const int rq = warp + q * G4_CONSUMERS;
         const bool vr = rq < nrow;
         rs[q] = vr ? (unsigned)(r_s[rq] - k0) : 0u;
         re[q] = vr ? (unsigned)(r_s[rq + 1] - k0) : 0u;
         ne[q] = vr ? n_s[rq] : -1;
         if (!SETUP && ne[q] < 0) re[q] = rs[q];
         longrow |= re[q] - rs[q] > 32u * G4_EPL;
#pragma unroll
         for (int e = 0; e < G4_EPL; ++e) {
            const unsigned k = rs[q] + lane + 32 * e;
            const bool v = k < re[q];
            a[q][e] = v ? a_s[k] : 0.0;
            cc[q][e] = v ? (int)c_s[k] : 0;
         }
      }
      if (!longrow) {
         if (SETUP) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
               const bool keep = __any_sync(0xffffffffu, a[q][0] > p.row_eps || a[q][1] > p.row_eps);
               const int rq = warp + q * G4_CONSUMERS;
               if (rq < nrow) {
                  if (lane == 0) { neff_glob[i0 + rq] = keep ? ne[q] : -1; tot += ne[q]; kept += keep; }
                  if (keep) {
                     if (a[q][0] != 0.0) my[cc[q][0]] += a[q][0];
                     if (a[q][1] != 0.0) my[cc[q][1]] += a[q][1];
                  }
               }
            }
         } else {
            double t[2][G4_EPL], d[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
               t[q][0] = th[cc[q][0]];
               t[q][1] = th[cc[q][1]];
               d[q] = a[q][0] * t[q][0] + a[q][1] * t[q][1];
            }
            // both normalisers in five shuffles: swap halves (lanes 0-15 collect row 0, lanes 16-31 row 1), then a
            // 4-step butterfly inside each half; lanes 0 and 16 divide; two shuffles broadcast r_0 and r_1
            const bool hi = lane >= 16;
            double v = (hi ? d[1] : d[0]) + __shfl_xor_sync(0xffffffffu, hi ? d[0] : d[1], 16);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            double rmine = 0.0;
            if ((lane & 15) == 0) {
               const int q = lane >> 4;
               if (re[q] > rs[q]) {
                  if (v == 0) zero = 1; else rmine = (double)ne[q] / v;
               }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
               const double rr = __shfl_sync(0xffffffffu, rmine, 16 * q);
               if (a[q][0] != 0.0) my[cc[q][0]] += a[q][0] * t[q][0] * rr;
               if (a[q][1] != 0.0) my[cc[q][1]] += a[q][1] * t[q][1] * rr;
            }
         }
      } else {
         // a row longer than 64 non-zeros inside a staged chunk: loop over it in shared memory
         for (int q = 0; q < 2; ++q) {
            const int rq = warp + q * G4_CONSUMERS;
            if (rq >= nrow) break;
            if (SETUP) {
               bool keep = false;
               for (unsigned k = rs[q] + lane; k < re[q]; k += 32) keep |= a_s[k] > p.row_eps;
               keep = __any_sync(0xffffffffu, keep);
               if (lane == 0) { neff_glob[i0 + rq] = keep ? ne[q] : -1; tot += ne[q]; kept += keep; }
               if (keep)
                  for (unsigned k = rs[q] + lane; k < re[q]; k += 32) my[c_s[k]] += a_s[k];
            } else {
               if (re[q] == rs[q]) continue;
               double d = 0.0;
               for (unsigned k = rs[q] + lane; k < re[q]; k += 32) d += a_s[k] * th[c_s[k]];
               d = warp_sum(d);
               if (d == 0) { zero = 1; continue; }
               const double rr = (double)ne[q] / d;
               for (unsigned k = rs[q] + lane; k < re[q]; k += 32) { const int c2 = c_s[k]; my[c2] += a_s[k] * th[c2] * rr; }
            }
            __syncwarp();
         }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&ring.empty[s]);
      ++use;
   }
}

__global__ void __launch_bounds__(G4_NT, 1)
em_grid_tma_kernel(DevParams p, const unsigned short* __restrict__ col16, const int32_t* __restrict__ list, int n_list, GridScratch gs,
                   double* cur_glob /* [n_cta][tstride] */) {
   cg::grid_group grid = cg::this_grid();
   extern __shared__ __align__(128) unsigned char g4_smem[];
   __shared__ double red[G4_NT / 32];
   __shared__ int s_rows[2];
   __shared__ __align__(8) uint64_t s_bar[2 * G4_NS];
   const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
   const int nb = gridDim.x, b = blockIdx.x;
   const bool producer = warp == G4_CONSUMERS;

   G4Ring ring;
   ring.stage = (char*)g4_smem;
   ring.full = s_bar;
   ring.empty = s_bar + G4_NS;
   if (tid == 0) {
      for (int s = 0; s < G4_NS; ++s) { mbar_init(&ring.full[s], 1); mbar_init(&ring.empty[s], G4_CONSUMERS); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   __syncthreads();
   unsigned use = 0;   // ring uses so far (same sequence on the producer and on every consumer warp)
   double* my_cur = cur_glob + (size_t)b * gs.tstride;   // this CTA's copy of theta (L2-resident, touched once per iteration)

   for (int item = 0; item < n_list; ++item) {
      const int l = list[item];
      const int64_t r0 = p.loc_row_off[l];
      const int R = (int)(p.loc_row_off[l + 1] - r0);
      const int64_t t0 = p.loc_iso_off[l];
      const int T = (int)(p.loc_iso_off[l + 1] - t0);
      unsigned* s_chunk = (unsigned*)(g4_smem + (size_t)G4_NS * G4_STAGE_BYTES);
      double* th = (double*)(g4_smem + (size_t)G4_NS * G4_STAGE_BYTES + G4_CHUNK_BYTES);
      double* acc = th + T;   // [G4_ACC][T]
      const int64_t* __restrict__ rp = p.row_ptr + r0;
      int32_t* neff = p.neff + r0;
      const int32_t* cnt = p.count + r0;
      double* my_partial = gs.partial + (size_t)b * gs.tstride;

      if (tid < 2) {
         const int64_t base = rp[0], nnz = rp[R] - base;
         const int64_t target = base + (nnz * (int64_t)(b + tid)) / nb;
         int lo = 0, hi = R;
         if (b + tid >= nb) lo = R;
         else if (b + tid == 0) hi = 0;
         while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (rp[mid] < target) lo = mid + 1; else hi = mid;
         }
         s_rows[tid] = lo;
      }
      for (int x = tid; x < G4_ACC * T; x += G4_NT) acc[x] = 0.0;
      __syncthreads();
      const int ra = s_rows[0], rb = s_rows[1];
      const int n_chunk = (rb - ra + G4_ROWS - 1) / G4_ROWS;
      const int64_t kb = rp[ra];
      for (int c = tid; c <= n_chunk; c += G4_NT) s_chunk[c] = (unsigned)(rp[min(ra + c * G4_ROWS, rb)] - kb);
      __syncthreads();
      double* my_acc = acc + (size_t)(producer ? 0 : warp) * T;

      // ---- setup pass
      long long tot = 0, kept = 0;
      int zero = 0;
      if (producer) {
         if (lane == 0) g4_produce(p, col16, p.count, r0, kb, s_chunk, ra, rb, n_chunk, ring, use);
         use = __shfl_sync(0xffffffffu, use, 0);
      } else {
         g4_consume<true>(p, col16, neff, cnt, rp, r0, kb, s_chunk, ra, rb, n_chunk, ring, use, th, my_acc, T, tot, kept, zero);
      }
      tot = warp_sum_ll(tot);
      kept = warp_sum_ll(kept);
      if (lane == 0 && (tot | kept)) {
         atomicAdd((unsigned long long*)&gs.ctr[2 * item], (unsigned long long)tot);
         atomicAdd((unsigned long long*)&gs.ctr[2 * item + 1], (unsigned long long)kept);
      }
      __threadfence();
      asm volatile("fence.proxy.async;" ::: "memory");
      __syncthreads();
      for (int j = tid; j < T; j += G4_NT) {
         double sj = 0.0;
         for (int w = 0; w < G4_ACC; ++w) { sj += acc[(size_t)w * T + j]; acc[(size_t)w * T + j] = 0.0; }
         my_partial[j] = sj;
      }
      grid.sync();
      for (int j = b * (G4_NT / 32) + warp; j < T; j += nb * (G4_NT / 32)) {
         double sj = 0.0;
         for (int cta = lane; cta < nb; cta += 32) sj += __ldcg(gs.partial + (size_t)cta * gs.tstride + j);
         sj = warp_sum(sj);
         if (lane == 0) gs.theta_next[j] = sj;
      }
      grid.sync();
      const double total = (double)__ldcg(gs.ctr + 2 * item);
      const long long kept_all = __ldcg(gs.ctr + 2 * item + 1);
      const double theta0 = total / (double)T;
      // s_j is kept in the upper half of this CTA's theta copy: my_cur[j] = theta_j, my_cur[tstride/2 + j] = s_j
      double* my_sdiv = my_cur + gs.tstride / 2;
      for (int j = tid; j < T; j += G4_NT) {
         my_sdiv[j] = __ldcg(gs.theta_next + j);
         my_cur[j] = theta0;
         th[j] = theta0;
      }
      grid.sync();   // theta_next is rewritten in iteration 0 only after everyone copied s_j out

      const double tol2 = p.tol * p.tol;
      int status = LOCUS_ITER_CAP, iters = 0;
      if (kept_all == 0) {
         status = LOCUS_NO_ROWS;
      } else {
         for (int it = 0; it < p.max_iter; ++it) {
            iters = it + 1;
            zero = 0;
            long long d0 = 0, d1 = 0;
            if (producer) {
               if (lane == 0) g4_produce(p, col16, p.neff, r0, kb, s_chunk, ra, rb, n_chunk, ring, use);
               use = __shfl_sync(0xffffffffu, use, 0);
            } else {
               g4_consume<false>(p, col16, neff, cnt, rp, r0, kb, s_chunk, ra, rb, n_chunk, ring, use, th, my_acc, T, d0, d1, zero);
            }
            zero = __syncthreads_or(zero);
            if (zero && tid == 0) atomicOr(&gs.zero_flag[item], 1);
            for (int j = tid; j < T; j += G4_NT) {
               double sj = 0.0;
               for (int w = 0; w < G4_ACC; ++w) { sj += acc[(size_t)w * T + j]; acc[(size_t)w * T + j] = 0.0; }
               my_partial[j] = sj;
            }
            grid.sync();
            for (int j = b * (G4_NT / 32) + warp; j < T; j += nb * (G4_NT / 32)) {
               double sj = 0.0;
               for (int cta = lane; cta < nb; cta += 32) sj += __ldcg(gs.partial + (size_t)cta * gs.tstride + j);
               sj = warp_sum(sj);
               if (lane == 0) gs.theta_next[j] = sj;
            }
            grid.sync();
            const int zf = *(volatile int*)&gs.zero_flag[item];
            double d2 = 0.0;
            for (int j = tid; j < T; j += G4_NT) {
               const double nj = __ldcg(gs.theta_next + j);
               const double diff = nj - my_cur[j];
               d2 += diff * diff;
               th[j] = nj;
            }
            d2 = block_sum<G4_NT>(d2, red);
            if (zf) { status = LOCUS_ZERO_DENOM; break; }
            if (d2 < tol2) { status = LOCUS_OK; break; }
            for (int j = tid; j < T; j += G4_NT) {
               const double nj = th[j];
               my_cur[j] = nj;
               const double sj = my_sdiv[j];
               th[j] = (sj != 0) ? nj / sj : 0.0;
            }
            __syncthreads();
         }
      }

      // ---- outputs + epilogue by CTA 0 (src/estimate.cpp:310-356)
      if (b == 0) {
         const bool uniform = status == LOCUS_ZERO_DENOM || status == LOCUS_NO_ROWS;
         double fsum = 0.0;
         for (int j = tid; j < T; j += G4_NT) {
            const double tj = uniform ? theta0 : my_cur[j];
            bool na = false;
            double f = 0.0;
            if (status != LOCUS_NO_ROWS) f = iso_fpkm(p, tj, p.iso_len[t0 + j], na);
            p.theta[t0 + j] = tj;
            p.fpkm[t0 + j] = f;
            th[j] = na ? -1.0 : 0.0;
            fsum += f;
         }
         fsum = block_sum<G4_NT>(fsum, red);
         double ksum = 0.0;
         for (int j = tid; j < T; j += G4_NT) {
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
         ksum = block_sum<G4_NT>(ksum, red);
         if (tid == 0) {
            p.iters[l] = iters;
            p.status[l] = status;
            p.locus_fpkm[l] = ksum;
         }
      }
      grid.sync();   // scratch (partial, theta_next) is reused by the next locus
   }
}

// Host launcher of the TMA path. col16_scratch: device buffer for the u16 columns of the whole batch (grown here).
inline int grid_tma_launch(const DevParams& dp, int64_t nnz_total, const int32_t* d_list, int n_list, int max_iso, const cudaDeviceProp& prop,
                           void** scratch, size_t* scratch_cap, void** col16_scratch, size_t* col16_cap, bool cols_ready, cudaStream_t st,
                           int* n_launch) {
   *n_launch = 0;
   if (n_list == 0) return 0;
   const size_t smem = grid_tma_smem_bytes(max_iso);
   if (cudaFuncSetAttribute(em_grid_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -3;
   int per_sm = 0;
   if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, em_grid_tma_kernel, G4_NT, smem) != cudaSuccess || per_sm < 1) return -3;
   const int nb = prop.multiProcessorCount;
   const int tstride = 8192;   // theta copy in the lower half, s_j in the upper half
   const size_t need = ((size_t)nb * tstride * 2 + tstride) * sizeof(double) + (size_t)n_list * (2 * sizeof(long long) + sizeof(int)) + 1024;
   if (need > *scratch_cap) {
      if (*scratch) cudaFree(*scratch);
      *scratch = nullptr;
      *scratch_cap = 0;
      if (cudaMalloc(scratch, need) != cudaSuccess) return -4;
      *scratch_cap = need;
   }
   const size_t need16 = (size_t)nnz_total * 2 + 256;
   if (need16 > *col16_cap) {
      if (*col16_scratch) cudaFree(*col16_scratch);
      *col16_scratch = nullptr;
      *col16_cap = 0;
      if (cudaMalloc(col16_scratch, need16) != cudaSuccess) return -4;
      *col16_cap = need16;
      cols_ready = false;
   }
   if (!cols_ready) {
      cols_to_u16_kernel<<<prop.multiProcessorCount * 8, 256, 0, st>>>(dp.col, (unsigned short*)*col16_scratch, nnz_total);
      ++*n_launch;
   }
   GridScratch gs;
   char* q = (char*)*scratch;
   gs.partial = (double*)q; q += (size_t)nb * tstride * sizeof(double);
   double* cur_glob = (double*)q; q += (size_t)nb * tstride * sizeof(double);
   gs.theta_next = (double*)q; q += (size_t)tstride * sizeof(double);
   gs.ctr = (long long*)q; q += (size_t)n_list * 2 * sizeof(long long);
   gs.zero_flag = (int*)q;
   gs.tstride = tstride;
   if (cudaMemsetAsync(gs.ctr, 0, (size_t)n_list * (2 * sizeof(long long) + sizeof(int)), st) != cudaSuccess) return -3;
   DevParams dpc = dp;
   const unsigned short* c16 = (const unsigned short*)*col16_scratch;
   void* args[] = {(void*)&dpc, (void*)&c16, (void*)&d_list, (void*)&n_list, (void*)&gs, (void*)&cur_glob};
   if (cudaLaunchCooperativeKernel((void*)em_grid_tma_kernel, dim3(nb), dim3(G4_NT), args, smem, st) != cudaSuccess) return -3;
   ++*n_launch;
   return 0;
}

}  // namespace sbq
