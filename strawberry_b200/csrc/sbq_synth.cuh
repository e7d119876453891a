// sbq_synth.cuh - giant-locus stress input generated ON THE DEVICE from a seed (BASELINE configs[3]).
//
// 200 loci x 1 M rows x ~48 non-zeros are ~115 GB of CSR: they are written straight into the context's device arrays and
// never cross PCIe. The generator is a pure function of (seed, GLOBAL locus id, row, entry) through a counter-based hash
// (splitmix64), so any partition of the locus ids over devices or waves yields the same loci, and its integer and
// floating-point steps are all exactly reproducible on a CPU: strawberry_b200/synth.py::giant_device restates it in numpy and
// tests/test_gpu_synth.py requires bit-equal row pointers, columns, weights and lengths.
//
//   key(seed, id, s) = sm64(sm64(seed ^ 0x5851F42D4C957F2D) + 4 id + s)          s = 0 (T), 1 (row degree), 2 (entries), 3 (lengths)
//   h(key, x)        = sm64(key + x)                  u(h) = (h >> 11) * 2^-53
//   T                = iso_lo + h(key0, 0) mod (iso_hi - iso_lo + 1)
//   row i            : k = 1 + Poisson(mean_extra) by CDF inversion of u(h(key1, i)), capped at min(T, 255); n_i = 1
//   entry m of row i : v = h(key2, 256 i + m); column = lo + floor(u(v) (hi - lo)), [lo, hi) = [m T / k, (m + 1) T / k)   (stratified: distinct, ascending)
//                      w = sm64(v); alpha = (1 + u(w)) * 2^-(6 + (w & 7))            (piecewise log-uniform on [1.2e-4, 3.1e-2): exact in fp64)
//   iso_len[j]       = 400 + h(key3, j) mod 7601
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sbq {

__host__ __device__ __forceinline__ uint64_t sm64(uint64_t x) {
   x += 0x9E3779B97F4A7C15ull;
   x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
   x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
   return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ uint64_t synth_key(uint64_t seed, uint64_t id, uint64_t stream) { return sm64(sm64(seed ^ 0x5851F42D4C957F2Dull) + 4 * id + stream); }
__host__ __device__ __forceinline__ double synth_u(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

constexpr int SYN_CDF = 255;          // k - 1 in 0..254
constexpr int SYN_SCAN_BLOCK = 4096;  // rows per block of the row-pointer scan

struct SynthLocus {                   // per generated locus (device array)
   uint64_t key_deg, key_ent, key_len;
   int32_t T;
   int32_t pad;
};

// degree of every row (k as a byte) + the grand total
__global__ void __launch_bounds__(256) synth_degree_kernel(const SynthLocus* __restrict__ loci, int n_loci, int64_t rows, const double* __restrict__ cdf_g,
                                                           unsigned char* __restrict__ deg, unsigned long long* total) {
   __shared__ double cdf[SYN_CDF + 1];
   __shared__ unsigned long long s_tot;
   for (int x = threadIdx.x; x < SYN_CDF; x += blockDim.x) cdf[x] = cdf_g[x];
   if (threadIdx.x == 0) s_tot = 0;
   __syncthreads();
   const int64_t n = (int64_t)n_loci * rows;
   unsigned long long mine = 0;
   for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (int64_t)gridDim.x * blockDim.x) {
      const int l = (int)(g / rows);
      const int64_t i = g - (int64_t)l * rows;
      const double u = synth_u(sm64(loci[l].key_deg + (uint64_t)i));
      int lo = 0, hi = SYN_CDF - 1;                      // #{j : cdf[j] <= u}, capped at 254
      while (lo < hi) {
         const int mid = (lo + hi) >> 1;
         if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
      }
      int k = 1 + lo;
      const int cap = loci[l].T < 255 ? loci[l].T : 255;
      if (k > cap) k = cap;
      deg[g] = (unsigned char)k;
      mine += (unsigned long long)k;
   }
   for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
   if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_tot, mine);
   __syncthreads();
   if (threadIdx.x == 0 && s_tot) atomicAdd(total, s_tot);
}

// exclusive scan of the byte degrees into int64 row pointers: per-block sums, scan of the block sums, per-block fill
__global__ void __launch_bounds__(1024) synth_scan_sums_kernel(const unsigned char* __restrict__ deg, int64_t n, long long* __restrict__ block_sum) {
   __shared__ long long red[32];
   const int64_t base = (int64_t)blockIdx.x * SYN_SCAN_BLOCK;
   long long v = 0;
   for (int x = threadIdx.x; x < SYN_SCAN_BLOCK; x += 1024)
      if (base + x < n) v += deg[base + x];
   for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
   if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
   __syncthreads();
   if (threadIdx.x == 0) {
      long long t = 0;
      for (int w = 0; w < 32; ++w) t += red[w];
      block_sum[blockIdx.x] = t;
   }
}
__global__ void __launch_bounds__(1024) synth_scan_blocks_kernel(long long* __restrict__ block_sum, int n_blocks) {
   // one CTA: every thread owns a contiguous chunk of the block sums
   __shared__ long long part[1024];
   const int chunk = (n_blocks + 1023) / 1024, b0 = threadIdx.x * chunk, b1 = min(n_blocks, b0 + chunk);
   long long s = 0;
   for (int b = b0; b < b1; ++b) s += block_sum[b];
   part[threadIdx.x] = s;
   __syncthreads();
   if (threadIdx.x == 0) {
      long long run = 0;
      for (int t = 0; t < 1024; ++t) { const long long v = part[t]; part[t] = run; run += v; }
   }
   __syncthreads();
   long long run = part[threadIdx.x];
   for (int b = b0; b < b1; ++b) { const long long v = block_sum[b]; block_sum[b] = run; run += v; }
}
__global__ void __launch_bounds__(1024) synth_scan_fill_kernel(const unsigned char* __restrict__ deg, int64_t n, const long long* __restrict__ block_off,
                                                               int64_t* __restrict__ row_ptr, int32_t* __restrict__ count) {
   __shared__ long long wsum[32];
   const int64_t base = (int64_t)blockIdx.x * SYN_SCAN_BLOCK + (int64_t)threadIdx.x * 4;
   int d[4];
   long long mine = 0;
#pragma unroll
   for (int e = 0; e < 4; ++e) { d[e] = base + e < n ? deg[base + e] : 0; mine += d[e]; }
   long long incl = mine;
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
   for (int o = 1; o < 32; o <<= 1) {
      const long long v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
   }
   if (lane == 31) wsum[warp] = incl;
   __syncthreads();
   if (warp == 0) {
      long long w = wsum[lane], wi = w;
      for (int o = 1; o < 32; o <<= 1) {
         const long long v = __shfl_up_sync(0xffffffffu, wi, o);
         if (lane >= o) wi += v;
      }
      wsum[lane] = wi - w;
   }
   __syncthreads();
   long long run = block_off[blockIdx.x] + wsum[warp] + incl - mine;
#pragma unroll
   for (int e = 0; e < 4; ++e) {
      if (base + e < n) { row_ptr[base + e] = run; count[base + e] = 1; }
      run += d[e];
      if (base + e == n - 1) row_ptr[n] = run;
   }
}

// columns and weights: one warp per row, lanes stride the row's entries (coalesced stores)
__global__ void __launch_bounds__(256) synth_fill_kernel(const SynthLocus* __restrict__ loci, int n_loci, int64_t rows, const int64_t* __restrict__ row_ptr,
                                                         int32_t* __restrict__ col, double* __restrict__ alpha) {
   const int lane = threadIdx.x & 31;
   const int64_t n = (int64_t)n_loci * rows;
   const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
   for (int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n; g += nw) {
      const int l = (int)(g / rows);
      const int64_t i = g - (int64_t)l * rows;
      const int64_t k0 = row_ptr[g];
      const int k = (int)(row_ptr[g + 1] - k0);
      const int T = loci[l].T;
      const uint64_t key = loci[l].key_ent;
      for (int m = lane; m < k; m += 32) {
         const uint64_t v = sm64(key + (uint64_t)(256 * i + m));
         const int lo = (int)(((int64_t)m * T) / k), hi = (int)(((int64_t)(m + 1) * T) / k);
         col[k0 + m] = lo + (int)(synth_u(v) * (double)(hi - lo));
         const uint64_t w = sm64(v);
         alpha[k0 + m] = ldexp(1.0 + synth_u(w), -(6 + (int)(w & 7)));
      }
   }
}

__global__ void synth_len_kernel(const SynthLocus* __restrict__ loci, int n_loci, const int64_t* __restrict__ loc_iso_off, int32_t* __restrict__ iso_len) {
   for (int l = blockIdx.x; l < n_loci; l += gridDim.x) {
      const int64_t t0 = loc_iso_off[l];
      for (int j = threadIdx.x; j < loci[l].T; j += blockDim.x) iso_len[t0 + j] = 400 + (int32_t)(sm64(loci[l].key_len + (uint64_t)j) % 7601ull);
   }
}

}  // namespace sbq
