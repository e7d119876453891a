// sbq_rawbuild.cuh - fragment-class assignment ON THE DEVICE (SURVEY section 8f.1, rows a3 / a7 / a9 of the scope table).
//
// What LocusContext::assign_exon_bin computes (reference src/estimate.cpp:135-198, with Contig::is_compatible
// src/contig.cpp:547-599, overlap_exons src/estimate.cpp:115-131, ExonBin::read_count include/isoform.h:285-296), for a whole
// batch of loci at once, from the collapsed hits' and the isoforms' feature lists:
//
//   raw_hit_kernel     one thread per hit: compatibility with every isoform of its locus (bit mask), the disjoint exon
//                      segments its aligned blocks touch (the class coordinates), hashes of both
//   raw_class_insert / raw_class_rep   hash table per locus keyed by the coordinate list: the FIRST hit (smallest index) of
//                      every distinct list is the class representative - the reference numbers classes in first-seen order,
//                      so class id = number of representatives before it (one prefix sum over the hits of the batch)
//   raw_member_kernel  class of every hit, union of the members' isoform masks (iso_2_bins_map), second hash table keyed by
//                      (class, code-blind feature list): the reference keeps the members in a std::set under a comparator that
//                      ignores the op code, so of several equivalent hits only the first inserted one counts
//   raw_mass_kernel    class mass = sum of the surviving members' masses (float), member count
//   raw_nnz_kernel / raw_fill_kernel   CSR rows (classes) x columns (isoforms, ascending), integer counts (int)mass, and per
//                      entry the descriptor of the class under the isoform (segment lengths of the span, implicit-segment mask,
//                      L_t: ExonBin::bin_under_iso, include/isoform.h:363-411) that weights_kernel turns into alpha
//
// Everything is order-free, so atomics are safe: class ids come from a prefix sum, set semantics from atomicMin on the hit
// index, and the float mass sum is EXACT (hence independent of the order the reference's std::set imposes) as long as every
// mass is a multiple of 1/2 and a class holds less than 2^23 of it - true for every BAM read without
// --allow-multimapped-hits (collapse masses are multiplicities of 0.5 + 0.5). The kernels flag what they cannot reproduce
// bit-exactly (fractional masses, more than RB_MAXC segments under one hit, a class spanning more than 32 segments of an
// isoform, a 64-bit hash collision) and the upload is refused with SBQ_ERR_UNSUPPORTED: the caller then uses the host builder.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sbq {

constexpr int RB_MAXC = 16;          // segments one hit may touch
constexpr int RB_POOL = 32;          // segment lengths per CSR entry (descriptor pool stride)
enum : int { RB_FLAG_COORDS = 1, RB_FLAG_COLLISION = 2, RB_FLAG_MASS = 4, RB_FLAG_NSEG = 8 };
constexpr unsigned long long RB_EMPTY = 0xffffffffffffffffull;

struct RawBatch {
   int64_t n_hit;
   int32_t n_loci;
   int32_t long_read;
   // input (device copies of the staged batch)
   const int32_t* hit_locus;       // [H]
   const int64_t* hit_feat_ptr;    // [H + 1]
   const uint32_t* hf_off;
   const uint32_t* hf_len;
   const uint8_t* hf_code;
   const float* hit_mass;          // [H] (PairedHit::collapse_mass read back as float, src/contig.cpp:306-309)
   const int32_t* hit_ref;         // [H]
   const int64_t* loc_hit_off;     // [L + 1]
   const int64_t* loc_iso_off;     // [L + 1]
   const int64_t* loc_seg_off;     // [L + 1]
   const int64_t* iso_feat_ptr;    // [T + 1]
   const uint32_t* if_off;
   const uint32_t* if_len;
   const uint8_t* if_code;
   const uint32_t* seg_left;       // [S] disjoint exon segments of every locus (closed intervals)
   const uint32_t* seg_right;
   const int64_t* iso_seg_ptr;     // [T + 1]
   const int32_t* iso_seg;         // segment indices (local to the locus) contained in each isoform, ascending
   const int32_t* iso_len;         // [T]
   // workspace
   uint16_t* coords;               // [H][RB_MAXC] segment indices (local) the hit touches, ascending
   uint8_t* ncoord;                // [H]
   uint8_t* valid;                 // [H] compatible with at least one isoform and touches at least one segment
   unsigned long long* chash;      // [H] hash of the coordinate list
   unsigned long long* fhash;      // [H] hash of (ref_id, code-blind feature list)
   const int64_t* loc_cm_off;      // [L + 1] first compatibility-mask word of the locus (hit i of the locus: + i * W, W = ceil(T / 32))
   uint32_t* cm;
   const int64_t* loc_tab_off;     // [L + 1] first hash-table slot of the locus; the capacity (a power of two) is the difference
   unsigned long long* keys1;      // class table
   uint32_t* min1;
   uint32_t* slot1;                // [H]
   unsigned long long* keys2;      // member (dedup) table
   uint32_t* min2;
   uint32_t* slot2;                // [H]
   int32_t* is_rep;                // [H]
   int64_t* cls_prefix;            // [H + 1] exclusive prefix sum of is_rep = global class index of a representative
   int32_t* hit_class;             // [H] class id local to the locus, -1 for invalid hits
   int* flags;
   // class-level arrays (allocated once the number of classes is known)
   const int64_t* loc_cls_off;     // [L + 1] first class of the locus
   const int64_t* loc_cmask_off;   // [L + 1] first isoform-mask word of the locus' classes (class c: + c * W)
   int64_t* class_rep;             // [R] representative hit (global index)
   float* class_mass;              // [R]
   int32_t* class_nfrag;           // [R]
   uint32_t* cmask;
   int32_t* class_nnz;             // [R]
};

__device__ __forceinline__ unsigned long long rb_mix(unsigned long long h, unsigned long long x) {
   h ^= x + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
   h *= 0xBF58476D1CE4E5B9ull;
   return h ^ (h >> 29);
}

// Contig::is_compatible(read, isoform) (src/contig.cpp:547-599). Transcripts alternate MATCH / INTRON features
// (Contig 6-argument ctor), so exon k is feature 2 k and "the isoform's next intron" is feature 2 k + 1.
__device__ inline bool rb_compatible(const uint32_t* ro, const uint32_t* rl, const uint8_t* rc, int rn, const uint32_t* io, const uint32_t* il, const uint8_t* ic, int in) {
   const int ne = (in + 1) >> 1;
   if (rn == 0 || ne == 0) return false;
   const uint32_t fl = ro[0], fr = ro[0] + rl[0] - 1;
   int lo = 0, hi = ne;   // first exon with right >= first.left
   while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (io[2 * mid] + il[2 * mid] - 1 < fl) lo = mid + 1; else hi = mid;
   }
   if (lo == ne) return false;
   if (!(io[2 * lo] <= fl && io[2 * lo] + il[2 * lo] - 1 >= fr)) return false;
   int it = lo;
   for (int i = 1; i < rn; ++i) {
      const uint8_t code = rc[i];
      if (code == 2) continue;                                   // S_GAP
      if (code == 1) {                                           // S_INTRON: must equal the isoform's next intron
         const int nx = 2 * it + 1;
         if (nx >= in) return false;
         if (!(ic[nx] == code && io[nx] == ro[i] && il[nx] == rl[i])) return false;
      } else {                                                   // S_MATCH: inside the current or a later exon
         const uint32_t l = ro[i], r = ro[i] + rl[i] - 1;
         while (it < ne && !(io[2 * it] <= l && io[2 * it] + il[2 * it] - 1 >= r)) ++it;
         if (it == ne) return false;
      }
   }
   return true;
}

__global__ void __launch_bounds__(128) raw_hit_kernel(RawBatch b) {
   for (int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; h < b.n_hit; h += (int64_t)gridDim.x * blockDim.x) {
      const int l = b.hit_locus[h];
      const int64_t f0 = b.hit_feat_ptr[h];
      const int rn = (int)(b.hit_feat_ptr[h + 1] - f0);
      b.hit_class[h] = -1;
      b.is_rep[h] = 0;
      b.valid[h] = 0;
      b.ncoord[h] = 0;
      if (rn == 0) continue;                                     // Contig with ref_id == -1: dropped (include/estimate.hpp:71-79)
      const uint32_t *ro = b.hf_off + f0, *rl = b.hf_len + f0;
      const uint8_t* rc = b.hf_code + f0;
      const int64_t t0 = b.loc_iso_off[l];
      const int T = (int)(b.loc_iso_off[l + 1] - t0), W = (T + 31) >> 5;
      uint32_t* cm = b.cm + b.loc_cm_off[l] + (h - b.loc_hit_off[l]) * W;
      int ncompat = 0;
      for (int w = 0; w < W; ++w) {
         uint32_t bits = 0;
         for (int t = 32 * w; t < min(T, 32 * w + 32); ++t) {
            const int64_t i0 = b.iso_feat_ptr[t0 + t];
            const int in = (int)(b.iso_feat_ptr[t0 + t + 1] - i0);
            if (rb_compatible(ro, rl, rc, rn, b.if_off + i0, b.if_len + i0, b.if_code + i0, in)) { bits |= 1u << (t & 31); ++ncompat; }
         }
         cm[w] = bits;
      }
      // overlap_exons(): segments that any MATCH block of the hit touches (closed intervals), ascending
      const int64_t s0 = b.loc_seg_off[l];
      const int S = (int)(b.loc_seg_off[l + 1] - s0);
      int nc = 0;
      bool overflow = false;
      unsigned long long ch = 1469598103934665603ull;
      for (int k = 0; k < rn && !overflow; ++k) {
         if (rc[k] != 0) continue;
         const uint32_t l_ = ro[k], r_ = ro[k] + rl[k] - 1;
         int lo = 0, hi = S;                                     // first segment with right >= block.left
         while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (b.seg_right[s0 + mid] < l_) lo = mid + 1; else hi = mid;
         }
         for (int s = lo; s < S && b.seg_left[s0 + s] <= r_; ++s) {
            if (nc > 0 && (int)b.coords[h * RB_MAXC + nc - 1] >= s) continue;   // blocks are ascending: already listed by the previous block
            if (nc == RB_MAXC) { overflow = true; break; }
            b.coords[h * RB_MAXC + nc++] = (uint16_t)s;
         }
      }
      if (overflow) { atomicOr(b.flags, RB_FLAG_COORDS); continue; }
      for (int i = 0; i < nc; ++i) ch = rb_mix(ch, b.coords[h * RB_MAXC + i]);
      b.ncoord[h] = (uint8_t)nc;
      b.chash[h] = ch == RB_EMPTY ? 0 : ch;
      unsigned long long fh = rb_mix(88172645463325252ull, (unsigned)b.hit_ref[h]);
      for (int k = 0; k < rn; ++k) fh = rb_mix(rb_mix(fh, ro[k]), rl[k]);
      b.fhash[h] = fh;
      const bool ok = ncompat > 0 && nc > 0;
      b.valid[h] = ok;
      if (ok) {
         const float m = b.hit_mass[h];
         const float m2 = m * 2.0f;
         if (!(m >= 0.0f && m < 4194304.0f && m2 == floorf(m2))) atomicOr(b.flags, RB_FLAG_MASS);   // not a multiple of 1/2: the order of the sum would matter
      }
   }
}

// open-addressing insert: returns the slot of `key` in the locus' table; min[slot] ends up as the smallest hit index with that key
__device__ __forceinline__ uint32_t rb_insert(unsigned long long* keys, uint32_t* mins, int64_t base, uint32_t cap, unsigned long long key, uint32_t h) {
   uint32_t s = (uint32_t)(key * 0x9E3779B97F4A7C15ull >> 32) & (cap - 1);
   for (;;) {
      const unsigned long long old = atomicCAS(&keys[base + s], RB_EMPTY, key);
      if (old == RB_EMPTY || old == key) {
         atomicMin(&mins[base + s], h);
         return s;
      }
      s = (s + 1) & (cap - 1);
   }
}

__global__ void __launch_bounds__(256) raw_class_insert_kernel(RawBatch b) {
   for (int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; h < b.n_hit; h += (int64_t)gridDim.x * blockDim.x) {
      if (!b.valid[h]) continue;
      const int l = b.hit_locus[h];
      const int64_t base = b.loc_tab_off[l];
      const uint32_t cap = (uint32_t)(b.loc_tab_off[l + 1] - base);
      b.slot1[h] = rb_insert(b.keys1, b.min1, base, cap, b.chash[h], (uint32_t)(h - b.loc_hit_off[l]));
   }
}

__global__ void __launch_bounds__(256) raw_class_rep_kernel(RawBatch b) {
   for (int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; h < b.n_hit; h += (int64_t)gridDim.x * blockDim.x) {
      if (!b.valid[h]) continue;
      const int l = b.hit_locus[h];
      const int64_t rep = b.loc_hit_off[l] + b.min1[b.loc_tab_off[l] + b.slot1[h]];
      if (rep == h) { b.is_rep[h] = 1; continue; }
      bool same = b.ncoord[rep] == b.ncoord[h];                  // a 64-bit collision of two different lists: refuse rather than merge classes
      for (int i = 0; same && i < b.ncoord[h]; ++i) same = b.coords[rep * RB_MAXC + i] == b.coords[h * RB_MAXC + i];
      if (!same) atomicOr(b.flags, RB_FLAG_COLLISION);
   }
}

__global__ void __launch_bounds__(256) raw_member_kernel(RawBatch b) {
   for (int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; h < b.n_hit; h += (int64_t)gridDim.x * blockDim.x) {
      if (!b.valid[h]) continue;
      const int l = b.hit_locus[h];
      const int64_t h0 = b.loc_hit_off[l];
      const int64_t rep = h0 + b.min1[b.loc_tab_off[l] + b.slot1[h]];
      const int64_t cg = b.cls_prefix[rep];                      // global class index = representatives before it
      const int cid = (int)(cg - b.loc_cls_off[l]);
      b.hit_class[h] = cid;
      if (rep == h) b.class_rep[cg] = h;
      const int T = (int)(b.loc_iso_off[l + 1] - b.loc_iso_off[l]), W = (T + 31) >> 5;
      const uint32_t* cm = b.cm + b.loc_cm_off[l] + (h - h0) * W;
      uint32_t* dst = b.cmask + b.loc_cmask_off[l] + (int64_t)cid * W;
      for (int w = 0; w < W; ++w)
         if (cm[w]) atomicOr(&dst[w], cm[w]);                    // iso_2_bins_map[t].insert(class) for every compatible (hit, isoform)
      const int64_t base = b.loc_tab_off[l];
      const uint32_t cap = (uint32_t)(b.loc_tab_off[l + 1] - base);
      unsigned long long key = rb_mix(b.fhash[h], (unsigned long long)cid + 1);
      if (key == RB_EMPTY) key = 0;
      b.slot2[h] = rb_insert(b.keys2, b.min2, base, cap, key, (uint32_t)(h - h0));
   }
}

__global__ void __launch_bounds__(256) raw_mass_kernel(RawBatch b) {
   for (int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; h < b.n_hit; h += (int64_t)gridDim.x * blockDim.x) {
      if (!b.valid[h]) continue;
      const int l = b.hit_locus[h];
      const int64_t h0 = b.loc_hit_off[l];
      const int64_t first = h0 + b.min2[b.loc_tab_off[l] + b.slot2[h]];
      if (first != h) {
         // an equivalent hit (same class, same ref_id, same (offset, len) list - op codes ignored) was inserted first: this one is
         // not a new set element, its mass does not count. Verify the equivalence (a hash collision must not drop a real member).
         const int64_t a0 = b.hit_feat_ptr[h], b0 = b.hit_feat_ptr[first];
         const int n = (int)(b.hit_feat_ptr[h + 1] - a0);
         bool same = n == (int)(b.hit_feat_ptr[first + 1] - b0) && b.hit_ref[h] == b.hit_ref[first] && b.hit_class[first] == b.hit_class[h];
         for (int k = 0; same && k < n; ++k) same = b.hf_off[a0 + k] == b.hf_off[b0 + k] && b.hf_len[a0 + k] == b.hf_len[b0 + k];
         if (!same) atomicOr(b.flags, RB_FLAG_COLLISION);
         continue;
      }
      const int64_t cg = b.loc_cls_off[l] + b.hit_class[h];
      atomicAdd(&b.class_mass[cg], b.hit_mass[h]);               // exact: multiples of 1/2, class total checked against 2^23 below
      atomicAdd(&b.class_nfrag[cg], 1);
   }
}

__global__ void __launch_bounds__(256) raw_nnz_kernel(RawBatch b, int64_t n_class, const int32_t* __restrict__ class_locus) {
   for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_class; c += (int64_t)gridDim.x * blockDim.x) {
      const int l = class_locus[c];
      const int T = (int)(b.loc_iso_off[l + 1] - b.loc_iso_off[l]), W = (T + 31) >> 5;
      const uint32_t* m = b.cmask + b.loc_cmask_off[l] + (c - b.loc_cls_off[l]) * W;
      int n = 0;
      for (int w = 0; w < W; ++w) n += __popc(m[w]);
      b.class_nnz[c] = n;
      if (!(b.class_mass[c] < 8388608.0f)) atomicOr(b.flags, RB_FLAG_MASS);
   }
}

// CSR rows of the batch + per-entry weight descriptors. row_ptr (exclusive scan of class_nnz) is already in place.
__global__ void __launch_bounds__(128)
raw_fill_kernel(RawBatch b, int64_t n_class, const int32_t* __restrict__ class_locus, const int64_t* __restrict__ row_ptr, int32_t* __restrict__ col, double* __restrict__ alpha,
                int32_t* __restrict__ count, int64_t* __restrict__ w_seg, uint8_t* __restrict__ w_n, uint32_t* __restrict__ w_mask, int32_t* __restrict__ w_len, uint32_t* __restrict__ w_pool) {
   for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_class; c += (int64_t)gridDim.x * blockDim.x) {
      const int l = class_locus[c];
      const int64_t t0 = b.loc_iso_off[l];
      const int T = (int)(b.loc_iso_off[l + 1] - t0), W = (T + 31) >> 5;
      const uint32_t* m = b.cmask + b.loc_cmask_off[l] + (c - b.loc_cls_off[l]) * W;
      count[c] = (int32_t)b.class_mass[c];                       // n_c = (int) float sum (src/estimate.cpp:288)
      const int64_t rep = b.class_rep[c];
      const uint16_t* cc = b.coords + rep * RB_MAXC;
      const int ncc = b.ncoord[rep];
      const int64_t s0 = b.loc_seg_off[l];
      int64_t k = row_ptr[c];
      for (int w = 0; w < W; ++w) {
         uint32_t bits = m[w];
         while (bits) {
            const int t = 32 * w + __ffs(bits) - 1;
            bits &= bits - 1;
            col[k] = t;
            w_len[k] = b.iso_len[t0 + t];
            if (b.long_read) {                                   // set_bin_weight_without_frag_dist: alpha = 1 / L_t (src/estimate.cpp:236-247)
               alpha[k] = 1.0 / (double)b.iso_len[t0 + t];
               w_seg[k] = -1; w_n[k] = 0; w_mask[k] = 0;
            } else {
               // ExonBin::bin_under_iso (include/isoform.h:363-411): the isoform's segments from the class' first to its last
               // coordinate; the inner ones the class does not list are implicit
               const int32_t* isegs = b.iso_seg + b.iso_seg_ptr[t0 + t];
               const int nis = (int)(b.iso_seg_ptr[t0 + t + 1] - b.iso_seg_ptr[t0 + t]);
               int lo = 0, hi = nis;
               while (lo < hi) { const int mid = (lo + hi) >> 1; if (isegs[mid] < (int)cc[0]) lo = mid + 1; else hi = mid; }
               int up = 0;
               hi = nis;
               while (up < hi) { const int mid = (up + hi) >> 1; if (isegs[mid] < (int)cc[ncc - 1]) up = mid + 1; else hi = mid; }
               const int nseg = up - lo + 1;
               alpha[k] = 0.0;
               if (lo >= nis || up >= nis || nseg > RB_POOL || nseg < 1) {
                  atomicOr(b.flags, RB_FLAG_NSEG);
                  w_seg[k] = -1; w_n[k] = 0; w_mask[k] = 0;
               } else {
                  uint32_t mask = 0;
                  int cpos = 1;
                  for (int i = 0; i < nseg; ++i) {
                     const int sg = isegs[lo + i];
                     w_pool[k * RB_POOL + i] = b.seg_right[s0 + sg] - b.seg_left[s0 + sg] + 1;
                     if (i >= 1 && i + 1 < nseg) {
                        if (cpos < ncc && sg == (int)cc[cpos]) ++cpos; else mask |= 1u << i;
                     }
                  }
                  w_seg[k] = k * RB_POOL;
                  w_n[k] = (uint8_t)nseg;
                  w_mask[k] = mask;
               }
            }
            ++k;
         }
      }
   }
}

// ---- generic exclusive scan int32 -> int64 (three kernels, 4096 elements per block)
constexpr int RB_SCAN_BLOCK = 4096;
__global__ void __launch_bounds__(1024) rb_scan_sums_kernel(const int32_t* __restrict__ v, int64_t n, long long* __restrict__ block_sum) {
   __shared__ long long red[32];
   const int64_t base = (int64_t)blockIdx.x * RB_SCAN_BLOCK;
   long long s = 0;
   for (int x = threadIdx.x; x < RB_SCAN_BLOCK; x += 1024)
      if (base + x < n) s += v[base + x];
   for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
   if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
   __syncthreads();
   if (threadIdx.x == 0) {
      long long t = 0;
      for (int w = 0; w < 32; ++w) t += red[w];
      block_sum[blockIdx.x] = t;
   }
}
__global__ void __launch_bounds__(1024) rb_scan_fill_kernel(const int32_t* __restrict__ v, int64_t n, const long long* __restrict__ block_off, int64_t* __restrict__ out) {
   __shared__ long long wsum[32];
   const int64_t base = (int64_t)blockIdx.x * RB_SCAN_BLOCK + (int64_t)threadIdx.x * 4;
   int d[4];
   long long mine = 0;
#pragma unroll
   for (int e = 0; e < 4; ++e) { d[e] = base + e < n ? v[base + e] : 0; mine += d[e]; }
   long long incl = mine;
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
   for (int o = 1; o < 32; o <<= 1) {
      const long long u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
   }
   if (lane == 31) wsum[warp] = incl;
   __syncthreads();
   if (warp == 0) {
      long long w = wsum[lane], wi = w;
      for (int o = 1; o < 32; o <<= 1) {
         const long long u = __shfl_up_sync(0xffffffffu, wi, o);
         if (lane >= o) wi += u;
      }
      wsum[lane] = wi - w;
   }
   __syncthreads();
   long long run = block_off[blockIdx.x] + wsum[warp] + incl - mine;
#pragma unroll
   for (int e = 0; e < 4; ++e) {
      if (base + e < n) out[base + e] = run;
      run += d[e];
      if (base + e == n - 1) out[n] = run;
   }
}

__global__ void rb_gather_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ idx, int64_t n, int64_t* __restrict__ out) {
   const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
   if (i < n) out[i] = src[idx[i]];
}

}  // namespace sbq
