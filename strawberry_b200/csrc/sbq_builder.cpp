// sbq_builder.cpp - host class-table builder (include/sbq_builder.h), part of libsbq.so.
//
// From-scratch implementation of what the reference's LocusContext constructor computes
// (include/estimate.hpp:61-109): disjoint exon segments, per-isoform segment lists, the
// fragment-class table (first-seen class ids, set-deduplicated float counts) and the class weights
// alpha, emitted as the CSR that sbq_submit() takes. Flat arrays and hash lookups instead of the
// reference's std::set / std::map / linear searches; the integer results are bit-identical
// (tests/test_builder.py checks them against the compiled reference and committed goldens).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <new>
#include <unordered_map>
#include <vector>

#include "../../include/sbq_builder.h"

namespace {

struct Feat {
   uint32_t off, len;
   uint8_t code;
   uint32_t left() const { return off; }
   uint32_t right() const { return off + len - 1; }
};
// GenomicFeature::operator< compares (offset, len) only - the op code is ignored (src/contig.cpp:186-193)
inline bool feat_lt(const Feat& a, const Feat& b) { return a.off != b.off ? a.off < b.off : a.len < b.len; }
inline bool feat_eq(const Feat& a, const Feat& b) { return a.code == b.code && a.off == b.off && a.len == b.len; }
// GenomicFeature::contains with small_extent = 0 (src/contig.cpp:120-125)
inline bool contains(const Feat& outer, const Feat& inner) { return outer.left() <= inner.left() && outer.right() >= inner.right(); }

struct FeatList {
   const Feat* p;
   int n;
};

// Contig::operator< for the _frags set: ref_id, then features lexicographically by (offset, len)
// (src/contig.cpp:342-347). Hits that compare equivalent are ONE set element: the first one wins.
struct FragKey {
   int32_t ref_id;
   FeatList f;
};
struct FragLess {
   bool operator()(const FragKey& a, const FragKey& b) const {
      if (a.ref_id != b.ref_id) return a.ref_id < b.ref_id;
      return std::lexicographical_compare(a.f.p, a.f.p + a.f.n, b.f.p, b.f.p + b.f.n, feat_lt);
   }
};

struct VecHash {
   size_t operator()(const std::vector<int32_t>& v) const {
      uint64_t h = 1469598103934665603ull;
      for (int32_t x : v) { h ^= (uint32_t)x; h *= 1099511628211ull; }
      return (size_t)h;
   }
};

// ---- a6: ExonBin::no_gap_ef / gap_ef / effective_len (include/isoform.h:105-129, :419-516) -------------
inline int no_gap_ef(int l_left, int l_right, int l_int, int fl) {
   if (fl < l_int + 2) return 0;
   if (fl > l_left + l_right + l_int) return 0;
   const int mid = fl - l_int - 1;
   return std::min(l_left, mid) + std::min(l_right, mid) - mid;
}
inline int gap_ef(int l_left, int l_right, int l_int, int rl, int gap) {
   if (2 * rl + gap < l_int + 2) return 0;
   if (2 * rl + gap > l_left + l_right + l_int) return 0;
   const int start = std::max(rl, l_left + l_int - gap - 1);
   const int end = std::min(l_left, l_left + l_right + l_int - gap - rl);
   return std::max(0, end - start);
}

int effective_len(const uint32_t* s, int n, const uint32_t* implicit, int n_imp, int fl, int rl) {
   const int gap = fl - 2 * rl;
   if (n == 1) return (int)(s[0] - (uint32_t)fl + 1u);
   if (n == 2) return no_gap_ef((int)s[0], (int)s[1], 0, fl);
   if (n == 3) {
      if (n_imp == 1) return gap_ef((int)s[0], (int)s[2], (int)s[1], rl, gap);
      return no_gap_ef((int)s[0], (int)s[2], (int)s[1], fl) - gap_ef((int)s[0], (int)s[2], (int)s[1], rl, gap);
   }
   if (n == 4) {
      const int hit14 = gap_ef((int)s[0], (int)s[3], (int)(s[2] + s[1]), rl, gap);
      const int hit24 = gap_ef((int)s[3], (int)s[1], (int)s[2], rl, gap);
      const int hit124 = gap_ef((int)(s[0] + s[1]), (int)s[3], (int)s[2], rl, gap);
      const int hit13 = gap_ef((int)s[0], (int)s[2], (int)s[1], rl, gap);
      const int hit134 = gap_ef((int)s[0], (int)(s[2] + s[3]), (int)s[1], rl, gap);
      if (n_imp == 0) {
         const int all124 = hit124 - hit14 - hit24, all134 = hit134 - hit14 - hit13;
         const int total = no_gap_ef((int)s[0], (int)s[3], (int)(s[1] + s[2]), fl);
         return total - all124 - all134 - hit14;
      }
      if (n_imp == 2) return hit14;
      if (implicit[0] == 1) return hit134 - hit14 - hit13;
      return hit124 - hit14 - hit24;
   }
   // more than four segments: enumerate the start offsets in the first segment (include/isoform.h:476-512)
   const uint32_t num_inners = (uint32_t)n - 2;
   uint32_t num_pos = 0;
   uint32_t target = (uint32_t)(uint64_t)(std::pow(2.0, (double)n) - 1.0);
   for (int k = 0; k < n_imp; ++k) target &= ~(1u << (implicit[k] & 31u));
   int inner_sum = 0;
   for (int k = 1; k < n - 1; ++k) inner_sum += (int)s[k];
   for (int i = 1; i != (int)(s[0] + 1u); ++i) {
      uint32_t hit = 1;
      const int bp_last = fl - i - inner_sum;
      if ((uint32_t)bp_last > s[n - 1]) continue;   // int vs uint comparison in the reference: negative => skipped
      if (bp_last == 0) break;
      hit |= 1u << ((uint32_t)(n - 1) & 31u);
      int last_rest = rl - bp_last;
      uint32_t j = num_inners;
      while (last_rest > 0 && j > 0) {
         hit |= 1u << (j & 31u);
         last_rest = (int)((uint32_t)last_rest - s[j]);
         --j;
      }
      int first_rest = rl - i;
      j = 1;
      while (first_rest > 0 && j <= num_inners) {
         hit |= 1u << (j & 31u);
         first_rest = (int)((uint32_t)first_rest - s[j]);
         ++j;
      }
      if (hit == target) ++num_pos;
   }
   return (int)num_pos;
}

// ---- a13: normal_pdf (include/common.h:92-99) and InsertSize::emp_dist_pdf (src/read.cpp:274-297) ------
inline double normal_pdf(double x, double m, double s) {
   static const double inv_sqrt_2pi = 0.3989422804014327;
   const double a = (x - m) / s;
   return inv_sqrt_2pi / s * std::exp(-0.5 * a * a);
}
double insert_pdf(const sbq_insert_model& m, uint32_t insert_size) {
   if (m.use_emp) {
      double ret = 0.0;
      if (!(insert_size < (uint32_t)m.start_offset || insert_size > (uint32_t)m.end_offset))
         ret = m.emp_dist[insert_size - (uint32_t)m.start_offset] / m.total_reads;
      if (ret != 0.0) return ret;
   }
   const double p = normal_pdf((double)insert_size, m.mean, m.sd);
   return p > 0 ? p : 0.0;
}

// ---- a9: Contig::is_compatible(read, isoform) (src/contig.cpp:547-599) -------------------------------
bool is_compatible(const FeatList& read, const FeatList& iso, const std::vector<int>& iso_exon_idx) {
   // iso_exon_idx: positions of the isoform's MATCH features inside iso.p
   const int ne = (int)iso_exon_idx.size();
   if (read.n == 0 || ne == 0) return false;
   const Feat& first = read.p[0];
   int lo = 0, hi = ne;   // lower_bound: first exon with right >= first.left
   while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (iso.p[iso_exon_idx[mid]].right() < first.left()) lo = mid + 1; else hi = mid;
   }
   if (lo == ne) return false;
   if (!contains(iso.p[iso_exon_idx[lo]], first)) return false;
   int it = lo;
   for (int i = 1; i < read.n; ++i) {
      const Feat& f = read.p[i];
      if (f.code == SBQ_FEAT_GAP) continue;
      if (f.code == SBQ_FEAT_INTRON) {
         const size_t next_intron = 2 * (size_t)it + 1;
         if (next_intron >= (size_t)iso.n) return false;
         if (!feat_eq(f, iso.p[next_intron])) return false;
      } else {
         while (it < ne && !contains(iso.p[iso_exon_idx[it]], f)) ++it;
         if (it == ne) return false;
      }
   }
   return true;
}

}  // namespace

struct sbq_table {
   std::vector<uint32_t> seg_left, seg_right;
   std::vector<int32_t> iso_seg_ptr, iso_seg;
   std::vector<int32_t> class_coord_ptr, class_coord, class_count, class_nfrag;
   std::vector<float> class_mass;
   std::vector<int64_t> row_ptr;
   std::vector<int32_t> col;
   std::vector<double> alpha;
   std::vector<int32_t> iso_len;
   std::vector<int32_t> hit_class;   // class of every input hit, -1 = dropped / compatible with nothing
   // deferred weights (alpha computed on the GPU): per CSR entry the offset of its segment lengths in w_pool (-1 = alpha
   // is final), their number, the bit mask of implicit segments and the isoform length
   std::vector<int64_t> w_seg_ptr;
   std::vector<uint8_t> w_nseg;
   std::vector<uint32_t> w_mask, w_pool;
   std::vector<int32_t> w_len;
   int32_t n_dropped = 0;
};

extern "C" {

int32_t sbq_effective_len(const uint32_t* seg_lens, int32_t n_seg, const uint32_t* implicit_idx, int32_t n_implicit, int32_t fl, int32_t rl) {
   if (!seg_lens || n_seg < 1) return 0;
   return effective_len(seg_lens, n_seg, implicit_idx, n_implicit, fl, rl);
}

double sbq_insert_pdf(const sbq_insert_model* model, uint32_t insert_size) { return model ? insert_pdf(*model, insert_size) : 0.0; }

int sbq_pair_features(uint32_t left_pos, const uint8_t* lop, const uint32_t* llen, int32_t ln, uint32_t right_pos, const uint8_t* rop,
                      const uint32_t* rlen, int32_t rn, uint32_t* feat_off, uint32_t* feat_len, uint8_t* feat_code, int32_t cap) {
   std::vector<Feat> g;
   // readhit_2_genomicFeats (src/contig.cpp:10-52); its bool result is ignored by the caller, so a bad
   // CIGAR simply stops contributing features. Returns the reference-span right end of the mate.
   auto mate = [&](uint32_t pos, const uint8_t* op, const uint32_t* len, int n) {
      uint32_t offset = pos;
      for (int i = 0; i < n; ++i) {
         switch (op[i]) {
            case SBQ_CIG_MATCH: g.push_back(Feat{offset, len[i], SBQ_FEAT_MATCH}); offset += len[i]; break;
            case SBQ_CIG_REF_SKIP: g.push_back(Feat{offset, len[i], SBQ_FEAT_INTRON}); offset += len[i]; break;
            case SBQ_CIG_DEL:
               if (i < 1 || i + 1 == n || op[i - 1] != SBQ_CIG_MATCH || op[i + 1] != SBQ_CIG_MATCH) return;
               g.back().len += len[i];
               offset += len[i];
               break;
            case SBQ_CIG_INS:
               if (i < 1 || i + 1 == n || op[i - 1] != SBQ_CIG_MATCH || op[i + 1] != SBQ_CIG_MATCH) return;
               break;
            case SBQ_CIG_SOFT_CLIP: break;
            default: return;
         }
      }
   };
   auto span_right = [](uint32_t pos, const uint8_t* op, const uint32_t* len, int n) {
      uint32_t s = 0;
      for (int i = 0; i < n; ++i)
         if (op[i] == SBQ_CIG_MATCH || op[i] == SBQ_CIG_REF_SKIP || op[i] == SBQ_CIG_DEL) s += len[i];
      return pos + s - 1;
   };
   if (ln > 0 && rn > 0) {
      mate(left_pos, lop, llen, ln);
      mate(right_pos, rop, rlen, rn);
      const uint32_t lr = span_right(left_pos, lop, llen, ln);
      const int gap_len = (int)right_pos - (int)lr - 1;
      if (gap_len > 0) {
         g.push_back(Feat{lr + 1, (uint32_t)gap_len, SBQ_FEAT_GAP});
      } else {
         // overlapping mates: sort + merge_genomicFeats (include/contig.h:111-138)
         std::sort(g.begin(), g.end(), feat_lt);
         std::vector<Feat> res;
         bool bad = false;
         for (size_t i = 0; i < g.size() && !bad; ++i) {
            res.push_back(g[i]);
            Feat& f = res.back();
            while (i + 1 < g.size() && f.code == g[i + 1].code) {
               if (f.code == SBQ_FEAT_INTRON) {
                  if (!feat_eq(f, g[i + 1])) { bad = true; break; }
               } else {
                  if (f.right() < g[i + 1].left()) { bad = true; break; }
                  const uint32_t right = std::max(f.right(), g[i + 1].right());
                  f.len = right - f.left() + 1;
               }
               ++i;
            }
         }
         if (bad) res.clear();
         g.swap(res);
      }
   } else {
      if (rn > 0) mate(right_pos, rop, rlen, rn);
      if (ln > 0) mate(left_pos, lop, llen, ln);
   }
   if (g.empty()) return 0;
   std::sort(g.begin(), g.end(), feat_lt);
   if ((int)g.size() > cap) return SBQ_ERR_INVALID;
   for (size_t i = 0; i < g.size(); ++i) { feat_off[i] = g[i].off; feat_len[i] = g[i].len; feat_code[i] = g[i].code; }
   return (int)g.size();
}

int sbq_build_locus(const sbq_locus_input* in, const sbq_insert_model* model, sbq_table** out) {
   if (!in || !out || in->n_iso < 1 || !in->iso_feat_ptr || (!model && !in->long_read)) return SBQ_ERR_INVALID;
   *out = nullptr;
   sbq_table* tb = new (std::nothrow) sbq_table();
   if (!tb) return SBQ_ERR_NOMEM;
   const int T = in->n_iso;

   // ---- feature lists
   std::vector<Feat> iso_feats((size_t)in->iso_feat_ptr[T]);
   for (size_t k = 0; k < iso_feats.size(); ++k) iso_feats[k] = Feat{in->iso_feat_off[k], in->iso_feat_len[k], in->iso_feat_code[k]};
   std::vector<Feat> hit_feats(in->n_hit > 0 ? (size_t)in->hit_feat_ptr[in->n_hit] : 0);
   for (size_t k = 0; k < hit_feats.size(); ++k) hit_feats[k] = Feat{in->hit_feat_off[k], in->hit_feat_len[k], in->hit_feat_code[k]};
   auto iso_list = [&](int t) { return FeatList{iso_feats.data() + in->iso_feat_ptr[t], in->iso_feat_ptr[t + 1] - in->iso_feat_ptr[t]}; };
   auto hit_list = [&](int h) { return FeatList{hit_feats.data() + in->hit_feat_ptr[h], in->hit_feat_ptr[h + 1] - in->hit_feat_ptr[h]}; };

   // ---- a12: exons = sorted unique MATCH features of all transcripts; disjoint() (include/interval.hpp:150-223)
   std::vector<Feat> exons;
   for (const Feat& f : iso_feats)
      if (f.code == SBQ_FEAT_MATCH) exons.push_back(f);
   if (exons.empty()) { delete tb; return SBQ_ERR_INVALID; }
   std::sort(exons.begin(), exons.end(), feat_lt);
   exons.erase(std::unique(exons.begin(), exons.end(), feat_eq), exons.end());
   {
      // breakpoints = sorted unique lefts and half-open rights; a piece [bar_k, bar_k+1) is a segment when it is
      // covered. Coverage only matters at the breakpoints, so a sweep over the exon ends replaces the per-base array.
      std::vector<uint32_t> bars;
      bars.reserve(exons.size() * 2);
      for (const Feat& e : exons) { bars.push_back(e.left()); bars.push_back(e.right() + 1); }
      std::sort(bars.begin(), bars.end());
      bars.erase(std::unique(bars.begin(), bars.end()), bars.end());
      std::vector<int> delta(bars.size(), 0);
      for (const Feat& e : exons) {
         delta[std::lower_bound(bars.begin(), bars.end(), e.left()) - bars.begin()] += 1;
         delta[std::lower_bound(bars.begin(), bars.end(), e.right() + 1) - bars.begin()] -= 1;
      }
      int cov = 0;
      for (size_t k = 0; k + 1 < bars.size(); ++k) {
         cov += delta[k];
         if (cov > 0) { tb->seg_left.push_back(bars[k]); tb->seg_right.push_back(bars[k + 1] - 1); }
      }
   }
   const int S = (int)tb->seg_left.size();

   // ---- a8: per isoform: exon positions, length, contained segments (src/contig.cpp:615-634)
   std::vector<std::vector<int>> iso_exon_idx(T);
   tb->iso_len.resize(T);
   tb->iso_seg_ptr.assign(1, 0);
   for (int t = 0; t < T; ++t) {
      const FeatList L = iso_list(t);
      int len = 0;
      for (int k = 0; k < L.n; ++k)
         if (L.p[k].code == SBQ_FEAT_MATCH) { iso_exon_idx[t].push_back(k); len += (int)L.p[k].len; }
      tb->iso_len[t] = len;
      const std::vector<int>& ex = iso_exon_idx[t];
      for (int s = 0; s < S; ++s) {
         const Feat seg{tb->seg_left[s], tb->seg_right[s] - tb->seg_left[s] + 1, SBQ_FEAT_MATCH};
         int lo = 0, hi = (int)ex.size();
         while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (L.p[ex[mid]].right() < seg.left()) lo = mid + 1; else hi = mid;
         }
         if (lo < (int)ex.size() && contains(L.p[ex[lo]], seg)) tb->iso_seg.push_back(s);
      }
      tb->iso_seg_ptr.push_back((int32_t)tb->iso_seg.size());
   }

   // ---- a3: class assignment (src/estimate.cpp:135-198). Classes in first-seen order; _frags deduplicated by the
   //      code-blind Contig comparator; iso -> classes map.
   std::unordered_map<std::vector<int32_t>, int32_t, VecHash> class_of;
   std::vector<std::vector<int32_t>> class_coords;
   std::vector<std::map<FragKey, float, FragLess>> class_frags;
   std::vector<std::vector<int32_t>> iso_classes(T);   // kept sorted & unique (std::set<int> in the reference)
   std::vector<int32_t> coords;
   tb->hit_class.assign(in->n_hit > 0 ? in->n_hit : 0, -1);
   for (int h = 0; h < in->n_hit; ++h) {
      const FeatList H = hit_list(h);
      if (H.n == 0) { ++tb->n_dropped; continue; }   // Contig with ref_id == -1 (include/estimate.hpp:71-79)
      bool have_coords = false;
      for (int t = 0; t < T; ++t) {
         if (!is_compatible(H, iso_list(t), iso_exon_idx[t])) continue;
         if (!have_coords) {   // overlap_exons(): segments overlapping any MATCH block of the hit (src/estimate.cpp:115-131)
            coords.clear();
            for (int s = 0; s < S; ++s) {
               for (int k = 0; k < H.n; ++k) {
                  if (H.p[k].code != SBQ_FEAT_MATCH) continue;
                  if (H.p[k].left() <= tb->seg_right[s] && tb->seg_left[s] <= H.p[k].right()) { coords.push_back(s); break; }
               }
            }
            have_coords = true;
         }
         if (coords.empty()) continue;
         auto ins = class_of.emplace(coords, (int32_t)class_coords.size());
         const int32_t cid = ins.first->second;
         tb->hit_class[h] = cid;
         if (ins.second) { class_coords.push_back(coords); class_frags.emplace_back(); }
         const int32_t ref_id = in->hit_ref_id ? in->hit_ref_id[h] : 0;
         class_frags[cid].emplace(FragKey{ref_id, H}, (float)in->hit_mass[h]);   // set::insert: first one wins
         std::vector<int32_t>& ic = iso_classes[t];
         auto pos = std::lower_bound(ic.begin(), ic.end(), cid);
         if (pos == ic.end() || *pos != cid) ic.insert(pos, cid);
      }
   }
   const int R = (int)class_coords.size();

   // ---- a7: counts = (int) float-sum of masses in set order (include/isoform.h:285-296, src/estimate.cpp:288)
   tb->class_coord_ptr.assign(1, 0);
   for (int c = 0; c < R; ++c) {
      tb->class_coord.insert(tb->class_coord.end(), class_coords[c].begin(), class_coords[c].end());
      tb->class_coord_ptr.push_back((int32_t)tb->class_coord.size());
      float sum = 0.0f;
      for (auto const& kv : class_frags[c]) sum += kv.second;
      tb->class_mass.push_back(sum);
      tb->class_count.push_back((int32_t)sum);
      tb->class_nfrag.push_back((int32_t)class_frags[c].size());
   }

   // ---- a4 / a5: weights per (isoform, class) (src/estimate.cpp:201-247), gathered per class row afterwards
   struct Entry { int32_t t; double w; int64_t seg_ptr; uint8_t nseg; uint32_t mask; };
   std::vector<std::vector<Entry>> row_entries(R);
   std::vector<uint32_t> seg_lens, implicit;
   for (int t = 0; t < T; ++t) {
      const int32_t* isegs = tb->iso_seg.data() + tb->iso_seg_ptr[t];
      const int nis = tb->iso_seg_ptr[t + 1] - tb->iso_seg_ptr[t];
      for (int32_t cid : iso_classes[t]) {
         double weight;
         Entry ent{t, 0.0, -1, 0, 0u};
         if (in->long_read) {
            weight = 1.0 / tb->iso_len[t];
         } else {
            // ExonBin::bin_under_iso (include/isoform.h:363-411): isoform segments from the class' first to its last
            // coordinate; the inner ones the class does not list are "implicit"
            const std::vector<int32_t>& cc = class_coords[cid];
            const int lo = (int)(std::lower_bound(isegs, isegs + nis, cc.front()) - isegs);
            const int up = (int)(std::lower_bound(isegs, isegs + nis, cc.back()) - isegs);
            if (lo >= nis || up >= nis) { delete tb; return SBQ_ERR_INVALID; }   // the reference asserts here
            seg_lens.clear();
            implicit.clear();
            for (int k = lo; k <= up; ++k) seg_lens.push_back(tb->seg_right[isegs[k]] - tb->seg_left[isegs[k]] + 1);
            size_t c = 1;
            for (int i = 1; i + 1 < (int)seg_lens.size(); ++i) {
               if (c < cc.size() && isegs[lo + i] == cc[c]) ++c; else implicit.push_back((uint32_t)i);
            }
            const int nseg = (int)seg_lens.size();
            int lmax = 0;
            for (uint32_t x : seg_lens) lmax += (int)x;
            int lmin = model->use_emp ? model->start_offset : in->read_len;
            if (nseg > 2) {
               int inner = 0;
               for (int k = 1; k + 1 < nseg; ++k) inner += (int)seg_lens[k];
               lmin = std::max(lmin, inner);
            }
            weight = 0.0;
            if (in->defer_weights && nseg <= 32) {
               // leave the sum to the GPU (weights_kernel): record what it needs
               ent.seg_ptr = (int64_t)tb->w_pool.size();
               ent.nseg = (uint8_t)nseg;
               for (uint32_t x : implicit) ent.mask |= 1u << x;
               tb->w_pool.insert(tb->w_pool.end(), seg_lens.begin(), seg_lens.end());
            } else {
               for (int fl = lmin; fl <= lmax; ++fl) {
                  const double le_eff = effective_len(seg_lens.data(), nseg, implicit.data(), (int)implicit.size(), fl, in->read_len);
                  weight += insert_pdf(*model, (uint32_t)fl) * le_eff / (tb->iso_len[t] - fl + 1);
               }
            }
         }
         ent.w = weight;
         row_entries[cid].push_back(ent);
      }
   }
   tb->row_ptr.assign(1, 0);
   for (int c = 0; c < R; ++c) {
      for (auto const& e : row_entries[c]) {   // t ascending by construction
         tb->col.push_back(e.t);
         tb->alpha.push_back(e.w);
         if (in->defer_weights) {
            tb->w_seg_ptr.push_back(e.seg_ptr);
            tb->w_nseg.push_back(e.nseg);
            tb->w_mask.push_back(e.mask);
            tb->w_len.push_back(tb->iso_len[e.t]);
         }
      }
      tb->row_ptr.push_back((int64_t)tb->col.size());
   }
   *out = tb;
   return SBQ_SUCCESS;
}

void sbq_table_free(sbq_table* t) { delete t; }

int sbq_table_locus(const sbq_table* t, sbq_locus* out) {
   if (!t || !out) return SBQ_ERR_INVALID;
   out->n_iso = (int32_t)t->iso_len.size();
   out->n_row = (int32_t)t->class_count.size();
   out->row_ptr = t->row_ptr.data();
   out->col = t->col.data();
   out->alpha = t->alpha.data();
   out->count = t->class_count.data();
   out->iso_len = t->iso_len.data();
   return SBQ_SUCCESS;
}

int sbq_table_get_dims(const sbq_table* t, sbq_table_dims* out) {
   if (!t || !out) return SBQ_ERR_INVALID;
   out->n_seg = (int32_t)t->seg_left.size();
   out->n_class = (int32_t)t->class_count.size();
   out->n_iso = (int32_t)t->iso_len.size();
   out->nnz = (int64_t)t->col.size();
   out->n_coord = (int64_t)t->class_coord.size();
   out->n_dropped_hits = t->n_dropped;
   return SBQ_SUCCESS;
}

int sbq_table_segments(const sbq_table* t, uint32_t* l, uint32_t* r) {
   if (!t) return SBQ_ERR_INVALID;
   if (l) memcpy(l, t->seg_left.data(), t->seg_left.size() * sizeof(uint32_t));
   if (r) memcpy(r, t->seg_right.data(), t->seg_right.size() * sizeof(uint32_t));
   return SBQ_SUCCESS;
}

int sbq_table_iso_segments(const sbq_table* t, int32_t* ptr, int32_t* seg) {
   if (!t) return SBQ_ERR_INVALID;
   if (ptr) memcpy(ptr, t->iso_seg_ptr.data(), t->iso_seg_ptr.size() * sizeof(int32_t));
   if (seg) memcpy(seg, t->iso_seg.data(), t->iso_seg.size() * sizeof(int32_t));
   return SBQ_SUCCESS;
}

int sbq_table_weight_desc(const sbq_table* t, sbq_weight_desc* out) {
   if (!t || !out) return SBQ_ERR_INVALID;
   out->n_entry = (int64_t)t->w_seg_ptr.size();
   out->seg_ptr = t->w_seg_ptr.data();
   out->n_seg = t->w_nseg.data();
   out->implicit_mask = t->w_mask.data();
   out->iso_len = t->w_len.data();
   out->pool = t->w_pool.data();
   out->n_pool = (int64_t)t->w_pool.size();
   return SBQ_SUCCESS;
}

int sbq_table_hit_classes(const sbq_table* t, int32_t* hit_class) {
   if (!t || !hit_class) return SBQ_ERR_INVALID;
   memcpy(hit_class, t->hit_class.data(), t->hit_class.size() * sizeof(int32_t));
   return SBQ_SUCCESS;
}

int sbq_table_classes(const sbq_table* t, int32_t* cptr, int32_t* coord, int32_t* count, float* mass, int32_t* nfrag) {
   if (!t) return SBQ_ERR_INVALID;
   if (cptr) memcpy(cptr, t->class_coord_ptr.data(), t->class_coord_ptr.size() * sizeof(int32_t));
   if (coord) memcpy(coord, t->class_coord.data(), t->class_coord.size() * sizeof(int32_t));
   if (count) memcpy(count, t->class_count.data(), t->class_count.size() * sizeof(int32_t));
   if (mass) memcpy(mass, t->class_mass.data(), t->class_mass.size() * sizeof(float));
   if (nfrag) memcpy(nfrag, t->class_nfrag.data(), t->class_nfrag.size() * sizeof(int32_t));
   return SBQ_SUCCESS;
}

}  // extern "C"
