// TEST INFRASTRUCTURE ONLY — never linked into the product (libsbq.so).
//
// Seam harness around the UNMODIFIED reference (ruolin/strawberry v1.1.2). This TU is ours; it
// #includes the reference headers from /root/reference/include at build time (oracle/Makefile) and
// is linked with the reference's own object files into oracle/_ref/libsbref.so. It exposes, over a
// plain C ABI callable from ctypes:
//
//   ref_em_solve        narrow seam: EmSolver::init + EmSolver::run     (src/estimate.cpp:366-488)
//   ref_em_solve_batch  same over a flat batch of CSR loci on a thread pool (the CPU baseline arm)
//   ref_locus_context   wide seam: LocusContext ctor + estimate_abundances
//                       (include/estimate.hpp:61-109, src/estimate.cpp:279-364), dumped as JSON
//
// Compiled with -fno-access-control so that private members (LocusContext::exon_bins,
// HitCluster::_uniq_hits, PairedHit::_collapse_mass, Sample::_total_mapped_reads) can be filled and
// dumped without touching reference sources.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sstream>
#include <thread>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <vector>
#include <memory>
#include "alignments.h"
#include "estimate.hpp"

namespace {

// A HitFactory with no file behind it; LocusContext only reads _reads_table.read_len_mode().
class NullHitFactory : public HitFactory {
 public:
   NullHitFactory(ReadTable& rt, RefSeqTable& ref) : HitFactory(rt, ref, "oracle") {}
   bool recordsRemain() const override { return false; }
   bool nextRecord(const char*&, size_t&) override { return false; }
   bool getHitFromBuf(const char*, ReadHit&) override { return false; }
   void undo_hit() override {}
   bool inspect_header() override { return true; }
   void reset() override {}
   void return2Pos(int64_t) override {}
   int64_t getCurrPos() override { return 0; }
};

void jdouble(std::ostringstream& os, double v) {
   char buf[64];
   if (v != v) { os << "\"nan\""; return; }
   if (v > 1.7e308) { os << "\"inf\""; return; }
   if (v < -1.7e308) { os << "\"-inf\""; return; }
   snprintf(buf, sizeof buf, "%.17g", v);
   os << buf;
}

}  // namespace

extern "C" {

// Narrow seam. alpha is dense row-major R x T. Returns bit0 = init() result, bit1 = run() result.
int ref_em_solve(int T, int R, const int* n, const double* alpha, double* theta_out) {
   std::vector<int> count(n, n + R);
   std::vector<std::vector<double>> model(R, std::vector<double>(T));
   for (int i = 0; i < R; ++i)
      for (int j = 0; j < T; ++j) model[i][j] = alpha[(size_t)i * T + j];
   EmSolver em;
   int rc = 0;
   bool ok = em.init(T, count, model);
   if (ok) {
      rc |= 1;
      if (em.run()) rc |= 2;
   }
   for (int j = 0; j < T; ++j) theta_out[j] = em._theta[j];
   return rc;
}

// Batched narrow seam over flat CSR loci, densified per locus exactly as
// LocusContext::estimate_abundances does (src/estimate.cpp:283-296) before calling EmSolver.
// loc_row_off/loc_iso_off: n_loci+1 prefix offsets; row_ptr: global, total_rows+1 entries.
// Returns wall seconds of the solve region (densify + init + run), threads = n_threads.
double ref_em_solve_batch(long n_loci, const long* loc_row_off, const long* loc_iso_off,
                          const long* row_ptr, const int* col, const double* alpha, const int* count,
                          double* theta_out, int* rc_out, int n_threads) {
   // Loci are handed out largest first (dense cost R x T, what EmSolver works on): with the natural order a big locus picked
   // up late leaves the other threads idle at the end. The sort is outside the timed region - the most favourable
   // schedule a thread pool over independent loci can get.
   std::vector<long> order(n_loci);
   for (long l = 0; l < n_loci; ++l) order[l] = l;
   std::stable_sort(order.begin(), order.end(), [&](long a, long b) {
      const long ca = (loc_row_off[a + 1] - loc_row_off[a]) * (loc_iso_off[a + 1] - loc_iso_off[a]);
      const long cb = (loc_row_off[b + 1] - loc_row_off[b]) * (loc_iso_off[b + 1] - loc_iso_off[b]);
      return ca > cb;
   });
   std::atomic<long> next(0);
   auto work = [&]() {
      for (;;) {
         const long w = next.fetch_add(1);
         if (w >= n_loci) break;
         const long l = order[w];
         long r0 = loc_row_off[l], r1 = loc_row_off[l + 1];
         long t0 = loc_iso_off[l], T = loc_iso_off[l + 1] - t0;
         long R = r1 - r0;
         std::vector<int> n(count + r0, count + r1);
         std::vector<std::vector<double>> model(R, std::vector<double>(T, 0.0));
         for (long i = 0; i < R; ++i)
            for (long k = row_ptr[r0 + i]; k < row_ptr[r0 + i + 1]; ++k) model[i][col[k]] = alpha[k];
         EmSolver em;
         int rc = 0;
         if (em.init((int)T, n, model)) {
            rc |= 1;
            if (em.run()) rc |= 2;
         }
         for (long j = 0; j < T; ++j) theta_out[t0 + j] = em._theta[j];
         if (rc_out) rc_out[l] = rc;
      }
   };
   auto t_begin = std::chrono::steady_clock::now();
   if (n_threads <= 1) {
      work();
   } else {
      std::vector<std::thread> pool;
      for (int t = 0; t < n_threads; ++t) pool.emplace_back(work);
      for (auto& th : pool) th.join();
   }
   return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
}

// Wide seam. Hits are given as mate pairs (slot 2h = left mate, 2h+1 = right mate; an empty CIGAR
// range means the mate is absent), in the order LocusContext would see them in
// HitCluster::uniq_hits(). Transcripts are alternating MATCH/INTRON feature lists.
// Writes a JSON dump to `out`; returns its length, or -needed if out_cap is too small.
long ref_locus_context(int use_emp, const int* frag_lens, int n_frag_lens, double mean, double sd,
                       int read_len, int long_read, double min_iso_frac, int eff_len_norm,
                       int total_mapped_reads,
                       int n_iso, const int* iso_feat_ptr, const unsigned* iso_feat_off,
                       const int* iso_feat_len, const int* iso_feat_code,
                       int n_hits, const double* hit_mass, const int* read_cig_ptr,
                       const unsigned* read_pos, const int* cig_type, const int* cig_len,
                       char* out, long out_cap) {
   long_read_sample = long_read != 0;
   kMinIsoformFrac = min_iso_frac;
   effective_len_norm = eff_len_norm != 0;
   infer_the_other_end = false;
   filter_by_expression = true;

   ReadTable rt;
   rt._read_len_abs[(uint)read_len] = 1;
   RefSeqTable ref(true);
   ref.set_id("chr1");
   std::shared_ptr<HitFactory> hf(new NullHitFactory(rt, ref));
   Sample sample(hf);
   if (use_emp)
      sample._insert_size_dist.reset(new InsertSize(std::vector<int>(frag_lens, frag_lens + n_frag_lens)));
   else
      sample._insert_size_dist.reset(new InsertSize(mean, sd));
   sample._total_mapped_reads = total_mapped_reads;

   std::vector<Contig> transcripts;
   for (int t = 0; t < n_iso; ++t) {
      std::vector<GenomicFeature> feats;
      for (int k = iso_feat_ptr[t]; k < iso_feat_ptr[t + 1]; ++k)
         feats.push_back(GenomicFeature((Match_t)iso_feat_code[k], iso_feat_off[k], iso_feat_len[k]));
      Contig c(0, (ReadID)(t + 1), Strand_t::StrandPlus, 1.0, feats, true);
      c.annotated_trans_id("T" + std::to_string(t));
      c.parent_id() = "G";
      transcripts.push_back(c);
   }

   std::shared_ptr<HitCluster> cluster(new HitCluster());
   cluster->_ref_id = 0;
   for (int h = 0; h < n_hits; ++h) {
      ReadHitPtr mates[2];
      for (int s = 0; s < 2; ++s) {
         int c0 = read_cig_ptr[2 * h + s], c1 = read_cig_ptr[2 * h + s + 1];
         if (c0 == c1) continue;
         std::vector<CigarOp> cig;
         uint span = 0;
         for (int k = c0; k < c1; ++k) {
            cig.push_back(CigarOp((CigarOpCode)cig_type[k], (uint32_t)cig_len[k]));
            if (cig_type[k] == MATCH || cig_type[k] == REF_SKIP || cig_type[k] == DEL) span += cig_len[k];
         }
         uint l = read_pos[2 * h + s];
         mates[s].reset(new ReadHit((ReadID)(h + 1), "r" + std::to_string(h),
                                    GenomicInterval(0, l, l + span - 1, Strand_t::StrandPlus), cig, 0,
                                    1 /*partner pos != 0 => not a singleton*/, 0, 1, s == 0 ? 99u : 147u,
                                    0.5, NULL));
      }
      PairedHit ph(mates[0], mates[1]);
      ph._collapse_mass = hit_mass[h];
      cluster->_uniq_hits.push_back(ph);
   }

   FILE* log = fopen("/dev/null", "w");
   std::ostringstream os;
   os << "{";
   // a11: Contig(PairedHit) feature lists (src/contig.cpp:216-267)
   os << "\"hits\":[";
   for (int h = 0; h < n_hits; ++h) {
      Contig c(cluster->_uniq_hits[h]);
      if (h) os << ",";
      os << "{\"ref_id\":" << c.ref_id() << ",\"mass\":";
      jdouble(os, (double)c.mass());
      os << ",\"feats\":[";
      for (size_t k = 0; k < c._genomic_feats.size(); ++k) {
         const GenomicFeature& f = c._genomic_feats[k];
         if (k) os << ",";
         os << "[" << (int)f._match_op._code << "," << f._genomic_offset << "," << f._match_op._len << "]";
      }
      os << "]}";
   }
   os << "],";

   auto t_ctor0 = std::chrono::steady_clock::now();
   LocusContext lc(sample, log, cluster, transcripts);
   const double ctor_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_ctor0).count();
   os << "\"ctor_seconds\":" << ctor_seconds << ",";

   os << "\"segs\":[";
   for (size_t i = 0; i < lc._exon_segs.size(); ++i) {
      if (i) os << ",";
      os << "[" << lc._exon_segs[i].left() << "," << lc._exon_segs[i].right() << "]";
   }
   os << "],\"iso_len\":[";
   for (size_t t = 0; t < lc._transcripts.size(); ++t) {
      if (t) os << ",";
      os << lc._transcripts[t]._length;
   }
   os << "],\"iso_segs\":[";
   for (size_t t = 0; t < lc._transcripts.size(); ++t) {
      if (t) os << ",";
      os << "[";
      for (size_t k = 0; k < lc._transcripts[t]._exon_segs.size(); ++k) {
         if (k) os << ",";
         os << "[" << lc._transcripts[t]._exon_segs[k].left() << "," << lc._transcripts[t]._exon_segs[k].right() << "]";
      }
      os << "]";
   }
   os << "],\"classes\":[";
   for (size_t c = 0; c < lc.exon_bins.size(); ++c) {
      const ExonBin& eb = lc.exon_bins[c];
      if (c) os << ",";
      os << "{\"coords\":[";
      bool first = true;
      for (auto const& p : eb._coords) {
         if (!first) os << ",";
         first = false;
         os << "[" << p.first << "," << p.second << "]";
      }
      os << "],\"nfrags\":" << eb._frags.size() << ",\"count_f\":";
      jdouble(os, (double)eb.read_count());
      int cnt = eb.read_count();   // src/estimate.cpp:288: n[i] = bin.read_count()  (float -> int)
      os << ",\"count\":" << cnt << ",\"weights\":{";
      first = true;
      for (auto const& w : eb._bin_weight_map) {
         if (!first) os << ",";
         first = false;
         os << "\"" << w.first << "\":";
         jdouble(os, w.second);
      }
      os << "},\"frag_lens\":{";
      first = true;
      for (auto const& fl : eb._iso_2_frag_lens) {
         if (!first) os << ",";
         first = false;
         os << "\"" << fl.first << "\":[";
         for (size_t k = 0; k < fl.second.size(); ++k) {
            if (k) os << ",";
            os << "[" << fl.second[k].first << ",";
            jdouble(os, (double)fl.second[k].second);
            os << "]";
         }
         os << "]";
      }
      os << "}}";
   }
   os << "],\"iso2bins\":{";
   {
      bool first = true;
      for (auto const& kv : lc.iso_2_bins_map) {
         if (!first) os << ",";
         first = false;
         os << "\"" << kv.first << "\":[";
         bool f2 = true;
         for (int b : kv.second) {
            if (!f2) os << ",";
            f2 = false;
            os << b;
         }
         os << "]";
      }
   }
   os << "},";

   // theta at full precision: the same densify + EmSolver calls estimate_abundances makes
   // (src/estimate.cpp:283-308); estimate_abundances itself only logs theta with %f.
   {
      size_t nrow = lc.exon_bins.size(), niso = lc._transcripts.size();
      std::vector<int> n(nrow);
      std::vector<std::vector<double>> alpha(nrow, std::vector<double>(niso, 0.0));
      for (size_t i = 0; i < nrow; ++i) {
         n[i] = lc.exon_bins[i].read_count();
         for (auto const& w : lc.exon_bins[i]._bin_weight_map) alpha[i][w.first] = w.second;
      }
      EmSolver em;
      bool ok = em.init((int)niso, n, alpha);
      bool ran = ok ? em.run() : false;
      os << "\"em_init\":" << (ok ? "true" : "false") << ",\"em_run\":" << (ran ? "true" : "false")
         << ",\"theta\":[";
      for (size_t j = 0; j < em._theta.size(); ++j) {
         if (j) os << ",";
         jdouble(os, em._theta[j]);
      }
      os << "],";
   }

   bool success = lc.estimate_abundances();
   os << "\"success\":" << (success ? "true" : "false") << ",\"isoforms\":[";
   if (success) {
      for (size_t t = 0; t < lc._transcripts.size(); ++t) {
         const Isoform& iso = lc._transcripts[t];
         if (t) os << ",";
         os << "{\"id\":" << iso.id() << ",\"fpkm\":";
         jdouble(os, iso._FPKM);
         os << ",\"frac\":";
         jdouble(os, iso._frac);
         os << ",\"fpkm_s\":\"" << iso._FPKM_s << "\",\"frac_s\":\"" << iso._frac_s << "\"}";
      }
   }
   os << "]}";
   fclose(log);

   std::string s = os.str();
   if ((long)s.size() + 1 > out_cap) return -(long)(s.size() + 1);
   memcpy(out, s.c_str(), s.size() + 1);
   return (long)s.size();
}

}  // extern "C"
