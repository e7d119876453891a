// procsample_sbq.cpp - the BATCHED drop-in: a replacement for Sample::procSample (reference src/alignments.cpp:1736-1835)
// that walks the BAM exactly like the reference does, but solves the numeric part of ALL loci in ONE sbq_run.
//
// Per-locus calls (the reference's own procSample driving integration/estimate_sbq.cpp) pay a kernel launch and a PCIe
// round trip per locus and serialise the -p N workers on one context. Here every locus is only STAGED while the BAM is
// read - finalizeCluster + the LocusContext constructor (class table and weights through sbq_build_locus) + sbq_submit,
// on the reference's worker threads when -p N is given - then sbq_batch_run() does EM, FPKM / frac / low-fraction filter
// and TPM for the whole sample on the GPU(s) (SBQ_N_GPUS devices: loci partitioned by non-zeros inside libsbq, one
// ncclAllReduce for the TPM denominator), and the reference's per-locus tail (theta log lines, strings, -f TSV, GTF)
// runs from the returned arrays in the original locus order.
//
// Compiled against the UNMODIFIED reference headers. Linked with the reference's own objects; the reference's definition
// of Sample::procSample in alignments.o is weakened with objcopy (integration/Makefile) so that this one is used - no
// reference source is patched. Everything before and after the quantification (BAM / GTF parsing, clustering, collapse,
// GTF printing) is the reference's code.
#include <atomic>
#include <chrono>
#include <climits>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "alignments.h"
#include "estimate.hpp"
#include "fasta.h"

using namespace std;

void sbq_batch_begin();                       // integration/estimate_sbq.cpp
void sbq_batch_run(int total_mapped_reads);
void sbq_batch_end();

namespace {
struct Pending {
   shared_ptr<HitCluster> cluster;
   unique_ptr<LocusContext> est;
   size_t order;                              // position of the cluster in the BAM walk (output order of -p 1)
};

// ---- single-pass host pipeline. The reference decodes and clusters the BAM twice in -g mode:
// Sample::preProcess (pass 1: insert-size histogram, total mapped reads) and Sample::procSample (pass 2: quantification),
// ~85 % of its wall time. Both passes build and finalize exactly the same clusters (collapseAndFilterHits depends on the
// cluster's own reads only), so the preProcess below keeps every finalized cluster and the batched procSample consumes
// them instead of walking the BAM again. SBQ_SINGLE_PASS=0 turns the cache off (two passes, as the reference).
vector<shared_ptr<HitCluster>> g_cache;
mutex g_cache_mu;
bool g_cache_complete = false;
bool single_pass_enabled() {
   const char* e = getenv("SBQ_SINGLE_PASS");
   return no_assembly && !(e && e[0] == '0');
}
}  // namespace

// Pass 1, same traversal and side effects as the reference (src/alignments.cpp:1189-1233): per cluster finalizeCluster +
// fragLenDist (fragment-length histogram, _total_mapped_reads). Additionally the finalized clusters are kept.
void Sample::preProcess(FILE* log) {
   const RefSeqTable& ref_t = _hit_factory->_ref_table;
   const bool keep = single_pass_enabled();
   atomic<int> workers(0);
   _num_cluster = 0;
   g_cache.clear();
   g_cache_complete = false;
   while (true) {
      shared_ptr<HitCluster> cluster(new HitCluster());
      if (-1 == nextClusterRefDemand(*cluster)) break;
      if (cluster->ref_id() == -1) continue;
      cluster->_id = ++_num_cluster;
      auto work = [this, &ref_t, cluster, log, keep] {
         finalizeCluster(cluster, true);
         fragLenDist(ref_t, cluster->ref_mRNAs(), cluster, log);
         if (keep) {
            lock_guard<mutex> lk(g_cache_mu);
            g_cache.push_back(cluster);
         }
      };
      if (use_threads && num_threads > 1) {
         while (workers.load() >= num_threads) this_thread::sleep_for(chrono::microseconds(200));
         ++workers;
         thread worker([work, &workers] {
            work();
            --workers;
         });
         worker.detach();
      } else {
         work();
      }
   }
   while (workers.load() != 0) this_thread::sleep_for(chrono::microseconds(200));
   if (keep) {
      sort(g_cache.begin(), g_cache.end(), [](const shared_ptr<HitCluster>& a, const shared_ptr<HitCluster>& b) { return a->_id < b->_id; });
      g_cache_complete = true;
   }
}

void Sample::procSample(FILE* pfile, FILE* plogfile, FILE* fragfile) {
   const auto t_start = chrono::steady_clock::now();
   if (!g_cache_complete) {                   // two passes: rewind the BAM and the reference-transcript cursor like the reference
      _hit_factory->reset();
      reset_refmRNAs();
   }
   vector<Isoform> isoforms;
   isoforms.reserve(1024);
   const RefSeqTable& ref_t = _hit_factory->_ref_table;
   int current_ref_id = INT_MAX;
   if (fragfile != NULL) {
      std::vector<string> header = {"sample", "sample_frag_count", "gene_id", "gene_frag_count",
                                    "transcripts", "FPKMs", "conditional_probabilities", "class_probabilities", "path_symbol", "path_count",
                                    "path_gc_content", "path_hexmer_entropy", "gc_stretch_0.8_20", "gc_stretch_0.9_20", "gc_stretch_0.8_40",
                                    "gc_stretch_0.9_40"};
      pretty_print(fragfile, header, "\t");
   }

   sbq_batch_begin();
   vector<Pending> pending;
   mutex pending_mu;
   atomic<int> workers(0);
   size_t order = 0;
   const bool cached = g_cache_complete;      // single pass: the clusters of pass 1 are reused, already finalized
   size_t next_cached = 0;
   // stage one locus: what Sample::quantifyCluster does up to the numeric part (src/alignments.cpp:1510-1526)
   auto stage = [&](shared_ptr<HitCluster> cluster, size_t ord) {
      if (!cached) finalizeCluster(cluster, true);
      unique_ptr<LocusContext> est(new LocusContext(*this, plogfile, cluster, cluster->ref_mRNAs()));
      est->estimate_abundances();             // stage mode: CSR of the class table -> sbq_submit
      lock_guard<mutex> lk(pending_mu);
      pending.push_back(Pending{cluster, std::move(est), ord});
   };

   while (true) {
      shared_ptr<HitCluster> cluster;
      if (cached) {
         if (next_cached == g_cache.size()) break;
         cluster = std::move(g_cache[next_cached++]);
      } else {
         cluster.reset(new HitCluster());
         if (-1 == nextClusterRefDemand(*cluster)) break;
         if (cluster->ref_id() == -1) continue;
      }
      if (current_ref_id != cluster->ref_id()) {
         current_ref_id = cluster->ref_id();
         if (BIAS_CORRECTION) {
            while (workers.load() != 0) this_thread::sleep_for(chrono::milliseconds(1));
            load_chrom_fasta(current_ref_id);
         }
      }
      if (use_threads && num_threads > 1) {
         while (workers.load() >= num_threads) this_thread::sleep_for(chrono::microseconds(200));
         ++workers;
         thread worker([&, cluster, order] {
            stage(cluster, order);
            --workers;
         });
         worker.detach();
      } else {
         stage(cluster, order);
      }
      ++order;
   }
   while (workers.load() != 0) this_thread::sleep_for(chrono::microseconds(200));
   const auto t_staged = chrono::steady_clock::now();

   sbq_batch_run(total_mapped_reads());      // EM + FPKM / frac / filter + TPM for every locus, one call
   const auto t_run = chrono::steady_clock::now();

   // the reference's per-locus tail, in BAM order (the order -p 1 of the reference emits)
   sort(pending.begin(), pending.end(), [](const Pending& a, const Pending& b) { return a.order < b.order; });
   for (auto& p : pending) {
      const bool success = p.est->estimate_abundances();   // finish mode: theta log, FPKM / frac / TPM strings, low-fraction erase
      if (!success) continue;
      vector<Isoform> iso = std::move(p.est->transcripts());
      cerr << ref_t.ref_real_name(p.cluster->ref_id()) << "\t" << p.cluster->left() << "\t" << p.cluster->right()
           << " finishes abundances estimation" << endl;
      if (fragfile != NULL) printContext(*p.est, p.cluster, _fasta_getter, fragfile);
      isoforms.insert(isoforms.end(), iso.begin(), iso.end());
   }
   sbq_batch_end();
   // TPM (src/alignments.cpp:1821-1829) was computed on the device with the all-reduced FPKM sum; _TPM / _TPM_s are set
   for (const auto& iso : isoforms) {
      iso._contig.print2gtf(pfile, _hit_factory->_ref_table, iso._FPKM_s, iso._frac_s, iso._TPM_s, iso._gene_str, iso._isoform_str, iso._ref_gene_id,
                            iso._ref_gene_name);
   }
   if (getenv("SBQ_TIMING")) {
      const auto t_end = chrono::steady_clock::now();
      auto ms = [](chrono::steady_clock::time_point a, chrono::steady_clock::time_point b) { return chrono::duration<double, milli>(b - a).count(); };
      fprintf(stderr, "SBQ_TIMING procSample(batched%s) loci %zu walk+stage_ms %.3f sbq_run_ms %.3f tail+gtf_ms %.3f total_ms %.3f\n", cached ? ", single pass" : "", pending.size(),
              ms(t_start, t_staged), ms(t_staged, t_run), ms(t_run, t_end), ms(t_start, t_end));
   }
}
