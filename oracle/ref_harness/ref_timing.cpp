// ref_timing.cpp - TEST INFRASTRUCTURE: wall-clock timers around the UNMODIFIED reference's quantification entry points.
//
// oracle/_ref/strawberry_ref_timed is the reference program linked from its own objects, with three symbols of estimate.o
// renamed by objcopy (no source patch): LocusContext::assign_exon_bin, ::set_theory_bin_weight and ::estimate_abundances become
// sbqref_*; this TU defines the original names as thin wrappers that time the call and forward to the renamed reference code.
// With SBQ_TIMING=1 the totals are printed at exit, in the same format the libsbq integration prints (integration/estimate_sbq.cpp),
// so that "quantification-only" time of both binaries can be tabulated next to their wall time.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "estimate.hpp"

extern "C" {
void sbqref_assign_exon_bin(LocusContext*, const std::vector<Contig>*, const std::vector<GenomicFeature>*);
void sbqref_set_theory_bin_weight(LocusContext*);
bool sbqref_estimate_abundances(LocusContext*);
}

namespace {
struct Timers {
   std::atomic<long long> assign_ns{0}, weight_ns{0}, est_ns{0}, loci{0};
   ~Timers() {
      if (!getenv("SBQ_TIMING")) return;
      fprintf(stderr, "SBQ_TIMING ref loci %lld class_table_ms %.3f (assign_exon_bin %.3f + set_theory_bin_weight %.3f) estimate_abundances_ms %.3f\n", loci.load(),
              (assign_ns + weight_ns) / 1e6, assign_ns / 1e6, weight_ns / 1e6, est_ns / 1e6);
   }
} g_tm;
long long now_ns() { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

void LocusContext::assign_exon_bin(const std::vector<Contig>& hits, const std::vector<GenomicFeature>& exon_segs) {
   const long long t = now_ns();
   sbqref_assign_exon_bin(this, &hits, &exon_segs);
   g_tm.assign_ns += now_ns() - t;
}
void LocusContext::set_theory_bin_weight() {
   const long long t = now_ns();
   sbqref_set_theory_bin_weight(this);
   g_tm.weight_ns += now_ns() - t;
}
bool LocusContext::estimate_abundances() {
   const long long t = now_ns();
   const bool ok = sbqref_estimate_abundances(this);
   g_tm.est_ns += now_ns() - t;
   g_tm.loci += 1;
   return ok;
}
