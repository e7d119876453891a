"""strawberry_b200 - B200-native quantification EM for ruolin/strawberry (host-side Python mirror over libsbq.so).

api        ctypes binding of include/sbq.h: Quantifier, EmSolver (mirror of the reference's EmSolver)
builder    ctypes binding of include/sbq_builder.h: host class-table builder (LocusContext constructor equivalent)
partition  multi-GPU host logic (LPT partition of loci, the single TPM all-reduce)
synth      seeded synthetic locus batches of the BASELINE configurations
build      compiles libsbq.so in-tree with nvcc for sm_100a
"""
__version__ = "0.1.0"
