// `extract_ref` — drop-in for LocalHGT's scripts/extract_ref (built there by Makefile:5-6 from
// src/extract_ref_normal_peak.cpp; exec'd by scripts/pipeline.sh:35).  Same 12 positional arguments;
// the whole body lives behind the C ABI so bindings and this binary run identical code.
#include "../../include/lhgt.h"

int main(int argc, char** argv) { return lhgt_main(argc, argv); }
