// get_bed_file.py + `samtools faidx -r` of LocalHGT's pipeline.sh:36-37 in one step.
#include "../../include/lhgt.h"
#include <cstdio>
int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: extract_regions <ref.fa> <interval_file> [extracted_ref.fasta]\n"
                        "       writes <interval_file>.bed (get_bed_file.py) and, when named, the extracted reference\n");
        return 2;
    }
    long len = 0;
    int rc = lhgt_extract_regions_files(argv[1], argv[2], argc > 3 ? argv[3] : nullptr, &len);
    if (rc) { fprintf(stderr, "extract_regions: error %d: %s\n", rc, lhgt_last_error()); return 1; }
    printf("extracted ref length is: %ld\n", len);       // get_bed_file.py:62
    return 0;
}
