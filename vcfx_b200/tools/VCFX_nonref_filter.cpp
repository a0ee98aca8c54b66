// VCFX_nonref_filter — drop-in replacement for the reference tool of the same name
// (src/VCFX_nonref_filter/VCFX_nonref_filter.cpp): same flags, messages, exit codes and output bytes;
// filterNonRefMmap (:458-548) / filterNonRef (:553-631) run on the GPU via libvcfx_cuda (VCFX_OP_NONREF_FILTER).
// a sibling tool on the same scan -> GT -> per-line predicate shape as the five of the hot path.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <getopt.h>
#include <string>
#include <sys/stat.h>
#include <unistd.h>

#include "vcfx_host.h"

static void display_help() {
    fputs("VCFX_nonref_filter: Exclude variants if all samples are homozygous reference.\n\n"
          "Usage:\n"
          "  VCFX_nonref_filter [options] [input.vcf]\n"
          "  VCFX_nonref_filter [options] < input.vcf > output.vcf\n\n"
          "Options:\n"
          "  -h, --help          Show this help message\n"
          "  -i, --input FILE    Input VCF file (uses fast memory-mapped I/O)\n\n"
          "Description:\n"
          "  Reads VCF lines. For each variant, we check each sample's genotype. If a\n"
          "  genotype is polyploid, all alleles must be '0'. If a genotype is missing\n"
          "  or partial, we consider it not guaranteed hom-ref => keep variant.\n"
          "  If we find at least one sample not hom-ref, we print the variant. Otherwise,\n"
          "  we skip it.\n\n"
          "Performance:\n"
          "  File input (-i) uses memory-mapped I/O for 100-1000x faster processing\n"
          "  compared to stdin. Features include:\n"
          "  - SIMD-optimized line scanning (AVX2/SSE2)\n"
          "  - Zero-copy string parsing with string_view\n"
          "  - 1MB output buffering\n"
          "  - Direct GT field extraction (avoids full sample parsing)\n"
          "  - Early termination on first non-homref sample\n\n"
          "Examples:\n"
          "  VCFX_nonref_filter -i input.vcf > filtered.vcf    # Fast (mmap)\n"
          "  VCFX_nonref_filter input.vcf > filtered.vcf       # Fast (mmap)\n"
          "  VCFX_nonref_filter < input.vcf > filtered.vcf     # Slower (stdin)\n\n", stdout);
}

int main(int argc, char *argv[]) {
    // vcfx::handle_common_flags (include/vcfx_core.h:31-67): --help / -h anywhere first, then --version / -v
    for (int i = 1; i < argc; ++i) if (!strcmp(argv[i], "--help") || !strcmp(argv[i], "-h")) { display_help(); return 0; }
    for (int i = 1; i < argc; ++i) if (!strcmp(argv[i], "--version") || !strcmp(argv[i], "-v")) { puts("VCFX_nonref_filter version 1.1.4"); return 0; }
    const char *input = nullptr;
    bool show_help = false;
    static struct option long_opts[] = {{"help", no_argument, nullptr, 'h'}, {"input", required_argument, nullptr, 'i'}, {nullptr, 0, nullptr, 0}};
    int c;
    while ((c = getopt_long(argc, argv, "hi:", long_opts, nullptr)) != -1) {
        switch (c) {
        case 'h': show_help = true; break;
        case 'i': input = optarg; break;
        default: show_help = true;
        }
    }
    if ((!input || !*input) && optind < argc) input = argv[optind];
    if (show_help) { display_help(); return 0; }

    vcfxh::RunOptions opt;
    opt.op = VCFX_OP_NONREF_FILTER;
    opt.rule = vcfxh::HeaderRule::ChromHeader;       // data lines in front of the first "#CHROM" line pass (with a warning)
    vcfxh::Totals tot;
    std::string err;
    int rc;
    if (input && *input && strcmp(input, "-") != 0) {
        int fd = open(input, O_RDONLY);
        struct stat st;
        if (fd < 0 || fstat(fd, &st) < 0) { fprintf(stderr, "Error: Cannot open file: %s\n", input); return 0; }   // (:459-463: message, exit code 0)
        opt.mode = VCFX_MODE_FILE;
        vcfxh::Source src(fd);
        rc = vcfxh::run_stream(src, opt, tot, err);
        close(fd);
    } else {
        opt.mode = VCFX_MODE_STDIN;
        vcfxh::Source src(0);
        rc = vcfxh::run_stream(src, opt, tot, err);
    }
    if (rc != VCFX_OK) { fprintf(stderr, "Error: %s\n", err.c_str()); vcfxh::finish(1); }
    for (uint64_t i = 0; i < tot.pre_header; ++i) fputs("Warning: VCF data line encountered before #CHROM. Passing line.\n", stderr);
    vcfxh::finish(0);
}
