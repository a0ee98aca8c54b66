// VCFX_allele_freq_calc — drop-in replacement for the reference tool of the same name
// (src/VCFX_allele_freq_calc/VCFX_allele_freq_calc.cpp): same flags, messages, exit codes and
// output bytes in both input modes; processMmap / processStdin run on the GPU via libvcfx_cuda.
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <getopt.h>
#include <string>
#include <sys/stat.h>
#include <unistd.h>

#include "vcfx_host.h"

static void print_help() {
    fputs("VCFX_allele_freq_calc v1.1 - High-performance allele frequency calculator\n\n"
          "Usage:\n"
          "  VCFX_allele_freq_calc [OPTIONS] [input.vcf]\n"
          "  VCFX_allele_freq_calc [OPTIONS] < input.vcf > output.tsv\n\n"
          "Options:\n"
          "  -i, --input FILE   Input VCF file (uses memory-mapping for best performance)\n"
          "  -q, --quiet        Suppress informational messages\n"
          "  -h, --help         Display this help message and exit\n"
          "  -v, --version      Show program version and exit\n\n"
          "Description:\n"
          "  Calculates allele frequency for each variant in a VCF file.\n"
          "  Allele frequency is computed as (#ALT alleles) / (total #alleles),\n"
          "  counting any non-zero numeric allele (1,2,3,...) as ALT.\n\n"
          "Output Format:\n"
          "  CHROM  POS  ID  REF  ALT  Allele_Frequency\n\n"
          "Performance:\n"
          "  - Memory-mapped I/O: Use -i flag for ~15-20x faster processing\n"
          "  - SIMD acceleration for line/field scanning\n"
          "  - Zero-copy parsing with string_view\n\n"
          "Examples:\n"
          "  VCFX_allele_freq_calc -i input.vcf > frequencies.tsv\n"
          "  VCFX_allele_freq_calc < input.vcf > frequencies.tsv\n", stdout);
}

static const char HEADER[] = "CHROM\tPOS\tID\tREF\tALT\tAllele_Frequency\n";

int main(int argc, char *argv[]) {
    const char *input = nullptr;
    bool quiet = false;
    static struct option long_opts[] = {{"input", required_argument, nullptr, 'i'}, {"quiet", no_argument, nullptr, 'q'},
                                        {"help", no_argument, nullptr, 'h'}, {"version", no_argument, nullptr, 'v'},
                                        {nullptr, 0, nullptr, 0}};
    int c;
    while ((c = getopt_long(argc, argv, "i:qhv", long_opts, nullptr)) != -1) {
        switch (c) {
        case 'i': input = optarg; break;
        case 'q': quiet = true; break;
        case 'h': print_help(); return 0;
        case 'v': puts("VCFX_allele_freq_calc v1.1"); return 0;
        default: print_help(); return 1;
        }
    }
    if (!input && optind < argc) input = argv[optind];

    vcfxh::RunOptions opt;
    opt.op = VCFX_OP_ALLELE_FREQ;
    opt.rule = vcfxh::HeaderRule::ChromHeader;
    vcfxh::Totals tot;
    std::string err;
    int rc;
    if (input) {
        int fd = open(input, O_RDONLY);
        struct stat st;
        if (fd < 0 || fstat(fd, &st) < 0) { fprintf(stderr, "Error: Cannot open file: %s\n", input); return 1; }
        if (!quiet) fprintf(stderr, "Processing %s (%llu MB)\n", input, (unsigned long long)st.st_size / (1024 * 1024));
        opt.mode = VCFX_MODE_FILE;
        vcfxh::write_all(1, HEADER, sizeof HEADER - 1);
        vcfxh::Source src(fd);
        rc = vcfxh::run_stream(src, opt, tot, err);
        close(fd);
    } else {
        vcfxh::Source src(0);
        if (src.at_eof_initially()) { print_help(); return 1; }
        opt.mode = VCFX_MODE_STDIN;
        vcfxh::write_all(1, HEADER, sizeof HEADER - 1);
        rc = vcfxh::run_stream(src, opt, tot, err);
    }
    if (rc != VCFX_OK) { fprintf(stderr, "Error: %s\n", err.c_str()); vcfxh::finish(1); }
    if (!quiet) {
        for (uint64_t i = 0; i < tot.pre_header; ++i) fputs("Warning: Data line encountered before #CHROM header. Skipping.\n", stderr);
        if (!input) for (uint64_t i = 0; i < tot.short_lines; ++i) fputs("Warning: Skipping invalid VCF line (fewer than 9 fields).\n", stderr);
        if (input) fprintf(stderr, "Processed %llu variants from %llu data lines\n", (unsigned long long)tot.rows, (unsigned long long)tot.data_lines);
    }
    vcfxh::finish(0);
}
