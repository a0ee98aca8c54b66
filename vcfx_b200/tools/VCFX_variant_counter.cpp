// VCFX_variant_counter — drop-in replacement for the reference tool of the same name
// (src/VCFX_variant_counter/VCFX_variant_counter.cpp): same flags, messages and exit codes; the
// per-line loop (countVariantsMmap / countVariants / countVariantsGzip) runs on the GPU through
// libvcfx_cuda instead.
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <getopt.h>
#include <string>
#include <unistd.h>

#include "vcfx_host.h"

static void display_help() {
    fputs("VCFX_variant_counter: Counts the total number of valid variants in a VCF.\n\n"
          "Usage:\n"
          "  VCFX_variant_counter [options] [input.vcf]\n"
          "  VCFX_variant_counter [options] < input.vcf\n\n"
          "Options:\n"
          "  -h, --help        Show this help.\n"
          "  -s, --strict      Fail on any data line with <8 columns.\n\n"
          "Description:\n"
          "  Reads a VCF from file argument or stdin. For each data line,\n"
          "  we check if it has >=8 columns; if it does, we count it; if fewer columns:\n"
          "   * if --strict => we exit with error,\n"
          "   * otherwise => we skip with a warning.\n"
          "  When a file is provided directly, uses memory-mapped I/O for faster processing.\n"
          "  Finally, we print 'Total Variants: X'.\n\n"
          "Example:\n"
          "  VCFX_variant_counter input.vcf          # Fast memory-mapped mode\n"
          "  VCFX_variant_counter < input.vcf        # Stdin mode\n"
          "  VCFX_variant_counter --strict input.vcf\n", stdout);
}

int main(int argc, char *argv[]) {
    // vcfx::handle_common_flags: --help/-h or --version/-v anywhere in argv (vcfx_core.h:31-62)
    for (int i = 1; i < argc; ++i)
        if (!strcmp(argv[i], "--help") || !strcmp(argv[i], "-h")) { display_help(); return 0; }
    for (int i = 1; i < argc; ++i)
        if (!strcmp(argv[i], "--version") || !strcmp(argv[i], "-v")) { puts("VCFX_variant_counter version 1.1.4"); return 0; }

    bool strict = false, show_help = false;
    static struct option long_opts[] = {{"help", no_argument, 0, 'h'}, {"strict", no_argument, 0, 's'}, {0, 0, 0, 0}};
    optind = 1;
    for (;;) {
        int c = getopt_long(argc, argv, "hs", long_opts, nullptr);
        if (c == -1) break;
        if (c == 's') strict = true; else show_help = true;
    }
    if (show_help) { display_help(); return 0; }

    vcfxh::RunOptions opt;
    opt.op = VCFX_OP_VARIANT_COUNT;
    opt.want_short_lines = !strict;
    opt.stop_at_first_short = strict;
    vcfxh::Totals tot;
    std::string err;
    int rc;
    if (optind < argc) {
        const char *fn = argv[optind];
        int fd = open(fn, O_RDONLY);
        if (fd < 0) { fprintf(stderr, "Error: cannot open file: %s\n", fn); return 1; }
        opt.mode = VCFX_MODE_FILE;
        vcfxh::Source src(fd);
        rc = vcfxh::run_stream(src, opt, tot, err);
        close(fd);
    } else {
        opt.mode = VCFX_MODE_STDIN;
        vcfxh::Source src(0);
        if (src.at_eof_initially()) { puts("Total Variants: 0"); return 0; }
        if (src.sniff_gzip()) {
            src.enable_gzip();
            if (src.failed()) { fputs("Error: inflateInit2 failed.\n", stderr); return 1; }
        }
        rc = vcfxh::run_stream(src, opt, tot, err);
        if (src.failed()) { fputs("Error: decompression failed.\n", stderr); vcfxh::finish(1); }
    }
    if (rc != VCFX_OK) { fprintf(stderr, "Error: %s\n", err.c_str()); vcfxh::finish(1); }
    if (strict && tot.short_lines) {
        fprintf(stderr, "Error: line %llu has <8 columns.\n", (unsigned long long)tot.first_short_line);
        vcfxh::finish(1);
    }
    for (uint64_t ln : tot.short_line_numbers)
        fprintf(stderr, "Warning: skipping line %llu with <8 columns.\n", (unsigned long long)ln);
    if (tot.short_lines > tot.short_line_numbers.size())
        fprintf(stderr, "Warning: skipping %llu more lines with <8 columns.\n",
                (unsigned long long)(tot.short_lines - tot.short_line_numbers.size()));
    printf("Total Variants: %d\n", (int)tot.rows);
    vcfxh::finish(0);
}
