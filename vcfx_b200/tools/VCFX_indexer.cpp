// VCFX_indexer — drop-in replacement for the reference tool of the same name (src/VCFX_indexer/VCFX_indexer.cpp): same
// flags, messages, exit codes and output bytes; createVCFIndexMmap (:205-322) / createVCFIndex (:329-443) run on the GPU
// via libvcfx_cuda (VCFX_OP_INDEX).  SURVEY.md §8 f4: the byte offsets fall out of the line table the scan kernel keeps.
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
    fputs("VCFX_indexer\n"
          "Usage: VCFX_indexer [options] [input.vcf]\n"
          "       VCFX_indexer [options] < input.vcf\n\n"
          "Description:\n"
          "  Reads a VCF from file argument or stdin and writes a 3-column index\n"
          "  (CHROM, POS, FILE_OFFSET) to stdout. FILE_OFFSET is the byte offset\n"
          "  from the start of the file to the beginning of each variant line.\n"
          "  When a file is provided directly, uses memory-mapped I/O for faster processing.\n\n"
          "Options:\n"
          "  -h, --help    Show this help message\n\n"
          "Example:\n"
          "  VCFX_indexer input.vcf > index.tsv       # Fast memory-mapped mode\n"
          "  VCFX_indexer < input.vcf > index.tsv     # Stdin mode\n", stdout);
}

int main(int argc, char *argv[]) {
    // vcfx::handle_common_flags (include/vcfx_core.h:31-67): --help / -h anywhere first, then --version / -v
    for (int i = 1; i < argc; ++i) if (!strcmp(argv[i], "--help") || !strcmp(argv[i], "-h")) { display_help(); return 0; }
    for (int i = 1; i < argc; ++i) if (!strcmp(argv[i], "--version") || !strcmp(argv[i], "-v")) { puts("VCFX_indexer version 1.1.4"); return 0; }
    static struct option long_opts[] = {{"help", no_argument, nullptr, 'h'}, {nullptr, 0, nullptr, 0}};
    int c;
    while ((c = getopt_long(argc, argv, "h", long_opts, nullptr)) != -1) {
        if (c == 'h') { display_help(); return 0; }
        display_help(); return 1;
    }
    vcfxh::RunOptions opt;
    opt.op = VCFX_OP_INDEX;
    opt.rule = vcfxh::HeaderRule::IndexChrom;
    opt.header_row = "CHROM\tPOS\tFILE_OFFSET\n";
    vcfxh::Totals tot;
    std::string err;
    int rc;
    if (optind < argc) {
        const char *input = argv[optind];
        int fd = open(input, O_RDONLY);
        if (fd < 0) { fprintf(stderr, "Error: cannot open file: %s\n", input); return 1; }
        struct stat st;
        if (fstat(fd, &st) < 0) { close(fd); fprintf(stderr, "Error: cannot stat file: %s\n", input); return 1; }
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
    if (tot.index_warned) fputs("Error: no #CHROM header found before variant lines.\n", stderr);
    vcfxh::finish(0);
}
