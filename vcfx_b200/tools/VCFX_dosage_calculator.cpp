// VCFX_dosage_calculator — drop-in replacement for the reference tool of the same name
// (src/VCFX_dosage_calculator/VCFX_dosage_calculator.cpp): same flags, messages, exit codes and output bytes.
// The per-line work (ten columns, GT index, a dosage per sample column, the row text; :421-591 / :228-356) runs on the GPU
// via libvcfx_cuda (VCFX_OP_DOSAGE).  The host prints the header row and the messages, and ends the run at a data line that
// comes before the "#CHROM" line.  SURVEY.md §8 f2: a sibling tool on the same scan -> GT -> per-line shape as the five
// of the hot path.
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
    fputs("VCFX_dosage_calculator: Calculate genotype dosage for each variant in a VCF file.\n\n"
          "Usage:\n"
          "  VCFX_dosage_calculator [options] [input.vcf]\n"
          "  VCFX_dosage_calculator [options] < input.vcf > dosage_output.txt\n\n"
          "Options:\n"
          "  -i, --input FILE  Input VCF file (uses mmap for best performance)\n"
          "  -q, --quiet       Suppress warning messages\n"
          "  -h, --help        Display this help message and exit\n\n"
          "Description:\n"
          "  For each variant in the input VCF, the tool computes the dosage for each sample\n"
          "  based on the genotype (GT) field. Dosage is defined as the number of alternate\n"
          "  alleles (i.e. each allele > 0 counts as 1). Thus:\n"
          "    0/0  => dosage 0\n"
          "    0/1  => dosage 1\n"
          "    1/1  => dosage 2\n"
          "    1/2  => dosage 2  (each alternate, regardless of numeric value, counts as 1)\n\n"
          "Performance:\n"
          "  When using -i/--input, the tool uses memory-mapped I/O for\n"
          "  ~10-15x faster processing of large files.\n\n"
          "Example:\n"
          "  VCFX_dosage_calculator -i input.vcf > dosage_output.txt\n"
          "  VCFX_dosage_calculator < input.vcf > dosage_output.txt\n", stdout);
}

static const char HEADER_ROW[] = "CHROM\tPOS\tID\tREF\tALT\tDosages\n";

int main(int argc, char *argv[]) {
    // vcfx::handle_common_flags (include/vcfx_core.h:31-67): --help / -h anywhere first, then --version / -v
    for (int i = 1; i < argc; ++i) if (!strcmp(argv[i], "--help") || !strcmp(argv[i], "-h")) { display_help(); return 0; }
    for (int i = 1; i < argc; ++i) if (!strcmp(argv[i], "--version") || !strcmp(argv[i], "-v")) { puts("VCFX_dosage_calculator version 1.1.4"); return 0; }
    const char *input = nullptr;
    bool show_help = false, quiet = false;
    static struct option long_opts[] = {{"help", no_argument, nullptr, 'h'}, {"input", required_argument, nullptr, 'i'},
                                        {"quiet", no_argument, nullptr, 'q'}, {nullptr, 0, nullptr, 0}};
    int c;
    while ((c = getopt_long(argc, argv, "hi:q", long_opts, nullptr)) != -1) {
        switch (c) {
        case 'h': show_help = true; break;
        case 'i': input = optarg; break;
        case 'q': quiet = true; break;
        default: show_help = true;
        }
    }
    if (!input && optind < argc) input = argv[optind];
    if (show_help) { display_help(); return 0; }

    int fd = 0;
    if (input) {
        fd = open(input, O_RDONLY);
        if (fd < 0) { fprintf(stderr, "Error: cannot open file '%s'\n", input); return 1; }
        struct stat st;
        if (fstat(fd, &st) < 0) { close(fd); fprintf(stderr, "Error: cannot stat file '%s'\n", input); return 1; }
        if (st.st_size == 0) { close(fd); return 0; }                   // (:388-391: nothing at all, not even the header row)
    }

    // ---- up to the first data line: a line starting with "#CHROM" must have come by then (:452-458 / :236-240; the
    // reference throws its buffered header row away, so nothing reaches stdout)
    vcfxh::Source src(fd);
    std::string head;
    {
        std::string buf(1 << 16, '\0');
        size_t scan = 0;
        bool eof = false, found_chrom = false;
        for (;;) {
            size_t nl;
            while ((nl = head.find('\n', scan)) == std::string::npos && !eof) {
                long r = src.read(&buf[0], buf.size());
                if (r <= 0) { eof = true; break; }
                head.append(buf.data(), (size_t)r);
            }
            if (scan >= head.size()) break;
            size_t end = (nl == std::string::npos) ? head.size() : nl;
            if (input && end > scan && head[end - 1] == '\r') --end;     // (stdin mode keeps a '\r')
            if (end > scan) {
                if (head[scan] != '#') {
                    if (!found_chrom) { fputs("Error: VCF header (#CHROM) not found before variant records.\n", stderr); return input ? 1 : 0; }
                    break;
                }
                if (end - scan >= 6 && head.compare(scan, 6, "#CHROM") == 0) found_chrom = true;
            }
            if (nl == std::string::npos) break;
            scan = nl + 1;
        }
    }

    vcfxh::RunOptions opt;
    opt.op = VCFX_OP_DOSAGE;
    opt.mode = input ? VCFX_MODE_FILE : VCFX_MODE_STDIN;
    opt.preface = head;
    vcfxh::write_all(1, HEADER_ROW, sizeof HEADER_ROW - 1);
    vcfxh::Totals tot;
    std::string err;
    const int rc = vcfxh::run_stream(src, opt, tot, err);
    if (input) close(fd);
    if (rc != VCFX_OK) { fprintf(stderr, "Error: %s\n", err.c_str()); vcfxh::finish(1); }
    if (!(quiet && input))                                               // (the stdin path prints its warnings whatever -q says, :264)
        for (uint64_t i = 0; i < tot.short_lines; ++i) fputs("Warning: Skipping VCF line with fewer than 10 fields.\n", stderr);
    vcfxh::finish(0);
}
