// VCFX_inbreeding_calculator — drop-in replacement for the reference tool of the same name
// (src/VCFX_inbreeding_calculator/VCFX_inbreeding_calculator.cpp): same flags, messages, exit codes and output bytes.
// The per-line work (ALT check, a genotype code per sample column, :531-600 / :718-757) and the per-sample sums over
// the sites in file order (:603-636 / :760-790) run on the GPU via libvcfx_cuda (VCFX_OP_INBREEDING); the rows
// "name \t F \n" come back with the last chunk.  The host reads the header for the sample names and prints the
// fixed lines.  SURVEY.md §8 f3: the sample-axis reduction on the same scan as the five tools of the hot path.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <getopt.h>
#include <string>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>

#include "vcfx_host.h"

static void display_help() {
    fputs("VCFX_inbreeding_calculator: Compute individual inbreeding coefficients (F)\n"
          "based on biallelic sites in a VCF.\n\n"
          "Usage:\n"
          "  VCFX_inbreeding_calculator [options] [input.vcf]\n"
          "  VCFX_inbreeding_calculator [options] < input.vcf\n\n"
          "Options:\n"
          "  -i, --input FILE          Input VCF file (uses memory-mapping for best performance)\n"
          "  -q, --quiet               Suppress informational messages\n"
          "  -h, --help                Show this help.\n"
          "  --freq-mode <mode>        'excludeSample' (default) or 'global'\n"
          "  --skip-boundary           Skip boundary freq sites. By default, they are used.\n"
          "  --count-boundary-as-used  If also skipping boundary, still increment usedCount.\n\n"
          "Description:\n"
          "  Reads a VCF in a single pass, ignoring multi-allelic lines (ALT with commas).\n"
          "  For each biallelic variant, we parse each sample's genotype code:\n"
          "       0/0 => 0,   0/1 => 1,   1/1 => 2, else => -1 (ignored)\n\n"
          "  Then, depending on --freq-mode:\n"
          "    * excludeSample => Each sample excludes its own genotype when computing p.\n"
          "    * global        => Compute a single global p from all samples' genotypes.\n\n"
          "  The --skip-boundary option, if set, ignores boundary freq p=0 or p=1.\n"
          "    BUT if you also specify --count-boundary-as-used, those boundary sites\n"
          "    increment usedCount (forcing F=1) without contributing to sumExp.\n\n"
          "  If sumExp=0 for a sample but usedCount>0, we output F=1.\n"
          "  If usedCount=0, we output NA.\n\n"
          "Performance:\n"
          "  Uses memory-mapped I/O and SIMD for ~20x speedup over stdin mode.\n"
          "  When a file is provided directly, uses mmap for faster processing.\n\n"
          "Example:\n"
          "  VCFX_inbreeding_calculator -i input.vcf > inbreeding.txt\n"
          "  VCFX_inbreeding_calculator < input.vcf > inbreeding.txt\n", stdout);
}

static const char HEADER_ROW[] = "Sample\tInbreedingCoefficient\n";

int main(int argc, char *argv[]) {
    // vcfx::handle_common_flags (include/vcfx_core.h:31-67): --help / -h anywhere first, then --version / -v
    for (int i = 1; i < argc; ++i) if (!strcmp(argv[i], "--help") || !strcmp(argv[i], "-h")) { display_help(); return 0; }
    for (int i = 1; i < argc; ++i) if (!strcmp(argv[i], "--version") || !strcmp(argv[i], "-v")) { puts("VCFX_inbreeding_calculator version 1.1.4"); return 0; }
    const char *input = nullptr;
    bool show_help = false, quiet = false;
    unsigned flags = 0;
    static struct option long_opts[] = {{"help", no_argument, nullptr, 'h'}, {"input", required_argument, nullptr, 'i'}, {"quiet", no_argument, nullptr, 'q'},
                                        {"freq-mode", required_argument, nullptr, 0}, {"skip-boundary", no_argument, nullptr, 0},
                                        {"count-boundary-as-used", no_argument, nullptr, 0}, {nullptr, 0, nullptr, 0}};
    for (;;) {
        int idx = 0;
        const int c = getopt_long(argc, argv, "hi:q", long_opts, &idx);
        if (c == -1) break;
        switch (c) {
        case 'h': show_help = true; break;
        case 'i': input = optarg; break;
        case 'q': quiet = true; break;
        case 0:
            if (idx == 3) {
                if (!strcmp(optarg, "global")) flags |= VCFX_F_IB_GLOBAL;
                else if (!strcmp(optarg, "excludeSample")) flags &= ~VCFX_F_IB_GLOBAL;
                else fprintf(stderr, "Warning: unrecognized freq-mode='%s'. Using 'excludeSample' by default.\n", optarg);
            } else if (idx == 4) flags |= VCFX_F_IB_SKIP_BOUNDARY;
            else if (idx == 5) flags |= VCFX_F_IB_COUNT_BOUNDARY;
            break;
        default: show_help = true;
        }
    }
    if (!input && optind < argc) input = argv[optind];
    if (show_help) { display_help(); return 0; }

    int fd = 0;
    if (input) {
        fd = open(input, O_RDONLY);
        struct stat st;
        if (fd < 0 || fstat(fd, &st) < 0) { fprintf(stderr, "Error: Cannot open file: %s\n", input); return 1; }      // (:868-871)
        if (!quiet) fprintf(stderr, "Processing %s (%lld bytes)...\n", input, (long long)st.st_size);
        if (st.st_size == 0) { fputs("Error: Empty file.\n", stderr); vcfxh::write_all(1, HEADER_ROW, sizeof HEADER_ROW - 1); return 0; }
    }

    // ---- the header: sample names, and where the data lines start
    // file mode (:474-519): the leading block of '#' and empty lines, every line in it that starts with "#CHROM" ADDS its
    // columns 10..; stdin mode (:689-706): the first '#' line that contains "#CHROM", whatever stands in front of it is skipped
    vcfxh::Source src(fd);
    std::string head;                      // bytes read and not yet given up
    std::vector<std::string> names;
    bool found_chrom = false;
    {
        std::string buf(1 << 16, '\0');
        size_t scan = 0;
        bool eof = false;
        for (;;) {
            size_t nl;
            while ((nl = head.find('\n', scan)) == std::string::npos && !eof) {
                long r = src.read(&buf[0], buf.size());
                if (r <= 0) { eof = true; break; }
                head.append(buf.data(), (size_t)r);
            }
            if (scan >= head.size()) { head.clear(); break; }             // the input ended inside the header
            size_t end = (nl == std::string::npos) ? head.size() : nl;
            size_t ce = end;
            if (ce > scan && head[ce - 1] == '\r') --ce;
            const size_t next = (nl == std::string::npos) ? head.size() : nl + 1;
            if (input) {
                if (ce > scan) {
                    if (head[scan] != '#') { head.erase(0, scan); break; }                    // the first data line: the body starts here
                    if (ce - scan >= 6 && head.compare(scan, 6, "#CHROM") == 0) {
                        found_chrom = true;
                        size_t p = scan; int idx = 0;
                        while (p < ce) {
                            size_t t = head.find('\t', p);
                            if (t == std::string::npos || t > ce) t = ce;
                            if (idx >= 9) names.emplace_back(head, p, t - p);
                            ++idx; p = t + 1;
                        }
                    }
                }
                scan = next;
            } else {
                if (ce > scan && head[scan] == '#' && head.substr(scan, ce - scan).find("#CHROM") != std::string::npos) {
                    found_chrom = true;
                    size_t p = scan; int idx = 0;
                    for (;;) {                                                                // split_tabs: a final empty field counts
                        size_t t = head.find('\t', p);
                        const bool last = (t == std::string::npos || t >= ce);
                        if (idx >= 9) names.emplace_back(head, p, (last ? ce : t) - p);
                        ++idx;
                        if (last) break;
                        p = t + 1;
                    }
                    head.erase(0, next);
                    break;
                }
                // nothing in front of the header line matters: give the bytes up as they go by (in bulk: a stream may have
                // millions of lines in front of a late header)
                scan = next;
                if (scan >= (1u << 20)) { head.erase(0, scan); scan = 0; }
            }
        }
    }
    if (input) {
        if (!found_chrom || names.empty()) {
            fputs("Error: No #CHROM line or no samples found.\n", stderr);
            vcfxh::write_all(1, HEADER_ROW, sizeof HEADER_ROW - 1);
            return 0;
        }
    } else {
        if (!found_chrom) { vcfxh::write_all(1, HEADER_ROW, sizeof HEADER_ROW - 1); fputs("Error: No #CHROM line found.\n", stderr); return 0; }
        if (names.empty()) { vcfxh::write_all(1, HEADER_ROW, sizeof HEADER_ROW - 1); fputs("Error: No sample columns found.\n", stderr); return 0; }
    }

    vcfxh::RunOptions opt;
    opt.op = VCFX_OP_INBREEDING;
    opt.mode = input ? VCFX_MODE_FILE : VCFX_MODE_STDIN;
    opt.flags = flags;
    opt.preface = head;
    opt.always_submit_final = true;        // the rows come with the final chunk, even when it is empty
    opt.single_device = true;              // the sums are carried from chunk to chunk, in order
    for (size_t i = 0; i < names.size(); ++i) { opt.sel_col.push_back((uint32_t)i); opt.sel_names.push_back(names[i]); }
    std::string body;
    opt.capture = &body;                   // one short row per sample
    vcfxh::Totals tot;
    std::string err;
    const int rc = vcfxh::run_stream(src, opt, tot, err);
    if (input) close(fd);
    if (rc != VCFX_OK) { fprintf(stderr, "Error: %s\n", err.c_str()); vcfxh::finish(1); }
    vcfxh::write_all(1, HEADER_ROW, sizeof HEADER_ROW - 1);
    if (tot.rows == 0 && !quiet) fputs("No biallelic variants found.\n", stderr);
    vcfxh::write_all(1, body.data(), body.size());
    vcfxh::finish(0);
}
