// VCFX_allele_counter — drop-in replacement for the reference tool of the same name
// (src/VCFX_allele_counter/VCFX_allele_counter.cpp): same flags, messages, exit codes and output
// bytes.  Host side: argument grammar (:352-395), header parse and sample selection (:806-868,
// :1139-1178, :1285-1349), the fixed header rows, optional gzip of the output (:431-514).  The
// per-line work (processChunk :550-642, countAllelesStream :1188-1246, countAllelesUnified
// :1371-1465) runs on the GPU through libvcfx_cuda.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <map>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>
#include <zlib.h>

#include "vcfx_host.h"

static void print_help() {
    fputs("VCFX_allele_counter - Count reference and alternate alleles per sample\n\n"
          "Usage: VCFX_allele_counter [OPTIONS] [FILE]\n\n"
          "Options:\n"
          "  -i, --input FILE      Input VCF file (uses mmap for best performance)\n"
          "  -t, --threads N       Number of threads (default: auto-detect CPU cores)\n"
          "  -s, --samples STR     Space-separated list of sample names to include\n"
          "  -l, --limit-samples N Limit to first N samples (useful for large cohorts)\n"
          "  -a, --aggregate       Output per-variant aggregates instead of per-sample\n"
          "  -z, --gzip            Compress output with gzip (~10x smaller)\n"
          "  -b, --binary          Output binary format (compact, for machine consumption)\n"
          "  -q, --quiet           Suppress informational messages\n"
          "  -h, --help            Display this help message\n"
          "  -v, --version         Display version information\n\n"
          "Examples:\n"
          "  VCFX_allele_counter -i input.vcf > counts.tsv              # Default per-sample\n"
          "  VCFX_allele_counter -a -i input.vcf > aggregate.tsv        # Per-variant aggregates\n"
          "  VCFX_allele_counter -z -i input.vcf > counts.tsv.gz        # Gzip compressed\n"
          "  VCFX_allele_counter -l 100 -i input.vcf > counts.tsv       # First 100 samples\n"
          "  VCFX_allele_counter -b -i input.vcf > counts.bin           # Binary format\n"
          "  VCFX_allele_counter -t 8 -i input.vcf > counts.tsv         # 8 threads\n\n"
          "Output formats:\n"
          "  Default:    CHROM  POS  ID  REF  ALT  Sample  Ref_Count  Alt_Count\n"
          "  Aggregate:  CHROM  POS  ID  REF  ALT  Total_Ref  Total_Alt  Sample_Count\n"
          "  Binary:     Compact binary with header (use -b flag)\n", stdout);
}

static const char TEXT_HEADER[] = "CHROM\tPOS\tID\tREF\tALT\tSample\tRef_Count\tAlt_Count\n";
static const char AGG_HEADER[] = "CHROM\tPOS\tID\tREF\tALT\tTotal_Ref\tTotal_Alt\tSample_Count\n";

enum class Fmt { TEXT, AGGREGATE, BINARY };

// names of fields 10.. of a "#CHROM" line (no entry for an empty tail after a final tab)
static void header_names(const std::string &line, std::vector<std::string> &names) {
    size_t p = 0;
    for (int i = 0; i < 9 && p < line.size(); ++i) {
        size_t t = line.find('\t', p);
        p = (t == std::string::npos) ? line.size() : t + 1;
    }
    while (p < line.size()) {
        size_t t = line.find('\t', p);
        size_t q = (t == std::string::npos) ? line.size() : t;
        names.emplace_back(line, p, q - p);
        p = (t == std::string::npos) ? line.size() : t + 1;
    }
}

int main(int argc, char *argv[]) {
    std::vector<std::string> req;
    const char *input = nullptr;
    bool quiet = false, gzip_out = false;
    int limit = 0, threads = 0;
    Fmt fmt = Fmt::TEXT;
    for (int i = 1; i < argc; ++i) {
        std::string arg = argv[i];
        if ((arg == "--samples" || arg == "-s") && i + 1 < argc) {
            std::string s = argv[++i];
            size_t start = 0, end;
            while ((end = s.find(' ', start)) != std::string::npos) {
                if (end > start) req.emplace_back(s.substr(start, end - start));
                start = end + 1;
            }
            if (start < s.size()) req.emplace_back(s.substr(start));
            for (auto &x : req) {
                size_t f = x.find_first_not_of(" \t\n\r"), l = x.find_last_not_of(" \t\n\r");
                if (f != std::string::npos) x = x.substr(f, l - f + 1);
            }
        } else if (arg == "--input" || arg == "-i") { if (i + 1 < argc) input = argv[++i]; }
        else if (arg == "--threads" || arg == "-t") { if (i + 1 < argc) threads = atoi(argv[++i]); }
        else if (arg == "--limit-samples" || arg == "-l") { if (i + 1 < argc) limit = atoi(argv[++i]); }
        else if (arg == "--gzip" || arg == "-z") gzip_out = true;
        else if (arg == "--aggregate" || arg == "-a") fmt = Fmt::AGGREGATE;
        else if (arg == "--binary" || arg == "-b") fmt = Fmt::BINARY;
        else if (arg == "--quiet" || arg == "-q") quiet = true;
        else if (arg == "--help" || arg == "-h") { print_help(); return 0; }
        else if (!arg.empty() && arg[0] != '-' && !input) input = argv[i];
    }
    for (int i = 1; i < argc; ++i)
        if (!strcmp(argv[i], "--version") || !strcmp(argv[i], "-v")) { puts("VCFX_allele_counter 2.0 (multi-threaded)"); return 0; }

    if (!quiet) {
        if (!req.empty()) {
            fputs("Info: Counting alleles for samples:", stderr);
            for (const auto &s : req) fprintf(stderr, " %s", s.c_str());
            fputc('\n', stderr);
        } else if (limit > 0) fprintf(stderr, "Info: Counting alleles for first %d samples\n", limit);
        else fputs("Info: Counting alleles for ALL samples\n", stderr);
        if (fmt == Fmt::AGGREGATE) fputs("Info: Output mode: aggregate (per-variant summaries)\n", stderr);
        else if (fmt == Fmt::BINARY) fputs("Info: Output mode: binary\n", stderr);
        if (gzip_out) fputs("Info: Output compression: gzip\n", stderr);
    }

    const bool unified = input && (fmt != Fmt::TEXT || gzip_out || limit > 0);
    int fd = 0;
    if (input) {
        if (!quiet) fprintf(stderr, "Info: Using mmap mode for file: %s\n", input);
        fd = open(input, O_RDONLY);
        struct stat st;
        if (fd < 0 || fstat(fd, &st) < 0) { fprintf(stderr, "Error: Cannot open file: %s\n", input); return 1; }
        if (st.st_size == 0) { fputs("Error: Empty file\n", stderr); return 1; }
    } else if (!quiet) fputs("Info: Using stdin streaming mode (single-threaded)\n", stderr);

    // ---- leading '#' block: sample names (and, on stdin, the selection as the reference grows it)
    vcfxh::Source src(fd);
    std::string head;                      // bytes consumed from the source so far
    std::vector<std::string> names;
    std::vector<uint32_t> cols;
    bool saw_chrom = false, saw_data_line = false, stream_lookup_failed = false;
    std::string failed_name;
    {
        std::string buf(1 << 16, '\0');
        size_t scan = 0;                   // start of the first line not yet classified
        bool in_block = true, eof = false;
        while (in_block) {
            size_t nl;
            while ((nl = head.find('\n', scan)) == std::string::npos && !eof) {
                long r = src.read(&buf[0], buf.size());
                if (r <= 0) { eof = true; break; }
                head.append(buf.data(), (size_t)r);
            }
            size_t end = (nl == std::string::npos) ? head.size() : nl;
            if (scan >= head.size()) break;                              // end of input inside the block
            const bool hash = head[scan] == '#';
            const bool blank = (end == scan);
            if (!hash && !(blank && !input)) { saw_data_line = true; break; }   // stdin skips blank lines (:1136)
            if (hash && end - scan >= 6 && head.compare(scan, 6, "#CHROM") == 0) {
                header_names(head.substr(scan, end - scan), names);
                saw_chrom = true;
                if (!input) {                                            // :1153-1176
                    if (!req.empty()) {
                        std::map<std::string, size_t> m;
                        for (size_t i = 0; i < names.size(); ++i) m[names[i]] = i;
                        for (const auto &s : req) {
                            auto it = m.find(s);
                            if (it == m.end()) { stream_lookup_failed = true; failed_name = s; break; }
                            cols.push_back((uint32_t)it->second);
                        }
                        if (stream_lookup_failed) break;
                    } else for (size_t i = 0; i < names.size(); ++i) cols.push_back((uint32_t)i);
                }
            }
            if (nl == std::string::npos) break;
            scan = nl + 1;
        }
    }
    if (stream_lookup_failed) { fprintf(stderr, "Error: Sample '%s' not found\n", failed_name.c_str()); return 1; }
    if (input) {
        if (names.empty()) { fputs("Error: No samples found in VCF\n", stderr); return 1; }
        if (!unified && !saw_data_line) { fputs("Error: No data lines found\n", stderr); return 1; }
        if (!req.empty()) {
            std::map<std::string, size_t> m;
            for (size_t i = 0; i < names.size(); ++i) m[names[i]] = i;
            for (const auto &s : req) {
                auto it = m.find(s);
                if (it == m.end()) { fprintf(stderr, "Error: Sample '%s' not found\n", s.c_str()); return 1; }
                cols.push_back((uint32_t)it->second);
            }
        } else for (size_t i = 0; i < names.size(); ++i) cols.push_back((uint32_t)i);
        if (unified && limit > 0 && cols.size() > (size_t)limit) {
            cols.resize((size_t)limit);
            if (!quiet) fprintf(stderr, "Info: Limiting to first %d samples\n", limit);
        }
        if (!unified && !quiet) {
            int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
            fprintf(stderr, "Info: Using %d threads\n", nt > 0 ? nt : 4);   // informational: the work runs on the GPU
        }
    } else if (!saw_chrom) {
        if (saw_data_line) { fputs("Error: No #CHROM header found before data\n", stderr); return 1; }
        vcfxh::write_all(1, TEXT_HEADER, sizeof TEXT_HEADER - 1);        // header row, then rc 1 (:1255-1259)
        return 1;
    }

    // ---- fixed header + body
    vcfxh::RunOptions opt;
    opt.op = VCFX_OP_ALLELE_COUNT;
    opt.mode = input ? VCFX_MODE_FILE : VCFX_MODE_STDIN;
    opt.preface = head;
    std::string header;
    if (!input || (unified && fmt == Fmt::TEXT)) { opt.flags = VCFX_F_AC_FORWARD; header = TEXT_HEADER; }
    else if (!unified) header = TEXT_HEADER;
    else if (fmt == Fmt::AGGREGATE) { opt.flags = VCFX_F_AC_AGGREGATE; header = AGG_HEADER; }
    else {
        opt.flags = VCFX_F_AC_BINARY;
        unsigned char h[20] = {'V', 'C', 'A', 'C', 1, 0, 0, 0};
        uint32_t ns = (uint32_t)cols.size(); memcpy(h + 8, &ns, 4); memset(h + 12, 0, 8);
        header.assign(reinterpret_cast<char *>(h), 20);
    }
    if (cols.empty()) {                    // "#CHROM" without sample columns: header row only
        vcfxh::write_all(1, header.data(), header.size());
        return 0;
    }
    opt.sel_col = cols;
    for (uint32_t c : cols) opt.sel_names.push_back(names[c]);
    // -z: the text is gzipped piece by piece as the chunks come back (gzdopen(dup(1), "wb6") like the reference's
    // GzipWriter, allele_counter.cpp:431-514), never held in memory as a whole
    gzFile gz = nullptr;
    bool gz_failed = false;
    if (gzip_out) {
        int dupfd = dup(1);
        gz = dupfd >= 0 ? gzdopen(dupfd, "wb6") : nullptr;
        if (!gz) { if (dupfd >= 0) close(dupfd); fprintf(stderr, "Error: cannot open gzip stream on stdout\n"); vcfxh::finish(1); }
        auto put = [&gz, &gz_failed](const char *p, size_t n) {
            while (n) {
                const unsigned k = (unsigned)std::min<size_t>(n, 1u << 30);
                const int w = gzwrite(gz, p, k);
                if (w <= 0) { gz_failed = true; return false; }
                p += w; n -= (size_t)w;
            }
            return true;
        };
        if (!put(header.data(), header.size())) { fprintf(stderr, "Error: gzip write failed\n"); vcfxh::finish(1); }
        opt.sink = put;
    } else vcfxh::write_all(1, header.data(), header.size());
    vcfxh::Totals tot;
    std::string err;
    int rc = vcfxh::run_stream(src, opt, tot, err);
    if (input) close(fd);
    if (gz) { if (gzclose(gz) != Z_OK) gz_failed = true; }
    if (rc != VCFX_OK || gz_failed) { fprintf(stderr, "Error: %s\n", gz_failed ? "gzip write failed" : err.c_str()); vcfxh::finish(1); }
    vcfxh::finish(0);
}
