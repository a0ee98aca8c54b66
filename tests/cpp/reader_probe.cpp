// reader_probe — exercises vcfxh::Source (single read and multi-threaded pread path) and
// vcfxh::write_all (write and multi-threaded pwrite path) without touching the GPU:
//   reader_probe <in> <out> <cap_bytes> [carry_bytes]   copies <in> to <out> through buffers of cap_bytes;
//   with carry_bytes the tail of every buffer is held back and put in front of the next one (run_stream's partial last line),
//   so the reads alternate between sizes above and below the threshold of the pread path
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <unistd.h>
#include <vector>

#include "vcfx_host.h"

int main(int argc, char **argv) {
    if (argc != 4 && argc != 5) return 2;
    const int in = open(argv[1], O_RDONLY);
    const int out = open(argv[2], O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (in < 0 || out < 0) return 3;
    const size_t cap = (size_t)strtoull(argv[3], nullptr, 10);
    const size_t carry = argc == 5 ? (size_t)strtoull(argv[4], nullptr, 10) : 0;
    std::vector<char> buf(cap);
    size_t kept = 0;
    vcfxh::Source src(in);
    unsigned long long total = 0;
    for (;;) {
        size_t have = kept;
        bool eof = false;
        while (have < cap) {                       // the fill loop of run_stream
            long r = src.read(buf.data() + have, cap - have);
            if (r < 0) return 4;
            if (r == 0) { eof = true; break; }
            have += (size_t)r;
        }
        kept = (eof || have <= carry) ? 0 : ((total / cap) % 2 ? carry : 0);     // every other buffer leaves a tail behind
        const size_t put = have - kept;
        if (put && !vcfxh::write_all(out, buf.data(), put)) return 5;
        total += put;
        if (kept) memmove(buf.data(), buf.data() + put, kept);
        if (eof) break;
    }
    printf("%llu\n", total);
    return 0;
}
