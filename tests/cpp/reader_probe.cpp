// reader_probe — exercises vcfxh::Source (single read and multi-threaded pread path) and
// vcfxh::write_all (write and multi-threaded pwrite path) without touching the GPU:
//   reader_probe <in> <out> <cap_bytes>   copies <in> to <out> through buffers of cap_bytes
#include <cstdio>
#include <cstdlib>
#include <fcntl.h>
#include <unistd.h>
#include <vector>

#include "vcfx_host.h"

int main(int argc, char **argv) {
    if (argc != 4) return 2;
    const int in = open(argv[1], O_RDONLY);
    const int out = open(argv[2], O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (in < 0 || out < 0) return 3;
    const size_t cap = (size_t)strtoull(argv[3], nullptr, 10);
    std::vector<char> buf(cap);
    vcfxh::Source src(in);
    unsigned long long total = 0;
    for (;;) {
        size_t have = 0;
        bool eof = false;
        while (have < cap) {                       // the fill loop of run_stream
            long r = src.read(buf.data() + have, cap - have);
            if (r < 0) return 4;
            if (r == 0) { eof = true; break; }
            have += (size_t)r;
        }
        if (have && !vcfxh::write_all(out, buf.data(), have)) return 5;
        total += have;
        if (eof) break;
    }
    printf("%llu\n", total);
    return 0;
}
