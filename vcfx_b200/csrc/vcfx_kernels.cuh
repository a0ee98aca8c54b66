// vcfx_kernels.cuh — sm_100a kernels of the VCFX hot path (one fused pass over the bytes).
//
// Decomposition: the chunk is cut into byte tiles of `tile_bytes`; one WARP
// owns every line that STARTS inside its tile — the reference's own thread-chunking rule
// (allele_counter.cpp:890-903, missing_detector.cpp:404-422) applied at warp granularity, so
// a line is always parsed from its first byte by one owner and no cross-tile stitching of
// parser state exists.  A warp streams its lines in 512-byte windows (16 B per lane, fully
// coalesced 128-bit loads, next window in flight + L2 prefetch further ahead) and does in
// that single pass what the reference does in its per-line loops:
//   stage 1  '\n' / '\t' / '\r' detection by SWAR byte compares + __ballot_sync/__popc
//   stage 2  tabs 1..9 of the record located with a warp prefix sum; FORMAT -> GT index
//   stage 3  every sample column parsed by the lane that holds its leading tab, straight
//            from registers (own 16 B + 4 B look-ahead from the neighbour lane); a lane
//            whose tabs form the period-4 "d|d\t" lattice takes a branch-free SWAR path
//   stage 4  warp REDUX of the per-lane tallies; rows are queued per tile, the tile's output
//            size goes through a decoupled look-back (single-pass chained scan), and the
//            warp then formats its rows at their final, file-ordered offsets.
// Tensor cores are not involved: this is byte scanning and integer reduction, HBM-bound.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "vcfx_numfmt.cuh"

namespace vcfx {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int WARPS_PER_CTA = 8;
constexpr int WINDOW = 512;

enum : int { OP_VC = 0, OP_AF = 1, OP_HWE = 2, OP_MD = 3, OP_AC = 4 };
enum : int { MODE_FILE = 0, MODE_STDIN = 1 };

struct Rec {                 // one queued output row (32 B)
    uint32_t ls_rel;         // line start - tile start
    uint32_t prefix_len;     // bytes of "CHROM\tPOS\tID\tREF\tALT\t" taken verbatim from the line
    uint32_t a, b, c, d;     // op-specific integers
    uint32_t e, f;
};

struct DevStats {
    unsigned long long lines, data_lines, rows, flagged, pre_header, short_lines;
    unsigned long long first_short_key;   // (tile << 32 | line index in tile), min
    unsigned long long n_events;
    unsigned long long dots_terminated;
    unsigned long long last_unterminated_flagged;
    unsigned long long bytes_out;
    unsigned long long overflow;
    unsigned long long first_short_line;  // filled by resolve_events_kernel
};

struct KParams {
    const uint8_t *in;       // chunk bytes; readable and '\n'-filled for >= 64 B past n
    uint64_t n;              // valid bytes
    uint64_t lo, hi;         // this launch owns lines starting in [lo, hi)
    uint32_t tile_bytes;
    uint32_t n_tiles;
    int32_t mode;
    uint32_t flags;
    uint64_t valid_from;
    int32_t is_final;
    uint8_t *out;
    uint64_t out_cap;
    unsigned long long *desc;    // [n_tiles] look-back descriptors (zeroed before launch)
    uint32_t *tile_lines;        // [n_tiles] lines started in each tile
    unsigned int *ticket;        // dynamic tile counter (zeroed before launch)
    Rec *scratch;                // [resident warps][qcap]
    uint32_t qcap;
    DevStats *stats;
    unsigned long long *events;  // short-line events (tile << 32 | index in tile)
    uint32_t ev_cap;
};

// ---------------------------------------------------------------------------------------
// byte-parallel compares on a 32-bit word; result has 0x80 in every matching byte (exact)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t zero_bytes(uint32_t x) {
    return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
}
__device__ __forceinline__ uint32_t eq_bytes(uint32_t w, uint32_t c4) { return zero_bytes(w ^ c4); }
// bytes below 0x10: '\t' '\n' '\r' live there and nothing else in a well-formed VCF
__device__ __forceinline__ uint32_t ctrl_bytes(uint32_t w) { return zero_bytes(w & 0xF0F0F0F0u); }

constexpr uint32_t C_TAB = 0x09090909u, C_NL = 0x0A0A0A0Au, C_CR = 0x0D0D0D0Du, C_DOT = 0x2E2E2E2Eu;

// keep only bytes whose absolute position p satisfies lo <= p < hi (word starts at `base`)
__device__ __forceinline__ uint32_t range_mask(uint64_t base, uint64_t lo, uint64_t hi) {
    uint32_t m = 0x80808080u;
    if (lo > base) { uint64_t s = lo - base; m = (s >= 4) ? 0u : (m << (8 * (uint32_t)s)); }
    if (hi < base + 4) { if (hi <= base) m = 0u; else m &= (0x80808080u >> (8 * (uint32_t)(base + 4 - hi))); }
    return m;
}

__device__ __forceinline__ uint4 ld16(const uint8_t *p) {
    return __ldg(reinterpret_cast<const uint4 *>(p));
}
__device__ __forceinline__ uint32_t ldb(const uint8_t *in, uint64_t p) { return (uint32_t)__ldg(in + p); }
__device__ __forceinline__ void prefetch_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__device__ __forceinline__ bool is_dig(uint32_t b) { return (b - 48u) <= 9u; }
__device__ __forceinline__ bool is_sep(uint32_t b) { return b == '/' || b == '|'; }

// ---------------------------------------------------------------------------------------
// exact scalar parsers (slow path; byte loads hit L1/L2 because the warp just streamed the line)
// ---------------------------------------------------------------------------------------

// allele_freq_calc.cpp:321-337 + 262-293 for the sample column starting at p.  The column ends
// at the next tab or at the line end; with strip_cr a '\r' directly before the '\n' is not
// content (:363-364).  The chunk is '\n'-padded past its last byte, so the scan always stops.
__device__ __noinline__ void af_sample_slow(const uint8_t *in, uint64_t p, bool strip_cr, int gt_index,
                                            uint32_t &alt, uint32_t &total) {
    uint64_t se = p;
    uint32_t c = ldb(in, se);
    while (c != '\t' && c != '\n') { ++se; c = ldb(in, se); }
    if (strip_cr && c == '\n' && se > p && ldb(in, se - 1) == '\r') --se;
    for (int i = 0; i < gt_index && p < se; ++i) {
        while (p < se && ldb(in, p) != ':') ++p;
        if (p < se) ++p;
    }
    if (p >= se) return;
    uint64_t ge = p;
    while (ge < se && ldb(in, ge) != ':') ++ge;
    while (p < ge) {
        while (p < ge && is_sep(ldb(in, p))) ++p;
        if (p >= ge) break;
        uint64_t q = p; bool numeric = true, zero = true; const uint32_t first = ldb(in, p);
        while (q < ge) {
            uint32_t d = ldb(in, q);
            if (is_sep(d)) break;
            if (numeric) { if (!is_dig(d)) numeric = false; else if (d != '0') zero = false; }
            ++q;
        }
        if (first != '.' && numeric) { ++total; if (!zero) ++alt; }
        p = q;
    }
}

// hwe_tester.cpp:339-378 for the sample column starting at p: first ':' piece only.  A '\r'
// before the '\n' never changes the class (it is neither digit nor separator), so the piece
// simply ends at ':' / tab / '\n'.
__device__ __noinline__ int hwe_sample_slow(const uint8_t *in, uint64_t p) {
    uint64_t e = p;
    for (;;) { uint32_t c = ldb(in, e); if (c == '\t' || c == ':' || c == '\n') break; ++e; }
    while (p < e) { uint32_t c = ldb(in, p); if (c == ' ' || c == '\r') ++p; else break; }
    if (p >= e) return -1;
    if (!is_dig(ldb(in, p))) return -1;
    int a1 = 0, a2 = 0;
    while (p < e && is_dig(ldb(in, p))) { a1 = (a1 > 1) ? 2 : a1 * 10 + (int)(ldb(in, p) - 48u); ++p; }
    if (p >= e || !is_sep(ldb(in, p))) return -1;
    ++p;
    if (p >= e || !is_dig(ldb(in, p))) return -1;
    while (p < e && is_dig(ldb(in, p))) { a2 = (a2 > 1) ? 2 : a2 * 10 + (int)(ldb(in, p) - 48u); ++p; }
    if (a1 > 1 || a2 > 1) return -1;       // saturating at 2 keeps ">1" without int overflow
    if (a1 == 0 && a2 == 0) return 0;
    if (a1 == 1 && a2 == 1) return 2;
    return 1;
}

// allele_freq_calc.cpp:298-316 on the FORMAT field [p, e)
__device__ __forceinline__ int gt_index_of(const uint8_t *in, uint64_t p, uint64_t e) {
    int idx = 0;
    while (p < e) {
        uint64_t q = p;
        while (q < e && ldb(in, q) != ':') ++q;
        if (q - p == 2 && ldb(in, p) == 'G' && ldb(in, p + 1) == 'T') return idx;
        ++idx;
        p = (q < e) ? q + 1 : q;
    }
    return -1;
}

// byte index (0..15) of the r-th (0-based) set 0x80-bit over the lane's four mask words
__device__ __forceinline__ int nth_byte(uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3, int r) {
    uint32_t m[4] = {m0, m1, m2, m3};
    int pos = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int c = __popc(m[j]);
        if (r >= 0 && r < c) {
            uint32_t x = m[j];
            for (int i = 0; i < r; ++i) x &= x - 1;
            pos = 4 * j + ((__ffs(x) - 1) >> 3);
            r = -1;
        } else if (r >= 0) r -= c;
    }
    return pos;
}
// clear the first d set bits over the four words
__device__ __forceinline__ void drop_first(uint32_t &m0, uint32_t &m1, uint32_t &m2, uint32_t &m3, int d) {
    while (d > 0 && m0) { m0 &= m0 - 1; --d; }
    while (d > 0 && m1) { m1 &= m1 - 1; --d; }
    while (d > 0 && m2) { m2 &= m2 - 1; --d; }
    while (d > 0 && m3) { m3 &= m3 - 1; --d; }
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, v, o); if (lane >= o) v += t; }
    return v;
}

// ---------------------------------------------------------------------------------------
// per-lane sample parsing from registers
// ---------------------------------------------------------------------------------------
struct Tally { uint32_t a, b, c; };   // AF: alt,total   HWE: homRef,het,homAlt

// AF, one sample whose first four bytes are q (b0 = first byte after the leading tab)
__device__ __forceinline__ void af_sample_reg(uint32_t q, const uint8_t *in, uint64_t pos, bool strip_cr,
                                              int gt_index, Tally &t) {
    uint32_t b0 = q & 0xFF, b1 = (q >> 8) & 0xFF, b2 = (q >> 16) & 0xFF, b3 = q >> 24;
    if (gt_index == 0) {
        bool term0 = (b0 == '\t' || b0 == ':' || b0 == '\n');
        if (term0) return;                                   // empty sample / empty GT
        bool d0 = is_dig(b0), dot0 = (b0 == '.');
        if (d0 || dot0) {
            bool term1 = (b1 == '\t' || b1 == ':' || b1 == '\n');
            if (term1) { if (d0) { t.b++; t.a += (b0 != '0'); } return; }
            if (is_sep(b1)) {
                bool d2 = is_dig(b2), dot2 = (b2 == '.');
                bool term3 = (b3 == '\t' || b3 == ':' || b3 == '\n');
                if ((d2 || dot2) && term3) {
                    if (d0) { t.b++; t.a += (b0 != '0'); }
                    if (d2) { t.b++; t.a += (b2 != '0'); }
                    return;
                }
            }
        }
    }
    af_sample_slow(in, pos, strip_cr, gt_index, t.a, t.b);
}

__device__ __forceinline__ void hwe_sample_reg(uint32_t q, const uint8_t *in, uint64_t pos, Tally &t) {
    uint32_t b0 = q & 0xFF, b1 = (q >> 8) & 0xFF, b2 = (q >> 16) & 0xFF, b3 = q >> 24;
    int cls;
    if (!is_dig(b0)) {
        if (b0 == ' ' || b0 == '\r') cls = hwe_sample_slow(in, pos); else return;
    } else if (is_sep(b1)) {
        if (!is_dig(b2)) return;
        if (is_dig(b3)) cls = hwe_sample_slow(in, pos);
        else { if (b0 > '1' || b2 > '1') return; cls = (int)(b0 - '0') + (int)(b2 - '0'); }
    } else if (is_dig(b1)) cls = hwe_sample_slow(in, pos);
    else return;
    if (cls == 0) t.a++; else if (cls == 1) t.b++; else if (cls == 2) t.c++;
}

// The lane's sample tabs (sm*) with its 16 bytes (w*) and 4 look-ahead bytes (la).
template <int OP>
__device__ __forceinline__ void lane_samples(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t la,
                                             uint32_t sm0, uint32_t sm1, uint32_t sm2, uint32_t sm3,
                                             const uint8_t *in, uint64_t pb, bool strip_cr, int gt_index, Tally &t) {
    if ((sm0 | sm1 | sm2 | sm3) == 0) return;
    // ---- period-4 lattice: tabs at the same byte of all four words, every sample "d s d"
    if (sm0 == sm1 && sm1 == sm2 && sm2 == sm3 && (sm0 & (sm0 - 1)) == 0 && gt_index == 0) {
        uint32_t sh = (uint32_t)(__ffs(sm0));               // 8,16,24,32 = 8*(tab byte + 1)
        uint32_t x0 = __funnelshift_rc(w0, w1, sh), x1 = __funnelshift_rc(w1, w2, sh);
        uint32_t x2 = __funnelshift_rc(w2, w3, sh), x3 = __funnelshift_rc(w3, la, sh);
        // x = [d0, sep, d1, tab]: check tab + separator, then both digits
        uint32_t bad = 0;
        uint32_t xs[4] = {x0, x1, x2, x3};
        uint32_t nz = 0, het = 0, ha = 0, hwe_ok = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t x = xs[j];
            uint32_t f = x & 0xFF00FF00u;
            bad |= (uint32_t)((f != 0x09007C00u) & (f != 0x09002F00u));
            uint32_t v = x & 0x00FF00FFu;                    // the two allele bytes, one per 16-bit lane
            // digit <=> v >= '0' and not v >= ':'  (bit 8 of v+0xD0 set, bit 8 of v+0xC6 clear; no carry leaves a lane)
            bad |= ((v + 0x00D000D0u) & ~(v + 0x00C600C6u) & 0x01000100u) ^ 0x01000100u;
            uint32_t u = v - 0x00300030u;                    // 0..9 per lane when both are digits
            if (OP == OP_AF) {
                nz += __popc((u + 0x000F000Fu) & 0x00100010u);
            } else {
                uint32_t a = u & 0xFFFFu, b = u >> 16;
                uint32_t ok = (uint32_t)((a | b) <= 1u);
                hwe_ok += ok; het += ok & (a ^ b); ha += ok & (a & b);
            }
        }
        if (bad == 0) {
            if (OP == OP_AF) { t.a += nz; t.b += 8; }
            else { t.a += hwe_ok - het - ha; t.b += het; t.c += ha; }
            return;
        }
    }
    // ---- generic: one sample per owned tab
    uint32_t ws[5] = {w0, w1, w2, w3, la};
    uint32_t sm[4] = {sm0, sm1, sm2, sm3};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t m = sm[j];
        while (m) {
            int k = (__ffs(m) - 1) >> 3;
            m &= m - 1;
            uint32_t q = __funnelshift_rc(ws[j], ws[j + 1], 8u * (uint32_t)(k + 1));
            uint64_t pos = pb + 4 * j + k + 1;
            if (OP == OP_AF) af_sample_reg(q, in, pos, strip_cr, gt_index, t);
            else hwe_sample_reg(q, in, pos, t);
        }
    }
}

// ---------------------------------------------------------------------------------------
// decoupled look-back over per-tile output sizes (status in bits 62..63)
// ---------------------------------------------------------------------------------------
constexpr unsigned long long ST_AGG = 1ULL << 62, ST_PFX = 2ULL << 62, ST_MASK = 3ULL << 62;

__device__ __forceinline__ unsigned long long ld_desc(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// returns the exclusive prefix of tile `t` and publishes its inclusive prefix
__device__ __forceinline__ unsigned long long lookback(unsigned long long *desc, uint32_t t,
                                                       unsigned long long own, int lane) {
    if (t == 0) { if (lane == 0) st_desc(desc, ST_PFX | own); return 0; }
    if (lane == 0) st_desc(desc + t, ST_AGG | own);
    unsigned long long excl = 0;
    long long idx = (long long)t - 1;
    for (;;) {
        long long my = idx - lane;
        unsigned long long v;
        do {
            v = (my >= 0) ? ld_desc(desc + my) : ST_PFX;
        } while (__any_sync(FULL, (v & ST_MASK) == 0));
        unsigned pf = __ballot_sync(FULL, (v & ST_MASK) == ST_PFX);
        int stop = pf ? (__ffs(pf) - 1) : 32;          // nearest predecessor with a full prefix
        unsigned long long val = (lane <= stop && lane < 32) ? (v & ~ST_MASK) : 0ULL;
        if (lane > stop) val = 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(FULL, val, o);
        excl += val;
        if (pf) break;
        idx -= 32;
    }
    if (lane == 0) st_desc(desc + t, ST_PFX | (excl + own));
    return excl;
}

// ---------------------------------------------------------------------------------------
// the fused scan / parse / reduce / format kernel
// ---------------------------------------------------------------------------------------
template <int OP>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
vcfx_scan_kernel(const KParams P) {
    const int lane = threadIdx.x & 31;
    const uint32_t gwarp = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    const uint8_t *__restrict__ in = P.in;
    const uint64_t n = P.n;
    const bool strip_cr = (OP == OP_HWE) || (P.mode == MODE_FILE && (OP == OP_AF || OP == OP_VC || OP == OP_MD));
    Rec *queue = P.scratch ? P.scratch + (size_t)gwarp * P.qcap : nullptr;

    unsigned long long s_lines = 0, s_data = 0, s_rows = 0, s_pre = 0, s_short = 0;

    for (;;) {
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(P.ticket, 1u);
        tile = __shfl_sync(FULL, tile, 0);
        if (tile >= P.n_tiles) break;

        const uint64_t a = P.lo + (uint64_t)tile * P.tile_bytes;
        const uint64_t b = min(a + (uint64_t)P.tile_bytes, P.hi);

        // ---- first line start in [a, b): byte 0 of the chunk, or one past a '\n' at >= a-1
        uint64_t ls;
        if (a == 0) ls = 0;
        else {
            ls = b;                                       // "none"
            uint64_t from = a - 1, wb = from & ~(uint64_t)15;
            while (wb < b - 1 + 0) {
                uint64_t pb = wb + 16 * lane;
                uint4 v = ld16(in + pb);
                uint32_t m0 = eq_bytes(v.x, C_NL) & range_mask(pb, from, b - 1);
                uint32_t m1 = eq_bytes(v.y, C_NL) & range_mask(pb + 4, from, b - 1);
                uint32_t m2 = eq_bytes(v.z, C_NL) & range_mask(pb + 8, from, b - 1);
                uint32_t m3 = eq_bytes(v.w, C_NL) & range_mask(pb + 12, from, b - 1);
                unsigned bal = __ballot_sync(FULL, (m0 | m1 | m2 | m3) != 0);
                if (bal) {
                    int src = __ffs(bal) - 1;
                    int k = nth_byte(m0, m1, m2, m3, 0);
                    k = __shfl_sync(FULL, k, src);
                    ls = wb + 16 * src + k + 1;
                    break;
                }
                wb += WINDOW;
            }
        }

        uint32_t nrows = 0, nlines = 0;
        unsigned long long out_bytes = 0;

        // ---- every line that starts in the tile
        while (ls < b && ls < n) {
            uint64_t wb = ls & ~(uint64_t)15;
            uint4 cur = ld16(in + wb + 16 * lane);
            uint4 nxt = ld16(in + wb + WINDOW + 16 * lane);
            int tabs = 0;                    // tabs seen so far in this line (uniform)
            uint64_t tp[9];                  // positions of tabs 0..8 (uniform)
#pragma unroll
            for (int k = 0; k < 9; ++k) tp[k] = 0;
            bool hdr_done = false, do_samples = false;
            int gt_index = -1;
            Tally tl = {0, 0, 0};
            uint64_t ee = 0, e = 0;          // content end (CR stripped) and '\n' position
            const uint32_t first = ldb(in, ls);
            // lines that need no field work: '#' lines (all ops) — still walked to find '\n'
            const bool hash = (first == '#');
            uint32_t wcount = 0;

            for (;;) {
                const uint64_t pb = wb + 16 * lane;
                if ((wcount & 7) == 0) { uint64_t pf = wb + 8 * WINDOW + 128 * lane; if (pf < n) prefetch_l2(in + pf); }
                ++wcount;
                // look-ahead: the 4 bytes after my 16
                uint32_t la = __shfl_down_sync(FULL, cur.x, 1);
                uint32_t nx0 = __shfl_sync(FULL, nxt.x, 0);
                if (lane == 31) la = nx0;

                // ---- stage 1: control bytes
                uint32_t c0 = ctrl_bytes(cur.x), c1 = ctrl_bytes(cur.y), c2 = ctrl_bytes(cur.z), c3 = ctrl_bytes(cur.w);
                uint32_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, eol0 = 0, eol1 = 0, eol2 = 0, eol3 = 0;
                bool eol_is_cr = false;
                if (c0 | c1 | c2 | c3) {
                    t0 = eq_bytes(cur.x, C_TAB); t1 = eq_bytes(cur.y, C_TAB);
                    t2 = eq_bytes(cur.z, C_TAB); t3 = eq_bytes(cur.w, C_TAB);
                    if ((c0 ^ t0) | (c1 ^ t1) | (c2 ^ t2) | (c3 ^ t3)) {
                        uint32_t n0 = eq_bytes(cur.x, C_NL), n1 = eq_bytes(cur.y, C_NL);
                        uint32_t n2 = eq_bytes(cur.z, C_NL), n3 = eq_bytes(cur.w, C_NL);
                        eol0 = n0; eol1 = n1; eol2 = n2; eol3 = n3;
                        if (strip_cr) {
                            // a '\r' directly before '\n' ends the content one byte early
                            uint32_t r0 = eq_bytes(cur.x, C_CR), r1 = eq_bytes(cur.y, C_CR);
                            uint32_t r2 = eq_bytes(cur.z, C_CR), r3 = eq_bytes(cur.w, C_CR);
                            uint32_t nla = eq_bytes(la, C_NL);
                            r0 &= __funnelshift_r(n0, n1, 8); r1 &= __funnelshift_r(n1, n2, 8);
                            r2 &= __funnelshift_r(n2, n3, 8); r3 &= __funnelshift_r(n3, nla, 8);
                            eol0 |= r0; eol1 |= r1; eol2 |= r2; eol3 |= r3;
                        }
                    }
                }
                // nothing before the line start counts (first window only)
                if (wb < ls) {
                    uint32_t k0 = range_mask(pb, ls, ~0ULL), k1 = range_mask(pb + 4, ls, ~0ULL);
                    uint32_t k2 = range_mask(pb + 8, ls, ~0ULL), k3 = range_mask(pb + 12, ls, ~0ULL);
                    t0 &= k0; t1 &= k1; t2 &= k2; t3 &= k3; eol0 &= k0; eol1 &= k1; eol2 &= k2; eol3 &= k3;
                }
                unsigned ebal = __ballot_sync(FULL, (eol0 | eol1 | eol2 | eol3) != 0);
                bool found = ebal != 0;
                uint64_t hi_clip = wb + WINDOW;
                if (found) {
                    int src = __ffs(ebal) - 1;
                    int k = nth_byte(eol0, eol1, eol2, eol3, 0);
                    k = __shfl_sync(FULL, k, src);
                    ee = wb + 16 * src + k;
                    eol_is_cr = (ldb(in, ee) == '\r');
                    e = eol_is_cr ? ee + 1 : ee;
                    hi_clip = ee;
                    uint32_t k0 = range_mask(pb, 0, ee), k1 = range_mask(pb + 4, 0, ee);
                    uint32_t k2 = range_mask(pb + 8, 0, ee), k3 = range_mask(pb + 12, 0, ee);
                    t0 &= k0; t1 &= k1; t2 &= k2; t3 &= k3;
                }

                if (!hash) {
                    uint32_t s0 = t0, s1 = t1, s2 = t2, s3 = t3;        // sample tabs of this lane
                    if (!hdr_done) {
                        // ---- stage 2: rank the tabs of this window, pick out tabs 0..8
                        int cnt = __popc(t0) + __popc(t1) + __popc(t2) + __popc(t3);
                        int incl = 0, excl = 0, total;
                        if (OP == OP_VC) total = (int)__reduce_add_sync(FULL, (unsigned)cnt);
                        else { incl = warp_incl_scan(cnt, lane); excl = incl - cnt; total = __shfl_sync(FULL, incl, 31); }
                        if (OP != OP_VC && total > 0) {
#pragma unroll
                            for (int k = 0; k < 9; ++k) {
                                int r = k - tabs;
                                if (r >= 0 && r < total) {
                                    bool mine = (r >= excl) && (r < incl);
                                    unsigned bal = __ballot_sync(FULL, mine);
                                    int owner = __ffs(bal) - 1;
                                    int byte = mine ? nth_byte(t0, t1, t2, t3, r - excl) : 0;
                                    byte = __shfl_sync(FULL, byte, owner);
                                    tp[k] = wb + 16 * owner + byte;
                                }
                            }
                            int drop = 8 - (tabs + excl);
                            if (drop > 0) drop_first(s0, s1, s2, s3, drop);
                        }
                        tabs += total;
                        if (tabs >= 9 && OP != OP_VC) {
                            hdr_done = true;
                            // FORMAT = [tp[7]+1, tp[8]); decide whether samples are parsed
                            if (OP == OP_AF) {
                                gt_index = gt_index_of(in, tp[7] + 1, tp[8]);
                                do_samples = gt_index >= 0 && (ls >= P.valid_from);
                            } else if (OP == OP_HWE) {
                                bool fmt_ok = (tp[8] - tp[7] - 1 >= 2) && ldb(in, tp[7] + 1) == 'G' && ldb(in, tp[7] + 2) == 'T';
                                do_samples = fmt_ok;
                                gt_index = 0;
                            }
                        }
                    }
                    // ---- stage 3: samples owned by this lane
                    if (do_samples && (OP == OP_AF || OP == OP_HWE)) {
                        lane_samples<OP>(cur.x, cur.y, cur.z, cur.w, la, s0, s1, s2, s3, in, pb,
                                         strip_cr, gt_index, tl);
                    }
                }
                (void)hi_clip;
                if (found) break;
                wb += WINDOW;
                cur = nxt;
                nxt = ld16(in + wb + WINDOW + 16 * lane);
            }
            // ---- end of line: ls .. ee (content) .. e ('\n')
            ++nlines;
            const bool empty = (OP == OP_VC) ? (e == ls) : (ee == ls);   // variant_counter tests the raw length (:364)
            if (OP == OP_VC) {
                if (!empty && !hash) {
                    if (tabs >= 7) ++s_rows;
                    else {
                        ++s_short;
                        if (lane == 0) {
                            unsigned long long key = ((unsigned long long)tile << 32) | (nlines - 1);
                            atomicMin(&P.stats->first_short_key, key);
                            unsigned long long slot = atomicAdd(&P.stats->n_events, 1ULL);
                            if (slot < P.ev_cap) P.events[slot] = key;
                        }
                    }
                }
            } else if ((OP == OP_AF || OP == OP_HWE) && !empty && !hash) {
                uint32_t ra = __reduce_add_sync(FULL, tl.a), rb = __reduce_add_sync(FULL, tl.b);
                uint32_t rc = (OP == OP_HWE) ? __reduce_add_sync(FULL, tl.c) : 0u;
                bool row = false;
                uint32_t prefix_len = 0;
                if (OP == OP_AF) {
                    if (ls < P.valid_from) ++s_pre;
                    else if (P.mode == MODE_FILE) {
                        ++s_data;
                        // FORMAT must exist and be non-empty (:396-401), GT must be one of its keys (:413)
                        if (tabs >= 8) {
                            uint64_t fs = tp[7] + 1, fe = (tabs >= 9) ? tp[8] : ee;
                            if (fs < fe) {
                                if (tabs < 9) gt_index = gt_index_of(in, fs, fe);
                                row = gt_index >= 0;
                            }
                        }
                    } else {
                        // stdin: fields = tabs + 1, minus a dropped empty tail (:509-518)
                        int nf = tabs + ((ldb(in, ee - 1) == '\t') ? 0 : 1);
                        if (nf < 9) ++s_short;
                        else {
                            ++s_data;
                            if (tabs < 9) gt_index = gt_index_of(in, tp[7] + 1, ee);
                            row = gt_index >= 0;
                        }
                    }
                } else {
                    ++s_data;
                    if (tabs >= 9) {
                        bool fmt_ok = do_samples;
                        bool alt_has_comma = false;
                        for (uint64_t q = tp[3] + 1; q < tp[4]; ++q) alt_has_comma |= (ldb(in, q) == ',');
                        row = fmt_ok && !alt_has_comma;
                        if (P.mode == MODE_FILE) {
                            // CHROM, POS, ALT non-empty (:497) and a non-empty remainder after tab 9 (:516-520)
                            row = row && (tp[0] > ls) && (tp[1] > tp[0] + 1) && (tp[4] > tp[3] + 1) && (tp[8] + 1 < ee);
                        }
                    }
                }
                if (row) {
                    prefix_len = (uint32_t)(tp[4] + 1 - ls);
                    if (lane == 0) {
                        if (nrows < P.qcap) {
                            Rec r; r.ls_rel = (uint32_t)(ls - a); r.prefix_len = prefix_len;
                            r.a = ra; r.b = rb; r.c = rc; r.d = 0; r.e = 0; r.f = 0;
                            queue[nrows] = r;
                        }
                    }
                    ++nrows; ++s_rows;
                    out_bytes += prefix_len + ((OP == OP_AF) ? 7 : 9);
                }
            }
            ls = e + 1;
        }
        s_lines += nlines;
        if (lane == 0) P.tile_lines[tile] = nlines;

        // ---- stage 4: place and write this tile's rows
        if (OP == OP_AF || OP == OP_HWE) {
            unsigned long long base = lookback(P.desc, tile, out_bytes, lane);
            if (tile == P.n_tiles - 1 && lane == 0) P.stats->bytes_out = base + out_bytes;
            if (nrows > P.qcap || base + out_bytes > P.out_cap) {
                if (lane == 0) atomicAdd(&P.stats->overflow, 1ULL);
            } else {
                __syncwarp();
                for (uint32_t r0 = 0; r0 < nrows; r0 += 32) {
                    uint32_t r = r0 + lane;
                    Rec rec; rec.prefix_len = 0; rec.ls_rel = 0; rec.a = rec.b = rec.c = 0;
                    uint32_t len = 0;
                    if (r < nrows) { rec = queue[r]; len = rec.prefix_len + ((OP == OP_AF) ? 7 : 9); }
                    int incl = warp_incl_scan((int)len, lane);
                    unsigned long long off = base + (unsigned long long)(incl - (int)len);
                    if (r < nrows) {
                        uint8_t *o = P.out + off;
                        const uint8_t *src = in + a + rec.ls_rel;
                        for (uint32_t i = 0; i < rec.prefix_len; ++i) o[i] = __ldg(src + i);
                        o += rec.prefix_len;
                        char num[24]; int nl;
                        if (OP == OP_AF) {
                            double v = af_value(rec.a, rec.b);
                            nl = (P.mode == MODE_FILE) ? fmt_af_file(v, num) : fmt_af_stdin(v, num);
                        } else {
                            double pv = hwe_pvalue((int)rec.a, (int)rec.b, (int)rec.c);
                            nl = (P.mode == MODE_FILE) ? fmt_p_file(pv, num) : fmt_p_stdin(pv, num);
                        }
                        for (int i = 0; i < nl; ++i) o[i] = (uint8_t)num[i];
                        o[nl] = '\n';
                    }
                    base += (unsigned long long)__shfl_sync(FULL, incl, 31);
                }
            }
        }
    }
    if (lane == 0) {
        if (s_lines) atomicAdd(&P.stats->lines, s_lines);
        if (s_data) atomicAdd(&P.stats->data_lines, s_data);
        if (s_rows) atomicAdd(&P.stats->rows, s_rows);
        if (s_pre) atomicAdd(&P.stats->pre_header, s_pre);
        if (s_short) atomicAdd(&P.stats->short_lines, s_short);
    }
}

// Turn (tile, index-in-tile) keys into 1-based line numbers: exclusive scan of the per-tile
// line counts, one CTA (n_tiles is at most a few hundred thousand).
__global__ void __launch_bounds__(1024)
resolve_events_kernel(const uint32_t *tile_lines, uint32_t n_tiles, unsigned long long *tile_base,
                      unsigned long long *events, uint32_t ev_cap, DevStats *stats) {
    __shared__ unsigned long long part[1024];
    const uint32_t tid = threadIdx.x;
    const uint32_t per = (n_tiles + 1023) / 1024;
    const uint32_t s = tid * per, e = min(s + per, n_tiles);
    unsigned long long sum = 0;
    for (uint32_t i = s; i < e; ++i) sum += tile_lines[i];
    part[tid] = sum;
    __syncthreads();
    if (tid == 0) { unsigned long long run = 0; for (int i = 0; i < 1024; ++i) { unsigned long long v = part[i]; part[i] = run; run += v; } }
    __syncthreads();
    unsigned long long run = part[tid];
    for (uint32_t i = s; i < e; ++i) { tile_base[i] = run; run += tile_lines[i]; }
    __syncthreads();
    unsigned long long nev = stats->n_events;
    if (nev > ev_cap) nev = ev_cap;
    for (unsigned long long i = tid; i < nev; i += 1024) {
        unsigned long long k = events[i];
        events[i] = tile_base[k >> 32] + (k & 0xFFFFFFFFULL) + 1ULL;
    }
    if (tid == 0) {
        unsigned long long k = stats->first_short_key;
        stats->first_short_line = (k == ~0ULL) ? 0ULL : tile_base[k >> 32] + (k & 0xFFFFFFFFULL) + 1ULL;
    }
}

}  // namespace vcfx
