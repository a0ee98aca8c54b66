// vcfx_numfmt.cuh — the numeric text of the hot path: bit-for-bit, with ONE stated exception (exp, below).
//
// Every function is __host__ __device__: the device build is what libvcfx_cuda's format
// stage runs; the host build (csrc/vcfx_numfmt_host.cpp) exists only so the "not gpu"
// tests can check the same source against the oracle without a GPU.
//
// What has to match (reference file:line):
//   AF   FILE  : writeDouble4, trunc(v*10000.0 + 0.5)           allele_freq_calc.cpp:119-143
//   AF   STDIN : iostream fixed << setprecision(4) == "%.4f"    allele_freq_calc.cpp:553-555
//   HWE  value : Yates chi-square + A&S 7.1.26 erfc * exp       hwe_tester.cpp:278-315
//   HWE  FILE  : appendDouble, six truncated digits             hwe_tester.cpp:236-268
//   HWE  STDIN : "%.6f"                                         hwe_tester.cpp:605-606
//
// The reference's x86 build has no FMA (no -march), so each double operation rounds on its
// own.  On the device that is enforced twice: the file is compiled with -fmad=false and the
// operations go through the __d*_rn intrinsics, which are never contracted.
//
// The exception: hwe_pvalue() ends in exp(), CUDA's on the device (<= 1 ulp) and glibc's in the reference.  Tolerance
// (BASELINE.json): 1e-12 relative.  Measured on the device against the oracle's libm arithmetic over 1,012,341 count
// triples (tests/test_gpu_hwe_pvalue.py, `parity.hwe_pvalue` in bench.py's line): max relative difference 3.9e-16,
// 2.2 % of the values differ in the last bit, no FILE-mode or stdin-mode text differs.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define VCFX_HD __host__ __device__ __forceinline__
#else
#define VCFX_HD inline
#endif

namespace vcfx {

#if defined(__CUDA_ARCH__)
VCFX_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
VCFX_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
VCFX_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
VCFX_HD double ddiv(double a, double b) { return __ddiv_rn(a, b); }
VCFX_HD double dsqrt(double a) { return __dsqrt_rn(a); }
#else
VCFX_HD double dmul(double a, double b) { volatile double r = a * b; return r; }
VCFX_HD double dadd(double a, double b) { volatile double r = a + b; return r; }
VCFX_HD double dsub(double a, double b) { volatile double r = a - b; return r; }
VCFX_HD double ddiv(double a, double b) { volatile double r = a / b; return r; }
VCFX_HD double dsqrt(double a) { return sqrt(a); }
#endif

// allele frequency as the reference computes it (allele_freq_calc.cpp:448-449)
VCFX_HD double af_value(uint32_t alt, uint32_t total) {
    return total > 0 ? ddiv((double)alt, (double)total) : 0.0;
}

VCFX_HD int put_u64(char *d, unsigned long long v) {
    char t[20]; int n = 0;
    do { t[n++] = (char)('0' + (int)(v % 10ULL)); v /= 10ULL; } while (v);
    for (int i = 0; i < n; ++i) d[i] = t[n - 1 - i];
    return n;
}

// FILE-mode AF text: two roundings (multiply, add), then truncation.
VCFX_HD int fmt_af_file(double v, char *d) {
    int n = 0;
    if (v < 0) { d[n++] = '-'; v = -v; }
    unsigned long long sc = (unsigned long long)dadd(dmul(v, 10000.0), 0.5);
    n += put_u64(d + n, sc / 10000ULL);
    unsigned fr = (unsigned)(sc % 10000ULL);
    d[n++] = '.';
    d[n++] = (char)('0' + (fr / 1000) % 10);
    d[n++] = (char)('0' + (fr / 100) % 10);
    d[n++] = (char)('0' + (fr / 10) % 10);
    d[n++] = (char)('0' + fr % 10);
    return n;
}

// "%.<digits>f" of a finite double 0 <= v < 2^10 with digits <= 6: the exact binary value
// times 10^digits, rounded half-to-even as glibc's printf does, in 128-bit integers.
VCFX_HD int fmt_fixed_exact(double v, int digits, char *d) {
    int n = 0;
    unsigned long long bits;
#if defined(__CUDA_ARCH__)
    bits = (unsigned long long)__double_as_longlong(v);
#else
    { union { double f; unsigned long long u; } cv; cv.f = v; bits = cv.u; }
#endif
    if (bits >> 63) { d[n++] = '-'; bits &= 0x7FFFFFFFFFFFFFFFULL; }
    int ex = (int)(bits >> 52);
    unsigned long long man = bits & 0xFFFFFFFFFFFFFULL;
    if (ex == 0) ex = 1; else man |= 1ULL << 52;      // subnormal / normal
    int sh = 1075 - ex;                               // value = man * 2^-sh
    unsigned long long p10 = 1;
    for (int i = 0; i < digits; ++i) p10 *= 10ULL;
    unsigned __int128 M = (unsigned __int128)man * p10;   // < 2^53 * 10^6 < 2^73
    unsigned long long q;
    if (sh <= 0) {
        q = (unsigned long long)(M << (-sh));         // callers keep v < 2^10, digits <= 6
    } else if (sh >= 127) {
        q = 0;                                        // M < 2^73 is far below half a unit
    } else {
        unsigned __int128 one = (unsigned __int128)1 << sh;
        unsigned __int128 rem = M & (one - 1);
        unsigned __int128 half = one >> 1;
        q = (unsigned long long)(M >> sh);
        if (rem > half || (rem == half && (q & 1ULL))) ++q;
    }
    n += put_u64(d + n, q / p10);
    if (digits > 0) {
        d[n++] = '.';
        unsigned long long fr = q % p10;
        for (int i = digits - 1; i >= 0; --i) { d[n + i] = (char)('0' + (int)(fr % 10ULL)); fr /= 10ULL; }
        n += digits;
    }
    return n;
}

VCFX_HD int fmt_af_stdin(double v, char *d) { return fmt_fixed_exact(v, 4, d); }
VCFX_HD int fmt_p_stdin(double v, char *d) { return fmt_fixed_exact(v, 6, d); }

// FILE-mode HWE text: integer part, then six times {x10, take the digit, subtract}.
VCFX_HD int fmt_p_file(double v, char *d) {
    int n = 0;
    if (v < 0) { d[n++] = '-'; v = -v; }
    long long ip = (long long)v;
    double fr = dsub(v, (double)ip);
    n += put_u64(d + n, (unsigned long long)ip);
    d[n++] = '.';
    for (int i = 0; i < 6; ++i) {
        fr = dmul(fr, 10.0);
        int dg = (int)fr;
        d[n++] = (char)('0' + dg);
        fr = dsub(fr, (double)dg);
    }
    return n;
}

VCFX_HD double yates_term(double obs, double ex) {
    if (ex <= 0.0) return 0.0;
    double df = dsub(fabs(dsub(obs, ex)), 0.5);
    if (df < 0.0) df = 0.0;
    return ddiv(dmul(df, df), ex);
}

// p-value of the Hardy-Weinberg chi-square test, operation for operation.
VCFX_HD double hwe_pvalue(int hr, int het, int ha) {
    int N = hr + het + ha;
    if (N < 1) return 1.0;
    double dN = (double)N;
    double p = ddiv(dadd(dmul(2.0, (double)hr), (double)het), dmul(2.0, dN));
    double q = dsub(1.0, p);
    if (p <= 0.0 || p >= 1.0) return 1.0;
    double e0 = dmul(dmul(dN, p), p);
    double e1 = dmul(dmul(dmul(dN, 2.0), p), q);
    double e2 = dmul(dmul(dN, q), q);
    double chi2 = dadd(dadd(yates_term((double)hr, e0), yates_term((double)het, e1)),
                       yates_term((double)ha, e2));
    if (chi2 <= 0.0) return 1.0;
    if (chi2 > 700.0) return 0.0;
    double x = dsqrt(dmul(chi2, 0.5));
    double t = ddiv(1.0, dadd(1.0, dmul(0.3275911, x)));
    double y = dmul(t, dadd(0.254829592, dmul(t, dadd(-0.284496736, dmul(t, dadd(1.421413741,
               dmul(t, dadd(-1.453152027, dmul(t, 1.061405429)))))))));
    // -x*x: the reference negates x first, then multiplies; the product is the same value
    return dmul(y, exp(dmul(-x, x)));
}

}  // namespace vcfx
