/*
 * vcfx_oracle.c — CPU restatement of the VCFX hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Not shipped, not linked into libvcfx_cuda, never used as a fallback: it exists so the
 * parity tests can compare the CUDA path against the reference's semantics on any input,
 * including on the GPU box where /root/reference is absent.  It is pinned against the
 * compiled reference tools (oracle/_ref/VCFX_*) by tests/test_oracle_golden.py and
 * against tests/golden/.
 *
 * Written from the behaviour of the reference (file:line cited per rule), organised
 * differently: one generic line walker + a per-line tab table, then one routine per tool
 * and input mode.  Build with -ffp-contract=off: the reference's x86 build has no FMA.
 */
#define _GNU_SOURCE
#include "vcfx_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ output buffer */
typedef struct { char *p; size_t n, cap; } obuf;

static void ob_need(obuf *o, size_t extra) {
    if (o->n + extra <= o->cap) return;
    size_t c = o->cap ? o->cap : 4096;
    while (c < o->n + extra) c *= 2;
    o->p = (char *)realloc(o->p, c);
    o->cap = c;
}
static void ob_put(obuf *o, const char *s, size_t n) {
    ob_need(o, n); memcpy(o->p + o->n, s, n); o->n += n;
}
static void ob_str(obuf *o, const char *s) { ob_put(o, s, strlen(s)); }
static void ob_ch(obuf *o, char c) { ob_need(o, 1); o->p[o->n++] = c; }

static void put_int(obuf *o, long long v);
static void res_init(oracle_result *r) { memset(r, 0, sizeof *r); }
static void res_take(oracle_result *r, obuf *o) {
    r->out = o->p ? o->p : (char *)calloc(1, 1); r->out_len = o->n;
}
void oracle_free(oracle_result *r) { if (r && r->out) { free(r->out); r->out = NULL; } }

/* ------------------------------------------------------------------ line walking
 * Both reader styles of the reference see the same line set: "while (p < end)" over an
 * mmap and std::getline both yield one line per '\n' plus a final unterminated remainder
 * when it is non-empty. */
typedef struct { const char *s, *e; int terminated; } line_t;

static int next_line(const char *buf, size_t n, size_t *pos, line_t *ln) {
    if (*pos >= n) return 0;
    const char *s = buf + *pos;
    const char *nl = (const char *)memchr(s, '\n', n - *pos);
    ln->s = s;
    if (nl) { ln->e = nl; ln->terminated = 1; *pos = (size_t)(nl - buf) + 1; }
    else    { ln->e = buf + n; ln->terminated = 0; *pos = n; }
    return 1;
}

/* k-th tab-separated field of [s,e) in the mmap tools' sense: it exists only when its first
 * byte lies before the line end (allele_freq_calc.cpp:247-256, hwe_tester.cpp:321-336),
 * so a field that would start exactly at the line end is "absent". Returns 1 if present. */
static int field_at(const char *s, const char *e, int k, const char **fs, const char **fe) {
    const char *p = s;
    for (int i = 0; i < k; ++i) {
        const char *t = (const char *)memchr(p, '\t', (size_t)(e - p));
        if (!t) { *fs = *fe = e; return 0; }
        p = t + 1;
    }
    if (p >= e) { *fs = *fe = e; return 0; }
    const char *t = (const char *)memchr(p, '\t', (size_t)(e - p));
    *fs = p; *fe = t ? t : e;
    return 1;
}

/* ------------------------------------------------------------------ allele_freq_calc */

/* allele_freq_calc.cpp:298-316 — index of the ':'-separated FORMAT key equal to "GT". */
int oracle_gt_index(const char *f, size_t n) {
    int idx = 0; size_t i = 0;
    while (i < n) {
        size_t j = i;
        while (j < n && f[j] != ':') ++j;
        if (j - i == 2 && f[i] == 'G' && f[i + 1] == 'T') return idx;
        ++idx;
        i = (j < n) ? j + 1 : j;
    }
    return -1;
}

/* allele_freq_calc.cpp:262-293 — every '/'- or '|'-separated token that is non-empty, does
 * not begin with '.', and is all digits counts once; it is ALT when any digit is not '0'. */
void oracle_af_counts(const char *g, size_t n, int *alt, int *total) {
    size_t i = 0;
    while (i < n) {
        while (i < n && (g[i] == '/' || g[i] == '|')) ++i;
        if (i >= n) break;
        size_t j = i;
        while (j < n && g[j] != '/' && g[j] != '|') ++j;
        if (g[i] != '.') {
            int numeric = 1, zero = 1;
            for (size_t k = i; k < j; ++k) {
                if (g[k] < '0' || g[k] > '9') { numeric = 0; break; }
                if (g[k] != '0') zero = 0;
            }
            if (numeric) { ++*total; if (!zero) ++*alt; }
        }
        i = j;
    }
}

/* allele_freq_calc.cpp:321-337 + :438-441 — gt_index-th ':' piece of one sample column. */
static void af_sample(const char *s, const char *e, int gt_index, int *alt, int *total) {
    const char *p = s;
    for (int i = 0; i < gt_index && p < e; ++i) {
        while (p < e && *p != ':') ++p;
        if (p < e) ++p;
    }
    if (p >= e) return;
    const char *q = p;
    while (q < e && *q != ':') ++q;
    oracle_af_counts(p, (size_t)(q - p), alt, total);
}

/* allele_freq_calc.cpp:119-143 — FILE-mode text: trunc(v*10000+0.5), two roundings. */
int oracle_fmt_af_file(double v, char *dst) {
    char *d = dst;
    if (v < 0) { *d++ = '-'; v = -v; }
    double scaled_d = v * 10000.0;
    scaled_d = scaled_d + 0.5;
    unsigned long long sc = (unsigned long long)scaled_d;
    d += sprintf(d, "%llu", sc / 10000ULL);
    unsigned fr = (unsigned)(sc % 10000ULL);
    *d++ = '.';
    *d++ = (char)('0' + (fr / 1000) % 10);
    *d++ = (char)('0' + (fr / 100) % 10);
    *d++ = (char)('0' + (fr / 10) % 10);
    *d++ = (char)('0' + fr % 10);
    return (int)(d - dst);
}
/* allele_freq_calc.cpp:553-555 — STDIN-mode text: iostream fixed/setprecision(4). */
int oracle_fmt_af_stdin(double v, char *dst) { return sprintf(dst, "%.4f", v); }

static const char AF_HEADER[] = "CHROM\tPOS\tID\tREF\tALT\tAllele_Frequency\n";

static int af_file(const char *in, size_t n, oracle_result *r) {
    obuf o = {0}; size_t pos = 0; line_t ln; int seen_chrom = 0;
    ob_str(&o, AF_HEADER);                                   /* :353 */
    while (next_line(in, n, &pos, &ln)) {
        const char *s = ln.s, *e = ln.e;
        if (e > s && e[-1] == '\r') --e;                     /* :363-364 */
        if (s == e) continue;                                /* :366 */
        if (*s == '#') {                                     /* :372-380 */
            if (e - s >= 6 && memcmp(s, "#CHROM", 6) == 0) seen_chrom = 1;
            continue;
        }
        if (!seen_chrom) { r->warnings++; continue; }        /* :382-386 */
        r->data_lines++;
        const char *fs[9], *fe[9];
        for (int k = 0; k < 5; ++k) field_at(s, e, k, &fs[k], &fe[k]);
        field_at(s, e, 8, &fs[8], &fe[8]);
        if (fs[8] == fe[8]) continue;                        /* :398 */
        int gi = oracle_gt_index(fs[8], (size_t)(fe[8] - fs[8]));
        if (gi < 0) continue;                                /* :413 */
        int alt = 0, total = 0;
        /* samples: everything after the 9th tab (:423-445) */
        const char *p = s; int tabs = 0;
        while (tabs < 9) {
            const char *t = (const char *)memchr(p, '\t', (size_t)(e - p));
            if (!t) { p = e; break; }
            p = t + 1; ++tabs;
        }
        while (p < e) {
            const char *t = (const char *)memchr(p, '\t', (size_t)(e - p));
            const char *se = t ? t : e;
            af_sample(p, se, gi, &alt, &total);
            if (!t) break;
            p = t + 1;
        }
        double f = total > 0 ? (double)alt / (double)total : 0.0;   /* :448-449 */
        for (int k = 0; k < 5; ++k) { ob_put(&o, fs[k], (size_t)(fe[k] - fs[k])); ob_ch(&o, '\t'); }
        char nb[40]; int l = oracle_fmt_af_file(f, nb);
        ob_put(&o, nb, (size_t)l); ob_ch(&o, '\n');
        r->rows++;
    }
    res_take(r, &o);
    return 0;
}

static int af_stdin(const char *in, size_t n, oracle_result *r) {
    obuf o = {0}; size_t pos = 0; line_t ln; int seen_chrom = 0;
    if (n == 0) {                      /* :638-641 empty stdin prints help, rc 1 (help text not restated) */
        r->rc = 1; res_take(r, &o); return 0;
    }
    ob_str(&o, AF_HEADER);                                   /* :487 */
    while (next_line(in, n, &pos, &ln)) {
        const char *s = ln.s, *e = ln.e;                     /* no '\r' handling in this mode */
        if (s == e) continue;                                /* :490 */
        if (*s == '#') {                                     /* :492-497 */
            if (e - s >= 6 && memcmp(s, "#CHROM", 6) == 0) seen_chrom = 1;
            continue;
        }
        if (!seen_chrom) { r->warnings++; continue; }        /* :499-502 */
        /* split on tabs; nothing is pushed for an empty tail after a final tab (:509-518) */
        const char *fs[16], *fe[16]; int nf = 0; size_t nfields = 0;
        const char *p = s; const char *samples = e;
        while (p < e) {
            const char *t = (const char *)memchr(p, '\t', (size_t)(e - p));
            const char *q = t ? t : e;
            if (nf < 9) { fs[nf] = p; fe[nf] = q; ++nf; }
            ++nfields;
            if (nfields == 9) samples = t ? t + 1 : e;
            if (!t) break;
            p = t + 1;
        }
        if (nfields < 9) { r->warnings++; continue; }        /* :520-523 */
        r->data_lines++;
        int gi = oracle_gt_index(fs[8], (size_t)(fe[8] - fs[8]));
        if (gi < 0) continue;                                /* :537 */
        int alt = 0, total = 0;
        p = samples;
        while (p < e) {
            const char *t = (const char *)memchr(p, '\t', (size_t)(e - p));
            const char *se = t ? t : e;
            af_sample(p, se, gi, &alt, &total);
            if (!t) break;
            p = t + 1;
        }
        double f = total > 0 ? (double)alt / (double)total : 0.0;
        for (int k = 0; k < 5; ++k) { ob_put(&o, fs[k], (size_t)(fe[k] - fs[k])); ob_ch(&o, '\t'); }
        char nb[64]; int l = oracle_fmt_af_stdin(f, nb);
        ob_put(&o, nb, (size_t)l); ob_ch(&o, '\n');
        r->rows++;
    }
    res_take(r, &o);
    return 0;
}

int oracle_allele_freq(const char *in, size_t n, int mode, oracle_result *r) {
    res_init(r);
    return mode == ORACLE_FILE ? af_file(in, n, r) : af_stdin(in, n, r);
}

/* ------------------------------------------------------------------ hwe_tester */

/* hwe_tester.cpp:339-378 — class of one sample column: 0 homRef, 1 het, 2 homAlt, -1 skip.
 * Only the first ':' piece is read; anything after the second allele is ignored. */
int oracle_hwe_class(const char *p, size_t n) {
    const char *e = p + n;
    if (p >= e) return -1;
    const char *c = (const char *)memchr(p, ':', (size_t)(e - p));
    if (c) e = c;
    while (p < e && (*p == ' ' || *p == '\r')) ++p;
    if (p >= e) return -1;
    if (*p < '0' || *p > '9') return -1;                 /* covers '.' */
    int a1 = 0;
    while (p < e && *p >= '0' && *p <= '9') { a1 = a1 * 10 + (*p - '0'); ++p; }
    if (p >= e || (*p != '/' && *p != '|')) return -1;
    ++p;
    if (p >= e || *p < '0' || *p > '9') return -1;
    int a2 = 0;
    while (p < e && *p >= '0' && *p <= '9') { a2 = a2 * 10 + (*p - '0'); ++p; }
    if (a1 > 1 || a2 > 1) return -1;                     /* int overflow on absurd digit runs is UB upstream */
    if (a1 == 0 && a2 == 0) return 0;
    if (a1 == 1 && a2 == 1) return 2;
    return 1;
}

/* hwe_tester.cpp:290-315 and :278-287 — Yates-corrected chi-square, 1-df p-value by the
 * Abramowitz-Stegun 7.1.26 erfc polynomial.  Operation order is kept term by term. */
/* how many of n value pairs print differently as FILE-mode text (truncated, hwe_tester.cpp:236-268) and as "%.6f" */
void oracle_p_text_diffs(const double *a, const double *b, size_t n, size_t *file_diffs, size_t *stdin_diffs) {
    size_t fd = 0, sd = 0;
    for (size_t i = 0; i < n; ++i) {
        char x[64], y[64];
        int lx = oracle_fmt_p_file(a[i], x), ly = oracle_fmt_p_file(b[i], y);
        if (lx != ly || memcmp(x, y, (size_t)lx) != 0) ++fd;
        lx = oracle_fmt_p_stdin(a[i], x); ly = oracle_fmt_p_stdin(b[i], y);
        if (lx != ly || memcmp(x, y, (size_t)lx) != 0) ++sd;
    }
    *file_diffs = fd; *stdin_diffs = sd;
}

void oracle_hwe_pvalues(const int *counts, size_t n, double *out) {
    for (size_t i = 0; i < n; ++i) out[i] = oracle_hwe_pvalue(counts[3 * i], counts[3 * i + 1], counts[3 * i + 2]);
}

double oracle_hwe_pvalue(int hr, int het, int ha) {
    int N = hr + het + ha;
    if (N < 1) return 1.0;
    double p = (2.0 * hr + het) / (2.0 * N);
    double q = 1.0 - p;
    if (p <= 0.0 || p >= 1.0) return 1.0;
    double ex[3]; double ob[3] = { (double)hr, (double)het, (double)ha };
    ex[0] = N * p * p;
    ex[1] = N * 2.0 * p * q;
    ex[2] = N * q * q;
    double chi2 = 0.0;
    for (int i = 0; i < 3; ++i) {
        double t = 0.0;
        if (ex[i] > 0.0) {
            double d = fabs(ob[i] - ex[i]) - 0.5;
            if (d < 0.0) d = 0.0;
            t = (d * d) / ex[i];
        }
        chi2 = (i == 0) ? t : chi2 + t;
    }
    if (chi2 <= 0.0) return 1.0;
    if (chi2 > 700.0) return 0.0;
    double x = sqrt(chi2 * 0.5);
    double t = 1.0 / (1.0 + 0.3275911 * x);
    double y = t * (0.254829592 + t * (-0.284496736 + t * (1.421413741 +
               t * (-1.453152027 + t * 1.061405429))));
    return y * exp(-x * x);
}

/* hwe_tester.cpp:236-268 — FILE-mode text: six truncated decimal digits. */
int oracle_fmt_p_file(double v, char *dst) {
    char *d = dst;
    if (v < 0) { *d++ = '-'; v = -v; }
    long long ip = (long long)v;
    double fr = v - (double)ip;
    d += sprintf(d, "%lld", ip);
    *d++ = '.';
    for (int i = 0; i < 6; ++i) {
        fr = fr * 10.0;
        int dg = (int)fr;
        *d++ = (char)('0' + dg);
        fr = fr - (double)dg;
    }
    return (int)(d - dst);
}
/* hwe_tester.cpp:605-606 — STDIN-mode text. */
int oracle_fmt_p_stdin(double v, char *dst) { return sprintf(dst, "%.6f", v); }

static const char HWE_HEADER[] = "CHROM\tPOS\tID\tREF\tALT\tHWE_pvalue\n";

static int has_comma(const char *s, const char *e) { return memchr(s, ',', (size_t)(e - s)) != NULL; }

static int hwe_file(const char *in, size_t n, oracle_result *r) {
    obuf o = {0}; size_t pos = 0; line_t ln;
    if (n == 0) { res_take(r, &o); return 0; }               /* :456 */
    ob_str(&o, HWE_HEADER);                                  /* :463 */
    while (next_line(in, n, &pos, &ln)) {
        const char *s = ln.s, *e = ln.e;
        if (e > s && e[-1] == '\r') --e;                     /* :480 */
        if (s == e || *s == '#') continue;                   /* :482, :466-472 */
        r->data_lines++;
        const char *fs[9], *fe[9];
        for (int k = 0; k < 5; ++k) field_at(s, e, k, &fs[k], &fe[k]);
        if (fs[0] == fe[0] || fs[1] == fe[1] || fs[4] == fe[4]) continue;   /* :497 */
        if (has_comma(fs[4], fe[4])) continue;               /* :503 */
        field_at(s, e, 8, &fs[8], &fe[8]);
        if (fe[8] - fs[8] < 2 || fs[8][0] != 'G' || fs[8][1] != 'T') continue;  /* :510 */
        const char *sp, *dummy;
        if (!field_at(s, e, 9, &sp, &dummy)) continue;       /* :516-520 */
        int c[3] = {0, 0, 0};
        while (sp < e) {                                     /* :526-536 */
            const char *t = (const char *)memchr(sp, '\t', (size_t)(e - sp));
            const char *se = t ? t : e;
            int k = oracle_hwe_class(sp, (size_t)(se - sp));
            if (k >= 0) c[k]++;
            sp = se + 1;
        }
        double pv = oracle_hwe_pvalue(c[0], c[1], c[2]);
        for (int k = 0; k < 5; ++k) { ob_put(&o, fs[k], (size_t)(fe[k] - fs[k])); ob_ch(&o, '\t'); }
        char nb[64]; int l = oracle_fmt_p_file(pv, nb);
        ob_put(&o, nb, (size_t)l); ob_ch(&o, '\n');
        r->rows++;
    }
    res_take(r, &o);
    return 0;
}

static int hwe_stdin(const char *in, size_t n, oracle_result *r) {
    obuf o = {0}; size_t pos = 0; line_t ln;
    ob_str(&o, HWE_HEADER);                                  /* :566, even for empty input */
    while (next_line(in, n, &pos, &ln)) {
        const char *s = ln.s, *e = ln.e;
        if (s == e) continue;                                /* :573 */
        if (e[-1] == '\r') --e;                              /* :574 */
        if (s == e || *s == '#') continue;                   /* :575 */
        r->data_lines++;
        /* vcfx::split_tabs keeps empty fields, including a trailing one (vcfx_io.h:59-76) */
        const char *fs[10], *fe[10]; size_t nfields = 0; const char *p = s;
        for (;;) {
            const char *t = (const char *)memchr(p, '\t', (size_t)(e - p));
            const char *q = t ? t : e;
            if (nfields < 10) { fs[nfields] = p; fe[nfields] = q; }
            ++nfields;
            if (!t) break;
            p = t + 1;
        }
        if (nfields < 10) continue;                          /* :578 */
        if (has_comma(fs[4], fe[4])) continue;               /* :586 */
        if (fe[8] - fs[8] < 2 || fs[8][0] != 'G' || fs[8][1] != 'T') continue;  /* :590 */
        int c[3] = {0, 0, 0};
        const char *sp = fs[9];
        for (;;) {                                           /* :595-601 every field incl. empty ones */
            const char *t = (const char *)memchr(sp, '\t', (size_t)(e - sp));
            const char *se = t ? t : e;
            int k = oracle_hwe_class(sp, (size_t)(se - sp));
            if (k >= 0) c[k]++;
            if (!t) break;
            sp = t + 1;
        }
        double pv = oracle_hwe_pvalue(c[0], c[1], c[2]);
        for (int k = 0; k < 5; ++k) { ob_put(&o, fs[k], (size_t)(fe[k] - fs[k])); ob_ch(&o, '\t'); }
        char nb[64]; int l = oracle_fmt_p_stdin(pv, nb);
        ob_put(&o, nb, (size_t)l); ob_ch(&o, '\n');
        r->rows++;
    }
    res_take(r, &o);
    return 0;
}

int oracle_hwe(const char *in, size_t n, int mode, oracle_result *r) {
    res_init(r);
    return mode == ORACLE_FILE ? hwe_file(in, n, r) : hwe_stdin(in, n, r);
}

/* ------------------------------------------------------------------ missing_detector */

/* pointer to the start of field k, or e when the line has fewer tabs
 * (missing_detector.cpp:277-283 skipToField) */
static const char *md_skip(const char *s, const char *e, int k) {
    const char *p = s;
    for (int i = 0; i < k && p < e; ++i) {
        const char *t = (const char *)memchr(p, '\t', (size_t)(e - p));
        p = t ? t + 1 : e;
    }
    return p;
}

/* missing_detector.cpp:290-336 — a sample is "missing" when the first ':' piece holds a '.'
 * that touches a piece boundary or a '/' '|' separator on either side. FORMAT is ignored. */
static int md_region_missing(const char *p, const char *e) {
    while (p < e) {
        const char *t = (const char *)memchr(p, '\t', (size_t)(e - p));
        const char *se = t ? t : e;
        const char *ge = (const char *)memchr(p, ':', (size_t)(se - p));
        if (!ge) ge = se;
        for (const char *g = p; g < ge; ++g) {
            if (*g != '.') continue;
            int prev_sep = (g == p) || g[-1] == '/' || g[-1] == '|';
            int next_sep = (g + 1 >= ge) || g[1] == '/' || g[1] == '|';
            if (prev_sep || next_sep) return 1;
        }
        if (!t) break;
        p = t + 1;
    }
    return 0;
}

static void md_emit_flagged(obuf *o, const char *s, const char *e) {
    /* INFO = field 7 (missing_detector.cpp:546-568, :893-906) */
    const char *is = md_skip(s, e, 7);
    const char *ie = (const char *)memchr(is, '\t', (size_t)(e - is));
    if (!ie) ie = e;
    ob_put(o, s, (size_t)(is - s));
    size_t il = (size_t)(ie - is);
    if (il == 0 || (il == 1 && is[0] == '.')) {
        ob_str(o, "MISSING_GENOTYPES=1");
    } else {
        ob_put(o, is, il);
        if (ie[-1] != ';') ob_ch(o, ';');
        ob_str(o, "MISSING_GENOTYPES=1");
    }
    ob_put(o, ie, (size_t)(e - ie));
    ob_ch(o, '\n');
}

static int md_file(const char *in, size_t n, oracle_result *r) {
    obuf o = {0}; size_t pos = 0; line_t ln;
    /* Pre-scan (missing_detector.cpp:371-392, 347-369, single-thread form = "-t 1"):
     * skip the leading '#' block, then look for ANY '.' after the 9th tab of every
     * newline-TERMINATED line; an unterminated last line is not looked at (:354). */
    int any_dot = 0;
    {
        size_t p2 = 0; line_t l2; int in_header = 1;
        while (next_line(in, n, &p2, &l2)) {
            if (in_header && *l2.s == '#') continue;   /* an empty line ends the '#' block */
            in_header = 0;
            if (!l2.terminated) break;
            r->data_lines++;          /* the fast path reports this count on stderr */
            const char *ss = md_skip(l2.s, l2.e, 9);
            if (ss < l2.e && memchr(ss, '.', (size_t)(l2.e - ss))) { any_dot = 1; break; }
        }
    }
    if (!any_dot) {                                          /* :456-477 verbatim copy */
        ob_put(&o, in, n);
        res_take(r, &o);
        return 0;
    }
    r->data_lines = 0;
    while (next_line(in, n, &pos, &ln)) {                    /* :499-579 */
        const char *s = ln.s, *e = ln.e;
        size_t raw = (size_t)(ln.e - ln.s) + (ln.terminated ? 1 : 0);
        if (e > s && e[-1] == '\r') --e;                     /* :505-506 */
        if (s == e || *s == '#') { ob_put(&o, s, raw); continue; }      /* :511 */
        r->data_lines++;
        const char *ss = md_skip(s, e, 9);
        if (ss >= e || !md_region_missing(ss, e)) { ob_put(&o, s, raw); continue; }
        r->flagged++;
        md_emit_flagged(&o, s, e);          /* '\r' dropped, '\n' always added (:571-574) */
    }
    res_take(r, &o);
    return 0;
}

static int md_stdin(const char *in, size_t n, oracle_result *r) {
    obuf o = {0}; size_t pos = 0; line_t ln;
    while (next_line(in, n, &pos, &ln)) {                    /* :864-910; no '\r' handling */
        const char *s = ln.s, *e = ln.e;
        if (s == e) { ob_ch(&o, '\n'); continue; }
        if (*s == '#') { ob_put(&o, s, (size_t)(e - s)); ob_ch(&o, '\n'); continue; }
        r->data_lines++;
        const char *ss = md_skip(s, e, 9);
        if (ss >= e || !md_region_missing(ss, e)) {
            ob_put(&o, s, (size_t)(e - s)); ob_ch(&o, '\n'); continue;
        }
        r->flagged++;
        md_emit_flagged(&o, s, e);
    }
    res_take(r, &o);
    return 0;
}

int oracle_missing(const char *in, size_t n, int mode, oracle_result *r) {
    res_init(r);
    return mode == ORACLE_FILE ? md_file(in, n, r) : md_stdin(in, n, r);
}

/* ------------------------------------------------------------------ nonref_filter (§8 f2: a sibling tool on the same
 * scan -> GT -> per-line predicate shape).  A data line is dropped when EVERY sample is "definitely" homozygous
 * reference; everything else passes through, each written line closed by '\n'. */

/* VCFX_nonref_filter.cpp:286-302 (file mode, inlined in allSamplesHomRefDirect): a GT of exactly three bytes must be
 * 0/0 or 0|0; any other non-empty GT is hom-ref when it holds nothing but '0', '/' and '|'. */
static int nr_homref_file(const char *g, const char *ge) {
    size_t n = (size_t)(ge - g);
    if (n == 3) return g[0] == '0' && (g[1] == '/' || g[1] == '|') && g[2] == '0';
    if (n == 0) return 0;
    for (const char *p = g; p < ge; ++p) if (*p != '/' && *p != '|' && *p != '0') return 0;
    return 1;
}
/* :419-449 isDefinitelyHomRef (stdin mode): empty is not; otherwise nothing but '0', '/' and '|' (the special cases in
 * front of the general scan decide the same way) */
static int nr_homref_stdin(const char *g, const char *ge) {
    if (g == ge) return 0;
    for (const char *p = g; p < ge; ++p) if (*p != '/' && *p != '|' && *p != '0') return 0;
    return 1;
}
/* :179-198 extractNthField: the n-th ':' piece of [s,e), empty when there are fewer */
static void nr_nth_piece(const char *s, const char *e, int k, const char **gs, const char **ge) {
    const char *fs = s; int idx = 0;
    for (const char *p = s; p <= e; ++p) {
        if (p == e || *p == ':') {
            if (idx == k) { *gs = fs; *ge = p; return; }
            ++idx; fs = p + 1;
        }
    }
    *gs = *ge = e;
}
/* :224-231 skipToField: start of field k (k tabs passed), NULL when the line has fewer tabs */
static const char *nr_skip(const char *p, const char *e, int k) {
    int idx = 0;
    while (p < e && idx < k) { if (*p == '\t') ++idx; ++p; }
    return idx == k ? p : NULL;
}

static int nr_file(const char *in, size_t n, oracle_result *r) {       /* :458-548 filterNonRefMmap */
    obuf o = {0}; size_t pos = 0; line_t ln; int header = 0;
    while (next_line(in, n, &pos, &ln)) {
        const char *s = ln.s, *e = ln.e;
        if (e > s && e[-1] == '\r') --e;                               /* :484-486 */
        if (s == e) { ob_ch(&o, '\n'); continue; }                     /* :489-493 */
        if (*s == '#') {                                               /* :496-504 */
            ob_put(&o, s, (size_t)(e - s)); ob_ch(&o, '\n');
            if (e - s >= 6 && memcmp(s + 1, "CHROM", 5) == 0) header = 1;
            continue;
        }
        r->data_lines++;
        int keep = 1;
        if (!header) r->warnings++;                                    /* :507-512 passes the line */
        else {
            const char *f = nr_skip(s, e, 8);                          /* :515 */
            if (f) {
                const char *fe = f; while (fe < e && *fe != '\t') ++fe;
                int gi = oracle_gt_index(f, (size_t)(fe - f));         /* :317-335 findGTIndex, :526-529 (the cache changes nothing) */
                if (gi >= 0) {
                    const char *p = nr_skip(s, e, 9);                  /* :248-309 allSamplesHomRefDirect */
                    if (p) {
                        int all = 1;
                        while (p < e) {
                            const char *se = p; while (se < e && *se != '\t') ++se;
                            if (se == p) { all = 0; break; }           /* empty sample: keep */
                            const char *gs, *ge;
                            if (gi == 0) { gs = p; ge = memchr(p, ':', (size_t)(se - p)); if (!ge) ge = se; }
                            else nr_nth_piece(p, se, gi, &gs, &ge);
                            if (!nr_homref_file(gs, ge)) { all = 0; break; }
                            p = se; if (p < e && *p == '\t') ++p;
                        }
                        if (all) keep = 0;
                    }
                }
            }
        }
        if (keep) { ob_put(&o, s, (size_t)(e - s)); ob_ch(&o, '\n'); r->rows++; }
    }
    res_take(r, &o);
    return 0;
}

static int nr_stdin(const char *in, size_t n, oracle_result *r) {      /* :553-631 filterNonRef */
    obuf o = {0}; size_t pos = 0; line_t ln; int header = 0;
    while (next_line(in, n, &pos, &ln)) {
        const char *s = ln.s, *e = ln.e;                               /* getline: '\r' stays */
        if (s == e) { ob_ch(&o, '\n'); continue; }
        if (*s == '#') {
            ob_put(&o, s, (size_t)(e - s)); ob_ch(&o, '\n');
            if (e - s >= 6 && memcmp(s, "#CHROM", 6) == 0) header = 1;
            continue;
        }
        r->data_lines++;
        int keep = 1;
        if (!header) r->warnings++;
        else {
            int tabs = 0; for (const char *p = s; p < e; ++p) tabs += (*p == '\t');
            if (tabs + 1 >= 10) {                                      /* :578-582 */
                const char *f = nr_skip(s, e, 8), *fe = f; while (fe < e && *fe != '\t') ++fe;
                /* FORMAT split on ':' with std::getline (:584-590): a piece equal to "GT"; the pieces std::getline
                 * drops (nothing after a final ':', nothing at all for an empty string) never equal "GT" */
                int gi = oracle_gt_index(f, (size_t)(fe - f));
                if (gi >= 0) {
                    int all = 1;
                    const char *p = fe + 1;                            /* field 9 (exists: >= 9 tabs) */
                    for (;;) {
                        const char *se = p; while (se < e && *se != '\t') ++se;
                        /* the sample split on ':' with std::getline (:610-616): pieces = ':' count + 1, minus one when
                         * the sample is empty or ends with ':' */
                        int pieces = 0;
                        if (se > p) { pieces = 1; for (const char *q = p; q < se; ++q) pieces += (*q == ':'); if (se[-1] == ':') --pieces; }
                        if (gi >= pieces) { all = 0; break; }          /* :618-621 */
                        const char *gs, *ge; nr_nth_piece(p, se, gi, &gs, &ge);
                        if (!nr_homref_stdin(gs, ge)) { all = 0; break; }
                        if (se >= e) break;
                        p = se + 1;
                    }
                    if (all) keep = 0;
                }
            }
        }
        if (keep) { ob_put(&o, s, (size_t)(e - s)); ob_ch(&o, '\n'); r->rows++; }
    }
    res_take(r, &o);
    return 0;
}

int oracle_nonref_filter(const char *in, size_t n, int mode, oracle_result *r) {
    res_init(r);
    return mode == ORACLE_FILE ? nr_file(in, n, r) : nr_stdin(in, n, r);
}

/* ------------------------------------------------------------------ indexer (§8 f4): CHROM, POS and the byte offset of
 * every data line behind the "#CHROM" line.  out: rows = data rows written; warnings = 1 when the tool prints
 * "Error: no #CHROM header found before variant lines." */

static int ix_space(int c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; }   /* std::isspace, C locale */

static int ix_file(const char *in, size_t n, oracle_result *r) {         /* VCFX_indexer.cpp:205-322 createVCFIndexMmap */
    obuf o = {0}; size_t pos = 0; line_t ln; int found = 0, warned = 0, saw_header = 0;
    while (next_line(in, n, &pos, &ln)) {
        const char *s = ln.s, *e = ln.e;
        long long off = (long long)(s - in);
        if (e > s && e[-1] == '\r') --e;                                   /* :263-266 */
        if (s == e) continue;
        const char *p = s;
        while (p < e && (*p == ' ' || *p == '\t')) ++p;                    /* :271-274 */
        if (p < e && *p == '#') {
            saw_header = 1;
            if (!found && e - s >= 6 && e - p >= 6 && memcmp(p, "#CHROM", 6) == 0) {   /* :108-122 isChromHeaderLine */
                found = 1; ob_str(&o, "CHROM\tPOS\tFILE_OFFSET\n");
            }
        } else if (found) {                                                /* :74-105 extractChromPos */
            if (p >= e) continue;
            const char *cs = p;
            while (p < e && *p != '\t') ++p;
            if (p == cs || p >= e) continue;
            const char *ce = p; ++p;
            unsigned long long v = 0;                                      /* int64 arithmetic that wraps, as compiled */
            while (p < e && *p >= '0' && *p <= '9') { v = v * 10ULL + (unsigned long long)(*p - '0'); ++p; }
            long long pv = (long long)v;
            if (pv <= 0) continue;
            ob_put(&o, cs, (size_t)(ce - cs)); ob_ch(&o, '\t'); put_int(&o, pv); ob_ch(&o, '\t'); put_int(&o, off); ob_ch(&o, '\n');
            r->rows++;
        } else if (!saw_header && !warned) { warned = 1; r->warnings = 1; }
    }
    res_take(r, &o);
    return 0;
}

static int ix_stdin(const char *in, size_t n, oracle_result *r) {        /* :329-443 createVCFIndex */
    obuf o = {0}; size_t pos = 0; line_t ln; int found = 0, warned = 0, saw_header = 0;
    while (next_line(in, n, &pos, &ln)) {
        const char *s = ln.s, *e = ln.e;
        long long off = (long long)(s - in);
        if (e > s && e[-1] == '\r') --e;                                   /* :421-423 / :437-439 */
        if (s == e) continue;
        const char *t = s;
        while (t < e && ix_space((unsigned char)*t)) ++t;                  /* ltrim */
        if (t < e && *t == '#') {
            saw_header = 1;
            if (!found) {                                                  /* fields[0] == "#CHROM" and a second field */
                const char *tab = memchr(t, '\t', (size_t)(e - t));
                if (tab && tab - t == 6 && memcmp(t, "#CHROM", 6) == 0) { found = 1; ob_str(&o, "CHROM\tPOS\tFILE_OFFSET\n"); }
            }
            continue;
        }
        if (!found) { if (!saw_header && !warned) { warned = 1; r->warnings = 1; } continue; }
        const char *tab = memchr(s, '\t', (size_t)(e - s));                /* splitTabs(line): the line as it is */
        if (!tab) continue;
        const char *ps = tab + 1, *pe = memchr(ps, '\t', (size_t)(e - ps));
        if (!pe) pe = e;
        /* std::stoll: leading isspace, an optional sign, at least one digit, out of range throws (line skipped) */
        const char *q = ps;
        while (q < pe && ix_space((unsigned char)*q)) ++q;
        int neg = 0;
        if (q < pe && (*q == '+' || *q == '-')) { neg = (*q == '-'); ++q; }
        if (q >= pe || *q < '0' || *q > '9') continue;
        unsigned long long v = 0; int over = 0;
        const unsigned long long lim = neg ? 9223372036854775808ULL : 9223372036854775807ULL;
        while (q < pe && *q >= '0' && *q <= '9') {
            unsigned d = (unsigned)(*q - '0');
            if (v > (lim - d) / 10ULL) over = 1;
            v = v * 10ULL + d; ++q;
        }
        if (over) continue;
        ob_put(&o, s, (size_t)(tab - s)); ob_ch(&o, '\t');
        if (neg) { if (v) ob_ch(&o, '-'); }
        { char nb[32]; int l = sprintf(nb, "%llu", v); ob_put(&o, nb, (size_t)l); }
        ob_ch(&o, '\t'); put_int(&o, off); ob_ch(&o, '\n');
        r->rows++;
    }
    res_take(r, &o);
    return 0;
}

int oracle_indexer(const char *in, size_t n, int mode, oracle_result *r) {
    res_init(r);
    return mode == ORACLE_FILE ? ix_file(in, n, r) : ix_stdin(in, n, r);
}

/* ------------------------------------------------------------------ phase_checker (§8 f2): only lines whose samples are ALL
 * fully phased pass; every other data line is dropped with a message on stderr (unless -q).  The restatement returns stdout in
 * `out`; the stderr text it would print without -q is appended to *err_text when that is given (oracle_phase_checker_err). */

/* VCFX_phase_checker.cpp:218-268 isFullyPhasedFast */
static int pc_phased(const char *g, const char *ge) {
    size_t len = (size_t)(ge - g);
    if (len == 0) return 0;
    if (len == 3) return g[1] == '|' && g[0] != '.' && g[2] != '.';
    if (len == 1) return 0;
    if (g[0] == '.' && len >= 3 && (g[1] == '/' || g[1] == '|') && g[2] == '.') return 0;
    int pipe = 0; const char *as = g;
    for (const char *p = g; p < ge; ++p) {
        if (*p == '|') {
            size_t al = (size_t)(p - as);
            if (al == 0 || (al == 1 && *as == '.')) return 0;
            pipe = 1; as = p + 1;
        } else if (*p == '/') return 0;
    }
    size_t al = (size_t)(ge - as);
    if (al == 0 || (al == 1 && *as == '.')) return 0;
    return pipe;
}

static void pc_msg_unphased(obuf *err, const char *s, const char *e) {     /* :541-555 / :645-648 */
    const char *t1 = memchr(s, '\t', (size_t)(e - s));
    if (!t1) return;
    const char *t2 = memchr(t1 + 1, '\t', (size_t)(e - t1 - 1));
    if (!t2) return;
    ob_str(err, "Unphased genotype found at CHROM="); ob_put(err, s, (size_t)(t1 - s));
    ob_str(err, ", POS="); ob_put(err, t1 + 1, (size_t)(t2 - t1 - 1)); ob_str(err, "; line skipped.\n");
}

static int pc_run(const char *in, size_t n, int file_mode, oracle_result *r, obuf *err) {
    obuf o = {0}; size_t pos = 0; line_t ln; int header = 0;
    /* file mode keeps the last FORMAT string and its GT index, and starts with ("", 0): an empty FORMAT column means "GT first"
     * until the first non-empty one has been looked at (:486-488, :313-316); stdin mode starts with ("", -1) (:572-574) */
    const char *cf = ""; size_t cfl = 0; int cgi = 0;
    while (next_line(in, n, &pos, &ln)) {
        const char *s = ln.s, *e = ln.e;
        if (file_mode && e > s && e[-1] == '\r') --e;                      /* :499-501 (stdin: getline keeps it) */
        if (s == e) { ob_ch(&o, '\n'); continue; }
        if (*s == '#') {
            ob_put(&o, s, (size_t)(e - s)); ob_ch(&o, '\n');
            if (e - s >= 6 && memcmp(s + 1, "CHROM", 5) == 0) header = 1;
            continue;
        }
        r->data_lines++;
        if (!header) { r->warnings++; if (err) ob_str(err, "Warning: Data line encountered before #CHROM header; skipping line.\n"); continue; }
        int tabs = 0; for (const char *p = s; p < e; ++p) tabs += (*p == '\t');
        const char *f = nr_skip(s, e, 8);
        int res;                                                           /* 1 all phased, 0 not, -1 invalid, -2 no GT (stdin) */
        if (file_mode) {                                                   /* :296-359 checkAllSamplesPhasedDirect */
            if (!f || f >= e) res = -1;
            else {
                const char *fe = f; while (fe < e && *fe != '\t') ++fe;
                if ((size_t)(fe - f) != cfl || memcmp(f, cf, cfl) != 0) { cgi = oracle_gt_index(f, (size_t)(fe - f)); cf = f; cfl = (size_t)(fe - f); }
                int gi = cgi;
                if (gi < 0) res = 0;
                else if (fe >= e) res = -1;
                else {
                    res = 1;
                    const char *p = fe + 1;
                    while (p < e) {
                        const char *se = p; while (se < e && *se != '\t') ++se;
                        if (se == p) { res = 0; break; }
                        const char *gs, *ge;
                        if (gi == 0) { gs = p; ge = memchr(p, ':', (size_t)(se - p)); if (!ge) ge = se; }
                        else nr_nth_piece(p, se, gi, &gs, &ge);
                        if (!pc_phased(gs, ge)) { res = 0; break; }
                        p = se; if (p < e) ++p;
                    }
                }
            }
        } else {                                                           /* :563-650 processVCF */
            if (tabs + 1 < 10) res = -1;
            else {
                const char *fe = f; while (fe < e && *fe != '\t') ++fe;
                int gi = oracle_gt_index(f, (size_t)(fe - f));
                if (gi < 0) res = -2;
                else {
                    res = 1;
                    const char *p = fe + 1;
                    for (;;) {
                        const char *se = p; while (se < e && *se != '\t') ++se;
                        const char *gs, *ge;
                        if (gi == 0) { gs = p; ge = memchr(p, ':', (size_t)(se - p)); if (!ge) ge = se; }
                        else nr_nth_piece(p, se, gi, &gs, &ge);
                        if (!pc_phased(gs, ge)) { res = 0; break; }
                        if (se >= e) break;
                        p = se + 1;
                    }
                }
            }
        }
        if (res == 1) { ob_put(&o, s, (size_t)(e - s)); ob_ch(&o, '\n'); r->rows++; }
        else {
            r->flagged++;
            if (err) {
                if (res == -1) ob_str(err, "Warning: Invalid VCF line with fewer than 10 columns; skipping line.\n");
                else if (res == -2) ob_str(err, "Warning: GT field not found; skipping line.\n");
                else if (file_mode) pc_msg_unphased(err, s, e);
                else {                                                     /* fields[0], fields[1]: they exist (>= 10 fields) */
                    pc_msg_unphased(err, s, e);
                }
            }
        }
    }
    res_take(r, &o);
    return 0;
}

int oracle_phase_checker(const char *in, size_t n, int mode, oracle_result *r) {
    res_init(r);
    return pc_run(in, n, mode == ORACLE_FILE, r, NULL);
}
/* the same, and the stderr text of a run without -q in a second result */
int oracle_phase_checker_err(const char *in, size_t n, int mode, oracle_result *r, oracle_result *err) {
    res_init(r); res_init(err);
    obuf eb = {0};
    int rc = pc_run(in, n, mode == ORACLE_FILE, r, &eb);
    res_take(err, &eb);
    return rc;
}

/* ------------------------------------------------------------------ inbreeding_calculator (§8 f3): per-sample F = 1 - observed
 * heterozygotes / expected heterozygotes over the biallelic sites, the expectation summed in file order in double. */

/* VCFX_inbreeding_calculator.cpp:296-339 parseGenotypeCode: the piece before the first ':', blanks / '\r' in front skipped,
 * digits, '/' or '|', digits (whatever follows them is not looked at); both alleles 0 or 1 -> 0, 1, 2; anything else -1.
 * (Allele numbers are accumulated in int; ten or more digits are outside what the restatement promises.) */
static int ib_code(const char *p, const char *end) {
    if (p >= end) return -1;
    const char *c = memchr(p, ':', (size_t)(end - p));
    if (c) end = c;
    while (p < end && (*p == ' ' || *p == '\r')) ++p;
    if (p >= end || *p < '0' || *p > '9') return -1;
    unsigned a1 = 0, a2 = 0;
    while (p < end && *p >= '0' && *p <= '9') { a1 = a1 * 10u + (unsigned)(*p - '0'); ++p; }
    if (p >= end || (*p != '/' && *p != '|')) return -1;
    ++p;
    if (p >= end || *p < '0' || *p > '9') return -1;
    while (p < end && *p >= '0' && *p <= '9') { a2 = a2 * 10u + (unsigned)(*p - '0'); ++p; }
    if ((int)a1 > 1 || (int)a2 > 1 || (int)a1 < 0 || (int)a2 < 0) return -1;
    return a1 == a2 ? (a1 == 0 ? 0 : 2) : 1;
}

typedef struct { double *sum; double *het; int *used; int *code; int n; long long variants; } ib_acc;

/* :603-636 / :760-790: one site's contribution to every sample that has a genotype code */
static void ib_site(ib_acc *a, int alt_sum, int n_good, int global, int skip_boundary, int count_boundary) {
    double global_p = (double)alt_sum / (2.0 * n_good);
    for (int s = 0; s < a->n; ++s) {
        int code = a->code[s];
        if (code < 0) continue;
        double freq;
        if (global) freq = global_p;
        else {
            int alt_ex = alt_sum - code, valid_ex = n_good - 1;
            if (valid_ex < 1) continue;
            freq = (double)alt_ex / (2.0 * valid_ex);
        }
        if (skip_boundary && (freq <= 0.0 || freq >= 1.0)) { if (count_boundary) a->used[s]++; continue; }
        a->used[s]++;
        double e_het = 2.0 * freq * (1.0 - freq);
        a->sum[s] += e_het;
        if (code == 1) a->het[s] += 1.0;
    }
}

static void ib_report(obuf *o, const ib_acc *a, const char **name, const size_t *name_len, int quiet, oracle_result *r) {   /* :641-667 / :806-826 */
    if (a->variants == 0) {
        if (!quiet) r->warnings = 3;                                       /* "No biallelic variants found." */
        for (int s = 0; s < a->n; ++s) { ob_put(o, name[s], name_len[s]); ob_str(o, "\tNA\n"); }
        return;
    }
    for (int s = 0; s < a->n; ++s) {
        ob_put(o, name[s], name_len[s]); ob_ch(o, '\t');
        if (a->used[s] == 0) { ob_str(o, "NA\n"); continue; }
        double e = a->sum[s];
        if (e <= 0.0) { ob_str(o, "1.000000\n"); continue; }
        double f = 1.0 - (a->het[s] / e);
        char nb[64]; int l = oracle_fmt_p_file(f, nb);                     /* :205-240 appendDouble: the same truncating digits as hwe_tester's */
        ob_put(o, nb, (size_t)l); ob_ch(o, '\n');
    }
}

static int ib_alloc(ib_acc *a, int n) {
    a->n = n; a->variants = 0;
    a->sum = calloc((size_t)n + 1, sizeof(double)); a->het = calloc((size_t)n + 1, sizeof(double));
    a->used = calloc((size_t)n + 1, sizeof(int)); a->code = calloc((size_t)n + 1, sizeof(int));
    return a->sum && a->het && a->used && a->code;
}
static void ib_free(ib_acc *a) { free(a->sum); free(a->het); free(a->used); free(a->code); }

/* flags: 1 = --freq-mode global, 2 = --skip-boundary, 4 = --count-boundary-as-used, 8 = -q.
 * r->warnings = which message goes to stderr: 1 "Error: Empty file.", 2 "Error: No #CHROM line or no samples found.",
 * 3 "No biallelic variants found." (not with -q), 4 "Error: No #CHROM line found.", 5 "Error: No sample columns found."
 * r->rows = sites used (at least two samples with a genotype code). */
int oracle_inbreeding(const char *in, size_t n, int mode, int flags, oracle_result *r) {
    res_init(r);
    obuf o = {0};
    const int global = flags & 1, skipb = (flags >> 1) & 1, countb = (flags >> 2) & 1, quiet = (flags >> 3) & 1;
    const char **name = NULL; size_t *name_len = NULL; int ns = 0, cap = 0;
    ib_acc a; memset(&a, 0, sizeof a);
    size_t pos = 0; line_t ln;
#define IB_PUSH(ps, pe) do { if (ns == cap) { cap = cap ? cap * 2 : 64; name = realloc(name, (size_t)cap * sizeof *name); name_len = realloc(name_len, (size_t)cap * sizeof *name_len); } \
                             name[ns] = (ps); name_len[ns] = (size_t)((pe) - (ps)); ++ns; } while (0)
    if (mode == ORACLE_FILE) {                                             /* :456-668 calculateInbreedingMmap */
        if (n == 0) { r->warnings = 1; ob_str(&o, "Sample\tInbreedingCoefficient\n"); res_take(r, &o); return 0; }
        int found_chrom = 0;
        size_t data_pos = n;
        /* the leading block of '#' and empty lines: every "#CHROM" line in it ADDS its columns 10.. to the sample list (:487-508) */
        for (;;) {
            size_t at = pos;
            if (!next_line(in, n, &pos, &ln)) { data_pos = n; break; }
            const char *s = ln.s, *e = ln.e;
            if (e > s && e[-1] == '\r') --e;
            if (s == e) continue;
            if (*s != '#') { data_pos = at; break; }
            if (e - s >= 6 && memcmp(s, "#CHROM", 6) == 0) {
                found_chrom = 1;
                const char *fs = s; int idx = 0;
                while (fs < e) {
                    const char *t = memchr(fs, '\t', (size_t)(e - fs)); if (!t) t = e;
                    if (idx >= 9) { const char *ne = t; if (ne > fs && ne[-1] == '\r') --ne; IB_PUSH(fs, ne); }
                    ++idx; fs = t + 1;
                }
            }
        }
        if (!found_chrom || ns == 0) { r->warnings = 2; ob_str(&o, "Sample\tInbreedingCoefficient\n"); res_take(r, &o); free(name); free(name_len); return 0; }
        if (!ib_alloc(&a, ns)) return -1;
        pos = data_pos;
        while (next_line(in, n, &pos, &ln)) {
            const char *s = ln.s, *e = ln.e;
            if (e > s && e[-1] == '\r') --e;
            if (s == e || *s == '#') continue;
            const char *fs, *fe;
            if (!field_at(s, e, 4, &fs, &fe) || fs == fe || memchr(fs, ',', (size_t)(fe - fs))) continue;     /* :549-553 */
            const char *sp, *spe;
            if (!field_at(s, e, 9, &sp, &spe)) continue;                   /* :557-561 */
            r->data_lines++;
            int alt_sum = 0, n_good = 0;
            /* a column the line does not have keeps the code of the last line that had it (the buffer is reused, :574-588) */
            for (int k = 0; k < ns && sp < e; ++k) {
                const char *t = memchr(sp, '\t', (size_t)(e - sp)); if (!t) t = e;
                int code = ib_code(sp, t);
                a.code[k] = code;
                if (code >= 0) { alt_sum += code; n_good++; }
                sp = t + 1;
            }
            if (n_good < 2) continue;
            a.variants++;
            ib_site(&a, alt_sum, n_good, global, skipb, countb);
        }
        ob_str(&o, "Sample\tInbreedingCoefficient\n");
        ib_report(&o, &a, name, name_len, quiet, r);
    } else {                                                               /* :670-826 calculateInbreedingStdin */
        int found_chrom = 0;
        while (next_line(in, n, &pos, &ln)) {
            const char *s = ln.s, *e = ln.e;
            if (s == e) continue;
            if (e[-1] == '\r') --e;
            if (s == e) continue;
            if (*s == '#') {
                if (!found_chrom) {
                    int hit = 0;
                    for (const char *q = s; q + 6 <= e; ++q) if (memcmp(q, "#CHROM", 6) == 0) { hit = 1; break; }
                    if (hit) {
                        found_chrom = 1;
                        const char *fs = s; int idx = 0;
                        for (;;) {                                         /* split_tabs: a final empty field counts */
                            const char *t = memchr(fs, '\t', (size_t)(e - fs));
                            const char *fe2 = t ? t : e;
                            if (idx >= 9) IB_PUSH(fs, fe2);
                            ++idx;
                            if (!t) break;
                            fs = t + 1;
                        }
                        if (!ib_alloc(&a, ns)) return -1;
                    }
                }
                continue;
            }
            if (!found_chrom) continue;
            int tabs = 0; for (const char *q = s; q < e; ++q) tabs += (*q == '\t');
            if (tabs + 1 < 10) continue;
            const char *fs, *fe;
            fs = nr_skip(s, e, 4); fe = fs; while (fe < e && *fe != '\t') ++fe;
            if (memchr(fs, ',', (size_t)(fe - fs))) continue;
            r->data_lines++;
            int alt_sum = 0, n_good = 0;
            const char *sp = nr_skip(s, e, 9);
            int gone = 0;
            for (int k = 0; k < ns; ++k) {
                int code = -1;
                if (!gone) {
                    const char *t = sp; while (t < e && *t != '\t') ++t;
                    code = ib_code(sp, t);
                    if (t >= e) gone = 1; else sp = t + 1;
                }
                a.code[k] = code;
                if (code >= 0) { alt_sum += code; n_good++; }
            }
            if (n_good < 2) continue;
            a.variants++;
            ib_site(&a, alt_sum, n_good, global, skipb, countb);
        }
        ob_str(&o, "Sample\tInbreedingCoefficient\n");
        if (!found_chrom) r->warnings = 4;
        else if (ns == 0) r->warnings = 5;
        else ib_report(&o, &a, name, name_len, quiet, r);
    }
#undef IB_PUSH
    r->rows = a.variants;
    ib_free(&a); free(name); free(name_len);
    res_take(r, &o);
    return 0;
}

/* ------------------------------------------------------------------ genotype_query (§8 f2): only lines in which at least one sample has the
 * queried genotype pass; '#' lines pass (stdin mode holds them back until the next data line comes), empty lines vanish;
 * no '\r' is cut anywhere. */

/* VCFX_genotype_query.cpp:246-272 parseDiploidAlleles; *a1 / *a2 keep whatever was assigned before a failure (the query is
 * parsed with them preset to -1 and used even when the parse fails, :640-644) */
static int gq_parse(const char *g, size_t n, int *a1, int *a2) {
    size_t sp = 0;
    while (sp < n && g[sp] != '|' && g[sp] != '/') ++sp;
    if (sp == n || sp == 0 || sp == n - 1) return 0;
    if (sp == 1 && g[0] == '.') return 0;
    *a1 = 0;
    for (size_t i = 0; i < sp; ++i) { if (g[i] < '0' || g[i] > '9') return 0; *a1 = (int)((unsigned)*a1 * 10u + (unsigned)(g[i] - '0')); }
    if (n - sp - 1 == 1 && g[sp + 1] == '.') return 0;
    *a2 = 0;
    for (size_t i = sp + 1; i < n; ++i) { if (g[i] < '0' || g[i] > '9') return 0; *a2 = (int)((unsigned)*a2 * 10u + (unsigned)(g[i] - '0')); }
    return 1;
}

/* :275-317 genotypeMatchesFast (qa <= qb: the caller sorted them) */
static int gq_match(const char *g, size_t n, const char *q, size_t qn, int qa, int qb, int strict) {
    if (strict) return n == qn && memcmp(g, q, n) == 0;
    if (n == 3 && qn == 3) {
        if (g[1] != '|' && g[1] != '/') return 0;
        if (g[0] < '0' || g[0] > '9' || g[2] < '0' || g[2] > '9') return 0;
        int ga = g[0] - '0', gb = g[2] - '0';
        if (ga > gb) { int t = ga; ga = gb; gb = t; }
        return ga == qa && gb == qb;
    }
    int a1, a2;
    if (!gq_parse(g, n, &a1, &a2)) return 0;
    if (a1 > a2) { int t = a1; a1 = a2; a2 = t; }
    return a1 == qa && a2 == qb;
}

/* :322-344 checkAnySampleMatches */
static int gq_any(const char *s, const char *e, int gi, const char *q, size_t qn, int qa, int qb, int strict) {
    const char *p = s;
    for (int i = 0; i < 9 && p < e; ++i) { const char *t = memchr(p, '\t', (size_t)(e - p)); if (!t) return 0; p = t + 1; }
    while (p < e) {
        const char *se = memchr(p, '\t', (size_t)(e - p)); if (!se) se = e;
        const char *gs, *ge;
        nr_nth_piece(p, se, gi, &gs, &ge);
        if (ge > gs && gq_match(gs, (size_t)(ge - gs), q, qn, qa, qb, strict)) return 1;
        p = se + 1;
    }
    return 0;
}

/* r: stdout; err: what goes to stderr without -q */
int oracle_genotype_query(const char *in, size_t n, int mode, const char *query, int strict, oracle_result *r, oracle_result *err) {
    res_init(r); res_init(err);
    obuf o = {0}, eb = {0};
    const size_t qn = strlen(query);
    int qa = -1, qb = -1;
    if (!strict) { gq_parse(query, qn, &qa, &qb); if (qa > qb) { int t = qa; qa = qb; qb = t; } }
    size_t pos = 0; line_t ln; int found = 0;
    /* stdin mode: '#' lines wait for the next data line (:559-571) */
    const char **hs = NULL; size_t *hl = NULL; size_t nh = 0, caph = 0;
    while (next_line(in, n, &pos, &ln)) {
        const char *s = ln.s, *e = ln.e;
        if (s == e) continue;
        if (*s == '#') {
            if (e - s >= 6 && memcmp(s, "#CHROM", 6) == 0) found = 1;
            if (mode == ORACLE_FILE) { ob_put(&o, s, (size_t)(e - s)); ob_ch(&o, '\n'); }
            else {
                if (nh == caph) { caph = caph ? caph * 2 : 64; hs = realloc(hs, caph * sizeof *hs); hl = realloc(hl, caph * sizeof *hl); }
                hs[nh] = s; hl[nh] = (size_t)(e - s); ++nh;
            }
            continue;
        }
        if (!found) { ob_str(&eb, "Error: No #CHROM header found before data lines.\n"); goto done; }     /* the run ends here */
        for (size_t i = 0; i < nh; ++i) { ob_put(&o, hs[i], hl[i]); ob_ch(&o, '\n'); }
        nh = 0;
        r->data_lines++;
        const char *f = s; int short_line = 0;
        for (int i = 0; i < 8 && f < e; ++i) { const char *t = memchr(f, '\t', (size_t)(e - f)); if (!t) { short_line = 1; break; } f = t + 1; }
        if (short_line) {
            r->warnings++;
            ob_str(&eb, "Warning: skipping line with <9 fields");
            if (mode != ORACLE_FILE) { ob_str(&eb, ": "); ob_put(&eb, s, (size_t)(e - s)); }
            ob_ch(&eb, '\n');
            continue;
        }
        const char *fe = memchr(f, '\t', (size_t)(e - f)); if (!fe) fe = e;
        int gi = oracle_gt_index(f, (size_t)(fe - f));
        if (gi < 0) continue;
        if (gq_any(s, e, gi, query, qn, qa, qb, strict)) { ob_put(&o, s, (size_t)(e - s)); ob_ch(&o, '\n'); r->rows++; }
    }
    if (mode != ORACLE_FILE && !found) ob_str(&eb, "Error: No #CHROM line found in VCF.\n");
done:
    free(hs); free(hl);
    res_take(r, &o); res_take(err, &eb);
    return 0;
}

/* ------------------------------------------------------------------ dosage_calculator (§8 f2): per data line "CHROM..ALT \t d,d,NA,.." with d =
 * the number of alleles > 0 of a two-allele GT */

/* VCFX_dosage_calculator.cpp:111-154 parseDosageInline */
static int ds_dosage(const char *g, size_t n) {
    if (n == 0) return -1;
    int dosage = 0, count = 0; size_t pos = 0;
    while (pos < n) {
        while (pos < n && (g[pos] == '/' || g[pos] == '|')) ++pos;
        if (pos >= n) break;
        if (g[pos] == '.') return -1;
        unsigned allele = 0; int digit = 0;
        while (pos < n && g[pos] >= '0' && g[pos] <= '9') { allele = allele * 10u + (unsigned)(g[pos] - '0'); digit = 1; ++pos; }
        if (!digit) return -1;
        if ((int)allele > 0) ++dosage;
        if (++count > 2) return -1;
    }
    return count == 2 ? dosage : -1;
}

/* r->rc = exit code; r->warnings = lines with fewer than ten fields (one message each; file mode: not with -q);
 * r->first_bad_line = 1 when the run ended at a data line in front of the "#CHROM" line (message, nothing on stdout) */
int oracle_dosage(const char *in, size_t n, int mode, oracle_result *r) {
    res_init(r);
    obuf o = {0};
    if (mode == ORACLE_FILE && n == 0) { res_take(r, &o); return 0; }      /* :388-391 */
    ob_str(&o, "CHROM\tPOS\tID\tREF\tALT\tDosages\n");
    size_t pos = 0; line_t ln; int header = 0;
    while (next_line(in, n, &pos, &ln)) {
        const char *s = ln.s, *e = ln.e;
        if (mode == ORACLE_FILE && e > s && e[-1] == '\r') --e;            /* :433-435 (stdin: getline keeps it) */
        if (s == e) continue;
        if (*s == '#') { if (e - s >= 6 && memcmp(s, "#CHROM", 6) == 0) header = 1; continue; }
        if (!header) {                                                     /* :452-458 / :236-240: the buffered output is thrown away */
            o.n = 0; r->first_bad_line = 1; r->rc = (mode == ORACLE_FILE) ? 1 : 0;
            res_take(r, &o); return 0;
        }
        r->data_lines++;
        const char *f[10]; size_t fl[10]; int nf = 0; size_t fs = 0, len = (size_t)(e - s);
        for (size_t i = 0; i <= len && nf < 10; ++i)
            if (i == len || s[i] == '\t') { f[nf] = s + fs; fl[nf] = i - fs; ++nf; fs = i + 1; }
        if (nf < 10) { r->warnings++; continue; }
        for (int k = 0; k < 5; ++k) { ob_put(&o, f[k], fl[k]); ob_ch(&o, '\t'); }
        int gi = oracle_gt_index(f[8], fl[8]);
        r->rows++;
        if (gi < 0) { ob_str(&o, "NA\n"); continue; }
        const char *sp = f[9]; int first = 1;
        while (sp < e) {
            const char *se = sp; while (se < e && *se != '\t') ++se;
            if (!first) ob_ch(&o, ',');
            first = 0;
            const char *gs, *ge; nr_nth_piece(sp, se, gi, &gs, &ge);
            int d = (ge > gs) ? ds_dosage(gs, (size_t)(ge - gs)) : -1;
            if (d < 0) ob_str(&o, "NA"); else ob_ch(&o, (char)('0' + d));
            sp = (se < e) ? se + 1 : e;
        }
        ob_ch(&o, '\n');
    }
    res_take(r, &o);
    return 0;
}

/* ------------------------------------------------------------------ variant_counter */

/* variant_counter.cpp:31-44 — at least 7 tabs. */
static int vc_eight_cols(const char *s, const char *e) {
    const char *p = s;
    if (s == e) return 0;
    for (int i = 0; i < 7; ++i) {
        p = (const char *)memchr(p, '\t', (size_t)(e - p));
        if (!p) return 0;
        ++p;
    }
    return 1;
}

int oracle_variant_count(const char *in, size_t n, int mode, int strict, oracle_result *r) {
    res_init(r);
    obuf o = {0}; size_t pos = 0; line_t ln; long long line_no = 0; int count = 0;
    while (next_line(in, n, &pos, &ln)) {
        const char *s = ln.s, *e = ln.e;
        ++line_no;                                           /* :360 / :216 */
        if (s == e || *s == '#') continue;                   /* :364 / :183-186 */
        if (mode == ORACLE_FILE && e[-1] == '\r') --e;       /* :366-368 (mmap only) */
        if (vc_eight_cols(s, e)) { ++count; continue; }
        if (strict) {                                        /* :373-377 / :195-197 */
            r->first_bad_line = line_no; r->rc = 1;
            res_take(r, &o);
            return 0;                                        /* nothing on stdout (:175-177) */
        }
        r->warnings++;
    }
    char nb[64]; int l = sprintf(nb, "Total Variants: %d\n", count);   /* :178 */
    ob_put(&o, nb, (size_t)l);
    r->rows = count;
    res_take(r, &o);
    return 0;
}

/* ------------------------------------------------------------------ allele_counter */

/* allele_counter.cpp:267-294 — on the first ':' piece: skip separators, '.' consumes one
 * byte, a digit run is one allele (0 = ref, else alt).  The reference never advances on
 * any other byte (infinite loop); outside [0-9./|] this restatement skips the byte, which
 * is outside the parity domain. */
void oracle_ac_counts(const char *g, size_t n, int *ref, int *alt) {
    size_t i = 0; *ref = 0; *alt = 0;
    while (i < n) {
        while (i < n && (g[i] == '/' || g[i] == '|')) ++i;
        if (i >= n) break;
        if (g[i] == '.') { ++i; continue; }
        unsigned a = 0; int has = 0;                /* unsigned: digit-run overflow is UB upstream */
        while (i < n && g[i] >= '0' && g[i] <= '9') { a = a * 10u + (unsigned)(g[i] - '0'); has = 1; ++i; }
        if (has) { if ((int)a == 0) ++*ref; else ++*alt; }
        else ++i;
    }
}

typedef struct { const char *s; size_t n; } sv;

static void put_int(obuf *o, long long v) { char nb[32]; int l = sprintf(nb, "%lld", v); ob_put(o, nb, (size_t)l); }

/* sample names = fields 9.. of a "#CHROM" line (allele_counter.cpp:810-821, 1139-1151) */
static void ac_header_names(const char *s, const char *e, sv **names, size_t *nn, size_t *cap) {
    const char *p = md_skip(s, e, 9);
    while (p < e) {
        const char *t = (const char *)memchr(p, '\t', (size_t)(e - p));
        const char *q = t ? t : e;
        if (*nn == *cap) { *cap = *cap ? *cap * 2 : 64; *names = (sv *)realloc(*names, *cap * sizeof(sv)); }
        (*names)[*nn].s = p; (*names)[*nn].n = (size_t)(q - p); ++*nn;
        p = t ? t + 1 : e;
    }
}

/* the -s argument: split on ' ', drop empties, trim (allele_counter.cpp:355-373) */
static size_t ac_parse_samples(const char *arg, sv **req) {
    size_t n = 0, cap = 0; *req = NULL;
    if (!arg) return 0;
    const char *p = arg, *end = arg + strlen(arg);
    while (p < end) {
        const char *q = (const char *)memchr(p, ' ', (size_t)(end - p));
        const char *te = q ? q : end;
        if (te > p) {
            const char *a = p, *b = te;
            while (a < b && strchr(" \t\n\r", *a)) ++a;
            while (b > a && strchr(" \t\n\r", b[-1])) --b;
            if (n == cap) { cap = cap ? cap * 2 : 8; *req = (sv *)realloc(*req, cap * sizeof(sv)); }
            if (a < b) { (*req)[n].s = a; (*req)[n].n = (size_t)(b - a); }
            else       { (*req)[n].s = p; (*req)[n].n = (size_t)(te - p); }   /* all-blank token kept as is */
            ++n;
        }
        if (!q) break;
        p = q + 1;
    }
    return n;
}

/* name -> column index; with duplicates the LAST column wins (map overwrite, :845-847) */
static long ac_lookup(const sv *names, size_t nn, sv key) {
    long hit = -1;
    for (size_t i = 0; i < nn; ++i)
        if (names[i].n == key.n && memcmp(names[i].s, key.s, key.n) == 0) hit = (long)i;
    return hit;
}

static void ac_prefix(obuf *pre, const char **pp, const char *e) {
    /* CHROM..ALT each followed by '\t', absent fields empty (:578-601) */
    const char *p = *pp;
    pre->n = 0;
    for (int k = 0; k < 5; ++k) {
        const char *t = (const char *)memchr(p, '\t', (size_t)(e - p));
        const char *q = t ? t : e;
        ob_put(pre, p, (size_t)(q - p)); ob_ch(pre, '\t');
        p = q; if (p < e) ++p;
    }
    *pp = p;
}

int oracle_allele_counter(const char *in, size_t n, int path, int format, int limit,
                          const char *samples, oracle_result *r) {
    res_init(r);
    obuf o = {0}, pre = {0};
    sv *names = NULL; size_t nn = 0, ncap = 0;
    sv *req = NULL; size_t nreq = ac_parse_samples(samples, &req);
    size_t *sel = NULL; size_t nsel = 0;
    size_t pos = 0; line_t ln;
    int have_header = 0, failed_early = 0;
    static const char TEXT_HDR[] = "CHROM\tPOS\tID\tREF\tALT\tSample\tRef_Count\tAlt_Count\n";
    static const char AGG_HDR[]  = "CHROM\tPOS\tID\tREF\tALT\tTotal_Ref\tTotal_Alt\tSample_Count\n";

#define AC_FAIL() do { r->rc = 1; failed_early = 1; goto done; } while (0)

    if (path != ORACLE_AC_STREAM) {
        if (n == 0) AC_FAIL();                               /* :793-796 / :1273-1276 "Empty file" */
        /* header block: every leading '#' line; names accumulate over #CHROM lines */
        size_t data_pos = n; int found_data = 0;
        for (;;) {
            size_t save = pos;
            if (!next_line(in, n, &pos, &ln)) break;
            if (*ln.s == '#') {   /* next_line never yields past the buffer, so *ln.s is readable */
                if (ln.e - ln.s >= 6 && memcmp(ln.s, "#CHROM", 6) == 0) {
                    ac_header_names(ln.s, ln.e, &names, &nn, &ncap); have_header = 1;
                }
                continue;
            }
            data_pos = save; found_data = 1; break;          /* first non-'#' line (even an empty one) */
        }
        if (nn == 0) AC_FAIL();                              /* "No samples found in VCF" */
        if (path == ORACLE_AC_MT_TEXT && !found_data) AC_FAIL();   /* :836-839 "No data lines found" */
        (void)have_header;
        /* selection */
        if (nreq) {
            sel = (size_t *)malloc(nreq * sizeof(size_t));
            for (size_t i = 0; i < nreq; ++i) {
                long k = ac_lookup(names, nn, req[i]);
                if (k < 0) AC_FAIL();                        /* "Sample 'x' not found" */
                sel[nsel++] = (size_t)k;
            }
        } else {
            sel = (size_t *)malloc((nn ? nn : 1) * sizeof(size_t));
            for (size_t i = 0; i < nn; ++i) sel[nsel++] = i;
        }
        if (path == ORACLE_AC_UNIFIED && limit > 0 && nsel > (size_t)limit) nsel = (size_t)limit;  /* :1336 */

        if (path == ORACLE_AC_MT_TEXT || format == ORACLE_AC_TEXT) ob_str(&o, TEXT_HDR);
        else if (format == ORACLE_AC_AGGREGATE) ob_str(&o, AGG_HDR);
        else {                                               /* BinaryHeader, packed (:326-333, 1361-1366) */
            unsigned char h[20] = { 'V', 'C', 'A', 'C', 1, 0, 0, 0 };
            uint32_t ns = (uint32_t)nsel; memcpy(h + 8, &ns, 4); memset(h + 12, 0, 8);
            ob_put(&o, (const char *)h, 20);
        }

        pos = data_pos;
        const char **starts = NULL; size_t scap = 0;
        while (next_line(in, n, &pos, &ln)) {
            const char *s = ln.s, *e = ln.e;                 /* no '\r' handling */
            if (s >= e || *s == '#') continue;               /* :571 / :1373 */
            r->data_lines++;
            const char *p = s;
            ac_prefix(&pre, &p, e);
            p = md_skip(p, e, 4);                            /* QUAL FILTER INFO FORMAT (:604) */
            if (path == ORACLE_AC_MT_TEXT) {
                /* every sample column start, always at least one entry (:528-544) */
                size_t ns = 0;
                const char *q = p;
                for (;;) {
                    if (ns == scap) { scap = scap ? scap * 2 : 4096; starts = (const char **)realloc(starts, scap * sizeof(*starts)); }
                    starts[ns++] = q;
                    if (q >= e) break;
                    const char *t = (const char *)memchr(q, '\t', (size_t)(e - q));
                    if (!t) break;
                    q = t + 1;
                }
                for (size_t i = 0; i < nsel; ++i) {
                    int rc_ = 0, ac_ = 0;
                    size_t idx = sel[i];
                    if (idx < ns) {
                        const char *gs = starts[idx];
                        const char *ge = (idx + 1 < ns) ? starts[idx + 1] - 1 : e;
                        const char *c = (gs < ge) ? (const char *)memchr(gs, ':', (size_t)(ge - gs)) : NULL;
                        if (c) ge = c;
                        if (gs < ge) oracle_ac_counts(gs, (size_t)(ge - gs), &rc_, &ac_);
                    }
                    ob_put(&o, pre.p, pre.n);
                    ob_put(&o, names[sel[i]].s, names[sel[i]].n); ob_ch(&o, '\t');
                    put_int(&o, (int8_t)rc_); ob_ch(&o, '\t');      /* int8_t storage (:621-622) */
                    put_int(&o, (int8_t)ac_); ob_ch(&o, '\n');
                    r->rows++;
                }
            } else {
                /* forward-only walk; stops at the first absent column (:1416-1451) */
                long long tr = 0, ta = 0; int sc = 0;
                size_t col = 0; const char *sp = p;
                for (size_t i = 0; i < nsel; ++i) {
                    while (col < sel[i] && sp < e) {
                        const char *t = (const char *)memchr(sp, '\t', (size_t)(e - sp));
                        sp = t ? t + 1 : e; ++col;
                    }
                    if (sp >= e) break;
                    const char *t = (const char *)memchr(sp, '\t', (size_t)(e - sp));
                    const char *ge = t ? t : e;
                    const char *c = (const char *)memchr(sp, ':', (size_t)(ge - sp));
                    if (c) ge = c;
                    int rc_ = 0, ac_ = 0;
                    oracle_ac_counts(sp, (size_t)(ge - sp), &rc_, &ac_);
                    if (format == ORACLE_AC_TEXT) {
                        ob_put(&o, pre.p, pre.n);
                        ob_put(&o, names[sel[i]].s, names[sel[i]].n); ob_ch(&o, '\t');
                        put_int(&o, rc_); ob_ch(&o, '\t'); put_int(&o, ac_); ob_ch(&o, '\n');
                        r->rows++;
                    } else if (format == ORACLE_AC_AGGREGATE) {
                        tr += rc_; ta += ac_; ++sc;
                    } else {
                        char b[2] = { (char)(int8_t)rc_, (char)(int8_t)ac_ };
                        ob_put(&o, b, 2);
                    }
                }
                if (format == ORACLE_AC_AGGREGATE) {
                    ob_put(&o, pre.p, pre.n);
                    put_int(&o, tr); ob_ch(&o, '\t'); put_int(&o, ta); ob_ch(&o, '\t');
                    put_int(&o, sc); ob_ch(&o, '\n');
                    r->rows++;
                }
            }
        }
        free(starts);
    } else {
        /* stdin: header row first, names/selection rebuilt at every #CHROM line
         * (allele_counter.cpp:1131, 1138-1181); -a/-b/-l are ignored on this path */
        ob_str(&o, TEXT_HDR);
        size_t selcap = 0;
        while (next_line(in, n, &pos, &ln)) {
            const char *s = ln.s, *e = ln.e;
            if (s == e) continue;
            if (*s == '#') {
                if (e - s >= 6 && memcmp(s, "#CHROM", 6) == 0) {
                    ac_header_names(s, e, &names, &nn, &ncap);
                    size_t add = nreq ? nreq : nn;
                    if (nsel + add > selcap) { selcap = (nsel + add) * 2 + 8; sel = (size_t *)realloc(sel, selcap * sizeof(size_t)); }
                    if (nreq) {
                        for (size_t i = 0; i < nreq; ++i) {
                            long k = ac_lookup(names, nn, req[i]);
                            if (k < 0) AC_FAIL();
                            sel[nsel++] = (size_t)k;
                        }
                    } else {
                        for (size_t i = 0; i < nn; ++i) sel[nsel++] = i;
                    }
                    have_header = 1;
                }
                continue;
            }
            if (!have_header) AC_FAIL();                     /* :1183-1186; nothing was flushed yet */
            r->data_lines++;
            const char *p = s;
            ac_prefix(&pre, &p, e);
            p = md_skip(p, e, 4);
            size_t col = 0; const char *sp = p;
            for (size_t i = 0; i < nsel; ++i) {
                while (col < sel[i] && sp < e) {
                    const char *t = (const char *)memchr(sp, '\t', (size_t)(e - sp));
                    sp = t ? t + 1 : e; ++col;
                }
                if (sp >= e) break;
                const char *t = (const char *)memchr(sp, '\t', (size_t)(e - sp));
                const char *ge = t ? t : e;
                const char *c = (const char *)memchr(sp, ':', (size_t)(ge - sp));
                if (c) ge = c;
                int rc_ = 0, ac_ = 0;
                oracle_ac_counts(sp, (size_t)(ge - sp), &rc_, &ac_);
                ob_put(&o, pre.p, pre.n);
                ob_put(&o, names[sel[i]].s, names[sel[i]].n); ob_ch(&o, '\t');
                put_int(&o, rc_); ob_ch(&o, '\t'); put_int(&o, ac_); ob_ch(&o, '\n');
                r->rows++;
            }
        }
        if (!have_header) r->rc = 1;                         /* :1259 */
    }
done:
    /* Every early failure happens before the first write(2): the file paths validate the
     * header first, and the stream path keeps its header row in a buffer that is dropped on
     * error (allele_counter.cpp:1162, :1185).  A header-less stream that reaches EOF does
     * write that row and then returns rc 1 (:1255-1259). */
    if (failed_early) o.n = 0;
    free(pre.p); free(names); free(req); free(sel);
    res_take(r, &o);
    return 0;
#undef AC_FAIL
}
