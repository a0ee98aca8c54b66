/*
 * vcfx_synth.c — deterministic synthetic VCF generator for the BASELINE.json configs.
 *
 * There is no network and the 1000 Genomes files the reference benchmarks on
 * (benchmarks/benchmark_config.yaml:18-36) cannot be fetched, so every test and bench
 * input is synthesised here with the shapes SURVEY.md §8(d) lists:
 *
 *   shape 1  C1  biallelic SNPs, FORMAT GT, unphased  a/b
 *   shape 2  C2  1000G chr21 shape: FORMAT GT, phased a|b, CHROM 21
 *   shape 3  C3  C2 + 5 % missing ('.', './.', '.|.', './1', '0/.'), 30 % unphased,
 *                3 % haploid; one line in three is kept fully called
 *   shape 4  C4  FORMAT GT:AD:DP:GQ:PL, 1..4 ALT alleles (85/10/4/1 %), 1 % missing
 *
 * Each data line depends only on (seed, shape, n_samples, variant index), so any range of
 * variants can be produced independently: threads, ranks and the CPU-baseline sample all
 * see identical bytes for the same variant.  POS is 1 + 100*index + jitter so that it
 * increases without a running sum.
 *
 * Plain C + pthreads, C ABI, loaded with ctypes (vcfx_b200/synth.py).  Not on the product
 * path: it only manufactures inputs.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint64_t seed;
    uint32_t shape;       /* 1..4 */
    uint32_t n_samples;
} vcfx_synth_cfg;

static inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
typedef struct { uint64_t s; } rng_t;
static inline uint64_t rnext(rng_t *r) { r->s += 0x9E3779B97F4A7C15ULL; uint64_t z = r->s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }

static inline char *put_u(char *d, uint64_t v) {
    char t[24]; int n = 0;
    do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *d++ = t[--n];
    return d;
}
static inline char *put_s(char *d, const char *s) { size_t n = strlen(s); memcpy(d, s, n); return d + n; }

static const char *FORMAT_OF[5] = { "", "GT", "GT", "GT", "GT:AD:DP:GQ:PL" };

/* upper bound of one data line, used to size scratch and caller buffers */
static size_t line_bound(const vcfx_synth_cfg *c) {
    size_t per = (c->shape == 4) ? 112 : 4;
    return 256 + per * (size_t)c->n_samples;
}

size_t vcfx_synth_line_bound(const vcfx_synth_cfg *c) { return line_bound(c); }

size_t vcfx_synth_header(const vcfx_synth_cfg *c, char *dst, size_t cap) {
    size_t need = 1024 + 8 * (size_t)c->n_samples;
    if (!dst || cap < need) return need;
    char *d = dst;
    d = put_s(d, "##fileformat=VCFv4.1\n");
    d = put_s(d, "##source=vcfx_b200_synth\n");
    d = put_s(d, "##contig=<ID=21>\n");
    d = put_s(d, "##INFO=<ID=AC,Number=A,Type=Integer,Description=\"Alt allele count\">\n");
    d = put_s(d, "##INFO=<ID=AN,Number=1,Type=Integer,Description=\"Allele number\">\n");
    d = put_s(d, "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n");
    if (c->shape == 4) {
        d = put_s(d, "##FORMAT=<ID=AD,Number=R,Type=Integer,Description=\"Allelic depths\">\n");
        d = put_s(d, "##FORMAT=<ID=DP,Number=1,Type=Integer,Description=\"Depth\">\n");
        d = put_s(d, "##FORMAT=<ID=GQ,Number=1,Type=Integer,Description=\"Genotype quality\">\n");
        d = put_s(d, "##FORMAT=<ID=PL,Number=G,Type=Integer,Description=\"Phred likelihoods\">\n");
    }
    d = put_s(d, "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT");
    for (uint32_t i = 0; i < c->n_samples; ++i) {
        *d++ = '\t'; *d++ = 'H'; *d++ = 'G';
        uint32_t v = 96 + i; char t[5];
        for (int k = 4; k >= 0; --k) { t[k] = (char)('0' + v % 10); v /= 10; }
        memcpy(d, t, 5); d += 5;
    }
    *d++ = '\n';
    return (size_t)(d - dst);
}

/* one data line for variant `vi` into d (no bounds check: caller sized with line_bound) */
static char *gen_line(const vcfx_synth_cfg *c, uint64_t vi, char *d, char *scratch) {
    rng_t r; r.s = mix64(c->seed ^ mix64(vi * 0xD6E8FEB86659FD93ULL + c->shape));
    const uint32_t S = c->n_samples;
    static const char BASES[4] = { 'A', 'C', 'G', 'T' };
    uint64_t h = rnext(&r);
    /* alt allele frequency: skewed towards rare, in [0, 0.5] */
    double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
    double f = 0.5 * u * u * u;
    if ((h & 7) == 0) f = 0.5 * u;                      /* some common variants */
    uint32_t thr = (uint32_t)(f * 65536.0);
    uint32_t n_alt = 1;
    if (c->shape == 4) { uint32_t k = (uint32_t)(rnext(&r) % 100); n_alt = k < 85 ? 1 : k < 95 ? 2 : k < 99 ? 3 : 4; }
    int fully_called = (c->shape != 3) || (vi % 3 == 0);

    /* ---- sample columns into scratch, counting alleles for INFO */
    char *s = scratch; uint32_t ac = 0, an = 0;
    if (c->shape <= 2) {
        const char sep = (c->shape == 1) ? '/' : '|';
        uint32_t i = 0;
        while (i < S) {
            uint64_t x = rnext(&r);
            for (int k = 0; k < 2 && i < S; ++k, ++i) {
                uint32_t a = ((uint32_t)(x & 0xFFFF) < thr); x >>= 16;
                uint32_t b = ((uint32_t)(x & 0xFFFF) < thr); x >>= 16;
                s[0] = (char)('0' + a); s[1] = sep; s[2] = (char)('0' + b); s[3] = '\t'; s += 4;
                ac += a + b;
            }
        }
        an = 2 * S;
    } else if (c->shape == 3) {
        for (uint32_t i = 0; i < S; ++i) {
            uint64_t x = rnext(&r);
            uint32_t a = ((uint32_t)(x & 0xFFFF) < thr), b = ((uint32_t)((x >> 16) & 0xFFFF) < thr);
            uint32_t sel = (uint32_t)((x >> 32) % 1000);
            char sep = (((x >> 44) % 100) < 30) ? '/' : '|';
            if (!fully_called && sel < 50) {
                switch (sel % 5) {
                case 0: *s++ = '.'; break;
                case 1: *s++ = '.'; *s++ = '/'; *s++ = '.'; break;
                case 2: *s++ = '.'; *s++ = '|'; *s++ = '.'; break;
                case 3: *s++ = '.'; *s++ = '/'; *s++ = '1'; ac++; an++; break;
                default: *s++ = '0'; *s++ = '/'; *s++ = '.'; an++; break;
                }
            } else if (sel >= 50 && sel < 80) {
                *s++ = (char)('0' + a); ac += a; an++;
            } else {
                *s++ = (char)('0' + a); *s++ = sep; *s++ = (char)('0' + b); ac += a + b; an += 2;
            }
            *s++ = '\t';
        }
    } else {
        const uint32_t n_pl = (n_alt + 1) * (n_alt + 2) / 2;
        for (uint32_t i = 0; i < S; ++i) {
            uint64_t x = rnext(&r), y = rnext(&r);
            if ((x >> 50) % 100 == 0) {                  /* 1 % missing */
                s = put_s(s, "./.:.:.:.:.\t");
                continue;
            }
            uint32_t a = ((uint32_t)(x & 0xFFFF) < thr) ? 1 + (uint32_t)((x >> 40) % n_alt) : 0;
            uint32_t b = ((uint32_t)((x >> 16) & 0xFFFF) < thr) ? 1 + (uint32_t)((x >> 44) % n_alt) : 0;
            *s++ = (char)('0' + a); *s++ = (((x >> 32) & 15) == 0) ? '|' : '/'; *s++ = (char)('0' + b);
            ac += (a != 0) + (b != 0); an += 2;
            *s++ = ':';
            uint32_t dp = 0;
            for (uint32_t k = 0; k <= n_alt; ++k) {
                uint32_t v = (uint32_t)((y >> (6 * k)) & 63);
                if (k != a && k != b) v &= 3;
                if (k) *s++ = ',';
                s = put_u(s, v); dp += v;
            }
            *s++ = ':'; s = put_u(s, dp);
            *s++ = ':'; s = put_u(s, (y >> 36) % 100);
            *s++ = ':';
            uint64_t z = mix64(y);
            for (uint32_t k = 0; k < n_pl; ++k) {
                if (k) *s++ = ',';
                if ((k & 7) == 7) z = mix64(z);
                s = put_u(s, ((z >> (8 * (k & 7))) & 255) * ((k * 7 + a + b) % 3 != 0));
            }
            *s++ = '\t';
        }
    }
    size_t slen = (size_t)(s - scratch);
    if (slen) --slen;                                    /* drop the trailing tab */

    /* ---- fixed fields */
    uint64_t pos = 1 + vi * 100 + (rnext(&r) % 100);
    if (c->shape == 1) { *d++ = '1'; } else { *d++ = '2'; *d++ = '1'; }
    *d++ = '\t'; d = put_u(d, pos); *d++ = '\t';
    uint64_t g = rnext(&r);
    if (g & 1) { *d++ = 'r'; *d++ = 's'; d = put_u(d, 1000000 + vi); } else { *d++ = '.'; }
    *d++ = '\t';
    uint32_t rb = (uint32_t)((g >> 8) & 3);
    *d++ = BASES[rb]; *d++ = '\t';
    for (uint32_t k = 0; k < n_alt; ++k) { if (k) *d++ = ','; *d++ = BASES[(rb + 1 + k) & 3]; if (k >= 3) *d++ = 'T'; }
    *d++ = '\t';
    d = put_s(d, "100\tPASS\tAC="); d = put_u(d, ac); d = put_s(d, ";AN="); d = put_u(d, an);
    *d++ = '\t'; d = put_s(d, FORMAT_OF[c->shape]);
    if (S) { *d++ = '\t'; memcpy(d, scratch, slen); d += slen; }
    *d++ = '\n';
    return d;
}

typedef struct {
    const vcfx_synth_cfg *cfg; uint64_t v0, v1; char *buf; size_t len; size_t cap;
} job_t;

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    size_t lb = line_bound(j->cfg);
    char *scratch = (char *)malloc(lb);
    char *d = j->buf;
    for (uint64_t v = j->v0; v < j->v1; ++v) d = gen_line(j->cfg, v, d, scratch);
    j->len = (size_t)(d - j->buf);
    free(scratch);
    return NULL;
}

/*
 * Generate variants [first, first+count) back to back into dst.
 * Returns the number of bytes written, or the capacity needed when dst is NULL / too small
 * (a safe upper bound: count * line_bound).
 */
size_t vcfx_synth_lines(const vcfx_synth_cfg *c, uint64_t first, uint64_t count,
                        char *dst, size_t cap, int threads) {
    size_t lb = line_bound(c);
    size_t need = (size_t)count * lb;
    if (count == 0) return 0;
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > count) threads = (int)count;
    if (threads == 1) {
        if (!dst || cap < need) return need;
        job_t j = { c, first, first + count, dst, 0, cap };
        worker(&j);
        return j.len;
    }
    /* threads write into private buffers, then the pieces are packed in order */
    job_t *jobs = (job_t *)calloc((size_t)threads, sizeof(job_t));
    pthread_t *tid = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    uint64_t per = (count + (uint64_t)threads - 1) / (uint64_t)threads;
    size_t total = 0; int nt = 0;
    for (int t = 0; t < threads; ++t) {
        uint64_t a = first + per * (uint64_t)t, b = a + per;
        if (a >= first + count) break;
        if (b > first + count) b = first + count;
        jobs[t].cfg = c; jobs[t].v0 = a; jobs[t].v1 = b;
        jobs[t].cap = (size_t)(b - a) * lb;
        jobs[t].buf = (char *)malloc(jobs[t].cap);
        pthread_create(&tid[t], NULL, worker, &jobs[t]);
        ++nt;
    }
    for (int t = 0; t < nt; ++t) { pthread_join(tid[t], NULL); total += jobs[t].len; }
    size_t ret = total;
    if (!dst || cap < total) ret = total > need ? total : need;
    else { char *d = dst; for (int t = 0; t < nt; ++t) { memcpy(d, jobs[t].buf, jobs[t].len); d += jobs[t].len; } }
    for (int t = 0; t < nt; ++t) free(jobs[t].buf);
    free(jobs); free(tid);
    return ret;
}
