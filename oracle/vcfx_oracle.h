/*
 * vcfx_oracle.h — CPU restatement of the VCFX hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity checker for vcfx_b200's CUDA path.  It is NOT part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  libvcfx_cuda never links or calls it and has no CPU fallback.
 *
 * Every function restates, in plain C, what one reference tool does to a whole VCF held in
 * memory; the reference file:line each rule comes from is cited in vcfx_oracle.c.
 * Pinning: tests/test_oracle_golden.py checks it byte-for-byte against the compiled
 * reference tools (oracle/_ref/VCFX_*, built by oracle/Makefile from /root/reference) on the
 * reference test-suite's cases and on seeded fuzz inputs, and tests/golden/ holds outputs
 * of those reference binaries.
 */
#ifndef VCFX_ORACLE_H
#define VCFX_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* semantics selector: the reference formats numbers and skips lines differently when it
 * mmaps a file (-i FILE / positional) and when it reads stdin (SURVEY.md finding 1). */
enum { ORACLE_FILE = 0, ORACLE_STDIN = 1 };

/* allele_counter code paths (allele_counter.cpp:1522-1533) */
enum { ORACLE_AC_MT_TEXT = 0,   /* -i FILE, default TEXT: countAllelesMmapMT + processChunk */
       ORACLE_AC_STREAM  = 1,   /* stdin: countAllelesStream */
       ORACLE_AC_UNIFIED = 2 }; /* -i FILE with -a / -b / -l / -z: countAllelesUnified */
enum { ORACLE_AC_TEXT = 0, ORACLE_AC_AGGREGATE = 1, ORACLE_AC_BINARY = 2 };

typedef struct {
    char     *out;            /* stdout bytes, malloc'd; release with oracle_free */
    size_t    out_len;
    int       rc;             /* process exit code the tool would return */
    long long data_lines;     /* tool specific: data lines seen */
    long long rows;           /* rows / variants emitted or counted */
    long long flagged;        /* missing_detector: lines rewritten */
    long long warnings;       /* lines that would print a warning on stderr */
    long long first_bad_line; /* variant_counter --strict: 1-based failing line, else 0 */
} oracle_result;

void oracle_free(oracle_result *r);

int oracle_allele_freq(const char *in, size_t n, int mode, oracle_result *r);
int oracle_hwe(const char *in, size_t n, int mode, oracle_result *r);
int oracle_missing(const char *in, size_t n, int mode, oracle_result *r);
int oracle_phase_checker(const char *in, size_t n, int mode, oracle_result *r);    /* VCFX_phase_checker (§8 f2) */
int oracle_phase_checker_err(const char *in, size_t n, int mode, oracle_result *r, oracle_result *err);   /* + stderr text without -q */
int oracle_indexer(const char *in, size_t n, int mode, oracle_result *r);          /* VCFX_indexer (§8 f4) */
/* VCFX_dosage_calculator (§8 f2); r->warnings = lines with fewer than ten fields, r->first_bad_line = 1: ended at a data line before the header */
int oracle_dosage(const char *in, size_t n, int mode, oracle_result *r);
/* VCFX_genotype_query (§8 f2): r = stdout, err = the stderr text of a run without -q */
int oracle_genotype_query(const char *in, size_t n, int mode, const char *query, int strict, oracle_result *r, oracle_result *err);
/* VCFX_inbreeding_calculator (§8 f3); flags: 1 --freq-mode global, 2 --skip-boundary, 4 --count-boundary-as-used, 8 -q */
int oracle_inbreeding(const char *in, size_t n, int mode, int flags, oracle_result *r);
int oracle_nonref_filter(const char *in, size_t n, int mode, oracle_result *r);   /* VCFX_nonref_filter (§8 f2) */
int oracle_variant_count(const char *in, size_t n, int mode, int strict, oracle_result *r);
/* samples: NULL/"" = all; else the -s argument (space separated names) */
int oracle_allele_counter(const char *in, size_t n, int path, int format, int limit_samples,
                          const char *samples, oracle_result *r);

/* scalar pieces, exported so kernel unit tests can be compared value by value */
int    oracle_fmt_af_file(double v, char *dst);   /* writeDouble4 twin; returns length */
int    oracle_fmt_af_stdin(double v, char *dst);  /* "%.4f" */
int    oracle_fmt_p_file(double v, char *dst);    /* appendDouble twin (truncating) */
int    oracle_fmt_p_stdin(double v, char *dst);   /* "%.6f" */
double oracle_hwe_pvalue(int hom_ref, int het, int hom_alt);
void   oracle_hwe_pvalues(const int *counts, size_t n, double *out);   /* the same over n triples */
void   oracle_p_text_diffs(const double *a, const double *b, size_t n, size_t *file_diffs, size_t *stdin_diffs);
void   oracle_af_counts(const char *gt, size_t n, int *alt, int *total);
int    oracle_hwe_class(const char *sample, size_t n);
void   oracle_ac_counts(const char *gt, size_t n, int *ref, int *alt);
int    oracle_gt_index(const char *format, size_t n);

#ifdef __cplusplus
}
#endif
#endif
