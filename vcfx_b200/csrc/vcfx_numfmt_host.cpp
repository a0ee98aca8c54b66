// vcfx_numfmt_host.cpp — host build of vcfx_numfmt.cuh for the "not gpu" tests only
// (tests/test_numfmt.py).  Compiled with g++ -ffp-contract=off into _numfmt_host.so.
// The product never loads this; the device build of the same header is what ships.
#include "vcfx_numfmt.cuh"

extern "C" {
int vcfx_host_fmt_af_file(unsigned alt, unsigned total, char *d) { return vcfx::fmt_af_file(vcfx::af_value(alt, total), d); }
int vcfx_host_fmt_af_stdin(unsigned alt, unsigned total, char *d) { return vcfx::fmt_af_stdin(vcfx::af_value(alt, total), d); }
int vcfx_host_fmt_fixed(double v, int digits, char *d) { return vcfx::fmt_fixed_exact(v, digits, d); }
int vcfx_host_fmt_p_file(double v, char *d) { return vcfx::fmt_p_file(v, d); }
int vcfx_host_fmt_p_stdin(double v, char *d) { return vcfx::fmt_p_stdin(v, d); }
double vcfx_host_hwe_pvalue(int hr, int het, int ha) { return vcfx::hwe_pvalue(hr, het, ha); }
}
