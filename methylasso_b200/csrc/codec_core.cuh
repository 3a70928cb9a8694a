// codec_core.cuh — parsing one line of a methylation count file (device logic, shared with the CPU emulation).
//
// The reference's CLI reads its inputs with data.table::fread(file, select = c(1, 2, a, b)) (MethyLasso.R:246-262):
// column 1 = chromosome, column 2 = position, and two numeric columns (coverage and methylation level, or methylated
// and unmethylated counts).  Here one thread parses one line.  Numbers are converted exactly as strtod / fread do
// whenever one correctly rounded operation suffices (Clinger's fast path: at most 15 significant digits and a decimal
// exponent within +-22, which covers counts, positions and percentages as Bismark / BED files print them); anything
// else is flagged and left to the host (status 1), never approximated.
#pragma once
#include "hd.h"

namespace ml {

enum : int32_t { TXT_OK = 0, TXT_HOST_PARSE = 1, TXT_MALFORMED = 2 };

struct TxtLine {
  int64_t pos;
  double a, b;
  int32_t chr_off, chr_len;   // chromosome field: offset from the line start, length
  int32_t status;
};

ML_HD bool txt_is_sep(char c, char sep) { return sep ? c == sep : (c == '\t' || c == ' '); }

ML_HD double txt_pow10(int k) {   // exact powers of ten in double: 10^0 .. 10^22
  const double t[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16,
                        1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  return t[k];
}

// [p, e) -> double.  TXT_OK: out is the correctly rounded value; TXT_HOST_PARSE: a number, but not convertible in one
// rounded operation; TXT_MALFORMED: not a decimal number (empty, NA, nan, inf, stray characters).
ML_HD int txt_parse_double(const char* p, const char* e, double& out) {
  out = 0.0;
  if (p >= e) return TXT_MALFORMED;
  bool neg = false;
  if (*p == '+' || *p == '-') { neg = *p == '-'; ++p; }
  uint64_t m = 0;
  int digits = 0, sig = 0, exp10 = 0;
  bool any = false, point = false;
  for (; p < e; ++p) {
    const char c = *p;
    if (c >= '0' && c <= '9') {
      any = true;
      if (sig > 0 || c != '0') ++sig;             // leading zeros are not significant
      if (sig <= 19) { m = m * 10 + (uint64_t)(c - '0'); if (point) --exp10; }
      else if (!point) ++exp10;                   // digits beyond 19 only shift the exponent (host parse anyway)
      ++digits;
    } else if (c == '.' && !point) {
      point = true;
    } else {
      break;
    }
  }
  if (!any) return TXT_MALFORMED;
  if (p < e && (*p == 'e' || *p == 'E')) {
    ++p;
    bool eneg = false;
    if (p < e && (*p == '+' || *p == '-')) { eneg = *p == '-'; ++p; }
    if (p >= e) return TXT_MALFORMED;
    int ex = 0;
    for (; p < e && *p >= '0' && *p <= '9'; ++p) ex = ex < 10000 ? ex * 10 + (*p - '0') : ex;
    exp10 += eneg ? -ex : ex;
  }
  if (p != e) return TXT_MALFORMED;
  if (m == 0) { out = neg ? -0.0 : 0.0; return TXT_OK; }
  if (sig > 15 || exp10 > 22 || exp10 < -22) return TXT_HOST_PARSE;
  const double v = exp10 >= 0 ? (double)m * txt_pow10(exp10) : (double)m / txt_pow10(-exp10);
  out = neg ? -v : v;
  return TXT_OK;
}

// one line [p, e) (no newline; a trailing '\r' is dropped): fields 0 (chromosome), 1 (position), col_a, col_b (0-based)
ML_HD TxtLine txt_parse_line(const char* p, const char* e, int col_a, int col_b, char sep) {
  TxtLine L{0, 0.0, 0.0, 0, 0, TXT_OK};
  const char* const line = p;
  if (e > p && e[-1] == '\r') --e;
  int field = 0, got = 0, status = TXT_OK;
  while (p <= e) {
    const char* f0 = p;
    while (p < e && !txt_is_sep(*p, sep)) ++p;
    if (field == 0) {
      L.chr_off = (int32_t)(f0 - line);
      L.chr_len = (int32_t)(p - f0);
      if (L.chr_len == 0) status = TXT_MALFORMED;
      ++got;
    } else if (field == 1 || field == col_a || field == col_b) {
      double v;
      int st = txt_parse_double(f0, p, v);
      if (field == 1) {
        if (st == TXT_OK && (v < 0 || v > 2147483647.0 || v != (double)(int64_t)v)) st = TXT_MALFORMED;
        L.pos = (int64_t)v;
      }
      if (field == col_a) L.a = v;
      if (field == col_b) L.b = v;       // col_a == col_b is allowed (the same column twice)
      status = st > status ? st : status;
      got += (field == 1) + (field == col_a) + (field == col_b);
    }
    ++field;
    if (p >= e) break;
    ++p;                                  // the separator
    if (!sep) while (p < e && txt_is_sep(*p, sep)) ++p;   // runs of blanks count as one separator
  }
  if (got < 4) status = TXT_MALFORMED;    // a selected column is missing
  L.status = status;
  return L;
}

}  // namespace ml
