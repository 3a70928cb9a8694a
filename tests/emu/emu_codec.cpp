// CPU build of methylasso_b200/csrc/codec_core.cuh for tests/test_codec.py.  Test infrastructure only.
#include <cstdint>
#include "../../methylasso_b200/csrc/codec_core.cuh"
extern "C" {
int emu_parse_double(const char* s, int n, double* out) { return ml::txt_parse_double(s, s + n, *out); }
int emu_parse_line(const char* s, int n, int col_a, int col_b, int sep, long long* pos, double* a, double* b, int* chr_off,
                   int* chr_len) {
  const ml::TxtLine L = ml::txt_parse_line(s, s + n, col_a, col_b, (char)sep);
  *pos = L.pos; *a = L.a; *b = L.b; *chr_off = L.chr_off; *chr_len = L.chr_len;
  return L.status;
}
}
