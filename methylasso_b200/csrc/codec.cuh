// codec.cuh — host interface of the text codec (codec.cu).
#pragma once
#include "common.h"

namespace ml {

enum : int32_t { TXT_EMPTY = 3 };   // blank line (next to TXT_OK / TXT_HOST_PARSE / TXT_MALFORMED of codec_core.cuh)

// One parsed file, resident: data line r is line r + skip of the text.
struct TextTable {
  int64_t n = 0, skip = 0, n_runs = 0;
  DevBuf<char> text;
  DevBuf<int64_t> line_start;       // byte offset of every line of the file (+ one past the last newline)
  DevBuf<int32_t> pos, chr_off, chr_len;
  DevBuf<double> a, b;
  DevBuf<int8_t> status;
  DevBuf<int64_t> run_first;        // first data line of every run of equal chromosome fields
};

void text_parse(Engine& eng, const char* text, int64_t nbytes, int64_t skip_lines, int col_a, int col_b, char sep,
                TextTable& out);

}  // namespace ml
