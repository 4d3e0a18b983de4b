// On-disk contract writer: the per-modality CSV families the test program writes with
// DataFrame.to_csv(index=False) -- normalized_*, reconstruction_*, reconstruction_error_*, reconstruction_error_roi_*,
// deviation_as_feature_importance_* (multimodal_kfold_test_cvae_supervised.py:116-178) -- as one host pass over a numeric
// block, threads over row ranges, one write().  Every number is printed the way pandas prints a float column
// (ndarray.astype(str), i.e. numpy's scalar str): the shortest digit string that round-trips at the column's precision,
// positional for 1e-4 <= |v| < 1e16 (float32 columns: < 1e6), otherwise scientific with a sign and at least two exponent digits, "x.0" for
// integral values, nan -> empty cell (na_rep), inf / -inf.  The files are byte-identical to pandas' (tests/test_host_cpu.py).
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "nmb_internal.h"

namespace nmb {
namespace {

template <class T>
inline void put_number(std::string& out, T v) {
  if (std::isnan(v)) return;                                   // na_rep = ''
  if (std::isinf(v)) { out += v < 0 ? "-inf" : "inf"; return; }
  char buf[48];
  // shortest round-trip digits as d[.ddd]e[+-]XX
  auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);
  const char* p = buf;
  if (*p == '-') { out += '-'; ++p; }
  char digits[24];
  int nd = 0;
  const char* e = p;
  for (; e < r.ptr && *e != 'e'; ++e)
    if (*e != '.') digits[nd++] = *e;
  int ex = 0;
  std::from_chars(e + 1 + (e[1] == '+'), r.ptr, ex);
  const double a = v < 0 ? -(double)v : (double)v;              // thresholds compared in higher precision, like numpy
  const double upper = sizeof(T) == 4 ? 1e6 : 1e16;             // (float32(1e-4) < 1e-4 prints as 1e-04)
  if (a == 0.0 || (a < upper && a >= 1e-4)) {                   // numpy: positional
    if (ex >= 0) {
      for (int i = 0; i <= ex; ++i) out += i < nd ? digits[i] : '0';
      out += '.';
      if (nd > ex + 1) out.append(digits + ex + 1, nd - ex - 1); else out += '0';
    } else {
      out += "0.";
      out.append(-ex - 1, '0');
      out.append(digits, nd);
    }
  } else {                                                      // scientific, trimmed mantissa, >= 2 exponent digits
    out += digits[0];
    if (nd > 1) { out += '.'; out.append(digits + 1, nd - 1); }
    out += 'e';
    out += ex < 0 ? '-' : '+';
    const int ax = ex < 0 ? -ex : ex;
    char eb[8];
    const int n = std::snprintf(eb, sizeof eb, "%02d", ax);
    out.append(eb, n);
  }
}

template <class T>
void format_rows(std::string& out, const char* const* prefix, const T* body, long long r0, long long r1, int n_cols, long long ld) {
  out.reserve((size_t)(r1 - r0) * (n_cols * 20 + 32));
  for (long long r = r0; r < r1; ++r) {
    bool first = true;
    if (prefix && prefix[r]) { out += prefix[r]; first = prefix[r][0] == 0; }
    const T* row = body + r * ld;
    for (int c = 0; c < n_cols; ++c) {
      if (!first) out += ',';
      first = false;
      put_number(out, row[c]);
    }
    out += '\n';
  }
}

}  // namespace

int csv_write(const char* path, const char* header, const char* const* prefix, const void* body, int is_f64, long long n_rows,
              int n_cols, long long ld, int threads, std::string& err) {
  int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  const long long min_rows = 256;                               // below this a thread costs more than it formats
  if ((long long)nt > (n_rows + min_rows - 1) / min_rows) nt = (int)((n_rows + min_rows - 1) / min_rows);
  if (nt < 1) nt = 1;
  std::vector<std::string> parts(nt);
  auto work = [&](int t) {
    const long long r0 = n_rows * t / nt, r1 = n_rows * (t + 1) / nt;
    if (is_f64) format_rows(parts[t], prefix, static_cast<const double*>(body), r0, r1, n_cols, ld);
    else format_rows(parts[t], prefix, static_cast<const float*>(body), r0, r1, n_cols, ld);
  };
  if (nt == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
  }
  FILE* f = std::fopen(path, "wb");
  if (!f) { err = std::string("cannot open ") + path + ": " + std::strerror(errno); return 1; }
  bool ok = true;
  if (header) ok = std::fputs(header, f) >= 0 && std::fputc('\n', f) != EOF;
  for (int t = 0; ok && t < nt; ++t) ok = parts[t].empty() || std::fwrite(parts[t].data(), 1, parts[t].size(), f) == parts[t].size();
  if (std::fclose(f) != 0) ok = false;
  if (!ok) { err = std::string("write failed: ") + path; return 1; }
  return 0;
}

}  // namespace nmb
