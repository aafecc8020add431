// Text writer for Kore's solution files (host code, no device work).
//
// The reference writes every field of every solution with np.savetxt and its defaults
// (/root/reference/bin/solve.py:283-306: fmt '%.18e', one space, '\n'); spin_doctor.py:27-61,
// koreviz/kmode.py:81-107 and tools/ read those files back with np.loadtxt, so the format
// has to stay.  At the E = 1e-8 size the flow field of 10 eigenvectors is 2 x 92 MB of text
// and np.savetxt needs 2.8 s per file -- several times the factorisation and the eigensolve
// together.  kb_savetxt formats row blocks on all host threads and writes them in order;
// the bytes are those of np.savetxt (both CPython's '%.18e' and glibc's are correctly
// rounded; non-finite values are spelled the way CPython spells them).
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <thread>
#include <vector>

#include "kb_internal.cuh"

namespace {
inline char* kb_fmt_e18(char* p, double v) {
  if (isnan(v)) {
    memcpy(p, "nan", 3);
    return p + 3;
  }
  if (isinf(v)) {
    if (v < 0) *p++ = '-';
    memcpy(p, "inf", 3);
    return p + 3;
  }
  return p + snprintf(p, 32, "%.18e", v);
}
}  // namespace

extern "C" int kb_savetxt(const char* path, const double* data, int64_t rows, int64_t cols, int64_t row_stride,
                          int64_t col_stride, int append, int nthreads) {
  if (!path || (!data && rows * cols > 0) || rows < 0 || cols < 0) return KB_EINVAL;
  FILE* f = fopen(path, append ? "ab" : "wb");
  if (!f) return KB_EINVAL;
  if (nthreads <= 0) {
    nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads <= 0) nthreads = 1;
  }
  if (nthreads > 64) nthreads = 64;
  // bounded memory: blocks of at most ~4 M numbers (100 MB of text) are formatted in parallel
  const int64_t maxrows = cols > 0 ? (((int64_t)1 << 22) / cols > 0 ? ((int64_t)1 << 22) / cols : 1) : rows;
  const size_t per = (size_t)(cols > 0 ? cols : 1) * 28 + 2;  // "-d.{18}e+ddd" = 26 bytes + separator
  std::vector<std::vector<char>> buf(nthreads);
  std::vector<size_t> len(nthreads);
  int rc = KB_OK;
  for (int64_t r0 = 0; r0 < rows && rc == KB_OK; r0 += maxrows) {
    const int64_t nr = rows - r0 < maxrows ? rows - r0 : maxrows;
    const int nt = nr < nthreads ? (int)(nr > 0 ? nr : 1) : nthreads;
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t) {
      pool.emplace_back([&, t]() {
        const int64_t lo = r0 + nr * t / nt, hi = r0 + nr * (t + 1) / nt;
        buf[t].resize((size_t)(hi - lo) * per + 8);
        char* p = buf[t].data();
        for (int64_t i = lo; i < hi; ++i) {
          const double* row = data + i * row_stride;
          for (int64_t j = 0; j < cols; ++j) {
            if (j) *p++ = ' ';
            p = kb_fmt_e18(p, row[j * col_stride]);
          }
          *p++ = '\n';
        }
        len[t] = (size_t)(p - buf[t].data());
      });
    }
    for (auto& th : pool) th.join();
    for (int t = 0; t < nt; ++t)
      if (len[t] && fwrite(buf[t].data(), 1, len[t], f) != len[t]) rc = KB_EINVAL;
  }
  if (fclose(f) != 0) rc = KB_EINVAL;
  return rc;
}
