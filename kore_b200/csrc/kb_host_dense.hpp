// Small dense complex kernels on the host for the projected (ncv x ncv) problem
// of the Krylov-Schur iteration.
//
// Replaces SLEPc's DS object (LAPACK zgehrd/zhseqr/ztrexc/ztrevc on the
// projected matrix, inside EPSSolve, /root/reference/bin/solve.py:123).  ncv is
// max(2 nev, nev+15) ~ 20..30, so this is latency, not throughput: it stays on
// the host.  Everything is built from one primitive, the complex
// Givens rotation  [c s; -conj(s) c] [f; g] = [r; 0]  with real c.
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <vector>

namespace kbd {

typedef std::complex<double> Z;

struct Mat {  // column-major
  int m = 0, n = 0;
  std::vector<Z> a;
  Mat() {}
  Mat(int m_, int n_) : m(m_), n(n_), a((size_t)m_ * n_, Z(0, 0)) {}
  Z& operator()(int i, int j) { return a[(size_t)i + (size_t)j * m]; }
  const Z& operator()(int i, int j) const { return a[(size_t)i + (size_t)j * m]; }
};

inline void givens(Z f, Z g, double& c, Z& s, Z& r) {
  double af = std::abs(f), ag = std::abs(g);
  if (ag == 0.0) {
    c = 1.0;
    s = Z(0, 0);
    r = f;
    return;
  }
  if (af == 0.0) {
    c = 0.0;
    s = std::conj(g) / ag;
    r = Z(ag, 0);
    return;
  }
  double d = std::hypot(af, ag);
  c = af / d;
  Z fs = f / af;
  s = fs * std::conj(g) / d;
  r = fs * d;
}

// rows (i, j) of A, columns [c0, c1):  [x; y] <- G [x; y]
inline void rot_rows(Mat& A, int i, int j, int c0, int c1, double c, Z s) {
  for (int k = c0; k < c1; ++k) {
    Z x = A(i, k), y = A(j, k);
    A(i, k) = c * x + s * y;
    A(j, k) = -std::conj(s) * x + c * y;
  }
}
// columns (i, j) of A, rows [r0, r1):  [x y] <- [x y] G^H
inline void rot_cols(Mat& A, int i, int j, int r0, int r1, double c, Z s) {
  for (int k = r0; k < r1; ++k) {
    Z x = A(k, i), y = A(k, j);
    A(k, i) = c * x + std::conj(s) * y;
    A(k, j) = -s * x + c * y;
  }
}

// Complex Schur decomposition H = Q T Q^H of a general m x m matrix.
// On exit H holds T (upper triangular), Q the unitary factor.  Returns false
// if the QR iteration failed to converge.
inline bool schur(Mat& H, Mat& Q) {
  const int m = H.m;
  Q = Mat(m, m);
  for (int i = 0; i < m; ++i) Q(i, i) = 1.0;
  // ---- Hessenberg form by Givens similarity
  for (int k = 0; k + 2 < m; ++k) {
    for (int i = m - 1; i >= k + 2; --i) {
      if (H(i, k) == Z(0, 0)) continue;
      double c;
      Z s, r;
      givens(H(i - 1, k), H(i, k), c, s, r);
      rot_rows(H, i - 1, i, k, m, c, s);
      H(i, k) = Z(0, 0);
      rot_cols(H, i - 1, i, 0, m, c, s);
      rot_cols(Q, i - 1, i, 0, m, c, s);
    }
  }
  // ---- implicit single-shift QR with deflation
  const double eps = 2.220446049250313e-16;
  double hn = 0.0;
  for (const Z& z : H.a) hn = std::max(hn, std::abs(z));
  const double small = std::max(hn, 1e-300) * 1e-300 / eps;
  int hi = m - 1;
  int iter = 0;
  while (hi > 0) {
    int l;
    for (l = hi; l > 0; --l) {
      double sref = std::abs(H(l - 1, l - 1)) + std::abs(H(l, l));
      if (sref == 0.0) sref = hn;
      if (std::abs(H(l, l - 1)) <= std::max(eps * sref, small)) {
        H(l, l - 1) = Z(0, 0);
        break;
      }
    }
    if (l == hi) {
      --hi;
      iter = 0;
      continue;
    }
    if (++iter > 60 * std::max(10, m)) return false;
    Z mu;
    if (iter % 10 == 0) {
      mu = H(hi, hi) + Z(0.75 * std::abs(H(hi, hi - 1)), 0);
    } else {
      Z a = H(hi - 1, hi - 1), b = H(hi - 1, hi), c_ = H(hi, hi - 1), d = H(hi, hi);
      Z half = 0.5 * (a - d);
      Z disc = std::sqrt(half * half + b * c_);
      Z mu1 = 0.5 * (a + d) + disc, mu2 = 0.5 * (a + d) - disc;
      mu = (std::abs(mu1 - d) <= std::abs(mu2 - d)) ? mu1 : mu2;
    }
    Z x = H(l, l) - mu, y = H(l + 1, l);
    for (int k = l; k < hi; ++k) {
      double c;
      Z s, r;
      givens(x, y, c, s, r);
      int c0 = (k > l) ? k - 1 : l;
      rot_rows(H, k, k + 1, c0, m, c, s);
      if (k > l) H(k + 1, k - 1) = Z(0, 0);
      int r1 = std::min(k + 3, hi + 1);
      rot_cols(H, k, k + 1, 0, r1, c, s);
      rot_cols(Q, k, k + 1, 0, m, c, s);
      if (k + 1 < hi) {
        x = H(k + 1, k);
        y = H(k + 2, k);
      }
    }
  }
  for (int j = 0; j < m; ++j)
    for (int i = j + 1; i < m; ++i) H(i, j) = Z(0, 0);
  return true;
}

// Swap the adjacent diagonal entries k, k+1 of the upper triangular T (and
// update Q) by a unitary similarity (the complex case of LAPACK's ztrexc).
inline void swap_adjacent(Mat& T, Mat& Q, int k) {
  const int m = T.m;
  Z t11 = T(k, k), t22 = T(k + 1, k + 1);
  double c;
  Z s, r;
  givens(T(k, k + 1), t22 - t11, c, s, r);
  if (k + 2 < m) rot_rows(T, k, k + 1, k + 2, m, c, s);
  rot_cols(T, k, k + 1, 0, k, c, s);
  T(k, k) = t22;
  T(k + 1, k + 1) = t11;
  rot_cols(Q, k, k + 1, 0, Q.m, c, s);
}

// Reorder so that the diagonal is sorted by ascending key (stable selection).
template <typename KeyFn>
inline void sort_schur(Mat& T, Mat& Q, KeyFn key) {
  const int m = T.m;
  for (int i = 0; i < m; ++i) {
    int best = i;
    double kb = key(T(i, i));
    for (int j = i + 1; j < m; ++j) {
      double kj = key(T(j, j));
      if (kj < kb) {
        kb = kj;
        best = j;
      }
    }
    for (int j = best; j > i; --j) swap_adjacent(T, Q, j - 1);
  }
}

// Unit-norm eigenvector z (length i+1) of the leading (i+1)x(i+1) block of the
// upper triangular T for the eigenvalue T(i,i).
inline std::vector<Z> tri_eigvec(const Mat& T, int i) {
  std::vector<Z> z(i + 1, Z(0, 0));
  z[i] = 1.0;
  double tn = 0.0;
  for (int c = 0; c <= i; ++c)
    for (int r = 0; r <= c; ++r) tn = std::max(tn, std::abs(T(r, c)));
  const double smin = std::max(2.220446049250313e-16 * tn, 1e-300);
  const Z th = T(i, i);
  for (int r = i - 1; r >= 0; --r) {
    Z acc = -T(r, i);
    for (int c = r + 1; c < i; ++c) acc -= T(r, c) * z[c];
    Z d = T(r, r) - th;
    if (std::abs(d) < smin) d = Z(smin, 0);
    z[r] = acc / d;
  }
  double nrm = 0.0;
  for (const Z& v : z) nrm += std::norm(v);
  nrm = std::sqrt(nrm);
  for (Z& v : z) v /= nrm;
  return z;
}

}  // namespace kbd
