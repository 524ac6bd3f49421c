"""wv_exp2_neg / wv_exp2_fast (csrc/wv_common.cuh; wv_exp2_neg is the 2^u of the squared-exponential leaves) compiled for
the host and compared with long-double exp2l: relative error below 2 ulp over the range the Gram kernels use, exact
special cases."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include <cstdio>
#include <cmath>
#include <random>
#include "%s/waveome_b200/csrc/wv_common.cuh"
int main() {
  double tab[WV_EXP2_TAB];
  for (int j = 0; j < WV_EXP2_TAB; ++j) wv_exp2_table_entry(j, &tab[j]);
  std::mt19937_64 g(1);
  std::uniform_real_distribution<double> U(-1100.0, 30.0), V(-3.0, 1.0);
  double maxulp = 0;
  for (int i = 0; i < 2000000; ++i) {
    double u = (i & 1) ? U(g) : V(g);
    double r = wv_exp2_fast(u, tab);
    long double ref = exp2l((long double)u);
    if (u <= 0.0 && u > -1021.0 && wv_exp2_neg(u, tab) != r) return 5;      // the device path shares the arithmetic
    if (u > -1021.0 && wv_exp2_lo(u, tab) != r) return 7;                   // integer clamp: identical inside the range
    if (u <= -1021.0 && !(wv_exp2_lo(u, tab) > 0.0 && wv_exp2_lo(u, tab) < 1e-307)) return 8;
    if (u <= -1021.0) { if (r != 0.0 || !(wv_exp2_neg(u, tab) < 1e-307)) return 2; continue; }
    double ulp = fabs((double)((r - ref) / ldexpl(1.0L, ilogbl(ref) - 52)));
    if (ulp > maxulp) maxulp = ulp;
  }
  printf("%%.4f\n", maxulp);
  if (wv_exp2_fast(0.0, tab) != 1.0 || wv_exp2_fast(-0.0, tab) != 1.0 || wv_exp2_fast(1.0, tab) != 2.0) return 3;
  if (wv_exp2_neg(0.0, tab) != 1.0 || !std::isnan(wv_exp2_neg(NAN, tab))) return 6;
  if (wv_exp2_lo(0.0, tab) != 1.0 || wv_exp2_lo(-0.0, tab) != 1.0 || !std::isnan(wv_exp2_lo(NAN, tab))) return 9;
  if (!(wv_exp2_lo(-INFINITY, tab) < 1e-307) || !(wv_exp2_lo(-1e300, tab) < 1e-307)) return 10;
  if (wv_exp2_fast(-INFINITY, tab) != 0.0 || !std::isnan(wv_exp2_fast(NAN, tab)) || !std::isinf(wv_exp2_fast(2000.0, tab))) return 4;
  return 0;
}
'''


def test_exp2_fast_accuracy(tmp_path):
    src = tmp_path / "exp2_test.cpp"
    src.write_text(SRC % ROOT)
    exe = tmp_path / "exp2_test"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", str(exe), str(src)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.returncode
    assert float(r.stdout.strip()) < 2.0


SRC12 = r'''
#include <cstdio>
#include <cmath>
#include <random>
#include <vector>
#include "%s/waveome_b200/csrc/wv_common.cuh"
int main() {
  std::vector<double> tab(WV_EXP2_BIG_TAB);
  for (int j = 0; j < WV_EXP2_BIG_TAB; ++j) tab[j] = (double)exp2l((long double)j / WV_EXP2_BIG_TAB);
  std::mt19937_64 g(2);
  std::uniform_real_distribution<double> U(-1020.0, 30.0), V(-3.0, 1.0);
  double maxulp = 0;
  for (int i = 0; i < 2000000; ++i) {
    double u = (i & 1) ? U(g) : V(g);
    double r = wv_exp2_big_lo(u, tab.data());
    long double ref = exp2l((long double)u);
    double ulp = fabs((double)((r - ref) / ldexpl(1.0L, ilogbl(ref) - 52)));
    if (ulp > maxulp) maxulp = ulp;
  }
  printf("%%.4f\n", maxulp);
  if (wv_exp2_big_lo(0.0, tab.data()) != 1.0 || wv_exp2_big_lo(-0.0, tab.data()) != 1.0 || wv_exp2_big_lo(3.0, tab.data()) != 8.0) return 3;
  if (!std::isnan(wv_exp2_big_lo(NAN, tab.data()))) return 4;
  if (!(wv_exp2_big_lo(-INFINITY, tab.data()) < 1e-307) || !(wv_exp2_big_lo(-1e300, tab.data()) < 1e-307)) return 5;
  if (!(wv_exp2_big_lo(-INFINITY, tab.data()) > 0.0)) return 6;
  return 0;
}
'''


def test_exp2_bigtable_table_accuracy(tmp_path):
    """wv_exp2_big_lo: the 2^u of the run-time specialised kernels (2048-entry table, degree-3 polynomial)."""
    src = tmp_path / "exp2_12_test.cpp"
    src.write_text(SRC12 % ROOT)
    exe = tmp_path / "exp2_12_test"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", str(exe), str(src)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.returncode
    assert float(r.stdout.strip()) < 1.5
