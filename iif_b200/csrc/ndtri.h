// Inverse of the standard normal CDF, float64: a restatement of the published Cephes `ndtri`
// algorithm (S. Moshier, Cephes Math Library 2.1) which is what scipy.special.ndtri evaluates --
// the reference calls it at classification/custom.py:4,20 for the `normit` variant.
// Three rational approximations: central region |y-0.5| <= 0.5-exp(-2) in (y-0.5)^2, and two tail
// regions in z = 1/sqrt(-2 ln y) split at sqrt(-2 ln y) = 8.  Shared by the CUDA weight kernel and
// the host build used by the CPU tests (tests/test_ndtri_host.py checks it against scipy).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define IIF_HD __host__ __device__ __forceinline__
#else
#define IIF_HD static inline
#endif

IIF_HD double iif_polevl(double x, const double* c, int n) {
  double r = c[0];
  for (int i = 1; i <= n; ++i) r = r * x + c[i];
  return r;
}
IIF_HD double iif_p1evl(double x, const double* c, int n) {  // leading coefficient 1 implied
  double r = x + c[0];
  for (int i = 1; i < n; ++i) r = r * x + c[i];
  return r;
}

IIF_HD double iif_ndtri(double y0) {
  const double P0[5] = {-5.99633501014107895267E1, 9.80010754185999661536E1, -5.66762857469070293439E1,
                        1.39312609387279679503E1, -1.23916583867381258016E0};
  const double Q0[8] = {1.95448858338141759834E0, 4.67627912898881538453E0, 8.63602421390890590575E1,
                        -2.25462687854119370527E2, 2.00260212380060660359E2, -8.20372256168333339912E1,
                        1.59056225126211695515E1, -1.18331621121330003142E0};
  const double P1[9] = {4.05544892305962419923E0, 3.15251094599893866154E1, 5.71628192246421288162E1,
                        4.40805073893200834700E1, 1.46849561928858024014E1, 2.18663306850790267539E0,
                        -1.40256079171354495875E-1, -3.50424626827848203418E-2, -8.57456785154685413611E-4};
  const double Q1[8] = {1.57799883256466749731E1, 4.53907635128879210584E1, 4.13172038254672030440E1,
                        1.50425385692907503408E1, 2.50464946208309415979E0, -1.42182922854787788574E-1,
                        -3.80806407691578277194E-2, -9.33259480895457427372E-4};
  const double P2[9] = {3.23774891776946035970E0, 6.91522889068984211695E0, 3.93881025292474443415E0,
                        1.33303460815807542389E0, 2.01485389549179081538E-1, 1.23716634817820021358E-2,
                        3.01581553508235416007E-4, 2.65806974686737550832E-6, 6.23974539184983293730E-9};
  const double Q2[8] = {6.02427039364742014255E0, 3.67983563856160859403E0, 1.37702099489081330271E0,
                        2.16236993594496635890E-1, 1.34204006088543189037E-2, 3.28014464682127739104E-4,
                        2.89247864745380683936E-6, 6.79019408009981274425E-9};
  const double s2pi = 2.50662827463100050242E0;      // sqrt(2 pi)
  const double expm2 = 0.13533528323661269189;       // exp(-2)
  if (y0 == 0.0) return -INFINITY;
  if (y0 == 1.0) return INFINITY;
  if (!(y0 > 0.0 && y0 < 1.0)) return NAN;
  int negate = 1;
  double y = y0;
  if (y > 1.0 - expm2) { y = 1.0 - y; negate = 0; }
  if (y > expm2) {
    y -= 0.5;
    const double y2 = y * y;
    double x = y + y * (y2 * iif_polevl(y2, P0, 4) / iif_p1evl(y2, Q0, 8));
    return x * s2pi;
  }
  double x = sqrt(-2.0 * log(y));
  const double x0 = x - log(x) / x;
  const double z = 1.0 / x;
  const double x1 = (x < 8.0) ? z * iif_polevl(z, P1, 8) / iif_p1evl(z, Q1, 8)
                              : z * iif_polevl(z, P2, 8) / iif_p1evl(z, Q2, 8);
  x = x0 - x1;
  return negate ? -x : x;
}
