// C entry points over the reference's own Basis recursions (src/Basis.cc),
// compiled from /root/reference in place.  TEST INFRASTRUCTURE ONLY.
#include "Basis.H"
extern "C" {
void ref_legendre_R(int lmax, double x, double *p)
{
  YAML::Node conf; Basis b(nullptr, conf);
  Eigen::MatrixXd P(lmax + 1, lmax + 1);
  b.legendre_R(lmax, x, P);
  for (int l = 0; l <= lmax; l++) for (int m = 0; m <= lmax; m++) p[l * (lmax + 1) + m] = P(l, m);
}
void ref_dlegendre_R(int lmax, double x, double *p, double *dp)
{
  YAML::Node conf; Basis b(nullptr, conf);
  Eigen::MatrixXd P(lmax + 1, lmax + 1), D(lmax + 1, lmax + 1);
  b.dlegendre_R(lmax, x, P, D);
  for (int l = 0; l <= lmax; l++) for (int m = 0; m <= lmax; m++) {
    p[l * (lmax + 1) + m] = P(l, m); dp[l * (lmax + 1) + m] = D(l, m);
  }
}
void ref_sinecosine_R(int mmax, double phi, double *c, double *s)
{
  YAML::Node conf; Basis b(nullptr, conf);
  Eigen::VectorXd C(mmax + 1), S(mmax + 1);
  b.sinecosine_R(mmax, phi, C, S);
  for (int m = 0; m <= mmax; m++) { c[m] = C[m]; s[m] = S[m]; }
}
}
