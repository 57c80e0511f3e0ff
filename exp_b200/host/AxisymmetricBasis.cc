// AxisymmetricBasis.cc -- see AxisymmetricBasis.H
#include "AxisymmetricBasis.H"

#include <algorithm>
#include <cmath>

#include "GenericError.H"
#include "expb200.h"

namespace exphost {

void jacobi_eigen(int n, const std::vector<double> &sym, std::vector<double> &eval, std::vector<double> &evec)
{
  std::vector<double> a(sym);
  evec.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) evec[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) (i == j ? diag : off) += a[(size_t)i * n + j] * a[(size_t)i * n + j];
    if (off <= 1e-30 * (diag + 1e-300)) break;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++) {
        const double apq = a[(size_t)p * n + q];
        if (apq == 0.0) continue;
        const double theta = (a[(size_t)q * n + q] - a[(size_t)p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; k++) {                  // A <- A J
          const double akp = a[(size_t)k * n + p], akq = a[(size_t)k * n + q];
          a[(size_t)k * n + p] = c * akp - s * akq;
          a[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {                  // A <- J^T A
          const double apk = a[(size_t)p * n + k], aqk = a[(size_t)q * n + k];
          a[(size_t)p * n + k] = c * apk - s * aqk;
          a[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {                  // V <- V J
          const double vkp = evec[(size_t)k * n + p], vkq = evec[(size_t)k * n + q];
          evec[(size_t)k * n + p] = c * vkp - s * vkq;
          evec[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  std::vector<int> order(n);
  for (int i = 0; i < n; i++) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int x, int y) { return a[(size_t)x * n + x] > a[(size_t)y * n + y]; });
  std::vector<double> v2((size_t)n * n);
  eval.resize(n);
  for (int j = 0; j < n; j++) {
    const int src = order[j];
    eval[j] = a[(size_t)src * n + src];
    int big = 0;
    for (int k = 1; k < n; k++)
      if (std::fabs(evec[(size_t)k * n + src]) > std::fabs(evec[(size_t)big * n + src])) big = k;
    const double sg = evec[(size_t)big * n + src] < 0.0 ? -1.0 : 1.0;
    for (int k = 0; k < n; k++) v2[(size_t)k * n + j] = sg * evec[(size_t)k * n + src];
  }
  evec.swap(v2);
}

void AxisymmetricBasis::update_compute(int mstep, int this_step)
{
  compute = false;
  if (nint) {
    if (mstep == 0 && this_step % nint == 0) { compute = true; requestSubsample = true; }
    else { requestSubsample = false; subsampleComputed = false; }
  }
  if (pcavar || pcaeof) {
    if (this_step >= npca0) compute = (mstep == 0) && !((this_step - npca0) % npca);
  }
}

void AxisymmetricBasis::allocate(size_t ntot)
{
  if (!massT.empty()) return;
  sampT = defSampT ? defSampT : (unsigned)std::floor(std::sqrt((double)ntot));
  if (sampT < 1) sampT = 1;
  const size_t L = (size_t)ldim();
  massT.assign(sampT, 0.0); massT1.assign(sampT, 0.0);
  countT.assign(sampT, 0); countT1.assign(sampT, 0);
  expcoefT.assign(sampT * L * nmax, 0.0); expcoefT1 = expcoefT;
  expcoefM.assign(sampT * L * nmax * nmax, 0.0); expcoefM1 = expcoefM;
  tvar.assign(L * nmax * nmax, 0.0);
}

void AxisymmetricBasis::zero_accumulation()
{
  muse0 = 0.0;
  used_cov = 0;
  if (pcavar) {
    std::fill(expcoefT1.begin(), expcoefT1.end(), std::complex<double>(0.0, 0.0));
    std::fill(expcoefM1.begin(), expcoefM1.end(), 0.0);
    std::fill(countT1.begin(), countT1.end(), 0u);
    std::fill(massT1.begin(), massT1.end(), 0.0);
  }
  if (pcaeof) std::fill(tvar.begin(), tvar.end(), 0.0);
}

void AxisymmetricBasis::add_level(const unsigned *cnt, const double *mass, const double *cT, const double *cM)
{
  const size_t L = (size_t)ldim();
  for (unsigned T = 0; T < sampT; T++) {
    countT1[T] += cnt[T];
    massT1[T] += mass[T];
    muse0 += mass[T];
    used_cov += (int)cnt[T];
  }
  for (size_t k = 0; k < sampT * L * nmax; k++) expcoefT1[k] += std::complex<double>(cT[2 * k], cT[2 * k + 1]);
  for (size_t k = 0; k < sampT * L * nmax * nmax; k++) expcoefM1[k] += cM[k];
}

void AxisymmetricBasis::add_tvar(const double *tv)
{
  for (size_t k = 0; k < tvar.size(); k++) tvar[k] += tv[k];
}

static void reduce(void *comm, double *buf, size_t n)
{
  if (!comm || n == 0) return;
  int rc = expb200_allreduce_sum_host(comm, buf, n);
  if (rc) throw GenericError(std::string("covariance all-reduce: ") + expb200_last_error(), __FILE__, __LINE__, rc);
}

void AxisymmetricBasis::gather(void *comm)
{
  // MPI_Allreduce(&muse0, &muse, ...) (src/SphericalBasis.cc:920-923) and
  // parallel_gather_coef2 (src/AxisymmetricBasis.cc:1204-1288)
  muse = muse0;
  reduce(comm, &muse, 1);
  if (pcavar) {
    massT = massT1;
    reduce(comm, massT.data(), massT.size());
    std::vector<double> cnt(countT1.begin(), countT1.end());
    reduce(comm, cnt.data(), cnt.size());
    for (size_t k = 0; k < cnt.size(); k++) countT[k] = (unsigned)std::llround(cnt[k]);
    expcoefT = expcoefT1;
    reduce(comm, reinterpret_cast<double *>(expcoefT.data()), 2 * expcoefT.size());
    expcoefM = expcoefM1;
    reduce(comm, expcoefM.data(), expcoefM.size());
  }
  if (pcaeof) reduce(comm, tvar.data(), tvar.size());
  requestSubsample = false;
  subsampleComputed = true;
}

// pca_hall_3d (src/AxisymmetricBasis.cc:198-588) without the diagnostic files and VTK output
void AxisymmetricBasis::pca_hall(bool compute_now, int used, const double *expcoef)
{
  (void)expcoef;
  if (muse <= 0.0 || !compute_now || !pcavar) return;
  const int L = ldim(), N = nmax;
  b_Hall.assign((size_t)L * N, 0.0);
  s_Hall.assign((size_t)L * N, 1.0);
  weight.assign((size_t)L * N, 1.0);
  evec.assign((size_t)L * N * N, 0.0);
  evalJK.assign((size_t)L * N, 0.0);
  std::vector<std::complex<double>> meanJK(N);
  std::vector<double> covr((size_t)N * N), ev, vec;
  for (int iC = 0; iC < L; iC++) {
    std::fill(meanJK.begin(), meanJK.end(), std::complex<double>(0.0, 0.0));
    std::fill(covr.begin(), covr.end(), 0.0);
    for (unsigned T = 0; T < sampT; T++)
      if (massT[T] > 0.0)
        for (int i = 0; i < N; i++) meanJK[i] += expcoefT[((size_t)T * L + iC) * N + i] / (massT[T] * sampT);
    for (unsigned T = 0; T < sampT; T++) {
      if (!(massT[T] > 0.0)) continue;
      for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) {
          if (subsamp) {
            const std::complex<double> di = expcoefT[((size_t)T * L + iC) * N + i] / massT[T] - meanJK[i];
            const std::complex<double> dj = expcoefT[((size_t)T * L + iC) * N + j] / massT[T] - meanJK[j];
            covr[(size_t)i * N + j] += (di * dj).real() / (double)sampT;
          } else {
            covr[(size_t)i * N + j] += expcoefM[(((size_t)T * L + iC) * N + i) * N + j] / (massT[T] * sampT);
          }
        }
    }
    if (!subsamp)
      for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) covr[(size_t)i * N + j] -= (meanJK[i] * meanJK[j]).real();
    // the reference hands covrJK.real() to a general eigen-solver; the matrix is symmetric up to
    // rounding: symmetrise and use Jacobi
    for (int i = 0; i < N; i++)
      for (int j = i + 1; j < N; j++) {
        const double s = 0.5 * (covr[(size_t)i * N + j] + covr[(size_t)j * N + i]);
        covr[(size_t)i * N + j] = covr[(size_t)j * N + i] = s;
      }
    jacobi_eigen(N, covr, ev, vec);
    for (int i = 0; i < N * N; i++) evec[(size_t)iC * N * N + i] = vec[i];
    double cum = 0.0, tot = 0.0;
    for (int n = 0; n < N; n++) tot += ev[n];
    for (int n = 0; n < N; n++) {
      evalJK[(size_t)iC * N + n] = ev[n];
      // tt = Tevec * meanJK
      std::complex<double> tt(0.0, 0.0);
      for (int k = 0; k < N; k++) tt += vec[(size_t)k * N + n] * meanJK[k];
      const double ttabs = std::abs(tt * std::conj(tt));
      double b, var;
      if (subsamp) { b = std::fabs(ev[n]) / ttabs / sampT; var = std::fabs(ev[n]) * sampT; }
      else         { b = std::fabs(ev[n]) / ttabs / used;  var = std::fabs(ev[n]); }
      b = std::max<double>(b, std::numeric_limits<double>::min());
      b_Hall[(size_t)iC * N + n] = b;
      s_Hall[(size_t)iC * N + n] = 1.0 / (1.0 + std::pow(snr * b, hexp));
      cum += ev[n];
      double w = 1.0;
      if (tk_type == VarianceCut) w = (tksmooth * var > ttabs) ? 0.0 : 1.0;
      else if (tk_type == CumulativeCut) w = (n == 1 || cum / tot <= tkcum) ? 1.0 : 0.0;
      else if (tk_type == VarianceWeighted) w = 1.0 / (1.0 + var / (ttabs + 1.0e-14));
      weight[(size_t)iC * N + n] = w;
    }
  }
}

// src/AxisymmetricBasis.cc:590-660: project onto the eigenvectors, weight, project back (or the
// Hall factor per coefficient)
void AxisymmetricBasis::smooth(double *expcoef) const
{
  if (!pcavar || tk_type == None || evec.empty()) return;
  const int N = nmax;
  std::vector<double> smth(N), inv(N);
  for (int l = 0, loffset = 0, loffC = 0; l <= Lmax; loffset += 2 * l + 1, loffC += l + 1, l++)
    for (int m = 0, moffset = 0; m <= l; m++) {
      const int indxC = loffC + m;
      const double *V = &evec[(size_t)indxC * N * N];
      for (int cs = 0; cs < (m ? 2 : 1); cs++) {
        double *row = expcoef + (size_t)(loffset + moffset + cs) * N;
        for (int n = 0; n < N; n++) {
          double dd = 0.0;
          for (int nn = 0; nn < N; nn++) dd += V[(size_t)nn * N + n] * row[nn];        // Tevec(n, nn)
          smth[n] = dd * weight[(size_t)indxC * N + n];
        }
        for (int n = 0; n < N; n++) {
          double s = 0.0;
          for (int k = 0; k < N; k++) s += V[(size_t)n * N + k] * smth[k];             // evec * smth
          inv[n] = s;
        }
        for (int n = 0; n < N; n++) {
          if (tk_type == Hall) row[n] *= s_Hall[(size_t)indxC * N + n];
          else                 row[n] = inv[n];
        }
      }
      moffset += m ? 2 : 1;
    }
}

} // namespace exphost
