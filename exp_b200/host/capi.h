/* capi.h -- flat C entry points of the C++ host driver (libexphost.so) so that the parity
 * tests and bench.py can assemble and run a simulation (components + forces + interactions,
 * begin_run, do_step) from Python.  This is a harness around the C++ classes, not a second
 * boundary: the drop-in boundary of the GPU path is include/expb200.h.
 * Every function returns 0 or a negative status; exphost_last_error() has the message. */
#ifndef EXPHOST_CAPI_H
#define EXPHOST_CAPI_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int lmax, nmax, numr, cmap;
  double rmin, rmax, rmap;             /* SLGridSph grid */
  const double *ev, *ef, *p0;
  double scale, rmin_exp, rmax_exp;    /* < 0: table range */
  int NO_L0, NO_L1, EVEN_L, EVEN_M, M0_only, self_consistent;
} exphost_sphere_desc;

typedef struct {
  int mmax, norder, numx, numy, cmapr, cmapz;
  double ascale, hscale, rmin, rmax;
  const double *potC, *potS, *rforceC, *rforceS, *zforceC, *zforceS;
  double rcylmax, acyl;                /* acyl < 0: ascale */
  int EVEN_M, MMIN, MLIM, NMIN, NLIM, self_consistent;
  const double *body, *orig;           /* orientation or NULL */
} exphost_cylinder_desc;

const char *exphost_last_error(void);
int exphost_reset(void);
/* globals of src/global.cc: dtime, multistep, dynfrac{S,D,V,A,P}, shiftlevl */
int exphost_set_params(double dtime, int multistep, double dynfracS, double dynfracD, double dynfracV,
                       double dynfracA, double dynfracP, int shiftlevl);
int exphost_set_comm(void *nccl_comm, int myid, int numprocs);
int exphost_add_component(const char *name, size_t n, const double *pos, const double *vel,
                          const double *mass, int device, int *index);
int exphost_component_options(int index, const double center[3], double rtrunc, int noswitch,
                              int freezeLev, int dtreset);
int exphost_set_sphere(int index, const exphost_sphere_desc *d);
int exphost_set_cylinder(int index, const exphost_cylinder_desc *d);
/* sharded runs: this rank's particles are global [indx_base, indx_base + n) of nbodies_tot */
int exphost_component_global(int index, uint64_t indx_base, uint64_t nbodies_tot);
/* covariance / PCA options of the sphereSL force (src/AxisymmetricBasis.cc:64-92): npca, npca0,
 * nint (nint > 0 implies pcavar), pcavar, pcaeof, samplesz (0: floor(sqrt(N))), subsamp,
 * tk_type (0 Hall, 1 VarianceCut, 2 CumulativeCut, 3 VarianceWeighted, 4 None), hexp, snr,
 * tksmooth, tkcum */
int exphost_sphere_pca_options(int index, int npca, int npca0, int nint, int pcavar, int pcaeof,
                               int samplesz, int subsamp, int tk_type, double hexp, double snr,
                               double tksmooth, double tkcum);
/* the gathered subsample arrays after the last compute step: sampT, then massT[sampT],
 * countT[sampT], coefT[sampT][ldim][nmax][2], coefM[sampT][ldim][nmax][nmax], tvar[ldim][nmax][nmax],
 * b_Hall / s_Hall / weight [ldim][nmax] (any pointer may be NULL) */
int exphost_sphere_get_covariance(int index, int *sampT, double *massT, unsigned *countT, double *coefT,
                                  double *coefM, double *tvar, double *b_Hall, double *s_Hall,
                                  double *weight, double *muse);
/* the same switches for the cylinder force (src/Cylinder.cc:1004-1025): nint turns the `covar`
 * accumulation on, pcavar / pcaeof the PCAVAR / PCAEOF ones; results after the last compute step
 * (any pointer may be NULL): numbT, massT [sampT]; VC [sampT][mmax+1][norder][2],
 * MV [sampT][mmax+1][norder][norder][2]; covV, covM (real); tvar [mmax+1][norder][norder] */
int exphost_cylinder_cov_options(int index, int nint, int npca, int npca0, int pcavar, int pcaeof, int samplesz);
int exphost_cylinder_get_covariance(int index, int *sampT, unsigned *numbT, double *massT, double *VC,
                                    double *MV, double *covV, double *covM, double *tvar);
/* Hall smoothing of the cylinder force (src/Cylinder.cc:231-241, 1115-1134): tk_type 0 Hall,
 * 1 Truncate, 2 None; results of the last analysis (any pointer may be NULL): meanJK, evalJK, ratio
 * [mmax+1][norder], evecJK [mmax+1][norder][norder] (eigenvectors = columns); returns 1 in *valid
 * when an analysis exists */
int exphost_cylinder_pca_options(int index, int tk_type, double snr, double hexp, int subsamp);
int exphost_cylinder_get_pca(int index, int *valid, double *meanJK, double *evalJK, double *ratio, double *evecJK,
                             double *raw_cos, double *raw_sin /* [mmax+1][norder]: before the last smoothing */);
/* Stateless EmpCylSL::pca_hall (PCAVAR part) + set_trimmed (exputil/EmpCylSL.cc:4577-4970, 5141-5207)
 * on given partition arrays (no GPU needed): in numbT, massT [sampT], covV [sampT][mmax+1][norder],
 * covM [sampT][mmax+1][norder][norder]; out as exphost_cylinder_get_pca; cosv, sinv
 * [mmax+1][norder] smoothed in place (may be NULL) */
int exphost_cyl_pca_hall(int mmax, int norder, int sampT, int subsamp, double nbodstot, int tk_type, double snr,
                         double hexp, const unsigned *numbT, const double *massT, const double *covV,
                         const double *covM, double *meanJK, double *evalJK, double *ratio, double *evecJK,
                         double *cosv, double *sinv);
/* Stateless AxisymmetricBasis::pca_hall_3d + the smoothing pass (src/AxisymmetricBasis.cc:198-660)
 * on given subsample arrays (no GPU needed): in massT[sampT], coefT[sampT][ldim][nmax][2],
 * coefM[sampT][ldim][nmax][nmax]; out b_Hall, s_Hall, weight, evalJK [ldim][nmax], evec
 * [ldim][nmax][nmax]; expcoef[(Lmax+1)^2][nmax] smoothed in place (may be NULL). */
int exphost_pca_hall(int Lmax, int nmax, int sampT, int used, int subsamp, int tk_type, double hexp,
                     double snr, double tksmooth, double tkcum, const double *massT, const double *coefT,
                     const double *coefM, double *b_Hall, double *s_Hall, double *weight, double *evalJK,
                     double *evec, double *expcoef);
int exphost_add_interaction(int from, int on);
int exphost_begin_run(void);
int exphost_do_steps(int nsteps);
/* device-timed: nsteps of do_step between two events on component 0's stream */
int exphost_do_steps_timed(int nsteps, double *ms);
int exphost_get_particles(int index, double *pos, double *vel, double *acc, double *pot,
                          unsigned *level, float *dtreq);
int exphost_get_sphere_coefs(int index, int which, int M, double *coef);
int exphost_get_cylinder_coefs(int index, double *cosv, double *sinv, double *cylmass);
int exphost_level_report(int index, uint64_t *examined, uint64_t *switched, uint64_t *offlo,
                         uint64_t *offhi, uint64_t *counts /* [multistep+1] */);
/* device binning diagnostics of a component (expb200_component_sort_counters) */
int exphost_sort_counters(int index, uint64_t counters[4]);
/* initialize_multistep tables for `multistep` levels (no GPU needed): mintvl[multistep+1],
 * mfirst[Mstep+1], dstepL/dstepN[(multistep+1)][Mstep] */
int exphost_multistep_tables(int multistep, int *mintvl, int *mfirst, int *dstepL, int *dstepN);
/* The force's coefficient output: kind 0 = dump_coefs stream (NewCoefs), 1 = old-style stream
 * (sphere), 2 = the SphStruct/CylStruct complex array.  Returns the byte count (buf may be
 * NULL) or a negative status. */
long exphost_dump_coefs(int index, int kind, char *buf, long cap);
/* Component::mtot/com/cov/coa after the last fix_positions: out[10] = {mtot, com[3], cov[3],
 * coa[3]}; lev[(multistep+1)][10] = {com_mas, com_lev, cov_lev, coa_lev} (may be NULL). */
int exphost_get_com(int index, double *out, double *lev);
int exphost_state(double *tnow, int *this_step, double *pos_check, double *vel_check);

#ifdef __cplusplus
}
#endif
#endif
