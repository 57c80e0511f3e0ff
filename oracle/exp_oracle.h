/*
 * exp_oracle.h -- C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the reference
 * (EXP-code/EXP v7.9.3) basis-function-expansion force loop, written from
 * the reference sources' behaviour, each function citing the file:line it
 * follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product (exp_b200/, libexpb200.so)
 * never does.
 *
 * PARITY PINNED to the reference's own code: the reference cannot be built as a whole here
 * (needs Eigen3, MPI, HDF5, gfortran), but its hot-path member functions compile on their
 * own.  oracle/ref_extract.py cuts them from /root/reference where they lie (into the
 * git-ignored oracle/_ref/gen/), oracle/ref_path_wrap.cc compiles them unmodified against
 * shim declarations (oracle/ref_shim/pathshim.H) into oracle/_ref/libexpref_path.so, and
 * tests/test_oracle_ref.py checks this restatement against them BIT FOR BIT: Basis
 * recursions, SLGridSph maps/get_pot/get_force/get_dens, the factorial table,
 * SphericalBasis::determine_coefficients_thread (+ pcavar/pcaeof) and
 * determine_acceleration_and_potential_thread, EmpCylSL r_to_xi/z_to_y/get_pot/accumulate/
 * accumulated_eval (the covar branch to 1e-14: Eigen expression order), incr_position/
 * incr_velocity threads, adjust_multistep_level_thread.  The reference's own N-body test case
 * (tests/Halo: SLGridSph.model, Lmax=2 nmax=10 numr=4000 multistep=4, 2T/VC check) runs from
 * tests/golden/halo_fixture.npz (tests/test_halo_reference_case.py).  The pyEXP twins
 * (oracle/pyexp_oracle.py) are held to the reference's own BasisClasses::Spherical::accumulate /
 * sph_eval / cyl_eval / crt_eval / computeAccel and BiorthBasis::legendre_R (expui/BiorthBasis.cc,
 * same mechanism) to 1e-13, and the Cylinder:: thread functions around EmpCylSL (centre / rtrunc /
 * rcylmax gates, orientation, erf blend to the monopole; src/Cylinder.cc:748-896, 1266-1446) bit
 * for bit, as are fix_positions_thread (src/Component.cc:3280-3352) and SphericalBasis::multistep_update
 * (src/SphericalBasis.cc:1156-1228) and Cylinder::multistep_update / CylEXP::multistep_update
 * (src/Cylinder.cc:1752-1773, src/CylEXP.cc:159-188) and the level-table interpolation
 * SphericalBasis::compute_multistep_coefficients (:1231-1474) with the cylinder's copy
 * CylEXP::compute_multistep_coefficients (src/CylEXP.cc:192-282).  Not pinned (restated only): the
 * cylindrical pyEXP twin (a thin wrapper over the pinned EmpCylSL::accumulate / accumulated_eval).
 *
 * All arrays are plain row-major C arrays.  Coefficient layout is the
 * reference's: row index = l*l + {0 (m=0), 2m-1 (cos), 2m (sin)}, nmax
 * doubles per row  (src/SphericalBasis.cc:513-590).
 */
#ifndef EXP_ORACLE_H
#define EXP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- a1-a4: recursions (src/Basis.cc:14-111, src/SphericalBasis.cc:330-335) */
void   orc_legendre_R (int lmax, double x, double *p);              /* p[l*(lmax+1)+m] */
void   orc_dlegendre_R(int lmax, double x, double *p, double *dp);
void   orc_sinecosine_R(int mmax, double phi, double *c, double *s);
double orc_factorial  (int l, int m);

/* ---- a5-a7: SLGridSph radial table (exputil/SLGridMP2.cc:711-765,872-989,1355-1382) */
typedef struct {
  int lmax, nmax, numr, cmap;
  double rmin, rmax, rmap;
  const double *ev;   /* [(lmax+1)][nmax]            */
  const double *ef;   /* [(lmax+1)][nmax][numr]      */
  const double *p0;   /* [numr]                      */
  const double *d0;   /* [numr] (may be NULL)        */
  /* derived by orc_slgrid_init */
  double xmin, xmax, dxi;
} orc_slgrid;

void   orc_slgrid_init (orc_slgrid *g);
double orc_r_to_xi     (const orc_slgrid *g, double r);
double orc_xi_to_r     (const orc_slgrid *g, double xi);
double orc_d_xi_to_r   (const orc_slgrid *g, double xi);
void   orc_slgrid_get_pot  (const orc_slgrid *g, double r, double *mat); /* [(lmax+1)][nmax] */
void   orc_slgrid_get_force(const orc_slgrid *g, double r, double *mat);
void   orc_slgrid_get_dens (const orc_slgrid *g, double r, double *mat);

/* ---- a8-a11: SphericalBasis / Sphere */
typedef struct {
  double scale;            /* SphericalBasis::scale  (src/SphericalBasis.cc:97-100) */
  double rmin, rmax;       /* expansion radii after Sphere ctor (src/Sphere.cc:65-66) */
  int NO_L0, NO_L1, EVEN_L, EVEN_M, M0_only;
} orc_sph_params;

/* Particle set: structure of plain arrays, n particles.  `center` is
 * com0+center, i.e. what Pos(.., Local|Centered) subtracts
 * (src/Component.H:747-757); rtrunc is Component::rtrunc (freeze test,
 * src/Component.cc:4194-4202). */
typedef struct {
  long n;
  double *pos;     /* [n][3] */
  double *vel;     /* [n][3] */
  double *acc;     /* [n][3] */
  double *pot;     /* [n]    */
  double *potext;  /* [n]    */
  double *mass;    /* [n]    */
  double *scale;   /* [n]   Particle::scale (may be NULL => -1) */
  unsigned *level; /* [n]    */
  float  *dtreq;   /* [n]    */
  double center[3];
  double rtrunc;
} orc_particles;

/* determine_coefficients_thread + thread sum (src/SphericalBasis.cc:429-599,855-858).
 * list/nlist = levlist[mlevel]; the slice of thread id is
 * [nlist*id/nthreads, nlist*(id+1)/nthreads).  coef: [(lmax+1)^2][nmax], overwritten.
 * Returns number of particles used (sum of use[id]). */
long orc_sph_coefficients(const orc_slgrid *g, const orc_sph_params *P,
                          const orc_particles *S, const int *list, long nlist,
                          double adb, int nthreads, double *coef);

/* Same with the pcavar covariance (src/SphericalBasis.cc:501-509,525-545,563-581):
 * whch = indx % sampT;  massT[sampT], countT[sampT],
 * coefT[sampT][(L+1)(L+2)/2][nmax][2 (re,im)],
 * coefM[sampT][(L+1)(L+2)/2][nmax][nmax] (real part; the imaginary part is
 * identically zero in the reference).  All accumulated into (caller zeroes). */
long orc_sph_coefficients_cov(const orc_slgrid *g, const orc_sph_params *P,
                              const orc_particles *S, const int *list, long nlist,
                              double adb, int sampT, double *coef, double *muse,
                              double *massT, unsigned *countT,
                              double *coefT, double *coefM);

/* determine_acceleration_and_potential_thread (src/SphericalBasis.cc:1476-1660)
 * for the particles in list (the caller concatenates levlist[lev>=mlevel]).
 * Adds into acc and pot. */
void orc_sph_accel(const orc_slgrid *g, const orc_sph_params *P,
                   const orc_particles *S, const int *list, long nlist,
                   const double *coef, int nthreads);

/* multistep_update for one particle (src/SphericalBasis.cc:1156-1228):
 * differ[(multistep+1)][(lmax+1)^2][nmax], from/to levels. */
void orc_sph_multistep_update(const orc_slgrid *g, const orc_sph_params *P,
                              const orc_particles *S, int indx, double adb,
                              int from, int to, double *differ);

/* compute_multistep_coefficients (src/SphericalBasis.cc:1231-1333):
 * coefL/coefN [(multistep+1)][ncoef]; out[ncoef]. */
void orc_multistep_interp(int multistep, int mdrft, const int *mfirst,
                          const int *dstepL, const int *dstepN, /* [(multistep+1)][Mstep] */
                          long ncoef, const double *coefL, const double *coefN,
                          double *out);

/* ---- a21: leapfrog (src/incpos.cc:15-70, src/incvel.cc:15-89) */
void orc_incr_position(orc_particles *S, const int *list, long nlist, double dt, int dim);
void orc_incr_velocity(orc_particles *S, const int *list, long nlist, double dt, int dim);

/* ---- a22-a23: multistep (src/multistep.cc:52-236,630-679) */
typedef struct {
  double dtime;
  double dynfracS, dynfracD, dynfracV, dynfracA, dynfracP;
  int multistep, shiftlevl;
  int mfirst_mdrft;        /* mfirst[mdrft] */
  int noswitch, reset_dtreq, apply;
  int dim;
} orc_level_params;

/* Returns number of level switches.  newlevel[i] for i in list is written
 * (and S->level updated, S->dtreq updated).  which[i] = index of the
 * controlling criterion (dseq.begin()->second).  offlo/offhi counters. */
long orc_adjust_levels(orc_particles *S, const int *list, long nlist,
                       const orc_level_params *P, int *which,
                       long *offlo, long *offhi);

/* initialize_multistep tables: mintvl[multistep+1], mfirst[Mstep+1],
 * dstepL/dstepN[(multistep+1)][Mstep]. */
void orc_multistep_tables(int multistep, int *mintvl, int *mfirst,
                          int *dstepL, int *dstepN);

/* f1: Component::fix_positions_cpu (src/Component.cc:3280-3590): lev_sums[(multistep+1)][10]
 * = {com_mas, com_lev[3], cov_lev[3], coa_lev[3]} (rows >= mlevel recomputed),
 * out[10] = {mtot, com[3], cov[3], coa[3]}. */
void orc_fix_positions(const orc_particles *S, int mlevel, int multistep, int dim,
                       double *lev_sums, double *out);

/* ---- a14-a19: cylinder (EmpCylSL tables + Cylinder force)
 * Tables are six arrays [(MMAX+1)][NORDER][(NUMX+1)][(NUMY+1)] (ix major, iy minor); the
 * m=0 slabs of the S arrays are unused.  exputil/EmpCylSL.cc:2123-2187 (setup_table). */
typedef struct {
  int MMAX, NORDER, NUMX, NUMY, CMAPR, CMAPZ;
  double ASCALE, HSCALE, RMIN, RMAX;       /* EmpCylSL statics; Rtable = RMAX/sqrt(2)        */
  int EVEN_M, MMIN, MLIM, NMIN, NLIM;      /* include/EmpCylSL.H:59-60 (MLIM/NLIM may be huge) */
  const double *potC, *potS, *rforceC, *rforceS, *zforceC, *zforceS;
  /* derived by orc_cyl_init */
  double Rtable, XMIN, XMAX, dX, YMIN, YMAX, dY;
} orc_cylgrid;

typedef struct {
  double rcylmax, acyl;        /* src/Cylinder.cc:752: Rmax2 = (rcylmax*acyl)^2              */
  const double *body;          /* orient->transformBody() 3x3 row-major, or NULL            */
  const double *orig;          /* orient->transformOrig() 3x3 row-major, or NULL            */
} orc_cyl_params;

void orc_cyl_init(orc_cylgrid *g);
double orc_cyl_r_to_xi(const orc_cylgrid *g, double r);    /* exputil/EmpCylSL.cc:6446-6463 */
double orc_cyl_z_to_y(const orc_cylgrid *g, double z);     /* exputil/EmpCylSL.cc:7109-7117 */
/* EmpCylSL::get_pot (exputil/EmpCylSL.cc:5557-5632): Vc, Vs [(MMAX+1)][NORDER] */
void orc_cyl_get_pot(const orc_cylgrid *g, double r, double z, double *Vc, double *Vs);
/* Cylinder::determine_coefficients_thread (src/Cylinder.cc:823-890) + EmpCylSL::accumulate
 * (exputil/EmpCylSL.cc:4049-4146) + thread sum of make_coefficients (:4151-4346).
 * cosN, sinN: [(MMAX+1)][NORDER], overwritten; *cylmass = sum of used masses. */
long orc_cyl_coefficients(const orc_cylgrid *g, const orc_cyl_params *P, const orc_particles *S,
                          const int *list, long nlist, double adb, int nthreads,
                          double *cosN, double *sinN, double *cylmass);
/* covariance variant (covar): VC[sampT][(MMAX+1)][NORDER][2], MV[sampT][(MMAX+1)][NORDER][NORDER][2],
 * massT[sampT], numbT[sampT]; accumulated into (caller zeroes). */
long orc_cyl_coefficients_cov(const orc_cylgrid *g, const orc_cyl_params *P, const orc_particles *S,
                              const int *list, long nlist, double adb, int sampT,
                              double *cosN, double *sinN, double *cylmass,
                              double *massT, unsigned *numbT, double *VC, double *MV);
/* Cylinder::multistep_update + CylEXP::multistep_update (src/Cylinder.cc:1751-1773,
 * src/CylEXP.cc:159-188): differC, differS [(multistep+1)][(MMAX+1)][NORDER] */
void orc_cyl_multistep_update(const orc_cylgrid *g, const orc_particles *S, int indx, double adb,
                              int from, int to, double *differC, double *differS);
/* EmpCylSL::accumulated_eval (exputil/EmpCylSL.cc:5256-5410) */
void orc_cyl_accumulated_eval(const orc_cylgrid *g, const double *accum_cos, const double *accum_sin,
                              double r, double z, double phi,
                              double *p0, double *p, double *fr, double *fz, double *fp);
/* Cylinder::determine_acceleration_and_potential_thread (src/Cylinder.cc:1266-1446) */
void orc_cyl_accel(const orc_cylgrid *g, const orc_cyl_params *P, const orc_particles *S,
                   const int *list, long nlist, const double *accum_cos, const double *accum_sin,
                   double cylmass, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
