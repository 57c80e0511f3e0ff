/*
 * expb200.h -- C ABI of the B200-native basis-function-expansion force loop.
 *
 * This is the drop-in boundary.  The reference has no C
 * ABI for primary forces; its GPU seams are C++ members compiled under
 * HAVE_LIBCUDA.  Every entry point below names the reference seam it replaces,
 * so that a maintainer can forward that member to the call (INTEGRATION.md
 * shows the stubs).  Plain pointers and sizes only; no exceptions cross the
 * boundary: every function returns 0 on success or a negative EXPB200_E* code,
 * and expb200_last_error() returns the message for the calling thread.
 *
 * Threading: one host thread per GPU (the reference calls these seams from the
 * rank's main thread, src/Component.H:1041-1079).  Handles are not re-entrant.
 *
 * Units / layout conventions
 *   - host particle arrays are the reference's AoS fields flattened:
 *     pos[n][3], vel[n][3], acc[n][3], mass[n], pot[n], potext[n],
 *     level[n] (unsigned), dtreq[n] (float), scale[n] (float)
 *     (include/Particle.H:31-67).  Slot i on the host is slot i on return:
 *     particle order is never changed across the ABI.
 *   - spherical coefficients: rows l*l + {0, 2m-1 (cos), 2m (sin)}, nmax
 *     doubles per row, exactly expcoef[row][n] (src/SphericalBasis.cc:513-590).
 */
#ifndef EXPB200_H
#define EXPB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EXPB200_OK          0
#define EXPB200_EINVAL     -1   /* bad argument                              */
#define EXPB200_ECUDA      -2   /* CUDA runtime error (message has details)  */
#define EXPB200_ENOMEM     -3
#define EXPB200_ESTATE     -4   /* call made in the wrong state              */
#define EXPB200_ENCCL      -5

typedef struct expb200_component_s *expb200_component;
typedef struct expb200_sphere_s    *expb200_sphere;
typedef struct expb200_cylinder_s  *expb200_cylinder;

const char *expb200_last_error(void);
/* Library/ABI version, number of visible devices, name + SM count of a device. */
int  expb200_version(void);
int  expb200_device_count(int *count);
int  expb200_device_info(int device, char *name, int name_len, int *sm_count, size_t *hbm_bytes);

/* Pinned host memory for asynchronous transfers (cudaHostAlloc / cudaHostRegister). */
int  expb200_host_alloc(void **ptr, size_t bytes);
int  expb200_host_free(void *ptr);
int  expb200_host_register(void *ptr, size_t bytes);
int  expb200_host_unregister(void *ptr);

/* ------------------------------------------------------------------------
 * Component: the device-resident particle store.
 * Replaces Component::ParticlesToCuda / CudaToParticles / HostToDev / DevToHost
 * (src/Component.H:1082-1139, src/cudaComponent.cu).
 * ---------------------------------------------------------------------- */
int expb200_component_create(int device, size_t capacity, expb200_component *out);
int expb200_component_destroy(expb200_component c);

/* Upload n particles (n <= capacity).  vel, level, scale, potext may be NULL
 * (zeros / level 0 / scale -1).  ~ Component::ParticlesToCuda. */
int expb200_component_upload(expb200_component c, size_t n,
                             const double *pos, const double *vel, const double *mass,
                             const unsigned *level, const float *scale, const double *potext);
/* Chunked upload/download for pipelined host<->device traffic: particles
 * [first, first+count) from host arrays that start at particle `first`.
 * Asynchronous on the component's copy stream when the host memory is pinned;
 * expb200_component_sync() waits for everything issued so far. */
int expb200_component_upload_range(expb200_component c, size_t first, size_t count,
                                   const double *pos, const double *vel, const double *mass);
int expb200_component_set_count(expb200_component c, size_t n);
int expb200_component_sync(expb200_component c);

/* Download; any pointer may be NULL.  ~ Component::CudaToParticles. */
int expb200_component_download(expb200_component c,
                               double *pos, double *vel, double *acc, double *pot,
                               unsigned *level, float *dtreq);
int expb200_component_download_range(expb200_component c, size_t first, size_t count,
                                     double *pos, double *vel, double *acc, double *pot);
size_t expb200_component_count(expb200_component c);

/* Device-side timing across both streams of a component (slot in [0,8)). */
int expb200_component_event_record(expb200_component c, int slot);
int expb200_component_event_elapsed(expb200_component c, int slot_a, int slot_b, double *ms);

/* Install levels (Particle::level) / acc+pot (restart, tests). */
int expb200_component_set_levels(expb200_component c, const unsigned *level);
int expb200_component_set_acc_pot(expb200_component c, const double *acc, const double *pot);

/* center = com0 + center, what Pos(..., Local|Centered) subtracts
 * (src/Component.H:747-757); rtrunc for Component::freeze
 * (src/Component.cc:4194-4202). */
int expb200_component_set_center(expb200_component c, const double center[3], double rtrunc);

/* Zero acc/pot(/potext) of particles with level >= minlev.
 * ~ Component::ZeroPotAccel (src/cudaComponent.cu:729-797). */
int expb200_component_zero_pot_acc(expb200_component c, int minlev);

/* Leapfrog.  mlevel < 0: all particles; else levlist[mlevel].
 * ~ incr_position_cuda / incr_velocity_cuda (src/cudaIncpos.cu, src/cudaIncvel.cu),
 * arithmetic of src/incpos.cc:63-66, src/incvel.cc:68-69. */
int expb200_incr_position(expb200_component c, double dt, int mlevel);
int expb200_incr_velocity(expb200_component c, double dt, int mlevel);
/* The per-level kicks and drifts of one sub-step (the loops of src/step.cc:215-237, 253-262 over
 * incr_velocity / incr_position) in one pass: a particle of level M in [mlevel_lo, nlev) gets
 * vel += acc * dt_kick[M] (dt_kick != NULL), then pos += vel * dt_drift[M] (dt_drift != NULL).
 * Same arithmetic per particle as the two calls above; only the active levels are visited. */
int expb200_move_levels(expb200_component c, const double *dt_kick, const double *dt_drift, int nlev,
                        int mlevel_lo);

/* Multistep level assignment.  ~ cuda_compute_levels (src/cudaMultistep.cu),
 * arithmetic of adjust_multistep_level_thread (src/multistep.cc:52-236),
 * bit-exact with the CPU path (dtreq rounds through float). */
typedef struct {
  double dtime;
  double dynfracS, dynfracD, dynfracV, dynfracA, dynfracP;
  int multistep;       /* highest level                                  */
  int shiftlevl;       /* max level shift per call (0 = unlimited)       */
  int mfirst_mdrft;    /* mfirst[mdrft]: lowest level that may be set    */
  int minlev;          /* examine particles with level >= minlev         */
  int noswitch;        /* Component::NoSwitch()                          */
  int reset_dtreq;     /* (DTreset && mstep==0) || firstCall             */
  int apply;           /* assign levels on this call                     */
  int dim;             /* Component::dim                                 */
} expb200_level_params;

typedef struct {
  uint64_t examined;   /* numtt */
  uint64_t switched;   /* numsw */
  uint64_t offlo;      /* wanted a level above multistep                 */
  uint64_t offhi;      /* wanted a step larger than dtime                */
  double   mindt, maxdt;
} expb200_level_stats;

int expb200_adjust_multistep_level(expb200_component c, const expb200_level_params *p,
                                   expb200_level_stats *stats);
/* Populations per level: counts[multistep+1].  ~ CudaGetLevelRange
 * (src/cudaComponent.cu:490-526). */
int expb200_component_level_counts(expb200_component c, int multistep, uint64_t *counts);
/* Diagnostics of the device-side binning (no reference counterpart): counters[0] = (level, cell)
 * sorts of the whole component, [1] = sorts of the tail of the active levels only, [2] = particles
 * those tail sorts moved, [3] = particles the whole-component sorts moved. */
int expb200_component_sort_counters(expb200_component c, uint64_t counters[4]);

/* Centre-of-mass sums per multistep level of the non-frozen particles with
 * level >= mlevel:  sums[M*10 + {0: mass, 1..3: mass*pos, 4..6: mass*vel,
 * 7..9: mass*acc}] for M in [mlevel, multistep]; rows below mlevel are left
 * untouched (the caller keeps com_lev/cov_lev/coa_lev/com_mas across calls).
 * Components k >= dim stay 0.  With nccl_comm != NULL the rows are all-reduced
 * (the reference's four MPI_Allreduce of src/Component.cc:3499-3502).
 * ~ Component::fix_positions_cuda (src/cudaComponent.cu:800-960, src/Component.H:1133),
 * arithmetic of fix_positions_thread (src/Component.cc:3280-3352); the consp/tidal
 * escape branch (:3314-3334) is not on this path. */
int expb200_component_fix_positions(expb200_component c, int mlevel, int multistep, int dim,
                                    double *sums, void *nccl_comm);

/* ------------------------------------------------------------------------
 * Sphere (force id "sphereSL"): SphericalBasis + SLGridSph tables.
 * ---------------------------------------------------------------------- */
typedef struct {
  int lmax, nmax, numr, cmap;      /* SLGridSph (include/SLGridMP2.H:28-240)      */
  double rmin_grid, rmax_grid;     /* radial table range                          */
  double rmap;                     /* coordinate-map scale (rmapping)             */
  double scale;                    /* SphericalBasis::scale                       */
  double rmin, rmax;               /* expansion radii (src/Sphere.cc:65-66)       */
  int NO_L0, NO_L1, EVEN_L, EVEN_M, M0_only;   /* src/SphericalBasis.cc:114-119   */
  int multistep;                   /* number of levels - 1 (0 = single level)     */
} expb200_sphere_config;

/* Tables as SLGridSph holds them: ev[(lmax+1)][nmax], ef[(lmax+1)][nmax][numr],
 * p0[numr].  ~ SLGridSph::initialize_cuda (exputil/cudaSLGridMP2.cu:52-140) +
 * SphericalBasis::initialize_cuda. */
int expb200_sphere_create(int device, const expb200_sphere_config *cfg,
                          const double *ev, const double *ef, const double *p0,
                          expb200_sphere *out);
int expb200_sphere_destroy(expb200_sphere s);
/* The component this force belongs to (PotAccel::component, set by Component::configure,
 * src/Component.cc:1070-1115).  When the force is evaluated on ANOTHER component's particles
 * (SetExternal, src/ComponentContainer.cc:785-823) positions are taken about the OWNER's centre
 * (cC->Pos(Inertial) then component->ConvertPos(Local|Centered), src/SphericalBasis.cc:1519-1540)
 * while Component::freeze keeps using the particles' own centre and rtrunc.  NULL (default):
 * the component passed to each call is its own origin.  The owner must outlive the force or be
 * reset to NULL. */
int expb200_sphere_set_owner(expb200_sphere s, expb200_component owner);

/* ~ SphericalBasis::determine_coefficients_cuda(compute) + DtoH_coefs
 * (src/SphericalBasis.cc:832-843).  Accumulates levlist[mlevel] (all particles
 * when the sphere was created with multistep == 0) into the device copy of
 * expcoef0; if coef_host != NULL the (lmax+1)^2*nmax doubles are copied back.
 * used (may be NULL) receives the number of particles inside [rmin,rmax].
 * adb = Component::Adiabatic(). */
int expb200_sphere_determine_coefficients(expb200_sphere s, expb200_component c, int mlevel,
                                          double adb, double *coef_host, uint64_t *used);

/* Device pointer of the last accumulated coefficient vector ((lmax+1)^2*nmax
 * doubles), e.g. to all-reduce it in place with NCCL on `stream`. */
int expb200_sphere_coef_device(expb200_sphere s, double **dptr, void **stream);

/* ~ SphericalBasis::HtoD_coefs(expcoef) (src/SphericalBasis.cc:1707): install
 * the coefficients used by the next force evaluation.  coef_host == NULL means
 * "use the vector accumulated on the device by determine_coefficients". */
int expb200_sphere_set_coefficients(expb200_sphere s, const double *coef_host);

/* ~ SphericalBasis::determine_acceleration_cuda (src/SphericalBasis.cc:1715):
 * adds this basis' acceleration and potential to particles of c with
 * level >= mlevel.  c may be a different component (use_external). */
int expb200_sphere_determine_acceleration(expb200_sphere s, expb200_component c, int mlevel);

/* Multistep coefficient tables (expcoefN / expcoefL per level) kept on the device.
 *  commit : swap L<->N for mlevel and store the vector just accumulated as expcoefN[mlevel]
 *           (src/SphericalBasis.cc:785-788, 864-903)
 *  compute_multistep_coefficients(mdrft): interpolate inactive levels, sum active ones and
 *           install the result for the next force evaluation (src/SphericalBasis.cc:1231-1333)
 *  multistep_update: move the contribution of every particle whose level changed in the last
 *           expb200_adjust_multistep_level from its old to its new level table
 *           (src/SphericalBasis.cc:1013-1228); finish=0 leaves the differences in the device
 *           buffer (expb200_sphere_differ_device) for an all-reduce before _finish. */
int expb200_sphere_multistep_commit(expb200_sphere s, int mlevel);
int expb200_sphere_compute_multistep_coefficients(expb200_sphere s, int mdrft);
int expb200_sphere_multistep_update(expb200_sphere s, expb200_component c, double adb, int finish);
int expb200_sphere_multistep_update_finish(expb200_sphere s);
int expb200_sphere_differ_device(expb200_sphere s, double **dptr, size_t *count, void **stream);
/* which = 0: expcoefN, 1: expcoefL, 2: differ (of level mlevel); 3: the installed vector
 * (expcoef, what the force uses), 4: the last accumulated vector (after any all-reduce) */
int expb200_sphere_get_level_coefficients(expb200_sphere s, int which, int mlevel, double *host);
int expb200_sphere_set_level_coefficients(expb200_sphere s, int which, int mlevel, const double *host);

/* Covariance of the coefficients (pcavar), the `compute` branch of
 * determine_coefficients_thread (src/SphericalBasis.cc:501-509, 525-545, 563-581;
 * ~ the covariance part of determine_coefficients_cuda, src/cudaSphericalBasis.cu) for
 * levlist[mlevel].  Bin T = (indx_base + host slot) % sampT.  Host outputs, any may be NULL:
 *   countT[sampT], massT[sampT]                    countT1 / massT1
 *   coefT[sampT][ldim][nmax][2]                    expcoefT1 (re, im)
 *   coefM[sampT][ldim][nmax][nmax]                 expcoefM1 (real; imaginary part is 0)
 * ldim = (lmax+1)(lmax+2)/2, index l(l+1)/2 + m.  lmax <= 6, nmax <= 32.  The component's
 * cell sort is replaced by the (bin, cell) sort of this call. */
int expb200_sphere_covariance(expb200_sphere s, expb200_component c, int mlevel, double adb,
                              int sampT, uint64_t indx_base, unsigned *countT, double *massT,
                              double *coefT, double *coefM);

/* The pcaeof accumulation (`compute && pcaeof`, src/SphericalBasis.cc:536-545, 573-581) for
 * levlist[mlevel]: tvar[ldim][nmax][nmax] (overwritten) = sum_p wk[n] wk[o] / mass, where the m > 0
 * products carry no angular factor. */
int expb200_sphere_covariance_eof(expb200_sphere s, expb200_component c, int mlevel, double adb,
                                  double *tvar);

/* One fused single-level KDK step for a self-gravitating component, the
 * sequence of do_step (src/step.cc:272-322) with the arithmetic of the
 * separate calls: kick(dt/2) drift(dt) coefficients force kick(dt/2).
 * (Coefficient sums run in the component's storage order, so two components
 * with different call histories agree to rounding, ~1e-14 relative; the same
 * history always gives the same bits.)
 * If comm != NULL it is an ncclComm_t and the coefficient vector is summed
 * across ranks (the reference's MPI_Allreduce, src/SphericalBasis.cc:864-903). */
int expb200_sphere_step(expb200_sphere s, expb200_component c, double dt, double adb,
                        void *nccl_comm);
/* nsteps such steps back to back (the main loop of src/expand.cc around do_step): the same
 * arithmetic, with the first kick + drift of step k+1 riding in the force kernel of step k. */
int expb200_sphere_steps(expb200_sphere s, expb200_component c, double dt, int nsteps, double adb,
                         void *nccl_comm);

/* ------------------------------------------------------------------------
 * Cylinder (force id "cylinder"): Cylinder + EmpCylSL tables.
 * Tables are the six EmpCylSL grids potC/potS/rforceC/rforceS/zforceC/zforceS
 * (exputil/EmpCylSL.cc:2123-2187), each laid out [(mmax+1)][norder][(numx+1)][(numy+1)]
 * (ix major, iy minor; the m=0 slabs of the S grids are unused).
 * ---------------------------------------------------------------------- */
typedef struct {
  int mmax, norder, numx, numy, cmapr, cmapz;  /* MMAX, NORDER, NUMX, NUMY, CMAPR, CMAPZ           */
  double ascale, hscale;                        /* ASCALE, HSCALE                                   */
  double rmin, rmax;                            /* EmpCylSL::RMIN, RMAX (units of ascale)           */
  double rcylmax, acyl;                         /* Cylinder gate Rmax2 = (rcylmax*acyl)^2           */
  int EVEN_M, MMIN, MLIM, NMIN, NLIM;           /* include/EmpCylSL.H:59-60                         */
  int multistep;
} expb200_cylinder_config;

/* ~ EmpCylSL::initialize_cuda (src/cudaEmpCylSL.cu:56-150) + Cylinder::initialize_cuda */
int expb200_cylinder_create(int device, const expb200_cylinder_config *cfg,
                            const double *potC, const double *potS, const double *rforceC,
                            const double *rforceS, const double *zforceC, const double *zforceS,
                            expb200_cylinder *out);
int expb200_cylinder_destroy(expb200_cylinder s);
/* as expb200_sphere_set_owner (src/Cylinder.cc:1343-1349); the cylinder's force pass has no
 * freeze test */
int expb200_cylinder_set_owner(expb200_cylinder s, expb200_component owner);
/* orient->transformBody() / transformOrig() (row-major 3x3); NULL, NULL = none
 * (src/Cylinder.cc:846-847, 1352-1353, 1418-1419) */
int expb200_cylinder_set_orientation(expb200_cylinder s, const double *body, const double *orig);
/* ~ Cylinder::determine_coefficients_cuda + DtoH_coefs (src/Cylinder.cc:1057-1069,
 * src/cudaCylinder.cu:1871-1924): accumulates levlist[mlevel]; cos/sin [(mmax+1)][norder]
 * as EmpCylSL::set_coef takes them; used = use[] count, mass = cylmass0 sum. */
int expb200_cylinder_determine_coefficients(expb200_cylinder s, expb200_component c, int mlevel,
                                            double adb, double *cos_host, double *sin_host,
                                            uint64_t *used, double *mass);
/* device vector [cos | sin | mass] of the last accumulation (count doubles), for one
 * in-place all-reduce (replaces exputil/EmpCylSL.cc:4192-4214 + src/Cylinder.cc:1093-1095) */
int expb200_cylinder_coef_device(expb200_cylinder s, double **dptr, size_t *count, void **stream);
/* ~ Cylinder::HtoD_coefs: install accum_cos/accum_sin and cylmass for the force; NULL,NULL =
 * use the device vector (cylmass < 0 keeps the accumulated mass). */
int expb200_cylinder_set_coefficients(expb200_cylinder s, const double *cos_host,
                                      const double *sin_host, double cylmass);
/* ~ Cylinder::determine_acceleration_cuda (src/Cylinder.cc:1481-1497) */
int expb200_cylinder_determine_acceleration(expb200_cylinder s, expb200_component c, int mlevel);
/* multistep tables cosL/cosN (exputil/EmpCylSL.cc:2004-2014, src/CylEXP.cc:192-282) */
int expb200_cylinder_multistep_commit(expb200_cylinder s, int mlevel);
int expb200_cylinder_compute_multistep_coefficients(expb200_cylinder s, int mdrft, double cylmass);
/* ~ CylEXP::multistep_update_begin/update/finish (src/CylEXP.cc:45-188,
 * src/Cylinder.cc:1751-1773), as expb200_sphere_multistep_update.  The differ buffer is
 * [nlev][cos | sin | pad] (count doubles) for one all-reduce before _finish. */
int expb200_cylinder_multistep_update(expb200_cylinder s, expb200_component c, double adb, int finish);
int expb200_cylinder_multistep_update_finish(expb200_cylinder s);
int expb200_cylinder_differ_device(expb200_cylinder s, double **dptr, size_t *count, void **stream);
/* which = 0: cosN/sinN, 1: cosL/sinL, 2: differ (of level mlevel); 3: installed accum_cos/sin,
 * 4: last accumulated (after any all-reduce); cos_host, sin_host [(mmax+1)][norder] */
int expb200_cylinder_get_level_coefficients(expb200_cylinder s, int which, int mlevel,
                                            double *cos_host, double *sin_host);
/* the mass slot of the same vectors (cylmass0 summed; src/Cylinder.cc:1081-1098) */
int expb200_cylinder_get_mass(expb200_cylinder s, int which, int mlevel, double *mass);

/* Covariance of the coefficients (`covar`), the compute branch of EmpCylSL::accumulate
 * (exputil/EmpCylSL.cc:4069-4074, 4123-4141; ~ the covariance part of
 * Cylinder::determine_coefficients_cuda, results as EmpCylSL::set_massT/set_numbT/set_VC/
 * set_MV take them, include/EmpCylSL.H:996-1110) for levlist[mlevel].
 * Bin T = (indx_base + host slot) % sampT.  Host outputs, any may be NULL:
 *   numbT[sampT], massT[sampT]
 *   VC[sampT][(mmax+1)][norder][2]                 mass * vec             (re, im)
 *   MV[sampT][(mmax+1)][norder][norder][2]         mass * vec vec^H       (re, im)
 * norder <= 32.  The component's cell sort is replaced by the (bin, cell) sort of this call. */
int expb200_cylinder_covariance(expb200_cylinder s, expb200_component c, int mlevel, double adb,
                                int sampT, uint64_t indx_base, unsigned *numbT, double *massT,
                                double *VC, double *MV);

/* The PCAVAR / PCAEOF variants of the same branch (exputil/EmpCylSL.cc:4089-4119; results as
 * EmpCylSL::set_covV/set_covM/set_tvar take them): which = 1 (PCAVAR): covV[T][m][n] += mass modu,
 * covM[T][m](n,o) += mass modu_n modu_o with modu = |(Vc cos m phi, Vs sin m phi)|; which = 2
 * (PCAEOF): tvar[m](n,o) += mass modu_n modu_o with modu = 4 pi |(Vc, Vs)| (use sampT = 1 and read
 * covM).  Real outputs covV[sampT][(mmax+1)][norder], covM[sampT][(mmax+1)][norder][norder]. */
int expb200_cylinder_covariance_pca(expb200_cylinder s, expb200_component c, int mlevel, double adb,
                                    int sampT, uint64_t indx_base, int which, unsigned *numbT,
                                    double *massT, double *covV, double *covM);

/* ------------------------------------------------------------------------
 * Multi-GPU: one process (MPI rank) per GPU; NCCL is loaded with dlopen.
 * Replaces the host MPI_Allreduce of the coefficient vector
 * (src/SphericalBasis.cc:864-903) by one in-place device all-reduce.
 * ---------------------------------------------------------------------- */
int expb200_comm_unique_id(char id[128]);                       /* ncclGetUniqueId (rank 0) */
int expb200_comm_create(int device, int nranks, int rank, const char id[128], void **comm);
int expb200_comm_destroy(void *comm);
/* Calls on one communicator may use different streams (one per component): each call is
 * ordered after the previous one.  ~ MPI_Allreduce(MPI_SUM, MPI_DOUBLE) in place. */
int expb200_allreduce_sum(void *comm, double *device_buf, size_t n, void *stream);
/* The same for a host vector (staged through a device buffer, blocking): the reference's
 * MPI_Allreduce of the covariance arrays (src/AxisymmetricBasis.cc:1224-1254,
 * exputil/EmpCylSL.cc:4272-4275, 4323-4327, 4567-4571). */
int expb200_allreduce_sum_host(void *comm, double *host_buf, size_t n);
/* Error state of the peer-memory all-reduces issued so far (after a sync): a rank that never
 * arrived within EXPB200_PEER_TIMEOUT_S seconds (default 600; 0 = wait for ever, like MPI) or
 * out-of-order sequence numbers fill the result with NaN and make this return EXPB200_ENCCL. */
int expb200_comm_check(void *comm);
/* Optional: all-reduces of up to 16384 doubles (every one on this path) done by one kernel
 * over NVLink peer memory instead of ncclAllReduce: each rank stores its vector into every
 * peer's receive buffer, then adds the world's vectors in rank order (bitwise identical on
 * all ranks).  _peer_handle exports this rank's buffer (a cudaIpcMemHandle_t); the host
 * gathers the handles of all ranks, in rank order, and passes them to _peer_connect.
 * Ranks must be processes on one node.  Without these calls NCCL is used. */
int expb200_comm_peer_handle(void *comm, char handle[64]);
int expb200_comm_peer_connect(void *comm, const char *handles /* [nranks][64] */);

/* ------------------------------------------------------------------------
 * Instrumentation (what bench.py reads): launch counters are always on; event
 * timing of each kernel family on its launching stream only while enabled.
 * ~ the reference's per-phase timers (src/step.cc:347-374, `timers: true`).
 * ---------------------------------------------------------------------- */
int expb200_profile_enable(int on);
int expb200_profile_reset(void);
int expb200_profile_kernels(void);                 /* number of kernel families */
const char *expb200_profile_name(int id);
int expb200_profile_read(unsigned long long *launches, double *ms, unsigned long long *timed);
/* Dependency-free DFMA microbenchmark: the FP64 roofline denominator. */
int expb200_fp64_peak(int device, double *tflops, double *sm_mhz_hint);
/* FP64 tensor-core (mma.sync f64) microbenchmark by shape (0: m8n8k4, 1: m16n8k4, 2: m16n8k8,
 * 3: m16n8k16): the roofline denominator of the covariance kernels. */
int expb200_dmma_peak(int device, int shape, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* EXPB200_H */
