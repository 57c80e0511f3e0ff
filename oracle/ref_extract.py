#!/usr/bin/env python
"""oracle/ref_extract.py -- TEST INFRASTRUCTURE ONLY.

Cuts the hot-path member functions out of the reference sources *where they lie*
(/root/reference, EXP v7.9.3 @ e60c27c) into oracle/_ref/gen/*.inc, which is
git-ignored: no reference source enters the repository.  oracle/ref_*_wrap.cc then
compiles those bodies, unmodified, against the shim declarations in
oracle/ref_shim/ (plain containers in place of Eigen, a flat Component), giving
oracle/_ref/libexpref_path.so -- the reference's own arithmetic for this path, used by
tests/test_oracle_ref.py to pin the restatement in oracle/*_oracle.cc.

Each piece is (file, first line, last line, regex the first line must match); a
mismatch (a different reference revision) aborts the build instead of silently
cutting the wrong lines.
"""
from __future__ import annotations

import os
import re
import sys

PIECES = {
    # --- Basis (src/Basis.cc): Legendre / trig recursions
    "basis_mineps":        ("src/Basis.cc", 7, 7, r"^constexpr double MINEPS = 3.0\*std::numeric_limits<double>::epsilon\(\);"),
    "basis_recursions":    ("src/Basis.cc", 14, 112, r"^void Basis::legendre_R\(int lmax, double x, Eigen::MatrixXd& p\)"),
    # --- SLGridSph (exputil/SLGridMP2.cc): coordinate map + table interpolation
    "slgrid_r_to_xi":      ("exputil/SLGridMP2.cc", 711, 726, r"^double SLGridSph::r_to_xi\(double r\)"),
    "slgrid_xi_to_r":      ("exputil/SLGridMP2.cc", 728, 746, r"^double SLGridSph::xi_to_r\(double xi\)"),
    "slgrid_d_xi_to_r":    ("exputil/SLGridMP2.cc", 748, 765, r"^double SLGridSph::d_xi_to_r\(double xi\)"),
    "slgrid_get_pot_mat":  ("exputil/SLGridMP2.cc", 872, 910, r"^void SLGridSph::get_pot\(Eigen::MatrixXd& mat, double x, int which\)"),
    "slgrid_get_dens_mat": ("exputil/SLGridMP2.cc", 913, 951, r"^void SLGridSph::get_dens\(Eigen::MatrixXd& mat, double x, int which\)"),
    "slgrid_get_force_mat": ("exputil/SLGridMP2.cc", 954, 989, r"^void SLGridSph::get_force\(Eigen::MatrixXd& mat, double x, int which\)"),
    "slgrid_init_grid":    ("exputil/SLGridMP2.cc", 1362, 1373, r"^  if \(cmap==1\) \{"),
    # --- Sphere (src/Sphere.cc): binds SLGridSph as the radial basis
    "sphere_get_dpotl":    ("src/Sphere.cc", 140, 145, r"^void Sphere::get_dpotl\("),
    "sphere_get_potl":     ("src/Sphere.cc", 147, 150, r"^void Sphere::get_potl\("),
    # --- SphericalBasis (src/SphericalBasis.cc)
    "sph_factorial":       ("src/SphericalBasis.cc", 328, 335, r"^  factorial.resize\(Lmax\+1, Lmax\+1\);"),
    "sph_coef_thread":     ("src/SphericalBasis.cc", 429, 599, r"^void \* SphericalBasis::determine_coefficients_thread\(void \* arg\)"),
    "sph_accel_thread":    ("src/SphericalBasis.cc", 1476, 1660, r"^void \* SphericalBasis::determine_acceleration_and_potential_thread\(void \* arg\)"),
    "sph_compute_multistep_coefficients": ("src/SphericalBasis.cc", 1231, 1474, r"^void SphericalBasis::compute_multistep_coefficients\(\)"),
    "sph_multistep_update": ("src/SphericalBasis.cc", 1156, 1228, r"^void SphericalBasis::multistep_update\(int from, int to, Component \*c, int i, int id\)"),
    "sph_get_pot_coefs_safe": ("src/SphericalBasis.cc", 1797, 1813, r"^void SphericalBasis::get_pot_coefs_safe\("),
    # --- EmpCylSL (exputil/EmpCylSL.cc)
    "cyl_accumulate":      ("exputil/EmpCylSL.cc", 4049, 4146, r"^void EmpCylSL::accumulate\(double r, double z, double phi, double mass,"),
    "cyl_accumulated_eval": ("exputil/EmpCylSL.cc", 5256, 5410, r"^void EmpCylSL::accumulated_eval\(double r, double z, double phi,"),
    "cyl_get_pot":         ("exputil/EmpCylSL.cc", 5557, 5632, r"^void EmpCylSL::get_pot\(Eigen::MatrixXd& Vc, Eigen::MatrixXd& Vs,"),
    "cyl_r_to_xi":         ("exputil/EmpCylSL.cc", 6446, 6463, r"^double EmpCylSL::r_to_xi\(double r\)"),
    "cyl_z_to_y":          ("exputil/EmpCylSL.cc", 7109, 7117, r"^double EmpCylSL::z_to_y\(double z\)"),
    # --- Cylinder (src/Cylinder.cc): the thread functions around EmpCylSL (orientation, rcylmax gate,
    # erf blend to the monopole)
    "cylinder_coef_thread": ("src/Cylinder.cc", 748, 896, r"^void \* Cylinder::determine_coefficients_thread\(void \* arg\)"),
    "cylinder_multistep_update": ("src/Cylinder.cc", 1752, 1773, r"^void Cylinder::multistep_update\(int from, int to, Component\* c, int i, int id\)"),
    "cylexp_multistep_update": ("src/CylEXP.cc", 159, 188, r"^void CylEXP::multistep_update\(int from, int to, double r, double z, double phi, double mass, int id\)"),
    "cylexp_compute_multistep_coefficients": ("src/CylEXP.cc", 192, 282, r"^void CylEXP::compute_multistep_coefficients\(unsigned mlevel\)"),
    "cylinder_accel_thread": ("src/Cylinder.cc", 1266, 1446, r"^void \* Cylinder::determine_acceleration_and_potential_thread\(void \* arg\)"),
    # --- centre-of-mass sums (src/Component.cc): thrd_pass_posn + fix_positions_thread
    "fix_positions_thread": ("src/Component.cc", 3266, 3352, r"^struct thrd_pass_posn"),
    # --- pyEXP twin (expui/BiorthBasis.cc): BasisClasses::Spherical accumulate / field evaluation.
    # accumulate is cut around its pcavar block (lines 652-659: Eigen expression templates on complex
    # vectors, the covariance of pyEXP's own subsampling, not on the path)
    "biorth_factorial":    ("expui/BiorthBasis.cc", 321, 329, r"^    factorial.resize\(lmax\+1, lmax\+1\);"),
    "biorth_accumulate_head": ("expui/BiorthBasis.cc", 583, 651, r"^  void Spherical::accumulate\(double x, double y, double z, double mass,"),
    "biorth_accumulate_tail": ("expui/BiorthBasis.cc", 660, 665, r"^$"),
    "biorth_sph_eval":     ("expui/BiorthBasis.cc", 711, 816, r"^  std::vector<double>"),
    "biorth_compute_accel": ("expui/BiorthBasis.cc", 818, 926, r"^  void Spherical::computeAccel\(double x, double y, double z,"),
    "biorth_cyl_eval":     ("expui/BiorthBasis.cc", 929, 942, r"^  std::vector<double>"),
    "biorth_crt_eval":     ("expui/BiorthBasis.cc", 945, 958, r"^  std::vector<double>"),
    "biorth_legendre":     ("expui/BiorthBasis.cc", 1034, 1122, r"^  void BiorthBasis::legendre_R\(int lmax, double x, Eigen::MatrixXd& p\)"),
    # --- leapfrog (src/incpos.cc, src/incvel.cc)
    "incpos_thread":       ("src/incpos.cc", 15, 70, r"^void \* incr_position_thread\(void \*ptr\)"),
    "incvel_thread":       ("src/incvel.cc", 15, 89, r"^void \* incr_velocity_thread\(void \*ptr\)"),
    # --- multistep level assignment (src/multistep.cc)
    "multistep_level_thread": ("src/multistep.cc", 52, 236, r"^void \* adjust_multistep_level_thread\(void \*ptr\)"),
}


def main(reference: str, outdir: str) -> int:
    os.makedirs(outdir, exist_ok=True)
    for name, (rel, a, b, rx) in PIECES.items():
        path = os.path.join(reference, rel)
        with open(path, "r", encoding="utf-8", errors="replace") as f:
            lines = f.readlines()
        if not re.search(rx, lines[a - 1]):
            sys.stderr.write(f"ref_extract: {rel}:{a} does not match /{rx}/ (got {lines[a - 1]!r}); "
                             f"different reference revision?\n")
            return 1
        with open(os.path.join(outdir, name + ".inc"), "w") as g:
            g.write(f"// generated from {rel}:{a}-{b} of the reference tree; not committed\n")
            g.write(f'#line {a} "{path}"\n')
            g.writelines(lines[a - 1:b])
    print(f"ref_extract: {len(PIECES)} pieces -> {outdir}")
    return 0


if __name__ == "__main__":
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "gen")
    sys.exit(main(ref, out))
