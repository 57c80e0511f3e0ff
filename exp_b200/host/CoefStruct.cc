// CoefStruct.cc -- see CoefStruct.H, and the dump_coefs members of the two forces.
#include "CoefStruct.H"

#include <cstdio>
#include <cstring>
#include <ostream>

#include "Cylinder.H"
#include "SphericalBasis.H"
#include "global.H"

namespace exphost {

std::string yaml_kv(const std::string &key, const std::string &v) { return key + ": " + v + "\n"; }
std::string yaml_kv(const std::string &key, double v)
{
  char b[64];
  std::snprintf(b, sizeof b, "%.17g", v);
  return key + ": " + b + "\n";
}
std::string yaml_kv(const std::string &key, int v) { return key + ": " + std::to_string(v) + "\n"; }
std::string yaml_kv(const std::string &key, bool v) { return key + ": " + (v ? "true" : "false") + "\n"; }

namespace {
void write_header(std::ostream &out, unsigned int magic, std::string y)
{
  if (!y.empty() && y.back() == '\n') y.pop_back();       // the emitter's string has no trailing newline
  const unsigned int hsize = (unsigned int)y.size();
  out.write(reinterpret_cast<const char *>(&magic), sizeof(unsigned int));
  out.write(reinterpret_cast<const char *>(&hsize), sizeof(unsigned int));
  out.write(y.data(), hsize);
}
} // namespace

// src/SphericalBasis.cc:1827-1904.  NewCoefs: magic 0xc0a57a2 (src/SphericalBasis.H:368), the
// YAML header {id, time, scale, nmax, lmax, normed}, then for every n the (l, m) terms: one
// double for m = 0, (cos, sin) for m > 0.  Old style: 64-byte id, tnow, scale, nmax, Lmax, then
// for every n the (Lmax+1)^2 rows.
void SphericalBasis::dump_coefs(std::ostream &out, bool NewCoefs)
{
  const std::vector<double> expcoef = get_coefs();           // [row][n]
  if (NewCoefs) {
    write_header(out, 0xc0a57a2u,
                 yaml_kv("id", id) + yaml_kv("time", tnow) + yaml_kv("scale", scale) + yaml_kv("nmax", nmax) +
                 yaml_kv("lmax", Lmax) + yaml_kv("normed", true));
    for (int ir = 0; ir < nmax; ir++)
      for (int l = 0, loffset = 0; l <= Lmax; loffset += 2 * l + 1, l++)
        for (int m = 0, moffset = 0; m <= l; m++) {
          const int nterm = m == 0 ? 1 : 2;
          for (int k = 0; k < nterm; k++) {
            const double z = expcoef[(size_t)(loffset + moffset + k) * nmax + ir];
            out.write(reinterpret_cast<const char *>(&z), sizeof(double));
          }
          moffset += nterm;
        }
  } else {
    char buf[64];
    for (int i = 0; i < 64; i++) buf[i] = i < (int)id.size() ? id[i] : '\0';
    out.write(buf, 64);
    const double t = tnow, sc = scale;
    out.write(reinterpret_cast<const char *>(&t), sizeof(double));
    out.write(reinterpret_cast<const char *>(&sc), sizeof(double));
    out.write(reinterpret_cast<const char *>(&nmax), sizeof(int));
    out.write(reinterpret_cast<const char *>(&Lmax), sizeof(int));
    for (int ir = 0; ir < nmax; ir++)
      for (int l = 0; l <= Lmax * (Lmax + 2); l++)
        out.write(reinterpret_cast<const char *>(&expcoef[(size_t)l * nmax + ir]), sizeof(double));
  }
}

// the SphStruct that dump_coefs_h5 fills (src/SphericalBasis.cc:1909-1945)
SphStruct SphericalBasis::coef_struct()
{
  const std::vector<double> expcoef = get_coefs();
  SphStruct cur;
  cur.time = tnow; cur.geom = "sphere"; cur.id = id;
  cur.lmax = Lmax; cur.nmax = nmax; cur.scale = scale; cur.normed = true;
  cur.allocate();
  for (int ir = 0; ir < nmax; ir++)
    for (int l = 0, L = 0, offset = 0; l <= Lmax; l++)
      for (int m = 0; m <= l; m++, L++) {
        if (m == 0) {
          cur(L, ir) = {expcoef[(size_t)offset * nmax + ir], 0.0};
          offset += 1;
        } else {
          cur(L, ir) = {expcoef[(size_t)offset * nmax + ir], expcoef[(size_t)(offset + 1) * nmax + ir]};
          offset += 2;
        }
      }
  if (component)
    for (int k = 0; k < 3; k++) cur.ctr[k] = component->com0[k] + component->center[k];
  return cur;
}

// src/Cylinder.cc:1618-1622 -> EmpCylSL::dump_coefs_binary (exputil/EmpCylSL.cc:5868-5919):
// magic 0xc0a57a3 (include/EmpCylSL.H:222), YAML {time, mmax, nmax}, then per m the cosine row
// and, for m > 0, the sine row.
void Cylinder::dump_coefs(std::ostream &out)
{
  std::vector<double> cosv, sinv;
  get_coefs(cosv, sinv);
  write_header(out, 0xc0a57a3u, yaml_kv("time", tnow) + yaml_kv("mmax", mmax) + yaml_kv("nmax", ncylorder));
  for (int mm = 0; mm <= mmax; mm++) {
    out.write(reinterpret_cast<const char *>(&cosv[(size_t)mm * ncylorder]), sizeof(double) * ncylorder);
    if (mm) out.write(reinterpret_cast<const char *>(&sinv[(size_t)mm * ncylorder]), sizeof(double) * ncylorder);
  }
}

// the CylStruct that dump_coefs_h5 fills (src/Cylinder.cc:1625-1656)
CylStruct Cylinder::coef_struct()
{
  std::vector<double> cosv, sinv;
  get_coefs(cosv, sinv);
  CylStruct cur;
  cur.time = tnow; cur.geom = "cylinder"; cur.id = id;
  cur.mmax = mmax; cur.nmax = ncylorder;
  cur.allocate();
  for (int m = 0; m <= mmax; m++)
    for (int ir = 0; ir < ncylorder; ir++)
      cur(m, ir) = {cosv[(size_t)m * ncylorder + ir], m ? sinv[(size_t)m * ncylorder + ir] : 0.0};
  if (component)
    for (int k = 0; k < 3; k++) cur.ctr[k] = component->com0[k] + component->center[k];
  return cur;
}

} // namespace exphost
