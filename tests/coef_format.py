"""TEST INFRASTRUCTURE: Python restatement of the reference's native coefficient streams, used
to check the host driver's dump_coefs byte for byte (header fields parsed, payload compared).
  sphere   src/SphericalBasis.cc:1827-1904 (cmagic 0xc0a57a2, src/SphericalBasis.H:368)
  cylinder exputil/EmpCylSL.cc:5868-5919   (cmagic 0xc0a57a3, include/EmpCylSL.H:222)
and of the complex packing of dump_coefs_h5 (src/SphericalBasis.cc:1925-1937,
src/Cylinder.cc:1640-1656)."""
import struct

import numpy as np
import yaml


def parse_header(buf):
    magic, hsize = struct.unpack_from("<II", buf, 0)
    node = yaml.safe_load(buf[8:8 + hsize].decode())
    return magic, node, 8 + hsize


def sphere_payload(expcoef, lmax, nmax):
    """expcoef[(lmax+1)^2][nmax] -> the doubles after the header, n outermost"""
    out = []
    for ir in range(nmax):
        loffset = 0
        for l in range(lmax + 1):
            moffset = 0
            for m in range(l + 1):
                if m == 0:
                    out.append(expcoef[loffset + moffset, ir]); moffset += 1
                else:
                    out.append(expcoef[loffset + moffset, ir]); out.append(expcoef[loffset + moffset + 1, ir])
                    moffset += 2
            loffset += 2 * l + 1
    return np.array(out)


def sphere_old_stream(expcoef, ident, tnow, scale, lmax, nmax):
    b = ident.encode()[:64].ljust(64, b"\0") + struct.pack("<ddii", tnow, scale, nmax, lmax)
    return b + np.ascontiguousarray(expcoef.T).tobytes()


def sphere_struct(expcoef, lmax, nmax):
    ldim = (lmax + 1) * (lmax + 2) // 2
    cof = np.zeros((ldim, nmax), dtype=np.complex128)
    L = offset = 0
    for l in range(lmax + 1):
        for m in range(l + 1):
            if m == 0:
                cof[L] = expcoef[offset]; offset += 1
            else:
                cof[L] = expcoef[offset] + 1j * expcoef[offset + 1]; offset += 2
            L += 1
    return cof


def cylinder_payload(cosv, sinv):
    out = []
    for mm in range(cosv.shape[0]):
        out.append(cosv[mm])
        if mm:
            out.append(sinv[mm])
    return np.concatenate(out)
