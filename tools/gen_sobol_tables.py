#!/usr/bin/env python
"""Derive the Sobol' tables used by the global Sobol sampler and write them as one binary blob.

The reference (src/pathtracer/sobolmatrices.rs:5-7, 53463, 54155) ships three tables:
  * SOBOL_MATRICES_32[1024*52]   - 32-bit generator-matrix columns, 52 columns per dimension
  * VD_C_SOBOL_MATRICES[25][..]  - per log2-resolution m: pixel-bit deltas caused by the frame bits
  * VD_C_SOBOL_MATRICES_INV[26][..] - per m: inverse of the (index low 2m bits -> pixel bits) map
They are not hand-made constants: the first is the Joe & Kuo (2008) "new-joe-kuo-6.21201" direction
numbers evaluated at 52 bits and truncated to the top 32 bits, the other two follow from the first
two dimensions by GF(2) linear algebra (Gruenschloss, "sobol.h" sample enumeration).  This script
re-derives all three from the Joe-Kuo numbers bundled with scipy (scipy.stats._sobol) and, when the
reference checkout is present, asserts bit-equality with its tables.

Output: pathtracer_rs_b200/data/sobol_tables.bin
  header: magic 'SOBL', u32 n_dims, u32 n_cols, u32 n_m (25), u32 n_minv (26)
  u32 matrices[n_dims*n_cols]
  for m in 1..=25: u32 len, u64[len]     (VdC)
  for m in 1..=26: u32 len, u64[len]     (VdC inverse; entry 0 is the m=1 table)
"""
import os, re, struct, sys
import numpy as np

NDIM, NCOL = 1024, 52
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "pathtracer_rs_b200", "data", "sobol_tables.bin")
REF = "/root/reference/src/pathtracer/sobolmatrices.rs"


def joe_kuo_52():
    import scipy.stats._sobol as s
    v = np.zeros((NDIM, NCOL), dtype=np.uint64)
    s._initialize_v(v, NDIM, NCOL)
    return v  # v[d, j] = 52-bit column j of dimension d


def gf2_inverse(cols, n):
    """cols[j] = image (n-bit int) of unit vector j. Returns columns of the inverse map."""
    a = list(cols)
    inv = [1 << j for j in range(n)]  # track: a[j] = A * inv[j]
    # Gaussian elimination on the set of (image, preimage) pairs
    pairs = list(zip(a, inv))
    basis = [None] * n  # basis[b] = (image with lowest set bit b ... ) reduce to unit vectors
    for bit in range(n):
        piv = None
        for k, (im, pre) in enumerate(pairs):
            if im >> bit & 1 and all((im >> lb & 1) == 0 for lb in range(bit)):
                piv = k
                break
        if piv is None:
            raise ValueError("singular")
        pim, ppre = pairs[piv]
        for k, (im, pre) in enumerate(pairs):
            if k != piv and im >> bit & 1:
                pairs[k] = (im ^ pim, pre ^ ppre)
    out = [0] * n
    for im, pre in pairs:
        assert im & (im - 1) == 0 and im != 0
        out[im.bit_length() - 1] = pre
    return out


def derive():
    v = joe_kuo_52()
    mats32 = (v >> np.uint64(20)).astype(np.uint32)
    c0 = [int(x) for x in v[0]]
    c1 = [int(x) for x in v[1]]
    vdc, vdc_inv = [], []
    for m in range(1, 27):
        def col(j):
            return ((c0[j] >> (52 - m)) << m) | (c1[j] >> (52 - m))
        a_cols = [col(j) for j in range(2 * m)]
        vdc_inv.append(gf2_inverse(a_cols, 2 * m))
        if m <= 25:
            vdc.append([col(j) for j in range(2 * m, 52)])
    return mats32, vdc, vdc_inv


def parse_reference():
    txt = open(REF).read()
    i0 = txt.index("SOBOL_MATRICES_32")
    body = txt[txt.index("= [", i0) + 3: txt.index("];", i0)]
    mats = np.array([int(x.replace("_", ""), 16) for x in re.findall(r"0x[0-9a-fA-F_]+", body)],
                    dtype=np.uint64).reshape(NDIM, NCOL).astype(np.uint32)
    consts = {}
    for mm in re.finditer(r"^const (MI?\d+): \[u64; (\d+)\] = \[(.*?)\];", txt, re.S | re.M):
        consts[mm.group(1)] = [int(x[:-4].replace("_", ""), 16)
                               for x in re.findall(r"0x[0-9a-fA-F_]+?_u64", mm.group(3))]
        assert len(consts[mm.group(1)]) == int(mm.group(2))
    def table(name):
        i = txt.index("pub const " + name)
        j = txt.index("= [", i) + 3
        body = txt[j: txt.index("];", j)]
        return [consts[n] for n in re.findall(r"&(MI?\d+)", body)]
    return mats, table("VD_C_SOBOL_MATRICES:"), table("VD_C_SOBOL_MATRICES_INV:")


def main():
    mats32, vdc, vdc_inv = derive()
    if os.path.exists(REF):
        rm, rv, ri = parse_reference()
        assert np.array_equal(rm, mats32), "SOBOL_MATRICES_32 mismatch"
        assert len(rv) == 25 and len(ri) == 26
        for m in range(25):
            assert rv[m] == vdc[m], f"VdC m={m+1} mismatch"
        for m in range(26):
            assert ri[m] == vdc_inv[m], f"VdC inv m={m+1} mismatch"
        print("derived tables are bit-identical to the reference's")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "wb") as f:
        f.write(b"SOBL" + struct.pack("<4I", NDIM, NCOL, 25, 26))
        f.write(mats32.astype("<u4").tobytes())
        for t in vdc + vdc_inv:
            f.write(struct.pack("<I", len(t)))
            f.write(np.array(t, dtype="<u8").tobytes())
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
