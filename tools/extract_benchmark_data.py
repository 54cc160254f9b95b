#!/usr/bin/env python
"""Extract the INPUT DATA of the reference's water benchmarks into small
fixtures the workload synthesiser can use on a box without /root/reference:

  benchmarks/QS/H2O-{32,64,128,256,512,1024}.inp  &COORD / &CELL ABC  ->  cp2k_b200/data/h2o_systems.npz
  benchmarks/QS/H2O-64_nonortho.inp               (+ ALPHA_BETA_GAMMA)
  data/GTH_BASIS_SETS  (TZV2P-GTH for H, O)        ->  cp2k_b200/data/basis_sets.json
  data/BASIS_MOLOPT    (DZVP-MOLOPT-SR-GTH for H, O)

Only numbers (coordinates, cell edges, exponents, contraction coefficients)
are stored; run in the build container:  python tools/extract_benchmark_data.py
"""
import json
import os
import re

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cp2k_b200", "data")


def read_system(path):
    elems, xyz, abc, angles = [], [], None, None
    in_coord = False
    for line in open(path):
        s = line.strip()
        if s.startswith("ABC"):
            abc = [float(x) for x in s.split()[1:4]]
        elif s.startswith("ALPHA_BETA_GAMMA"):
            angles = [float(x) for x in s.split()[1:4]]
        elif s.startswith("&COORD"):
            in_coord = True
        elif s.startswith("&END COORD"):
            in_coord = False
        elif in_coord and s and not s.startswith("#"):
            p = s.split()
            elems.append(p[0])
            xyz.append([float(x) for x in p[1:4]])
    return elems, np.array(xyz), np.array(abc), (np.array(angles) if angles else np.array([90.0, 90.0, 90.0]))


def read_basis(path, element, name):
    lines = [l.rstrip("\n") for l in open(path)]
    for i, l in enumerate(lines):
        p = l.split()
        if len(p) >= 2 and p[0] == element and name in p[1:]:
            break
    else:
        raise KeyError((element, name))
    i += 1
    nset = int(lines[i].split()[0])
    i += 1
    sets = []
    for _ in range(nset):
        head = [int(x) for x in lines[i].split()]
        i += 1
        _, lmin, lmax, npgf = head[:4]
        nshell = head[4:4 + (lmax - lmin + 1)]
        zet, coef = [], []
        for _ in range(npgf):
            vals = [float(x) for x in lines[i].split()]
            i += 1
            zet.append(vals[0])
            coef.append(vals[1:1 + sum(nshell)])
        sets.append({"lmin": lmin, "lmax": lmax, "nshell": nshell, "zet": zet, "coef": coef})
    return sets


def main():
    os.makedirs(OUT, exist_ok=True)
    systems = {}
    for tag in ("H2O-32", "H2O-64", "H2O-128", "H2O-256", "H2O-512", "H2O-1024", "H2O-64_nonortho"):
        elems, xyz, abc, ang = read_system(os.path.join(REF, "benchmarks", "QS", tag + ".inp"))
        key = re.sub(r"[^A-Za-z0-9]", "_", tag)
        systems[key + "__xyz_angstrom"] = xyz
        systems[key + "__is_oxygen"] = np.array([e == "O" for e in elems])
        systems[key + "__abc_angstrom"] = abc
        systems[key + "__alpha_beta_gamma"] = ang
        print(tag, len(elems), "atoms", abc, ang)
    np.savez_compressed(os.path.join(OUT, "h2o_systems.npz"), **systems)
    basis = {}
    for el in ("H", "O"):
        basis[f"{el}:TZV2P-GTH"] = read_basis(os.path.join(REF, "data", "GTH_BASIS_SETS"), el, "TZV2P-GTH")
        basis[f"{el}:DZVP-MOLOPT-SR-GTH"] = read_basis(os.path.join(REF, "data", "BASIS_MOLOPT"), el,
                                                       "DZVP-MOLOPT-SR-GTH")
    json.dump(basis, open(os.path.join(OUT, "basis_sets.json"), "w"), indent=1)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
