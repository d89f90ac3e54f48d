#!/usr/bin/env python
"""Extract the reference's golden tables for the CoreRT hot path into a small
JSON fixture (tests/golden/reference_tables.json).

Run in the build container only (needs /root/reference):
    python tests/golden/extract_golden.py

Sources (numeric tables only; no reference code is copied):
  test/benchmarks/natraj_trues.jl                      Natraj et al. (2009) Rayleigh I/Q/U, tau=0.5, mu0=0.2
  test/benchmarks/6SV1_R_trues.jl                      6SV1 scalar Rayleigh reflectances (6 cases)
  test/vlidort_baseline/reference_data/siewert2000_IIA_greek.jl   Siewert (2000) problem IIA Greek constants
  test/vlidort_baseline/reference_data/siewert2000_IIA_truth.jl   VLIDORT 2.8.3 tables 2-9
  test/vlidort_baseline/reference_data/solar_tester_{atmosphere,truth}.jl   VLIDORT 2.8.3 solar_tester (23 layers,
                                                        Rayleigh + gas + HG aerosol), task 1, TOA up / BOA down
  test/benchmarks/natraj_cds/CDS/{I,Q,U}_{UP,DN}_TAU_*   Natraj et al. (2009) improved Coulson-Dave-Sekera tables
                                                        (a subset -> tests/golden/natraj_cds_subset.json)
"""
import json
import os
import re

REF = os.environ.get("VSMOM_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_tables.json")

_num = r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?"


def julia_matrix(text: str, name: str):
    """Parse `name = [ a b c ; d e f ; ... ]` (rows separated by ';')."""
    mt = re.search(r"(?:const\s+)?" + re.escape(name) + r"\s*=\s*\[(.*?)\]", text, re.S)
    assert mt, name
    rows = []
    for row in mt.group(1).split(";"):
        vals = [float(v) for v in re.findall(_num, row)]
        if vals:
            rows.append(vals)
    return rows


def julia_vector(text: str, name: str):
    mt = re.search(re.escape(name) + r"\s*=\s*\[(.*?)\]", text, re.S)
    assert mt, name
    return [float(v) for v in re.findall(_num, mt.group(1))]


def nested_brackets(text: str):
    """Parse a Julia nested-array literal of numbers into nested Python lists."""
    text = re.sub(r"#.*", "", text)
    s = text[text.index("["):text.rindex("]") + 1]
    s = re.sub(r"\s+", " ", s)
    s = re.sub(r",\s*\]", "]", s)
    return json.loads(s)


CDS_TAUS = ["0.1", "0.5", "1"]
CDS_ALBEDOS = [0.0, 0.25, 0.8]
CDS_MU0S = [0.2, 0.6, 1.0]


def parse_cds(path):
    """{(albedo, mu0, mu): [7 azimuth values]} from one CDS table file (format: 00README)."""
    tab, alb = {}, None
    for ln in open(path):
        mt = re.match(r"\s*albedo\s*=\s*(" + _num + ")", ln)
        if mt:
            alb = float(mt.group(1))
            continue
        parts = ln.split()
        if len(parts) == 9 and alb is not None:
            try:
                v = [float(x) for x in parts]
            except ValueError:
                continue
            tab[(alb, v[0], v[1])] = v[2:]
    return tab


def extract_cds():
    """Subset of the CDS tables used by test/benchmarks/benchmark_natraj_cds.jl: 3 optical depths x 3 albedos
    x 3 solar cosines, all 16 mu x 7 phi, I/Q/U, upwelling (TOA) and downwelling (BOA)."""
    d = os.path.join(REF, "test/benchmarks/natraj_cds/CDS")
    mus = [0.02, 0.06, 0.1, 0.16, 0.2, 0.28, 0.32, 0.4, 0.52, 0.64, 0.72, 0.84, 0.92, 0.96, 0.98, 1.0]
    out = {"tau": [float(t) for t in CDS_TAUS], "albedo": CDS_ALBEDOS, "mu0": CDS_MU0S, "mu": mus,
           "phi": [0.0, 30.0, 60.0, 90.0, 120.0, 150.0, 180.0], "nstreams": 11, "depol": 0.0, "tables": {}}
    for st in "IQU":
        for di in ("UP", "DN"):
            arr = []
            for t in CDS_TAUS:
                tab = parse_cds(os.path.join(d, f"{st}_{di}_TAU_{t}"))
                arr.append([[[tab[(a, m0, mu)] for mu in mus] for m0 in CDS_MU0S] for a in CDS_ALBEDOS])
            out["tables"][f"{st}_{di}"] = arr          # [tau][albedo][mu0][mu][phi]
    path = os.path.join(os.path.dirname(OUT), "natraj_cds_subset.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


def extract_solar_tester():
    """VLIDORT solar_tester scalar case as used by test/vlidort_baseline/cases/case_B_solar_tester.jl: the
    23-layer atmosphere and the task-1 (plane-parallel, no delta-M) Stokes-I at TOA (upwelling) and BOA
    (downwelling) for all 36 geometries."""
    d = os.path.join(REF, "test/vlidort_baseline/reference_data")
    ta = open(os.path.join(d, "solar_tester_atmosphere.jl")).read()
    tt = open(os.path.join(d, "solar_tester_truth.jl")).read()
    vals = {}
    for mt in re.finditer(r"^\s*(" + _num + r"),?\s*#\s*geom=(\d+)\s+level=(\d+)\s+dir=(\d+)\s+task=(\d+)", tt, re.M):
        v, g, lv, di, tk = float(mt.group(1)), *(int(mt.group(k)) for k in range(2, 6))
        vals[(g, lv, di, tk)] = v
    levels = julia_vector(tt, "SOLAR_TESTER_TAU_LEVELS")
    return {
        "height_km": julia_vector(ta, "SOLAR_TESTER_HEIGHT_KM"), "molext": julia_vector(ta, "SOLAR_TESTER_MOLEXT"),
        "molomg": julia_vector(ta, "SOLAR_TESTER_MOLOMG"),
        "sza": julia_vector(tt, "SOLAR_TESTER_SZA_DEG"), "vza": julia_vector(tt, "SOLAR_TESTER_VZA_DEG"),
        "raz": julia_vector(tt, "SOLAR_TESTER_RAZ_DEG"), "albedo": 0.05, "depol": 0.01072, "nstreams": 8,
        "aerosol": {"g": 0.8, "omega": 0.95, "tau_total": 0.5, "nmoments": 15, "layers": 6},   # case_B_solar_tester.jl:19-22,44-54
        "toa_up": [vals[(g, 1, 1, 1)] for g in range(1, 37)],               # geom = (isza*3 + ivza)*3 + iraz
        "boa_dn": [vals[(g, len(levels), 2, 1)] for g in range(1, 37)],
        "rtol": 1e-3,                                                       # case_B_solar_tester.jl:166-170
    }


def extract_solar_tester_vector():
    """VLIDORT solar_tester, NSTOKES = 3 with the Problem III Mie aerosol, as used by
    test/vlidort_baseline/cases/case_C_solar_tester_vector.jl: first 16 expansion moments of the F-matrix and the
    task-1 I/Q/U at TOA (up) and BOA (down) for all 36 geometries."""
    d = os.path.join(REF, "test/vlidort_baseline/reference_data")
    ta = open(os.path.join(d, "solar_tester_problemIII_aerosol.jl")).read()
    tt = open(os.path.join(d, "solar_tester_vector_truth.jl")).read()
    out = {"moms": {k: julia_vector(ta, "PROBLEMIII_" + k)[:16] for k in ("a1", "b1", "a2", "a3", "b2", "a4")},
           "omega": 0.99999, "tau_total": 0.5, "layers": 6, "nstreams": 8, "albedo": 0.05,
           "sza": julia_vector(tt, "SOLAR_TESTER_VECTOR_SZA_DEG"), "vza": julia_vector(tt, "SOLAR_TESTER_VECTOR_VZA_DEG"),
           "raz": julia_vector(tt, "SOLAR_TESTER_VECTOR_RAZ_DEG"), "rtol_up": 1e-3, "rtol_dn": 2e-3}
    for st in "IQU":
        a = tt.index(f"const SOLAR_TESTER_VECTOR_{st} = reshape(")
        b = tt.index("], ", a) if "], " in tt[a:a + 400000] else len(tt)
        blk = tt[a:tt.index("const SOLAR_TESTER_VECTOR_", a + 10)]
        vals = {}
        for mt in re.finditer(r"^\s*(" + _num + r"),?\s*#\s*geom=(\d+)\s+level=(\d+)\s+dir=(\d+)\s+task=(\d+)", blk, re.M):
            vals[tuple(int(mt.group(k)) for k in range(2, 6))] = float(mt.group(1))
        out[st + "_toa_up"] = [vals[(g, 1, 1, 1)] for g in range(1, 37)]
        out[st + "_boa_dn"] = [vals[(g, 5, 2, 1)] for g in range(1, 37)]
    return out


def main():
    extract_cds()
    out = {"solar_tester": extract_solar_tester(), "solar_tester_vector": extract_solar_tester_vector()}
    t = open(os.path.join(REF, "test/benchmarks/natraj_trues.jl")).read()
    out["natraj"] = {
        "mu": [0.02, 0.06, 0.10, 0.16, 0.20, 0.28, 0.32, 0.40, 0.52, 0.64, 0.72,
               0.84, 0.92, 0.96, 0.98, 1.00],                       # test_CoreRT.jl:119
        "phi": [0.0, 30.0, 60.0, 90.0, 120.0, 150.0, 180.0],        # test_CoreRT.jl:120
        "tau": 0.5, "mu0": 0.2, "nstreams": 11, "depol": 0.0, "albedo": 0.0,
        "I": julia_matrix(t, "I_trues"),
        "Q": julia_matrix(t, "Q_trues"),
        "U": julia_matrix(t, "U_trues"),
        "tol": {"I": 5e-4, "Q": 2.5e-3, "U": 5e-4},                 # test_CoreRT.jl:153-155
    }
    t = open(os.path.join(REF, "test/benchmarks/6SV1_R_trues.jl")).read()
    out["sixsv1"] = {
        "R_trues": nested_brackets(t),
        # test_CoreRT.jl:36-41: (azs, szas, lambda_nm, tau, rho)
        "cases": [
            [[180, 90, 0], [23.0739, 53.1301, 78.4630], 530, 0.1, 0.0],
            [[180, 90, 0], [0.0001, 36.8699, 66.4218], 530, 0.1, 0.25],
            [[180, 90, 0], [0.0001, 36.8699, 66.4218], 440, 0.25, 0.0],
            [[180, 90, 0], [23.0739, 53.1301, 78.4630], 440, 0.25, 0.25],
            [[180, 90, 0], [23.0739, 53.1301, 78.4630], 360, 0.50, 0.0],
            [[180, 90, 0], [0.0001, 36.8699, 66.4218], 360, 0.50, 0.25],
        ],
        # test/benchmarks/6SV1_1.yaml
        "vza": [0.0, 11.4783, 16.2602, 23.0739, 32.8599, 43.9455, 50.2082, 58.6677,
                66.4218, 71.3371, 73.7398, 78.463, 80.7931, 84.2608, 86.5602, 88.854],
        "nstreams": 11, "depol": 0.0, "tol": 0.006,
    }
    t = open(os.path.join(REF, "test/vlidort_baseline/reference_data/siewert2000_IIA_greek.jl")).read()
    rows = {k: julia_vector(t, "row%d" % k) for k in range(1, 7)}
    t = open(os.path.join(REF, "test/vlidort_baseline/reference_data/siewert2000_IIA_truth.jl")).read()
    out["siewert"] = {
        # mapping documented in siewert2000_IIA_greek.jl header: γ and ϵ sign-flipped
        "alpha": rows[1], "beta": rows[2], "gamma": [-v for v in rows[3]],
        "delta": rows[4], "epsilon": [-v for v in rows[5]], "zeta": rows[6],
        "tau": 1.0, "omega": 0.973527, "sza": 53.130102354065, "nstreams": 16,
        "vza": [0.0001, 25.841932763, 36.869897646, 45.572995999, 53.130102354, 60.0,
                66.421821522, 72.542396876, 78.463040967, 84.260829523, 89.9999],
        "cosines": julia_vector(t, "SIEWERT_TABLE2_COSINES"),
        "tables": {str(k): julia_matrix(t, "SIEWERT_TABLE%d" % k) for k in range(2, 10)},
        # (az, stokes index) -> table number   (siewert2000_IIA_truth.jl:233-242)
        "map": [[0.0, 0, 2], [0.0, 1, 3], [90.0, 0, 4], [90.0, 1, 5], [90.0, 2, 6],
                [90.0, 3, 7], [180.0, 0, 8], [180.0, 1, 9]],
    }
    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
