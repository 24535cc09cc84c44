#!/usr/bin/env python3
"""Convert the reference's ELM (Fortran) golden dumps (test_CanHydro, test_CanFlux, test_SurfAlb); first: test_CanHydro - test/data/CanopyHydrology_{IN,OUT}.txt,
the fixture of BASELINE.json config 1 - into tests/golden/elm_canopy_hydrology.npz.  The 48 records (NSTEP 1..48,
the range test/test_CanHydro.cc:154 loops over) become 48 columns.  Data only; format described in
src/utils/read_test_input.cc:14-23 ("NSTEP n" / "name v0 v1 ..." / "!!! n")."""
import pathlib, sys
import numpy as np

R = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
ROOT = pathlib.Path(__file__).resolve().parent.parent


def parse(path):
    recs, cur = {}, None
    for line in open(path):
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "NSTEP":
            cur = recs.setdefault(int(tok[1]), {})
        elif tok[0] == "!!!":
            cur = None
        elif cur is not None:
            cur[tok[0]] = np.array([float(x) for x in tok[1:]])
    return recs


def stack(recs, steps, name):
    return np.stack([recs[s][name] for s in steps])


def main():
    i = parse(R / "test/data/CanopyHydrology_IN.txt")
    o = parse(R / "test/data/CanopyHydrology_OUT.txt")
    steps = list(range(1, 49))
    out = {"steps": np.array(steps)}
    for name in i[1]:
        out["in_" + name] = stack(i, steps, name)
    for name in o[1]:
        out["out_" + name] = stack(o, steps, name)
    dst = ROOT / "tests/golden/elm_canopy_hydrology.npz"
    np.savez_compressed(dst, **out)
    print(dst, dst.stat().st_size, "bytes;", len(steps), "records;",
          "oldfflag", np.unique(out["in_oldfflag"]), "dewmx", np.unique(out["in_dewmx"]),
          "snl", np.unique(out["in_snl"]), "h2osfc max", out["in_h2osfc"].max())


def canopy_fluxes():
    """test/data/CanopyFluxes_{IN,OUT}.txt, records 0..96 (test/test_CanFlux.cc:328)."""
    i = parse(R / "test/data/CanopyFluxes_IN.txt")
    o = parse(R / "test/data/CanopyFluxes_OUT.txt")
    steps = list(range(0, 97))
    out = {"steps": np.array(steps)}
    for name in i[0]:
        out["in_" + name] = stack(i, steps, name)
    for name in o[0]:
        out["out_" + name] = stack(o, steps, name)
    dst = ROOT / "tests/golden/elm_canopy_fluxes.npz"
    np.savez_compressed(dst, **out)
    print(dst, dst.stat().st_size, "bytes;", len(steps), "records; dayl values", len(np.unique(out["in_dayl"])),
          "night records", int((out["in_parsun_z"][:, 0] <= 0).sum()))


def surface_albedo():
    """test/data/SurfaceAlbedo_{IN,OUT}.txt, records 2..96 (test/test_SurfAlb.cc:391 loops over 2..48)."""
    i = parse(R / "test/data/SurfaceAlbedo_IN.txt")
    o = parse(R / "test/data/SurfaceAlbedo_OUT.txt")
    steps = sorted(set(i) & set(o))
    out = {"steps": np.array(steps)}
    for name in i[steps[0]]:
        out["in_" + name] = stack(i, steps, name)
    for name in o[steps[0]]:
        out["out_" + name] = stack(o, steps, name)
    dst = ROOT / "tests/golden/elm_surface_albedo.npz"
    np.savez_compressed(dst, **out)
    print(dst, dst.stat().st_size, "bytes;", len(steps), "records; sunlit", int((out["in_coszen"][:, 0] > 0).sum()),
          "snl", np.unique(out["in_snl"]), "h2osno", np.unique(out["in_h2osno"]))


if __name__ == "__main__":
    main()
    canopy_fluxes()
    surface_albedo()
