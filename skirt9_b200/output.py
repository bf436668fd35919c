"""The output files of a simulation run through the Python host mirror, in the reference's formats:

* `<prefix>_<instrument>_sed.dat` / `_sedstats.dat` -- FluxRecorder::calibrateAndWrite, FluxRecorder.cpp:672-735: calibrated
  flux density per wavelength bin, one column per recorded component (TextOutFile.cpp:81-98: '#' header lines, %.9e columns);
* `<prefix>_<instrument>_<component>.fits` -- FluxRecorder.cpp:738-800 + FITSInOut::write, FITSInOut.cpp:127-215: the surface
  brightness cube as 32-bit big-endian floats with the reference's header cards, followed by the ASCII table extension that
  lists the wavelengths;
* `<prefix>_<probe>_J.dat` -- RadiationFieldProbe with the per-cell form: the mean intensity of every cell and bin.

Units are the reference's ExtragalacticUnits with wavelengthOutputStyle Wavelength and fluxOutputStyle Frequency (micron, Jy,
MJy/sr, W/m2/Hz/sr, Mpc, arcsec), the style of every ski file under tests/golden/ski.  The drop-in binary does not need this
module: there the reference's own writers run on the arrays the engine hands back (shim/GpuLifeCycle.cpp).
"""
import math
import time

import numpy as np

from . import abi
from . import host as H

MPC = 1e6 * H.PC
ARCSEC = math.pi / 180.0 / 3600.0


def _g(x):
    """StringUtils::toString(double): the shortest of %.10g-like renderings the reference uses in header lines."""
    return ("%.10g" % x)


def _grid_of(sim, instrument):
    i = sim.instruments[instrument]
    oligo = sim.oligoWavelengths is not None
    return i.wavelengthGrid if (i.wavelengthGrid is not None and not oligo) else sim.defaultWavelengthGrid


def _sed_columns(sim, instrument):
    """Column names and engine components in the reference's order (FluxRecorder.cpp:514-560)."""
    i = sim.instruments[instrument]
    cols = [("total flux", abi.SK_COMP_TOTAL)]
    if i.recordComponents:
        emission = sim.dustEmissionWLG is not None
        cols += [("transparent flux", abi.SK_COMP_TRANSPARENT), ("direct primary flux", abi.SK_COMP_PRIMARY_DIRECT),
                 ("scattered primary flux", abi.SK_COMP_PRIMARY_SCATTERED),
                 ("direct secondary flux", abi.SK_COMP_SECONDARY_DIRECT if emission else None),
                 ("scattered secondary flux", abi.SK_COMP_SECONDARY_SCATTERED if emission else None),
                 ("transparent secondary flux", abi.SK_COMP_SECONDARY_TRANSPARENT if emission else None)]
        cols += [("%d-times scattered primary flux" % (k + 1), abi.SK_COMP_PRIMARY_SCATTERED_LEVEL + k)
                 for k in range(i.numScatteringLevels)]
    return cols


def _observer_header(kind, ins):
    """The first line of the text files, FluxRecorder.cpp:620-643."""
    text = "# %s at inclination %s deg, azimuth %s deg" % (kind, _g(ins.inclination * 180 / math.pi), _g(ins.azimuth * 180 / math.pi))
    if ins.redshift:
        return text + ", redshift %s, luminosity distance %s Mpc" % (_g(ins.redshift), _g(ins.luminosityDistance / MPC))
    return text + ", distance %s Mpc" % _g(ins.distance / MPC)


def _write_columns(path, header_lines, names, rows):
    with open(path, "w") as f:
        for line in header_lines:
            f.write(line + "\n")
        for k, name in enumerate(names):
            f.write("# column %d: %s\n" % (k + 1, name))
        for row in rows:
            f.write(" ".join("%.9e" % v for v in row) + "\n")


def write_sed(sim, engine, instrument, path):
    """<prefix>_<instrument>_sed.dat (FluxRecorder.cpp:672-706)."""
    g = _grid_of(sim, instrument)
    cols = _sed_columns(sim, instrument)
    data = [g.lambdav * 1e6] + [sim.sed_flux_density(engine, instrument, c) if c is not None else np.zeros(g.num_bins)
                                for _, c in cols]
    names = ["wavelength; lambda (micron)"] + [n + "; F_nu (Jy)" for n, _ in cols]
    _write_columns(path, [_observer_header("SED", sim.instruments[instrument])], names, np.stack(data, axis=1))


def write_sed_statistics(sim, engine, instrument, path):
    """<prefix>_<instrument>_sedstats.dat (FluxRecorder.cpp:709-732): Sum w^k per bin, k = 0..4, uncalibrated (W)."""
    g = _grid_of(sim, instrument)
    st = engine.read_sed_stats(instrument)
    names = ["wavelength; lambda (micron)"] + ["Sum[w_i**%d] (1)" % k for k in range(5)]
    with open(path, "w") as f:
        for k, name in enumerate(names):
            f.write("# column %d: %s\n" % (k + 1, name))
        f.write("# --> w_i is luminosity contribution (in W) from i_th launched photon\n")
        for ell in range(g.num_bins):
            f.write(" ".join("%.9e" % v for v in [g.lambdav[ell] * 1e6] + [st[k][ell] for k in range(5)]) + "\n")


def _card(key, value, comment):
    """One 80-character header card in the fixed format cfitsio writes (ffpkys / ffpkyd with 9 decimals / ffpkyj / ffpkyl)."""
    if isinstance(value, bool):
        v = "%20s" % ("T" if value else "F")
    elif isinstance(value, int):
        v = "%20d" % value
    elif isinstance(value, float):
        v = "%20s" % ("%.9E" % value)
    else:
        v = "%-20s" % ("'%-8s'" % value)
    return ("%-8s= %s / %s" % (key, v, comment))[:80].ljust(80)


def _pad(b, fill):
    return b + fill * (-len(b) % 2880)


def write_fits_cube(path, cube, wavelengths_micron, ins, units="MJy/sr"):
    """FITSInOut::write (FITSInOut.cpp:127-215): cube[nz][ny][nx] as BITPIX -32 with the reference's cards, then the ASCII table
    extension 'Z-axis coordinate values' with one E16.9 column."""
    cube = np.asarray(cube, dtype=float)
    nz, ny, nx = cube.shape
    d_ang = ins.angularDiameterDistance if ins.redshift else ins.distance
    d_lum = ins.luminosityDistance if ins.redshift else ins.distance
    incx = 2.0 * math.atan(0.5 * ins.fieldOfViewX / nx / d_ang) / ARCSEC   # FluxRecorder.cpp:776-781
    incy = 2.0 * math.atan(0.5 * ins.fieldOfViewY / ny / d_ang) / ARCSEC
    xc = 2.0 * math.atan(0.5 * ins.centerX / d_ang) / ARCSEC
    yc = 2.0 * math.atan(0.5 * ins.centerY / d_ang) / ARCSEC
    cards = [_card("SIMPLE", True, "file does conform to FITS standard"), _card("BITPIX", -32, "number of bits per data pixel"),
             _card("NAXIS", 3, "number of data axes"), _card("NAXIS1", nx, "length of data axis 1"),
             _card("NAXIS2", ny, "length of data axis 2"), _card("NAXIS3", nz, "length of data axis 3"),
             _card("EXTEND", True, "FITS dataset may contain extensions"),
             "COMMENT   FITS (Flexible Image Transport System) format is defined in 'Astronomy".ljust(80),
             "COMMENT   and Astrophysics', volume 376, page 359; bibcode: 2001A&A...376..359H".ljust(80),
             _card("BSCALE", 1, "Array value scale"), _card("BZERO", 0, "Array value offset"),
             _card("DATE", time.strftime("%Y-%m-%dT%H:%M:%S", time.gmtime()), "Date and time of creation (UTC)"),
             _card("ORIGIN", "SKIRT simulation", "Astronomical Observatory, Ghent University"),
             _card("BUNIT", units, "Physical unit of the array values"),
             _card("CRPIX1", (nx + 1) / 2.0, "X-axis coordinate system reference pixel"),
             _card("CRVAL1", xc, "Coordinate value at X-axis reference pixel"),
             _card("CDELT1", incx, "Coordinate increment along X-axis"), _card("CUNIT1", "arcsec", "Physical units of the X-axis"),
             _card("CTYPE1", " ", "Linear X coordinates"),
             _card("CRPIX2", (ny + 1) / 2.0, "Y-axis coordinate system reference pixel"),
             _card("CRVAL2", yc, "Coordinate value at Y-axis reference pixel"),
             _card("CDELT2", incy, "Coordinate increment along Y-axis"), _card("CUNIT2", "arcsec", "Physical units of the Y-axis"),
             _card("CTYPE2", " ", "Linear Y coordinates"), _card("CUNIT3", "micron", "Physical units of the Z-axis"),
             _card("CROTA1", ins.inclination * 180 / math.pi, "Inclination angle, in deg"),
             _card("CROTA2", ins.azimuth * 180 / math.pi, "Azimuth angle, in deg"),
             _card("CROTA3", ins.roll * 180 / math.pi, "Roll angle, in deg"),
             _card("REDSHIFT", float(ins.redshift), "Redshift (if zero, distances are equal)"),
             _card("DISTLUMI", d_lum / MPC, "Luminosity distance"), _card("DISTANGD", d_ang / MPC, "Angular diameter distance"),
             _card("DISTUNIT", "Mpc", "Units of distances"), "END".ljust(80)]
    out = _pad("".join(cards).encode("ascii"), b" ")
    out += _pad(cube.astype(">f4").tobytes(), b"\0")
    ext = [_card("XTENSION", "TABLE", "ASCII table extension"), _card("BITPIX", 8, "8-bit ASCII characters"),
           _card("NAXIS", 2, "2-dimensional ASCII table"), _card("NAXIS1", 16, "width of table in characters"),
           _card("NAXIS2", nz, "number of rows in table"), _card("PCOUNT", 0, "no group parameters (required keyword)"),
           _card("GCOUNT", 1, "one data group (required keyword)"), _card("TFIELDS", 1, "number of fields in each row"),
           _card("TTYPE1", "GRID_POINTS", "label for field   1"), _card("TBCOL1", 1, "beginning column of field   1"),
           _card("TFORM1", "E16.9", "Fortran-77 format of field"), _card("TUNIT1", "micron", "physical unit of field"),
           _card("EXTNAME", "Z-axis coordinate values", "name of this ASCII table extension"), "END".ljust(80)]
    out += _pad("".join(ext).encode("ascii"), b" ")
    out += _pad("".join("%16s" % ("%.9E" % w) for w in wavelengths_micron).encode("ascii"), b" ")
    with open(path, "wb") as f:
        f.write(out)


def write_frames(sim, engine, instrument, prefix):
    """<prefix>_total.fits and, with recordComponents, one file per component that holds any flux (FluxRecorder.cpp:562-617,
    738-800).  Returns the list of files written."""
    ins = sim.instruments[instrument]
    g = _grid_of(sim, instrument)
    files = [("total", abi.SK_COMP_TOTAL)]
    if ins.recordComponents:
        files += [("transparent", abi.SK_COMP_TRANSPARENT), ("primarydirect", abi.SK_COMP_PRIMARY_DIRECT),
                  ("primaryscattered", abi.SK_COMP_PRIMARY_SCATTERED)]
        if sim.dustEmissionWLG is not None:
            files += [("secondarytransparent", abi.SK_COMP_SECONDARY_TRANSPARENT), ("secondarydirect", abi.SK_COMP_SECONDARY_DIRECT),
                      ("secondaryscattered", abi.SK_COMP_SECONDARY_SCATTERED)]
        files += [("primaryscattered%d" % (k + 1), abi.SK_COMP_PRIMARY_SCATTERED_LEVEL + k) for k in range(ins.numScatteringLevels)]
    written = []
    for name, comp in files:
        cube = sim.surface_brightness(engine, instrument, comp)
        if name != "total" and not np.any(cube):
            continue   # (the reference skips component files without any flux: "empty arrays will be ignored")
        path = "%s_%s.fits" % (prefix, name)
        write_fits_cube(path, cube, g.lambdav * 1e6, ins)
        written.append(path)
    return written


def write_radiation_field(sim, engine, path):
    """RadiationFieldProbe with PerCellForm (RadiationFieldProbe.cpp:32-90): J_nu of every cell at every bin of the radiation
    field grid, rf1 + rf2 (MediumSystem::meanIntensity, MediumSystem.cpp:1370-1380)."""
    g = sim.radiationFieldWLG
    J = sim.mean_intensity_nu(engine, 0)
    if sim.dustEmissionWLG is not None:
        J = J + sim.mean_intensity_nu(engine, 1)
    names = ["spatial cell index (1)"] + ["J_nu at lambda = %s micron (W/m2/Hz/sr)" % _g(lam * 1e6) for lam in g.lambdav]
    with open(path, "w") as f:
        f.write("# Mean intensity per spatial cell\n")
        for k, name in enumerate(names):
            f.write("# column %d: %s\n" % (k + 1, name))
        for m in range(J.shape[0]):
            f.write("%d " % m + " ".join("%.9e" % v for v in J[m]) + "\n")


def write_all(sim, engine, prefix):
    """Every instrument's files and, when the radiation field is stored, `<prefix>_rf_J.dat`; returns the paths."""
    paths = []
    for j, ins in enumerate(sim.instruments):
        base = "%s_%s" % (prefix, ins.instrumentName)
        if ins.kind in (abi.SK_INSTR_SED, abi.SK_INSTR_FULL):
            write_sed(sim, engine, j, base + "_sed.dat")
            paths.append(base + "_sed.dat")
            if ins.recordStatistics:
                write_sed_statistics(sim, engine, j, base + "_sedstats.dat")
                paths.append(base + "_sedstats.dat")
        if ins.kind in (abi.SK_INSTR_FRAME, abi.SK_INSTR_FULL):
            paths += write_frames(sim, engine, j, base)
    if sim.storeRadiationField:
        write_radiation_field(sim, engine, prefix + "_rf_J.dat")
        paths.append(prefix + "_rf_J.dat")
    return paths
