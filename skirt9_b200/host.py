"""Host-side mirror of the SKIRT 9 simulation items that feed the photon life cycle.

Class and attribute names follow the reference (`SKIRT/core/*.hpp`) so that a model is written the way the
corresponding ``.ski`` file reads.  Everything here is *setup* code that runs once on the host and ends in flat
tables handed to the engine through the C ABI (include/sk_engine.h); the life cycle itself is never executed
here -- there is no Python or CPU implementation of the hot path in the product.

All quantities are SI, like the reference's internal units (`SKIRT/utils/Constants.hpp`).
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
from scipy.special import erf, gamma as _gamma

from . import abi

# SKIRT/utils/Constants.hpp
C_LIGHT = 2.99792458e8
H_PLANCK = 6.62606957e-34
K_BOLTZ = 1.3806488e-23
PC = 3.08567758e16
LSUN = 3.839e26
MSUN = 1.9891e30
MICRON = 1e-6


# ---------------------------------------------------------------------------------------------------
# numerical helpers (SKIRT/utils/NR.hpp, SpecialFunctions.cpp)
# ---------------------------------------------------------------------------------------------------

def gln(p, x):
    """SpecialFunctions::gln, SpecialFunctions.cpp:798-811 (vectorised over x, scalar or array p)."""
    p = np.asarray(p, dtype=float)
    x = np.asarray(x, dtype=float)
    q = 1.0 - p
    lnx = np.log(x)
    s = q * lnx
    small = lnx * (1.0 + 0.5 * s + s * s / 6.0 + s * s * s / 24.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        big = (np.power(x, q) - 1.0) / q
    return np.where(q == 0.0, lnx, np.where(np.abs(q) < 1e-3, small, big))


def gln2(p, x1, x2):
    """SpecialFunctions::gln2, SpecialFunctions.cpp:815-818."""
    return x2 ** (1.0 - p) * gln(p, x1 / x2)


def cdf2_loglog(xv, pv):
    """NR::cdf2(loglog=true), SKIRT/utils/NR.cpp:25-60: returns (normalised pv, Pv, norm)."""
    xv = np.asarray(xv, dtype=float)
    pv = np.asarray(pv, dtype=float).copy()
    n = len(xv) - 1
    area = np.zeros(n)
    ok = (pv[:-1] > 0) & (pv[1:] > 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        alpha = np.log(pv[1:] / pv[:-1]) / np.log(xv[1:] / xv[:-1])
        a = pv[:-1] * xv[:-1] * gln(-alpha, xv[1:] / xv[:-1])
    area[ok] = a[ok]
    Pv = np.concatenate([[0.0], np.cumsum(area)])
    norm = Pv[-1]
    if norm > 0:
        pv /= norm
        Pv /= norm
    Pv[-1] = 1.0
    return pv, Pv, norm


def planck(lam, T):
    """PlanckFunction::value, SKIRT/utils/PlanckFunction.cpp:24-27."""
    f1 = H_PLANCK * C_LIGHT / (K_BOLTZ * T)
    f2 = 2.0 * H_PLANCK * C_LIGHT * C_LIGHT
    with np.errstate(over="ignore"):   # (far on the Wien side expm1 overflows to inf and the value is 0, as in the reference)
        return f2 / np.power(lam, 5) / np.expm1(f1 / lam)


# ---------------------------------------------------------------------------------------------------
# wavelength grids (DisjointWavelengthGrid.cpp:22-136)
# ---------------------------------------------------------------------------------------------------

class DisjointWavelengthGrid:
    def __init__(self):
        self.lambdav = self.borderv = self.ellv = self.dlambdav = None

    def _set_range(self, lambdav, log_scale):
        lam = np.sort(np.asarray(lambdav, dtype=float))
        n = len(lam)
        border = np.empty(n + 1)
        if n == 1:
            border[0], border[1] = lam[0] * 0.999, lam[0] * 1.001
        elif log_scale:
            border[0] = math.sqrt(lam[0] ** 3 / lam[1])
            border[1:n] = np.sqrt(lam[:-1] * lam[1:])
            border[n] = math.sqrt(lam[-1] ** 3 / lam[-2])
        else:
            border[0] = (3 * lam[0] - lam[1]) / 2
            border[1:n] = (lam[:-1] + lam[1:]) / 2
            border[n] = (3 * lam[-1] - lam[-2]) / 2
        self.lambdav, self.borderv = lam, border
        self.dlambdav = border[1:] - border[:-1]
        self.ellv = np.concatenate([[-1], np.arange(n), [-1]]).astype(np.int32)

    def _set_bins(self, lambdav, rel_half_width, constant_width):
        lam = np.sort(np.asarray(lambdav, dtype=float))
        n = len(lam)
        if constant_width:
            delta = lam[0] * rel_half_width
            left, right = lam - delta, lam + delta
        else:
            left, right = lam * (1 - rel_half_width), lam * (1 + rel_half_width)
        border = np.empty(2 * n)
        border[0::2], border[1::2] = left, right
        ell = np.full(2 * n + 1, -1, dtype=np.int32)
        ell[1::2] = np.arange(n)
        self.lambdav, self.borderv, self.ellv = lam, border, ell
        self.dlambdav = right - left

    @property
    def num_bins(self):
        return len(self.lambdav)

    def wavelength_range(self):
        return float(self.borderv[0]), float(self.borderv[-1])

    def table(self):
        return {"borders": self.borderv, "ell": self.ellv, "lambda": self.lambdav, "dlambda": self.dlambdav}


class LogWavelengthGrid(DisjointWavelengthGrid):
    """LogWavelengthGrid.cpp:12-24: N characteristic wavelengths from min to max inclusive."""

    def __init__(self, minWavelength, maxWavelength, numWavelengths):
        super().__init__()
        n = numWavelengths - 1
        lam = np.exp(math.log(minWavelength) + np.arange(n + 1) * (math.log(maxWavelength / minWavelength) / n))
        self._set_range(lam, True)


class ListWavelengthGrid(DisjointWavelengthGrid):
    def __init__(self, wavelengths, log=True):
        super().__init__()
        self._set_range(wavelengths, log)


class OligoWavelengthGrid(DisjointWavelengthGrid):
    """OligoWavelengthGrid.cpp:25-31."""

    def __init__(self, wavelengths):
        super().__init__()
        self._set_bins(wavelengths, 1e-3, True)


# ---------------------------------------------------------------------------------------------------
# geometries
# ---------------------------------------------------------------------------------------------------

class ShellGeometry:
    """ShellGeometry.cpp:14-57."""

    def __init__(self, minRadius, maxRadius, exponent):
        self.rmin, self.rmax, self.p = minRadius, maxRadius, exponent
        self.smin = float(gln(self.p - 2.0, self.rmin))
        self.sdiff = float(gln2(self.p - 2.0, self.rmax, self.rmin))
        self.tmin = self.rmin ** (3.0 - self.p)
        self.tmax = self.rmax ** (3.0 - self.p)
        self.A = 0.25 / math.pi / self.sdiff

    def density(self, x, y, z):
        r = np.sqrt(x * x + y * y + z * z)
        with np.errstate(divide="ignore"):
            rho = self.A * np.power(r, -self.p)
        return np.where((r < self.rmin) | (r > self.rmax), 0.0, rho)

    def SigmaZ(self):
        return 2.0 * self.A * float(gln2(self.p, self.rmax, self.rmin))

    def density_fields(self):
        return abi.SK_GEOM_SHELL, [self.rmin, self.rmax, self.p, self.A]

    def source_fields(self):
        return {"geometry": abi.SK_GEOM_SHELL,
                "geom_params": [self.rmin, self.rmax, self.p, self.smin, self.sdiff, self.tmin, self.tmax]}


class ExpDiskGeometry:
    """ExpDiskGeometry.cpp:13-90."""

    def __init__(self, scaleLength, scaleHeight, minRadius=0.0, maxRadius=0.0, maxZ=0.0):
        self.hR, self.hz, self.Rmin, self.Rmax, self.zmax = scaleLength, scaleHeight, minRadius, maxRadius, maxZ
        intz = -2.0 * self.hz * math.expm1(-self.zmax / self.hz) if self.zmax > 0 else 2.0 * self.hz
        tmin = math.exp(-self.Rmin / self.hR) * (1 + self.Rmin / self.hR) if self.Rmin > 0 else 1.0
        tmax = math.exp(-self.Rmax / self.hR) * (1 + self.Rmax / self.hR) if self.Rmax > 0 else 0.0
        self.rho0 = 1.0 / (self.hR ** 2 * (tmin - tmax) * 2 * math.pi * intz)

    def density_Rz(self, R, z):
        absz = np.abs(z)
        rho = self.rho0 * np.exp(-R / self.hR) * np.exp(-absz / self.hz)
        out = np.zeros_like(rho)
        ok = R >= self.Rmin
        if self.Rmax > 0:
            ok &= R <= self.Rmax
        if self.zmax > 0:
            ok &= absz <= self.zmax
        out[ok] = rho[ok]
        return out

    def density(self, x, y, z):
        return self.density_Rz(np.sqrt(x * x + y * y), z)

    def SigmaZ(self):
        if self.Rmin > 0:
            return 0.0
        if self.zmax > 0:
            return -2.0 * self.rho0 * self.hz * math.expm1(-self.zmax / self.hz)
        return 2.0 * self.rho0 * self.hz

    def density_fields(self):
        return abi.SK_GEOM_EXPDISK, [self.hR, self.hz, self.Rmin, self.Rmax, self.zmax, self.rho0]

    def source_fields(self):
        return {"geometry": abi.SK_GEOM_EXPDISK, "geom_params": [self.hR, self.hz, self.Rmin, self.Rmax, self.zmax]}


class RingGeometry:
    """RingGeometry.cpp:14-80."""

    def __init__(self, ringRadius, width, height):
        self.R0, self.w, self.hz = ringRadius, width, height
        t = self.R0 / self.w / math.sqrt(2.0)
        intz = 2.0 * self.hz
        intR = self.w ** 2 * (math.exp(-t * t) + math.sqrt(math.pi) * t * (1.0 + math.erf(t)))
        self.A = 1.0 / (2.0 * math.pi * intz * intR)
        NRr = 330
        self.Rv = np.linspace(max(0.0, self.R0 - 8 * self.w), self.R0 + 8 * self.w, NRr)
        u = (self.R0 - self.Rv) / self.w / math.sqrt(2.0)
        self.Xv = 4.0 * math.pi * self.A * self.hz * self.w ** 2 * (
            math.exp(-t * t) - np.exp(-u * u) + math.sqrt(math.pi) * t * (math.erf(t) - erf(u)))
        self.Xv[0], self.Xv[-1] = 0.0, 1.0

    def density(self, x, y, z):
        R = np.sqrt(x * x + y * y)
        u = (R - self.R0) / (math.sqrt(2.0) * self.w)
        return self.A * np.exp(-u * u) * np.exp(-np.abs(z) / self.hz)

    def SigmaZ(self):
        t = self.R0 / (math.sqrt(2.0) * self.w)
        return 2.0 * self.A * self.hz * math.exp(-t * t)

    def density_fields(self):
        return abi.SK_GEOM_RING, [self.R0, self.w, self.hz, self.A]

    def source_fields(self):
        return {"geometry": abi.SK_GEOM_RING, "geom_params": [self.R0, self.w, self.hz],
                "geom_table_x": self.Rv, "geom_table_P": self.Xv}


class SpiralStructureGeometryDecorator:
    """SpiralStructureGeometryDecorator.cpp:12-76 decorating an ExpDiskGeometry."""

    def __init__(self, geometry: ExpDiskGeometry, numArms, pitchAngle, radiusZeroPoint, phaseZeroPoint,
                 perturbationWeight, index):
        self.geometry = geometry
        self.m, self.p, self.R0, self.phi0, self.w, self.N = (numArms, pitchAngle, radiusZeroPoint, phaseZeroPoint,
                                                              perturbationWeight, index)
        self.tanp = math.tan(self.p)
        self.cn = math.sqrt(math.pi) * _gamma(self.N + 1.0) / _gamma(self.N + 0.5)

    def perturbation(self, R, phi):
        with np.errstate(divide="ignore"):
            g = np.log(R / self.R0) / self.tanp + self.phi0 + 0.5 * math.pi / self.m
        return (1.0 - self.w) + self.w * self.cn * np.power(np.sin(0.5 * self.m * (g - phi)), 2 * self.N)

    def density(self, x, y, z):
        R = np.sqrt(x * x + y * y)
        return self.geometry.density_Rz(R, z) * self.perturbation(R, np.arctan2(y, x))

    def SigmaZ(self):
        return self.geometry.SigmaZ()

    def density_fields(self):
        g = self.geometry
        return abi.SK_GEOM_SPIRAL_EXPDISK, [g.hR, g.hz, g.Rmin, g.Rmax, g.zmax, g.rho0, self.m, self.tanp, self.R0,
                                            self.phi0, self.w, self.N, self.cn]

    def source_fields(self):
        g = self.geometry
        return {"geometry": abi.SK_GEOM_SPIRAL_EXPDISK,
                "geom_params": [g.hR, g.hz, g.Rmin, g.Rmax, g.zmax, self.m, self.tanp, self.R0, self.phi0, self.w, self.N,
                                self.cn]}


# ---------------------------------------------------------------------------------------------------
# dust mix, medium
# ---------------------------------------------------------------------------------------------------

def _clamped_resample(lam, inlam, inval, loglog):
    """NR::clampedResample<interpolateLogLog|interpolateLogLin>, NR.hpp (values clamped outside the table)."""
    lam = np.asarray(lam, dtype=float)
    inlam = np.asarray(inlam, dtype=float)
    inval = np.asarray(inval, dtype=float)
    if len(inlam) == 1:
        return np.full_like(lam, inval[0])
    lx = np.log(np.clip(lam, inlam[0], inlam[-1]))
    if loglog and np.all(inval > 0):
        return np.exp(np.interp(lx, np.log(inlam), np.log(inval)))
    return np.interp(lx, np.log(inlam), inval)


class MeanListDustMix:
    """MeanListDustMix.cpp:12-27 + TabulatedDustMix.cpp:12-44 + DustMix::setupSelfAfter (DustMix.cpp:47-246)."""
    MU = 1.5e-29  # kg per hydrogen atom

    def __init__(self, wavelengths, extinctionCoefficients, albedos, asymmetryParameters):
        o = np.argsort(wavelengths)
        self.inlam = np.asarray(wavelengths, dtype=float)[o]
        self.inkappa = np.asarray(extinctionCoefficients, dtype=float)[o]
        self.inalbedo = np.asarray(albedos, dtype=float)[o]
        self.ing = np.asarray(asymmetryParameters, dtype=float)[o]
        self.lambda_border = None

    def setup(self, wl_range, extra_wavelengths):
        per_dex = 1000.0
        kmin = math.floor(per_dex * math.log10(wl_range[0]))
        kmax = math.ceil(per_dex * math.log10(wl_range[1]))
        lam = np.power(10.0, np.arange(kmin, kmax + 1) / per_dex)
        lam = np.unique(np.concatenate([lam, np.asarray(list(extra_wavelengths), dtype=float)]))
        dm = 0.1
        cutoff = lam[-1] > dm
        if cutoff:
            lam = lam[lam <= dm]
            if lam[-1] != dm:
                lam = np.append(lam, dm)
            lam = np.append(lam, dm * 1.001)
        self.lambdav = lam
        self.lambda_border = lam.copy()
        self.lambda_border[1:] = np.sqrt(lam[1:] * lam[:-1])
        mu = self.MU
        self.sigma_abs = _clamped_resample(lam, self.inlam, mu * self.inkappa * (1 - self.inalbedo), True)
        self.sigma_sca = _clamped_resample(lam, self.inlam, mu * self.inkappa * self.inalbedo, True)
        g = _clamped_resample(lam, self.inlam, self.ing, False)
        self.asymmpar = np.where(np.abs(g) > 0.999999, np.copysign(0.999999, g), g)
        self.sigma_abs_unclipped = self.sigma_abs.copy()
        if cutoff:
            self.sigma_abs[-2:] = 0.0
            self.sigma_sca[-2:] = 0.0
        self.mu = mu

    def precalculate(self, rf_grid, em_grid):
        """EquilibriumDustEmissionCalculator::precalculate, EquilibriumDustEmissionCalculator.cpp:18-93 (one dust
        population, no CMB heating): sigma_abs resampled log-log on the radiation-field and (extended) emission grids and
        the Planck-integrated absorption cross section on the power-law temperature grid 0..5000 K."""
        lam, sig = self.lambdav, self.sigma_abs_unclipped

        def resample_loglog(x):  # NR::resample<interpolateLogLog>: zero outside the table, NR.hpp:406-413
            x = np.asarray(x, dtype=float)
            out = np.zeros_like(x)
            ok = (x >= lam[0]) & (x <= lam[-1])
            i = np.clip(np.searchsorted(lam, x[ok], side="right") - 1, 0, len(lam) - 2)
            with np.errstate(divide="ignore", invalid="ignore"):
                v = sig[i] * np.exp(np.log(x[ok] / lam[i]) / np.log(lam[i + 1] / lam[i]) * np.log(sig[i + 1] / sig[i]))
            bad = (sig[i] <= 0) | (sig[i + 1] <= 0)
            v[bad] = np.where(x[ok][bad] == lam[i][bad], sig[i][bad], np.where(x[ok][bad] == lam[i + 1][bad], sig[i + 1][bad], 0.0))
            out[ok] = v
            return out

        self.rf_sigma_abs = resample_loglog(rf_grid.lambdav)
        self.em_lambda = np.concatenate([[em_grid.borderv[0]], em_grid.lambdav, [em_grid.borderv[-1]]])  # extlambdav
        self.em_sigma_abs = resample_loglog(self.em_lambda)
        n, ratio, xmax = 1000, 500.0, 5000.0       # NR::buildPowerLawGrid(_Tv, 0., 5000., 1000, 500.), NR.hpp:221-235
        q = ratio ** (1.0 / (n - 1))
        self.Tv = (1.0 - q ** np.arange(n + 1)) / (1.0 - q ** n) * xmax
        dl = lam[1:] - lam[:-1]
        planckabs = np.zeros(n + 1)
        with np.errstate(over="ignore", divide="ignore"):
            for p_ in range(1, n + 1):
                planckabs[p_] = np.sum(sig[1:] * planck(lam[1:], self.Tv[p_]) * dl)
        self.planck_abs = planckabs

    def index_for_lambda(self, lam):
        """DustMix::indexForLambda = NR::locateClip(_lambdav, lambda), DustMix.cpp:276-279."""
        i = np.searchsorted(self.lambda_border, lam, side="right") - 1
        return np.clip(i, 0, len(self.lambda_border) - 2)

    def section_ext(self, lam):
        i = self.index_for_lambda(lam)
        return self.sigma_abs[i] + self.sigma_sca[i]


class _PowerLawVectorField:
    """Common part of RadialVectorField / CylindricalVectorField: a unit vector times (r/unityRadius)^exponent inside
    (exponent > 0) or outside (exponent < 0) the unity radius, 1 elsewhere; null on the axis / at the origin."""
    kind = abi.SK_VEL_NONE

    def __init__(self, unityRadius=0.0, exponent=0.0):
        self.unityRadius, self.exponent = unityRadius, exponent

    def _unit(self, r):
        raise NotImplementedError

    def vector(self, r):
        r = np.atleast_2d(np.asarray(r, dtype=float))
        u = self._unit(r)
        rr = np.sqrt((u * u).sum(axis=1))
        with np.errstate(divide="ignore", invalid="ignore"):
            u = np.where(rr[:, None] > 0, u / rr[:, None], 0.0)
            f = np.ones_like(rr)
            if self.unityRadius > 0:
                sel = (rr < self.unityRadius) if self.exponent > 0 else (rr > self.unityRadius) if self.exponent < 0 \
                    else np.zeros_like(rr, dtype=bool)
                f = np.where(sel & (rr > 0), (rr / self.unityRadius) ** self.exponent, 1.0)
        return f[:, None] * u

    def source_fields(self, magnitude):
        return {"velocity_kind": self.kind, "velocity": (magnitude, self.unityRadius, self.exponent)}


class RadialVectorField(_PowerLawVectorField):
    """RadialVectorField.cpp:19-37."""
    kind = abi.SK_VEL_RADIAL

    def _unit(self, r):
        return r.copy()


class CylindricalVectorField(_PowerLawVectorField):
    """CylindricalVectorField.cpp:19-38: rotation about the z axis."""
    kind = abi.SK_VEL_CYLINDRICAL

    def _unit(self, r):
        return np.stack([-r[:, 1], r[:, 0], np.zeros(len(r))], axis=1)


class UnidirectionalVectorField:
    """UnidirectionalVectorField: the same unit vector everywhere."""

    def __init__(self, x=0.0, y=0.0, z=1.0):
        n = math.sqrt(x * x + y * y + z * z)
        self.u = (x / n, y / n, z / n)

    def vector(self, r):
        r = np.atleast_2d(np.asarray(r, dtype=float))
        return np.tile(np.array(self.u), (len(r), 1))

    def source_fields(self, magnitude):
        return {"velocity_kind": abi.SK_VEL_CONSTANT, "velocity": tuple(magnitude * c for c in self.u)}


class GeometricMedium:
    """GeometricMedium.cpp:14-20 with OpticalDepthMaterialNormalization (axis Z), .cpp:12-31; optionally with a bulk velocity
    velocityMagnitude * velocityDistribution(r) (GeometricMedium.cpp:51-68)."""

    def __init__(self, geometry, materialMix, opticalDepth, wavelength, velocityMagnitude=0.0, velocityDistribution=None):
        self.geometry, self.mix = geometry, materialMix
        self.tau, self.norm_wavelength = opticalDepth, wavelength
        self.number = None
        self.velocityMagnitude, self.velocityDistribution = velocityMagnitude, velocityDistribution

    def has_velocity(self):
        return self.velocityDistribution is not None and self.velocityMagnitude != 0

    def bulk_velocity(self, r):
        return self.velocityMagnitude * self.velocityDistribution.vector(r)

    def setup(self):
        section = float(self.mix.section_ext(self.norm_wavelength))
        self.number = (self.tau / section) / self.geometry.SigmaZ()
        self.mass = self.number * self.mix.mu

    def number_density(self, x, y, z):
        return self.number * self.geometry.density(x, y, z)

    def density_geometry(self):
        """The medium as sk_engine_build_octree / sk_engine_sample_medium take it (include/sk_engine.h)."""
        kind, params = self.geometry.density_fields()
        return abi.SkDensityGeometry.make(kind, params, self.number, self.mass)


class ParticleMedium:
    """ParticleMedium (ImportedMedium.cpp:198-203) on a ParticleSnapshot with the CubicSplineSmoothingKernel: particles[n][5] =
    x y z h M (m, m, m, m, kg).  The mirror does not evaluate the smoothed density itself: the cell densities are imported
    (MonteCarloSimulation.density, e.g. from a reference run) or sampled by the engine (deviceSetup:
    sk_engine_sample_medium_particles)."""

    def __init__(self, particles, materialMix, massFraction=1.0):
        self.particles = np.ascontiguousarray(particles, dtype=float).reshape(-1, 5)
        self.mix, self.massFraction = materialMix, massFraction
        self.norm_wavelength = None

    def setup(self):
        self.mass = float(self.particles[:, 4].sum()) * self.massFraction
        self.number = self.mass / self.mix.mu

    @property
    def density_scale(self):
        return self.massFraction / self.mix.mu

    def number_density(self, x, y, z):
        raise NotImplementedError("the smoothed-particle density is sampled on the engine's side (deviceSetup)")


# ---------------------------------------------------------------------------------------------------
# spatial grids
# ---------------------------------------------------------------------------------------------------

class CartesianSpatialGrid:
    """CartesianSpatialGrid.cpp:22-60 with LinMesh; cell index m = k + Nz*j + Nz*Ny*i (.cpp:210-213)."""

    def __init__(self, minX, maxX, minY, maxY, minZ, maxZ, numX, numY, numZ):
        self.xv = minX + (maxX - minX) * np.arange(numX + 1) / numX
        self.yv = minY + (maxY - minY) * np.arange(numY + 1) / numY
        self.zv = minZ + (maxZ - minZ) * np.arange(numZ + 1) / numZ
        self.extent = (minX, minY, minZ, maxX, maxY, maxZ)

    def setup(self, media, num_density_samples, rng):
        pass

    @property
    def num_cells(self):
        return (len(self.xv) - 1) * (len(self.yv) - 1) * (len(self.zv) - 1)

    def cell_boxes(self):
        nx, ny, nz = len(self.xv) - 1, len(self.yv) - 1, len(self.zv) - 1
        i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
        i, j, k = i.ravel(), j.ravel(), k.ravel()
        return np.stack([self.xv[i], self.yv[j], self.zv[k], self.xv[i + 1], self.yv[j + 1], self.zv[k + 1]], axis=1)

    def configure(self, engine):
        engine.set_grid_cartesian(self.xv, self.yv, self.zv)


class DensityTreePolicy:
    """DensityTreePolicy.cpp:116-227, dust criteria: mass fraction, optical depth across the node's diagonal at
    `wavelength`, density dispersion over the samples (a value of 0 disables a criterion, DensityTreePolicy.cpp:64-66)."""

    def __init__(self, minLevel, maxLevel, maxDustFraction, maxDustOpticalDepth=0.0, wavelength=0.55e-6,
                 maxDustDensityDispersion=0.0):
        self.minLevel, self.maxLevel, self.maxDustFraction = minLevel, maxLevel, maxDustFraction
        self.maxDustOpticalDepth, self.wavelength = maxDustOpticalDepth, wavelength
        self.maxDustDensityDispersion = maxDustDensityDispersion

    def dust_kappa(self, media):
        """_dustKappa, DensityTreePolicy.cpp:75-84: sum of the extinction sections over the sum of the masses per entity."""
        if not self.maxDustOpticalDepth > 0:
            return 0.0
        return sum(float(m.mix.section_ext(self.wavelength)) for m in media) / sum(m.mix.mu for m in media)


class PolicyTreeSpatialGrid:
    """PolicyTreeSpatialGrid with treeType OctTree: TreeSpatialGrid.cpp:23-49, DensityTreePolicy.cpp:242-309,
    OctTreeNode.cpp:22-35.  Nodes are kept in the reference's breadth-first `_nodev` order."""

    def __init__(self, minX, maxX, minY, maxY, minZ, maxZ, policy: DensityTreePolicy):
        self.extent = (minX, minY, minZ, maxX, maxY, maxZ)
        self.policy = policy
        self.first_child = None
        self.boxes = None

    def setup(self, media, num_density_samples, rng):
        pol = self.policy
        boxes = [np.array([self.extent], dtype=float)]
        first_child = [np.array([-1], dtype=np.int64)]
        level, lbeg, total = 0, 0, 1
        cur = boxes[0]
        while len(cur):
            n = len(cur)
            if level < pol.minLevel:
                divide = np.ones(n, dtype=bool)
            elif level >= pol.maxLevel:
                divide = np.zeros(n, dtype=bool)
            else:
                # needsSubdivide, DensityTreePolicy.cpp:130-210: dust mass density at numSamples random positions
                mass = sum(med.mass for med in media)
                rho, rhomin, rhomax = np.zeros(n), np.full(n, np.inf), np.zeros(n)
                for _ in range(num_density_samples):
                    u = rng.random((n, 3))
                    p = cur[:, :3] + u * (cur[:, 3:] - cur[:, :3])
                    rhoi = np.zeros(n)
                    for med in media:
                        rhoi += med.mass * med.geometry.density(p[:, 0], p[:, 1], p[:, 2])
                    rho += rhoi
                    rhomin, rhomax = np.minimum(rhomin, rhoi), np.maximum(rhomax, rhoi)
                rho /= num_density_samples
                size = cur[:, 3:] - cur[:, :3]
                divide = np.zeros(n, dtype=bool)
                if pol.maxDustFraction > 0:
                    divide |= rho * np.prod(size, axis=1) / mass > pol.maxDustFraction
                if pol.maxDustOpticalDepth > 0:
                    divide |= pol.dust_kappa(media) * rho * np.linalg.norm(size, axis=1) > pol.maxDustOpticalDepth
                if pol.maxDustDensityDispersion > 0:
                    with np.errstate(invalid="ignore", divide="ignore"):
                        q = np.where(rhomax > 0, (rhomax - rhomin) / rhomax, 0.0)
                    divide |= q > pol.maxDustDensityDispersion
            idx = np.nonzero(divide)[0]
            fc = np.full(n, -1, dtype=np.int64)
            fc[idx] = total + 8 * np.arange(len(idx))
            first_child[-1] = fc
            par = cur[idx]
            c = 0.5 * (par[:, :3] + par[:, 3:])  # Box::center, Box.hpp:135
            kids = np.empty((len(idx), 8, 6))
            for ch in range(8):
                for ax in range(3):
                    hi = (ch >> ax) & 1
                    kids[:, ch, ax] = np.where(hi, c[:, ax], par[:, ax])
                    kids[:, ch, ax + 3] = np.where(hi, par[:, ax + 3], c[:, ax])
            cur = kids.reshape(-1, 6)
            if len(cur):
                boxes.append(cur)
                first_child.append(np.full(len(cur), -1, dtype=np.int64))
            lbeg = total
            total += len(cur)
            level += 1
        self.boxes = np.concatenate(boxes)
        self.first_child = np.concatenate(first_child).astype(np.int32)

    @property
    def num_cells(self):
        return int((self.first_child < 0).sum())

    def cell_boxes(self):
        return self.boxes[self.first_child < 0]

    def configure(self, engine):
        engine.set_grid_octree(self.extent, self.first_child)

    def tree_policy(self, num_density_samples, media=()):
        pol = self.policy
        return abi.SkTreePolicy(pol.minLevel, pol.maxLevel, num_density_samples, 0, pol.maxDustFraction,
                                pol.maxDustOpticalDepth, pol.maxDustDensityDispersion, pol.dust_kappa(media))

    def adopt(self, first_child):
        """Takes over a node list built elsewhere (sk_engine_build_octree) and derives the node boxes from it."""
        self.first_child = np.asarray(first_child, dtype=np.int32)
        self.boxes = boxes_from_first_child(self.extent, self.first_child)


def boxes_from_first_child(extent, first_child):
    """Node boxes of a breadth-first octree node list by Box::center recursion (Box.hpp:135, OctTreeNode.cpp:22-35)."""
    fc = np.asarray(first_child, dtype=np.int64)
    boxes = np.empty((len(fc), 6))
    boxes[0] = extent
    level = np.array([0])
    while len(level):
        par = level[fc[level] >= 0]
        if not len(par):
            break
        b = boxes[par]
        c = 0.5 * (b[:, :3] + b[:, 3:])
        kids = np.empty((len(par), 8, 6))
        for ch in range(8):
            for ax in range(3):
                hi = (ch >> ax) & 1
                kids[:, ch, ax] = c[:, ax] if hi else b[:, ax]
                kids[:, ch, ax + 3] = b[:, ax + 3] if hi else c[:, ax]
        idx = (fc[par][:, None] + np.arange(8)[None, :]).ravel()
        boxes[idx] = kids.reshape(-1, 6)
        level = idx
    return boxes


class FileTreeSpatialGrid(PolicyTreeSpatialGrid):
    """FileTreeSpatialGrid.cpp:22-78: rebuilds an octree from the 0/1 pre-order topology stream written by
    TreeSpatialGridTopologyProbe (TreeSpatialGrid.cpp:232-251); nodes are created depth-first."""

    def __init__(self, minX, maxX, minY, maxY, minZ, maxZ, topology: Sequence[int], policyOrder: bool = False):
        """policyOrder=True renumbers the nodes breadth-first, level by level, the way DensityTreePolicy::constructTree
        (DensityTreePolicy.cpp:242-309) created them in the run that exported the topology, so that cell indices match
        that run's per-cell output files."""
        super().__init__(minX, maxX, minY, maxY, minZ, maxZ, None)
        self.topology = list(topology)
        self.policyOrder = policyOrder

    def setup(self, media, num_density_samples, rng):
        topo = self.topology
        boxes = [list(self.extent)]
        first_child = [-1]
        pos = [0]

        def subdivide_if_needed(node):
            # subdivideNodeIfNeeded, FileTreeSpatialGrid.cpp:20-37: one flag per node, pre-order
            flag = topo[pos[0]]
            pos[0] += 1
            if not flag:
                return
            b = boxes[node]
            c = [0.5 * (b[a] + b[a + 3]) for a in range(3)]
            fc = len(boxes)
            first_child[node] = fc
            for ch in range(8):
                nb = [0.0] * 6
                for ax in range(3):
                    hi = (ch >> ax) & 1
                    nb[ax] = c[ax] if hi else b[ax]
                    nb[ax + 3] = b[ax + 3] if hi else c[ax]
                boxes.append(nb)
                first_child.append(-1)
            for ch in range(8):
                subdivide_if_needed(fc + ch)

        # the first item is the number of children of the root (8 or 0), then the root's own flag follows
        if topo[0] not in (0, 8):
            raise ValueError("only octree topologies are supported")
        pos[0] = 1
        subdivide_if_needed(0)
        self.boxes = np.asarray(boxes, dtype=float)
        self.first_child = np.asarray(first_child, dtype=np.int32)
        if self.policyOrder:
            fc = self.first_child
            order = [0]
            for node in order:  # breadth-first queue: children are appended when their parent is visited
                if fc[node] >= 0:
                    order.extend(range(fc[node], fc[node] + 8))
            order = np.asarray(order)
            newid = np.empty(len(order), dtype=np.int64)
            newid[order] = np.arange(len(order))
            self.boxes = self.boxes[order]
            self.first_child = np.where(fc[order] >= 0, newid[np.maximum(fc[order], 0)], -1).astype(np.int32)


class VoronoiMeshSpatialGrid:
    """VoronoiMeshSpatialGrid with policy ImportedSites / given sites (VoronoiMeshSpatialGrid.cpp:52-144): a Voronoi
    tessellation of the domain box generated by the given sites.  The reference builds it with the vendored voro++
    (VoronoiMeshSnapshot::buildMesh, VoronoiMeshSnapshot.cpp:491-730); this mirror takes the neighbour relation from the
    Delaunay triangulation of the sites (scipy / Qhull) -- the same tessellation -- and lists all six domain walls for
    every cell (a wall that does not bound a cell is never the nearest exit, so this is equivalent to the reference's
    lists, which hold only the walls that do).  Cell volumes, which the reference gets from voro++, can be supplied."""

    def __init__(self, minX, maxX, minY, maxY, minZ, maxZ, sites, volumes=None):
        self.extent = (minX, minY, minZ, maxX, maxY, maxZ)
        sites = np.asarray(sites, dtype=float).reshape(-1, 3)
        ext = np.asarray(self.extent)
        inside = np.all((sites > ext[:3]) & (sites < ext[3:]), axis=1)   # sites outside the domain are discarded
        sites = sites[inside]
        self.sites = sites[np.argsort(sites[:, 0], kind="stable")]       # cells in order of increasing x, .cpp:507-508
        self.volumes = None if volumes is None else np.asarray(volumes, dtype=float)
        self.cell_extents = None   # enclosing boxes of the cells (VoronoiMeshSnapshot::Cell is a Box), [n,6]
        self.nbr_offset = self.nbr_index = None

    def setup(self, media, num_density_samples, rng):
        from scipy.spatial import Delaunay
        n = len(self.sites)
        tri = Delaunay(self.sites)
        indptr, indices = tri.vertex_neighbor_vertices
        # The domain walls a cell lists (VoronoiMeshSnapshot.cpp:1134-1143: voro++ reports the walls that bound the cell): the
        # vertices of a bounded Voronoi cell are the circumcentres of the Delaunay tetrahedra around its site, so the cell
        # reaches beyond a wall exactly when one of them does; the unbounded cells of the hull sites list all six.
        S = tri.simplices
        A = self.sites[S[:, 0]]
        B, C, D = self.sites[S[:, 1]] - A, self.sites[S[:, 2]] - A, self.sites[S[:, 3]] - A
        mat = np.stack([B, C, D], axis=1)
        rhs = 0.5 * np.stack([(B * B).sum(1), (C * C).sum(1), (D * D).sum(1)], axis=1)
        good = np.abs(np.linalg.det(mat)) > 1e-12 * np.prod(np.linalg.norm(mat, axis=2), axis=1)
        cc = np.full_like(A, np.nan)
        cc[good] = np.linalg.solve(mat[good], rhs[good][..., None])[..., 0] + A[good]
        lo, hi = np.full((n, 3), np.inf), np.full((n, 3), -np.inf)
        sliver = np.zeros(n, dtype=bool)     # a flat tetrahedron has its circumcentre (nearly) at infinity: all walls
        for j in range(4):
            np.minimum.at(lo, S[good, j], cc[good])
            np.maximum.at(hi, S[good, j], cc[good])
            sliver[S[~good, j]] = True
        every = sliver
        every[np.unique(tri.convex_hull)] = True
        ext = np.asarray(self.extent)
        touch = np.concatenate([(lo < ext[:3]) | every[:, None], (hi > ext[3:]) | every[:, None]], axis=1)  # xmin ymin zmin xmax ymax zmax
        touch = touch[:, [0, 3, 1, 4, 2, 5]]                                                              # walls -1 .. -6
        counts = np.diff(indptr) + touch.sum(axis=1)
        self.nbr_offset = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        idx = np.empty(self.nbr_offset[-1], dtype=np.int32)
        walls = np.array([-1, -2, -3, -4, -5, -6], dtype=np.int32)
        for m in range(n):
            a, b = self.nbr_offset[m], self.nbr_offset[m + 1]
            k = indptr[m + 1] - indptr[m]
            idx[a:a + k] = indices[indptr[m]:indptr[m + 1]]
            idx[a + k:b] = walls[touch[m]]
        self.nbr_index = idx

    @property
    def num_cells(self):
        return len(self.sites)

    def cell_boxes(self):
        """Not boxes: the mirror only needs volumes (for mean intensities) and sample positions (cell densities)."""
        raise NotImplementedError

    def cell_volumes(self):
        if self.volumes is not None:
            return self.volumes
        ext = np.asarray(self.extent)
        return np.full(self.num_cells, np.prod(ext[3:] - ext[:3]) / self.num_cells)  # placeholder: equal shares

    def compute_cell_geometry(self):
        """Volumes and enclosing boxes of the cells clipped to the domain -- what the reference takes from voro++
        (VoronoiMeshSnapshot::Cell::init, VoronoiMeshSnapshot.cpp:104-135).  The sites are mirrored in the six domain
        walls, which makes every cell of the (unbounded) tessellation of all points equal to the bounded cell."""
        from scipy.spatial import ConvexHull, Voronoi
        ext = np.asarray(self.extent)
        lo, hi = ext[:3], ext[3:]
        pts = [self.sites]
        for ax in range(3):
            for wall in (lo[ax], hi[ax]):
                r = self.sites.copy()
                r[:, ax] = 2.0 * wall - r[:, ax]
                pts.append(r)
        vor = Voronoi(np.concatenate(pts))
        n = len(self.sites)
        boxes, vols = np.empty((n, 6)), np.empty(n)
        for m in range(n):
            reg = vor.regions[vor.point_region[m]]
            if -1 in reg or not reg:
                raise RuntimeError("unbounded Voronoi cell inside the mirrored point set")
            v = np.clip(vor.vertices[reg], lo, hi)
            boxes[m, :3], boxes[m, 3:] = v.min(axis=0), v.max(axis=0)
            vols[m] = ConvexHull(v).volume
        self.cell_extents, self.volumes = boxes, vols
        return self

    def configure(self, engine):
        if self.nbr_offset is None:
            # no tessellation yet (MonteCarloSimulation.deviceSetup): VoronoiMeshSnapshot::buildMesh on the engine's side
            engine.build_voronoi(self.extent, self.sites)
            self.nbr_offset, self.nbr_index, self.volumes, self.cell_extents = engine.read_voronoi()
            return
        engine.set_grid_voronoi(self.extent, self.sites, self.nbr_offset, self.nbr_index)
        if self.cell_extents is not None:
            engine.set_voronoi_extents(self.cell_extents)


# ---------------------------------------------------------------------------------------------------
# sources, SEDs
# ---------------------------------------------------------------------------------------------------

class BlackBodySED:
    """BlackBodySED.cpp:12-65 + PlanckFunction::cdf (PlanckFunction.cpp:31-42)."""

    def __init__(self, temperature):
        self.T = temperature

    def setup(self, source_range):
        lo, hi = source_range
        n = max(100, int(1000.0 * math.log10(hi / lo)))
        self.lambdav = np.exp(math.log(lo) + np.arange(n + 1) * (math.log(hi / lo) / n))
        pv = planck(self.lambdav, self.T)
        self.pv, self.Pv, self.Ltot = cdf2_loglog(self.lambdav, pv)

    def source_fields(self):
        return {"sed_kind": abi.SK_SED_BLACKBODY, "sed_lambda": self.lambdav, "sed_p": self.pv, "sed_P": self.Pv,
                "sed_temperature": self.T, "sed_norm": self.Ltot}

    def specific_luminosity(self, lam):
        return planck(np.asarray(lam, dtype=float), self.T) / self.Ltot


def cdf_loglog_range(inx, inp, lo, hi):
    """NR::cdf<interpolateLogLog>(xv, pv, Pv, inxv, inpv, range), NR.hpp:494-520: the tabulated pdf restricted to [lo, hi]
    with log-log interpolated end points; returns (xv, pv, Pv, norm)."""
    inx, inp = np.asarray(inx, dtype=float), np.asarray(inp, dtype=float)

    def interp(x, i):  # NR::interpolateLogLog between points i-1 and i
        x1, x2, f1, f2 = inx[i - 1], inx[i], inp[i - 1], inp[i]
        if f1 <= 0 or f2 <= 0:
            return f1 if x == x1 else f2 if x == x2 else 0.0
        return f1 * math.exp(math.log(x / x1) / math.log(x2 / x1) * math.log(f2 / f1))

    min_right = int(np.searchsorted(inx, lo, side="right"))
    max_right = int(np.searchsorted(inx, hi, side="left"))
    xv = np.concatenate([[lo], inx[min_right:max_right], [hi]])
    pv = np.empty(len(xv))
    pv[0] = 0.0 if min_right == 0 else interp(lo, min_right)
    pv[1:-1] = inp[min_right:max_right]
    pv[-1] = 0.0 if max_right == len(inx) else interp(hi, max_right)
    pv, Pv, norm = cdf2_loglog(xv, pv)
    return xv, pv, Pv, norm


class ListSED:
    """ListSED.cpp:12-24 + TabulatedSED.cpp:12-35: a tabulated SED (wavelengths in m, specific luminosities per unit
    wavelength in arbitrary units), normalised over the source wavelength range."""

    def __init__(self, wavelengths, specificLuminosities):
        o = np.argsort(wavelengths)
        self.inlam = np.asarray(wavelengths, dtype=float)[o]
        self.inp = np.asarray(specificLuminosities, dtype=float)[o]

    def setup(self, source_range):
        self.lambdav, self.pv, self.Pv, self.norm = cdf_loglog_range(self.inlam, self.inp, *source_range)

    def source_fields(self):
        return {"sed_kind": abi.SK_SED_TABULATED, "sed_lambda": self.lambdav, "sed_p": self.pv, "sed_P": self.Pv}

    def specific_luminosity(self, lam):
        lam = np.asarray(lam, dtype=float)
        return np.exp(np.interp(np.log(lam), np.log(self.inlam), np.log(self.inp))) / self.norm


@dataclass
class PointSource:
    position: Sequence[float]
    sed: BlackBodySED
    luminosity: float  # IntegratedLuminosityNormalization over the source range (W)
    sourceWeight: float = 1.0
    wavelengthBias: float = 0.5
    velocity: Optional[Sequence[float]] = None  # SpecialtySource velocityX/Y/Z (m/s)

    def has_velocity(self):
        return self.velocity is not None and any(self.velocity)

    def fields(self):
        d = {"kind": abi.SK_SRC_POINT, "position": tuple(self.position)}
        if self.has_velocity():
            d.update({"velocity_kind": abi.SK_VEL_CONSTANT, "velocity": tuple(self.velocity)})
        return d


@dataclass
class GeometricSource:
    geometry: object
    sed: BlackBodySED
    luminosity: float
    sourceWeight: float = 1.0
    wavelengthBias: float = 0.5
    velocityMagnitude: float = 0.0  # GeometricSource::velocityMagnitude / velocityDistribution
    velocityDistribution: object = None

    def has_velocity(self):
        return self.velocityDistribution is not None and self.velocityMagnitude != 0

    def fields(self):
        d = {"kind": abi.SK_SRC_GEOMETRIC}
        d.update(self.geometry.source_fields())
        if self.has_velocity():
            d.update(self.velocityDistribution.source_fields(self.velocityMagnitude))
        return d


# ---------------------------------------------------------------------------------------------------
# instruments
# ---------------------------------------------------------------------------------------------------

@dataclass
class Instrument:
    instrumentName: str
    distance: float
    inclination: float = 0.0
    azimuth: float = 0.0
    roll: float = 0.0
    kind: int = abi.SK_INSTR_SED
    radius: float = 0.0
    fieldOfViewX: float = 0.0
    numPixelsX: int = 0
    centerX: float = 0.0
    fieldOfViewY: float = 0.0
    numPixelsY: int = 0
    centerY: float = 0.0
    recordComponents: bool = False
    numScatteringLevels: int = 0
    recordStatistics: bool = False
    wavelengthGrid: Optional[DisjointWavelengthGrid] = None
    redshift: float = 0.0  # FluxRecorder::setObserverFrameRedshift (from the simulation's Cosmology) ...
    luminosityDistance: float = 0.0  # ... with the distances the Cosmology derives from it (used for calibration only)
    angularDiameterDistance: float = 0.0

    def fields(self, grid_index):
        return {"kind": self.kind, "wavelength_grid": grid_index, "inclination": self.inclination,
                "azimuth": self.azimuth, "roll": self.roll, "distance": self.distance, "radius": self.radius,
                "num_pixels_x": self.numPixelsX, "num_pixels_y": self.numPixelsY,
                "field_of_view_x": self.fieldOfViewX, "field_of_view_y": self.fieldOfViewY,
                "center_x": self.centerX, "center_y": self.centerY, "record_components": self.recordComponents,
                "num_scattering_levels": self.numScatteringLevels, "record_statistics": self.recordStatistics,
                "redshift": self.redshift}


def SEDInstrument(**kw):
    return Instrument(kind=abi.SK_INSTR_SED, **kw)


def FrameInstrument(**kw):
    return Instrument(kind=abi.SK_INSTR_FRAME, **kw)


def FullInstrument(**kw):
    return Instrument(kind=abi.SK_INSTR_FULL, **kw)


# ---------------------------------------------------------------------------------------------------
# the simulation object
# ---------------------------------------------------------------------------------------------------

@dataclass
class MonteCarloSimulation:
    """Mirror of MonteCarloSimulation + Configuration for the accelerated modes
    (OligoExtinctionOnly / ExtinctionOnly, with or without a stored radiation field)."""
    sources: List[object]
    medium: GeometricMedium
    grid: object
    instruments: List[Instrument]
    numPackets: float
    oligoWavelengths: Optional[Sequence[float]] = None  # oligochromatic when given
    minWavelength: float = 0.09e-6  # SourceSystem::minWavelength / maxWavelength (panchromatic)
    maxWavelength: float = 100e-6
    defaultWavelengthGrid: Optional[DisjointWavelengthGrid] = None
    radiationFieldWLG: Optional[DisjointWavelengthGrid] = None
    storeRadiationField: bool = False
    forceScattering: bool = True
    explicitAbsorption: bool = False   # PhotonPacketOptions::explicitAbsorption
    minWeightReduction: float = 1e4
    minScattEvents: int = 0
    pathLengthBias: float = 0.5
    sourceBias: float = 0.5
    numDensitySamples: int = 100
    seed: int = 0
    # DustEmission mode (MonteCarloSimulation.hpp:223-273, DustEmissionOptions, SecondaryEmissionOptions, IterationOptions)
    dustEmissionWLG: Optional[DisjointWavelengthGrid] = None   # given => simulationMode DustEmission
    iterateSecondaryEmission: bool = False
    minSecondaryIterations: int = 1
    maxSecondaryIterations: int = 10
    maxFractionOfPrimary: float = 0.01
    maxFractionOfPrevious: float = 0.03
    secondarySpatialBias: float = 0.5
    dustEmissionWavelengthBias: float = 0.5
    storeEmissionRadiationField: bool = False
    secondaryPacketsMultiplier: float = 1.0
    secondaryIterationPacketsMultiplier: float = 1.0
    # dynamic medium state (DynamicStateOptions with a ClearDensityRecipe; IterationOptions): the recipe runs on the host between
    # the segments, on the radiation field the engine hands back, and the new densities go to the engine
    includeHeatingByCMB: bool = False  # DustEmissionOptions::includeHeatingByCMB, with the redshift of the simulation's Cosmology
    cosmologyRedshift: float = 0.0
    clearDensityThreshold: Optional[float] = None   # ClearDensityRecipe::fieldStrengthThreshold; given => the recipe is present
    iteratePrimaryEmission: bool = False            # MonteCarloSimulation::iteratePrimaryEmission
    includePrimaryEmission: bool = False            # IterationOptions::includePrimaryEmission (merged iterations)
    minPrimaryIterations: int = 1
    maxPrimaryIterations: int = 10
    primaryIterationPacketsMultiplier: float = 1.0
    primaryIterationInitialPacketsFraction: float = 1.0
    primaryIterationPacketsRamp: float = 1.0
    setup_seed: int = 12345  # host-side sampling of densities / tree policy (numpy RNG)
    # build the octree and sample the medium state on the engine's side
    # (sk_engine_build_octree / sk_engine_sample_medium) instead of with the numpy code of setup()
    deviceSetup: bool = False
    density: Optional[np.ndarray] = field(default=None, repr=False)
    # further medium components, each with its own material mix (Configuration::hasMultipleConstantSectionMedia); `density`
    # is then [component][cell]
    extraMedia: List[GeometricMedium] = field(default_factory=list)

    @property
    def media(self):
        return [self.medium] + list(self.extraMedia)

    def setup(self):
        """Simulation::setup(): digests the hierarchy into the flat tables the engine needs."""
        rng = np.random.default_rng(self.setup_seed)
        oligo = self.oligoWavelengths is not None
        if oligo:
            self.defaultWavelengthGrid = OligoWavelengthGrid(self.oligoWavelengths)
            self.source_range = self.defaultWavelengthGrid.wavelength_range()  # Configuration.cpp:60-65
            if self.storeRadiationField:
                self.radiationFieldWLG = self.defaultWavelengthGrid  # Configuration.cpp:249
        else:
            self.source_range = (self.minWavelength, self.maxWavelength)
        # wavelength grids addressed by index
        self.grids = []
        for g in [self.defaultWavelengthGrid, self.radiationFieldWLG, self.dustEmissionWLG] \
                + [i.wavelengthGrid for i in self.instruments]:
            if g is not None and all(g is not h for h in self.grids):
                self.grids.append(g)
        # Configuration::simulationWavelengthRange / simulationWavelengths, Configuration.cpp:566-662
        lo, hi = self.source_range
        extra = [md.norm_wavelength for md in self.media if md.norm_wavelength is not None]
        for g in self.grids:
            a, b = g.wavelength_range()
            lo, hi = min(lo, a), max(hi, b)
            extra += list(g.lambdav)
        # kinematics (Configuration.cpp:81, 320, 573): a wide margin for Doppler shifts
        self.hasMovingMedia = any(getattr(md, "has_velocity", lambda: False)() for md in self.media)
        self.hasMovingSources = any(s.has_velocity() for s in self.sources)
        if (self.hasMovingMedia or self.hasMovingSources) and oligo:
            raise ValueError("velocities are refused in oligochromatic simulations (GeometricMedium.cpp:55-60)")
        if self.hasMovingSources or self.hasMovingMedia:
            lo0, hi0 = self.source_range
            if self.dustEmissionWLG is not None:
                a, b = self.dustEmissionWLG.wavelength_range()
                lo0, hi0 = min(lo0, a), max(hi0, b)
            lo, hi = min(lo, lo0 / (1.0 + 1.0 / 3.0)), max(hi, hi0 * (1.0 + 1.0 / 3.0))
        if self.storeRadiationField and not oligo:
            lo, hi = min(lo, 0.09e-6), max(hi, 2000e-6)
        lo, hi = lo / 1.01, hi * 1.01
        for md in self.media:
            md.mix.setup((lo, hi), extra)
        if self.dustEmissionWLG is not None:
            if not self.storeRadiationField or self.radiationFieldWLG is None:
                raise ValueError("DustEmission mode needs a stored radiation field")
            for md in self.media:
                md.mix.precalculate(self.radiationFieldWLG, self.dustEmissionWLG)
        for md in self.media:
            md.setup()
        for s in self.sources:
            s.sed.setup(self.source_range)
        # grid and medium state (MediumSystem.cpp:286-399)
        if self.deviceSetup and isinstance(self.grid, VoronoiMeshSpatialGrid):
            # the tessellation (neighbour lists, volumes, enclosing boxes) is built by the engine in configure()
            # (sk_engine_build_voronoi); the medium state is the density at the sites (numDensitySamples = 1)
            if self.density is None and not isinstance(self.medium, ParticleMedium):
                st = self.grid.sites
                self.density = np.stack([md.number_density(st[:, 0], st[:, 1], st[:, 2]) for md in self.media])
                if not self.extraMedia:
                    self.density = self.density[0]
            self.volume = None
            return self
        if self.deviceSetup:
            if not isinstance(self.grid, (CartesianSpatialGrid, PolicyTreeSpatialGrid)) \
                    or isinstance(self.grid, FileTreeSpatialGrid):
                raise ValueError("deviceSetup needs a Cartesian, policy octree or Voronoi grid")
            if self.extraMedia:
                raise ValueError("deviceSetup samples a single medium component")
            if isinstance(self.grid, CartesianSpatialGrid):
                self.grid.setup([self.medium], self.numDensitySamples, rng)
            self.density = self.volume = None  # both come from the engine in configure()
            return self
        self.grid.setup(self.media, self.numDensitySamples, rng)
        if isinstance(self.grid, VoronoiMeshSpatialGrid):
            if self.dustEmissionWLG is not None and self.grid.cell_extents is None:
                self.grid.compute_cell_geometry()  # emission positions need the cells' enclosing boxes, J their volumes
            self.volume = self.grid.cell_volumes()
            n = self.grid.num_cells
            boxes = None
            if self.density is None:
                # MediumSystem.cpp:286-399 with numDensitySamples = 1: the density at the cell's site
                st = self.grid.sites
                self.density = np.stack([md.number_density(st[:, 0], st[:, 1], st[:, 2]) for md in self.media])
                if not self.extraMedia:
                    self.density = self.density[0]
        else:
            boxes = self.grid.cell_boxes()
            self.volume = np.prod(boxes[:, 3:] - boxes[:, :3], axis=1)
            n = len(boxes)
        nmed = len(self.media)
        dens = np.zeros((nmed, n))
        if self.density is not None:
            # medium state imported from a SpatialCellPropertiesProbe / DensityProbe file of a reference run (tests/golden)
            dens = np.asarray(self.density, dtype=float).reshape(nmed, -1)
            if dens.shape[1] != n:
                raise ValueError("imported density does not match the grid")
        elif self.numDensitySamples == 1:
            c = 0.5 * (boxes[:, :3] + boxes[:, 3:])
            for h, md in enumerate(self.media):
                dens[h] = md.number_density(c[:, 0], c[:, 1], c[:, 2])
        else:
            for _ in range(self.numDensitySamples):
                p = boxes[:, :3] + rng.random((n, 3)) * (boxes[:, 3:] - boxes[:, :3])
                for h, md in enumerate(self.media):
                    dens[h] += md.number_density(p[:, 0], p[:, 1], p[:, 2])
            dens /= self.numDensitySamples
        self.density = dens if self.extraMedia else dens[0]
        self.velocity = None
        if self.hasMovingMedia:
            # MediumSystem.cpp:330-365 with numPropertySamples = 1 and AggregatePolicy::Average: the number-density weighted
            # mean of the components' bulk velocities at the central position of the cell; zero where there is no material
            if boxes is None:
                raise ValueError("moving media need a Cartesian or tree grid here (central positions of Voronoi cells are centroids)")
            c = 0.5 * (boxes[:, :3] + boxes[:, 3:])
            ntot = dens.sum(axis=0)
            v = np.zeros((n, 3))
            for h, md in enumerate(self.media):
                if md.has_velocity():
                    v += dens[h][:, None] * md.bulk_velocity(c)
            with np.errstate(divide="ignore", invalid="ignore"):
                self.velocity = np.where(ntot[:, None] > 0, v / ntot[:, None], 0.0)
        return self

    def config_struct(self, device=0):
        xi = self.pathLengthBias if self.forceScattering else 0.0  # Configuration.cpp:497-504
        if getattr(self, "hasMovingMedia", False):
            xi = 0.0  # "Disabling path length stretching to allow Doppler shifts to be properly sampled", Configuration.cpp:492-498
        force = self.forceScattering or self.storeRadiationField   # Configuration.cpp:476-482
        return abi.SkConfig(self.seed, int(force), self.minScattEvents, xi, self.minWeightReduction, device,
                            int(self.explicitAbsorption))

    def configure(self, engine: abi.Engine):
        """Hands every table to the engine (the extractor step of INTEGRATION.md)."""
        oligo = self.oligoWavelengths is not None
        marks = [("start", time.perf_counter())]

        def mark(name):
            marks.append((name, time.perf_counter()))
        if self.deviceSetup and isinstance(self.grid, VoronoiMeshSpatialGrid):
            self.grid.configure(engine)          # builds the tessellation on the device
            self.volume = self.grid.volumes
            mark("grid")
            if isinstance(self.medium, ParticleMedium) and self.density is None:
                # the cell loop of MediumSystem::setupSelfAfter for a ParticleMedium, on the engine's side
                engine.sample_medium_particles(self.medium.particles, self.medium.density_scale, self.numDensitySamples,
                                               self.grid.num_cells)
            elif self.extraMedia:
                engine.set_media(self.density, self.volume)
            else:
                engine.set_medium(self.density, self.volume)
        elif self.deviceSetup:
            # DensityTreePolicy::constructTree + the cell loop of MediumSystem::setupSelfAfter on the engine's side
            if isinstance(self.medium, ParticleMedium):
                self.grid.configure(engine)      # (the tree policy needs a geometric medium: Cartesian grid, or a tree given)
                mark("grid")
                engine.sample_medium_particles(self.medium.particles, self.medium.density_scale, self.numDensitySamples,
                                               self.grid.num_cells)
            else:
                geom = self.medium.density_geometry()
                if isinstance(self.grid, PolicyTreeSpatialGrid):
                    _, ncells = engine.build_octree(self.grid.extent,
                                                    self.grid.tree_policy(self.numDensitySamples, [self.medium]), [geom])
                    self.grid.first_child = None  # fetched on demand: fetch_device_setup()
                else:
                    self.grid.configure(engine)
                    ncells = self.grid.num_cells
                mark("grid")
                engine.sample_medium(geom, self.numDensitySamples, ncells)
        else:
            self.grid.configure(engine)
            mark("grid")
            if self.extraMedia:
                engine.set_media(self.density, self.volume)
            else:
                engine.set_medium(self.density, self.volume)
        if getattr(self, "velocity", None) is not None:
            engine.set_velocities(self.velocity)
        mark("medium")
        mix = self.medium.mix
        if self.extraMedia:
            engine.set_dustmixes([(md.mix.lambda_border, md.mix.sigma_abs, md.mix.sigma_sca, md.mix.asymmpar, md.mix.mu)
                                  for md in self.media])
        else:
            engine.set_dustmix(mix.lambda_border, mix.sigma_abs, mix.sigma_sca, mix.asymmpar, mix.mu)
        rf = -1
        if self.storeRadiationField:
            rf = [k for k, g in enumerate(self.grids) if g is self.radiationFieldWLG][0]
        engine.set_wavelength_grids([g.table() for g in self.grids], rf)
        srcs = []
        for s in self.sources:
            d = s.fields()
            d.update(s.sed.source_fields())
            d["luminosity"] = s.luminosity
            d["source_weight"] = s.sourceWeight
            if oligo:
                # NormalizedSource.cpp:27-31: always use the bias distribution; OligoWavelengthDistribution.cpp:13-38
                g = self.defaultWavelengthGrid
                d.update({"wavelength_bias": 1.0, "bias_kind": abi.SK_BIAS_OLIGO, "oligo_lambda": g.lambdav,
                          "oligo_probability": 1.0 / g.num_bins / g.dlambdav[0]})
            else:
                d.update({"wavelength_bias": s.wavelengthBias, "bias_kind": abi.SK_BIAS_LOGUNIFORM,
                          "bias_min": self.source_range[0], "bias_max": self.source_range[1]})
            srcs.append(d)
        engine.set_sources(srcs, self.sourceBias)
        mark("tables")
        instr = []
        for i in self.instruments:
            g = i.wavelengthGrid if (i.wavelengthGrid is not None and not oligo) else self.defaultWavelengthGrid
            instr.append(i.fields([k for k, h in enumerate(self.grids) if h is g][0]))
        engine.set_instruments(instr, self.dustEmissionWLG is not None)
        mark("instruments")
        if self.dustEmissionWLG is not None:
            mix, eg = self.medium.mix, self.dustEmissionWLG
            lo, hi = eg.wavelength_range()
            # the CMB source term of the energy balance, EquilibriumDustEmissionCalculator.cpp:37-44 (Constants::Tcmb = 2.725 K)
            cmb = planck(self.radiationFieldWLG.lambdav, 2.725 * (1.0 + self.cosmologyRedshift)) if self.includeHeatingByCMB else None
            if self.extraMedia:
                engine.set_secondary_media([k for k, h in enumerate(self.grids) if h is eg][0], self.secondarySpatialBias,
                                           self.dustEmissionWavelengthBias, lo, hi,
                                           [(md.mix.Tv, md.mix.planck_abs, md.mix.rf_sigma_abs, md.mix.em_sigma_abs)
                                            for md in self.media], rf_cmb=cmb)
            else:
                engine.set_secondary([k for k, h in enumerate(self.grids) if h is eg][0], self.secondarySpatialBias,
                                     self.dustEmissionWavelengthBias, lo, hi, mix.Tv, mix.planck_abs, mix.rf_sigma_abs,
                                     mix.em_sigma_abs, rf_cmb=cmb)
        mark("secondary")
        # seconds spent per group of engine calls (bench.py reports them under e2e.parts)
        self.last_configure_parts = {marks[k][0]: marks[k][1] - marks[k - 1][1] for k in range(1, len(marks))}
        return engine

    def fetch_device_setup(self, engine: abi.Engine):
        """Reads back what configure() built on the engine's side (deviceSetup): node list, densities, volumes."""
        if isinstance(self.grid, PolicyTreeSpatialGrid):
            self.grid.adopt(engine.read_octree())
        self.density, self.volume = engine.read_medium()
        return self

    def run(self, engine: abi.Engine, first=0, count=None, stream_id=0, comm=None):
        """MonteCarloSimulation::runSimulation, MonteCarloSimulation.cpp:58-100: primary emission and, in DustEmission
        mode, the secondary emission iterations and the final secondary emission.  `comm` (skirt9_b200.parallel.Comm)
        shards the histories of every segment over the ranks and performs the reference's reductions."""
        dynamic = self.clearDensityThreshold is not None                      # Configuration.cpp:139-226
        primary_iterations = self.iteratePrimaryEmission and dynamic
        merged = (self.dustEmissionWLG is not None and self.iterateSecondaryEmission and self.includePrimaryEmission and dynamic)
        if primary_iterations:
            self.run_primary_emission_iterations(engine, stream_id + 3000, comm)
        if merged:
            self.run_merged_emission_iterations(engine, stream_id + 4000, comm)
        self.run_primary_emission(engine, first, count, stream_id, comm)
        if self.dustEmissionWLG is not None:
            if self.iterateSecondaryEmission and not merged:
                self.run_secondary_emission_iterations(engine, stream_id + 1000, comm)
            self.run_secondary_emission(engine, stream_id + 2000, comm)
        if comm is not None:
            comm.allreduce_detectors(engine)   # FluxRecorder::calibrateAndWrite, FluxRecorder.cpp:487-493

    def run_primary_emission(self, engine, first=0, count=None, stream_id=0, comm=None):
        """MonteCarloSimulation::runPrimaryEmission, MonteCarloSimulation.cpp:104-138."""
        n = int(self.numPackets)
        store = self.storeRadiationField
        if store:
            engine.clear_rf(True)
        engine.prepare_primary(n)
        if comm is not None:
            first, count = comm.block(engine, n)
        engine.run_segment(first, n if count is None else count, primary=True, peel=True, store=store,
                           stream_id=stream_id)
        if store:
            if comm is not None:
                comm.allreduce_rf(engine, True)
            engine.communicate_rf(True)

    # -- dynamic medium state ------------------------------------------------------------------------
    def update_dynamic_state(self, engine):
        """MediumSystem::updateDynamicStateRecipes (MediumSystem.cpp:1498-1558) with a ClearDensityRecipe
        (ClearDensityRecipe.cpp:17-35): a cell whose radiation field strength U = Sum_l J_l dlambda_l / 1.7623e-6 W/m2/sr exceeds
        the threshold loses its material.  Returns (cells updated, converged); the new densities go to the engine."""
        rf = engine.read_rf(0)
        if self.dustEmissionWLG is not None:
            rf = rf + engine.read_rf(1)                      # MediumSystem::radiationField, MediumSystem.cpp:1360-1366
        U = rf.sum(axis=1) / (4 * math.pi * self.volume) / 1.7623e-06
        dens = np.asarray(self.density, dtype=float)
        total = dens if dens.ndim == 1 else dens.sum(axis=0)
        hit = (U > self.clearDensityThreshold)
        updated = int(np.count_nonzero(hit & (total > 0))) if dens.ndim == 1 else int(np.count_nonzero(hit[None, :] & (dens > 0)))
        cells = int(np.count_nonzero(hit & (total > 0)))
        if dens.ndim == 1:
            dens = np.where(hit, 0.0, dens)
        else:
            dens = np.where(hit[None, :], 0.0, dens)
        self.density = dens
        if self.extraMedia:
            engine.set_media(self.density, self.volume)
        else:
            engine.set_medium(self.density, self.volume)
        if getattr(self, "velocity", None) is not None:
            engine.set_velocities(self.velocity)
        return cells, updated == 0                            # DynamicStateRecipe::endUpdate with maxNotConvergedCells = 0

    @staticmethod
    def _loop_ends(converged, it, min_iters, max_iters):
        """logLoopConvergence, MonteCarloSimulation.cpp:233-261."""
        return (converged and it >= min_iters) or (not converged and it >= max_iters)

    def run_primary_emission_iterations(self, engine, stream_id=3000, comm=None):
        """MonteCarloSimulation::runPrimaryEmissionIterations, MonteCarloSimulation.cpp:266-330."""
        n_it = self.numPackets * self.primaryIterationPacketsMultiplier
        min_n = max(1.0, n_it * self.primaryIterationInitialPacketsFraction)
        max_n = max(1.0, n_it)
        self.primary_iterations = []
        it, prev = 0, 0
        while True:
            it += 1
            n = int(min(max_n, min_n * self.primaryIterationPacketsRamp ** (it - 1)))
            if n != prev:
                engine.prepare_primary(n)
                prev = n
            engine.clear_rf(True)
            first, count = comm.block(engine, n) if comm is not None else (0, n)
            engine.run_segment(first, count, primary=True, peel=False, store=True, stream_id=stream_id + it)
            if comm is not None:
                comm.allreduce_rf(engine, True)
            engine.communicate_rf(True)
            updated, converged = self.update_dynamic_state(engine)
            self.primary_iterations.append({"iteration": it, "packets": n, "updated_cells": updated, "converged": converged})
            if self._loop_ends(converged, it, self.minPrimaryIterations, self.maxPrimaryIterations):
                break

    def run_merged_emission_iterations(self, engine, stream_id=4000, comm=None):
        """MonteCarloSimulation::runMergedEmissionIterations, MonteCarloSimulation.cpp:407-496."""
        n1 = int(self.numPackets * self.primaryIterationPacketsMultiplier)
        n2 = int(self.numPackets * self.secondaryIterationPacketsMultiplier)
        engine.prepare_primary(n1)
        self.convergence = []
        prev, it = 0.0, 0
        while True:
            it += 1
            engine.clear_rf(True)
            first, count = comm.block(engine, n1) if comm is not None else (0, n1)
            engine.run_segment(first, count, primary=True, peel=False, store=True, stream_id=stream_id + 2 * it)
            if comm is not None:
                comm.allreduce_rf(engine, True)
            engine.communicate_rf(True)
            engine.clear_rf(False)
            lum = engine.prepare_secondary(n2)
            if not lum > 0:
                return
            first, count = comm.block(engine, n2) if comm is not None else (0, n2)
            engine.run_segment(first, count, primary=False, peel=False, store=True, stream_id=stream_id + 2 * it + 1)
            if comm is not None:
                comm.allreduce_rf(engine, False)
            engine.communicate_rf(False)
            updated, converged = self.update_dynamic_state(engine)
            # (the reference evaluates the absorbed luminosities after the update, with the new densities)
            Lprim, Lseco = engine.absorbed_luminosity(True), engine.absorbed_luminosity(False)
            with np.errstate(divide="ignore", invalid="ignore"):
                dust = (Lprim <= 0 or Lseco <= 0 or Lseco / Lprim < self.maxFractionOfPrimary
                        or abs((Lseco - prev) / Lseco) < self.maxFractionOfPrevious)
            prev = Lseco
            converged = bool(converged and dust)
            self.convergence.append({"iteration": it, "dust_luminosity": lum, "absorbed_primary": Lprim,
                                     "absorbed_secondary": Lseco, "updated_cells": updated, "converged": converged})
            if self._loop_ends(converged, it, self.minSecondaryIterations, self.maxSecondaryIterations):
                break

    def run_secondary_emission(self, engine, stream_id=2000, comm=None):
        """MonteCarloSimulation::runSecondaryEmission, MonteCarloSimulation.cpp:142-173."""
        store = self.storeEmissionRadiationField
        if store:
            engine.clear_rf(False)
        n = int(self.numPackets * self.secondaryPacketsMultiplier)
        self.dust_luminosity = engine.prepare_secondary(n)
        if self.dust_luminosity > 0:
            first, count = comm.block(engine, n) if comm is not None else (0, n)
            engine.run_segment(first, count, primary=False, peel=True, store=store, stream_id=stream_id)
        if store:
            if comm is not None:
                comm.allreduce_rf(engine, False)
            engine.communicate_rf(False)

    def run_secondary_emission_iterations(self, engine, stream_id=1000, comm=None):
        """MonteCarloSimulation::runSecondaryEmissionIterations (.cpp:335-403) with DustAbsorptionConvergence
        (.cpp:180-227) and logLoopConvergence (.cpp:233-261).  Records the log values in self.convergence."""
        n = int(self.numPackets * self.secondaryIterationPacketsMultiplier)
        self.convergence = []
        prev = 0.0
        it = 0
        while True:
            it += 1
            engine.clear_rf(False)
            lum = engine.prepare_secondary(n)
            if not lum > 0:
                return
            first, count = comm.block(engine, n) if comm is not None else (0, n)
            engine.run_segment(first, count, primary=False, peel=False, store=True, stream_id=stream_id + it)
            if comm is not None:
                comm.allreduce_rf(engine, False)
            engine.communicate_rf(False)
            Lprim, Lseco = engine.absorbed_luminosity(True), engine.absorbed_luminosity(False)
            with np.errstate(divide="ignore", invalid="ignore"):
                converged = (Lprim <= 0 or Lseco <= 0 or Lseco / Lprim < self.maxFractionOfPrimary
                             or abs((Lseco - prev) / Lseco) < self.maxFractionOfPrevious)
            prev = Lseco
            self.convergence.append({"iteration": it, "dust_luminosity": lum, "absorbed_primary": Lprim,
                                     "absorbed_secondary": Lseco, "converged": bool(converged)})
            if converged and it >= self.minSecondaryIterations:
                break
            if not converged and it >= self.maxSecondaryIterations:
                break

    # -- calibration of raw tallies to the units the reference writes ------------------------------
    def sed_flux_density(self, engine, instrument=0, component=abi.SK_COMP_TOTAL):
        """FluxRecorder::calibrateAndWrite, FluxRecorder.cpp:672-708: F_lambda = L/(4 pi d^2)/dlambda (W/m2/m);
        returned as F_nu in Jy (fluxOutputStyle Frequency, Units.cpp:557-567)."""
        i = self.instruments[instrument]
        oligo = self.oligoWavelengths is not None
        g = i.wavelengthGrid if (i.wavelengthGrid is not None and not oligo) else self.defaultWavelengthGrid
        L = engine.read_sed(instrument, component)
        d = i.luminosityDistance if i.redshift else i.distance  # FluxRecorder.cpp:503
        flam = L / (4 * math.pi * d ** 2) / g.dlambdav
        return flam * g.lambdav ** 2 / C_LIGHT * 1e26

    def surface_brightness(self, engine, instrument=0, component=abi.SK_COMP_TOTAL):
        """FluxRecorder::calibrateAndWrite, FluxRecorder.cpp:503-506,740-749: per-pixel F_lambda / Omega_pixel with
        Omega = 4 atan(dx/2d) atan(dy/2d); returned as [ell][j][i] in MJy/sr (fluxOutputStyle Frequency,
        Units.cpp:599-609), the layout and unit of the reference's FITS cubes."""
        i = self.instruments[instrument]
        oligo = self.oligoWavelengths is not None
        g = i.wavelengthGrid if (i.wavelengthGrid is not None and not oligo) else self.defaultWavelengthGrid
        L = engine.read_ifu(instrument, component).reshape(g.num_bins, i.numPixelsY, i.numPixelsX)
        omega = 4.0 * math.atan(0.5 * i.fieldOfViewX / i.numPixelsX / i.distance) \
            * math.atan(0.5 * i.fieldOfViewY / i.numPixelsY / i.distance)
        flam = L / (4 * math.pi * i.distance ** 2) / g.dlambdav[:, None, None] / omega
        return flam * (g.lambdav ** 2)[:, None, None] / C_LIGHT * 1e26 * 1e-6

    def mean_intensity_nu(self, engine, which=0):
        """J_nu in W/m2/Hz/sr as RadiationFieldProbe writes it with fluxOutputStyle Frequency (Units.cpp:641-651)."""
        g = self.radiationFieldWLG
        return self.mean_intensity(engine, which) * (g.lambdav ** 2)[None, :] / C_LIGHT

    def mean_intensity(self, engine, which=0):
        """MediumSystem::meanIntensity, MediumSystem.cpp:1370-1380: J_lambda = rf/(4 pi V dlambda)."""
        rf = engine.read_rf(which)
        g = self.radiationFieldWLG
        return rf / (4 * math.pi * self.volume[:, None] * g.dlambdav[None, :])
