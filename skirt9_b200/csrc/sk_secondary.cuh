// sk_secondary.cuh -- dust re-emission on the device: SecondarySourceSystem / DustSecondarySource
// (SKIRT/core/SecondarySourceSystem.cpp:84-142, DustSecondarySource.cpp:26-146,511-581) with the
// EquilibriumDustEmissionCalculator (EquilibriumDustEmissionCalculator.cpp:120-150).
//
// The reference computes a cell's emission spectrum lazily, per thread, the first time a packet is launched from it
// (DustCellEmission::calculateIfNeeded, DustSecondarySource.cpp:187-285).  Here the spectra of ALL emitting cells are
// computed by one batched kernel per segment (one thread per cell, ~100 wavelengths each) and kept in HBM
// ([ncells][N_em+2] pdf and cdf), so that the launch of a history is a pair of binary searches.
#pragma once
#include "sk_blocks.cuh"

// MediumSystem::dustLuminosity, MediumSystem.cpp:1452-1462 (single dust medium with constant cross sections:
// opacityAbs = n * sigma_abs(lambda_ell), radiationField = rf1 + rf2)
template <int GRID>
__global__ void sk_dust_luminosity_kernel(const SkDevModel M)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M.ncells) return;
    double nh[SK_MAX_MEDIA];
    for (int h = 0; h < M.nmed; ++h) nh[h] = sk_component_density<GRID>(M, m, h);
    double Labs = 0.;
    for (int ell = 0; ell < M.nrf; ++ell)
    {
        // MediumSystem::opacityAbs(lambda, m, Dust): the sum over the dust components (MediumSystem.cpp:619-630)
        double opacity = nh[0] * M.sec_kabs_rf[ell];
        for (int h = 1; h < M.nmed; ++h) opacity += nh[h] * M.sec_kabs_rf[h * M.nrf + ell];
        double rf = 0.;
        rf += M.rf1[SK_RF_INDEX(M, m, ell)];
        rf += M.rf2[SK_RF_INDEX(M, m, ell)];
        Labs += opacity * rf;
    }
    M.sec_Lv[m] = Labs;
}

// SpecialFunctions::gln, SpecialFunctions.cpp:798-811
__device__ __forceinline__ double sk_gln(double p, double x)
{
    const double q = 1.0 - p;
    if (q == 0.0)
        return log(x);
    else if (fabs(q) < 1e-3)
    {
        double lnx = log(x);
        double s = q * lnx;
        return lnx * (1.0 + 0.5 * s + 1.0 / 6.0 * s * s + 1.0 / 24.0 * s * s * s);
    }
    else
        return (pow(x, q) - 1.0) / q;
}

// The normalised emission spectrum and its cumulative distribution for every emitting cell:
// MediumSystem::meanIntensity (MediumSystem.cpp:1370-1380) -> EquilibriumDustEmissionCalculator::equilibriumTemperature
// (.cpp:120-131, NR::clampedValue<interpolateLinLin>, NR.hpp:391-399) -> emissivity (.cpp:135-150) times the number
// density (DustMix::emissionSpectrum, DustMix.cpp:650-653) -> NR::cdf<interpolateLogLog> over the range of the emission
// grid (NR.hpp:494-520 + NR::cdf2, NR.cpp:25-60).  sec_Lv holds the normalised luminosities here (only their sign is used).
template <int GRID>
__global__ void sk_emission_spectrum_kernel(const SkDevModel M)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M.ncells) return;
    const int nem = M.sec_nem, nrf = M.nrf, nT = M.sec_nT;
    double* pv = M.sec_pv + (size_t)m * nem;
    double* Pv = M.sec_Pv + (size_t)m * nem;
    if (!(M.sec_Lv[m] > 0))
    {
        for (int i = 0; i < nem; ++i)
        {
            pv[i] = 0.;
            Pv[i] = 0.;
        }
        return;
    }
    const SkDevWlg& rfg = M.wlg[M.rf_grid];
    const double factor = 1. / (4. * M_PI * M.volume[m]);
    const double* x = M.sec_lambda;
    // MediumSystem::dustEmissionSpectrum (MediumSystem.cpp:1466-1476): the sum over the dust components of the number
    // density times the emissivity of the component's mix at its own equilibrium temperature
    for (int h = 0; h < M.nmed; ++h)
    {
        const double* __restrict__ rfsig = M.sec_rfsig + h * nrf;
        const double* __restrict__ planckabs = M.sec_planckabs + h * nT;
        const double* __restrict__ emsig = M.sec_emsig + h * nem;
        double inputabs = 0.;
        for (int ell = 0; ell < nrf; ++ell)
        {
            double rf = 0.;
            rf += M.rf1[SK_RF_INDEX(M, m, ell)];
            rf += M.rf2[SK_RF_INDEX(M, m, ell)];
            double J = rf * factor / rfg.dlambda[ell];
            inputabs += rfsig[ell] * (J + M.sec_cmb[ell]) * rfg.dlambda[ell];  // (Jv + _Bcmbv), .cpp:123
        }
        double T = 0.;
        if (inputabs > 0.)
        {
            int i = inputabs == planckabs[nT - 1] ? nT - 2 : sk_locate_basic(planckabs, inputabs, nT);
            if (i < 0)
                T = M.sec_T[0];
            else if (i >= nT - 1)
                T = M.sec_T[nT - 1];
            else
                T = sk_interp_linlin(inputabs, planckabs[i], planckabs[i + 1], M.sec_T[i], M.sec_T[i + 1]);
        }
        const double n = sk_component_density<GRID>(M, m, h);
        if (h == 0)
            for (int i = 0; i < nem; ++i) pv[i] = n * (emsig[i] * sk_planck(x[i], T));
        else
            for (int i = 0; i < nem; ++i) pv[i] += n * (emsig[i] * sk_planck(x[i], T));
    }
    {
        double first = sk_interp_loglog(x[0], x[0], x[1], pv[0], pv[1]);
        double last = sk_interp_loglog(x[nem - 1], x[nem - 2], x[nem - 1], pv[nem - 2], pv[nem - 1]);
        pv[0] = first;
        pv[nem - 1] = last;
        Pv[0] = 0.;
        double P = 0.;
        for (int i = 0; i != nem - 1; ++i)
        {
            double area = 0.;
            if (pv[i] > 0 && pv[i + 1] > 0)
            {
                double alpha = log(pv[i + 1] / pv[i]) / log(x[i + 1] / x[i]);
                area = pv[i] * x[i] * sk_gln(-alpha, x[i + 1] / x[i]);
            }
            P = P + area;
            Pv[i + 1] = P;
        }
        const double norm = P;
        if (norm > 0.)
            for (int i = 0; i < nem; ++i)
            {
                pv[i] /= norm;
                Pv[i] /= norm;
            }
    }
}

// SecondarySourceSystem::launch (SecondarySourceSystem.cpp:130-142) + DustSecondarySource::launch
// (DustSecondarySource.cpp:511-581) without velocities and polarisation; SpatialGrid::randomPositionInCell
// (TreeSpatialGrid.cpp:125-128, CartesianSpatialGrid.cpp:80-83) = Random::position(box) (Random.cpp:168-176);
// VoronoiMeshSpatialGrid::randomPositionInCell = VoronoiMeshSnapshot::generatePosition(m) (VoronoiMeshSnapshot.cpp:976-989).
template <int GRID>
__device__ __noinline__ void sk_launch_secondary(const SkDevModel* __restrict__ Mg, const SkSmemTables& T, SkRng& g,
                                                 unsigned long long history, SkLaunch& pp)
{
    const SkDevModel& M = *Mg;
    const int nem = M.sec_nem;
    int lo = 0, hi = M.ncells + 1;  // std::upper_bound(_Iv, historyIndex) - 1
    while (lo < hi)
    {
        int mid = (lo + hi) >> 1;
        if (history < M.sec_Iv[mid])
            hi = mid;
        else
            lo = mid + 1;
    }
    const int m = lo - 1;  // launch order = cell order for the AllCellsLibrary (AllCellsLibrary.cpp:26-32)
    const double ws = M.sec_ws[m];
    const double* xv = M.sec_lambda;
    const double* pv = M.sec_pv + (size_t)m * nem;
    const double* Pv = M.sec_Pv + (size_t)m * nem;
    double lambda, w;
    const double xi = M.sec_xi;
    if (!xi)
    {
        lambda = sk_sample_cdf_loglog(g, xv, pv, Pv, nem);
        w = 1.;
    }
    else
    {
        const double logMin = log(M.sec_bias_min);
        const double logWidth = log(M.sec_bias_max) - log(M.sec_bias_min);
        if (sk_uniform(g) > xi)
            lambda = sk_sample_cdf_loglog(g, xv, pv, Pv, nem);
        else
            lambda = exp(logMin + logWidth * sk_uniform(g));  // DefaultWavelengthDistribution.cpp:37-40
        double sl = 0.;  // NR::value<interpolateLogLog>, NR.hpp:372-378
        int i = sk_locate_fail(xv, nem, lambda);
        if (i >= 0) sl = sk_interp_loglog(lambda, xv[i], xv[i + 1], pv[i], pv[i + 1]);
        if (!sl)
            w = 0.;
        else
        {
            double b;
            if (lambda >= M.sec_bias_min * (1 - 1e-14) && lambda <= M.sec_bias_max * (1 + 1e-14))  // Range.hpp:56
                b = 1. / (logWidth * lambda);
            else
                b = 0.;
            w = sl / ((1 - xi) * sl + xi * b);
        }
    }
    double b0, b1, b2, b3, b4, b5;
    if (GRID == 3)
    {
        const double* b = M.vbox + 6 * (size_t)m;  // VoronoiMeshSnapshot::Cell is a Box: the cell's enclosing box
        b0 = b[0];
        b1 = b[1];
        b2 = b[2];
        b3 = b[3];
        b4 = b[4];
        b5 = b[5];
    }
    else if (GRID == 1)
    {
        int k = m % M.nz, j = (m / M.nz) % M.ny, i = m / (M.nz * M.ny);
        b0 = T.X[i];
        b1 = T.Y[j];
        b2 = T.Z[k];
        b3 = T.X[i + 1];
        b4 = T.Y[j + 1];
        b5 = T.Z[k + 1];
    }
    else
    {
        const uint4 c = reinterpret_cast<const uint4*>(M.cell_coord)[m];
        const int size = 1 << (M.maxlevel - (int)c.w);
        b0 = T.X[c.x];
        b1 = T.Y[c.y];
        b2 = T.Z[c.z];
        b3 = T.X[c.x + size];
        b4 = T.Y[c.y + size];
        b5 = T.Z[c.z + size];
    }
    if (GRID == 3)
    {
        // VoronoiMeshSnapshot::generatePosition(m), VoronoiMeshSnapshot.cpp:976-989: random points in the enclosing box
        // until one is closest to site m among the sites of m's neighbours (isPointClosestTo, .cpp:848-856)
        const double4 sm = sk_ld_rec(&M.vrec[m]);
        const long long i0 = M.vnbr_off[m], i1 = M.vnbr_off[m + 1];
        bool found = false;
        for (int it = 0; it < 10000 && !found; ++it)
        {
            const double ux = sk_uniform(g), uy = sk_uniform(g), uz = sk_uniform(g);
            const double x = b0 + ux * (b3 - b0);
            const double y = b1 + uy * (b4 - b1);
            const double z = b2 + uz * (b5 - b2);
            double dx = x - sm.x, dy = y - sm.y, dz = z - sm.z;
            const double target = dx * dx + dy * dy + dz * dz;
            found = true;
            for (long long i = i0; i < i1; ++i)
            {
                const int id = __ldg(&M.vnbr[i]);
                if (id < 0) continue;
                const double4 t = sk_ld_rec(&M.vrec[id]);
                double ex = x - t.x, ey = y - t.y, ez = z - t.z;
                if (ex * ex + ey * ey + ez * ez < target)
                {
                    found = false;
                    break;
                }
            }
            pp.rx = x;
            pp.ry = y;
            pp.rz = z;
        }
        if (!found)
        {
            // the reference throws a fatal error here; the engine emits from the site, which lies in the cell
            pp.rx = sm.x;
            pp.ry = sm.y;
            pp.rz = sm.z;
        }
    }
    else
    {
        const double ux = sk_uniform(g), uy = sk_uniform(g), uz = sk_uniform(g);
        pp.rx = b0 + ux * (b3 - b0);  // Box::fracPos, SKIRT/utils/Box.hpp:151-154
        pp.ry = b1 + uy * (b4 - b1);
        pp.rz = b2 + uz * (b5 - b2);
    }
    sk_random_direction(g, pp.kx, pp.ky, pp.kz);
    const double L = M.sec_Lpp * 1.;  // _Lv[s]/_Wv[s] = 1 for the single secondary source
    pp.lambda = lambda;
    pp.W = (L * ws * w) * lambda;
    pp.ilam = sk_locate_clip(M.lam_border, M.nlam, lambda);
    // the bulk velocity of the emitting cell (DustSecondarySource.cpp:271, 562-580)
    pp.vx = M.vel ? M.vel[m].x : 0.;
    pp.vy = M.vel ? M.vel[m].y : 0.;
    pp.vz = M.vel ? M.vel[m].z : 0.;
}
