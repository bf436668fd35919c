// sk_lifecycle.cuh -- the photon life cycle on the device: MonteCarloSimulation::performLifeCycle
// (SKIRT/core/MonteCarloSimulation.cpp:538-613) and everything it calls.
//
// Execution model: a persistent kernel in which ONE WARP OWNS A POOL of SK_POOL in-flight
// photon packets (state in a per-warp structure-of-arrays region of HBM/L2).  The warp advances its whole pool in
// lock-step through the stages of the reference's forced-scattering loop; stages are of two kinds:
//   * event stages (launch, peel-off set-up and detection, scattering, optical-depth sampling, interaction): every
//     lane handles one packet of the pool, all lanes execute the same code -> no divergence;
//   * trace stages (forward path to the boundary, re-walk to the interaction point, peel-off path to an observer):
//     lanes are NOT bound to packets; each lane walks one ray cell by cell and, when its ray ends, immediately takes
//     the next unprocessed ray of the stage from the warp's list -> the cell-crossing loop, which is >95 % of the work,
//     stays converged although path lengths vary from 1 to several hundred segments.
// Paths are never materialised (the reference stores vector<Segment>, SpatialGridPath.hpp:93-115): the forward path
// is walked once for the total optical depth (radiation-field deposits fused in) and re-walked up to the sampled
// interaction point; both walks use identical arithmetic so they see identical segments.
#pragma once
#include "sk_device.cuh"

#ifndef SK_POOL
#define SK_POOL 256
#endif

// per-packet state: field-major arrays of SK_POOL entries per warp
enum {
    D_RX, D_RY, D_RZ, D_KX, D_KY, D_KZ, D_LAMBDA, D_W, D_LTHR, D_SIGEXT, D_TAUPATH, D_TAUINT, D_STOT, D_SINT,
    D_PEELW, D_PTAU, D_LIMIT, D_HISTW0, SK_ND = D_HISTW0 + SK_MAX_INSTR
};
enum {
    I_HLO, I_HHI, I_DRAW, I_NSCATT, I_STATE, I_ILAM, I_M, I_IX, I_IY, I_IZ, I_LEV, I_MINT, I_MIX, I_MIY, I_MIZ,
    I_MLEV, I_NSEG, I_HELL0, SK_NI = I_HELL0 + SK_MAX_INSTR
};
// I_STATE bits
#define SK_ST_LIVE 1
#define SK_ST_SCATTER 2    // a scattering event is pending (peel-off of kind "scattering", then new direction)
#define SK_ST_FOUND 4      // non-forced propagation: interaction point found

struct SkPoolView {
    double* d;
    int32_t* i;
    __device__ __forceinline__ double& D(int f, int s) const { return d[f * SK_POOL + s]; }
    __device__ __forceinline__ int32_t& I(int f, int s) const { return i[f * SK_POOL + s]; }
};

struct SkLocalCounters {
    unsigned int packets, fwd_paths, fwd_segs, replay_segs, peel_paths, peel_segs, scatt, rf, det, fallbacks;
};

struct SkSmemTables {
    const double *X, *Y, *Z;
};

struct SkCellPos {
    int m;           // cell index, -1 = outside / unknown
    int ix, iy, iz;  // octree: lattice coordinates of the lower corner; Cartesian: bin indices i,j,k
    int lev;
};

// ---------------------------------------------------------------------------------------------------
// Geometry helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool sk_box_contains(const double* b, double x, double y, double z)
{
    return x >= b[0] && x <= b[3] && y >= b[1] && y <= b[4] && z >= b[2] && z <= b[5];  // Box.hpp:99-109
}
__device__ __forceinline__ bool sk_box_strictly_inside(const double* b, double x, double y, double z)
{
    return x > b[0] && x < b[3] && y > b[1] && y < b[4] && z > b[2] && z < b[5];
}

// PathSegmentGenerator::moveInside, SKIRT/utils/PathSegmentGenerator.cpp:11-112
__device__ __noinline__ bool sk_move_inside(double& rx, double& ry, double& rz, double kx, double ky, double kz,
                                            const double* box, double eps, double& cumds_out)
{
    double cumds = 0.;
    if (rx <= box[0])
    {
        if (kx <= 0.0) return false;
        double ds = (box[0] - rx) / kx;
        rx = box[0] + eps;
        ry += ky * ds;
        rz += kz * ds;
        cumds += ds;
    }
    else if (rx >= box[3])
    {
        if (kx >= 0.0) return false;
        double ds = (box[3] - rx) / kx;
        rx = box[3] - eps;
        ry += ky * ds;
        rz += kz * ds;
        cumds += ds;
    }
    if (ry <= box[1])
    {
        if (ky <= 0.0) return false;
        double ds = (box[1] - ry) / ky;
        rx += kx * ds;
        ry = box[1] + eps;
        rz += kz * ds;
        cumds += ds;
    }
    else if (ry >= box[4])
    {
        if (ky >= 0.0) return false;
        double ds = (box[4] - ry) / ky;
        rx += kx * ds;
        ry = box[4] - eps;
        rz += kz * ds;
        cumds += ds;
    }
    if (rz <= box[2])
    {
        if (kz <= 0.0) return false;
        double ds = (box[2] - rz) / kz;
        rx += kx * ds;
        ry += ky * ds;
        rz = box[2] + eps;
        cumds += ds;
    }
    else if (rz >= box[5])
    {
        if (kz >= 0.0) return false;
        double ds = (box[5] - rz) / kz;
        rx += kx * ds;
        ry += ky * ds;
        rz = box[5] - eps;
        cumds += ds;
    }
    if (!sk_box_contains(box, rx, ry, rz)) return false;
    cumds_out = cumds;
    return true;
}

// Octree cell location: TreeNode::leafChild (TreeNode.cpp:65-76) + OctTreeNode::child (OctTreeNode.cpp:37-42) on the
// integer lattice: a node is (ix,iy,iz,level); its centre (CHILD_0->rmax) is the lattice border at +half size.
// `fc` is the node id of the first child of the node to descend from.
__device__ __forceinline__ void sk_tree_descend(const int32_t* __restrict__ node_child, int maxlevel,
                                                const SkSmemTables& T, int fc, int ix, int iy, int iz, int lev,
                                                double x, double y, double z, SkCellPos& out)
{
    while (fc >= 0)
    {
        int half = 1 << (maxlevel - lev - 1);
        int l = 0;
        if (!(x < T.X[ix + half]))
        {
            l |= 1;
            ix += half;
        }
        if (!(y < T.Y[iy + half]))
        {
            l |= 2;
            iy += half;
        }
        if (!(z < T.Z[iz + half]))
        {
            l |= 4;
            iz += half;
        }
        lev++;
        fc = __ldg(&node_child[fc + l]);
    }
    out.m = -(fc + 1);
    out.ix = ix;
    out.iy = iy;
    out.iz = iz;
    out.lev = lev;
}

// Cold path: locates the cell holding (x,y,z) from scratch.  Takes the model through a pointer to its copy in
// global memory so that the kernel-parameter copy never has its address taken (that would force it into local memory).
template <int GRID>
__device__ __noinline__ void sk_locate(const SkDevModel* __restrict__ Mg, const SkSmemTables& T, double x, double y,
                                       double z, SkCellPos& out)
{
    if (!sk_box_contains(Mg->ext, x, y, z))
    {
        out.m = -1;
        return;
    }
    if (GRID == 1)
    {
        out.ix = sk_locate_clip(T.X, Mg->nx + 1, x);  // CartesianSpatialGrid.cpp:105-107
        out.iy = sk_locate_clip(T.Y, Mg->ny + 1, y);
        out.iz = sk_locate_clip(T.Z, Mg->nz + 1, z);
        out.lev = 0;
        out.m = out.iz + Mg->nz * out.iy + Mg->nz * Mg->ny * out.ix;
    }
    else
        sk_tree_descend(Mg->node_child, Mg->maxlevel, T, __ldg(&Mg->node_child[0]), 0, 0, 0, 0, x, y, z, out);
}

// Cold path of the octree step: the full neighbour search of TreeSpatialGrid.cpp:190-207 for the situations the
// fast path does not decide itself (domain boundary, near-ties between exit walls, grazing directions):
// the leaf containing the new position (TreeNode::neighbor + root()->leafChild fall-back), the nextafter escape
// when stuck in the same cell, and termination.
__device__ __noinline__ void sk_tree_step_rare(const SkDevModel* __restrict__ Mg, const SkSmemTables& T, double& rx,
                                               double& ry, double& rz, double kx, double ky, double kz, int m_old,
                                               SkCellPos& q)
{
    sk_locate<2>(Mg, T, rx, ry, rz, q);
    if (q.m == m_old)
    {
        // PathSegmentGenerator::propagateToNextAfter, PathSegmentGenerator.hpp:148-153
        rx = nextafter(rx, (kx < 0.) ? -DBL_MAX : DBL_MAX);
        ry = nextafter(ry, (ky < 0.) ? -DBL_MAX : DBL_MAX);
        rz = nextafter(rz, (kz < 0.) ? -DBL_MAX : DBL_MAX);
        sk_locate<2>(Mg, T, rx, ry, rz, q);
        if (q.m == m_old) q.m = -1;
    }
}

// ---------------------------------------------------------------------------------------------------
// One cell crossing.  Returns the segment (m, dens, ds) and moves (r, p) to the next cell; `p.m < 0` afterwards
// means the path has left the grid.
//   Cartesian: CartesianSpatialGrid::MySegmentGenerator::next, CartesianSpatialGrid.cpp:95-162
//   Octree:    TreeSpatialGrid::MySegmentGenerator::next, TreeSpatialGrid.cpp:140-216, with TreeNode::neighbor
//              (TreeNode.cpp:103-112) served by the per-cell links
// The ray carries the reciprocals of its direction components (0 where the reference treats the component as
// zero, fabs(k) <= 1e-15), so the exit distances cost a multiplication instead of a division per axis.
// ---------------------------------------------------------------------------------------------------
struct SkRayDir {
    double kx, ky, kz;
    double ikx, iky, ikz;
    __device__ __forceinline__ void set(double x, double y, double z)
    {
        kx = x;
        ky = y;
        kz = z;
        ikx = (fabs(x) > 1e-15) ? 1.0 / x : 0.;
        iky = (fabs(y) > 1e-15) ? 1.0 / y : 0.;
        ikz = (fabs(z) > 1e-15) ? 1.0 / z : 0.;
    }
};

template <int GRID>
__device__ __forceinline__ void sk_step(const SkDevModel& M, const SkDevModel* __restrict__ Mg, const SkSmemTables& T,
                                        SkLocalCounters& cnt, double& rx, double& ry, double& rz, const SkRayDir& k,
                                        SkCellPos& p, int& m_out, double& dens_out, double& ds_out)
{
    if (GRID == 1)
    {
        int m = p.m;
        double dens = __ldg(&M.dens[m]);
        double xE = T.X[p.ix + ((k.kx < 0.0) ? 0 : 1)];
        double yE = T.Y[p.iy + ((k.ky < 0.0) ? 0 : 1)];
        double zE = T.Z[p.iz + ((k.kz < 0.0) ? 0 : 1)];
        double dsx = (k.ikx != 0.) ? (xE - rx) * k.ikx : DBL_MAX;
        double dsy = (k.iky != 0.) ? (yE - ry) * k.iky : DBL_MAX;
        double dsz = (k.ikz != 0.) ? (zE - rz) * k.ikz : DBL_MAX;
        double ds;
        bool outside;
        if (dsx <= dsy && dsx <= dsz)
        {
            ds = dsx;
            rx = xE;
            ry += k.ky * dsx;
            rz += k.kz * dsx;
            p.ix += (k.kx < 0.0) ? -1 : 1;
            outside = (p.ix >= M.nx || p.ix < 0);
        }
        else if (dsy < dsx && dsy <= dsz)
        {
            ds = dsy;
            ry = yE;
            rx += k.kx * dsy;
            rz += k.kz * dsy;
            p.iy += (k.ky < 0.0) ? -1 : 1;
            outside = (p.iy >= M.ny || p.iy < 0);
        }
        else
        {
            ds = dsz;
            rz = zE;
            rx += k.kx * dsz;
            ry += k.ky * dsz;
            p.iz += (k.kz < 0.0) ? -1 : 1;
            outside = (p.iz >= M.nz || p.iz < 0);
        }
        m_out = m;
        dens_out = dens;
        ds_out = ds;
        p.m = outside ? -1 : p.iz + M.nz * p.iy + M.nz * M.ny * p.ix;
    }
    else
    {
        // one 32-byte sector: density + the six neighbour links of the current cell
        const int4* rp = reinterpret_cast<const int4*>(&M.cells[p.m]);
        const int4 a = __ldg(rp), b = __ldg(rp + 1);
        const int size = 1 << (M.maxlevel - p.lev);
        const bool nx = k.kx < 0.0, ny = k.ky < 0.0, nz = k.kz < 0.0;
        const double xnext = T.X[p.ix + (nx ? 0 : size)];
        const double ynext = T.Y[p.iy + (ny ? 0 : size)];
        const double znext = T.Z[p.iz + (nz ? 0 : size)];
        const double dsx = (k.ikx != 0.) ? (xnext - rx) * k.ikx : DBL_MAX;
        const double dsy = (k.iky != 0.) ? (ynext - ry) * k.iky : DBL_MAX;
        const double dsz = (k.ikz != 0.) ? (znext - rz) * k.ikz : DBL_MAX;
        // exit wall: x if dsx<=dsy && dsx<=dsz, else y if dsy<=dsx && dsy<=dsz, else z (TreeSpatialGrid.cpp:160-178)
        const bool takex = dsx <= dsy && dsx <= dsz;
        const bool takey = !takex && dsy <= dsx && dsy <= dsz;
        const double ds = takex ? dsx : takey ? dsy : dsz;
        const double other = takex ? fmin(dsy, dsz) : takey ? fmin(dsx, dsz) : fmin(dsx, dsy);
        const double kexit = takex ? k.kx : takey ? k.ky : k.kz;
        const int lx = nx ? a.z : a.w, ly = ny ? b.x : b.y, lz = nz ? b.z : b.w;
        const int link = takex ? lx : takey ? ly : lz;
        const double adv = ds + M.eps;
        rx += k.kx * adv;
        ry += k.ky * adv;
        rz += k.kz * adv;
        m_out = p.m;
        dens_out = __hiloint2double(a.y, a.x);
        ds_out = ds;
        // The link decides the next cell unless the new position may also have crossed a second wall (the exit
        // distances of two walls differ by less than a few eps), the direction grazes the exit wall (the eps advance may
        // be lost to rounding), or the path reaches the domain boundary; those cases take the reference's full search.
        const bool rare = link < 0 || !(other - ds > 4. * M.eps) || !(fabs(kexit) > 1e-3);
        if (!rare)
        {
            // step to the lattice point just across the exit wall, then align to the neighbour's level
            int ix = p.ix, iy = p.iy, iz = p.iz;
            if (takex)
                ix += nx ? -1 : size;
            else if (takey)
                iy += ny ? -1 : size;
            else
                iz += nz ? -1 : size;
            const int nlev = (link >> SK_LINK_LEVEL_SHIFT) & 15;
            const int mask = ~((1 << (M.maxlevel - nlev)) - 1);
            ix &= mask;
            iy &= mask;
            iz &= mask;
            const int idx = link & SK_LINK_INDEX_MASK;
            if (!(link & SK_LINK_INTERNAL))
            {
                p.m = idx;  // leaf neighbour at the same or a coarser level
                p.ix = ix;
                p.iy = iy;
                p.iz = iz;
                p.lev = nlev;
            }
            else
                sk_tree_descend(M.node_child, M.maxlevel, T, idx, ix, iy, iz, nlev, rx, ry, rz, p);
        }
        else
        {
            // (copies confine the address-taken variables, which live in local memory, to this cold branch)
            cnt.fallbacks++;
            double tx = rx, ty = ry, tz = rz;
            SkCellPos q;
            sk_tree_step_rare(Mg, T, tx, ty, tz, k.kx, k.ky, k.kz, p.m, q);
            rx = tx;
            ry = ty;
            rz = tz;
            p = q;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Sources: SourceSystem::launch (SourceSystem.cpp:101-113), NormalizedSource::launch (NormalizedSource.cpp:73-110),
// GeometricSource::launchNormalized (GeometricSource.cpp:66-82), PointSource::launchSpecialty (PointSource.cpp:32-42)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double sk_sample_cdf_loglog(SkRng& g, const double* xv, const double* pv, const double* Pv,
                                                       int n)
{
    double X = sk_uniform(g);  // Random::cdfLogLog, Random.cpp:210-216
    int i = sk_locate_clip(Pv, n, X);
    double alpha = log(pv[i + 1] / pv[i]) / log(xv[i + 1] / xv[i]);
    return xv[i] * sk_gexp(-alpha, (X - Pv[i]) / (pv[i] * xv[i]));
}
__device__ __forceinline__ double sk_sample_cdf_linlin(SkRng& g, const double* xv, const double* Pv, int n)
{
    double X = sk_uniform(g);  // Random::cdfLinLin, Random.cpp:201-206
    int i = sk_locate_clip(Pv, n, X);
    return sk_interp_linlin(X, Pv[i], Pv[i + 1], xv[i], xv[i + 1]);
}
__device__ __forceinline__ double sk_specific_luminosity(const SkDevSource& s, double lambda)
{
    if (s.sed_kind == SK_SED_BLACKBODY) return sk_planck(lambda, s.sed_temperature) / s.sed_norm;
    int i = sk_locate_fail(s.sed_lambda, s.sed_n, lambda);
    if (i < 0) return 0.;
    return sk_interp_loglog(lambda, s.sed_lambda[i], s.sed_lambda[i + 1], s.sed_p[i], s.sed_p[i + 1]);
}
// ExpDiskGeometry::randomCylRadius / randomZ, ExpDiskGeometry.cpp:46-68
__device__ __forceinline__ double sk_expdisk_R(SkRng& g, double hR, double Rmin, double Rmax)
{
    double R, X;
    do
    {
        X = sk_uniform(g);
        R = hR * (-1.0 - sk_lambert_w1((X - 1.0) / M_E));
    } while ((Rmax > 0.0 && R >= Rmax) || R <= Rmin);
    return R;
}
__device__ __forceinline__ double sk_expdisk_z(SkRng& g, double hz, double zmax)
{
    double z, X;
    do
    {
        X = sk_uniform(g);
        z = (X <= 0.5) ? hz * log(2.0 * X) : -hz * log(2.0 * (1.0 - X));
    } while (zmax > 0.0 && fabs(z) >= zmax);
    return z;
}
__device__ __noinline__ void sk_generate_position(SkRng& g, const SkDevSource& s, double& x, double& y, double& z)
{
    const double* p = s.gp;
    switch (s.geometry)
    {
        case SK_GEOM_SHELL:
        {
            // ShellGeometry::randomRadius (ShellGeometry.cpp:43-57) + SpheGeometry::generatePosition (SpheGeometry.cpp:26-33)
            double pe = p[2], smin = p[3], sdiff = p[4], tmin = p[5], tmax = p[6];
            double X = sk_uniform(g);
            double rad;
            if (fabs(pe - 3.0) < 1e-2)
                rad = sk_gexp(pe - 2.0, smin + X * sdiff);
            else
            {
                double zz = (1.0 - X) * tmin + X * tmax;
                rad = pow(zz, 1.0 / (3.0 - pe));
            }
            double kx, ky, kz;
            sk_random_direction(g, kx, ky, kz);
            x = rad * kx;
            y = rad * ky;
            z = rad * kz;
            break;
        }
        case SK_GEOM_EXPDISK:
        {
            double R = sk_expdisk_R(g, p[0], p[2], p[3]);  // SepAxGeometry::generatePosition, SepAxGeometry.cpp:12-20
            double phi = 2.0 * M_PI * sk_uniform(g);
            double zz = sk_expdisk_z(g, p[1], p[4]);
            x = R * cos(phi);
            y = R * sin(phi);
            z = zz;
            break;
        }
        case SK_GEOM_RING:
        {
            double R = sk_sample_cdf_linlin(g, s.geom_table_x, s.geom_table_P, s.geom_table_n);  // RingGeometry.cpp:56-68
            double phi = 2.0 * M_PI * sk_uniform(g);
            double X = sk_uniform(g);
            double zz = (X <= 0.5) ? p[2] * log(2.0 * X) : -p[2] * log(2.0 * (1.0 - X));
            x = R * cos(phi);
            y = R * sin(phi);
            z = zz;
            break;
        }
        case SK_GEOM_SPIRAL_EXPDISK:
        {
            // SpiralStructureGeometryDecorator::generatePosition / perturbation, .cpp:33-45,72-76
            double R0 = sk_expdisk_R(g, p[0], p[2], p[3]);
            double phi0 = 2.0 * M_PI * sk_uniform(g);
            double zz = sk_expdisk_z(g, p[1], p[4]);
            double x0 = R0 * cos(phi0), y0 = R0 * sin(phi0);
            double R = sqrt(x0 * x0 + y0 * y0);
            double m = p[5], pitch = p[6], Rz = p[7], phiz = p[8], w = p[9], N = p[10];
            double tanp = tan(pitch);
            double cn = sqrt(M_PI) * tgamma(N + 1.0) / tgamma(N + 0.5);
            double c = 1.0 + (cn - 1.0) * w;
            double phi, t;
            do
            {
                phi = 2.0 * M_PI * sk_uniform(g);
                double gamma = log(R / Rz) / tanp + phiz + 0.5 * M_PI / m;
                double perturbation = (1.0 - w) + w * cn * pow(sin(0.5 * m * (gamma - phi)), 2 * N);
                t = sk_uniform(g) * c / perturbation;
            } while (t > 1);
            x = R * cos(phi);
            y = R * sin(phi);
            z = zz;
            break;
        }
        default: x = y = z = 0.; break;
    }
}

struct SkLaunch {
    double lambda, W, rx, ry, rz, kx, ky, kz;
    int ilam;
};

__device__ __noinline__ void sk_launch_primary(const SkDevModel* __restrict__ Mg, SkRng& g, unsigned long long history,
                                               SkLaunch& pp)
{
    const SkDevModel& M = *Mg;
    int lo = 0, hi = M.nsrc + 1;  // std::upper_bound(_Iv, historyIndex) - 1
    while (lo < hi)
    {
        int mid = (lo + hi) >> 1;
        if (history < M.Iv[mid])
            hi = mid;
        else
            lo = mid + 1;
    }
    const SkDevSource& s = M.src[lo - 1];
    double L = M.Lpp * s.Lw;
    double lambda, w;
    double xi = s.wavelength_bias;
    if (!xi)
    {
        lambda = sk_sample_cdf_loglog(g, s.sed_lambda, s.sed_p, s.sed_P, s.sed_n);
        w = 1.;
    }
    else
    {
        if (sk_uniform(g) > xi)
            lambda = sk_sample_cdf_loglog(g, s.sed_lambda, s.sed_p, s.sed_P, s.sed_n);
        else if (s.bias_kind == SK_BIAS_OLIGO)
        {
            size_t index = (size_t)(sk_uniform(g) * s.oligo_n);  // OligoWavelengthDistribution.cpp:34-38
            lambda = s.oligo_lambda[index];
        }
        else
        {
            double logMin = log(s.bias_min);
            double logWidth = log(s.bias_max) - log(s.bias_min);
            lambda = exp(logMin + logWidth * sk_uniform(g));  // DefaultWavelengthDistribution.cpp:37-40
        }
        double sl = sk_specific_luminosity(s, lambda);
        if (!sl)
            w = 0.;
        else
        {
            double b;
            if (s.bias_kind == SK_BIAS_OLIGO)
                b = s.oligo_probability;
            else
            {
                double logWidth = log(s.bias_max) - log(s.bias_min);
                if (lambda >= s.bias_min * (1 - 1e-14) && lambda <= s.bias_max * (1 + 1e-14))  // Range.hpp:56
                    b = 1. / (logWidth * lambda);
                else
                    b = 0.;
            }
            w = sl / ((1 - xi) * sl + xi * b);
        }
    }
    double Lw = L * w;
    if (s.kind == SK_SRC_POINT)
    {
        pp.rx = s.position[0];
        pp.ry = s.position[1];
        pp.rz = s.position[2];
    }
    else
        sk_generate_position(g, s, pp.rx, pp.ry, pp.rz);
    sk_random_direction(g, pp.kx, pp.ky, pp.kz);
    pp.lambda = lambda;
    pp.W = Lw * lambda;  // PhotonPacket::launch, PhotonPacket.cpp:18-40
    pp.ilam = sk_locate_clip(M.lam_border, M.nlam, lambda);
}

// ---------------------------------------------------------------------------------------------------
// Instruments
// ---------------------------------------------------------------------------------------------------
// First half of Instrument::detect / FluxRecorder::detect: does this instrument record a packet at (x,y,z) with
// wavelength lambda?  SEDInstrument.cpp:22-25 + ApertureInstrument.cpp:24-43, FrameInstrument.cpp:45-64,
// FluxRecorder.cpp:306-313.  Returns false when nothing is recorded (and no optical depth is needed).
__device__ __forceinline__ bool sk_detect_geometry(const SkDevModel& M, const SkDevInstr& q, double x, double y,
                                                   double z, double lambda, int& l, int& ell)
{
    l = 0;
    if (q.kind == SK_INSTR_SED)
    {
        if (q.radius2)
        {
            double xpp = -q.sinphi * x + q.cosphi * y;
            double ypp = -q.cosphi * q.costheta * x - q.sinphi * q.costheta * y + q.sintheta * z;
            double radius2 = xpp * xpp + ypp * ypp;
            if (radius2 > q.radius2) return false;
        }
    }
    else
    {
        double xpp = -q.sinphi * x + q.cosphi * y;
        double ypp = -q.cosphi * q.costheta * x - q.sinphi * q.costheta * y + q.sintheta * z;
        double xp = q.cosomega * xpp - q.sinomega * ypp;
        double yp = q.sinomega * xpp + q.cosomega * ypp;
        int i = (int)floor((xp - q.xpmin) / q.xpsiz);
        int jj = (int)floor((yp - q.ypmin) / q.ypsiz);
        if (i < 0 || i >= q.nx || jj < 0 || jj >= q.ny)
            l = -1;
        else
            l = i + q.nx * jj;
    }
    if (!q.include_sed && l < 0) return false;
    ell = sk_wlg_bin(M.wlg[q.wlg], lambda);
    return ell >= 0;
}

// Second half of FluxRecorder::detect (FluxRecorder.cpp:320-433): component routing and the tallies.
__device__ __forceinline__ void sk_record(const SkDevInstr& q, int l, int ell, double L, double Lext, int nscatt,
                                          bool primary_origin)
{
    int c_ext, c_tr = -1, c_lev = -1;
    if (q.record_total_only)
        c_ext = SK_COMP_TOTAL;
    else if (primary_origin)
    {
        if (nscatt == 0)
        {
            c_tr = SK_COMP_TRANSPARENT;
            c_ext = SK_COMP_PRIMARY_DIRECT;
        }
        else
        {
            c_ext = SK_COMP_PRIMARY_SCATTERED;
            if (nscatt <= q.num_levels) c_lev = SK_COMP_PRIMARY_SCATTERED_LEVEL + nscatt - 1;
        }
    }
    else
    {
        if (nscatt == 0)
        {
            c_tr = SK_COMP_SECONDARY_TRANSPARENT;
            c_ext = SK_COMP_SECONDARY_DIRECT;
        }
        else
            c_ext = SK_COMP_SECONDARY_SCATTERED;
    }
    if (q.include_sed)
    {
        atomicAdd(&q.sed[c_ext][ell], Lext);  // LockFree::add, LockFree.hpp:23-37 -> native fp64 RED
        if (c_tr >= 0) atomicAdd(&q.sed[c_tr][ell], L);
        if (c_lev >= 0) atomicAdd(&q.sed[c_lev][ell], Lext);
    }
    if (q.include_ifu && l >= 0)
    {
        size_t index = (size_t)l + (size_t)ell * q.npix;  // FluxRecorder.cpp:433
        atomicAdd(&q.ifu[c_ext][index], Lext);
        if (c_tr >= 0) atomicAdd(&q.ifu[c_tr][index], L);
        if (c_lev >= 0) atomicAdd(&q.ifu[c_lev][index], Lext);
    }
}

// ---------------------------------------------------------------------------------------------------
// Trace stage: walks all rays of `list[0..n)` with dynamic lane refill.
//   MODE 0  forward path to the boundary: MediumSystem::setExtinctionOpticalDepths (MediumSystem.cpp:849-871); with
//           STORE fused with MonteCarloSimulation::storeRadiationField (.cpp:638-665)
//   MODE 1  walk to the interaction point: SpatialGridPath::findInteractionPoint (SpatialGridPath.cpp:164-206), or for
//           non-forced scattering MediumSystem::setInteractionPointUsingExtinction (MediumSystem.cpp:978-1010)
//   MODE 2  optical depth to the observer: MediumSystem::getExtinctionOpticalDepth (MediumSystem.cpp:1192-1219)
// Structure: a compact inner loop that only crosses cells, and an outer service block (store the results of finished
// rays, load new rays) that is entered when at least SK_REFILL_MIN lanes are idle, so that its cost is shared.
// ---------------------------------------------------------------------------------------------------
#ifndef SK_REFILL_MIN
#define SK_REFILL_MIN 6
#endif

template <int GRID, int MODE, bool STORE>
__device__ __forceinline__ void sk_trace_stage(const SkDevModel& M, const SkDevModel* __restrict__ Mg,
                                               const SkSmemTables& T, const SkRunArgs& A, const SkPoolView& P,
                                               const int* list, int n, const SkRayDir& obs, SkLocalCounters& cnt)
{
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const bool forced = M.force_scattering != 0;
    int head = 0;
    bool active = false, pending = false;
    int slot = 0;
    double rx = 0, ry = 0, rz = 0;
    SkRayDir k;
    k.set(0., 0., 1.);
    SkCellPos p{-1, 0, 0, 0, 0};
    double tau = 0, s = 0, limit = 0, section = 0;
    int nseg = 0;
    // MODE 0 + STORE extras
    double lum = 0, lnExtBeg = 0, extBeg = 1;
    int rf_ell = -1;
    double* rf = nullptr;
    // MODE 1 results
    SkCellPos hit{-1, 0, 0, 0, 0};
    double s_int = 0;
    bool found = false;

    while (true)
    {
        // ---------------- service block: results of finished rays out, new rays in
        if (pending)
        {
            pending = false;
            if (MODE == 0)
            {
                P.D(D_TAUPATH, slot) = tau;
                P.D(D_STOT, slot) = s;
                P.I(I_NSEG, slot) = nseg;
                cnt.fwd_paths++;
                cnt.fwd_segs += nseg;
            }
            else if (MODE == 1)
            {
                P.D(D_SINT, slot) = s_int;
                P.I(I_MINT, slot) = hit.m;
                P.I(I_MIX, slot) = hit.ix;
                P.I(I_MIY, slot) = hit.iy;
                P.I(I_MIZ, slot) = hit.iz;
                P.I(I_MLEV, slot) = hit.lev;
                if (found) P.I(I_STATE, slot) |= SK_ST_FOUND;
                if (forced)
                    cnt.replay_segs += nseg;
                else
                {
                    cnt.fwd_paths++;
                    cnt.fwd_segs += nseg;
                }
            }
            else
            {
                P.D(D_PTAU, slot) = tau;
                cnt.peel_paths++;
                cnt.peel_segs += nseg;
            }
        }
        {
            unsigned idle = __ballot_sync(0xffffffffu, !active);
            int idx = head + __popc(idle & lt_mask);
            head += __popc(idle);
            if (!active && idx < n)
            {
                slot = list[idx];
                active = true;
                rx = P.D(D_RX, slot);
                ry = P.D(D_RY, slot);
                rz = P.D(D_RZ, slot);
                if (MODE == 2)
                    k = obs;
                else
                    k.set(P.D(D_KX, slot), P.D(D_KY, slot), P.D(D_KZ, slot));
                p.m = P.I(I_M, slot);
                p.ix = P.I(I_IX, slot);
                p.iy = P.I(I_IY, slot);
                p.iz = P.I(I_IZ, slot);
                p.lev = P.I(I_LEV, slot);
                section = P.D(D_SIGEXT, slot);
                tau = 0.;
                s = 0.;
                nseg = 0;
                if (MODE == 0 && STORE)
                {
                    lnExtBeg = 0.;
                    extBeg = 1.;
                    double lambda = P.D(D_LAMBDA, slot);
                    rf_ell = sk_wlg_bin(M.wlg[M.rf_grid], lambda);  // MonteCarloSimulation.cpp:643
                    rf = A.primary ? M.rf1 : M.rf2c;
                    lum = P.D(D_W, slot) / lambda;
                }
                if (MODE == 1)
                {
                    limit = P.D(D_TAUINT, slot);
                    hit = p;
                    found = false;
                    s_int = 0.;
                }
                if (MODE == 2) limit = P.D(D_LIMIT, slot);
                if (p.m < 0)
                {
                    // the path starts outside (or exactly on the border of) the grid: PathSegmentGenerator::moveInside
                    double cumds = 0.;
                    double tx = rx, ty = ry, tz = rz;
                    if (sk_move_inside(tx, ty, tz, k.kx, k.ky, k.kz, Mg->ext, M.eps, cumds))
                    {
                        SkCellPos q;
                        sk_locate<GRID>(Mg, T, tx, ty, tz, q);
                        rx = tx;
                        ry = ty;
                        rz = tz;
                        p = q;
                        if (cumds > 0.)
                        {
                            // the empty segment in front of the grid (m = -1)
                            nseg++;
                            s += cumds;
                        }
                    }
                    if (p.m < 0)
                    {
                        // the path misses the grid: no segments
                        active = false;
                        pending = true;
                        if (MODE == 1) s_int = s;
                    }
                }
            }
            if (!__any_sync(0xffffffffu, active || pending)) break;
        }
        // ---------------- inner loop: cross cells until enough lanes have finished their ray
        const int want_idle = head < n ? SK_REFILL_MIN : 32;
        int nidle;
        do
        {
            if (active)
            {
                int m;
                double dens, ds;
                const SkCellPos cur = p;
                sk_step<GRID>(M, Mg, T, cnt, rx, ry, rz, k, p, m, dens, ds);
                bool done = false;
                if (MODE == 0)
                {
                    if (ds > 0.)  // SpatialGridPath::addSegment, SpatialGridPath.cpp:41-48
                    {
                        nseg++;
                        s += ds;
                        tau += section * dens * ds;
                        if (STORE && rf_ell >= 0)
                        {
                            double lnExtEnd = -tau;
                            double extEnd = exp(lnExtEnd);
                            double extMean = sk_lnmean4(extEnd, extBeg, lnExtEnd, lnExtBeg);
                            double Lds = lum * extMean * ds;
                            atomicAdd(&rf[(size_t)m * M.nrf + rf_ell], Lds);  // MediumSystem.cpp:1294-1300
                            cnt.rf++;
                            lnExtBeg = lnExtEnd;
                            extBeg = extEnd;
                        }
                    }
                }
                else if (MODE == 1)
                {
                    if (forced ? (ds > 0.) : true)
                    {
                        nseg++;
                        double tau0 = tau, s0 = s;
                        s += ds;
                        tau += section * dens * ds;
                        hit = cur;
                        if (limit < tau)
                        {
                            s_int = sk_interp_linlin(limit, tau0, tau, s0, s);  // interaction inside this segment
                            found = true;
                            done = true;
                        }
                    }
                }
                else
                {
                    nseg++;
                    tau += section * dens * ds;
                    if (tau >= limit)
                    {
                        tau = INFINITY;  // MediumSystem.cpp:1215
                        done = true;
                    }
                }
                if (!done && p.m < 0)
                {
                    // the path has left the grid; MODE 1: at or beyond the exit optical depth of the last segment ->
                    // use the last segment (SpatialGridPath.cpp:199-205); non-forced: no interaction
                    done = true;
                    if (MODE == 1) s_int = s;
                }
                if (done)
                {
                    active = false;
                    pending = true;
                }
            }
            nidle = __popc(__ballot_sync(0xffffffffu, !active));
        } while (nidle < want_idle);
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------
// Event stages
// ---------------------------------------------------------------------------------------------------
// Ends a history: FluxRecorder::recordContributions for the SED arrays (FluxRecorder.cpp:962-986); frees the slot.
__device__ __forceinline__ void sk_finish_history(const SkDevModel& M, const SkPoolView& P, int slot)
{
    for (int j = 0; j < M.ninstr; ++j)
    {
        const SkDevInstr& q = M.instr[j];
        if (!q.record_stats) continue;
        int ell = P.I(I_HELL0 + j, slot);
        if (ell >= 0)
        {
            double w = P.D(D_HISTW0 + j, slot);
            double wn = 1.;
            for (int kk = 0; kk <= 4; ++kk)
            {
                atomicAdd(&q.wsed[kk][ell], wn);
                wn *= w;
            }
        }
    }
    P.I(I_STATE, slot) = 0;
}

// Peel-off towards the observer group [j0, j1): MonteCarloSimulation::peelOffEmission (.cpp:617-634) /
// peelOffScattering (.cpp:784-842, consolidated branch) split around the trace stage.
template <int GRID>
__device__ __forceinline__ void sk_peel_group(const SkDevModel& M, const SkDevModel* __restrict__ Mg,
                                              const SkSmemTables& T, const SkRunArgs& A,
                                              const SkPoolView& P, int* list, int j0, int j1, SkLocalCounters& cnt)
{
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const SkDevInstr& q0 = M.instr[j0];
    const double ox = q0.kobs[0], oy = q0.kobs[1], oz = q0.kobs[2];
    int n = 0;
    // ---- set-up: weight of the peel-off packet and which packets need an optical depth at all
#pragma unroll 1
    for (int base = 0; base < SK_POOL; base += 32)
    {
        int slot = base + lane;
        int st = P.I(I_STATE, slot);
        bool need = false;
        if (st & SK_ST_LIVE)
        {
            double W = P.D(D_W, slot);
            double lambda = P.D(D_LAMBDA, slot);
            double peelW;
            if (st & SK_ST_SCATTER)
            {
                // DustMix::peeloffScattering HG branch (DustMix.cpp:430-445); MediumSystem::peelOffScattering
                // (MediumSystem.cpp:734-767) with the single-medium weight 1; launchScatteringPeelOff (PhotonPacket.cpp:89-103)
                double costheta = P.D(D_KX, slot) * ox + P.D(D_KY, slot) * oy + P.D(D_KZ, slot) * oz;
                double gp = M.gpar[P.I(I_ILAM, slot)];
                double value = fabs(gp) > 0.95 ? sk_mean_hg(gp, costheta) : sk_value_hg(gp, costheta);
                double I = 0.;
                I += value * 1.;
                peelW = W * I;
            }
            else
                peelW = W;  // launchEmissionPeelOff, PhotonPacket.cpp:66-85 (isotropic emission)
            P.D(D_PEELW, slot) = peelW;
            double x = P.D(D_RX, slot), y = P.D(D_RY, slot), z = P.D(D_RZ, slot);
            for (int j = j0; j < j1; ++j)
            {
                int l, ell;
                if (sk_detect_geometry(M, M.instr[j], x, y, z, lambda, l, ell)) need = true;
            }
            if (need)
            {
                double L = peelW / lambda;
                if (L <= 0)
                {
                    P.D(D_PTAU, slot) = INFINITY;  // MediumSystem.cpp:1196
                    need = false;
                }
                else
                    P.D(D_LIMIT, slot) = log(L) + 745;  // MediumSystem.cpp:1199
            }
        }
        unsigned mask = __ballot_sync(0xffffffffu, need);
        if (need) list[n + __popc(mask & lt_mask)] = slot;
        n += __popc(mask);
    }
    __syncwarp();
    SkRayDir obs;
    obs.set(ox, oy, oz);
    sk_trace_stage<GRID, 2, false>(M, Mg, T, A, P, list, n, obs, cnt);
    // ---- detection: FluxRecorder::detect, FluxRecorder.cpp:304-468
#pragma unroll 1
    for (int base = 0; base < SK_POOL; base += 32)
    {
        int slot = base + lane;
        int st = P.I(I_STATE, slot);
        if (st & SK_ST_LIVE)
        {
            double lambda = P.D(D_LAMBDA, slot);
            double x = P.D(D_RX, slot), y = P.D(D_RY, slot), z = P.D(D_RZ, slot);
            double L = P.D(D_PEELW, slot) / lambda;
            int nscatt = (st & SK_ST_SCATTER) ? P.I(I_NSCATT, slot) + 1 : 0;
            for (int j = j0; j < j1; ++j)
            {
                const SkDevInstr& q = M.instr[j];
                int l, ell;
                if (!sk_detect_geometry(M, q, x, y, z, lambda, l, ell)) continue;
                double Lext = L * exp(-P.D(D_PTAU, slot));
                cnt.det++;
                sk_record(q, l, ell, L, Lext, nscatt, A.primary != 0);
                if (q.record_stats && q.include_sed)
                {
                    P.D(D_HISTW0 + j, slot) += Lext;  // FluxRecorder.cpp:457-466
                    P.I(I_HELL0 + j, slot) = ell;
                }
            }
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------
// The life cycles of one warp's pool.
// ---------------------------------------------------------------------------------------------------
template <int GRID>
__device__ __forceinline__ void sk_warp_life_cycles(const SkDevModel& M, const SkDevModel* __restrict__ Mg,
                                                    const SkSmemTables& T, const SkRunArgs& A,
                                                    const SkPoolView& P, int* list, SkLocalCounters& cnt)
{
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const bool forced = M.force_scattering != 0;
#pragma unroll 1
    for (int base = 0; base < SK_POOL; base += 32) P.I(I_STATE, base + lane) = 0;
    __syncwarp();
    bool more = true;

    while (true)
    {
        // ---- stage A: launch a history into every free slot (SourceSystem::launch)
        int nlive = 0;
    #pragma unroll 1
    for (int base = 0; base < SK_POOL; base += 32)
        {
            int slot = base + lane;
            int st = P.I(I_STATE, slot);
            bool want = more && !(st & SK_ST_LIVE);
            unsigned mask = __ballot_sync(0xffffffffu, want);
            if (mask)
            {
                int cntw = __popc(mask);
                unsigned long long b = 0;
                if (lane == 0) b = atomicAdd(A.work_counter, (unsigned long long)cntw);
                b = __shfl_sync(0xffffffffu, b, 0);
                unsigned long long h = b + __popc(mask & lt_mask);
                if (b + cntw >= A.count) more = false;
                if (want && h < A.count)
                {
                    unsigned long long history = A.first + h;
                    SkRng g;
                    sk_rng_init(g, M.seed, A.stream_id, history, 0);
                    SkLaunch pp;
                    sk_launch_primary(Mg, g, history, pp);
                    if (pp.W / pp.lambda > 0)  // MonteCarloSimulation.cpp:553
                    {
                        cnt.packets++;
                        SkCellPos c;
                        c.m = -1;
                        c.ix = c.iy = c.iz = c.lev = 0;
                        if (sk_box_strictly_inside(M.ext, pp.rx, pp.ry, pp.rz)) sk_locate<GRID>(Mg, T, pp.rx, pp.ry, pp.rz, c);
                        P.D(D_RX, slot) = pp.rx;
                        P.D(D_RY, slot) = pp.ry;
                        P.D(D_RZ, slot) = pp.rz;
                        P.D(D_KX, slot) = pp.kx;
                        P.D(D_KY, slot) = pp.ky;
                        P.D(D_KZ, slot) = pp.kz;
                        P.D(D_LAMBDA, slot) = pp.lambda;
                        P.D(D_W, slot) = pp.W;
                        P.D(D_LTHR, slot) = (pp.W / pp.lambda) / M.min_weight_reduction;  // .cpp:563
                        P.D(D_SIGEXT, slot) = M.sig_ext[pp.ilam];
                        P.I(I_HLO, slot) = (int)(uint32_t)history;
                        P.I(I_HHI, slot) = (int)(uint32_t)(history >> 32);
                        P.I(I_DRAW, slot) = (int)g.draw;
                        P.I(I_NSCATT, slot) = 0;
                        P.I(I_ILAM, slot) = pp.ilam;
                        P.I(I_M, slot) = c.m;
                        P.I(I_IX, slot) = c.ix;
                        P.I(I_IY, slot) = c.iy;
                        P.I(I_IZ, slot) = c.iz;
                        P.I(I_LEV, slot) = c.lev;
                        for (int j = 0; j < M.ninstr; ++j)
                        {
                            P.D(D_HISTW0 + j, slot) = 0.;
                            P.I(I_HELL0 + j, slot) = -1;
                        }
                        st = SK_ST_LIVE;
                        P.I(I_STATE, slot) = st;
                    }
                }
            }
            nlive += __popc(__ballot_sync(0xffffffffu, (st & SK_ST_LIVE) != 0));
        }
        __syncwarp();
        if (nlive == 0)
        {
            if (!more) break;
            continue;
        }

        // ---- stage B: peel-off (emission for fresh packets, scattering for the others), one trace per observer
        if (A.peel)
        {
            int j0 = 0;
            while (j0 < M.ninstr)
            {
                int j1 = j0 + 1;
                while (j1 < M.ninstr && M.instr[j1].same_as_preceding) j1++;
                sk_peel_group<GRID>(M, Mg, T, A, P, list, j0, j1, cnt);
                j0 = j1;
            }
        }

        // ---- stage C: the pending scattering events: MediumSystem::simulateScattering (.cpp:796-823) +
        //      DustMix::performScattering HG branch (DustMix.cpp:496-511); then the list of all live packets
        int n = 0;
    #pragma unroll 1
    for (int base = 0; base < SK_POOL; base += 32)
        {
            int slot = base + lane;
            int st = P.I(I_STATE, slot);
            bool live = (st & SK_ST_LIVE) != 0;
            if (live && (st & SK_ST_SCATTER))
            {
                SkRng g;
                sk_rng_init(g, M.seed, A.stream_id,
                            ((unsigned long long)(uint32_t)P.I(I_HHI, slot) << 32) | (uint32_t)P.I(I_HLO, slot),
                            (uint32_t)P.I(I_DRAW, slot));
                double gp = M.gpar[P.I(I_ILAM, slot)];
                double kx = P.D(D_KX, slot), ky = P.D(D_KY, slot), kz = P.D(D_KZ, slot);
                if (fabs(gp) < 1e-6)
                    sk_random_direction(g, kx, ky, kz);
                else
                {
                    double f = ((1.0 - gp) * (1.0 + gp)) / (1.0 - gp + 2.0 * gp * sk_uniform(g));
                    double costheta = (1.0 + gp * gp - f * f) / (2.0 * gp);
                    sk_random_direction_about(g, kx, ky, kz, costheta);
                }
                P.D(D_KX, slot) = kx;
                P.D(D_KY, slot) = ky;
                P.D(D_KZ, slot) = kz;
                P.I(I_DRAW, slot) = (int)g.draw;
                P.I(I_NSCATT, slot) += 1;
                P.I(I_STATE, slot) = st & ~SK_ST_SCATTER;
                cnt.scatt++;
            }
            unsigned mask = __ballot_sync(0xffffffffu, live);
            if (live) list[n + __popc(mask & lt_mask)] = slot;
            n += __popc(mask);
        }
        __syncwarp();

        // ---- stage D: forward paths (forced scattering only)
        SkRayDir nodir;
        nodir.set(0., 0., 1.);
        if (forced)
        {
            if (A.store)
                sk_trace_stage<GRID, 0, true>(M, Mg, T, A, P, list, n, nodir, cnt);
            else
                sk_trace_stage<GRID, 0, false>(M, Mg, T, A, P, list, n, nodir, cnt);
        }

        // ---- stage E: sample the interaction optical depth: simulateForcedPropagation (.cpp:696-722) or
        //      Random::expon for simulateNonForcedPropagation (.cpp:749)
        n = 0;
    #pragma unroll 1
    for (int base = 0; base < SK_POOL; base += 32)
        {
            int slot = base + lane;
            int st = P.I(I_STATE, slot);
            bool live = (st & SK_ST_LIVE) != 0;
            if (live)
            {
                SkRng g;
                sk_rng_init(g, M.seed, A.stream_id,
                            ((unsigned long long)(uint32_t)P.I(I_HHI, slot) << 32) | (uint32_t)P.I(I_HLO, slot),
                            (uint32_t)P.I(I_DRAW, slot));
                if (forced)
                {
                    double taupath = P.D(D_TAUPATH, slot);
                    if (!(P.I(I_NSEG, slot) > 0 && taupath > 0.))
                    {
                        // no extinction along the path: the packet cannot scatter, terminate it (.cpp:702-706)
                        sk_finish_history(M, P, slot);
                        live = false;
                    }
                    else
                    {
                        double xi = M.path_length_bias;
                        double tauint;
                        if (xi == 0.)
                            tauint = sk_expon_cutoff(g, taupath);
                        else
                        {
                            tauint = sk_uniform(g) < xi ? sk_uniform(g) * taupath : sk_expon_cutoff(g, taupath);
                            double pw = -exp(-tauint) / expm1(-taupath);
                            double qw = (1.0 - xi) * pw + xi / taupath;
                            P.D(D_W, slot) *= pw / qw;
                        }
                        P.D(D_TAUINT, slot) = tauint;
                    }
                }
                else
                    P.D(D_TAUINT, slot) = -log(sk_uniform(g));
                if (live)
                {
                    P.I(I_DRAW, slot) = (int)g.draw;
                    P.I(I_STATE, slot) = st & ~SK_ST_FOUND;
                }
            }
            unsigned mask = __ballot_sync(0xffffffffu, live);
            if (live) list[n + __popc(mask & lt_mask)] = slot;
            n += __popc(mask);
        }
        __syncwarp();

        // ---- stage F: walk to the interaction point
        sk_trace_stage<GRID, 1, false>(M, Mg, T, A, P, list, n, nodir, cnt);

        // ---- stage G: the interaction: albedo weight, move, termination test (.cpp:724-741, 576-580)
    #pragma unroll 1
    for (int base = 0; base < SK_POOL; base += 32)
        {
            int slot = base + lane;
            int st = P.I(I_STATE, slot);
            if (st & SK_ST_LIVE)
            {
                bool alive = true;
                if (!forced && !(st & SK_ST_FOUND)) alive = false;  // escaped, MonteCarloSimulation.cpp:594
                if (alive)
                {
                    int m = P.I(I_MINT, slot);
                    int ilam = P.I(I_ILAM, slot);
                    // MediumSystem::albedoForScattering, MediumSystem.cpp:678-693
                    double albedo = 0.;
                    if (m >= 0)
                    {
                        double dn = GRID == 1 ? M.dens[m] : M.cells[m].dens;
                        double ksca = dn * M.sig_sca[ilam];
                        double kext = dn * P.D(D_SIGEXT, slot);
                        albedo = kext > 0. ? ksca / kext : 0.;
                    }
                    double W = P.D(D_W, slot);
                    if (forced)
                        W *= -expm1(-P.D(D_TAUPATH, slot)) * albedo;
                    else
                        W *= albedo;
                    double sint = P.D(D_SINT, slot);
                    double kx = P.D(D_KX, slot), ky = P.D(D_KY, slot), kz = P.D(D_KZ, slot);
                    double x = P.D(D_RX, slot) + sint * kx;  // PhotonPacket::propagate, PhotonPacket.cpp:107-111
                    double y = P.D(D_RY, slot) + sint * ky;
                    double z = P.D(D_RZ, slot) + sint * kz;
                    P.D(D_W, slot) = W;
                    P.D(D_RX, slot) = x;
                    P.D(D_RY, slot) = y;
                    P.D(D_RZ, slot) = z;
                    double L = W / P.D(D_LAMBDA, slot);
                    if (forced)
                    {
                        if (L <= 0 || (L <= P.D(D_LTHR, slot) && P.I(I_NSCATT, slot) >= M.min_scatt_events)) alive = false;
                    }
                    else if (L <= 0)
                        alive = false;
                    if (alive)
                    {
                        // the next paths start in the interaction cell unless rounding moved the point out of it
                        SkCellPos c{m, P.I(I_MIX, slot), P.I(I_MIY, slot), P.I(I_MIZ, slot), P.I(I_MLEV, slot)};
                        bool inside = m >= 0 && sk_box_strictly_inside(M.ext, x, y, z);
                        if (inside)
                        {
                            if (GRID == 1)
                                inside = x >= T.X[c.ix] && x < T.X[c.ix + 1] && y >= T.Y[c.iy] && y < T.Y[c.iy + 1]
                                         && z >= T.Z[c.iz] && z < T.Z[c.iz + 1];
                            else
                            {
                                int size = 1 << (M.maxlevel - c.lev);
                                inside = x >= T.X[c.ix] && x < T.X[c.ix + size] && y >= T.Y[c.iy] && y < T.Y[c.iy + size]
                                         && z >= T.Z[c.iz] && z < T.Z[c.iz + size];
                            }
                            if (!inside)
                            {
                                inside = sk_box_strictly_inside(M.ext, x, y, z);
                                if (inside) sk_locate<GRID>(Mg, T, x, y, z, c);
                            }
                        }
                        if (!inside) c.m = -1;
                        P.I(I_M, slot) = c.m;
                        P.I(I_IX, slot) = c.ix;
                        P.I(I_IY, slot) = c.iy;
                        P.I(I_IZ, slot) = c.iz;
                        P.I(I_LEV, slot) = c.lev;
                        P.I(I_STATE, slot) = SK_ST_LIVE | SK_ST_SCATTER;
                    }
                }
                if (!alive) sk_finish_history(M, P, slot);
            }
        }
        __syncwarp();
    }
}
