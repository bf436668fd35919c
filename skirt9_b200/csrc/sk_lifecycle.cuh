// sk_lifecycle.cuh -- the photon life cycle on the device: MonteCarloSimulation::performLifeCycle
// (SKIRT/core/MonteCarloSimulation.cpp:538-613) and everything it calls, restated for one CUDA thread per
// in-flight history.  Paths are never materialised (the reference stores vector<Segment>, SpatialGridPath.hpp:93-115):
// the forward path is walked once for the total optical depth (with the radiation-field deposits fused in) and
// re-walked up to the sampled interaction point; both walks use identical arithmetic so they see identical segments.
#pragma once
#include "sk_device.cuh"

struct SkPacket {
    double lambda, W;  // PhotonPacket::_lambda, _W = L*lambda (PhotonPacket.hpp:337-340)
    double rx, ry, rz, kx, ky, kz;
    int nscatt;
    int primary_origin;
    int ilam;          // DustMix::indexForLambda(lambda) (DustMix.cpp:276-279); constant without kinematics
    double sig_ext;    // sectionExt at ilam
};

struct SkLocalCounters {
    unsigned int packets, fwd_paths, fwd_segs, replay_segs, peel_paths, peel_segs, scatt, rf, det, fallbacks;
};

// Shared-memory staging of the per-axis border tables (Cartesian borders or octree lattice tables).
struct SkSmemTables {
    const double *X, *Y, *Z;
};

// ---------------------------------------------------------------------------------------------------
// Ray state + PathSegmentGenerator::moveInside (SKIRT/utils/PathSegmentGenerator.cpp:11-112)
// ---------------------------------------------------------------------------------------------------
struct SkRay {
    double rx, ry, rz, kx, ky, kz;
};

__device__ __forceinline__ bool sk_box_contains(const double* b, double x, double y, double z)
{
    return x >= b[0] && x <= b[3] && y >= b[1] && y <= b[4] && z >= b[2] && z <= b[5];  // Box.hpp:99-109
}

__device__ __noinline__ bool sk_move_inside(SkRay& g, const double* box, double eps, double& cumds_out)
{
    double cumds = 0.;
    if (g.rx <= box[0])
    {
        if (g.kx <= 0.0) return false;
        double ds = (box[0] - g.rx) / g.kx;
        g.rx = box[0] + eps;
        g.ry += g.ky * ds;
        g.rz += g.kz * ds;
        cumds += ds;
    }
    else if (g.rx >= box[3])
    {
        if (g.kx >= 0.0) return false;
        double ds = (box[3] - g.rx) / g.kx;
        g.rx = box[3] - eps;
        g.ry += g.ky * ds;
        g.rz += g.kz * ds;
        cumds += ds;
    }
    if (g.ry <= box[1])
    {
        if (g.ky <= 0.0) return false;
        double ds = (box[1] - g.ry) / g.ky;
        g.rx += g.kx * ds;
        g.ry = box[1] + eps;
        g.rz += g.kz * ds;
        cumds += ds;
    }
    else if (g.ry >= box[4])
    {
        if (g.ky >= 0.0) return false;
        double ds = (box[4] - g.ry) / g.ky;
        g.rx += g.kx * ds;
        g.ry = box[4] - eps;
        g.rz += g.kz * ds;
        cumds += ds;
    }
    if (g.rz <= box[2])
    {
        if (g.kz <= 0.0) return false;
        double ds = (box[2] - g.rz) / g.kz;
        g.rx += g.kx * ds;
        g.ry += g.ky * ds;
        g.rz = box[2] + eps;
        cumds += ds;
    }
    else if (g.rz >= box[5])
    {
        if (g.kz >= 0.0) return false;
        double ds = (box[5] - g.rz) / g.kz;
        g.rx += g.kx * ds;
        g.ry += g.ky * ds;
        g.rz = box[5] - eps;
        cumds += ds;
    }
    if (!sk_box_contains(box, g.rx, g.ry, g.rz)) return false;
    cumds_out = cumds;
    return true;
}

// ---------------------------------------------------------------------------------------------------
// Octree cell location: TreeNode::leafChild (TreeNode.cpp:65-76) + OctTreeNode::child (OctTreeNode.cpp:37-42),
// expressed on the integer lattice: a node is (ix,iy,iz,level), its centre is the lattice border at +half size.
// ---------------------------------------------------------------------------------------------------
struct SkTreePos {
    int m;             // cell index, -1 = outside
    int ix, iy, iz;    // lattice coordinates of the cell's lower corner
    int lev;
};

__device__ __forceinline__ void sk_tree_descend(const SkDevModel& M, const SkSmemTables& T, int node, int ix, int iy,
                                                int iz, int lev, double x, double y, double z, SkTreePos& out)
{
    int fc = __ldg(&M.node_child[node]);
    while (fc >= 0)
    {
        int half = 1 << (M.maxlevel - lev - 1);
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
        node = fc + l;
        fc = __ldg(&M.node_child[node]);
    }
    out.m = -(fc + 1);
    out.ix = ix;
    out.iy = iy;
    out.iz = iz;
    out.lev = lev;
}

__device__ __noinline__ void sk_tree_locate_root(const SkDevModel& M, const SkSmemTables& T, double x, double y,
                                                 double z, SkTreePos& out)
{
    if (!sk_box_contains(M.ext, x, y, z))
    {
        out.m = -1;
        return;
    }
    sk_tree_descend(M, T, 0, 0, 0, 0, 0, x, y, z, out);
}

__device__ __forceinline__ bool sk_tree_cell_contains(const SkDevModel& M, const SkSmemTables& T, const SkTreePos& p,
                                                      double x, double y, double z)
{
    int size = 1 << (M.maxlevel - p.lev);
    return x >= T.X[p.ix] && x <= T.X[p.ix + size] && y >= T.Y[p.iy] && y <= T.Y[p.iy + size] && z >= T.Z[p.iz]
           && z <= T.Z[p.iz + size];
}

// ---------------------------------------------------------------------------------------------------
// Path traversal.  `visit(m, dens, ds)` is called for every segment the reference generator would return
// (m = -1 for the empty segment in front of the grid); it returns false to stop the walk early.
//   Cartesian: CartesianSpatialGrid::MySegmentGenerator::next, CartesianSpatialGrid.cpp:95-162
//   Octree:    TreeSpatialGrid::MySegmentGenerator::next, TreeSpatialGrid.cpp:140-216, with
//              TreeNode::neighbor (TreeNode.cpp:103-112) served by the per-cell links
// ---------------------------------------------------------------------------------------------------
template <class Visit>
__device__ __forceinline__ void sk_trace_cartesian(const SkDevModel& M, const SkSmemTables& T, double rx, double ry,
                                                   double rz, double kx, double ky, double kz, Visit&& visit)
{
    SkRay g{rx, ry, rz, kx, ky, kz};
    double cumds = 0.;
    if (!sk_move_inside(g, M.ext, M.eps, cumds)) return;
    int i = sk_locate_clip(T.X, M.nx + 1, g.rx);
    int j = sk_locate_clip(T.Y, M.ny + 1, g.ry);
    int k = sk_locate_clip(T.Z, M.nz + 1, g.rz);
    if (cumds > 0.)
        if (!visit(-1, 0., cumds)) return;
    const int di = (kx < 0.0) ? -1 : 1, dj = (ky < 0.0) ? -1 : 1, dk = (kz < 0.0) ? -1 : 1;
    const int oi = (kx < 0.0) ? 0 : 1, oj = (ky < 0.0) ? 0 : 1, ok = (kz < 0.0) ? 0 : 1;
    const bool ux = fabs(kx) > 1e-15, uy = fabs(ky) > 1e-15, uz = fabs(kz) > 1e-15;
    while (true)
    {
        int m = k + M.nz * j + M.nz * M.ny * i;
        double dens = __ldg(&M.dens[m]);
        double xE = T.X[i + oi];
        double yE = T.Y[j + oj];
        double zE = T.Z[k + ok];
        double dsx = ux ? (xE - g.rx) / kx : DBL_MAX;
        double dsy = uy ? (yE - g.ry) / ky : DBL_MAX;
        double dsz = uz ? (zE - g.rz) / kz : DBL_MAX;
        double ds;
        bool outside;
        if (dsx <= dsy && dsx <= dsz)
        {
            ds = dsx;
            g.rx = xE;
            g.ry += ky * dsx;
            g.rz += kz * dsx;
            i += di;
            outside = (i >= M.nx || i < 0);
        }
        else if (dsy < dsx && dsy <= dsz)
        {
            ds = dsy;
            g.ry = yE;
            g.rx += kx * dsy;
            g.rz += kz * dsy;
            j += dj;
            outside = (j >= M.ny || j < 0);
        }
        else
        {
            ds = dsz;
            g.rz = zE;
            g.rx += kx * dsz;
            g.ry += ky * dsz;
            k += dk;
            outside = (k >= M.nz || k < 0);
        }
        if (!visit(m, dens, ds)) return;
        if (outside) return;
    }
}

template <class Visit>
__device__ __forceinline__ void sk_trace_tree(const SkDevModel& M, const SkSmemTables& T, SkLocalCounters& cnt,
                                              double rx, double ry, double rz, double kx, double ky, double kz,
                                              Visit&& visit)
{
    SkRay g{rx, ry, rz, kx, ky, kz};
    double cumds = 0.;
    if (!sk_move_inside(g, M.ext, M.eps, cumds)) return;
    SkTreePos p;
    sk_tree_locate_root(M, T, g.rx, g.ry, g.rz, p);
    if (cumds > 0.)
        if (!visit(-1, 0., cumds)) return;
    if (p.m < 0) return;  // cannot happen after a successful moveInside; the reference would dereference null
    const bool ux = fabs(kx) > 1e-15, uy = fabs(ky) > 1e-15, uz = fabs(kz) > 1e-15;
    const int wx = (kx < 0.0) ? 0 : 1, wy = (ky < 0.0) ? 2 : 3, wz = (kz < 0.0) ? 4 : 5;
    const double eps = M.eps;
    while (true)
    {
        // one 32-byte sector: density + the six neighbour links of the current cell
        const int4* rp = reinterpret_cast<const int4*>(&M.cells[p.m]);
        int4 a = __ldg(rp), b = __ldg(rp + 1);
        double dens = __hiloint2double(a.y, a.x);
        int size = 1 << (M.maxlevel - p.lev);
        double xnext = T.X[p.ix + ((kx < 0.0) ? 0 : size)];
        double ynext = T.Y[p.iy + ((ky < 0.0) ? 0 : size)];
        double znext = T.Z[p.iz + ((kz < 0.0) ? 0 : size)];
        double dsx = ux ? (xnext - g.rx) / kx : DBL_MAX;
        double dsy = uy ? (ynext - g.ry) / ky : DBL_MAX;
        double dsz = uz ? (znext - g.rz) / kz : DBL_MAX;
        double ds;
        int wall;
        if (dsx <= dsy && dsx <= dsz)
        {
            ds = dsx;
            wall = wx;
        }
        else if (dsy <= dsx && dsy <= dsz)
        {
            ds = dsy;
            wall = wy;
        }
        else
        {
            ds = dsz;
            wall = wz;
        }
        double adv = ds + eps;
        g.rx += kx * adv;
        g.ry += ky * adv;
        g.rz += kz * adv;
        const int m_old = p.m;
        bool go_on = visit(m_old, dens, ds);
        if (!go_on) return;

        // neighbour lookup through the link of the exit wall
        int link = wall == 0 ? a.z : wall == 1 ? a.w : wall == 2 ? b.x : wall == 3 ? b.y : wall == 4 ? b.z : b.w;
        SkTreePos q;
        q.m = -1;
        if (link >= 0)
        {
            int axis = wall >> 1;
            bool neg = !(wall & 1);
            if (!(link & SK_LINK_INTERNAL))
            {
                // leaf neighbour at the same or a coarser level
                q.lev = (link >> SK_LINK_LEVEL_SHIFT) & 15;
                q.m = link & SK_LINK_INDEX_MASK;
                int nsize = 1 << (M.maxlevel - q.lev);
                int mask = ~(nsize - 1);
                q.ix = p.ix & mask;
                q.iy = p.iy & mask;
                q.iz = p.iz & mask;
                int c = (axis == 0 ? p.ix : axis == 1 ? p.iy : p.iz);
                c = neg ? c - nsize : c + size;
                if (axis == 0)
                    q.ix = c;
                else if (axis == 1)
                    q.iy = c;
                else
                    q.iz = c;
            }
            else
            {
                // internal neighbour node of the same level: descend to the leaf that holds the new position
                int ix = p.ix, iy = p.iy, iz = p.iz;
                int shift = neg ? -size : size;
                if (axis == 0)
                    ix += shift;
                else if (axis == 1)
                    iy += shift;
                else
                    iz += shift;
                sk_tree_descend(M, T, link & SK_LINK_INDEX_MASK, ix, iy, iz, p.lev, g.rx, g.ry, g.rz, q);
            }
            if (!sk_tree_cell_contains(M, T, q, g.rx, g.ry, g.rz)) q.m = -1;
        }
        if (q.m < 0)
        {
            // `if (!_node) _node = _grid->root()->leafChild(r())`, TreeSpatialGrid.cpp:193
            if (sk_box_contains(M.ext, g.rx, g.ry, g.rz))
            {
                cnt.fallbacks++;
                sk_tree_locate_root(M, T, g.rx, g.ry, g.rz, q);
            }
        }
        if (q.m == m_old)
        {
            // PathSegmentGenerator::propagateToNextAfter, PathSegmentGenerator.hpp:148-153
            g.rx = nextafter(g.rx, (kx < 0.) ? -DBL_MAX : DBL_MAX);
            g.ry = nextafter(g.ry, (ky < 0.) ? -DBL_MAX : DBL_MAX);
            g.rz = nextafter(g.rz, (kz < 0.) ? -DBL_MAX : DBL_MAX);
            sk_tree_locate_root(M, T, g.rx, g.ry, g.rz, q);
        }
        if (q.m < 0 || q.m == m_old) return;
        p = q;
    }
}

template <int GRID, class Visit>
__device__ __forceinline__ void sk_trace(const SkDevModel& M, const SkSmemTables& T, SkLocalCounters& cnt, double rx,
                                         double ry, double rz, double kx, double ky, double kz, Visit&& visit)
{
    if (GRID == 1)
        sk_trace_cartesian(M, T, rx, ry, rz, kx, ky, kz, visit);
    else
        sk_trace_tree(M, T, cnt, rx, ry, rz, kx, ky, kz, visit);
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

__device__ __noinline__ void sk_launch_primary(const SkDevModel& M, SkRng& g, unsigned long long history, SkPacket& pp)
{
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
    pp.nscatt = 0;
    pp.primary_origin = 1;
    pp.ilam = sk_locate_clip(M.lam_border, M.nlam, lambda);
    pp.sig_ext = M.sig_ext[pp.ilam];
}

// ---------------------------------------------------------------------------------------------------
// Instruments: Instrument::detect + FluxRecorder::detect (FluxRecorder.cpp:304-468)
// ---------------------------------------------------------------------------------------------------
struct SkPeel {
    double W;       // weight of the peel-off packet
    int nscatt;
    bool has_tau;
    double tau;
};

template <int GRID>
__device__ __forceinline__ double sk_observed_optical_depth(const SkDevModel& M, const SkSmemTables& T,
                                                            SkLocalCounters& cnt, const SkPacket& pp, double W,
                                                            const double* kobs)
{
    // MediumSystem::getExtinctionOpticalDepth(pp, infinity), MediumSystem.cpp:1192-1219
    double L = W / pp.lambda;
    if (L <= 0) return INFINITY;
    double taumax = log(L) + 745;
    double tau = 0.;
    const double section = pp.sig_ext;
    bool inf = false;
    unsigned int nseg = 0;
    sk_trace<GRID>(M, T, cnt, pp.rx, pp.ry, pp.rz, kobs[0], kobs[1], kobs[2], [&](int m, double dens, double ds) {
        nseg++;
        if (m >= 0)
        {
            tau += section * dens * ds;
            if (tau >= taumax)
            {
                inf = true;
                return false;
            }
        }
        return true;
    });
    cnt.peel_paths++;
    cnt.peel_segs += nseg;
    return inf ? INFINITY : tau;
}

template <int GRID>
__device__ __forceinline__ void sk_detect(const SkDevModel& M, const SkSmemTables& T, SkLocalCounters& cnt,
                                          const SkDevInstr& q, const SkPacket& pp, SkPeel& peel, double* hist_w,
                                          int* hist_ell, int j)
{
    int l = 0;
    double x = pp.rx, y = pp.ry, z = pp.rz;
    if (q.kind == SK_INSTR_SED)
    {
        if (q.radius2)
        {
            double xpp = -q.sinphi * x + q.cosphi * y;  // ApertureInstrument.cpp:24-43
            double ypp = -q.cosphi * q.costheta * x - q.sinphi * q.costheta * y + q.sintheta * z;
            double radius2 = xpp * xpp + ypp * ypp;
            if (radius2 > q.radius2) return;
        }
    }
    else
    {
        double xpp = -q.sinphi * x + q.cosphi * y;  // FrameInstrument::pixelOnDetector, FrameInstrument.cpp:45-64
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
    if (!q.include_sed && l < 0) return;
    int ell = sk_wlg_bin(M.wlg[q.wlg], pp.lambda);
    if (ell < 0) return;

    double L = peel.W / pp.lambda;
    if (!peel.has_tau)
    {
        peel.tau = sk_observed_optical_depth<GRID>(M, T, cnt, pp, peel.W, q.kobs);
        peel.has_tau = true;
    }
    double Lext = L * exp(-peel.tau);
    cnt.det++;

    // component routing, FluxRecorder.cpp:345-380
    int c_ext, c_tr = -1, c_lev = -1;
    if (q.record_total_only)
        c_ext = SK_COMP_TOTAL;
    else if (pp.primary_origin)
    {
        if (peel.nscatt == 0)
        {
            c_tr = SK_COMP_TRANSPARENT;
            c_ext = SK_COMP_PRIMARY_DIRECT;
        }
        else
        {
            c_ext = SK_COMP_PRIMARY_SCATTERED;
            if (peel.nscatt <= q.num_levels) c_lev = SK_COMP_PRIMARY_SCATTERED_LEVEL + peel.nscatt - 1;
        }
    }
    else
    {
        if (peel.nscatt == 0)
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
    if (q.record_stats && q.include_sed)
    {
        hist_w[j] += Lext;  // FluxRecorder.cpp:457-466: contributions of one history are summed per bin first
        hist_ell[j] = ell;
    }
}

// MonteCarloSimulation::peelOffEmission (.cpp:617-634) and peelOffScattering (.cpp:784-842, consolidated branch)
template <int GRID>
__device__ __forceinline__ void sk_peel_off(const SkDevModel& M, const SkSmemTables& T, SkLocalCounters& cnt,
                                            const SkPacket& pp, bool scattering, double* hist_w, int* hist_ell)
{
    SkPeel peel;
    peel.W = 0.;
    peel.nscatt = 0;
    peel.has_tau = false;
    peel.tau = 0.;
    for (int j = 0; j < M.ninstr; ++j)
    {
        const SkDevInstr& q = M.instr[j];
        if (!q.same_as_preceding)
        {
            if (scattering)
            {
                // DustMix::peeloffScattering HG branch (DustMix.cpp:430-445); MediumSystem::peelOffScattering
                // (MediumSystem.cpp:734-767) with the single-medium weight 1; launchScatteringPeelOff (PhotonPacket.cpp:89-103)
                double costheta = pp.kx * q.kobs[0] + pp.ky * q.kobs[1] + pp.kz * q.kobs[2];
                double gp = M.gpar[pp.ilam];
                double value = fabs(gp) > 0.95 ? sk_mean_hg(gp, costheta) : sk_value_hg(gp, costheta);
                double I = 0.;
                I += value * 1.;
                peel.W = pp.W * I;
                peel.nscatt = pp.nscatt + 1;
            }
            else
            {
                peel.W = pp.W;  // launchEmissionPeelOff, PhotonPacket.cpp:66-85 (isotropic emission)
                peel.nscatt = 0;
            }
            peel.has_tau = false;
        }
        sk_detect<GRID>(M, T, cnt, q, pp, peel, hist_w, hist_ell, j);
    }
}

// ---------------------------------------------------------------------------------------------------
// One complete history.
// ---------------------------------------------------------------------------------------------------
template <int GRID>
__device__ __forceinline__ void sk_life_cycle(const SkDevModel& M, const SkSmemTables& T, const SkRunArgs& A,
                                              SkLocalCounters& cnt, unsigned long long history)
{
    SkRng g;
    sk_rng_init(g, M.seed, A.stream_id, history);
    SkPacket pp;
    sk_launch_primary(M, g, history, pp);

    double hist_w[SK_MAX_INSTR];
    int hist_ell[SK_MAX_INSTR];
#pragma unroll
    for (int j = 0; j < SK_MAX_INSTR; ++j)
    {
        hist_w[j] = 0.;
        hist_ell[j] = -1;
    }

    if (pp.W / pp.lambda > 0)
    {
        cnt.packets++;
        if (A.peel) sk_peel_off<GRID>(M, T, cnt, pp, false, hist_w, hist_ell);

        const double sig_sca = M.sig_sca[pp.ilam];
        const double gp = M.gpar[pp.ilam];
        int rf_ell = -1;
        double* rf = nullptr;
        if (A.store)
        {
            rf_ell = sk_wlg_bin(M.wlg[M.rf_grid], pp.lambda);  // MonteCarloSimulation.cpp:643
            rf = A.primary ? M.rf1 : M.rf2c;
        }

        if (M.force_scattering)
        {
            const double Lthreshold = (pp.W / pp.lambda) / M.min_weight_reduction;
            while (true)
            {
                // ---- pass 1: MediumSystem::setExtinctionOpticalDepths (MediumSystem.cpp:849-871) fused with
                //      MonteCarloSimulation::storeRadiationField (.cpp:638-665)
                double tau = 0., s = 0.;
                int last_m = -1;
                unsigned int nseg = 0, nrf = 0;
                const double section = pp.sig_ext;
                const double luminosity = pp.W / pp.lambda;
                double lnExtBeg = 0., extBeg = 1.;
                const bool deposit = rf_ell >= 0;
                sk_trace<GRID>(M, T, cnt, pp.rx, pp.ry, pp.rz, pp.kx, pp.ky, pp.kz, [&](int m, double dens, double ds) {
                    if (ds > 0.)  // SpatialGridPath::addSegment, SpatialGridPath.cpp:41-48
                    {
                        nseg++;
                        s += ds;
                        if (m >= 0) tau += section * dens * ds;
                        last_m = m;
                        if (deposit)
                        {
                            double lnExtEnd = -tau;
                            double extEnd = exp(lnExtEnd);
                            if (m >= 0)
                            {
                                double extMean = sk_lnmean4(extEnd, extBeg, lnExtEnd, lnExtBeg);
                                double Lds = luminosity * extMean * ds;
                                atomicAdd(&rf[(size_t)m * M.nrf + rf_ell], Lds);  // MediumSystem.cpp:1294-1300
                                nrf++;
                            }
                            lnExtBeg = lnExtEnd;
                            extBeg = extEnd;
                        }
                    }
                    return true;
                });
                cnt.fwd_paths++;
                cnt.fwd_segs += nseg;
                cnt.rf += nrf;

                // ---- MonteCarloSimulation::simulateForcedPropagation, .cpp:696-742
                const double taupath = tau;
                if (!(nseg > 0 && taupath > 0.))
                {
                    pp.W *= 0.;
                    break;
                }
                double xi = M.path_length_bias;
                double tauint;
                if (xi == 0.)
                    tauint = sk_expon_cutoff(g, taupath);
                else
                {
                    tauint = sk_uniform(g) < xi ? sk_uniform(g) * taupath : sk_expon_cutoff(g, taupath);
                    double p = -exp(-tauint) / expm1(-taupath);
                    double q = (1.0 - xi) * p + xi / taupath;
                    pp.W *= p / q;
                }
                // ---- pass 2: SpatialGridPath::findInteractionPoint (SpatialGridPath.cpp:164-206) by re-walking
                int m_int = last_m;
                double s_int = s;
                {
                    double tau2 = 0., s2 = 0.;
                    unsigned int nre = 0;
                    sk_trace<GRID>(M, T, cnt, pp.rx, pp.ry, pp.rz, pp.kx, pp.ky, pp.kz,
                                   [&](int m, double dens, double ds) {
                                       if (ds > 0.)
                                       {
                                           nre++;
                                           double tau0 = tau2, s0 = s2;
                                           s2 += ds;
                                           if (m >= 0) tau2 += section * dens * ds;
                                           if (tauint < tau2)
                                           {
                                               m_int = m;
                                               s_int = sk_interp_linlin(tauint, tau0, tau2, s0, s2);
                                               return false;
                                           }
                                       }
                                       return true;
                                   });
                    cnt.replay_segs += nre;
                }
                // MediumSystem::albedoForScattering, MediumSystem.cpp:678-693
                double albedo = 0.;
                if (m_int >= 0)
                {
                    double n = GRID == 1 ? M.dens[m_int] : M.cells[m_int].dens;
                    double ksca = n * sig_sca;
                    double kext = n * pp.sig_ext;
                    albedo = kext > 0. ? ksca / kext : 0.;
                }
                pp.W *= -expm1(-taupath) * albedo;
                pp.rx += s_int * pp.kx;  // PhotonPacket::propagate, PhotonPacket.cpp:107-111
                pp.ry += s_int * pp.ky;
                pp.rz += s_int * pp.kz;

                double L = pp.W / pp.lambda;
                if (L <= 0 || (L <= Lthreshold && pp.nscatt >= M.min_scatt_events)) break;

                if (A.peel) sk_peel_off<GRID>(M, T, cnt, pp, true, hist_w, hist_ell);

                // MediumSystem::simulateScattering (.cpp:796-823) + DustMix::performScattering HG (DustMix.cpp:496-511)
                if (fabs(gp) < 1e-6)
                    sk_random_direction(g, pp.kx, pp.ky, pp.kz);
                else
                {
                    double f = ((1.0 - gp) * (1.0 + gp)) / (1.0 - gp + 2.0 * gp * sk_uniform(g));
                    double costheta = (1.0 + gp * gp - f * f) / (2.0 * gp);
                    sk_random_direction_about(g, pp.kx, pp.ky, pp.kz, costheta);
                }
                pp.nscatt++;
                cnt.scatt++;
            }
        }
        else
        {
            // non-forced scattering: simulateNonForcedPropagation (.cpp:746-780) with
            // MediumSystem::setInteractionPointUsingExtinction (MediumSystem.cpp:978-1010)
            while (true)
            {
                double tauinteract = -log(sk_uniform(g));  // Random::expon
                double tau = 0., s = 0.;
                bool found = false;
                int m_int = -1;
                double s_int = 0.;
                unsigned int nseg = 0;
                const double section = pp.sig_ext;
                sk_trace<GRID>(M, T, cnt, pp.rx, pp.ry, pp.rz, pp.kx, pp.ky, pp.kz, [&](int m, double dens, double ds) {
                    nseg++;
                    double tau0 = tau, s0 = s;
                    if (m >= 0) tau += section * dens * ds;
                    s += ds;
                    if (tauinteract < tau)
                    {
                        found = true;
                        m_int = m;
                        s_int = sk_interp_linlin(tauinteract, tau0, tau, s0, s);
                        return false;
                    }
                    return true;
                });
                cnt.fwd_paths++;
                cnt.fwd_segs += nseg;
                if (!found) break;
                double n = GRID == 1 ? M.dens[m_int] : M.cells[m_int].dens;
                double ksca = n * sig_sca;
                double kext = n * pp.sig_ext;
                pp.W *= kext > 0. ? ksca / kext : 0.;
                pp.rx += s_int * pp.kx;
                pp.ry += s_int * pp.ky;
                pp.rz += s_int * pp.kz;
                if (pp.W / pp.lambda <= 0) break;
                if (A.peel) sk_peel_off<GRID>(M, T, cnt, pp, true, hist_w, hist_ell);
                if (fabs(gp) < 1e-6)
                    sk_random_direction(g, pp.kx, pp.ky, pp.kz);
                else
                {
                    double f = ((1.0 - gp) * (1.0 + gp)) / (1.0 - gp + 2.0 * gp * sk_uniform(g));
                    double costheta = (1.0 + gp * gp - f * f) / (2.0 * gp);
                    sk_random_direction_about(g, pp.kx, pp.ky, pp.kz, costheta);
                }
                pp.nscatt++;
                cnt.scatt++;
            }
        }
    }

    // FluxRecorder::recordContributions for the SED arrays, FluxRecorder.cpp:962-986
    for (int j = 0; j < M.ninstr; ++j)
    {
        if (hist_ell[j] >= 0)
        {
            const SkDevInstr& q = M.instr[j];
            double wn = 1.;
            for (int k = 0; k <= 4; ++k)
            {
                atomicAdd(&q.wsed[k][hist_ell[j]], wn);
                wn *= hist_w[j];
            }
        }
    }
}
