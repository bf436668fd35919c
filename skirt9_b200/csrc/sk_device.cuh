// sk_device.cuh -- device-side data layout and building blocks of the photon life-cycle kernel (sm_100a).
//
// Layout in HBM:
//   * octree: one 32-byte record per CELL {double density; int32 link[6]} -- exactly one 32 B sector per cell
//     crossing; the cell geometry is NOT stored per cell: a cell is identified by its integer lattice coordinates
//     (ix,iy,iz at the finest level, carried in registers) and its level, and a ray is walked in lattice coordinates
//     u = (r - min) / pitch, in which the cell borders are those integers (no border look-ups in the crossing loop; the
//     per-axis border tables X/Y/Z[2^maxLevel+1] remain for the set-up kernels).  link[w] is the same-level-or-coarser
//     neighbour across wall w: a cell index (+ its level) when that neighbour is a leaf, an internal node id
//     otherwise (then a short descent through the 4-byte-per-node child table follows), -1 at the domain boundary.
//   * Cartesian grid: border arrays in shared memory, density[m] 8 B per crossing, neighbours by index arithmetic.
//   * tallies: fp64 arrays updated with native RED.F64 atomics.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#include "../../include/sk_engine.h"

#define SK_MAX_LEVELS 8
#define SK_NUM_COMP (SK_COMP_PRIMARY_SCATTERED_LEVEL + SK_MAX_LEVELS)
#define SK_MAX_INSTR 8
#ifndef SK_PIX_K
#define SK_PIX_K 32  // distinct (pixel, bin) entries of a history kept in its bank slot; longer lists continue in chunks of
#endif
#ifndef SK_PIX_C
#define SK_PIX_C 32  // this many entries taken from a pool in device memory (SkBank::pool_*)
#endif
#define SK_MAX_TREE_LEVEL 15
#define SK_LINK_INTERNAL 0x40000000
#define SK_LINK_LEVEL_SHIFT 26
#define SK_LINK_INDEX_MASK 0x03FFFFFF

struct __align__(32) SkCellRec {
    double dens;
    int32_t link[6];
};

struct SkDevWlg {
    int32_t num_bins, num_borders;
    const double* borders;
    const int32_t* ell;
    const double* lambda;
    const double* dlambda;
};

struct SkDevSource {
    int32_t kind, geometry, sed_kind, bias_kind;
    double position[3];
    double gp[SK_GEOM_MAX_PARAMS];
    int32_t geom_table_n, sed_n, oligo_n, pad;
    const double *geom_table_x, *geom_table_P;
    const double *sed_lambda, *sed_p, *sed_P;
    const double* oligo_lambda;
    double sed_temperature, sed_norm, wavelength_bias, bias_min, bias_max, oligo_probability;
    double Lw;  // _Lv[h]/_Wv[h]  (SourceSystem.cpp:105)
    int32_t velocity_kind, pad3;  // sk_velocity_kind and its parameters (sk_source_t::velocity)
    double velocity[3];
};

struct SkDevInstr {
    int32_t kind, wlg, nx, ny, num_levels, record_total_only, record_stats, same_as_preceding;
    int32_t include_sed, include_ifu, nl, pad;
    unsigned long long npix;
    double kobs[3];
    double costheta, sintheta, cosphi, sinphi, cosomega, sinomega;
    double xpmin, xpsiz, ypmin, ypsiz, radius2;
    double zp1;         // 1 + observer-frame redshift (FluxRecorder.cpp:310)
    double* sed[SK_NUM_COMP];
    double* ifu[SK_NUM_COMP];
    double* wsed[5];
    double* wifu[5];    // per-pixel statistics (FluxRecorder::_wifu) or null
    int32_t pix_slot;   // index of this instrument's per-history pixel list in the bank, or -1
    int32_t sed_slot;   // index of its per-history list of SED bins, used in runs with kinematics (the peel-off packets of one
                        // history then differ in wavelength, FluxRecorder.cpp:962-986), or -1
};

struct SkDevModel {
    // configuration
    uint32_t seed;
    int32_t force_scattering, min_scatt_events;
    int32_t explicit_absorption;  // interaction points in scattering optical depth, weight exp(-tau_abs) (.cpp:567-570,727-731)
    double path_length_bias, min_weight_reduction;
    // grid
    int32_t grid_kind;  // 1 cartesian, 2 octree, 3 voronoi
    int32_t nx, ny, nz; // cartesian bins; octree: lattice size per axis (2^maxlevel) in nx
    int32_t maxlevel;
    int32_t lattice_in_smem;
    double ext[6];
    double eps;
    double eps4;                 // 4 eps: exit distances of two walls closer than this take the full neighbour search
    double lat_h[3];             // octree: pitch of the finest lattice per axis, (max - min) / 2^maxlevel
    double lat_invh[3];          // and its reciprocal
    const double *xv, *yv, *zv;  // cartesian borders or octree lattice tables
    const double* dens;          // cartesian: density per cell
    const SkCellRec* cells;      // octree
    const int32_t* node_child;   // octree: first child node id, or -(cell+1) for a leaf
    const uint32_t* cell_coord;  // octree: 4 x uint32 per cell {ix,iy,iz,level}
    int32_t ncells, nnodes;
    // voronoi mesh (grid_kind 3): one 32-byte record per cell {site x,y,z; density}, CSR neighbour lists (cell index, or
    // -1..-6 for the domain walls), and the start-cell table of the nearest-site walk
    const double4* vrec;
    const long long* vnbr_off;
    const int32_t* vnbr;
    const double4* vnrec;  // per neighbour entry (same order as vnbr): {site x,y,z of the neighbour; its index in the low 32 bits
                           // of w}, so that a crossing reads its ~15 candidates as one contiguous run of 32-byte records
                           // instead of an index list plus as many scattered site records
    const int32_t* vblock;
    const double* vbox;  // [6*ncells] enclosing boxes of the Voronoi cells (dust emission only), or null
    int32_t vnb;
    const double* volume;        // cell volumes (MediumState::volume)
    // several medium components with their own material mixes (sk_engine_set_media): the density of component 0 lives where
    // the single medium's does (cell record / dens[]), the others in densx[(h-1)*ncells + m]
    int32_t nmed;
    const double* densx;
    // dust: one table set per component, [h*nlam + i]; lam_border is common to all mixes (DustMix.cpp:52-98)
    int32_t nlam;
    const double *lam_border, *sig_abs, *sig_sca, *sig_ext, *gpar;
    // wavelength grids
    const SkDevWlg* wlg;
    int32_t nwlg, rf_grid, nrf;
    double *rf1, *rf2, *rf2c;    // radiation field tables, wavelength-major on the device: SK_RF_INDEX
    // sources
    const SkDevSource* src;
    const unsigned long long* Iv;
    int32_t nsrc;
    double Lpp;
    // secondary (dust) emission, sk_secondary.cuh
    int32_t sec_nem, sec_nT;         // points of the extended emission grid; size of the temperature grid
    const double *sec_lambda, *sec_emsig, *sec_rfsig, *sec_T, *sec_planckabs;  // per component: [h*sec_nem+i], [h*nrf+ell], -, [h*sec_nT+i]
    const double* sec_cmb;           // [nrf] CMB source term of the energy balance (zeros without CMB heating)
    const double* sec_kabs_rf;       // sigma_abs at the characteristic wavelengths of the radiation field grid, [h*nrf + ell]
    double *sec_pv, *sec_Pv;         // [ncells][sec_nem] normalised emission spectrum and its cdf
    double *sec_Lv, *sec_ws;         // [ncells] absorbed luminosity; launch weight _Lv[m]/_Wv[m]
    unsigned long long* sec_Iv;      // [ncells+1] history index -> cell map
    double sec_Lpp, sec_xi, sec_bias_min, sec_bias_max;
    // instruments
    const SkDevInstr* instr;
    int32_t ninstr;
    int32_t pix_base_d, pix_base_i;  // first bank field of the per-history pixel lists: per instrument SK_PIX_K doubles and
                                     // SK_PIX_K + 2 ints (pixel-bin indices, number of entries, newest pool chunk or -1)
    // counters
    unsigned long long* counters;
    // kinematics (sk_engine_set_velocities; also set when only sources move): vel[m] = MediumState::bulkVelocity(m),
    // all zero for media at rest.  The walks then run in the several-component instantiation with per-cell section look-ups,
    // and the bank holds the extra per-packet fields SK_KD_* / SK_KI_* from kin_base_d / kin_base_i on.
    int32_t kin, kin_base_d, kin_base_i, kin_pad;
    const double4* vel;  // {vx, vy, vz, 0} per cell: one 32-byte request per crossing, like the cell record
};

// Radiation field tables on the device are wavelength-major, rf[ell * ncells + m] (the reference's Table<2> is [m][ell],
// MediumSystem.hpp:883-885; sk_engine_read_rf transposes): the forward trace walks its rays in order of their wavelength
// bin, so that the deposits of the rays in flight go to a few slices of ncells doubles each that stay in the L2 cache,
// instead of to random 32-byte sectors of a table several times the size of the L2 (a read-modify-write in HBM each).
#define SK_RF_INDEX(M, m, ell) ((size_t)(ell) * (size_t)(M).ncells + (size_t)(m))

struct SkRunArgs {
    unsigned long long first, count;   // histories of this engine: count indices h, mapped to first + sk_history_of(h)
    unsigned long long il_stride, il_offset;  // interleaved sharding: h -> (h >> il_shift) * il_stride + il_offset + (h & mask)
    int32_t il_shift;                  // log2 of the block length, or -1 for a contiguous range
    int32_t primary, peel, store;
    uint32_t stream_id;
    unsigned long long* work_counter;  // dynamic history dispenser
    const SkDevModel* model;           // copy of the model in global memory for the cold, non-inlined paths
};

// the history index (relative to A.first) of the h-th history this engine runs (sk_engine_set_history_interleave)
__device__ __forceinline__ unsigned long long sk_history_of(const SkRunArgs& A, unsigned long long h)
{
    if (A.il_shift < 0) return h;
    return (h >> A.il_shift) * A.il_stride + A.il_offset + (h & ((1ull << A.il_shift) - 1ull));
}

// ---------------------------------------------------------------------------------------------------
// TMA bulk copy global -> shared memory (cp.async.bulk, SASS UBLKCP) completing on an mbarrier: stages the border tables
// of the trace kernels.  Sizes and both addresses are multiples of 16 bytes.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sk_shared_addr(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void sk_mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sk_shared_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void sk_mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sk_shared_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sk_tma_load_bulk(void* dst_shared, const void* src_global, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     sk_shared_addr(dst_shared)),
                 "l"(src_global), "r"(bytes), "r"(sk_shared_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void sk_mbar_wait(unsigned long long* bar, unsigned phase)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SK_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SK_MBAR_DONE;\n"
        "bra SK_MBAR_WAIT;\n"
        "SK_MBAR_DONE:\n"
        "}\n" ::"r"(sk_shared_addr(bar)),
        "r"(phase)
        : "memory");
}

// ---------------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator (replaces Random.cpp:20-56); identical to oracle/sk_oracle.c
// ---------------------------------------------------------------------------------------------------
struct SkRng {
    uint32_t k0, k1, c0, c1, draw;
};

__device__ __forceinline__ void sk_philox(uint32_t c[4], uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r)
    {
        uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0;
        uint32_t n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0;
        c[1] = lo1;
        c[2] = n2;
        c[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ double sk_u01(uint32_t lo, uint32_t hi)
{
    unsigned long long x = ((unsigned long long)hi << 32) | lo;
    return ((double)(x >> 12) + 0.5) * (1.0 / 4503599627370496.0);
}
__device__ __forceinline__ void sk_rng_init(SkRng& g, uint32_t seed, uint32_t stream, unsigned long long history,
                                            uint32_t draw)
{
    g.k0 = seed;
    g.k1 = stream;
    g.c0 = (uint32_t)history;
    g.c1 = (uint32_t)(history >> 32);
    g.draw = draw;
}
// Random::uniform, Random.cpp:70-73: the n-th deviate of a history is Philox(counter = (history, n, 0))
#ifndef SK_UNIFORM_INLINE
#define SK_UNIFORM_INLINE __noinline__
#endif
__device__ SK_UNIFORM_INLINE double sk_uniform(SkRng& g)
{
    uint32_t c[4] = {g.c0, g.c1, g.draw, 0u};
    sk_philox(c, g.k0, g.k1);
    g.draw++;
    return sk_u01(c[0], c[1]);
}
// Random::exponCutoff, Random.cpp:105-117
__device__ __forceinline__ double sk_expon_cutoff(SkRng& g, double xmax)
{
    if (xmax == 0.0)
        return 0.0;
    else if (xmax < 1e-10)
        return sk_uniform(g) * xmax;
    double x = -log(1.0 - sk_uniform(g) * (1.0 - exp(-xmax)));
    while (x > xmax) x = -log(1.0 - sk_uniform(g) * (1.0 - exp(-xmax)));
    return x;
}
// Direction::Direction(theta,phi), SKIRT/utils/Direction.cpp:10-35
__device__ __forceinline__ void sk_direction_from_angles(double theta, double phi, double& kx, double& ky, double& kz)
{
    const double eps = 1e-8;
    if (theta <= eps)
    {
        kx = 0;
        ky = 0;
        kz = 1;
    }
    else if (theta >= M_PI - eps)
    {
        kx = 0;
        ky = 0;
        kz = -1;
    }
    else
    {
        double sintheta = sin(theta);
        kx = sintheta * cos(phi);
        ky = sintheta * sin(phi);
        kz = cos(theta);
    }
}
// Random::direction(), Random.cpp:121-126
__device__ __forceinline__ void sk_random_direction(SkRng& g, double& kx, double& ky, double& kz)
{
    double theta = acos(2.0 * sk_uniform(g) - 1.0);
    double phi = 2.0 * M_PI * sk_uniform(g);
    sk_direction_from_angles(theta, phi, kx, ky, kz);
}
// Random::direction(bfk, costheta), Random.cpp:130-164
__device__ __forceinline__ void sk_random_direction_about(SkRng& g, double& kx, double& ky, double& kz, double costheta)
{
    double phi = 2.0 * M_PI * sk_uniform(g);
    double cosphi = cos(phi);
    double sinphi = sin(phi);
    double sintheta = sqrt(fabs((1.0 - costheta) * (1.0 + costheta)));
    double nx, ny, nz;
    if (kz > 0.99999)
    {
        nx = cosphi * sintheta;
        ny = sinphi * sintheta;
        nz = costheta;
    }
    else if (kz < -0.99999)
    {
        nx = cosphi * sintheta;
        ny = sinphi * sintheta;
        nz = -costheta;
    }
    else
    {
        double root = sqrt((1.0 - kz) * (1.0 + kz));
        nx = sintheta / root * (-kx * kz * cosphi + ky * sinphi) + kx * costheta;
        ny = -sintheta / root * (ky * kz * cosphi + kx * sinphi) + ky * costheta;
        nz = root * sintheta * cosphi + kz * costheta;
    }
    kx = nx;
    ky = ny;
    kz = nz;
}

// ---------------------------------------------------------------------------------------------------
// numerical helpers (NR.hpp:130-191,328-358; SpecialFunctions.cpp:578-627,822-880)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int sk_locate_basic(const double* __restrict__ xv, double x, int n)
{
    int jl = -1, ju = n;
    while (ju - jl > 1)
    {
        int jm = (ju + jl) >> 1;
        if (x < xv[jm])
            ju = jm;
        else
            jl = jm;
    }
    return jl;
}
__device__ __forceinline__ int sk_locate_clip(const double* __restrict__ xv, int n, double x)
{
    if (x < xv[0]) return 0;
    return sk_locate_basic(xv, x, n - 1);
}
__device__ __forceinline__ int sk_locate_fail(const double* __restrict__ xv, int n, double x)
{
    if (x > xv[n - 1]) return -1;
    return sk_locate_basic(xv, x, n - 1);
}
__device__ __forceinline__ double sk_interp_linlin(double x, double x1, double x2, double f1, double f2)
{
    return f1 + ((x - x1) / (x2 - x1)) * (f2 - f1);
}
__device__ __forceinline__ double sk_interp_loglog(double x, double x1, double x2, double f1, double f2)
{
    if (f1 <= 0 || f2 <= 0)
    {
        if (x == x1) return f1;
        if (x == x2) return f2;
        return 0;
    }
    return f1 * exp(log(x / x1) / log(x2 / x1) * (log(f2 / f1)));
}
// x^e for the small even integer exponents of SpiralStructureGeometryDecorator::perturbation (2N, N = index): x*x is the
// correctly rounded pow(x, 2); other exponents go through pow()
__device__ __forceinline__ double sk_pow_even(double x, double e)
{
    return e == 2.0 ? x * x : pow(x, e);
}
__device__ __forceinline__ double sk_gexp(double p, double x)
{
    const double q = 1.0 - p;
    if (q == 0.0)
        return exp(x);
    else if (fabs(q) < 1e-3)
    {
        double x2 = x * x;
        return exp(x)
               * (1.0 - 0.5 * x2 * q + 1.0 / 24.0 * x * x2 * (8.0 + 3.0 * x) * q * q
                  - 1.0 / 48.0 * x2 * x2 * (12.0 + 8.0 * x + x2) * q * q * q);
    }
    else
        return pow(1.0 + q * x, 1.0 / q);
}
__device__ __forceinline__ double sk_lnmean4(double x1, double x2, double lnx1, double lnx2)
{
    if (x1 > x2)
    {
        double t = x1;
        x1 = x2;
        x2 = t;
        t = lnx1;
        lnx1 = lnx2;
        lnx2 = t;
    }
    if (x1 <= 0) return 0.;
    double x = x2 / x1 - 1.;
    if (x < 1e-3)
    {
        return x1
               / (1. - 1. / 2. * x + 1. / 3. * x * x - 1. / 4. * x * x * x + 1. / 5. * x * x * x * x
                  - 1. / 6. * x * x * x * x * x);
    }
    else
        return (x2 - x1) / (lnx2 - lnx1);
}
__device__ __noinline__ double sk_lambert_w1(double z)
{
    const double eps = 1.0e-12;
    const double em1 = 0.3678794411714423215955237701614608;
    const double c[12] = {-1.0,
                          2.331643981597124203363536062168,
                          -1.812187885639363490240191647568,
                          1.936631114492359755363277457668,
                          -2.353551201881614516821543561516,
                          3.066858901050631912893148922704,
                          -4.175335600258177138854984177460,
                          5.858023729874774148815053846119,
                          -8.401032217523977370984161688514,
                          12.250753501314460424,
                          -18.100697012472442755,
                          27.029044799010561650};
    if (z == 0.0) return -DBL_MAX;
    double q = z + em1;
    double r = -sqrt(q);
    double t8 = c[8] + r * (c[9] + r * (c[10] + r * c[11]));
    double t5 = c[5] + r * (c[6] + r * (c[7] + r * t8));
    double t1 = c[1] + r * (c[2] + r * (c[3] + r * (c[4] + r * t5)));
    double w0 = c[0] + r * t1;
    if (q < 3.0e-3) return w0;
    double w, e, p, t;
    if (z < -1e-6)
        w = w0;
    else
    {
        double l1 = log(-z);
        double l2 = log(-l1);
        w = l1 - l2 + l2 / l1;
    }
    for (int i = 0; i < 10; i++)
    {
        e = exp(w);
        t = w * e - z;
        p = w + 1.0;
        t /= e * p - 0.5 * (p + 1.0) * t / p;
        w -= t;
        if (fabs(t) < eps * (1.0 + fabs(w))) return w;
    }
    return w;
}
// PlanckFunction::value, SKIRT/utils/PlanckFunction.cpp:24-27
__device__ __forceinline__ double sk_planck(double lambda, double T)
{
    const double h = 6.62606957e-34, c = 2.99792458e8, k = 1.3806488e-23;
    double f1 = h * c / (k * T);
    double f2 = 2.0 * h * c * c;
    const double l2 = lambda * lambda;  // lambda^5 by multiplication (the oracle does the same): pow() is the most
    return f2 / (l2 * l2 * lambda) / (exp(f1 / lambda) - 1.0);  // expensive call of the launch kernel
}
// the same with a guess: i is returned when it is the answer (two comparisons), else the search runs.  The walks with
// kinematics look the dust tables up in every cell, and the perceived wavelength moves by parts in a thousand from cell to cell.
// (a few steps up or down the table first: the answer is the index i with (i == 0 or xv[i] <= x) and (i == n-2 or x < xv[i+1]),
//  however it is found)
__device__ __forceinline__ int sk_locate_clip_hint(const double* __restrict__ xv, int n, double x, int i)
{
#pragma unroll 1
    for (int step = 0; step < 6; ++step)
    {
        if (i > 0 && x < xv[i])
            --i;
        else if (i < n - 2 && !(x < xv[i + 1]))
            ++i;
        else
            return i;
    }
    return sk_locate_clip(xv, n, x);
}
// Doppler shifts, PhotonPacket.cpp:133-151 (no Hubble flow): the wavelength a packet of rest wavelength lambda leaves with in
// direction k from an emitter moving with v, and the wavelength a receiver moving with v perceives
#define SK_C_LIGHT 299792458.  // Constants::c()
__device__ __forceinline__ double sk_shifted_emission(double lambda, double kx, double ky, double kz, double vx, double vy,
                                                      double vz)
{
    return lambda * (1 - (kx * vx + ky * vy + kz * vz) / SK_C_LIGHT);
}
__device__ __forceinline__ double sk_perceived(double lambda, double kx, double ky, double kz, double vx, double vy, double vz)
{
    return lambda / (1 - (kx * vx + ky * vy + kz * vz) / SK_C_LIGHT);
}
// DisjointWavelengthGrid::bin, DisjointWavelengthGrid.cpp:332-341
// (the border index std::upper_bound returns, with a guess: see sk_locate_clip_hint)
__device__ __forceinline__ int sk_wlg_upper_hint(const SkDevWlg& g, double lambda, int lo0)
{
    // (the answer is the index lo with (lo == 0 or borders[lo-1] <= lambda) and (lo == num_borders or lambda < borders[lo]))
#pragma unroll 1
    for (int step = 0; step < 3; ++step)
    {
        if (lo0 > 0 && lambda < g.borders[lo0 - 1])
            --lo0;
        else if (lo0 < g.num_borders && !(lambda < g.borders[lo0]))
            ++lo0;
        else
            return lo0;
    }
    int lo = 0, hi = g.num_borders;
    while (lo < hi)
    {
        int mid = (lo + hi) >> 1;
        if (lambda < g.borders[mid])
            hi = mid;
        else
            lo = mid + 1;
    }
    return lo;
}
__device__ __forceinline__ int sk_wlg_bin(const SkDevWlg& g, double lambda)
{
    int lo = 0, hi = g.num_borders;
    while (lo < hi)
    {
        int mid = (lo + hi) >> 1;
        if (lambda < g.borders[mid])
            hi = mid;
        else
            lo = mid + 1;
    }
    return g.ell[lo];
}
// DustMix.cpp:391-425
__device__ __forceinline__ double sk_value_hg(double g, double costheta)
{
    double t = 1. + g * g - 2. * g * costheta;
    return (1. - g) * (1. + g) / sqrt(t * t * t);
}
__device__ __forceinline__ double sk_integral_hg(double g, double cosalpha, double cosbeta)
{
    double ta = sqrt(1. + g * g - 2. * g * cosalpha);
    double tb = sqrt(1. + g * g - 2. * g * cosbeta);
    double f1 = (1. - g) * (1. + g) / g;
    double f2 = (tb - ta) / (tb * ta);
    return f1 * f2;
}
__device__ __noinline__ double sk_mean_hg(double g, double costheta)
{
    const double delta = 4. * M_PI / 180.;
    double theta = acos(costheta);
    double cosalpha = cos(theta - delta);
    double cosbeta = cos(theta + delta);
    if (theta < delta)
        return (sk_integral_hg(g, 1., cosalpha) + sk_integral_hg(g, 1., cosbeta)) / (2. - cosalpha - cosbeta);
    if (theta > M_PI - delta)
        return (sk_integral_hg(g, cosalpha, -1.) + sk_integral_hg(g, cosbeta, -1.)) / (2. + cosalpha + cosbeta);
    return sk_integral_hg(g, cosalpha, cosbeta) / (cosalpha - cosbeta);
}
