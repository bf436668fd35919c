/* sk_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY) for the photon-packet life cycle.
 *
 * A plain-C, single-threaded restatement of the reference algorithm on the hot path named by
 * BASELINE.json (SKIRT 9, MonteCarloSimulation::performLifeCycle and what it calls).  It exists so
 * that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can check / time against it.
 * Nothing under skirt9_b200/ (the product) may import, link or call this file.
 *
 * Parity pinning: the reference has no golden vectors or tests for this path.  The
 * oracle is pinned instead against outputs of the unmodified reference compiled here into
 * oracle/_ref (see oracle/ref.mk, tests/golden/make_golden.py, tests/test_golden_reference.py):
 * agreement is statistical (the reference draws from per-thread MT19937-64, this file and the CUDA
 * engine draw from Philox4x32-10 keyed by history index), within the Monte-Carlo error the reference
 * itself reports through recordStatistics.  Against the CUDA engine the agreement is deterministic
 * (same draws, same arithmetic; differences only from summation order and libm ulps).
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 * The structs and the function set mirror include/sk_engine.h with the prefix sko_.
 */
#include "../include/sk_engine.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------------------ */
/* engine object                                                                                    */
/* ------------------------------------------------------------------------------------------------ */

typedef struct {
    sk_wavelength_grid_t g;
} wlg_t;

typedef struct {
    sk_source_t s;
} src_t;

#define MAX_LEVELS 8
#define NUM_COMP (SK_COMP_PRIMARY_SCATTERED_LEVEL + MAX_LEVELS)

typedef struct {
    sk_instrument_t d;
    double kobs[3];
    double costheta, sintheta, cosphi, sinphi, cosomega, sinomega;
    double xpmin, xpsiz, ypmin, ypsiz, radius2;
    int same_as_preceding;
    int include_sed, include_ifu;
    int record_total_only;
    int nl;
    size_t npix;
    double* sed[NUM_COMP];
    double* ifu[NUM_COMP];
    double* wsed[5];
    double* wifu[5];
    /* per-history statistics accumulator (FluxRecorder::ContributionList, FluxRecorder.hpp:327-339) */
    /* SED: one entry per wavelength bin the history has reached (with kinematics its peel-off packets differ in wavelength),
       FluxRecorder.cpp:962-986 */
    int hist_nsed, hist_capsed;
    int* hist_sed_ell;
    double* hist_sed_w;
    /* the history's contributions per frame pixel: index l + ell*Npix and summed weight, in order of first detection */
    int hist_npix, hist_cappix;
    size_t* hist_lell;
    double* hist_wpix;
} instr_t;

typedef struct {
    int m;
    double ds, s;
    double tau;    /* cumulative extinction optical depth, or scattering optical depth with explicit absorption */
    double tauabs; /* cumulative absorption optical depth with explicit absorption, else zero (SpatialGridPath.hpp:98-99) */
} seg_t;

typedef struct sko_engine {
    sk_config_t cfg;
    uint64_t il_block; /* interleaved sharding (sk_engine_set_history_interleave): block length, parts, this part */
    uint32_t il_parts, il_part;
    /* grid */
    int grid_kind; /* 1 cartesian, 2 octree, 3 voronoi */
    int nx, ny, nz;
    double *xv, *yv, *zv;
    double extent[6];
    double eps;
    int nnodes;
    int32_t* first_child;
    double* node_box;    /* [6*nnodes] */
    int32_t* cell_of_node;
    int32_t* node_of_cell;
    /* voronoi mesh (grid_kind 3) */
    double* vsite;     /* [3*ncells] */
    int64_t* vnbr_off; /* [ncells+1] */
    int32_t* vnbr;     /* neighbour cell indices, -1..-6 for the domain walls */
    int vcells, vnb;   /* number of cells; blocks per axis of the start-cell table */
    int32_t* vblock;   /* [vnb^3] a cell whose site lies in (or near) the block: start of the walk to the nearest site */
    double* vbox;      /* [6*ncells] enclosing boxes of the cells (VoronoiMeshSnapshot::Cell is a Box), or NULL */
    double* vvol;      /* [ncells] cell volumes when the tessellation was built here (sko_build_voronoi), or NULL */
    /* medium: nmed components (Configuration::hasMultipleConstantSectionMedia when > 1); dens[h*ncells + m] */
    int ncells, nmed;
    double *dens, *vol;
    /* kinematics: MediumState::bulkVelocity(m), [3*m + c], or NULL for media at rest; kin = moving media or moving sources */
    double* vel;
    int kin;
    /* dust: one table set per component, [h*nlam + i]; the wavelength grid lam_border is common (DustMix.cpp:52-98) */
    int nlam, nmix;
    double *lam_border, *sig_abs, *sig_sca, *sig_ext, *gpar;
    double mu;
    /* wavelength grids */
    int nwlg;
    wlg_t* wlg;
    int rf_grid;
    int nrf;
    double *rf1, *rf2, *rf2c;
    /* sources */
    int nsrc;
    src_t* src;
    double source_bias;
    double *Lv, *Wv;
    double Ltot;
    uint64_t* Iv;
    double Lpp;
    uint64_t npackets;
    /* instruments */
    int ninstr;
    instr_t* instr;
    int has_medium_emission;
    /* secondary (dust) emission */
    int has_secondary;
    sk_secondary_t sec;
    int sec_nmed;
    double* sec_cmb; /* [nrf] the CMB source term of the energy balance (zeros without CMB heating) */
    double *sec_T, *sec_planckabs, *sec_rfsig, *sec_emsig; /* EquilibriumDustEmissionCalculator tables, one set per dust
                                                              component: [h*nT + i], [h*nrf + ell], [h*sec_nem + i] */
    int sec_nem;             /* N_em + 2 points of DisjointWavelengthGrid::extlambdav() */
    double* sec_lambda;      /* [sec_nem] */
    double *sec_pv, *sec_Pv; /* [ncells][sec_nem] normalised emission spectrum and cdf of every cell */
    double *sec_Lv, *sec_Wv; /* [ncells] DustSecondarySource::_Lv, _Wv */
    uint64_t* sec_Iv;        /* [ncells+1] */
    double sec_Lpp;
    int secondary_ready;
    /* path buffer */
    seg_t* segs;
    int nsegs, capsegs;
    sk_counters_t cnt;
} sko_engine_t;

static char g_err[512] = "";
static void update_kin(sko_engine_t* e);
static int fail(int code, const char* msg)
{
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}
const char* sko_last_error(void)
{
    return g_err;
}
int sko_abi_version(void)
{
    return SK_ABI_VERSION;
}

static double* dupd(const double* p, size_t n)
{
    double* q = (double*)malloc((n ? n : 1) * sizeof(double));
    if (p) memcpy(q, p, n * sizeof(double));
    return q;
}
static int32_t* dupi(const int32_t* p, size_t n)
{
    int32_t* q = (int32_t*)malloc((n ? n : 1) * sizeof(int32_t));
    if (p) memcpy(q, p, n * sizeof(int32_t));
    return q;
}

/* ------------------------------------------------------------------------------------------------ */
/* random numbers: Philox4x32-10 (Salmon et al. 2011), replacing Random.cpp:20-56 (MT19937-64)      */
/*   key = (seed, stream_id); counter = (history lo, history hi, draw index, 0); each draw gives one */
/*   uniform deviate in the open interval (0,1) (Random.cpp:26-27 excludes 0 and 1 as well).        */
/* ------------------------------------------------------------------------------------------------ */

typedef struct {
    uint32_t k0, k1;
    uint32_t c0, c1, draw;
} rng_t;

static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; ++r)
    {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0;
        c[1] = n1;
        c[2] = n2;
        c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
static double u01(uint32_t lo, uint32_t hi)
{
    uint64_t x = ((uint64_t)hi << 32) | lo;
    return ((double)(x >> 12) + 0.5) * (1.0 / 4503599627370496.0); /* 2^-52; result in (0,1) */
}
static void rng_init(rng_t* g, uint32_t seed, uint32_t stream, uint64_t history)
{
    g->k0 = seed;
    g->k1 = stream;
    g->c0 = (uint32_t)history;
    g->c1 = (uint32_t)(history >> 32);
    g->draw = 0;
}
/* Random::uniform, Random.cpp:70-73: the n-th deviate of a history is Philox(counter = (history, n, 0)) */
static double uniform(rng_t* g)
{
    uint32_t c[4] = {g->c0, g->c1, g->draw, 0u};
    philox4x32_10(c, g->k0, g->k1);
    g->draw++;
    return u01(c[0], c[1]);
}
/* Random::exponCutoff, Random.cpp:105-117 */
static double expon_cutoff(rng_t* g, double xmax)
{
    if (xmax == 0.0)
        return 0.0;
    else if (xmax < 1e-10)
        return uniform(g) * xmax;
    double x = -log(1.0 - uniform(g) * (1.0 - exp(-xmax)));
    while (x > xmax)
    {
        x = -log(1.0 - uniform(g) * (1.0 - exp(-xmax)));
    }
    return x;
}
/* Direction::Direction(theta,phi), SKIRT/utils/Direction.cpp:10-35 */
static void direction_from_angles(double theta, double phi, double k[3])
{
    const double eps = 1e-8;
    if (theta <= eps)
    {
        k[0] = 0;
        k[1] = 0;
        k[2] = 1;
    }
    else if (theta >= M_PI - eps)
    {
        k[0] = 0;
        k[1] = 0;
        k[2] = -1;
    }
    else
    {
        double sintheta = sin(theta);
        k[0] = sintheta * cos(phi);
        k[1] = sintheta * sin(phi);
        k[2] = cos(theta);
    }
}
/* Random::direction(), Random.cpp:121-126 */
static void random_direction(rng_t* g, double k[3])
{
    double theta = acos(2.0 * uniform(g) - 1.0);
    double phi = 2.0 * M_PI * uniform(g);
    direction_from_angles(theta, phi, k);
}
/* Random::direction(bfk, costheta), Random.cpp:130-164 */
static void random_direction_about(rng_t* g, const double k[3], double costheta, double knew[3])
{
    double phi = 2.0 * M_PI * uniform(g);
    double cosphi = cos(phi);
    double sinphi = sin(phi);
    double sintheta = sqrt(fabs((1.0 - costheta) * (1.0 + costheta)));
    double kx = k[0], ky = k[1], kz = k[2];
    if (kz > 0.99999)
    {
        knew[0] = cosphi * sintheta;
        knew[1] = sinphi * sintheta;
        knew[2] = costheta;
    }
    else if (kz < -0.99999)
    {
        knew[0] = cosphi * sintheta;
        knew[1] = sinphi * sintheta;
        knew[2] = -costheta;
    }
    else
    {
        double root = sqrt((1.0 - kz) * (1.0 + kz));
        knew[0] = sintheta / root * (-kx * kz * cosphi + ky * sinphi) + kx * costheta;
        knew[1] = -sintheta / root * (ky * kz * cosphi + kx * sinphi) + ky * costheta;
        knew[2] = root * sintheta * cosphi + kz * costheta;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* numerical helpers                                                                                */
/* ------------------------------------------------------------------------------------------------ */

/* NR::locateBasicImpl / locateClip / locateFail, SKIRT/utils/NR.hpp:130-191 */
static int locate_basic(const double* xv, double x, int n)
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
static int locate_clip(const double* xv, int n, double x)
{
    if (x < xv[0]) return 0;
    return locate_basic(xv, x, n - 1);
}
static int locate_fail(const double* xv, int n, double x)
{
    if (x > xv[n - 1]) return -1;
    return locate_basic(xv, x, n - 1);
}
/* NR::interpolateLinLin, NR.hpp:328-331 */
static double interp_linlin(double x, double x1, double x2, double f1, double f2)
{
    return f1 + ((x - x1) / (x2 - x1)) * (f2 - f1);
}
/* NR::interpolateLogLog, NR.hpp:349-358 */
static double interp_loglog(double x, double x1, double x2, double f1, double f2)
{
    if (f1 <= 0 || f2 <= 0)
    {
        if (x == x1) return f1;
        if (x == x2) return f2;
        return 0;
    }
    return f1 * exp(log(x / x1) / log(x2 / x1) * (log(f2 / f1)));
}
/* SpecialFunctions::gexp, SKIRT/utils/SpecialFunctions.cpp:822-836 */
static double gexp(double p, double x)
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
/* SpecialFunctions::gln, SpecialFunctions.cpp:798-811 */
static double gln(double p, double x)
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

/* SpecialFunctions::lnmean(x1,x2,lnx1,lnx2), SpecialFunctions.cpp:860-880 */
static double lnmean4(double x1, double x2, double lnx1, double lnx2)
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
    {
        return (x2 - x1) / (lnx2 - lnx1);
    }
}
/* SpecialFunctions::LambertW1, SpecialFunctions.cpp:578-627 */
static double lambert_w1(double z)
{
    const double eps = 1.0e-12;
    const double em1 = 0.3678794411714423215955237701614608;
    static const double c[12] = {-1.0,
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
    return w; /* the reference throws here; unreachable for valid arguments */
}
/* PlanckFunction::value, SKIRT/utils/PlanckFunction.cpp:24-27 with Constants h, c, k */
static double planck(double lambda, double T)
{
    const double h = 6.62606957e-34, c = 2.99792458e8, k = 1.3806488e-23;
    double f1 = h * c / (k * T);
    double f2 = 2.0 * h * c * c;
    const double l2 = lambda * lambda; /* lambda^5 by multiplication, as in the CUDA engine (sk_device.cuh sk_planck) */
    return f2 / (l2 * l2 * lambda) / (exp(f1 / lambda) - 1.0);
}

/* ------------------------------------------------------------------------------------------------ */
/* setters                                                                                          */
/* ------------------------------------------------------------------------------------------------ */

int sko_create(const sk_config_t* config, sko_engine_t** out)
{
    if (!config || !out) return fail(SK_ERR_INVALID, "null argument");
    sko_engine_t* e = (sko_engine_t*)calloc(1, sizeof *e);
    e->cfg = *config;
    e->rf_grid = -1;
    e->capsegs = 1000; /* SpatialGridPath.cpp:14 INITIAL_CAPACITY */
    e->segs = (seg_t*)malloc(e->capsegs * sizeof(seg_t));
    *out = e;
    return SK_OK;
}

static void free_instruments(sko_engine_t* e)
{
    for (int i = 0; i < e->ninstr; ++i)
    {
        for (int c = 0; c < NUM_COMP; ++c)
        {
            free(e->instr[i].sed[c]);
            free(e->instr[i].ifu[c]);
        }
        for (int k = 0; k < 5; ++k) free(e->instr[i].wsed[k]);
        for (int k = 0; k < 5; ++k) free(e->instr[i].wifu[k]);
        free(e->instr[i].hist_lell);
        free(e->instr[i].hist_wpix);
        free(e->instr[i].hist_sed_ell);
        free(e->instr[i].hist_sed_w);
    }
    free(e->instr);
    e->instr = NULL;
    e->ninstr = 0;
}
static void free_sources(sko_engine_t* e)
{
    for (int i = 0; i < e->nsrc; ++i)
    {
        sk_source_t* s = &e->src[i].s;
        free((void*)s->geom_table_x);
        free((void*)s->geom_table_P);
        free((void*)s->sed_lambda);
        free((void*)s->sed_p);
        free((void*)s->sed_P);
        free((void*)s->oligo_lambda);
    }
    free(e->src);
    free(e->Lv);
    free(e->Wv);
    free(e->Iv);
    e->src = NULL;
    e->Lv = e->Wv = NULL;
    e->Iv = NULL;
    e->nsrc = 0;
}
static void free_wlg(sko_engine_t* e)
{
    for (int i = 0; i < e->nwlg; ++i)
    {
        free((void*)e->wlg[i].g.borders);
        free((void*)e->wlg[i].g.ell);
        free((void*)e->wlg[i].g.lambda);
        free((void*)e->wlg[i].g.dlambda);
    }
    free(e->wlg);
    e->wlg = NULL;
    e->nwlg = 0;
}
static void free_grid(sko_engine_t* e)
{
    free(e->xv);
    free(e->yv);
    free(e->zv);
    free(e->first_child);
    free(e->node_box);
    free(e->cell_of_node);
    free(e->node_of_cell);
    free(e->vsite);
    free(e->vnbr_off);
    free(e->vnbr);
    free(e->vblock);
    free(e->vbox);
    free(e->vvol);
    e->vbox = NULL;
    e->vvol = NULL;
    e->vsite = NULL;
    e->vnbr_off = NULL;
    e->vnbr = e->vblock = NULL;
    e->xv = e->yv = e->zv = e->node_box = NULL;
    e->first_child = e->cell_of_node = e->node_of_cell = NULL;
    e->grid_kind = 0;
}

static void free_secondary(sko_engine_t* e);
void sko_destroy(sko_engine_t* e)
{
    if (!e) return;
    free_instruments(e);
    free_sources(e);
    free_secondary(e);
    free_wlg(e);
    free_grid(e);
    free(e->dens);
    free(e->vol);
    free(e->vel);
    free(e->lam_border);
    free(e->sig_abs);
    free(e->sig_sca);
    free(e->sig_ext);
    free(e->gpar);
    free(e->rf1);
    free(e->rf2);
    free(e->rf2c);
    free(e->segs);
    free(e);
}

/* CartesianSpatialGrid setup, CartesianSpatialGrid.cpp:22-60 */
int sko_set_grid_cartesian(sko_engine_t* e, int32_t nx, int32_t ny, int32_t nz, const double* xv, const double* yv,
                           const double* zv)
{
    if (!e || nx < 1 || ny < 1 || nz < 1 || !xv || !yv || !zv) return fail(SK_ERR_INVALID, "bad cartesian grid");
    free_grid(e);
    e->grid_kind = 1;
    e->nx = nx;
    e->ny = ny;
    e->nz = nz;
    e->xv = dupd(xv, nx + 1);
    e->yv = dupd(yv, ny + 1);
    e->zv = dupd(zv, nz + 1);
    e->extent[0] = xv[0];
    e->extent[1] = yv[0];
    e->extent[2] = zv[0];
    e->extent[3] = xv[nx];
    e->extent[4] = yv[ny];
    e->extent[5] = zv[nz];
    double dx = e->extent[3] - e->extent[0], dy = e->extent[4] - e->extent[1], dz = e->extent[5] - e->extent[2];
    e->eps = 1e-12 * sqrt(dx * dx + dy * dy + dz * dz); /* CartesianSpatialGrid.cpp:102, Box.hpp:125-129 */
    return SK_OK;
}

/* TreeSpatialGrid::setupSelfAfter, TreeSpatialGrid.cpp:23-49; OctTreeNode::createChildren, OctTreeNode.cpp:22-35 */
int sko_set_grid_octree(sko_engine_t* e, const double extent[6], int32_t num_nodes, const int32_t* first_child)
{
    if (!e || !extent || num_nodes < 1 || !first_child) return fail(SK_ERR_INVALID, "bad octree");
    free_grid(e);
    e->grid_kind = 2;
    memcpy(e->extent, extent, 6 * sizeof(double));
    double dx = extent[3] - extent[0], dy = extent[4] - extent[1], dz = extent[5] - extent[2];
    e->eps = 1e-12 * sqrt(dx * dx + dy * dy + dz * dz); /* TreeSpatialGrid.cpp:28 */
    e->nnodes = num_nodes;
    e->first_child = dupi(first_child, num_nodes);
    e->node_box = (double*)malloc(6 * (size_t)num_nodes * sizeof(double));
    e->cell_of_node = (int32_t*)malloc((size_t)num_nodes * sizeof(int32_t));
    e->node_of_cell = (int32_t*)malloc((size_t)num_nodes * sizeof(int32_t));
    char* seen = (char*)calloc(num_nodes, 1);
    memcpy(e->node_box, extent, 6 * sizeof(double));
    seen[0] = 1;
    int m = 0;
    for (int l = 0; l < num_nodes; ++l)
    {
        if (!seen[l])
        {
            free(seen);
            return fail(SK_ERR_INVALID, "octree node list is not parent-before-child");
        }
        int fc = first_child[l];
        if (fc < 0)
        {
            e->cell_of_node[l] = m;
            e->node_of_cell[m] = l;
            m++;
            continue;
        }
        if (fc <= l || fc + 8 > num_nodes)
        {
            free(seen);
            return fail(SK_ERR_INVALID, "octree child index out of range");
        }
        e->cell_of_node[l] = -1;
        const double* b = e->node_box + 6 * (size_t)l;
        double xmin = b[0], ymin = b[1], zmin = b[2], xmax = b[3], ymax = b[4], zmax = b[5];
        double cx = 0.5 * (xmin + xmax), cy = 0.5 * (ymin + ymax), cz = 0.5 * (zmin + zmax); /* Box.hpp:135 */
        for (int c = 0; c < 8; ++c)
        {
            double* cb = e->node_box + 6 * (size_t)(fc + c);
            cb[0] = (c & 1) ? cx : xmin;
            cb[3] = (c & 1) ? xmax : cx;
            cb[1] = (c & 2) ? cy : ymin;
            cb[4] = (c & 2) ? ymax : cy;
            cb[2] = (c & 4) ? cz : zmin;
            cb[5] = (c & 4) ? zmax : cz;
            seen[fc + c] = 1;
        }
    }
    free(seen);
    e->ncells = 0; /* medium must be (re)set */
    e->nx = m;     /* remember leaf count for validation */
    return SK_OK;
}

/* The start-cell table used by the nearest-site search: the domain is cut into nb^3 blocks, nb = clamp(cbrt(N),3,250) as in
 * VoronoiMeshSnapshot.cpp:544, and every block remembers one cell whose site lies inside it (blocks without a site inherit
 * the previous block's cell).  Identical in the CUDA engine (engine.cu). */
static void voronoi_build_blocks(sko_engine_t* e)
{
    int nb = (int)cbrt((double)e->vcells);
    if (nb < 3) nb = 3;
    if (nb > 250) nb = 250;
    e->vnb = nb;
    size_t n3 = (size_t)nb * nb * nb;
    e->vblock = (int32_t*)malloc(n3 * sizeof(int32_t));
    for (size_t b = 0; b < n3; ++b) e->vblock[b] = -1;
    const double* x = e->extent;
    for (int m = 0; m < e->vcells; ++m)
    {
        int i = (int)((e->vsite[3 * m] - x[0]) / (x[3] - x[0]) * nb);
        int j = (int)((e->vsite[3 * m + 1] - x[1]) / (x[4] - x[1]) * nb);
        int k = (int)((e->vsite[3 * m + 2] - x[2]) / (x[5] - x[2]) * nb);
        i = i < 0 ? 0 : i >= nb ? nb - 1 : i;
        j = j < 0 ? 0 : j >= nb ? nb - 1 : j;
        k = k < 0 ? 0 : k >= nb ? nb - 1 : k;
        size_t b = ((size_t)i * nb + j) * nb + k;
        if (e->vblock[b] < 0) e->vblock[b] = m;
    }
    int32_t last = 0;
    for (size_t b = 0; b < n3; ++b)
    {
        if (e->vblock[b] < 0)
            e->vblock[b] = last;
        else
            last = e->vblock[b];
    }
}

int sko_set_grid_voronoi(sko_engine_t* e, const double extent[6], int32_t num_cells, const double* sites,
                         const int64_t* nbr_offset, const int32_t* nbr_index)
{
    if (!e || !extent || num_cells < 1 || !sites || !nbr_offset || !nbr_index) return fail(SK_ERR_INVALID, "bad voronoi mesh");
    if (nbr_offset[0] != 0) return fail(SK_ERR_INVALID, "neighbour offsets must start at zero");
    for (int m = 0; m < num_cells; ++m)
    {
        if (nbr_offset[m + 1] < nbr_offset[m]) return fail(SK_ERR_INVALID, "neighbour offsets must ascend");
        for (int64_t i = nbr_offset[m]; i < nbr_offset[m + 1]; ++i)
            if (nbr_index[i] < -6 || nbr_index[i] >= num_cells) return fail(SK_ERR_INVALID, "neighbour index out of range");
    }
    free_grid(e);
    e->grid_kind = 3;
    e->vcells = num_cells;
    memcpy(e->extent, extent, 6 * sizeof(double));
    double dx = extent[3] - extent[0], dy = extent[4] - extent[1], dz = extent[5] - extent[2];
    e->eps = 1e-12 * sqrt(dx * dx + dy * dy + dz * dz); /* VoronoiMeshSnapshot.cpp:396 */
    e->vsite = dupd(sites, 3 * (size_t)num_cells);
    e->vnbr_off = (int64_t*)malloc(((size_t)num_cells + 1) * sizeof(int64_t));
    memcpy(e->vnbr_off, nbr_offset, ((size_t)num_cells + 1) * sizeof(int64_t));
    e->vnbr = dupi(nbr_index, (size_t)nbr_offset[num_cells]);
    voronoi_build_blocks(e);
    return SK_OK;
}

static int grid_num_cells(const sko_engine_t* e)
{
    if (e->grid_kind == 3) return e->vcells;
    if (e->grid_kind == 1) return e->nx * e->ny * e->nz;
    if (e->grid_kind == 2) return e->nx;
    return 0;
}

/* The enclosing boxes of the Voronoi cells: VoronoiMeshSnapshot::Cell::init, VoronoiMeshSnapshot.cpp:104-135 */
int sko_set_voronoi_extents(sko_engine_t* e, int32_t num_cells, const double* boxes)
{
    if (!e || !boxes) return fail(SK_ERR_INVALID, "null argument");
    if (e->grid_kind != 3) return fail(SK_ERR_STATE, "set the Voronoi grid before its cell extents");
    if (num_cells != e->vcells) return fail(SK_ERR_INVALID, "extents do not match the grid");
    for (int m = 0; m < num_cells; ++m)
        for (int a = 0; a < 3; ++a)
            if (!(boxes[6 * (size_t)m + a] <= boxes[6 * (size_t)m + a + 3])) return fail(SK_ERR_INVALID, "empty cell extent");
    free(e->vbox);
    e->vbox = dupd(boxes, 6 * (size_t)num_cells);
    return SK_OK;
}

/* ------------------------------------------------------------------------------------------------ */
/* Voronoi tessellation: VoronoiMeshSnapshot::buildMesh (VoronoiMeshSnapshot.cpp:491-730), which the  */
/* reference delegates to the vendored voro++ (SKIRT/voro: container::compute_cell for every site, then                    */
/* voronoicell_neighbor::neighbors / volume / vertices).  voro++'s published algorithm (Rycroft 2009) starts from the       */
/* domain box around the site and cuts it with the bisecting planes towards the other sites in order of increasing        */
/* distance, visiting the blocks of a uniform search grid outwards, until no unvisited site can cut the cell any more     */
/* (every unvisited site is farther than twice the cell's largest vertex distance).  Restated here with the cell held as  */
/* a list of vertices, each the intersection of three planes (the cell is simple for sites in general position), in       */
/* coordinates relative to the site.  Output per cell: the neighbours whose plane carries a face (in order of cutting),    */
/* then the domain walls that do (-1..-6 = xmin,xmax,ymin,ymax,zmin,zmax), the volume and the enclosing box.               */
/* ------------------------------------------------------------------------------------------------ */
#define VC_MAXP 96  /* planes kept per cell (6 walls + the bisectors that have cut it so far) */
#define VC_MAXT 192 /* vertices */
typedef struct {
    int np, nt;
    double pn[VC_MAXP][3], pd[VC_MAXP]; /* plane j: pn.x <= pd, x relative to the site */
    int pid[VC_MAXP];                   /* neighbour index, or -1..-6 */
    unsigned char ta[VC_MAXT], tb[VC_MAXT], tc[VC_MAXT];
    double vx[VC_MAXT], vy[VC_MAXT], vz[VC_MAXT];
    double rmax2; /* largest squared vertex distance */
} vcell_t;

/* the point where three planes meet (Cramer's rule); returns 0 when they do not meet in a point */
static int vc_vertex(const vcell_t* c, int a, int b, int d, double* x, double* y, double* z)
{
    const double* A = c->pn[a];
    const double* B = c->pn[b];
    const double* C = c->pn[d];
    const double bcx = B[1] * C[2] - B[2] * C[1], bcy = B[2] * C[0] - B[0] * C[2], bcz = B[0] * C[1] - B[1] * C[0];
    const double det = A[0] * bcx + A[1] * bcy + A[2] * bcz;
    if (det == 0.) return 0;
    const double cax = C[1] * A[2] - C[2] * A[1], cay = C[2] * A[0] - C[0] * A[2], caz = C[0] * A[1] - C[1] * A[0];
    const double abx = A[1] * B[2] - A[2] * B[1], aby = A[2] * B[0] - A[0] * B[2], abz = A[0] * B[1] - A[1] * B[0];
    const double da = c->pd[a], db = c->pd[b], dd = c->pd[d];
    *x = (da * bcx + db * cax + dd * abx) / det;
    *y = (da * bcy + db * cay + dd * aby) / det;
    *z = (da * bcz + db * caz + dd * abz) / det;
    return 1;
}
static void vc_rmax(vcell_t* c)
{
    double r = 0.;
    for (int t = 0; t < c->nt; ++t)
    {
        double q = c->vx[t] * c->vx[t] + c->vy[t] * c->vy[t] + c->vz[t] * c->vz[t];
        if (q > r) r = q;
    }
    c->rmax2 = r;
}
static void vc_init(vcell_t* c, const double extent[6], const double p[3])
{
    c->np = 6;
    c->nt = 0;
    for (int w = 0; w < 6; ++w)
    {
        const int axis = w >> 1, upper = w & 1;
        c->pn[w][0] = c->pn[w][1] = c->pn[w][2] = 0.;
        c->pn[w][axis] = upper ? 1. : -1.;
        c->pd[w] = upper ? extent[axis + 3] - p[axis] : -(extent[axis] - p[axis]);
        c->pid[w] = -(w + 1);
    }
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            for (int k = 0; k < 2; ++k)
            {
                const int t = c->nt++;
                c->ta[t] = (unsigned char)i;
                c->tb[t] = (unsigned char)(2 + j);
                c->tc[t] = (unsigned char)(4 + k);
                vc_vertex(c, i, 2 + j, 4 + k, &c->vx[t], &c->vy[t], &c->vz[t]);
            }
    vc_rmax(c);
}
/* cuts the cell with the plane n.x <= d of neighbour `id`; returns 0 on success, a negative code when the cell is not simple
 * or outgrows the buffers */
static int vc_clip(vcell_t* c, double nx, double ny, double nz, double d, int id)
{
    unsigned char out[VC_MAXT];
    int nout = 0;
    for (int t = 0; t < c->nt; ++t)
    {
        out[t] = (nx * c->vx[t] + ny * c->vy[t] + nz * c->vz[t] - d) > 0.;
        nout += out[t];
    }
    if (!nout) return 0;             /* the plane does not reach the cell */
    if (nout == c->nt) return -1;    /* cannot happen for a bisector: the site itself is inside */
    if (c->np >= VC_MAXP)
    {
        /* drop the planes that no longer carry a vertex (cut away again by nearer neighbours) and renumber */
        unsigned char map[VC_MAXP], keep[VC_MAXP];
        for (int j = 0; j < c->np; ++j) keep[j] = j < 6;
        for (int t = 0; t < c->nt; ++t) keep[c->ta[t]] = keep[c->tb[t]] = keep[c->tc[t]] = 1;
        int np = 0;
        for (int j = 0; j < c->np; ++j)
        {
            map[j] = (unsigned char)np;
            if (!keep[j]) continue;
            c->pn[np][0] = c->pn[j][0];
            c->pn[np][1] = c->pn[j][1];
            c->pn[np][2] = c->pn[j][2];
            c->pd[np] = c->pd[j];
            c->pid[np] = c->pid[j];
            np++;
        }
        c->np = np;
        for (int t = 0; t < c->nt; ++t)
        {
            c->ta[t] = map[c->ta[t]];
            c->tb[t] = map[c->tb[t]];
            c->tc[t] = map[c->tc[t]];
        }
        if (c->np >= VC_MAXP) return -2;
    }
    const int P = c->np++;
    c->pn[P][0] = nx;
    c->pn[P][1] = ny;
    c->pn[P][2] = nz;
    c->pd[P] = d;
    c->pid[P] = id;
    /* the edges (pairs of planes) of the removed vertices that lead to a kept vertex: those that occur once among them */
    unsigned char ea[3 * VC_MAXT], eb[3 * VC_MAXT], eo[3 * VC_MAXT];
    int ne = 0;
    for (int t = 0; t < c->nt; ++t)
    {
        if (!out[t]) continue;
        const unsigned char pa[3] = {c->ta[t], c->ta[t], c->tb[t]}, pb[3] = {c->tb[t], c->tc[t], c->tc[t]};
        for (int k = 0; k < 3; ++k)
        {
            int found = -1;
            for (int q = 0; q < ne; ++q)
                if (ea[q] == pa[k] && eb[q] == pb[k]) found = q;
            if (found >= 0)
                eo[found]++;
            else
            {
                ea[ne] = pa[k];
                eb[ne] = pb[k];
                eo[ne] = 1;
                ne++;
            }
        }
    }
    /* compact the kept vertices, then one new vertex per boundary edge */
    int nt = 0;
    for (int t = 0; t < c->nt; ++t)
        if (!out[t])
        {
            c->ta[nt] = c->ta[t];
            c->tb[nt] = c->tb[t];
            c->tc[nt] = c->tc[t];
            c->vx[nt] = c->vx[t];
            c->vy[nt] = c->vy[t];
            c->vz[nt] = c->vz[t];
            nt++;
        }
    for (int q = 0; q < ne; ++q)
    {
        if (eo[q] > 2) return -3; /* more than three planes through a vertex: not a simple cell */
        if (eo[q] != 1) continue;
        if (nt >= VC_MAXT) return -2;
        c->ta[nt] = ea[q];
        c->tb[nt] = eb[q];
        c->tc[nt] = (unsigned char)P;
        if (!vc_vertex(c, ea[q], eb[q], P, &c->vx[nt], &c->vy[nt], &c->vz[nt])) return -3;
        nt++;
    }
    c->nt = nt;
    vc_rmax(c);
    return 0;
}
/* volume and enclosing box (relative to the site) from the faces: the vertices of a face are ordered by walking from one to
 * the next across the plane they share besides the face's own; the volume is the sum of the pyramids site-face */
static int vc_measure(const vcell_t* c, double* volume, double box[6], unsigned char used[VC_MAXP])
{
    for (int j = 0; j < c->np; ++j) used[j] = 0;
    box[0] = box[1] = box[2] = DBL_MAX;
    box[3] = box[4] = box[5] = -DBL_MAX;
    for (int t = 0; t < c->nt; ++t)
    {
        used[c->ta[t]] = used[c->tb[t]] = used[c->tc[t]] = 1;
        if (c->vx[t] < box[0]) box[0] = c->vx[t];
        if (c->vy[t] < box[1]) box[1] = c->vy[t];
        if (c->vz[t] < box[2]) box[2] = c->vz[t];
        if (c->vx[t] > box[3]) box[3] = c->vx[t];
        if (c->vy[t] > box[4]) box[4] = c->vy[t];
        if (c->vz[t] > box[5]) box[5] = c->vz[t];
    }
    double V = 0.;
    for (int f = 0; f < c->np; ++f)
    {
        if (!used[f]) continue;
        int t0 = -1;
        for (int t = 0; t < c->nt && t0 < 0; ++t)
            if (c->ta[t] == f || c->tb[t] == f || c->tc[t] == f) t0 = t;
        /* the two other planes of the first vertex; leave through the second */
        int o1 = c->ta[t0] == f ? c->tb[t0] : c->ta[t0];
        int via = c->tc[t0] == f ? c->tb[t0] : c->tc[t0];
        if (via == o1) via = c->tc[t0];
        int cur = t0, steps = 0;
        double sum = 0.;
        double px = 0., py = 0., pz = 0.; /* previous vertex of the fan */
        while (1)
        {
            int next = -1;
            for (int t = 0; t < c->nt; ++t)
            {
                if (t == cur) continue;
                const int a = c->ta[t], b = c->tb[t], d = c->tc[t];
                if ((a == f || b == f || d == f) && (a == via || b == via || d == via)) next = t;
            }
            if (next < 0 || ++steps > c->nt) return -3;
            const int a = c->ta[next], b = c->tb[next], d = c->tc[next];
            const int other = (a != f && a != via) ? a : (b != f && b != via) ? b : d;
            if (next == t0) break;
            if (steps >= 2)
            {
                /* triangle (t0, previous, next) of the fan: 6 x signed volume of the pyramid with the site */
                const double ax = c->vx[t0], ay = c->vy[t0], az = c->vz[t0];
                const double bx = c->vx[next], by = c->vy[next], bz = c->vz[next];
                sum += ax * (py * bz - pz * by) + ay * (pz * bx - px * bz) + az * (px * by - py * bx);
            }
            px = c->vx[next];
            py = c->vy[next];
            pz = c->vz[next];
            via = other;
            cur = next;
        }
        V += fabs(sum);
    }
    *volume = V / 6.;
    return 0;
}

int sko_build_voronoi(sko_engine_t* e, const double extent[6], int32_t num_sites, const double* sites, uint64_t* num_entries)
{
    if (!e || !extent || num_sites < 1 || !sites) return fail(SK_ERR_INVALID, "bad voronoi sites");
    const int n = num_sites;
    for (int m = 0; m < n; ++m)
        for (int a = 0; a < 3; ++a)
            if (!(sites[3 * (size_t)m + a] > extent[a] && sites[3 * (size_t)m + a] < extent[a + 3]))
                return fail(SK_ERR_INVALID, "site outside the domain");
    /* search grid: cubic blocks holding two sites on average; sites of a block in ascending index */
    const double wx = extent[3] - extent[0], wy = extent[4] - extent[1], wz = extent[5] - extent[2];
    const double w = cbrt(2. * wx * wy * wz / n);
    int gx = (int)ceil(wx / w), gy = (int)ceil(wy / w), gz = (int)ceil(wz / w);
    gx = gx < 1 ? 1 : gx;
    gy = gy < 1 ? 1 : gy;
    gz = gz < 1 ? 1 : gz;
    const size_t ng = (size_t)gx * gy * gz;
    int32_t* start = (int32_t*)calloc(ng + 1, sizeof(int32_t));
    int32_t* blk = (int32_t*)malloc((size_t)n * sizeof(int32_t));
    int32_t* order = (int32_t*)malloc((size_t)n * sizeof(int32_t));
    for (int m = 0; m < n; ++m)
    {
        int i = (int)((sites[3 * (size_t)m] - extent[0]) / w), j = (int)((sites[3 * (size_t)m + 1] - extent[1]) / w),
            k = (int)((sites[3 * (size_t)m + 2] - extent[2]) / w);
        i = i >= gx ? gx - 1 : i;
        j = j >= gy ? gy - 1 : j;
        k = k >= gz ? gz - 1 : k;
        blk[m] = (int32_t)(((size_t)i * gy + j) * gz + k);
        start[blk[m] + 1]++;
    }
    for (size_t b = 0; b < ng; ++b) start[b + 1] += start[b];
    {
        int32_t* fill = (int32_t*)malloc(ng * sizeof(int32_t));
        memcpy(fill, start, ng * sizeof(int32_t));
        for (int m = 0; m < n; ++m) order[fill[blk[m]]++] = m;
        free(fill);
    }
    int64_t* off = (int64_t*)malloc(((size_t)n + 1) * sizeof(int64_t));
    int32_t* idx = (int32_t*)malloc((size_t)n * VC_MAXP * sizeof(int32_t));
    double* vol = (double*)malloc((size_t)n * sizeof(double));
    double* box = (double*)malloc(6 * (size_t)n * sizeof(double));
    int rc = 0;
    off[0] = 0;
    vcell_t c;
    for (int m = 0; m < n && !rc; ++m)
    {
        const double* p = sites + 3 * (size_t)m;
        vc_init(&c, extent, p);
        const int bi = blk[m] / (gy * gz), bj = (blk[m] / gz) % gy, bk = blk[m] % gz;
        const int smax = (gx > gy ? (gx > gz ? gx : gz) : (gy > gz ? gy : gz));
        for (int s = 0; s <= smax && !rc; ++s)
        {
            /* every site in a block at Chebyshev distance s or more is at least (s-1) w away */
            if (s >= 2 && (double)(s - 1) * w * (double)(s - 1) * w >= 4. * c.rmax2) break;
            for (int i = bi - s; i <= bi + s && !rc; ++i)
            {
                if (i < 0 || i >= gx) continue;
                for (int j = bj - s; j <= bj + s && !rc; ++j)
                {
                    if (j < 0 || j >= gy) continue;
                    const int shell = (i == bi - s || i == bi + s || j == bj - s || j == bj + s);
                    for (int k = bk - s; k <= bk + s && !rc; k += (shell || s == 0) ? 1 : 2 * s)
                    {
                        if (k < 0 || k >= gz) continue;
                        const size_t b = ((size_t)i * gy + j) * gz + k;
                        for (int32_t q = start[b]; q < start[b + 1] && !rc; ++q)
                        {
                            const int mi = order[q];
                            if (mi == m) continue;
                            const double nx = sites[3 * (size_t)mi] - p[0], ny = sites[3 * (size_t)mi + 1] - p[1],
                                         nz = sites[3 * (size_t)mi + 2] - p[2];
                            const double n2 = nx * nx + ny * ny + nz * nz;
                            if (n2 >= 4. * c.rmax2) continue; /* its bisector lies beyond the farthest vertex */
                            if (n2 == 0.) rc = -4;
                            else rc = vc_clip(&c, nx, ny, nz, 0.5 * n2, mi);
                        }
                    }
                }
            }
        }
        if (rc) break;
        unsigned char used[VC_MAXP];
        double b6[6];
        rc = vc_measure(&c, &vol[m], b6, used);
        if (rc) break;
        int64_t o = off[m];
        for (int j = 6; j < c.np; ++j)
            if (used[j]) idx[o++] = c.pid[j];
        for (int j = 0; j < 6; ++j)
            if (used[j]) idx[o++] = c.pid[j];
        off[m + 1] = o;
        for (int a = 0; a < 3; ++a)
        {
            box[6 * (size_t)m + a] = p[a] + b6[a];
            box[6 * (size_t)m + a + 3] = p[a] + b6[a + 3];
        }
    }
    free(start);
    free(blk);
    free(order);
    int out = SK_OK;
    if (rc)
        out = fail(SK_ERR_UNSUPPORTED, rc == -2   ? "Voronoi cell with more faces than the builder holds"
                                       : rc == -4 ? "coinciding Voronoi sites"
                                                  : "Voronoi sites in degenerate position (a cell that is not simple)");
    else
    {
        out = sko_set_grid_voronoi(e, extent, n, sites, off, idx);
        if (!out) out = sko_set_voronoi_extents(e, n, box);
        if (!out)
        {
            free(e->vvol);
            e->vvol = vol;
            vol = NULL;
            if (num_entries) *num_entries = (uint64_t)off[n];
        }
    }
    free(off);
    free(idx);
    free(vol);
    free(box);
    return out;
}
/* the tessellation the engine holds: neighbour lists, cell volumes and enclosing boxes (any pointer may be NULL) */
int sko_read_voronoi(sko_engine_t* e, int64_t* nbr_offset, int32_t* nbr_index, double* volume, double* boxes)
{
    if (!e || e->grid_kind != 3) return fail(SK_ERR_STATE, "the engine holds no Voronoi grid");
    if (nbr_offset) memcpy(nbr_offset, e->vnbr_off, ((size_t)e->vcells + 1) * sizeof(int64_t));
    if (nbr_index) memcpy(nbr_index, e->vnbr, (size_t)e->vnbr_off[e->vcells] * sizeof(int32_t));
    if (volume)
    {
        if (!e->vvol) return fail(SK_ERR_STATE, "the grid was not built by build_voronoi: no volumes");
        memcpy(volume, e->vvol, (size_t)e->vcells * sizeof(double));
    }
    if (boxes)
    {
        if (!e->vbox) return fail(SK_ERR_STATE, "no cell extents");
        memcpy(boxes, e->vbox, 6 * (size_t)e->vcells * sizeof(double));
    }
    return SK_OK;
}

int sko_set_media(sko_engine_t* e, int32_t num_cells, int32_t num_media, const double* number_density, const double* volume)
{
    if (!e || !number_density) return fail(SK_ERR_INVALID, "null argument");
    if (num_media < 1 || num_media > SK_MAX_MEDIA) return fail(SK_ERR_UNSUPPORTED, "number of medium components");
    if (!e->grid_kind) return fail(SK_ERR_STATE, "set the grid before the medium");
    if (num_cells != grid_num_cells(e)) return fail(SK_ERR_INVALID, "medium size does not match the grid");
    free(e->dens);
    free(e->vol);
    free(e->vel); /* a new medium state is at rest until sko_set_velocities says otherwise */
    e->vel = NULL;
    update_kin(e);
    e->ncells = num_cells;
    e->nmed = num_media;
    e->dens = dupd(number_density, (size_t)num_media * num_cells);
    e->vol = volume ? dupd(volume, num_cells) : NULL;
    return SK_OK;
}
int sko_set_medium(sko_engine_t* e, int32_t num_cells, const double* number_density, const double* volume)
{
    return sko_set_media(e, num_cells, 1, number_density, volume);
}

/* ------------------------------------------------------------------------------------------------ */
/* setup: octree construction by the density policy and medium-state sampling  */
/* ------------------------------------------------------------------------------------------------ */

/* x^e for the even exponents 2N of SpiralStructureGeometryDecorator::perturbation: x*x is the correctly rounded
 * pow(x, 2); same rule in the CUDA engine (sk_device.cuh sk_pow_even) */
static double pow_even(double x, double e)
{
    return e == 2.0 ? x * x : pow(x, e);
}

/* Geometry::density(Position): ShellGeometry.cpp:30-36 (through SpheGeometry), ExpDiskGeometry.cpp:32-42 and
 * RingGeometry.cpp:39-43 (through AxGeometry), SpiralStructureGeometryDecorator.cpp:24-29,71-75 */
static double geom_density(const sk_density_geometry_t* g, double x, double y, double z)
{
    const double* p = g->p;
    switch (g->geometry)
    {
        case SK_GEOM_SHELL:
        {
            double r = sqrt(x * x + y * y + z * z);
            if (r < p[0] || r > p[1]) return 0.0;
            return p[3] * pow(r, -p[2]);
        }
        case SK_GEOM_EXPDISK:
        case SK_GEOM_SPIRAL_EXPDISK:
        {
            double R = sqrt(x * x + y * y);
            double absz = fabs(z);
            double rho;
            if (p[3] > 0.0 && R > p[3])
                rho = 0.0;
            else if (p[4] > 0.0 && absz > p[4])
                rho = 0.0;
            else if (R < p[2])
                rho = 0.0;
            else
                rho = p[5] * exp(-R / p[0]) * exp(-absz / p[1]);
            if (g->geometry == SK_GEOM_EXPDISK) return rho;
            double phi = atan2(y, x);
            double m = p[6], tanp = p[7], R0 = p[8], phi0 = p[9], w = p[10], N = p[11], cn = p[12];
            double gamma = log(R / R0) / tanp + phi0 + 0.5 * M_PI / m;
            double perturbation = (1.0 - w) + w * cn * pow_even(sin(0.5 * m * (gamma - phi)), 2 * N);
            return rho * perturbation;
        }
        case SK_GEOM_RING:
        {
            double R = sqrt(x * x + y * y);
            double u = (R - p[0]) / (M_SQRT2 * p[1]);
            return p[3] * exp(-u * u) * exp(-fabs(z) / p[2]);
        }
    }
    return 0.0;
}

/* DensityTreePolicy::needsSubdivide, DensityTreePolicy.cpp:116-227 (dust criteria; samples from the stream
 * (seed, "TREE") with counter = node index) */
static int needs_subdivide(const sk_tree_policy_t* pol, int nmedia, const sk_density_geometry_t* media, double dust_mass,
                           uint32_t seed, int node, int level, const double* box)
{
    if (level < pol->min_level) return 1;
    if (level >= pol->max_level) return 0;
    rng_t g;
    rng_init(&g, seed, 0x54524545u, (uint64_t)node);
    double rhosum = 0., rhomin = DBL_MAX, rhomax = 0.;
    for (int i = 0; i != pol->num_samples; ++i)
    {
        /* Random::position(Box), Random.cpp:168-176 */
        double ux = uniform(&g);
        double uy = uniform(&g);
        double uz = uniform(&g);
        double x = box[0] + ux * (box[3] - box[0]);
        double y = box[1] + uy * (box[4] - box[1]);
        double z = box[2] + uz * (box[5] - box[2]);
        double rhoi = 0.;
        for (int h = 0; h < nmedia; ++h) rhoi += media[h].mass * geom_density(&media[h], x, y, z);
        rhosum += rhoi;
        if (rhoi < rhomin) rhomin = rhoi;
        if (rhoi > rhomax) rhomax = rhoi;
    }
    double rho = rhosum / pol->num_samples;
    double dx = box[3] - box[0], dy = box[4] - box[1], dz = box[5] - box[2];
    double V = dx * dy * dz;
    double M = rho * V;
    if (pol->max_dust_fraction > 0. && M / dust_mass > pol->max_dust_fraction) return 1;
    if (pol->max_dust_optical_depth > 0. && pol->dust_kappa * rho * sqrt(dx * dx + dy * dy + dz * dz) > pol->max_dust_optical_depth)
        return 1;
    if (pol->max_dust_density_dispersion > 0.)
    {
        double q = rhomax > 0 ? (rhomax - rhomin) / rhomax : 0.;
        if (q > pol->max_dust_density_dispersion) return 1;
    }
    return 0;
}

/* DensityTreePolicy::constructTree, DensityTreePolicy.cpp:242-309: breadth-first, level by level; then the grid is set
 * as by sko_set_grid_octree */
int sko_build_octree(sko_engine_t* e, const double extent[6], const sk_tree_policy_t* pol, int32_t num_media,
                     const sk_density_geometry_t* media, uint64_t* num_nodes, uint64_t* num_cells)
{
    if (!e || !extent || !pol || !media || num_media < 1) return fail(SK_ERR_INVALID, "bad tree policy");
    double dust_mass = 0.;
    for (int h = 0; h < num_media; ++h) dust_mass += media[h].mass;
    size_t cap = 1024, nn = 1;
    int32_t* fc = (int32_t*)malloc(cap * sizeof(int32_t));
    int32_t* lev = (int32_t*)malloc(cap * sizeof(int32_t));
    double* box = (double*)malloc(6 * cap * sizeof(double));
    memcpy(box, extent, 6 * sizeof(double));
    lev[0] = 0;
    size_t lbeg = 0, lend = 1;
    while (lend != lbeg)
    {
        for (size_t l = lbeg; l != lend; ++l)
        {
            fc[l] = -1;
            if (!needs_subdivide(pol, num_media, media, dust_mass, (uint32_t)e->cfg.seed, (int)l, lev[l], box + 6 * l)) continue;
            if (nn + 8 > cap)
            {
                cap *= 2;
                fc = (int32_t*)realloc(fc, cap * sizeof(int32_t));
                lev = (int32_t*)realloc(lev, cap * sizeof(int32_t));
                box = (double*)realloc(box, 6 * cap * sizeof(double));
            }
            fc[l] = (int32_t)nn;
            const double* b = box + 6 * l;
            double cx = 0.5 * (b[0] + b[3]), cy = 0.5 * (b[1] + b[4]), cz = 0.5 * (b[2] + b[5]); /* Box.hpp:135 */
            for (int c = 0; c < 8; ++c) /* OctTreeNode::createChildren, OctTreeNode.cpp:22-35 */
            {
                double* cb = box + 6 * (nn + c);
                cb[0] = (c & 1) ? cx : b[0];
                cb[3] = (c & 1) ? b[3] : cx;
                cb[1] = (c & 2) ? cy : b[1];
                cb[4] = (c & 2) ? b[4] : cy;
                cb[2] = (c & 4) ? cz : b[2];
                cb[5] = (c & 4) ? b[5] : cz;
                lev[nn + c] = lev[l] + 1;
            }
            nn += 8;
        }
        lbeg = lend;
        lend = nn;
    }
    int rc = sko_set_grid_octree(e, extent, (int32_t)nn, fc);
    free(fc);
    free(lev);
    free(box);
    if (rc) return rc;
    if (num_nodes) *num_nodes = nn;
    if (num_cells) *num_cells = (uint64_t)e->nx;
    return SK_OK;
}

int sko_read_octree(sko_engine_t* e, int32_t* first_child)
{
    if (!e || !first_child || e->grid_kind != 2) return fail(SK_ERR_STATE, "the engine holds no octree");
    memcpy(first_child, e->first_child, (size_t)e->nnodes * sizeof(int32_t));
    return SK_OK;
}

/* MediumSystem::setupSelfAfter cell loop, MediumSystem.cpp:286-330, with PropertySampler (MediumSystem.cpp:46-106):
 * samples from the stream (seed, "CELL") with counter = cell index */
int sko_sample_medium(sko_engine_t* e, const sk_density_geometry_t* medium, int32_t num_samples)
{
    if (!e || !medium || num_samples < 1) return fail(SK_ERR_INVALID, "bad medium");
    if (e->grid_kind != 1 && e->grid_kind != 2) return fail(SK_ERR_UNSUPPORTED, "density sampling needs a Cartesian or octree grid");
    int nc = grid_num_cells(e);
    free(e->dens);
    free(e->vol);
    free(e->vel);
    e->vel = NULL;
    update_kin(e);
    e->dens = (double*)malloc((size_t)nc * sizeof(double));
    e->vol = (double*)malloc((size_t)nc * sizeof(double));
    e->ncells = nc;
    e->nmed = 1;
    for (int m = 0; m < nc; ++m)
    {
        double b[6];
        if (e->grid_kind == 1)
        {
            int k = m % e->nz, j = (m / e->nz) % e->ny, i = m / (e->nz * e->ny);
            b[0] = e->xv[i];
            b[3] = e->xv[i + 1];
            b[1] = e->yv[j];
            b[4] = e->yv[j + 1];
            b[2] = e->zv[k];
            b[5] = e->zv[k + 1];
        }
        else
            memcpy(b, e->node_box + 6 * (size_t)e->node_of_cell[m], sizeof b);
        double n;
        if (num_samples == 1)
            n = medium->number * geom_density(medium, 0.5 * (b[0] + b[3]), 0.5 * (b[1] + b[4]), 0.5 * (b[2] + b[5]));
        else
        {
            rng_t g;
            rng_init(&g, (uint32_t)e->cfg.seed, 0x43454c4cu, (uint64_t)m);
            double sum = 0.;
            for (int i = 0; i != num_samples; ++i)
            {
                double ux = uniform(&g);
                double uy = uniform(&g);
                double uz = uniform(&g);
                double x = b[0] + ux * (b[3] - b[0]);
                double y = b[1] + uy * (b[4] - b[1]);
                double z = b[2] + uz * (b[5] - b[2]);
                sum += medium->number * geom_density(medium, x, y, z);
            }
            n = sum / num_samples;
        }
        e->dens[m] = n;
        e->vol[m] = (b[3] - b[0]) * (b[4] - b[1]) * (b[5] - b[2]);
    }
    return SK_OK;
}

/* ParticleMedium: the smoothed-particle density of ParticleSnapshot::density(Position) (ParticleSnapshot.cpp:248-258) with the
 * CubicSplineSmoothingKernel (CubicSplineSmoothingKernel.cpp:39-47), sampled per cell like any medium by the cell loop of
 * MediumSystem::setupSelfAfter (MediumSystem.cpp:286-330, PropertySampler :46-106).  The reference sums over the particles
 * of the search block that holds the position (BoxSearch::entitiesFor, BoxSearch.cpp:201-207: ascending particle index);
 * particles that do not reach the position contribute an exact zero, so the sum over ALL particles in ascending index is
 * the same number -- which is what this restatement does.  particles[5*m..] = x y z h M. */
static double sph_kernel_density(double u)
{
    if (u < 0.0 || u >= 1.0)
        return 0.0;
    else if (u < 0.5)
        return 8.0 / M_PI * (1.0 - 6.0 * u * u * (1.0 - u));
    else
        return 8.0 / M_PI * 2.0 * (1.0 - u) * (1.0 - u) * (1.0 - u);
}
static double sph_density(const double* particles, int np, double x, double y, double z)
{
    double sum = 0.;
    for (int m = 0; m < np; ++m)
    {
        const double* q = particles + 5 * (size_t)m;
        const double dx = x - q[0], dy = y - q[1], dz = z - q[2];
        const double h = q[3];
        const double r2 = dx * dx + dy * dy + dz * dz;
        if (r2 >= h * h) continue; /* u >= 1: an exact zero */
        const double u = sqrt(r2) / h;
        sum += sph_kernel_density(u) * (q[4] / (h * h * h)); /* Particle::density(), ParticleSnapshot.cpp:45 */
    }
    return sum > 0. ? sum : 0.;
}
int sko_sample_medium_particles(sko_engine_t* e, int32_t num_particles, const double* particles, double density_scale,
                                int32_t num_samples)
{
    if (!e || !particles || num_particles < 1 || num_samples < 1) return fail(SK_ERR_INVALID, "bad particle medium");
    if (!e->grid_kind) return fail(SK_ERR_STATE, "set the grid before the medium");
    if (e->grid_kind == 3 && !e->vbox) return fail(SK_ERR_STATE, "sampling in Voronoi cells needs their extents");
    if (e->grid_kind == 3 && num_samples == 1)
        return fail(SK_ERR_UNSUPPORTED, "the density at the centroid of a Voronoi cell (numDensitySamples = 1)");
    for (int m = 0; m < num_particles; ++m)
        if (!(particles[5 * (size_t)m + 3] > 0.)) return fail(SK_ERR_INVALID, "smoothing length must be positive");
    const int nc = grid_num_cells(e);
    double* dens = (double*)malloc((size_t)nc * sizeof(double));
    double* vol = NULL;
    if (e->grid_kind != 3)
        vol = (double*)malloc((size_t)nc * sizeof(double));
    else if (e->vvol)
        vol = dupd(e->vvol, nc);
    for (int m = 0; m < nc; ++m)
    {
        double b[6];
        if (e->grid_kind == 1)
        {
            int k = m % e->nz, j = (m / e->nz) % e->ny, i = m / (e->nz * e->ny);
            b[0] = e->xv[i];
            b[3] = e->xv[i + 1];
            b[1] = e->yv[j];
            b[4] = e->yv[j + 1];
            b[2] = e->zv[k];
            b[5] = e->zv[k + 1];
        }
        else if (e->grid_kind == 2)
            memcpy(b, e->node_box + 6 * (size_t)e->node_of_cell[m], sizeof b);
        else
            memcpy(b, e->vbox + 6 * (size_t)m, sizeof b);
        double n;
        if (num_samples == 1)
            n = sph_density(particles, num_particles, 0.5 * (b[0] + b[3]), 0.5 * (b[1] + b[4]), 0.5 * (b[2] + b[5])) * density_scale;
        else
        {
            rng_t g;
            rng_init(&g, (uint32_t)e->cfg.seed, 0x43454c4cu, (uint64_t)m);
            double sum = 0.;
            for (int i = 0; i != num_samples; ++i)
            {
                double x, y, z;
                if (e->grid_kind == 3)
                {
                    /* VoronoiMeshSnapshot::generatePosition(m), VoronoiMeshSnapshot.cpp:976-989 */
                    const double* sm = e->vsite + 3 * (size_t)m;
                    int found = 0;
                    x = sm[0];
                    y = sm[1];
                    z = sm[2];
                    for (int it = 0; it < 10000 && !found; ++it)
                    {
                        double ux = uniform(&g), uy = uniform(&g), uz = uniform(&g);
                        double px = b[0] + ux * (b[3] - b[0]);
                        double py = b[1] + uy * (b[4] - b[1]);
                        double pz = b[2] + uz * (b[5] - b[2]);
                        double dx = px - sm[0], dy = py - sm[1], dz = pz - sm[2];
                        double target = dx * dx + dy * dy + dz * dz;
                        found = 1;
                        for (int64_t q = e->vnbr_off[m]; q < e->vnbr_off[m + 1]; ++q)
                        {
                            int id = e->vnbr[q];
                            if (id < 0) continue;
                            const double* t = e->vsite + 3 * (size_t)id;
                            double ex = px - t[0], ey = py - t[1], ez = pz - t[2];
                            if (ex * ex + ey * ey + ez * ez < target)
                            {
                                found = 0;
                                break;
                            }
                        }
                        if (found)
                        {
                            x = px;
                            y = py;
                            z = pz;
                        }
                    }
                }
                else
                {
                    double ux = uniform(&g);
                    double uy = uniform(&g);
                    double uz = uniform(&g);
                    x = b[0] + ux * (b[3] - b[0]);
                    y = b[1] + uy * (b[4] - b[1]);
                    z = b[2] + uz * (b[5] - b[2]);
                }
                sum += sph_density(particles, num_particles, x, y, z) * density_scale;
            }
            n = sum / num_samples;
        }
        dens[m] = n;
        if (e->grid_kind != 3) vol[m] = (b[3] - b[0]) * (b[4] - b[1]) * (b[5] - b[2]);
    }
    free(e->dens);
    free(e->vol);
    free(e->vel);
    e->vel = NULL;
    update_kin(e);
    e->dens = dens;
    e->vol = vol;
    e->ncells = nc;
    e->nmed = 1;
    return SK_OK;
}

int sko_read_medium(sko_engine_t* e, double* number_density, double* volume)
{
    if (!e || !e->ncells) return fail(SK_ERR_STATE, "the engine holds no medium state");
    if (number_density) memcpy(number_density, e->dens, (size_t)e->ncells * sizeof(double));
    if (volume)
    {
        if (!e->vol) return fail(SK_ERR_STATE, "the engine holds no cell volumes");
        memcpy(volume, e->vol, (size_t)e->ncells * sizeof(double));
    }
    return SK_OK;
}

static void update_kin(sko_engine_t* e)
{
    e->kin = e->vel != NULL;
    for (int h = 0; h < e->nsrc; ++h)
        if (e->src[h].s.velocity_kind != SK_VEL_NONE) e->kin = 1;
}

/* MediumState::bulkVelocity(m) for every cell (sk_engine_set_velocities) */
int sko_set_velocities(sko_engine_t* e, int32_t num_cells, const double* velocity)
{
    if (!e) return fail(SK_ERR_INVALID, "null engine");
    free(e->vel);
    e->vel = NULL;
    if (velocity && num_cells > 0)
    {
        if (!e->dens || num_cells != e->ncells) return fail(SK_ERR_INVALID, "velocities do not match the medium state");
        e->vel = dupd(velocity, 3 * (size_t)num_cells);
    }
    update_kin(e);
    return SK_OK;
}

int sko_set_dustmixes(sko_engine_t* e, int32_t num_media, const sk_dustmix_t* mixes)
{
    if (!e || !mixes || num_media < 1 || mixes[0].num_lambda < 2) return fail(SK_ERR_INVALID, "bad dust mix");
    if (num_media > SK_MAX_MEDIA) return fail(SK_ERR_UNSUPPORTED, "number of medium components");
    int n = mixes[0].num_lambda;
    for (int h = 1; h < num_media; ++h)
        if (mixes[h].num_lambda != n || memcmp(mixes[h].lambda_border, mixes[0].lambda_border, n * sizeof(double)))
            return fail(SK_ERR_INVALID, "the dust mixes of one simulation share one wavelength grid (DustMix.cpp:52-98)");
    free(e->lam_border);
    free(e->sig_abs);
    free(e->sig_sca);
    free(e->sig_ext);
    free(e->gpar);
    e->nlam = n;
    e->nmix = num_media;
    e->lam_border = dupd(mixes[0].lambda_border, n);
    e->sig_abs = (double*)malloc((size_t)num_media * n * sizeof(double));
    e->sig_sca = (double*)malloc((size_t)num_media * n * sizeof(double));
    e->gpar = (double*)malloc((size_t)num_media * n * sizeof(double));
    e->sig_ext = (double*)malloc((size_t)num_media * n * sizeof(double));
    for (int h = 0; h < num_media; ++h)
    {
        memcpy(e->sig_abs + (size_t)h * n, mixes[h].sigma_abs, n * sizeof(double));
        memcpy(e->sig_sca + (size_t)h * n, mixes[h].sigma_sca, n * sizeof(double));
        memcpy(e->gpar + (size_t)h * n, mixes[h].asymmpar, n * sizeof(double));
    }
    for (int i = 0; i < num_media * n; ++i) e->sig_ext[i] = e->sig_abs[i] + e->sig_sca[i]; /* DustMix.cpp:160-163 */
    e->mu = mixes[0].mu;
    return SK_OK;
}
int sko_set_dustmix(sko_engine_t* e, const sk_dustmix_t* mix)
{
    return sko_set_dustmixes(e, 1, mix);
}

static void alloc_rf(sko_engine_t* e)
{
    free(e->rf1);
    free(e->rf2);
    free(e->rf2c);
    e->rf1 = e->rf2 = e->rf2c = NULL;
    e->nrf = 0;
    if (e->rf_grid >= 0 && e->ncells > 0)
    {
        e->nrf = e->wlg[e->rf_grid].g.num_bins;
        size_t n = (size_t)e->ncells * e->nrf;
        e->rf1 = (double*)calloc(n, sizeof(double));
        e->rf2 = (double*)calloc(n, sizeof(double));
        e->rf2c = (double*)calloc(n, sizeof(double));
    }
}

int sko_set_wavelength_grids(sko_engine_t* e, int32_t n, const sk_wavelength_grid_t* grids, int32_t rf_grid)
{
    if (!e || n < 0 || (n && !grids) || rf_grid >= n) return fail(SK_ERR_INVALID, "bad wavelength grids");
    if (rf_grid >= 0 && e->ncells <= 0) return fail(SK_ERR_STATE, "set the medium before a radiation field grid");
    free_wlg(e);
    e->wlg = (wlg_t*)calloc(n ? n : 1, sizeof(wlg_t));
    e->nwlg = n;
    for (int i = 0; i < n; ++i)
    {
        sk_wavelength_grid_t* g = &e->wlg[i].g;
        *g = grids[i];
        g->borders = dupd(grids[i].borders, g->num_borders);
        g->ell = dupi(grids[i].ell, g->num_borders + 1);
        g->lambda = dupd(grids[i].lambda, g->num_bins);
        g->dlambda = dupd(grids[i].dlambda, g->num_bins);
    }
    e->rf_grid = rf_grid;
    alloc_rf(e);
    return SK_OK;
}

/* SourceSystem::setupSelfAfter, SourceSystem.cpp:14-41 */
int sko_set_sources(sko_engine_t* e, int32_t n, const sk_source_t* sources, double source_bias)
{
    if (!e || n < 1 || !sources) return fail(SK_ERR_INVALID, "bad sources");
    free_sources(e);
    e->src = (src_t*)calloc(n, sizeof(src_t));
    e->nsrc = n;
    e->source_bias = source_bias;
    for (int i = 0; i < n; ++i)
    {
        sk_source_t* s = &e->src[i].s;
        *s = sources[i];
        s->geom_table_x = s->geom_table_n ? dupd(sources[i].geom_table_x, s->geom_table_n) : NULL;
        s->geom_table_P = s->geom_table_n ? dupd(sources[i].geom_table_P, s->geom_table_n) : NULL;
        s->sed_lambda = s->sed_n ? dupd(sources[i].sed_lambda, s->sed_n) : NULL;
        s->sed_p = s->sed_n ? dupd(sources[i].sed_p, s->sed_n) : NULL;
        s->sed_P = s->sed_n ? dupd(sources[i].sed_P, s->sed_n) : NULL;
        s->oligo_lambda = s->oligo_n ? dupd(sources[i].oligo_lambda, s->oligo_n) : NULL;
    }
    e->Lv = (double*)malloc(n * sizeof(double));
    e->Wv = (double*)malloc(n * sizeof(double));
    e->Iv = (uint64_t*)calloc(n + 1, sizeof(uint64_t));
    double L = 0.;
    for (int h = 0; h < n; ++h) L += sources[h].luminosity;
    e->Ltot = L;
    update_kin(e);
    if (!L) return SK_OK;
    double wLsum = 0., wsum = 0.;
    for (int h = 0; h < n; ++h)
    {
        e->Lv[h] = sources[h].luminosity / L;
        wLsum += sources[h].source_weight * e->Lv[h];
        wsum += sources[h].source_weight;
    }
    double xi = source_bias;
    for (int h = 0; h < n; ++h)
        e->Wv[h] = (1 - xi) * (sources[h].source_weight * e->Lv[h]) / wLsum + xi * sources[h].source_weight / wsum;
    return SK_OK;
}

/* DistantInstrument::setupSelfBefore (DistantInstrument.cpp:39-50), FrameInstrument::setupSelfBefore
 * (FrameInstrument.cpp:12-32), FluxRecorder::finalizeConfiguration (FluxRecorder.cpp:185-300) */
int sko_set_instruments(sko_engine_t* e, int32_t n, const sk_instrument_t* instruments, int32_t has_medium_emission)
{
    if (!e || n < 0 || (n && !instruments)) return fail(SK_ERR_INVALID, "bad instruments");
    free_instruments(e);
    e->instr = (instr_t*)calloc(n ? n : 1, sizeof(instr_t));
    e->ninstr = n;
    e->has_medium_emission = has_medium_emission;
    for (int i = 0; i < n; ++i)
    {
        instr_t* q = &e->instr[i];
        q->d = instruments[i];
        const sk_instrument_t* d = &q->d;
        if (d->wavelength_grid < 0 || d->wavelength_grid >= e->nwlg)
            return fail(SK_ERR_INVALID, "instrument wavelength grid index out of range");
        if (d->num_scattering_levels > MAX_LEVELS) return fail(SK_ERR_UNSUPPORTED, "too many scattering levels");
        q->costheta = cos(d->inclination);
        q->sintheta = sin(d->inclination);
        q->cosphi = cos(d->azimuth);
        q->sinphi = sin(d->azimuth);
        q->cosomega = cos(d->roll);
        q->sinomega = sin(d->roll);
        direction_from_angles(d->inclination, d->azimuth, q->kobs);
        q->include_sed = (d->kind == SK_INSTR_SED || d->kind == SK_INSTR_FULL);
        q->include_ifu = (d->kind == SK_INSTR_FRAME || d->kind == SK_INSTR_FULL);
        q->radius2 = d->radius * d->radius;
        if (q->include_ifu)
        {
            q->xpmin = d->center_x - 0.5 * d->field_of_view_x;
            q->xpsiz = d->field_of_view_x / d->num_pixels_x;
            q->ypmin = d->center_y - 0.5 * d->field_of_view_y;
            q->ypsiz = d->field_of_view_y / d->num_pixels_y;
            q->npix = (size_t)d->num_pixels_x * d->num_pixels_y;
        }
        q->nl = e->wlg[d->wavelength_grid].g.num_bins;
        q->record_total_only = !d->record_components;
        /* DistantInstrument::determineSameObserverAsPreceding, DistantInstrument.cpp:54-62 */
        q->same_as_preceding = 0;
        if (i > 0)
        {
            const sk_instrument_t* p = &e->instr[i - 1].d;
            if (d->distance == p->distance && d->inclination == p->inclination && d->azimuth == p->azimuth
                && d->roll == p->roll)
                q->same_as_preceding = 1;
        }
        size_t lensed = q->include_sed ? (size_t)q->nl : 0;
        size_t lenifu = q->include_ifu ? q->npix * q->nl : 0;
        for (int c = 0; c < NUM_COMP; ++c)
        {
            int need;
            if (q->record_total_only)
                need = (c == SK_COMP_TOTAL);
            else if (c == SK_COMP_TRANSPARENT || c == SK_COMP_PRIMARY_DIRECT || c == SK_COMP_PRIMARY_SCATTERED)
                need = 1;
            else if (c == SK_COMP_SECONDARY_DIRECT || c == SK_COMP_SECONDARY_SCATTERED
                     || c == SK_COMP_SECONDARY_TRANSPARENT)
                need = has_medium_emission;
            else if (c >= SK_COMP_PRIMARY_SCATTERED_LEVEL)
                need = (c - SK_COMP_PRIMARY_SCATTERED_LEVEL) < d->num_scattering_levels;
            else
                need = 0;
            if (need && lensed) q->sed[c] = (double*)calloc(lensed, sizeof(double));
            if (need && lenifu) q->ifu[c] = (double*)calloc(lenifu, sizeof(double));
        }
        if (d->record_statistics && lensed)
            for (int k = 0; k < 5; ++k) q->wsed[k] = (double*)calloc(lensed, sizeof(double));
        if (d->record_statistics && lenifu)
            for (int k = 0; k < 5; ++k) q->wifu[k] = (double*)calloc(lenifu, sizeof(double));
    }
    return SK_OK;
}

static void free_secondary(sko_engine_t* e)
{
    free(e->sec_T);
    free(e->sec_planckabs);
    free(e->sec_rfsig);
    free(e->sec_cmb);
    e->sec_cmb = NULL;
    free(e->sec_emsig);
    free(e->sec_lambda);
    free(e->sec_pv);
    free(e->sec_Pv);
    free(e->sec_Lv);
    free(e->sec_Wv);
    free(e->sec_Iv);
    e->sec_T = e->sec_planckabs = e->sec_rfsig = e->sec_emsig = e->sec_lambda = e->sec_pv = e->sec_Pv = NULL;
    e->sec_Lv = e->sec_Wv = NULL;
    e->sec_Iv = NULL;
    e->has_secondary = e->secondary_ready = 0;
}

int sko_set_secondary_media(sko_engine_t* e, int32_t num_media, const sk_secondary_t* sec)
{
    if (!e || !sec || num_media < 1) return fail(SK_ERR_INVALID, "null argument");
    if (num_media > SK_MAX_MEDIA) return fail(SK_ERR_UNSUPPORTED, "number of medium components");
    if (e->rf_grid < 0) return fail(SK_ERR_STATE, "dust emission needs a radiation field grid");
    if (e->grid_kind == 3 && !e->vbox)
        return fail(SK_ERR_STATE, "dust emission from a Voronoi grid needs the cell extents (sko_set_voronoi_extents)");
    if (sec->emission_grid < 0 || sec->emission_grid >= e->nwlg) return fail(SK_ERR_INVALID, "bad emission grid index");
    for (int h = 0; h < num_media; ++h)
    {
        if (sec[h].num_temperatures < 2 || !sec[h].temperature || !sec[h].planck_abs || !sec[h].rf_sigma_abs || !sec[h].em_sigma_abs)
            return fail(SK_ERR_INVALID, "missing emission calculator tables");
        if (sec[h].emission_grid != sec->emission_grid || sec[h].num_temperatures != sec->num_temperatures
            || memcmp(sec[h].temperature, sec->temperature, sec->num_temperatures * sizeof(double)))
            return fail(SK_ERR_INVALID, "the emission calculators of the dust components must share their grids");
    }
    free_secondary(e);
    e->sec = *sec;
    e->sec_nmed = num_media;
    const sk_wavelength_grid_t* g = &e->wlg[sec->emission_grid].g;
    int n = g->num_bins;
    const int nT = sec->num_temperatures;
    e->sec_nem = n + 2;
    e->sec_T = dupd(sec->temperature, nT);
    e->sec_planckabs = (double*)malloc((size_t)num_media * nT * sizeof(double));
    e->sec_rfsig = (double*)malloc((size_t)num_media * e->nrf * sizeof(double));
    e->sec_cmb = (double*)calloc(e->nrf ? e->nrf : 1, sizeof(double));
    if (sec->rf_cmb) memcpy(e->sec_cmb, sec->rf_cmb, e->nrf * sizeof(double));
    e->sec_emsig = (double*)malloc((size_t)num_media * (n + 2) * sizeof(double));
    for (int h = 0; h < num_media; ++h)
    {
        memcpy(e->sec_planckabs + (size_t)h * nT, sec[h].planck_abs, nT * sizeof(double));
        memcpy(e->sec_rfsig + (size_t)h * e->nrf, sec[h].rf_sigma_abs, e->nrf * sizeof(double));
        memcpy(e->sec_emsig + (size_t)h * (n + 2), sec[h].em_sigma_abs, (n + 2) * sizeof(double));
    }
    /* DisjointWavelengthGrid::extlambdav, DisjointWavelengthGrid.cpp:346-356 */
    e->sec_lambda = (double*)malloc((n + 2) * sizeof(double));
    e->sec_lambda[0] = g->borders[0];
    for (int ell = 0; ell < n; ++ell) e->sec_lambda[ell + 1] = g->lambda[ell];
    e->sec_lambda[n + 1] = g->borders[g->num_borders - 1];
    e->has_secondary = 1;
    return SK_OK;
}
int sko_set_secondary(sko_engine_t* e, const sk_secondary_t* sec)
{
    return sko_set_secondary_media(e, 1, sec);
}

int sko_clear_instruments(sko_engine_t* e)
{
    if (!e) return fail(SK_ERR_INVALID, "null engine");
    for (int i = 0; i < e->ninstr; ++i)
    {
        instr_t* q = &e->instr[i];
        size_t lensed = q->include_sed ? (size_t)q->nl : 0, lenifu = q->include_ifu ? q->npix * q->nl : 0;
        for (int c = 0; c < NUM_COMP; ++c)
        {
            if (q->sed[c]) memset(q->sed[c], 0, lensed * sizeof(double));
            if (q->ifu[c]) memset(q->ifu[c], 0, lenifu * sizeof(double));
        }
        for (int k = 0; k < 5; ++k)
        {
            if (q->wsed[k]) memset(q->wsed[k], 0, lensed * sizeof(double));
            if (q->wifu[k]) memset(q->wifu[k], 0, lenifu * sizeof(double));
        }
    }
    return SK_OK;
}

/* MediumSystem::clearRadiationField, MediumSystem.cpp:1279-1290 */
int sko_clear_rf(sko_engine_t* e, int32_t primary)
{
    if (!e) return fail(SK_ERR_INVALID, "null engine");
    size_t n = (size_t)e->ncells * e->nrf;
    if (!n) return SK_OK;
    if (primary)
    {
        memset(e->rf1, 0, n * sizeof(double));
        memset(e->rf2, 0, n * sizeof(double));
    }
    else
        memset(e->rf2c, 0, n * sizeof(double));
    return SK_OK;
}

/* SourceSystem::prepareForLaunch, SourceSystem.cpp:75-97 */
int sko_prepare_primary(sko_engine_t* e, uint64_t num_packets)
{
    if (!e || !e->nsrc) return fail(SK_ERR_STATE, "no sources");
    if (!e->Ltot) return fail(SK_ERR_INVALID, "Cannot launch primary source photon packets when total luminosity is zero");
    int Ns = e->nsrc;
    e->Iv[0] = 0;
    double W = 0.;
    for (int h = 1; h != Ns; ++h)
    {
        W += e->Wv[h - 1];
        uint64_t idx = (uint64_t)round(W * (double)num_packets);
        e->Iv[h] = idx < num_packets ? idx : num_packets;
    }
    e->Iv[Ns] = num_packets;
    e->Lpp = e->Ltot / (double)num_packets;
    e->npackets = num_packets;
    return SK_OK;
}

static int index_for_lambda(const sko_engine_t* e, double lambda);
static double interp_loglog(double x, double x1, double x2, double f1, double f2);
static double interp_linlin(double x, double x1, double x2, double f1, double f2);
static double planck(double lambda, double T);
static double gln(double p, double x);

/* SecondarySourceSystem::prepareForLaunch (SecondarySourceSystem.cpp:84-126) for the single DustSecondarySource:
 * DustSecondarySource::prepareLuminosities / preparePacketMap (DustSecondarySource.cpp:26-146) with the AllCellsLibrary
 * mapping (identity, AllCellsLibrary.cpp:26-32), followed by the emission spectrum of every emitting cell, which the
 * reference computes lazily: DustCellEmission::calculateSingleSpectrum (DustSecondarySource.cpp:277-285) =
 * MediumSystem::dustEmissionSpectrum (MediumSystem.cpp:1466-1476) <- meanIntensity (:1370-1380) <- DustMix::emissionSpectrum
 * (DustMix.cpp:650-653) <- EquilibriumDustEmissionCalculator::emissivity / equilibriumTemperature (.cpp:120-150), then
 * NR::cdf<interpolateLogLog> over the range of the emission grid (NR.hpp:494-520, NR.cpp:25-60). */
int sko_prepare_secondary(sko_engine_t* e, uint64_t num_packets, double* luminosity)
{
    if (!e || !luminosity) return fail(SK_ERR_INVALID, "null argument");
    if (!e->has_secondary) return fail(SK_ERR_STATE, "call set_secondary first");
    if (!num_packets) return fail(SK_ERR_INVALID, "zero packets");
    const int M = e->ncells, nrf = e->nrf, nem = e->sec_nem;
    const sk_wavelength_grid_t* rfg = &e->wlg[e->rf_grid].g;
    if (!e->sec_Lv)
    {
        e->sec_Lv = (double*)calloc(M, sizeof(double));
        e->sec_Wv = (double*)calloc(M, sizeof(double));
        e->sec_Iv = (uint64_t*)calloc((size_t)M + 1, sizeof(uint64_t));
        e->sec_pv = (double*)calloc((size_t)M * nem, sizeof(double));
        e->sec_Pv = (double*)calloc((size_t)M * nem, sizeof(double));
    }
    /* luminosities 1: MediumSystem::dustLuminosity, MediumSystem.cpp:1452-1462 */
    for (int m = 0; m < M; ++m)
    {
        double Labs = 0.;
        for (int ell = 0; ell < nrf; ++ell)
        {
            /* MediumSystem::opacityAbs(lambda, m, Dust): sum over the dust components, MediumSystem.cpp:619-630 */
            double opacity = 0.;
            for (int h = 0; h < e->nmed; ++h)
                opacity += e->dens[(size_t)h * M + m] * e->sig_abs[(size_t)h * e->nlam + index_for_lambda(e, rfg->lambda[ell])];
            double rf = 0.;
            rf += e->rf1[(size_t)m * nrf + ell];
            rf += e->rf2[(size_t)m * nrf + ell];
            Labs += opacity * rf;
        }
        e->sec_Lv[m] = Labs;
    }
    /* luminosities 2 */
    double L = 0.;
    for (int m = 0; m < M; ++m) L += e->sec_Lv[m];
    *luminosity = L;
    e->secondary_ready = 0;
    if (!L) return SK_OK; /* SecondarySourceSystem.cpp:93-94: the caller skips the segment */
    for (int m = 0; m < M; ++m) e->sec_Lv[m] /= L;
    /* preparePacketMap: composite-biased launch weights and the history index map (cells in launch order = cell order) */
    double wsum = 0.;
    for (int m = 0; m < M; ++m) wsum += e->sec_Lv[m] > 0 ? 1. : 0.;
    double xi = e->sec.spatial_bias;
    for (int m = 0; m < M; ++m)
    {
        double w = (e->sec_Lv[m] > 0 ? 1. : 0.) / wsum;
        e->sec_Wv[m] = (1 - xi) * e->sec_Lv[m] + xi * w;
    }
    e->sec_Iv[0] = 0;
    double W = 0.;
    for (int p = 1; p != M; ++p)
    {
        W += e->sec_Wv[p - 1];
        uint64_t idx = (uint64_t)round(W * (double)num_packets);
        e->sec_Iv[p] = idx < num_packets ? idx : num_packets;
    }
    e->sec_Iv[M] = num_packets;
    e->sec_Lpp = L / (double)num_packets; /* SecondarySourceSystem.cpp:119 with a single source of weight 1 */
    /* the emission spectrum of every emitting cell */
    const int nT = e->sec.num_temperatures;
    for (int m = 0; m < M; ++m)
    {
        double* pv = e->sec_pv + (size_t)m * nem;
        double* Pv = e->sec_Pv + (size_t)m * nem;
        if (!(e->sec_Lv[m] > 0))
        {
            memset(pv, 0, nem * sizeof(double));
            memset(Pv, 0, nem * sizeof(double));
            continue;
        }
        /* MediumSystem::dustEmissionSpectrum (MediumSystem.cpp:1466-1476): the sum over the dust components of
           DustMix::emissionSpectrum = number density times the emissivity of the component's own mix at its own equilibrium
           temperature in the cell's radiation field (meanIntensity + equilibriumTemperature) */
        double factor = 1. / (4. * M_PI * e->vol[m]);
        for (int i = 0; i < nem; ++i) pv[i] = 0.;
        for (int h = 0; h < e->nmed; ++h)
        {
            const double* rfsig = e->sec_rfsig + (size_t)h * nrf;
            const double* planckabs = e->sec_planckabs + (size_t)h * nT;
            const double* emsig = e->sec_emsig + (size_t)h * nem;
            double inputabs = 0.;
            for (int ell = 0; ell < nrf; ++ell)
            {
                double rf = 0.;
                rf += e->rf1[(size_t)m * nrf + ell];
                rf += e->rf2[(size_t)m * nrf + ell];
                double J = rf * factor / rfg->dlambda[ell];
                inputabs += rfsig[ell] * (J + e->sec_cmb[ell]) * rfg->dlambda[ell]; /* (Jv + _Bcmbv), .cpp:123 */
            }
            double T = 0.;
            if (inputabs > 0.)
            {
                /* NR::clampedValue<interpolateLinLin>(inputabs, _planckabsvv, _Tv), NR.hpp:391-399 with NR::locate */
                int i = inputabs == planckabs[nT - 1] ? nT - 2 : locate_basic(planckabs, inputabs, nT);
                if (i < 0)
                    T = e->sec_T[0];
                else if (i >= nT - 1)
                    T = e->sec_T[nT - 1];
                else
                    T = interp_linlin(inputabs, planckabs[i], planckabs[i + 1], e->sec_T[i], e->sec_T[i + 1]);
            }
            /* emissivity times number density on the extended emission grid */
            const double n = e->dens[(size_t)h * M + m];
            for (int i = 0; i < nem; ++i) pv[i] += n * (emsig[i] * planck(e->sec_lambda[i], T));
        }
        /* NR::cdf<interpolateLogLog>(xv,pv,Pv, extlambdav, ev, range of the grid): the range is [ext[0], ext[n+1]], so
           the axis is the extended grid itself and the two outer values are the log-log interpolants evaluated at the
           end points of the first and last interval */
        {
            const double* x = e->sec_lambda;
            double first = interp_loglog(x[0], x[0], x[1], pv[0], pv[1]);
            double last = interp_loglog(x[nem - 1], x[nem - 2], x[nem - 1], pv[nem - 2], pv[nem - 1]);
            pv[0] = first;
            pv[nem - 1] = last;
            Pv[0] = 0.;
            for (int i = 0; i != nem - 1; ++i)
            {
                double area = 0.;
                if (pv[i] > 0 && pv[i + 1] > 0)
                {
                    double alpha = log(pv[i + 1] / pv[i]) / log(x[i + 1] / x[i]);
                    area = pv[i] * x[i] * gln(-alpha, x[i + 1] / x[i]);
                }
                Pv[i + 1] = Pv[i] + area;
            }
            double norm = Pv[nem - 1];
            if (norm > 0.)
                for (int i = 0; i < nem; ++i)
                {
                    pv[i] /= norm;
                    Pv[i] /= norm;
                }
        }
    }
    e->secondary_ready = 1;
    return SK_OK;
}

/* ------------------------------------------------------------------------------------------------ */
/* path segment generators                                                                          */
/* ------------------------------------------------------------------------------------------------ */

typedef struct {
    int state; /* 0 Unknown, 1 Inside, 2 Outside (PathSegmentGenerator.hpp State) */
    double rx, ry, rz, kx, ky, kz;
    int m;
    double ds;
    int i, j, k; /* cartesian */
    int node;    /* tree: current node or -1; voronoi: current cell */
    int hint;    /* voronoi: a cell near the start of the path (-1 = none) to start the nearest-site walk from */
} gen_t;

static int box_contains(const double* b, double x, double y, double z)
{
    /* Box::contains, SKIRT/utils/Box.hpp:99-109 (closed on all sides) */
    return x >= b[0] && x <= b[3] && y >= b[1] && y <= b[4] && z >= b[2] && z <= b[5];
}

/* PathSegmentGenerator::moveInside, SKIRT/utils/PathSegmentGenerator.cpp:11-112 */
static int move_inside(gen_t* g, const double* box, double eps)
{
    g->m = -1;
    g->ds = 0.;
    g->state = 2;
    double cumds = 0.;
    if (g->rx <= box[0])
    {
        if (g->kx <= 0.0)
            return 0;
        else
        {
            double ds = (box[0] - g->rx) / g->kx;
            g->rx = box[0] + eps;
            g->ry += g->ky * ds;
            g->rz += g->kz * ds;
            cumds += ds;
        }
    }
    else if (g->rx >= box[3])
    {
        if (g->kx >= 0.0)
            return 0;
        else
        {
            double ds = (box[3] - g->rx) / g->kx;
            g->rx = box[3] - eps;
            g->ry += g->ky * ds;
            g->rz += g->kz * ds;
            cumds += ds;
        }
    }
    if (g->ry <= box[1])
    {
        if (g->ky <= 0.0)
            return 0;
        else
        {
            double ds = (box[1] - g->ry) / g->ky;
            g->rx += g->kx * ds;
            g->ry = box[1] + eps;
            g->rz += g->kz * ds;
            cumds += ds;
        }
    }
    else if (g->ry >= box[4])
    {
        if (g->ky >= 0.0)
            return 0;
        else
        {
            double ds = (box[4] - g->ry) / g->ky;
            g->rx += g->kx * ds;
            g->ry = box[4] - eps;
            g->rz += g->kz * ds;
            cumds += ds;
        }
    }
    if (g->rz <= box[2])
    {
        if (g->kz <= 0.0)
            return 0;
        else
        {
            double ds = (box[2] - g->rz) / g->kz;
            g->rx += g->kx * ds;
            g->ry += g->ky * ds;
            g->rz = box[2] + eps;
            cumds += ds;
        }
    }
    else if (g->rz >= box[5])
    {
        if (g->kz >= 0.0)
            return 0;
        else
        {
            double ds = (box[5] - g->rz) / g->kz;
            g->rx += g->kx * ds;
            g->ry += g->ky * ds;
            g->rz = box[5] - eps;
            cumds += ds;
        }
    }
    if (!box_contains(box, g->rx, g->ry, g->rz)) return 0;
    g->m = -1;
    g->ds = cumds;
    g->state = 1;
    return 1;
}

static void gen_start(gen_t* g, const double r[3], const double k[3])
{
    g->state = 0;
    g->rx = r[0];
    g->ry = r[1];
    g->rz = r[2];
    g->kx = k[0];
    g->ky = k[1];
    g->kz = k[2];
    g->node = -1;
    g->hint = -1;
}

/* CartesianSpatialGrid::MySegmentGenerator::next, CartesianSpatialGrid.cpp:95-162 */
static int next_cartesian(const sko_engine_t* e, gen_t* g)
{
    switch (g->state)
    {
        case 0:
        {
            if (!move_inside(g, e->extent, e->eps)) return 0;
            g->i = locate_clip(e->xv, e->nx + 1, g->rx);
            g->j = locate_clip(e->yv, e->ny + 1, g->ry);
            g->k = locate_clip(e->zv, e->nz + 1, g->rz);
            if (g->ds > 0.) return 1;
        }
        /* fall through */
        case 1:
        {
            int m = g->k + e->nz * g->j + e->nz * e->ny * g->i; /* CartesianSpatialGrid.cpp:210-213 */
            double xE = (g->kx < 0.0) ? e->xv[g->i] : e->xv[g->i + 1];
            double yE = (g->ky < 0.0) ? e->yv[g->j] : e->yv[g->j + 1];
            double zE = (g->kz < 0.0) ? e->zv[g->k] : e->zv[g->k + 1];
            double dsx = (fabs(g->kx) > 1e-15) ? (xE - g->rx) / g->kx : DBL_MAX;
            double dsy = (fabs(g->ky) > 1e-15) ? (yE - g->ry) / g->ky : DBL_MAX;
            double dsz = (fabs(g->kz) > 1e-15) ? (zE - g->rz) / g->kz : DBL_MAX;
            if (dsx <= dsy && dsx <= dsz)
            {
                g->m = m;
                g->ds = dsx;
                g->rx = xE;
                g->ry += g->ky * dsx;
                g->rz += g->kz * dsx;
                g->i += (g->kx < 0.0) ? -1 : 1;
                if (g->i >= e->nx || g->i < 0) g->state = 2;
            }
            else if (dsy < dsx && dsy <= dsz)
            {
                g->m = m;
                g->ds = dsy;
                g->ry = yE;
                g->rx += g->kx * dsy;
                g->rz += g->kz * dsy;
                g->j += (g->ky < 0.0) ? -1 : 1;
                if (g->j >= e->ny || g->j < 0) g->state = 2;
            }
            else
            {
                g->m = m;
                g->ds = dsz;
                g->rz = zE;
                g->rx += g->kx * dsz;
                g->ry += g->ky * dsz;
                g->k += (g->kz < 0.0) ? -1 : 1;
                if (g->k >= e->nz || g->k < 0) g->state = 2;
            }
            return 1;
        }
        default: break;
    }
    return 0;
}

/* TreeNode::leafChild (TreeNode.cpp:65-76) with OctTreeNode::child (OctTreeNode.cpp:37-42) */
static int tree_leaf_child(const sko_engine_t* e, double x, double y, double z)
{
    if (!box_contains(e->node_box, x, y, z)) return -1;
    int node = 0;
    while (e->first_child[node] >= 0)
    {
        int fc = e->first_child[node];
        const double* c0 = e->node_box + 6 * (size_t)fc; /* CHILD_0->rmax() is the centre of the node */
        int l = (x < c0[3] ? 0 : 1) + (y < c0[4] ? 0 : 2) + (z < c0[5] ? 0 : 4);
        node = fc + l;
    }
    return node;
}

/* TreeSpatialGrid::MySegmentGenerator::next, TreeSpatialGrid.cpp:140-216.
 * TreeNode::neighbor(wall, r) (TreeNode.cpp:103-112) returns the neighbouring leaf that contains r and the
 * caller falls back to the top-down search when there is none; both give the leaf containing r (they can
 * differ only when r lies exactly on a face shared by two candidate leaves), so the lookup is restated as
 * the top-down search alone. */
static int next_tree(sko_engine_t* e, gen_t* g)
{
    switch (g->state)
    {
        case 0:
        {
            if (!move_inside(g, e->extent, e->eps)) return 0;
            g->node = tree_leaf_child(e, g->rx, g->ry, g->rz);
            if (g->ds > 0.) return 1;
        }
        /* fall through */
        case 1:
        {
            const double* b = e->node_box + 6 * (size_t)g->node;
            double xnext = (g->kx < 0.0) ? b[0] : b[3];
            double ynext = (g->ky < 0.0) ? b[1] : b[4];
            double znext = (g->kz < 0.0) ? b[2] : b[5];
            double dsx = (fabs(g->kx) > 1e-15) ? (xnext - g->rx) / g->kx : DBL_MAX;
            double dsy = (fabs(g->ky) > 1e-15) ? (ynext - g->ry) / g->ky : DBL_MAX;
            double dsz = (fabs(g->kz) > 1e-15) ? (znext - g->rz) / g->kz : DBL_MAX;
            double ds;
            if (dsx <= dsy && dsx <= dsz)
                ds = dsx;
            else if (dsy <= dsx && dsy <= dsz)
                ds = dsy;
            else
                ds = dsz;
            double adv = ds + e->eps;
            g->rx += g->kx * adv;
            g->ry += g->ky * adv;
            g->rz += g->kz * adv;
            g->m = e->cell_of_node[g->node];
            g->ds = ds;
            int oldnode = g->node;
            g->node = tree_leaf_child(e, g->rx, g->ry, g->rz);
            if (g->node == oldnode)
            {
                /* PathSegmentGenerator::propagateToNextAfter, PathSegmentGenerator.hpp:148-153 */
                g->rx = nextafter(g->rx, (g->kx < 0.) ? -DBL_MAX : DBL_MAX);
                g->ry = nextafter(g->ry, (g->ky < 0.) ? -DBL_MAX : DBL_MAX);
                g->rz = nextafter(g->rz, (g->kz < 0.) ? -DBL_MAX : DBL_MAX);
                g->node = tree_leaf_child(e, g->rx, g->ry, g->rz);
            }
            if (g->node < 0 || g->node == oldnode) g->state = 2;
            return 1;
        }
        default: break;
    }
    return 0;
}

/* VoronoiMeshSnapshot::cellIndex (VoronoiMeshSnapshot.cpp:1006-1040): the cell whose site is nearest to the position.  The
 * reference searches block lists and k-d trees; here the search walks the neighbour graph: from cell m move to the
 * neighbour whose site is closest to the position as long as that is closer than the own site.  The segment from a site to
 * a query point inside the (convex) domain leaves the site's cell through a face inside the domain, and the neighbour
 * across that face is closer to the query, so the walk ends exactly in the cell that contains the position.  `hint` < 0
 * starts from the block table. */
static int voronoi_walk(const sko_engine_t* e, double x, double y, double z, int hint)
{
    int m = hint;
    if (m < 0)
    {
        const double* b = e->extent;
        int nb = e->vnb;
        int i = (int)((x - b[0]) / (b[3] - b[0]) * nb);
        int j = (int)((y - b[1]) / (b[4] - b[1]) * nb);
        int k = (int)((z - b[2]) / (b[5] - b[2]) * nb);
        i = i < 0 ? 0 : i >= nb ? nb - 1 : i;
        j = j < 0 ? 0 : j >= nb ? nb - 1 : j;
        k = k < 0 ? 0 : k >= nb ? nb - 1 : k;
        m = e->vblock[((size_t)i * nb + j) * nb + k];
    }
    const double* s = e->vsite;
    double dx = x - s[3 * m], dy = y - s[3 * m + 1], dz = z - s[3 * m + 2];
    double d = dx * dx + dy * dy + dz * dz;
    while (1)
    {
        int best = -1;
        double dbest = d;
        for (int64_t i = e->vnbr_off[m]; i < e->vnbr_off[m + 1]; ++i)
        {
            int mi = e->vnbr[i];
            if (mi < 0) continue;
            double ex = x - s[3 * mi], ey = y - s[3 * mi + 1], ez = z - s[3 * mi + 2];
            double di = ex * ex + ey * ey + ez * ez;
            if (di < dbest)
            {
                dbest = di;
                best = mi;
            }
        }
        if (best < 0) return m;
        m = best;
        d = dbest;
    }
}
static int voronoi_cell_index(const sko_engine_t* e, double x, double y, double z, int hint)
{
    if (!box_contains(e->extent, x, y, z)) return -1;
    return voronoi_walk(e, x, y, z, hint);
}

/* VoronoiMeshSnapshot::MySegmentGenerator::next, VoronoiMeshSnapshot.cpp:1058-1188 */
static int next_voronoi(sko_engine_t* e, gen_t* g)
{
    switch (g->state)
    {
        case 0:
        {
            if (!move_inside(g, e->extent, e->eps)) return 0;
            g->node = voronoi_cell_index(e, g->rx, g->ry, g->rz, g->hint);
            if (g->ds > 0.) return 1;
        }
        /* fall through */
        case 1:
        {
            while (1)
            {
                int mr = g->node;
                const double* pr = e->vsite + 3 * (size_t)mr;
                double sq = DBL_MAX;
                const int NO_INDEX = -99;
                int mq = NO_INDEX;
                for (int64_t i = e->vnbr_off[mr]; i < e->vnbr_off[mr + 1]; ++i)
                {
                    int mi = e->vnbr[i];
                    double si = 0;
                    if (mi >= 0)
                    {
                        const double* pi = e->vsite + 3 * (size_t)mi;
                        double nx = pi[0] - pr[0], ny = pi[1] - pr[1], nz = pi[2] - pr[2];
                        double ndotk = nx * g->kx + ny * g->ky + nz * g->kz;
                        if (ndotk > 0)
                        {
                            double px = 0.5 * (pi[0] + pr[0]), py = 0.5 * (pi[1] + pr[1]), pz = 0.5 * (pi[2] + pr[2]);
                            si = (nx * (px - g->rx) + ny * (py - g->ry) + nz * (pz - g->rz)) / ndotk;
                        }
                    }
                    else
                    {
                        switch (mi)
                        {
                            case -1: si = (e->extent[0] - g->rx) / g->kx; break;
                            case -2: si = (e->extent[3] - g->rx) / g->kx; break;
                            case -3: si = (e->extent[1] - g->ry) / g->ky; break;
                            case -4: si = (e->extent[4] - g->ry) / g->ky; break;
                            case -5: si = (e->extent[2] - g->rz) / g->kz; break;
                            default: si = (e->extent[5] - g->rz) / g->kz; break;
                        }
                    }
                    if (si > 0 && si < sq)
                    {
                        sq = si;
                        mq = mi;
                    }
                }
                if (mq == NO_INDEX)
                {
                    g->rx += g->kx * e->eps;
                    g->ry += g->ky * e->eps;
                    g->rz += g->kz * e->eps;
                    g->node = voronoi_cell_index(e, g->rx, g->ry, g->rz, mr);
                    if (g->node < 0)
                    {
                        g->state = 2;
                        return 0;
                    }
                }
                else
                {
                    double adv = sq + e->eps;
                    g->rx += g->kx * adv;
                    g->ry += g->ky * adv;
                    g->rz += g->kz * adv;
                    g->m = mr;
                    g->ds = sq;
                    g->node = mq;
                    if (mq < 0) g->state = 2;
                    return 1;
                }
            }
        }
        default: break;
    }
    return 0;
}

static int gen_next(sko_engine_t* e, gen_t* g)
{
    return e->grid_kind == 1 ? next_cartesian(e, g) : e->grid_kind == 2 ? next_tree(e, g) : next_voronoi(e, g);
}

/* ------------------------------------------------------------------------------------------------ */
/* photon packet and medium functions                                                               */
/* ------------------------------------------------------------------------------------------------ */

typedef struct {
    double lambda, W; /* PhotonPacket::_lambda, _W = L*lambda (PhotonPacket.hpp:337-340) */
    double r[3], k[3];
    int nscatt;
    int primary_origin;
    uint64_t history;
    int has_tau;
    double tau_obs;
    int ilam; /* DustMix::indexForLambda(lambda); without kinematics constant during the life cycle */
    int m_int; /* PhotonPacket::interactionCellIndex(): the cell of the last interaction point (SpatialGridPath.hpp:150) */
    /* kinematics: the rest-frame emission wavelength and the velocity of the emitter (PhotonPacket::_lambda0, _bvi) */
    double lambda0, vsrc[3];
    int has_vsrc;
} packet_t;

/* ---- kinematics: PhotonPacket::shiftedEmissionWavelength / shiftedReceptionWavelength / perceivedWavelength,
 * PhotonPacket.cpp:133-151 (no Hubble flow) */
#define SK_C_LIGHT 299792458. /* Constants::c() */
static double shifted_emission_wavelength(double lambda, const double k[3], const double v[3])
{
    return lambda * (1 - (k[0] * v[0] + k[1] * v[1] + k[2] * v[2]) / SK_C_LIGHT);
}
/* the wavelength cell m perceives for a packet of wavelength lambda travelling along k */
static double perceived_wavelength(const sko_engine_t* e, double lambda, const double k[3], int m)
{
    if (!e->vel || m < 0) return lambda;
    const double* v = e->vel + 3 * (size_t)m;
    return lambda / (1 - (k[0] * v[0] + k[1] * v[1] + k[2] * v[2]) / SK_C_LIGHT);
}

/* the opacity sum over the medium components in cell m with the sections sig[h*nlam + ilam]: MediumSystem::opacitySca/Ext
 * for spatially constant sections (MediumSystem.cpp:632-662), n_h * sigma_h each (MaterialMix::opacity*) */
static double opacity_sum(const sko_engine_t* e, const double* sig, int ilam, int m)
{
    double result = 0.;
    for (int h = 0; h < e->nmed; ++h) result += e->dens[(size_t)h * e->ncells + m] * sig[(size_t)h * e->nlam + ilam];
    return result;
}

/* DustMix::indexForLambda, DustMix.cpp:276-279 */
static int index_for_lambda(const sko_engine_t* e, double lambda)
{
    return locate_clip(e->lam_border, e->nlam, lambda);
}

/* the index of the dust property tables for the packet in cell m: fixed for the packet without kinematics, else looked up
 * at the wavelength the cell perceives (the "spatially variable cross sections" branches of MediumSystem.cpp) */
static int ilam_in_cell(const sko_engine_t* e, const packet_t* pp, int m)
{
    if (!e->kin) return pp->ilam;
    return index_for_lambda(e, perceived_wavelength(e, pp->lambda, pp->k, m));
}

/* DisjointWavelengthGrid::bin, DisjointWavelengthGrid.cpp:332-341 */
static int wlg_bin(const sk_wavelength_grid_t* g, double lambda)
{
    /* std::upper_bound: first border strictly greater than lambda */
    int lo = 0, hi = g->num_borders;
    while (lo < hi)
    {
        int mid = (lo + hi) >> 1;
        if (lambda < g->borders[mid])
            hi = mid;
        else
            lo = mid + 1;
    }
    return g->ell[lo];
}

/* MediumSystem::setExtinctionOpticalDepths, single constant-section medium branch, MediumSystem.cpp:849-871 -- or, with
 * explicit absorption, MediumSystem::setScatteringAndAbsorptionOpticalDepths (MonteCarloSimulation.cpp:567-570);
 * SpatialGridPath::addSegment, SpatialGridPath.cpp:41-48 */
static void set_extinction_optical_depths(sko_engine_t* e, const packet_t* pp)
{
    gen_t g;
    gen_start(&g, pp->r, pp->k);
    e->nsegs = 0;
    double s = 0.;
    while (gen_next(e, &g))
    {
        if (g.ds > 0.)
        {
            if (e->nsegs == e->capsegs)
            {
                e->capsegs *= 2;
                e->segs = (seg_t*)realloc(e->segs, e->capsegs * sizeof(seg_t));
            }
            s += g.ds;
            seg_t* sg = &e->segs[e->nsegs++];
            sg->m = g.m;
            sg->ds = g.ds;
            sg->s = s;
            sg->tau = 0.;
            sg->tauabs = 0.;
        }
    }
    if (e->cfg.explicit_absorption)
    {
        /* MediumSystem::setScatteringAndAbsorptionOpticalDepths, single constant-section medium, MediumSystem.cpp:905-934 */
        /* (several components: MediumSystem.cpp:937-955) */
        double tauSca = 0., tauAbs = 0.;
        for (int n = 0; n < e->nsegs; ++n)
        {
            seg_t* sg = &e->segs[n];
            if (sg->m >= 0)
            {
                const int il = ilam_in_cell(e, pp, sg->m); /* (kinematics: MediumSystem.cpp:958-972) */
                for (int h = 0; h < e->nmed; ++h)
                {
                    double ns = e->dens[(size_t)h * e->ncells + sg->m] * sg->ds;
                    tauSca += e->sig_sca[(size_t)h * e->nlam + il] * ns;
                    tauAbs += e->sig_abs[(size_t)h * e->nlam + il] * ns;
                }
            }
            sg->tau = tauSca;
            sg->tauabs = tauAbs;
        }
    }
    else
    {
        /* single medium, MediumSystem.cpp:863-871; several media with constant sections, MediumSystem.cpp:874-885 */
        double tau = 0.;
        for (int n = 0; n < e->nsegs; ++n)
        {
            seg_t* sg = &e->segs[n];
            if (sg->m >= 0)
            {
                const int il = ilam_in_cell(e, pp, sg->m); /* (kinematics: MediumSystem.cpp:888-900) */
                for (int h = 0; h < e->nmed; ++h)
                    tau += e->sig_ext[(size_t)h * e->nlam + il] * e->dens[(size_t)h * e->ncells + sg->m] * sg->ds;
            }
            sg->tau = tau;
        }
    }
    e->cnt.forward_paths++;
    e->cnt.forward_segments += e->nsegs;
}

/* MediumSystem::getExtinctionOpticalDepth(pp, infinity), single-medium branch, MediumSystem.cpp:1192-1219 */
static double get_extinction_optical_depth(sko_engine_t* e, const packet_t* ppp)
{
    double L = ppp->W / ppp->lambda;
    if (L <= 0) return INFINITY;
    double taumax = log(L) + 745;
    gen_t g;
    gen_start(&g, ppp->r, ppp->k);
    double tau = 0.;
    e->cnt.peel_paths++;
    while (gen_next(e, &g))
    {
        e->cnt.peel_segments++;
        if (g.m >= 0)
        {
            /* (several media: MediumSystem.cpp:1222-1240; kinematics: MediumSystem.cpp:1242-1258) */
            const int il = ilam_in_cell(e, ppp, g.m);
            for (int h = 0; h < e->nmed; ++h)
                tau += e->sig_ext[(size_t)h * e->nlam + il] * e->dens[(size_t)h * e->ncells + g.m] * g.ds;
            if (tau >= taumax) return INFINITY;
        }
    }
    return tau;
}

/* MonteCarloSimulation::storeRadiationField (constant perceived wavelength branch), MonteCarloSimulation.cpp:638-665;
 * MediumSystem::storeRadiationField, MediumSystem.cpp:1294-1300 */
static void store_radiation_field(sko_engine_t* e, int primary, const packet_t* pp)
{
    if (e->rf_grid < 0) return;
    if (e->kin)
    {
        /* the branch for perceived wavelengths that vary along the path, MonteCarloSimulation.cpp:667-691 */
        double lnExtBeg = 0.;
        double extBeg = 1.;
        double* rf = primary ? e->rf1 : e->rf2c;
        for (int n = 0; n < e->nsegs; ++n)
        {
            const seg_t* sg = &e->segs[n];
            double lnExtEnd = -(sg->tau + sg->tauabs);
            double extEnd = exp(lnExtEnd);
            if (sg->m >= 0)
            {
                double lambda = perceived_wavelength(e, pp->lambda, pp->k, sg->m);
                int ell = wlg_bin(&e->wlg[e->rf_grid].g, lambda);
                if (ell >= 0)
                {
                    double extMean = lnmean4(extEnd, extBeg, lnExtEnd, lnExtBeg);
                    double Lds = (pp->W / lambda) * extMean * sg->ds; /* PhotonPacket::perceivedLuminosity */
                    rf[(size_t)sg->m * e->nrf + ell] += Lds;
                    e->cnt.rf_deposits++;
                }
            }
            lnExtBeg = lnExtEnd;
            extBeg = extEnd;
        }
        return;
    }
    int ell = wlg_bin(&e->wlg[e->rf_grid].g, pp->lambda);
    if (ell < 0) return;
    double luminosity = pp->W / pp->lambda;
    double lnExtBeg = 0.;
    double extBeg = 1.;
    double* rf = primary ? e->rf1 : e->rf2c;
    for (int n = 0; n < e->nsegs; ++n)
    {
        const seg_t* sg = &e->segs[n];
        double lnExtEnd = -(sg->tau + sg->tauabs); /* Segment::tauExt(), SpatialGridPath.hpp:108 */
        double extEnd = exp(lnExtEnd);
        if (sg->m >= 0)
        {
            double extMean = lnmean4(extEnd, extBeg, lnExtEnd, lnExtBeg);
            double Lds = luminosity * extMean * sg->ds;
            rf[(size_t)sg->m * e->nrf + ell] += Lds;
            e->cnt.rf_deposits++;
        }
        lnExtBeg = lnExtEnd;
        extBeg = extEnd;
    }
}

/* SpatialGridPath::findInteractionPoint, SpatialGridPath.cpp:164-206 (extinction-only members) */
static void find_interaction_point(const sko_engine_t* e, double tauinteract, int* m_out, double* s_out, double* tauabs_out)
{
    *tauabs_out = 0.;
    if (e->nsegs == 0)
    {
        *m_out = -1;
        *s_out = 0.;
        return;
    }
    /* std::upper_bound on the cumulative optical depth: first segment with tau > tauinteract */
    int lo = 0, hi = e->nsegs;
    while (lo < hi)
    {
        int mid = (lo + hi) >> 1;
        if (tauinteract < e->segs[mid].tau)
            hi = mid;
        else
            lo = mid + 1;
    }
    if (lo == 0)
    {
        *m_out = e->segs[0].m;
        *s_out = interp_linlin(tauinteract, 0., e->segs[0].tau, 0., e->segs[0].s);
        *tauabs_out = interp_linlin(tauinteract, 0., e->segs[0].tau, 0., e->segs[0].tauabs);
    }
    else if (lo < e->nsegs)
    {
        *m_out = e->segs[lo].m;
        *s_out = interp_linlin(tauinteract, e->segs[lo - 1].tau, e->segs[lo].tau, e->segs[lo - 1].s, e->segs[lo].s);
        *tauabs_out = interp_linlin(tauinteract, e->segs[lo - 1].tau, e->segs[lo].tau, e->segs[lo - 1].tauabs, e->segs[lo].tauabs);
    }
    else
    {
        *m_out = e->segs[lo - 1].m;
        *s_out = e->segs[lo - 1].s;
        *tauabs_out = e->segs[lo - 1].tauabs;
    }
}

/* DustMix.cpp:391-425: valueHG / integralHG / meanHG */
static double value_hg(double g, double costheta)
{
    double t = 1. + g * g - 2. * g * costheta;
    return (1. - g) * (1. + g) / sqrt(t * t * t);
}
static double integral_hg(double g, double cosalpha, double cosbeta)
{
    double ta = sqrt(1. + g * g - 2. * g * cosalpha);
    double tb = sqrt(1. + g * g - 2. * g * cosbeta);
    double f1 = (1. - g) * (1. + g) / g;
    double f2 = (tb - ta) / (tb * ta);
    return f1 * f2;
}
static double mean_hg(double g, double costheta)
{
    const double delta = 4. * M_PI / 180.;
    double theta = acos(costheta);
    double cosalpha = cos(theta - delta);
    double cosbeta = cos(theta + delta);
    if (theta < delta) return (integral_hg(g, 1., cosalpha) + integral_hg(g, 1., cosbeta)) / (2. - cosalpha - cosbeta);
    if (theta > M_PI - delta)
        return (integral_hg(g, cosalpha, -1.) + integral_hg(g, cosbeta, -1.)) / (2. + cosalpha + cosbeta);
    return integral_hg(g, cosalpha, cosbeta) / (cosalpha - cosbeta);
}

/* ------------------------------------------------------------------------------------------------ */
/* instruments                                                                                      */
/* ------------------------------------------------------------------------------------------------ */

/* FluxRecorder::recordContributions for the SED arrays, FluxRecorder.cpp:962-986 */
static void flush_history_stats(instr_t* q)
{
    for (int i = 0; i < q->hist_nsed && q->wsed[0]; ++i)
    {
        double wn = 1.;
        for (int k = 0; k <= 4; ++k)
        {
            q->wsed[k][q->hist_sed_ell[i]] += wn;
            wn *= q->hist_sed_w[i];
        }
    }
    q->hist_nsed = 0;
    /* FluxRecorder::recordContributions for the frame, FluxRecorder.cpp:990-1013 (the list already holds one entry per
       pixel and wavelength bin, which is what sorting and grouping the raw contributions produces) */
    for (int i = 0; i < q->hist_npix; ++i)
    {
        double wn = 1.;
        for (int k = 0; k <= 4; ++k)
        {
            q->wifu[k][q->hist_lell[i]] += wn;
            wn *= q->hist_wpix[i];
        }
    }
    q->hist_npix = 0;
}

/* Instrument::detect for SED/Frame/Full instruments: SEDInstrument.cpp:22-25 + ApertureInstrument.cpp:24-43,
 * FrameInstrument.cpp:37-64, FluxRecorder::detect FluxRecorder.cpp:304-468 */
static void detect(sko_engine_t* e, instr_t* q, packet_t* ppp)
{
    int l = 0;
    double x = ppp->r[0], y = ppp->r[1], z = ppp->r[2];
    if (q->d.kind == SK_INSTR_SED)
    {
        if (q->radius2)
        {
            double xpp = -q->sinphi * x + q->cosphi * y;
            double ypp = -q->cosphi * q->costheta * x - q->sinphi * q->costheta * y + q->sintheta * z;
            double radius2 = xpp * xpp + ypp * ypp;
            if (radius2 > q->radius2) return;
        }
        l = 0;
    }
    else
    {
        double xpp = -q->sinphi * x + q->cosphi * y;
        double ypp = -q->cosphi * q->costheta * x - q->sinphi * q->costheta * y + q->sintheta * z;
        double xp = q->cosomega * xpp - q->sinomega * ypp;
        double yp = q->sinomega * xpp + q->cosomega * ypp;
        int i = (int)floor((xp - q->xpmin) / q->xpsiz);
        int j = (int)floor((yp - q->ypmin) / q->ypsiz);
        if (i < 0 || i >= q->d.num_pixels_x || j < 0 || j >= q->d.num_pixels_y)
            l = -1;
        else
            l = i + q->d.num_pixels_x * j;
    }
    if (!q->include_sed && l < 0) return;

    /* the packet's redshifted wavelength, FluxRecorder.cpp:309-310 */
    int ell = wlg_bin(&e->wlg[q->d.wavelength_grid].g, ppp->lambda * (1. + q->d.redshift));
    if (ell < 0) return;

    double L = ppp->W / ppp->lambda;
    double tau;
    if (ppp->has_tau)
        tau = ppp->tau_obs;
    else
    {
        tau = get_extinction_optical_depth(e, ppp);
        ppp->tau_obs = tau;
        ppp->has_tau = 1;
    }
    double Lext = L * exp(-tau);
    e->cnt.detections++;

    for (int pass = 0; pass < 2; ++pass)
    {
        double** arrays = pass == 0 ? q->sed : q->ifu;
        size_t index;
        if (pass == 0)
        {
            if (!q->include_sed) continue;
            index = (size_t)ell;
        }
        else
        {
            if (!q->include_ifu || l < 0) continue;
            index = (size_t)l + (size_t)ell * q->npix;
        }
        if (q->record_total_only)
            arrays[SK_COMP_TOTAL][index] += Lext;
        else if (ppp->primary_origin)
        {
            if (ppp->nscatt == 0)
            {
                arrays[SK_COMP_TRANSPARENT][index] += L;
                arrays[SK_COMP_PRIMARY_DIRECT][index] += Lext;
            }
            else
            {
                arrays[SK_COMP_PRIMARY_SCATTERED][index] += Lext;
                if (ppp->nscatt <= q->d.num_scattering_levels)
                    arrays[SK_COMP_PRIMARY_SCATTERED_LEVEL + ppp->nscatt - 1][index] += Lext;
            }
        }
        else
        {
            if (ppp->nscatt == 0)
            {
                arrays[SK_COMP_SECONDARY_TRANSPARENT][index] += L;
                arrays[SK_COMP_SECONDARY_DIRECT][index] += Lext;
            }
            else
                arrays[SK_COMP_SECONDARY_SCATTERED][index] += Lext;
        }
    }
    if (q->d.record_statistics && q->include_sed)
    {
        int i = 0;
        while (i < q->hist_nsed && q->hist_sed_ell[i] != ell) i++;
        if (i == q->hist_nsed)
        {
            if (q->hist_nsed == q->hist_capsed)
            {
                q->hist_capsed = q->hist_capsed ? 2 * q->hist_capsed : 16;
                q->hist_sed_ell = (int*)realloc(q->hist_sed_ell, q->hist_capsed * sizeof(int));
                q->hist_sed_w = (double*)realloc(q->hist_sed_w, q->hist_capsed * sizeof(double));
            }
            q->hist_sed_ell[i] = ell;
            q->hist_sed_w[i] = 0.;
            q->hist_nsed++;
        }
        q->hist_sed_w[i] += Lext;
    }
    if (q->d.record_statistics && q->include_ifu && l >= 0)
    {
        size_t lell = (size_t)l + (size_t)ell * q->npix;
        int i = 0;
        while (i < q->hist_npix && q->hist_lell[i] != lell) i++;
        if (i == q->hist_npix)
        {
            if (q->hist_npix == q->hist_cappix)
            {
                q->hist_cappix = q->hist_cappix ? 2 * q->hist_cappix : 64;
                q->hist_lell = (size_t*)realloc(q->hist_lell, q->hist_cappix * sizeof(size_t));
                q->hist_wpix = (double*)realloc(q->hist_wpix, q->hist_cappix * sizeof(double));
            }
            q->hist_lell[i] = lell;
            q->hist_wpix[i] = 0.;
            q->hist_npix++;
        }
        q->hist_wpix[i] += Lext;
    }
}

/* MonteCarloSimulation::peelOffEmission, MonteCarloSimulation.cpp:617-634; PhotonPacket::launchEmissionPeelOff,
 * PhotonPacket.cpp:66-85 (isotropic emission: no angular bias) */
static void peel_off_emission(sko_engine_t* e, const packet_t* pp)
{
    packet_t ppp;
    memset(&ppp, 0, sizeof ppp);
    for (int j = 0; j < e->ninstr; ++j)
    {
        instr_t* q = &e->instr[j];
        if (!q->same_as_preceding)
        {
            ppp = *pp;
            ppp.nscatt = 0;
            memcpy(ppp.k, q->kobs, sizeof ppp.k);
            ppp.has_tau = 0;
            if (pp->has_vsrc) /* PhotonPacket.cpp:77 */
            {
                ppp.lambda = shifted_emission_wavelength(pp->lambda0, q->kobs, pp->vsrc);
                ppp.ilam = index_for_lambda(e, ppp.lambda);
            }
        }
        detect(e, q, &ppp);
    }
}

/* MonteCarloSimulation::peelOffScattering (consolidated branch), MonteCarloSimulation.cpp:784-842;
 * MediumSystem::peelOffScattering, MediumSystem.cpp:734-767; DustMix::peeloffScattering HG branch,
 * DustMix.cpp:430-445; PhotonPacket::launchScatteringPeelOff, PhotonPacket.cpp:89-103 */
static void peel_off_scattering(sko_engine_t* e, const packet_t* pp)
{
    const double glarge = 0.95;
    packet_t ppp;
    memset(&ppp, 0, sizeof ppp);
    for (int j = 0; j < e->ninstr; ++j)
    {
        instr_t* q = &e->instr[j];
        if (!q->same_as_preceding)
        {
            double costheta = pp->k[0] * q->kobs[0] + pp->k[1] * q->kobs[1] + pp->k[2] * q->kobs[2];
            /* MediumSystem::weightsForScattering (MediumSystem.cpp:697-730): 1 for a single medium, else the scattering
               opacities of the components in the interaction cell, normalised; none scatters: no peel-off at all
               (MonteCarloSimulation.cpp:790-792) */
            /* the wavelength the interaction cell perceives, MediumSystem::perceivedWavelengthForScattering
               (MediumSystem.cpp:667-674) */
            const double lamp = perceived_wavelength(e, pp->lambda, pp->k, pp->m_int);
            const int ilp = e->kin ? index_for_lambda(e, lamp) : pp->ilam;
            double wv[SK_MAX_MEDIA] = {1., 0., 0., 0.};
            if (e->nmed > 1)
            {
                double sum = 0.;
                for (int h = 0; h < e->nmed; ++h)
                {
                    wv[h] = e->dens[(size_t)h * e->ncells + pp->m_int] * e->sig_sca[(size_t)h * e->nlam + ilp];
                    sum += wv[h];
                }
                if (!(sum > 0.)) return;
                for (int h = 0; h < e->nmed; ++h) wv[h] /= sum;
            }
            double I = 0.;
            for (int h = 0; h < e->nmed; ++h)
                if (wv[h] > 0.)
                {
                    double g = e->gpar[(size_t)h * e->nlam + ilp];
                    double value = fabs(g) > glarge ? mean_hg(g, costheta) : value_hg(g, costheta);
                    I += value * wv[h]; /* MediumSystem.cpp:745-754 */
                }
            ppp = *pp;
            if (e->kin)
            {
                /* PhotonPacket::launchScatteringPeelOff (PhotonPacket.cpp:89-103): the perceived wavelength, shifted by the
                   bulk velocity of the cell for the direction towards the observer */
                ppp.lambda = lamp;
                if (e->vel && pp->m_int >= 0)
                {
                    const double* v = e->vel + 3 * (size_t)pp->m_int;
                    if (v[0] != 0. || v[1] != 0. || v[2] != 0.) ppp.lambda = shifted_emission_wavelength(lamp, q->kobs, v);
                }
                ppp.ilam = index_for_lambda(e, ppp.lambda);
            }
            ppp.W = pp->W * I;
            ppp.nscatt = pp->nscatt + 1;
            memcpy(ppp.k, q->kobs, sizeof ppp.k);
            ppp.has_tau = 0;
        }
        detect(e, q, &ppp);
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* sources                                                                                          */
/* ------------------------------------------------------------------------------------------------ */

/* Random::cdfLogLog, Random.cpp:210-216 */
static double sample_cdf_loglog(rng_t* g, const double* xv, const double* pv, const double* Pv, int n)
{
    double X = uniform(g);
    int i = locate_clip(Pv, n, X);
    double alpha = log(pv[i + 1] / pv[i]) / log(xv[i + 1] / xv[i]);
    return xv[i] * gexp(-alpha, (X - Pv[i]) / (pv[i] * xv[i]));
}
/* Random::cdfLinLin, Random.cpp:201-206 */
static double sample_cdf_linlin(rng_t* g, const double* xv, const double* Pv, int n)
{
    double X = uniform(g);
    int i = locate_clip(Pv, n, X);
    return interp_linlin(X, Pv[i], Pv[i + 1], xv[i], xv[i + 1]);
}

/* ContSED::specificLuminosity: BlackBodySED.cpp:38-41 / TabulatedSED.cpp:46-49 */
static double specific_luminosity(const sk_source_t* s, double lambda)
{
    if (s->sed_kind == SK_SED_BLACKBODY) return planck(lambda, s->sed_temperature) / s->sed_norm;
    int i = locate_fail(s->sed_lambda, s->sed_n, lambda);
    if (i < 0) return 0.;
    return interp_loglog(lambda, s->sed_lambda[i], s->sed_lambda[i + 1], s->sed_p[i], s->sed_p[i + 1]);
}

/* ExpDiskGeometry::randomCylRadius / randomZ, ExpDiskGeometry.cpp:46-68 */
static double expdisk_random_R(rng_t* g, double hR, double Rmin, double Rmax)
{
    double R, X;
    do
    {
        X = uniform(g);
        R = hR * (-1.0 - lambert_w1((X - 1.0) / M_E));
    } while ((Rmax > 0.0 && R >= Rmax) || R <= Rmin);
    return R;
}
static double expdisk_random_z(rng_t* g, double hz, double zmax)
{
    double z, X;
    do
    {
        X = uniform(g);
        z = (X <= 0.5) ? hz * log(2.0 * X) : -hz * log(2.0 * (1.0 - X));
    } while (zmax > 0.0 && fabs(z) >= zmax);
    return z;
}

/* Geometry::generatePosition for the supported geometries */
static void generate_position(rng_t* g, const sk_source_t* s, double r[3])
{
    const double* p = s->geom_params;
    switch (s->geometry)
    {
        case SK_GEOM_SHELL:
        {
            /* ShellGeometry::randomRadius (ShellGeometry.cpp:43-57) + SpheGeometry::generatePosition (SpheGeometry.cpp:26-33) */
            double pe = p[2], smin = p[3], sdiff = p[4], tmin = p[5], tmax = p[6];
            double X = uniform(g);
            double rad;
            if (fabs(pe - 3.0) < 1e-2)
            {
                double sv = smin + X * sdiff;
                rad = gexp(pe - 2.0, sv);
            }
            else
            {
                double zz = (1.0 - X) * tmin + X * tmax;
                rad = pow(zz, 1.0 / (3.0 - pe));
            }
            double k[3];
            random_direction(g, k);
            r[0] = rad * k[0];
            r[1] = rad * k[1];
            r[2] = rad * k[2];
            break;
        }
        case SK_GEOM_EXPDISK:
        {
            /* SepAxGeometry::generatePosition, SepAxGeometry.cpp:12-20 */
            double R = expdisk_random_R(g, p[0], p[2], p[3]);
            double phi = 2.0 * M_PI * uniform(g);
            double z = expdisk_random_z(g, p[1], p[4]);
            r[0] = R * cos(phi);
            r[1] = R * sin(phi);
            r[2] = z;
            break;
        }
        case SK_GEOM_RING:
        {
            /* RingGeometry::randomCylRadius / randomZ, RingGeometry.cpp:56-68 */
            double R = sample_cdf_linlin(g, s->geom_table_x, s->geom_table_P, s->geom_table_n);
            double phi = 2.0 * M_PI * uniform(g);
            double X = uniform(g);
            double z = (X <= 0.5) ? p[2] * log(2.0 * X) : -p[2] * log(2.0 * (1.0 - X));
            r[0] = R * cos(phi);
            r[1] = R * sin(phi);
            r[2] = z;
            break;
        }
        case SK_GEOM_SPIRAL_EXPDISK:
        {
            /* SpiralStructureGeometryDecorator::generatePosition / perturbation,
               SpiralStructureGeometryDecorator.cpp:33-45,72-76 */
            double R0 = expdisk_random_R(g, p[0], p[2], p[3]);
            double phi0 = 2.0 * M_PI * uniform(g);
            double z = expdisk_random_z(g, p[1], p[4]);
            double x0 = R0 * cos(phi0), y0 = R0 * sin(phi0);
            double R = sqrt(x0 * x0 + y0 * y0); /* Position::cylindrical, Position.cpp:104-109 */
            double m = p[5], tanp = p[6], Rz = p[7], phiz = p[8], w = p[9], N = p[10], cn = p[11];
            double c = 1.0 + (cn - 1.0) * w;
            double phi, t;
            do
            {
                phi = 2.0 * M_PI * uniform(g);
                double gamma = log(R / Rz) / tanp + phiz + 0.5 * M_PI / m;
                double perturbation = (1.0 - w) + w * cn * pow_even(sin(0.5 * m * (gamma - phi)), 2 * N);
                t = uniform(g) * c / perturbation;
            } while (t > 1);
            r[0] = R * cos(phi);
            r[1] = R * sin(phi);
            r[2] = z;
            break;
        }
        default: r[0] = r[1] = r[2] = 0.; break;
    }
}

/* the bulk velocity of a source at the launch position: PointSource velocityX/Y/Z, or GeometricSource::velocityMagnitude()
 * times the vector field (GeometricSource.cpp:73-79; RadialVectorField.cpp:19-37, CylindricalVectorField.cpp:19-38) */
static void source_velocity(const sk_source_t* s, const double r[3], double v[3])
{
    v[0] = v[1] = v[2] = 0.;
    if (s->velocity_kind == SK_VEL_CONSTANT)
    {
        v[0] = s->velocity[0];
        v[1] = s->velocity[1];
        v[2] = s->velocity[2];
    }
    else if (s->velocity_kind == SK_VEL_RADIAL || s->velocity_kind == SK_VEL_CYLINDRICAL)
    {
        const double mag = s->velocity[0], unity = s->velocity[1], expon = s->velocity[2];
        double u[3];
        if (s->velocity_kind == SK_VEL_RADIAL)
        {
            u[0] = r[0];
            u[1] = r[1];
            u[2] = r[2];
        }
        else
        {
            u[0] = -r[1];
            u[1] = r[0];
            u[2] = 0.;
        }
        double rr = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        if (rr == 0.) return;
        u[0] /= rr;
        u[1] /= rr;
        u[2] /= rr;
        double f = 1.;
        if (unity > 0.)
            if ((expon > 0. && rr < unity) || (expon < 0. && rr > unity)) f = pow(rr / unity, expon);
        v[0] = mag * (f * u[0]);
        v[1] = mag * (f * u[1]);
        v[2] = mag * (f * u[2]);
    }
}

/* SourceSystem::launch (SourceSystem.cpp:101-113), NormalizedSource::launch (NormalizedSource.cpp:73-110),
 * GeometricSource::launchNormalized (GeometricSource.cpp:66-82), PointSource::launchSpecialty (PointSource.cpp:32-42),
 * PhotonPacket::launch (PhotonPacket.cpp:18-40) */
static void launch_primary(sko_engine_t* e, rng_t* g, uint64_t history, packet_t* pp)
{
    /* std::upper_bound(_Iv, historyIndex) - 1 */
    int lo = 0, hi = e->nsrc + 1;
    while (lo < hi)
    {
        int mid = (lo + hi) >> 1;
        if (history < e->Iv[mid])
            hi = mid;
        else
            lo = mid + 1;
    }
    int h = lo - 1;
    const sk_source_t* s = &e->src[h].s;
    double weight = e->Lv[h] / e->Wv[h];
    double L = e->Lpp * weight;

    double lambda, w;
    double xi = s->wavelength_bias;
    if (!xi)
    {
        lambda = sample_cdf_loglog(g, s->sed_lambda, s->sed_p, s->sed_P, s->sed_n);
        w = 1.;
    }
    else
    {
        if (uniform(g) > xi)
            lambda = sample_cdf_loglog(g, s->sed_lambda, s->sed_p, s->sed_P, s->sed_n);
        else if (s->bias_kind == SK_BIAS_OLIGO)
        {
            size_t index = (size_t)(uniform(g) * s->oligo_n); /* OligoWavelengthDistribution.cpp:34-38 */
            lambda = s->oligo_lambda[index];
        }
        else
        {
            double logMin = log(s->bias_min);
            double logWidth = log(s->bias_max) - log(s->bias_min);
            lambda = exp(logMin + logWidth * uniform(g)); /* DefaultWavelengthDistribution.cpp:37-40 */
        }
        double sl = specific_luminosity(s, lambda);
        if (!sl)
            w = 0.;
        else
        {
            double b;
            if (s->bias_kind == SK_BIAS_OLIGO)
                b = s->oligo_probability;
            else
            {
                double logWidth = log(s->bias_max) - log(s->bias_min);
                /* Range::containsFuzzy, SKIRT/utils/Range.hpp:56 */
                if (lambda >= s->bias_min * (1 - 1e-14) && lambda <= s->bias_max * (1 + 1e-14))
                    b = 1. / (logWidth * lambda);
                else
                    b = 0.;
            }
            w = sl / ((1 - xi) * sl + xi * b);
        }
    }
    double Lw = L * w;
    if (s->kind == SK_SRC_POINT)
    {
        pp->r[0] = s->position[0];
        pp->r[1] = s->position[1];
        pp->r[2] = s->position[2];
    }
    else
        generate_position(g, s, pp->r);
    random_direction(g, pp->k);
    pp->lambda = lambda;
    pp->W = Lw * lambda;
    pp->nscatt = 0;
    pp->primary_origin = 1;
    pp->history = history;
    pp->has_tau = 0;
    pp->lambda0 = lambda;
    pp->has_vsrc = 0;
    if (s->velocity_kind != SK_VEL_NONE)
    {
        /* PhotonPacket::launch with a velocity interface (PhotonPacket.cpp:33): the wavelength is Doppler-shifted for the
           launch direction, the weight keeps the rest-frame wavelength */
        source_velocity(s, pp->r, pp->vsrc);
        pp->has_vsrc = 1;
        pp->lambda = shifted_emission_wavelength(lambda, pp->k, pp->vsrc);
    }
    pp->ilam = index_for_lambda(e, pp->lambda);
}

/* SecondarySourceSystem::launch (SecondarySourceSystem.cpp:130-142) + DustSecondarySource::launch
 * (DustSecondarySource.cpp:511-581) without velocities and polarisation; cell box from the grid for
 * SpatialGrid::randomPositionInCell (TreeSpatialGrid.cpp:125-128, CartesianSpatialGrid.cpp:80-83) + Random::position
 * (Random.cpp:168-176) + Box::fracPos */
static void cell_box(const sko_engine_t* e, int m, double b[6])
{
    if (e->grid_kind == 1)
    {
        int k = m % e->nz, j = (m / e->nz) % e->ny, i = m / (e->nz * e->ny);
        b[0] = e->xv[i];
        b[1] = e->yv[j];
        b[2] = e->zv[k];
        b[3] = e->xv[i + 1];
        b[4] = e->yv[j + 1];
        b[5] = e->zv[k + 1];
    }
    else
        memcpy(b, e->node_box + 6 * (size_t)e->node_of_cell[m], 6 * sizeof(double));
}

static void launch_secondary(sko_engine_t* e, rng_t* g, uint64_t history, packet_t* pp)
{
    const int M = e->ncells, nem = e->sec_nem;
    /* std::upper_bound(_Iv, historyIndex) - 1 */
    int lo = 0, hi = M + 1;
    while (lo < hi)
    {
        int mid = (lo + hi) >> 1;
        if (history < e->sec_Iv[mid])
            hi = mid;
        else
            lo = mid + 1;
    }
    int m = lo - 1; /* launch order = cell order for the AllCellsLibrary */
    double ws = e->sec_Lv[m] / e->sec_Wv[m];
    const double* xv = e->sec_lambda;
    const double* pv = e->sec_pv + (size_t)m * nem;
    const double* Pv = e->sec_Pv + (size_t)m * nem;
    double lambda, w;
    double xi = e->sec.wavelength_bias;
    if (!xi)
    {
        lambda = sample_cdf_loglog(g, xv, pv, Pv, nem);
        w = 1.;
    }
    else
    {
        double logMin = log(e->sec.bias_min);
        double logWidth = log(e->sec.bias_max) - log(e->sec.bias_min);
        if (uniform(g) > xi)
            lambda = sample_cdf_loglog(g, xv, pv, Pv, nem);
        else
            lambda = exp(logMin + logWidth * uniform(g)); /* DefaultWavelengthDistribution.cpp:37-40 */
        /* NR::value<interpolateLogLog>(lambda, _lambdav, _pv), NR.hpp:372-378 */
        double sl = 0.;
        int i = locate_fail(xv, nem, lambda);
        if (i >= 0) sl = interp_loglog(lambda, xv[i], xv[i + 1], pv[i], pv[i + 1]);
        if (!sl)
            w = 0.;
        else
        {
            double b;
            if (lambda >= e->sec.bias_min * (1 - 1e-14) && lambda <= e->sec.bias_max * (1 + 1e-14))
                b = 1. / (logWidth * lambda);
            else
                b = 0.;
            w = sl / ((1 - xi) * sl + xi * b);
        }
    }
    if (e->grid_kind == 3)
    {
        /* VoronoiMeshSnapshot::generatePosition(m), VoronoiMeshSnapshot.cpp:976-989: random points in the cell's enclosing
         * box until one is closest to site m among the sites of m's neighbours (isPointClosestTo, .cpp:848-856) */
        const double* box = e->vbox + 6 * (size_t)m;
        const double* sm = e->vsite + 3 * (size_t)m;
        int found = 0;
        for (int it = 0; it < 10000 && !found; ++it)
        {
            double ux = uniform(g), uy = uniform(g), uz = uniform(g);
            double x = box[0] + ux * (box[3] - box[0]);
            double y = box[1] + uy * (box[4] - box[1]);
            double z = box[2] + uz * (box[5] - box[2]);
            double dx = x - sm[0], dy = y - sm[1], dz = z - sm[2];
            double target = dx * dx + dy * dy + dz * dz;
            found = 1;
            for (int64_t i = e->vnbr_off[m]; i < e->vnbr_off[m + 1]; ++i)
            {
                int id = e->vnbr[i];
                if (id < 0) continue;
                const double* t = e->vsite + 3 * (size_t)id;
                double ex = x - t[0], ey = y - t[1], ez = z - t[2];
                if (ex * ex + ey * ey + ez * ez < target)
                {
                    found = 0;
                    break;
                }
            }
            pp->r[0] = x;
            pp->r[1] = y;
            pp->r[2] = z;
        }
        if (!found) /* the reference throws a fatal error here; emit from the site */
        {
            pp->r[0] = sm[0];
            pp->r[1] = sm[1];
            pp->r[2] = sm[2];
        }
    }
    else
    {
        double box[6];
        cell_box(e, m, box);
        double ux = uniform(g), uy = uniform(g), uz = uniform(g);
        pp->r[0] = box[0] + ux * (box[3] - box[0]); /* Box::fracPos, Box.hpp */
        pp->r[1] = box[1] + uy * (box[4] - box[1]);
        pp->r[2] = box[2] + uz * (box[5] - box[2]);
    }
    random_direction(g, pp->k);
    double L = e->sec_Lpp * 1.; /* _Lv[s]/_Wv[s] = 1 for the single secondary source */
    pp->lambda = lambda;
    pp->W = (L * ws * w) * lambda;
    pp->nscatt = 0;
    pp->primary_origin = 0;
    pp->history = history;
    pp->has_tau = 0;
    pp->lambda0 = lambda;
    pp->has_vsrc = 0;
    if (e->vel)
    {
        /* the bulk velocity of the emitting cell, if it is nonzero (DustSecondarySource.cpp:271, 562-563) */
        const double* v = e->vel + 3 * (size_t)m;
        if (v[0] != 0. || v[1] != 0. || v[2] != 0.)
        {
            memcpy(pp->vsrc, v, sizeof pp->vsrc);
            pp->has_vsrc = 1;
            pp->lambda = shifted_emission_wavelength(lambda, pp->k, pp->vsrc);
        }
    }
    pp->ilam = index_for_lambda(e, pp->lambda);
}

/* ------------------------------------------------------------------------------------------------ */
/* the life cycle                                                                                   */
/* ------------------------------------------------------------------------------------------------ */

/* MonteCarloSimulation::simulateForcedPropagation, MonteCarloSimulation.cpp:696-742 */
static void simulate_forced_propagation(sko_engine_t* e, rng_t* g, packet_t* pp, int* m_int)
{
    double taupath = e->nsegs ? e->segs[e->nsegs - 1].tau : 0.;
    if (taupath <= 0.)
    {
        pp->W *= 0.;
        *m_int = -1;
        return;
    }
    double xi = e->cfg.path_length_bias;
    double tau = 0.;
    if (xi == 0.)
        tau = expon_cutoff(g, taupath);
    else
    {
        tau = uniform(g) < xi ? uniform(g) * taupath : expon_cutoff(g, taupath);
        double p = -exp(-tau) / expm1(-taupath);
        double q = (1.0 - xi) * p + xi / taupath;
        double weight = p / q;
        pp->W *= weight;
    }
    int m;
    double s, tauAbs;
    find_interaction_point(e, tau, &m, &s, &tauAbs);
    if (e->cfg.explicit_absorption)
        pp->W *= -expm1(-taupath) * exp(-tauAbs); /* MonteCarloSimulation.cpp:727-731 */
    else
    {
        /* MediumSystem::albedoForScattering, MediumSystem.cpp:678-693 (single dust medium: ksca/kext) */
        double albedo = 0.;
        if (m >= 0)
        {
            const int il = ilam_in_cell(e, pp, m); /* perceivedWavelengthForScattering */
            double ksca = opacity_sum(e, e->sig_sca, il, m);
            double kext = opacity_sum(e, e->sig_ext, il, m);
            albedo = kext > 0. ? ksca / kext : 0.;
        }
        pp->W *= -expm1(-taupath) * albedo;
    }
    pp->m_int = m;
    /* PhotonPacket::propagate, PhotonPacket.cpp:107-111 */
    pp->r[0] += s * pp->k[0];
    pp->r[1] += s * pp->k[1];
    pp->r[2] += s * pp->k[2];
    *m_int = m;
}

/* MonteCarloSimulation::simulateNonForcedPropagation (MonteCarloSimulation.cpp:746-780) with
 * MediumSystem::setInteractionPointUsingExtinction (MediumSystem.cpp:978-1010) */
static int simulate_nonforced_propagation(sko_engine_t* e, rng_t* g, packet_t* pp)
{
    double tauinteract = -log(uniform(g)); /* Random::expon, Random.cpp:98-101 */
    gen_t gen;
    gen_start(&gen, pp->r, pp->k);
    double tau = 0., s = 0., tauabs = 0.;
    /* with explicit absorption the walk is in scattering optical depth, setInteractionPointUsingScatteringAndAbsorption
       (MediumSystem.cpp:1075-1110; several media :1112-1150: the absorption optical depth is accumulated next to it and
       interpolated at the interaction point), and the weight is the absorption along the way instead of the albedo
       (MonteCarloSimulation.cpp:757-762) */
    const int explicit_abs = e->cfg.explicit_absorption;
    const double* section = explicit_abs ? e->sig_sca : e->sig_ext; /* (several media: MediumSystem.cpp:1012-1040) */
    e->cnt.forward_paths++;
    while (gen_next(e, &gen))
    {
        e->cnt.forward_segments++;
        double tau0 = tau, s0 = s, tauabs0 = tauabs;
        double ds = gen.ds;
        int m = gen.m;
        const int il = m >= 0 ? ilam_in_cell(e, pp, m) : pp->ilam; /* (kinematics: MediumSystem.cpp:1042-1070, 1152-1188) */
        if (m >= 0 && explicit_abs && (e->nmed > 1 || e->kin))
            for (int h = 0; h < e->nmed; ++h)
            {
                double ns = e->dens[(size_t)h * e->ncells + m] * ds;
                tau += e->sig_sca[(size_t)h * e->nlam + il] * ns;
                tauabs += e->sig_abs[(size_t)h * e->nlam + il] * ns;
            }
        else if (m >= 0)
            for (int h = 0; h < e->nmed; ++h)
                tau += section[(size_t)h * e->nlam + il] * e->dens[(size_t)h * e->ncells + m] * gen.ds;
        s += ds;
        if (tauinteract < tau)
        {
            double sint = interp_linlin(tauinteract, tau0, tau, s0, s);
            double ksca = opacity_sum(e, e->sig_sca, il, m);
            double kext = opacity_sum(e, e->sig_ext, il, m);
            double albedo = kext > 0. ? ksca / kext : 0.;
            pp->m_int = m;
            if (explicit_abs)
                albedo = (e->nmed > 1 || e->kin) ? exp(-interp_linlin(tauinteract, tau0, tau, tauabs0, tauabs))
                                     : exp(-(tauinteract * e->sig_abs[pp->ilam] / e->sig_sca[pp->ilam]));
            pp->W *= albedo;
            pp->r[0] += sint * pp->k[0];
            pp->r[1] += sint * pp->k[1];
            pp->r[2] += sint * pp->k[2];
            return 1;
        }
    }
    return 0;
}

/* MediumSystem::simulateScattering (MediumSystem.cpp:796-823), DustMix::performScattering HG branch
 * (DustMix.cpp:496-511), PhotonPacket::scatter (PhotonPacket.cpp:115-122) */
static void simulate_scattering(sko_engine_t* e, rng_t* g, packet_t* pp)
{
    /* select a medium component within the cell: NR::cdf over the scattering opacities (NR.hpp:446-463: cumulative sums
       divided by the total, first element zero) and NR::locateClip of one uniform deviate, MediumSystem.cpp:805-818 */
    int hsel = 0;
    const double lamp = perceived_wavelength(e, pp->lambda, pp->k, pp->m_int); /* MediumSystem.cpp:802 */
    const int ilp = e->kin ? index_for_lambda(e, lamp) : pp->ilam;
    if (e->nmed > 1)
    {
        double Xv[SK_MAX_MEDIA + 1];
        Xv[0] = 0.;
        for (int h = 0; h < e->nmed; ++h)
            Xv[h + 1] = Xv[h] + e->dens[(size_t)h * e->ncells + pp->m_int] * e->sig_sca[(size_t)h * e->nlam + ilp];
        for (int h = 0; h <= e->nmed; ++h) Xv[h] /= Xv[e->nmed];
        hsel = locate_clip(Xv, e->nmed + 1, uniform(g));
    }
    double gp = e->gpar[(size_t)hsel * e->nlam + ilp];
    double knew[3];
    if (fabs(gp) < 1e-6)
        random_direction(g, knew);
    else
    {
        double f = ((1.0 - gp) * (1.0 + gp)) / (1.0 - gp + 2.0 * gp * uniform(g));
        double costheta = (1.0 + gp * gp - f * f) / (2.0 * gp);
        random_direction_about(g, pp->k, costheta, knew);
    }
    pp->nscatt++;
    memcpy(pp->k, knew, sizeof knew);
    if (e->kin)
    {
        /* PhotonPacket::scatter(bfk, bfv, lambda), PhotonPacket.cpp:115-122: the packet leaves at the perceived wavelength,
           shifted by the bulk velocity of the cell for its new direction */
        pp->lambda = lamp;
        if (e->vel && pp->m_int >= 0)
        {
            const double* v = e->vel + 3 * (size_t)pp->m_int;
            if (v[0] != 0. || v[1] != 0. || v[2] != 0.) pp->lambda = shifted_emission_wavelength(lamp, knew, v);
        }
        pp->ilam = index_for_lambda(e, pp->lambda);
    }
    e->cnt.scatterings++;
}

/* MonteCarloSimulation::performLifeCycle for one history index, MonteCarloSimulation.cpp:538-613 */
static void life_cycle(sko_engine_t* e, uint64_t history, int primary, int peel, int store, uint32_t stream_id)
{
    rng_t g;
    rng_init(&g, e->cfg.seed, stream_id, history);
    packet_t pp;
    memset(&pp, 0, sizeof pp);
    if (primary)
        launch_primary(e, &g, history, &pp);
    else
        launch_secondary(e, &g, history, &pp);
    if (pp.W / pp.lambda > 0)
    {
        e->cnt.packets++;
        if (peel) peel_off_emission(e, &pp);
        if (e->cfg.force_scattering)
        {
            double Lthreshold = (pp.W / pp.lambda) / e->cfg.min_weight_reduction;
            int minScattEvents = e->cfg.min_scatt_events;
            while (1)
            {
                set_extinction_optical_depths(e, &pp);
                if (store) store_radiation_field(e, primary, &pp);
                int m;
                simulate_forced_propagation(e, &g, &pp, &m);
                double L = pp.W / pp.lambda;
                if (L <= 0 || (L <= Lthreshold && pp.nscatt >= minScattEvents)) break;
                if (peel) peel_off_scattering(e, &pp);
                simulate_scattering(e, &g, &pp);
            }
        }
        else
        {
            while (1)
            {
                if (!simulate_nonforced_propagation(e, &g, &pp)) break;
                if (pp.W / pp.lambda <= 0) break;
                if (peel) peel_off_scattering(e, &pp);
                simulate_scattering(e, &g, &pp);
            }
        }
    }
    /* per-history statistics: FluxRecorder::detect (FluxRecorder.cpp:457-466) flushes when the history changes */
    for (int j = 0; j < e->ninstr; ++j) flush_history_stats(&e->instr[j]);
}

int sko_run_segment(sko_engine_t* e, uint64_t first, uint64_t count, int32_t primary, int32_t peel, int32_t store,
                    uint32_t stream_id)
{
    if (!e) return fail(SK_ERR_INVALID, "null engine");
    if (!e->grid_kind || !e->dens || !e->nlam) return fail(SK_ERR_STATE, "engine is not fully configured");
    if (primary && !e->nsrc) return fail(SK_ERR_STATE, "engine is not fully configured");
    if (primary && !e->npackets) return fail(SK_ERR_STATE, "call prepare_primary first");
    if (!primary && !e->secondary_ready) return fail(SK_ERR_STATE, "call prepare_secondary first");
    if (e->nmed != e->nmix) return fail(SK_ERR_STATE, "the number of dust mixes does not match the number of medium components");
    if (!primary && e->sec_nmed != e->nmed) return fail(SK_ERR_STATE, "emission tables do not match the medium components");
    if (store && e->rf_grid < 0) return fail(SK_ERR_STATE, "no radiation field grid configured");
    if (store && !e->cfg.force_scattering)
        return fail(SK_ERR_INVALID, "storing the radiation field requires forced scattering (Configuration.cpp:476-482)");
    for (uint64_t i = 0; i != count; ++i)
    {
        /* sk_engine_set_history_interleave: every num_parts-th block of `block` histories */
        if (e->il_parts > 1 && (i / e->il_block) % e->il_parts != e->il_part) continue;
        life_cycle(e, first + i, primary, peel, store, stream_id);
    }
    return SK_OK;
}

int sko_set_history_interleave(sko_engine_t* e, uint64_t block, uint32_t num_parts, uint32_t part)
{
    if (!e) return fail(SK_ERR_INVALID, "null engine");
    if (num_parts < 1 || part >= num_parts) return fail(SK_ERR_INVALID, "part must be below num_parts");
    if (num_parts > 1 && (block == 0 || (block & (block - 1)))) return fail(SK_ERR_INVALID, "block must be a power of two");
    e->il_block = block;
    e->il_parts = num_parts;
    e->il_part = part;
    return SK_OK;
}

/* MediumSystem::communicateRadiationField (single process), MediumSystem.cpp:1304-1313 */
int sko_communicate_rf(sko_engine_t* e, int32_t primary)
{
    if (!e) return fail(SK_ERR_INVALID, "null engine");
    if (!primary && e->rf2) memcpy(e->rf2, e->rf2c, (size_t)e->ncells * e->nrf * sizeof(double));
    return SK_OK;
}

/* MediumSystem::totalDustAbsorbedLuminosity, MediumSystem.cpp:1317-1356 (single dust medium) */
int sko_absorbed_luminosity(sko_engine_t* e, int32_t primary, double* out)
{
    if (!e || !out) return fail(SK_ERR_INVALID, "null argument");
    if (e->rf_grid < 0) return fail(SK_ERR_STATE, "no radiation field");
    const sk_wavelength_grid_t* g = &e->wlg[e->rf_grid].g;
    const double* rf = primary ? e->rf1 : e->rf2;
    double Labs = 0.;
    for (int m = 0; m < e->ncells; ++m)
    {
        double sum = 0.;
        for (int ell = 0; ell < e->nrf; ++ell)
        {
            int il = index_for_lambda(e, g->lambda[ell]);
            double opacity = 0.; /* MediumSystem::opacityAbs over the dust components */
            for (int h = 0; h < e->nmed; ++h) opacity += e->sig_abs[(size_t)h * e->nlam + il] * e->dens[(size_t)h * e->ncells + m];
            sum += opacity * rf[(size_t)m * e->nrf + ell];
        }
        Labs += sum;
    }
    *out = Labs;
    return SK_OK;
}

int sko_read_rf(sko_engine_t* e, int32_t which, double* out)
{
    if (!e || !out) return fail(SK_ERR_INVALID, "null argument");
    const double* src = which == 0 ? e->rf1 : which == 1 ? e->rf2 : e->rf2c;
    if (!src) return fail(SK_ERR_STATE, "no radiation field");
    memcpy(out, src, (size_t)e->ncells * e->nrf * sizeof(double));
    return SK_OK;
}

static int read_array(sko_engine_t* e, int instrument, int component, int ifu, double* out)
{
    if (!e || !out || instrument < 0 || instrument >= e->ninstr || component < 0 || component >= NUM_COMP)
        return fail(SK_ERR_INVALID, "bad instrument/component");
    instr_t* q = &e->instr[instrument];
    size_t len = ifu ? q->npix * q->nl : (size_t)q->nl;
    double** arrays = ifu ? q->ifu : q->sed;
    if (ifu ? !q->include_ifu : !q->include_sed) return fail(SK_ERR_INVALID, "instrument does not record this");
    if (component == SK_COMP_TOTAL && !q->record_total_only)
    {
        /* FluxRecorder::calibrateAndWrite: total = direct + scattered (+ secondary), FluxRecorder.cpp:540-570 */
        for (size_t i = 0; i < len; ++i)
        {
            double t = arrays[SK_COMP_PRIMARY_DIRECT][i] + arrays[SK_COMP_PRIMARY_SCATTERED][i];
            if (arrays[SK_COMP_SECONDARY_DIRECT])
                t += arrays[SK_COMP_SECONDARY_DIRECT][i] + arrays[SK_COMP_SECONDARY_SCATTERED][i];
            out[i] = t;
        }
        return SK_OK;
    }
    if (!arrays[component]) return fail(SK_ERR_INVALID, "component not recorded");
    memcpy(out, arrays[component], len * sizeof(double));
    return SK_OK;
}
int sko_read_sed(sko_engine_t* e, int32_t instrument, int32_t component, double* out)
{
    return read_array(e, instrument, component, 0, out);
}
int sko_read_ifu(sko_engine_t* e, int32_t instrument, int32_t component, double* out)
{
    return read_array(e, instrument, component, 1, out);
}
int sko_read_sed_stats(sko_engine_t* e, int32_t instrument, int32_t k, double* out)
{
    if (!e || !out || instrument < 0 || instrument >= e->ninstr || k < 0 || k > 4)
        return fail(SK_ERR_INVALID, "bad instrument/power");
    instr_t* q = &e->instr[instrument];
    if (!q->wsed[k]) return fail(SK_ERR_INVALID, "statistics not recorded");
    memcpy(out, q->wsed[k], (size_t)q->nl * sizeof(double));
    return SK_OK;
}
int sko_read_ifu_stats(sko_engine_t* e, int32_t instrument, int32_t k, double* out)
{
    if (!e || !out || instrument < 0 || instrument >= e->ninstr || k < 0 || k > 4)
        return fail(SK_ERR_INVALID, "bad instrument/power");
    instr_t* q = &e->instr[instrument];
    if (!q->wifu[k]) return fail(SK_ERR_INVALID, "statistics not recorded");
    memcpy(out, q->wifu[k], q->npix * (size_t)q->nl * sizeof(double));
    return SK_OK;
}
int sko_counters(sko_engine_t* e, sk_counters_t* out, int32_t reset)
{
    if (!e || !out) return fail(SK_ERR_INVALID, "null argument");
    *out = e->cnt;
    if (reset) memset(&e->cnt, 0, sizeof e->cnt);
    return SK_OK;
}

/* Counterpart of sk_engine_device_buffer for the radiation field tables (host memory here), so that the multi-rank tests
 * can all-reduce them in place exactly where the reference calls ProcessManager::sumToAll (MediumSystem.cpp:1304-1313). */
int sko_device_buffer(sko_engine_t* e, int32_t which, void** ptr, uint64_t* num_doubles)
{
    if (!e || !ptr || !num_doubles) return fail(SK_ERR_INVALID, "null argument");
    if (which < 0 || which > 2) return fail(SK_ERR_UNSUPPORTED, "the oracle exposes only the radiation field tables");
    *ptr = which == 0 ? e->rf1 : which == 1 ? e->rf2 : e->rf2c;
    *num_doubles = (uint64_t)e->ncells * e->nrf;
    return SK_OK;
}

/* test hooks: expose single building blocks so that unit tests can pin them one by one */
double sko_test_uniform(uint32_t seed, uint32_t stream, uint64_t history, int index)
{
    rng_t g;
    rng_init(&g, seed, stream, history);
    double u = 0.;
    for (int i = 0; i <= index; ++i) u = uniform(&g);
    return u;
}
double sko_test_lnmean(double x1, double x2)
{
    return lnmean4(x1, x2, log(x1), log(x2));
}
double sko_test_lambert_w1(double z)
{
    return lambert_w1(z);
}
double sko_test_mean_hg(double g, double costheta)
{
    return mean_hg(g, costheta);
}
/* traces one path from r along k and returns the number of segments; m/ds are written up to cap entries */
int sko_test_trace(sko_engine_t* e, const double r[3], const double k[3], int32_t* m, double* ds, int cap)
{
    gen_t g;
    gen_start(&g, r, k);
    int n = 0;
    while (gen_next(e, &g))
    {
        if (n < cap)
        {
            m[n] = g.m;
            ds[n] = g.ds;
        }
        n++;
    }
    return n;
}
/* nearest-site search of the Voronoi grid: the walk, and a brute-force scan to check it against */
int sko_test_voronoi_cell_index(sko_engine_t* e, const double r[3], int brute)
{
    if (!e || e->grid_kind != 3) return -2;
    if (!brute) return voronoi_cell_index(e, r[0], r[1], r[2], -1);
    if (!box_contains(e->extent, r[0], r[1], r[2])) return -1;
    int best = -1;
    double dbest = DBL_MAX;
    for (int m = 0; m < e->vcells; ++m)
    {
        double dx = r[0] - e->vsite[3 * m], dy = r[1] - e->vsite[3 * m + 1], dz = r[2] - e->vsite[3 * m + 2];
        double d = dx * dx + dy * dy + dz * dz;
        if (d < dbest)
        {
            dbest = d;
            best = m;
        }
    }
    return best;
}
/* the launch position of one history of the prepared secondary segment and the cell it was launched for */
int sko_test_secondary_launch(sko_engine_t* e, uint64_t history, uint32_t stream_id, double r[3])
{
    if (!e || !e->secondary_ready) return -2;
    rng_t g;
    rng_init(&g, (uint32_t)e->cfg.seed, stream_id, history);
    packet_t pp;
    launch_secondary(e, &g, history, &pp);
    memcpy(r, pp.r, 3 * sizeof(double));
    int lo = 0, hi = e->ncells + 1;
    while (lo < hi)
    {
        int mid = (lo + hi) >> 1;
        if (history < e->sec_Iv[mid])
            hi = mid;
        else
            lo = mid + 1;
    }
    return lo - 1;
}
void sko_test_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    philox4x32_10(c, key[0], key[1]);
    memcpy(out, c, sizeof c);
}
