// engine.cu -- host side of libskirt9_b200.so: the C ABI of include/sk_engine.h, the device memory layout
// builder (octree links / lattice tables) and the kernel launches.  sm_100a only; no CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <numeric>
#include <string>
#include <vector>

#include "sk_secondary.cuh"
#include "sk_setup.cuh"
#include "sk_wavefront.cuh"

// ---------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}
#define CK(call)                                                                                            \
    do                                                                                                      \
    {                                                                                                       \
        cudaError_t _e = (call);                                                                            \
        if (_e != cudaSuccess)                                                                              \
            return fail(SK_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e));                   \
    } while (0)

// SK_DEBUG_TIMING=1: host wall time of the phases of the set-up calls, to stderr
#include <chrono>
namespace
{
    struct PhaseTimer {
        const char* name;
        std::chrono::steady_clock::time_point t0;
        bool on;
        explicit PhaseTimer(const char* n) : name(n), t0(std::chrono::steady_clock::now()), on(getenv("SK_DEBUG_TIMING") != nullptr) {}
        void lap(const char* what)
        {
            if (!on) return;
            auto t1 = std::chrono::steady_clock::now();
            fprintf(stderr, "[sk timing] %s: %s %.2f ms\n", name, what, std::chrono::duration<double, std::milli>(t1 - t0).count());
            t0 = t1;
        }
    };
}

extern "C" const char* sk_last_error(void)
{
    return g_err.c_str();
}
extern "C" int sk_abi_version(void)
{
    return SK_ABI_VERSION;
}


// ---------------------------------------------------------------------------------------------------
// device memory: stream-ordered allocations from the device's default pool on the engine's stream.  A plain cudaFree
// synchronises the whole device and was measured at 0.4 s once the process holds a multi-GB bank; cudaFreeAsync returns
// the block to the pool (release threshold = never), where the next set_* call or the next engine finds it again.
// ---------------------------------------------------------------------------------------------------
static thread_local cudaStream_t g_stream = nullptr;  // stream of the engine the calling thread is working on
template <class T>
static cudaError_t dev_malloc(T** p, size_t bytes)
{
    return cudaMallocAsync((void**)p, bytes, g_stream);
}
static void dev_free(const void* p)
{
    if (p) cudaFreeAsync(const_cast<void*>(p), g_stream);
}

// ---------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------
// Octree neighbour links, one thread per cell: for the lattice point just across each wall find the deepest node of
// level <= the cell's level that holds it (replaces TreeNode::_neighbors built by OctTreeNode::addNeighbors,
// OctTreeNode.cpp:46-138)
__global__ void sk_build_links_kernel(const int32_t* __restrict__ first_child, const int32_t* __restrict__ node_child,
                                      const uint32_t* __restrict__ node_coord, const int32_t* __restrict__ node_of_cell,
                                      int ncells, int N, int maxlev, SkCellRec* __restrict__ cells,
                                      uint32_t* __restrict__ cell_coord)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= ncells) return;
    const int l = node_of_cell[m];
    const uint4 c = reinterpret_cast<const uint4*>(node_coord)[l];
    reinterpret_cast<uint4*>(cell_coord)[m] = c;
    const int lev = (int)c.w, size = N >> lev;
    SkCellRec r;
    r.dens = 0.;
    for (int w = 0; w < 6; ++w)
    {
        int x = (int)c.x, y = (int)c.y, z = (int)c.z;
        const int axis = w >> 1;
        int& t = axis == 0 ? x : axis == 1 ? y : z;
        t += (w & 1) ? size : -1;
        if (t < 0 || t >= N)
        {
            r.link[w] = -1;
            continue;
        }
        int node = 0, nlev = 0, nx = 0, ny = 0, nz = 0;
        while (first_child[node] >= 0 && nlev < lev)
        {
            const int half = N >> (nlev + 1);
            const int ch = ((x - nx) >= half ? 1 : 0) + ((y - ny) >= half ? 2 : 0) + ((z - nz) >= half ? 4 : 0);
            nx += (ch & 1) ? half : 0;
            ny += (ch & 2) ? half : 0;
            nz += (ch & 4) ? half : 0;
            node = first_child[node] + ch;
            nlev++;
        }
        const int nc = node_child[node];
        r.link[w] = nc < 0 ? ((nlev << SK_LINK_LEVEL_SHIFT) | (-(nc + 1)))
                           : (SK_LINK_INTERNAL | (nlev << SK_LINK_LEVEL_SHIFT) | nc);
    }
    cells[m] = r;
}
// device layout [ell][m] -> the reference's Table<2> layout [m][ell] (sk_engine_read_rf)
__global__ void sk_rf_transpose_kernel(const double* __restrict__ in, double* __restrict__ out, int ncells, int nrf)
{
    __shared__ double tile[32][33];
    const int m0 = blockIdx.x * 32, l0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y)
    {
        const int ell = l0 + j, m = m0 + threadIdx.x;
        if (ell < nrf && m < ncells) tile[j][threadIdx.x] = in[(size_t)ell * ncells + m];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y)
    {
        const int m = m0 + j, ell = l0 + threadIdx.x;
        if (ell < nrf && m < ncells) out[(size_t)m * nrf + ell] = tile[threadIdx.x][j];
    }
}
__global__ void sk_pool_reset_kernel(int32_t* __restrict__ pool_free, int* __restrict__ pool_ctl, int nchunks)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nchunks) pool_free[c] = c;
    if (c == 0)
    {
        pool_ctl[0] = nchunks;
        pool_ctl[1] = 0;
    }
}
__global__ void sk_set_density_kernel(SkCellRec* __restrict__ cells, const double* __restrict__ dens, int ncells)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < ncells) cells[m].dens = dens[m];
}
__global__ void sk_set_density_voronoi_kernel(double4* __restrict__ rec, const double* __restrict__ dens, int ncells)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < ncells) rec[m].w = dens[m];
}
__global__ void sk_sum_kernel(double* __restrict__ out, const double* __restrict__ a, const double* __restrict__ b,
                              const double* __restrict__ c, const double* __restrict__ d, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        double v = a[i] + b[i];
        if (c) v += c[i] + d[i];
        out[i] = v;
    }
}

// MediumSystem::totalDustAbsorbedLuminosity, MediumSystem.cpp:1317-1356 (single dust medium, constant sections)
__global__ void sk_absorbed_kernel(const double* __restrict__ rf, const double* __restrict__ dens_or_null,
                                   const SkCellRec* __restrict__ cells, const double4* __restrict__ vrec,
                                   const double* __restrict__ densx, int nmed,
                                   const double* __restrict__ kabs, int ncells, int nrf, double* out)
{
    double sum = 0.;
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < ncells; m += gridDim.x * blockDim.x)
    {
        double n = dens_or_null ? dens_or_null[m] : cells ? cells[m].dens : vrec[m].w;
        double s = 0.;
        for (int ell = 0; ell < nrf; ++ell)
        {
            // MediumSystem::opacityAbs over the dust components; kabs[h*nrf + ell]
            double opacity = kabs[ell] * n;
            for (int h = 1; h < nmed; ++h) opacity += kabs[h * nrf + ell] * densx[(size_t)(h - 1) * ncells + m];
            s += opacity * rf[(size_t)ell * ncells + m];  // SK_RF_INDEX
        }
        sum += s;
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, sum);
}

// ---------------------------------------------------------------------------------------------------
// engine object
// ---------------------------------------------------------------------------------------------------
struct HostInstr {
    sk_instrument_t d;
    int nl;
    size_t npix;
    bool include_sed, include_ifu, record_total_only;
    long long sed_off[SK_NUM_COMP], ifu_off[SK_NUM_COMP];  // offsets (doubles) into the detector block, -1 = absent
    long long wsed_off[5], wifu_off[5];
    int pix_slot = -1;
    int sed_slot = -1;  // relative to the pixel lists: the list of SED bins a run with kinematics keeps per history
};

struct sk_engine {
    sk_config_t cfg;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    SkDevModel M;
    // owned device allocations by group
    std::vector<void*> grid_allocs, medium_allocs, dust_allocs, wlg_allocs, src_allocs, instr_allocs, rf_allocs, sec_allocs;
    std::vector<void*> vel_allocs;  // kinematics: the per-cell bulk velocities (given, or zeros when only sources move)
    bool vel_given = false, src_moving = false;
    // host mirrors
    int grid_kind = 0;
    int grid_cells = 0;
    std::vector<double> dens_host;
    std::vector<double> dust_lam_border, dust_sig_abs;
    std::vector<sk_wavelength_grid_t> wlg_host;
    std::vector<std::vector<double>> wlg_lambda;
    std::vector<double> wlg_border_lo, wlg_border_hi;
    std::vector<double> Lv, Wv;
    double Ltot = 0.;
    uint64_t npackets = 0;
    std::vector<HostInstr> instr;
    double* det_block = nullptr;
    size_t det_count = 0;
    double* stat_block = nullptr;
    size_t stat_count = 0;
    unsigned long long* work_counter = nullptr;
    SkBank bank = {};  // the in-flight packets (sk_wavefront.cuh)
    int pool_chunks = 0;                                    // chunks in the pool of per-history pixel lists
    unsigned long long il_block = 0;                       // interleaved sharding (sk_engine_set_history_interleave)
    unsigned il_parts = 1, il_part = 0;
    unsigned int* sort_cursor = nullptr;                    // bin cursors of the wavelength sort of the forward rays
    unsigned long long pixel_overflows = 0;
    int bank_fields_d = 0, bank_fields_i = 0;
    unsigned int* ctl_host = nullptr;                       // pinned copy of the control words + work counter
    cudaEvent_t ev_ctl = nullptr;
    SkDevModel* model_dev = nullptr;
    int num_sms = 0;
    double* scalar = nullptr;
    int table_len[3] = {0, 0, 0};
    const int32_t* first_child_dev = nullptr;  // octree: first_child per node (owned by grid_allocs)
    std::map<const void*, int> occupancy;       // resident blocks per SM of the trace-kernel instantiations ...
    size_t occupancy_smem = (size_t)-1;         // ... for this dynamic shared-memory size
    size_t smem_bytes = 0;
    float last_ms = 0.f;
    bool timing_pending = false;
    std::vector<int> instr_same_observer;
    std::vector<std::array<double, 3>> instr_kobs;
    bool secondary_ready = false, has_secondary = false;
    // the tessellation built by sk_engine_build_voronoi, kept for sk_engine_read_voronoi
    std::vector<int64_t> voronoi_off;
    std::vector<int32_t> voronoi_idx;
    std::vector<double> voronoi_volume, voronoi_box;
    int num_mixes = 0;      // dust mixes given by sk_engine_set_dustmixes; must equal M.nmed when a segment runs
    int sec_num_media = 0;  // dust components the secondary-emission tables were given for
    int num_pix_lists = 0;
    int num_sed_lists = 0;  // SED instruments with statistics: one more list each in runs with kinematics
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    cudaEvent_t pin_ev[2] = {nullptr, nullptr};
    double* scratch = nullptr;
    size_t scratch_len = 0;
    sk_secondary_t sec;
    // page-locked staging of the per-cell arrays of sk_engine_prepare_secondary (luminosities down, launch weights and packet
    // map up: 32 bytes per cell and call), allocated once per cell count
    double* sec_pin = nullptr;
    size_t sec_pin_cells = 0;
    unsigned long long rounds_total = 0, launches_total = 0;
    // per-stage timing of the last segment: event pairs around every launch
    std::vector<cudaEvent_t> stage_events;  // pool, grown on demand
    std::vector<int> stage_of_pair;         // stage id of pair i (events 2i, 2i+1) in the current segment
    float stage_ms[SK_STAGE_COUNT] = {0};
};

static int stage_begin(sk_engine* e, int stage)
{
    size_t i = e->stage_of_pair.size();
    while (e->stage_events.size() < 2 * (i + 1))
    {
        cudaEvent_t ev;
        CK(cudaEventCreate(&ev));
        e->stage_events.push_back(ev);
    }
    e->stage_of_pair.push_back(stage);
    CK(cudaEventRecord(e->stage_events[2 * i], e->stream));
    e->launches_total++;
    return SK_OK;
}
static int stage_end(sk_engine* e)
{
    size_t i = e->stage_of_pair.size() - 1;
    CK(cudaEventRecord(e->stage_events[2 * i + 1], e->stream));
    return SK_OK;
}

static int bind(sk_engine* e)
{
    CK(cudaSetDevice(e->cfg.device));
    g_stream = e->stream;
    return SK_OK;
}
static void free_group(std::vector<void*>& v)
{
    for (void* p : v) dev_free(p);
    v.clear();
}
// kinematics are on when the media or any source move; a new medium state is at rest until the caller says otherwise
static void drop_velocities(sk_engine* e)
{
    free_group(e->vel_allocs);
    e->M.vel = nullptr;
    e->vel_given = false;
    e->M.kin = e->src_moving ? 1 : 0;
}

template <class T>
static int upload(std::vector<void*>& group, const T* host, size_t n, T** out)
{
    T* d = nullptr;
    CK(dev_malloc(&d, std::max<size_t>(n, 1) * sizeof(T)));
    group.push_back(d);
    if (n)
    {
        CK(cudaMemcpyAsync(d, host, n * sizeof(T), cudaMemcpyHostToDevice, g_stream));
        CK(cudaStreamSynchronize(g_stream));  // the caller's array may go away
    }
    *out = d;
    return SK_OK;
}
template <class T>
static int dalloc_zero(std::vector<void*>& group, size_t n, T** out)
{
    T* d = nullptr;
    CK(dev_malloc(&d, std::max<size_t>(n, 1) * sizeof(T)));
    group.push_back(d);
    CK(cudaMemsetAsync(d, 0, std::max<size_t>(n, 1) * sizeof(T), g_stream));
    *out = d;
    return SK_OK;
}

extern "C" int sk_engine_create(const sk_config_t* config, sk_engine_t** out)
{
    if (!config || !out) return fail(SK_ERR_INVALID, "null argument");
    int ndev = 0;
    cudaError_t err = cudaGetDeviceCount(&ndev);
    if (err != cudaSuccess || ndev == 0)
        return fail(SK_ERR_CUDA, "no CUDA device available (the engine has no CPU fallback)");
    if (config->device < 0 || config->device >= ndev) return fail(SK_ERR_INVALID, "device ordinal out of range");
    CK(cudaSetDevice(config->device));
    sk_engine* e = new sk_engine();
    e->cfg = *config;
    memset(&e->M, 0, sizeof e->M);
    e->M.seed = config->seed;
    e->M.force_scattering = config->force_scattering;
    e->M.min_scatt_events = config->min_scatt_events;
    e->M.explicit_absorption = config->explicit_absorption != 0;
    e->M.path_length_bias = config->path_length_bias;
    e->M.min_weight_reduction = config->min_weight_reduction;
    e->M.rf_grid = -1;
    e->M.nmed = 1;
    CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    g_stream = e->stream;
    {
        // keep freed blocks in the pool instead of returning them to the driver at the next synchronisation
        cudaMemPool_t pool;
        CK(cudaDeviceGetDefaultMemPool(&pool, config->device));
        uint64_t threshold = UINT64_MAX;
        CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
    }
    CK(cudaEventCreate(&e->ev0));
    CK(cudaEventCreate(&e->ev1));
    CK(dev_malloc(&e->work_counter, sizeof(unsigned long long)));
    CK(dev_malloc(&e->scalar, sizeof(double)));
    CK(dev_malloc(&e->M.counters, 16 * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(e->M.counters, 0, 16 * sizeof(unsigned long long), e->stream));
    *out = e;
    return SK_OK;
}

extern "C" void sk_engine_destroy(sk_engine_t* e)
{
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    g_stream = e->stream;
    cudaStreamSynchronize(e->stream);
    free_group(e->grid_allocs);
    free_group(e->medium_allocs);
    free_group(e->vel_allocs);
    free_group(e->dust_allocs);
    free_group(e->wlg_allocs);
    free_group(e->src_allocs);
    free_group(e->instr_allocs);
    free_group(e->rf_allocs);
    free_group(e->sec_allocs);
    dev_free(e->work_counter);
    dev_free(e->bank.d);
    dev_free(e->bank.i);
    dev_free(e->bank.list);
    dev_free(e->bank.free_list);
    dev_free(e->bank.pool_lell);
    dev_free(e->bank.pool_w);
    dev_free(e->bank.pool_next);
    dev_free(e->bank.pool_free);
    dev_free(e->bank.pool_ctl);
    dev_free(e->sort_cursor);
    dev_free(e->bank.ctl);
    cudaFreeHost(e->ctl_host);
    if (e->ev_ctl) cudaEventDestroy(e->ev_ctl);
    for (cudaEvent_t ev : e->stage_events) cudaEventDestroy(ev);
    dev_free(e->model_dev);
    dev_free(e->scratch);
    if (e->pinned) cudaFreeHost(e->pinned);
    if (e->sec_pin) cudaFreeHost(e->sec_pin);
    if (e->pin_ev[0]) cudaEventDestroy(e->pin_ev[0]);
    if (e->pin_ev[1]) cudaEventDestroy(e->pin_ev[1]);
    dev_free(e->scalar);
    dev_free(e->M.counters);
    cudaEventDestroy(e->ev0);
    cudaEventDestroy(e->ev1);
    cudaStreamSynchronize(e->stream);  // the stream-ordered frees above
    cudaStreamDestroy(e->stream);
    g_stream = nullptr;
    delete e;
}

// A new grid invalidates everything that is sized by its number of cells or points into it: the medium state, the
// radiation field tables (allocated by sk_engine_set_wavelength_grids) and the secondary-emission tables
// (sk_engine_set_secondary).  The caller repeats those calls for the new grid; until then the engine reports "not configured"
// instead of writing beyond buffers that were sized for the old grid.
static void drop_grid(sk_engine* e)
{
    free_group(e->grid_allocs);
    free_group(e->medium_allocs);
    free_group(e->rf_allocs);
    free_group(e->sec_allocs);
    drop_velocities(e);
    e->grid_kind = 0;
    e->grid_cells = 0;
    e->M.grid_kind = 0;
    e->M.ncells = 0;
    e->M.cells = nullptr;
    e->M.dens = nullptr;
    e->M.densx = nullptr;
    e->M.nmed = 1;
    e->M.volume = nullptr;
    e->M.vrec = nullptr;
    e->M.vnrec = nullptr;
    e->M.vbox = nullptr;
    e->M.node_child = nullptr;
    e->M.cell_coord = nullptr;
    e->M.xv = e->M.yv = e->M.zv = nullptr;
    e->M.rf1 = e->M.rf2 = e->M.rf2c = nullptr;
    e->M.rf_grid = -1;
    e->M.nmed = 1;
    e->M.nrf = 0;
    e->has_secondary = false;
    e->secondary_ready = false;
    e->first_child_dev = nullptr;
    e->dens_host.clear();
    e->voronoi_off.clear();
    e->voronoi_idx.clear();
    e->voronoi_volume.clear();
    e->voronoi_box.clear();
}

static int set_tables(sk_engine* e, const double* xv, int nx1, const double* yv, int ny1, const double* zv, int nz1)
{
    // device copies padded to an even number of entries: the trace kernels stage them with 16-byte-granular bulk copies
    double *dx, *dy, *dz;
    const double* src[3] = {xv, yv, zv};
    const int len[3] = {nx1, ny1, nz1};
    double** dst[3] = {&dx, &dy, &dz};
    for (int a = 0; a < 3; ++a)
    {
        std::vector<double> padded(SK_TABLE_PAD(len[a]), src[a][len[a] - 1]);
        std::copy(src[a], src[a] + len[a], padded.begin());
        if (int rc = upload(e->grid_allocs, padded.data(), padded.size(), dst[a])) return rc;
    }
    e->M.xv = dx;
    e->M.yv = dy;
    e->M.zv = dz;
    e->table_len[0] = nx1;
    e->table_len[1] = ny1;
    e->table_len[2] = nz1;
    size_t bytes = (size_t)(SK_TABLE_PAD(nx1) + SK_TABLE_PAD(ny1) + SK_TABLE_PAD(nz1)) * sizeof(double);
    // keep the tables in shared memory when they leave room for >= 4 CTAs per SM (227 KB per SM)
    e->M.lattice_in_smem = bytes <= 48 * 1024 ? 1 : 0;
    e->smem_bytes = e->M.lattice_in_smem ? bytes : 0;
    return SK_OK;
}

extern "C" int sk_engine_set_grid_cartesian(sk_engine_t* e, int32_t nx, int32_t ny, int32_t nz, const double* xv,
                                            const double* yv, const double* zv)
{
    if (!e || nx < 1 || ny < 1 || nz < 1 || !xv || !yv || !zv) return fail(SK_ERR_INVALID, "bad cartesian grid");
    if (int rc_bind = bind(e)) return rc_bind;
    drop_grid(e);
    e->grid_kind = 1;
    e->M.grid_kind = 1;
    e->M.nx = nx;
    e->M.ny = ny;
    e->M.nz = nz;
    e->grid_cells = nx * ny * nz;
    double ext[6] = {xv[0], yv[0], zv[0], xv[nx], yv[ny], zv[nz]};
    memcpy(e->M.ext, ext, sizeof ext);
    double dx = ext[3] - ext[0], dy = ext[4] - ext[1], dz = ext[5] - ext[2];
    e->M.eps = 1e-12 * sqrt(dx * dx + dy * dy + dz * dz);  // CartesianSpatialGrid.cpp:102
    e->M.ncells = 0;
    e->M.cells = nullptr;
    e->M.vrec = nullptr;
    e->M.vnrec = nullptr;
    e->M.vbox = nullptr;
    return set_tables(e, xv, nx + 1, yv, ny + 1, zv, nz + 1);
}

// Builds the device layout of an octree from the reference's node list (TreeSpatialGrid::_nodev order):
// integer lattice coordinates per node, per-axis lattice border tables computed with the reference's own
// midpoint arithmetic (Box::center, Box.hpp:135, applied recursively as OctTreeNode::createChildren does),
// and for every cell the six same-level-or-coarser neighbour links that replace TreeNode::_neighbors
// (OctTreeNode::addNeighbors, OctTreeNode.cpp:46-138).
// lattice border tables by recursive midpoints: the doubles Box::center (Box.hpp:135) yields level by level
static void midpoint_tables(const double extent[6], int N, std::vector<double> T[3])
{
    for (int a = 0; a < 3; ++a)
    {
        T[a].assign(N + 1, 0.);
        T[a][0] = extent[a];
        T[a][N] = extent[a + 3];
        for (int s = N; s > 1; s >>= 1)
            for (int lo = 0; lo < N; lo += s) T[a][lo + s / 2] = 0.5 * (T[a][lo] + T[a][lo + s]);
    }
}

// Common tail of sk_engine_set_grid_octree and sk_engine_build_octree: the node arrays are on the device (d_first and
// d_child already owned by grid_allocs, d_nodecoord = {ix,iy,iz,level} in units of the finest level, d_nodeofcell scratch)
static int finish_octree(sk_engine* e, const double extent[6], int nn, int nc, int maxlev, const int32_t* d_first,
                         const int32_t* d_child, const uint32_t* d_nodecoord, const int32_t* d_nodeofcell)
{
    const int N = 1 << maxlev;
    PhaseTimer pt("finish_octree");
    std::vector<double> T[3];
    midpoint_tables(extent, N, T);
    e->M.nx = N;
    e->M.ny = N;
    e->M.nz = N;
    e->M.maxlevel = maxlev;
    e->M.nnodes = nn;
    e->grid_cells = nc;
    e->M.ncells = 0;
    memcpy(e->M.ext, extent, 6 * sizeof(double));
    double dx = extent[3] - extent[0], dy = extent[4] - extent[1], dz = extent[5] - extent[2];
    e->M.eps = 1e-12 * sqrt(dx * dx + dy * dy + dz * dz);  // TreeSpatialGrid.cpp:28
    e->M.eps4 = 4. * e->M.eps;
    e->M.lat_h[0] = dx / N;  // pitch of the finest lattice: the trace kernels walk rays in lattice coordinates
    e->M.lat_h[1] = dy / N;
    e->M.lat_h[2] = dz / N;
    for (int a = 0; a < 3; ++a) e->M.lat_invh[a] = 1. / e->M.lat_h[a];
    uint32_t* d_coord;
    SkCellRec* d_cells;
    if (int rc = dalloc_zero(e->grid_allocs, 4 * (size_t)nc, &d_coord)) return rc;
    if (int rc = dalloc_zero(e->grid_allocs, (size_t)nc, &d_cells)) return rc;
    pt.lap("alloc+zero cells");
    sk_build_links_kernel<<<(nc + 127) / 128, 128, 0, e->stream>>>(d_first, d_child, d_nodecoord, d_nodeofcell, nc, N, maxlev,
                                                                  d_cells, d_coord);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    if (err != cudaSuccess) return fail(SK_ERR_CUDA, std::string("octree link builder: ") + cudaGetErrorString(err));
    pt.lap("link kernel + sync");
    e->M.node_child = d_child;
    e->M.cell_coord = d_coord;
    e->M.cells = d_cells;
    e->M.vrec = nullptr;
    e->M.vnrec = nullptr;
    e->M.vbox = nullptr;
    e->first_child_dev = d_first;
    if (int rc = set_tables(e, T[0].data(), N + 1, T[1].data(), N + 1, T[2].data(), N + 1)) return rc;
    e->grid_kind = 2;  // only a completely built grid is announced
    e->M.grid_kind = 2;
    return SK_OK;
}

static int set_grid_octree_device(sk_engine* e, const double extent[6], int nn, const int32_t* first_child);
extern "C" int sk_engine_set_grid_octree(sk_engine_t* e, const double extent[6], int32_t num_nodes,
                                         const int32_t* first_child)
{
    if (!e || !extent || num_nodes < 1 || !first_child) return fail(SK_ERR_INVALID, "bad octree");
    if (num_nodes > SK_LINK_INDEX_MASK) return fail(SK_ERR_UNSUPPORTED, "octree with more than 2^26 nodes");
    if (int rc_bind = bind(e)) return rc_bind;
    return set_grid_octree_device(e, extent, num_nodes, first_child);
}

// VoronoiMeshSnapshot as built by the reference's setup (sites + neighbour lists); the engine adds the start-cell table of
// its nearest-site walk (same construction as oracle/sk_oracle.c voronoi_build_blocks).
extern "C" int sk_engine_set_grid_voronoi(sk_engine_t* e, const double extent[6], int32_t num_cells, const double* sites,
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
    if (nbr_offset[num_cells] > 0x7fffffffLL) return fail(SK_ERR_UNSUPPORTED, "Voronoi mesh with more than 2^31 neighbour entries");
    if (int rc_bind = bind(e)) return rc_bind;
    drop_grid(e);
    const int nc = num_cells;
    std::vector<double4> rec(nc);
    for (int m = 0; m < nc; ++m) rec[m] = make_double4(sites[3 * (size_t)m], sites[3 * (size_t)m + 1], sites[3 * (size_t)m + 2], 0.);
    int nb = (int)std::cbrt((double)nc);
    nb = std::max(3, std::min(250, nb));  // VoronoiMeshSnapshot.cpp:544
    std::vector<int32_t> block((size_t)nb * nb * nb, -1);
    for (int m = 0; m < nc; ++m)
    {
        int i = (int)((rec[m].x - extent[0]) / (extent[3] - extent[0]) * nb);
        int j = (int)((rec[m].y - extent[1]) / (extent[4] - extent[1]) * nb);
        int k = (int)((rec[m].z - extent[2]) / (extent[5] - extent[2]) * nb);
        i = i < 0 ? 0 : i >= nb ? nb - 1 : i;
        j = j < 0 ? 0 : j >= nb ? nb - 1 : j;
        k = k < 0 ? 0 : k >= nb ? nb - 1 : k;
        size_t b = ((size_t)i * nb + j) * nb + k;
        if (block[b] < 0) block[b] = m;
    }
    int32_t last = 0;
    for (size_t b = 0; b < block.size(); ++b)
    {
        if (block[b] < 0)
            block[b] = last;
        else
            last = block[b];
    }
    std::vector<long long> off(nbr_offset, nbr_offset + nc + 1);
    // the neighbour records of the crossing loop: the neighbour's site next to its index, contiguous per cell
    std::vector<double4> nrec((size_t)nbr_offset[nc]);
    for (size_t i = 0; i < nrec.size(); ++i)
    {
        const int32_t mi = nbr_index[i];
        const long long bits = (long long)mi;
        double w;
        memcpy(&w, &bits, sizeof w);
        nrec[i] = mi >= 0 ? make_double4(rec[mi].x, rec[mi].y, rec[mi].z, w) : make_double4(0., 0., 0., w);
    }
    double4* d_nrec;
    if (int rc = upload(e->grid_allocs, nrec.data(), nrec.size(), &d_nrec)) return rc;
    double4* d_rec;
    long long* d_off;
    int32_t *d_idx, *d_block;
    if (int rc = upload(e->grid_allocs, rec.data(), (size_t)nc, &d_rec)) return rc;
    if (int rc = upload(e->grid_allocs, off.data(), (size_t)nc + 1, &d_off)) return rc;
    if (int rc = upload(e->grid_allocs, nbr_index, (size_t)nbr_offset[nc], &d_idx)) return rc;
    if (int rc = upload(e->grid_allocs, block.data(), block.size(), &d_block)) return rc;
    e->grid_kind = 3;
    e->M.grid_kind = 3;
    e->M.nx = e->M.ny = e->M.nz = 0;
    e->M.maxlevel = 0;
    e->M.nnodes = 0;
    e->M.cells = nullptr;
    e->M.xv = e->M.yv = e->M.zv = nullptr;
    e->M.lattice_in_smem = 0;
    e->smem_bytes = 0;
    e->M.vrec = d_rec;
    e->M.vnbr_off = d_off;
    e->M.vnbr = d_idx;
    e->M.vnrec = d_nrec;
    e->M.vblock = d_block;
    e->M.vbox = nullptr;
    e->M.vnb = nb;
    e->grid_cells = nc;
    e->M.ncells = 0;
    memcpy(e->M.ext, extent, 6 * sizeof(double));
    double dx = extent[3] - extent[0], dy = extent[4] - extent[1], dz = extent[5] - extent[2];
    e->M.eps = 1e-12 * sqrt(dx * dx + dy * dy + dz * dz);  // VoronoiMeshSnapshot.cpp:396
    return SK_OK;
}

extern "C" int sk_engine_set_voronoi_extents(sk_engine_t* e, int32_t num_cells, const double* boxes)
{
    if (!e || !boxes) return fail(SK_ERR_INVALID, "null argument");
    if (e->grid_kind != 3) return fail(SK_ERR_STATE, "set the Voronoi grid before its cell extents");
    if (num_cells != e->grid_cells) return fail(SK_ERR_INVALID, "extents do not match the grid");
    for (int m = 0; m < num_cells; ++m)
        for (int a = 0; a < 3; ++a)
            if (!(boxes[6 * (size_t)m + a] <= boxes[6 * (size_t)m + a + 3])) return fail(SK_ERR_INVALID, "empty cell extent");
    if (int rc_bind = bind(e)) return rc_bind;
    double* d;
    if (int rc = upload(e->grid_allocs, boxes, 6 * (size_t)num_cells, &d)) return rc;
    e->M.vbox = d;
    return SK_OK;
}

static int exclusive_scan(sk_engine* e, const int32_t* in, int32_t* out, int n, int32_t* sums, int32_t* total);
// VoronoiMeshSnapshot::buildMesh on the device (sk_setup.cuh): search grid on the host (a counting sort of the sites into cubic
// blocks of two sites on average), one thread per cell, the neighbour slots packed
// into lists on the device; the lists then take the same route as a tessellation handed in by the caller.
extern "C" int sk_engine_build_voronoi(sk_engine_t* e, const double extent[6], int32_t num_sites, const double* sites,
                                       uint64_t* num_entries)
{
    if (!e || !extent || num_sites < 1 || !sites) return fail(SK_ERR_INVALID, "bad voronoi sites");
    const int n = num_sites;
    for (int m = 0; m < n; ++m)
        for (int a = 0; a < 3; ++a)
            if (!(sites[3 * (size_t)m + a] > extent[a] && sites[3 * (size_t)m + a] < extent[a + 3]))
                return fail(SK_ERR_INVALID, "site outside the domain");
    if (int rc_bind = bind(e)) return rc_bind;
    PhaseTimer pt("build_voronoi");
    const double wx = extent[3] - extent[0], wy = extent[4] - extent[1], wz = extent[5] - extent[2];
    const double w = cbrt(2. * wx * wy * wz / n);
    int gx = std::max(1, (int)ceil(wx / w)), gy = std::max(1, (int)ceil(wy / w)), gz = std::max(1, (int)ceil(wz / w));
    const size_t ng = (size_t)gx * gy * gz;
    std::vector<int32_t> start(ng + 1, 0), blk(n), order(n);
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
        std::vector<int32_t> fill(start.begin(), start.end() - 1);
        for (int m = 0; m < n; ++m) order[fill[blk[m]]++] = m;
    }
    pt.lap("search grid (host)");
    std::vector<void*> scratch;
    struct Guard {
        std::vector<void*>& v;
        ~Guard() { free_group(v); }
    } guard{scratch};
    SkVoronoiBuild B;
    memcpy(B.ext, extent, sizeof B.ext);
    B.w = w;
    B.gx = gx;
    B.gy = gy;
    B.gz = gz;
    B.n = n;
    double *d_sites, *d_vol, *d_box;
    int32_t *d_start, *d_order, *d_blk, *d_nbr, *d_count, *d_off, *d_sums, *d_total, *d_idx;
    int* d_err;
    if (int rc = upload(scratch, sites, 3 * (size_t)n, &d_sites)) return rc;
    if (int rc = upload(scratch, start.data(), start.size(), &d_start)) return rc;
    if (int rc = upload(scratch, order.data(), order.size(), &d_order)) return rc;
    if (int rc = upload(scratch, blk.data(), blk.size(), &d_blk)) return rc;
    if (int rc = dalloc_zero(scratch, (size_t)n * SK_VC_MAXNB, &d_nbr)) return rc;
    if (int rc = dalloc_zero(scratch, (size_t)n, &d_count)) return rc;
    if (int rc = dalloc_zero(scratch, (size_t)n, &d_off)) return rc;
    if (int rc = dalloc_zero(scratch, (size_t)(n + SK_SCAN_BLOCK - 1) / SK_SCAN_BLOCK + 1, &d_sums)) return rc;
    if (int rc = dalloc_zero(scratch, (size_t)1, &d_total)) return rc;
    if (int rc = dalloc_zero(scratch, (size_t)n, &d_vol)) return rc;
    if (int rc = dalloc_zero(scratch, 6 * (size_t)n, &d_box)) return rc;
    if (int rc = dalloc_zero(scratch, (size_t)1, &d_err)) return rc;
    B.sites = d_sites;
    B.start = d_start;
    B.order = d_order;
    B.blk = d_blk;
    B.nbr = d_nbr;
    B.count = d_count;
    B.volume = d_vol;
    B.box = d_box;
    B.error = d_err;
    sk_voronoi_build_kernel<<<(n + 63) / 64, 64, 0, e->stream>>>(B);
    CK(cudaGetLastError());
    if (int rc = exclusive_scan(e, d_count, d_off, n, d_sums, d_total)) return rc;
    int32_t total = 0;
    int err = 0;
    CK(cudaMemcpyAsync(&total, d_total, sizeof total, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(&err, d_err, sizeof err, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    pt.lap("cells (device)");
    if (err)
        return fail(SK_ERR_UNSUPPORTED, err == -2   ? "Voronoi cell with more faces than the builder holds"
                                        : err == -4 ? "coinciding Voronoi sites"
                                        : err == -5 ? "Voronoi cell with more than 64 faces"
                                                    : "Voronoi sites in degenerate position (a cell that is not simple)");
    if (int rc = dalloc_zero(scratch, (size_t)std::max(total, 1), &d_idx)) return rc;
    sk_voronoi_compact_kernel<<<(n + 255) / 256, 256, 0, e->stream>>>(d_nbr, d_count, d_off, n, d_idx);
    CK(cudaGetLastError());
    std::vector<int32_t> off32(n), idx((size_t)total);
    std::vector<double> vol(n), box(6 * (size_t)n);
    CK(cudaMemcpyAsync(off32.data(), d_off, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(idx.data(), d_idx, (size_t)total * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(vol.data(), d_vol, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(box.data(), d_box, 6 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    std::vector<int64_t> off((size_t)n + 1);
    for (int m = 0; m < n; ++m) off[m] = off32[m];
    off[n] = total;
    pt.lap("lists to the host");
    if (int rc = sk_engine_set_grid_voronoi(e, extent, n, sites, off.data(), idx.data())) return rc;
    if (int rc = sk_engine_set_voronoi_extents(e, n, box.data())) return rc;
    e->voronoi_off = std::move(off);
    e->voronoi_idx = std::move(idx);
    e->voronoi_volume = std::move(vol);
    e->voronoi_box = std::move(box);
    if (num_entries) *num_entries = (uint64_t)total;
    return SK_OK;
}

extern "C" int sk_engine_read_voronoi(sk_engine_t* e, int64_t* nbr_offset, int32_t* nbr_index, double* volume, double* boxes)
{
    if (!e) return fail(SK_ERR_INVALID, "null engine");
    if (e->grid_kind != 3 || e->voronoi_off.empty()) return fail(SK_ERR_STATE, "the engine holds no tessellation built by sk_engine_build_voronoi");
    if (nbr_offset) memcpy(nbr_offset, e->voronoi_off.data(), e->voronoi_off.size() * sizeof(int64_t));
    if (nbr_index) memcpy(nbr_index, e->voronoi_idx.data(), e->voronoi_idx.size() * sizeof(int32_t));
    if (volume) memcpy(volume, e->voronoi_volume.data(), e->voronoi_volume.size() * sizeof(double));
    if (boxes) memcpy(boxes, e->voronoi_box.data(), e->voronoi_box.size() * sizeof(double));
    return SK_OK;
}

// MediumState::bulkVelocity(m) (MediumSystem.cpp:330-365) for the walks with kinematics (sk_wavefront.cuh)
extern "C" int sk_engine_set_velocities(sk_engine_t* e, int32_t num_cells, const double* velocity)
{
    if (!e) return fail(SK_ERR_INVALID, "null engine");
    if (int rc_bind = bind(e)) return rc_bind;
    drop_velocities(e);
    if (!velocity || num_cells <= 0) return SK_OK;
    if (!e->grid_kind || !e->M.ncells) return fail(SK_ERR_STATE, "set the medium state before the velocities");
    if (num_cells != e->M.ncells) return fail(SK_ERR_INVALID, "velocities do not match the medium state");
    std::vector<double4> v4((size_t)num_cells);
    for (int m = 0; m < num_cells; ++m) v4[m] = make_double4(velocity[3 * (size_t)m], velocity[3 * (size_t)m + 1], velocity[3 * (size_t)m + 2], 0.);
    double4* v;
    if (int rc = upload(e->vel_allocs, v4.data(), (size_t)num_cells, &v)) return rc;
    e->M.vel = v;
    e->vel_given = true;
    e->M.kin = 1;
    return SK_OK;
}

extern "C" int sk_engine_set_medium(sk_engine_t* e, int32_t num_cells, const double* number_density,
                                    const double* volume)
{
    return sk_engine_set_media(e, num_cells, 1, number_density, volume);
}

extern "C" int sk_engine_set_media(sk_engine_t* e, int32_t num_cells, int32_t num_media, const double* number_density,
                                   const double* volume)
{
    if (!e || !number_density) return fail(SK_ERR_INVALID, "null argument");
    if (num_media < 1 || num_media > SK_MAX_MEDIA)
        return fail(SK_ERR_UNSUPPORTED, "between 1 and " + std::to_string(SK_MAX_MEDIA) + " medium components are supported");
    if (!e->grid_kind) return fail(SK_ERR_STATE, "set the grid before the medium");
    if (num_cells != e->grid_cells) return fail(SK_ERR_INVALID, "medium size does not match the grid");
    if (int rc_bind = bind(e)) return rc_bind;
    free_group(e->medium_allocs);
    drop_velocities(e);
    e->M.densx = nullptr;
    e->M.nmed = num_media;
    if (num_media > 1)
    {
        // the components beyond the first: densx[(h-1)*ncells + m]
        double* dx;
        if (int rc = upload(e->medium_allocs, number_density + num_cells, (size_t)(num_media - 1) * num_cells, &dx)) return rc;
        e->M.densx = dx;
    }
    e->dens_host.assign(number_density, number_density + num_cells);
    if (e->grid_kind == 1)
    {
        double* d;
        if (int rc = upload(e->medium_allocs, number_density, (size_t)num_cells, &d)) return rc;
        e->M.dens = d;
    }
    else if (e->grid_kind == 3)
    {
        double* d;
        if (int rc = upload(e->medium_allocs, number_density, (size_t)num_cells, &d)) return rc;
        sk_set_density_voronoi_kernel<<<(num_cells + 255) / 256, 256, 0, e->stream>>>(const_cast<double4*>(e->M.vrec), d, num_cells);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(e->stream));
        e->M.dens = nullptr;
    }
    else
    {
        // the cell records (links) live with the grid; the density is written into them in place
        double* d;
        if (int rc = upload(e->medium_allocs, number_density, (size_t)num_cells, &d)) return rc;
        sk_set_density_kernel<<<(num_cells + 255) / 256, 256, 0, e->stream>>>(const_cast<SkCellRec*>(e->M.cells), d, num_cells);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(e->stream));
        e->M.dens = nullptr;
    }
    e->M.ncells = num_cells;
    e->M.volume = nullptr;
    if (volume)
    {
        double* v;
        if (int rc = upload(e->medium_allocs, volume, (size_t)num_cells, &v)) return rc;
        e->M.volume = v;
    }
    return SK_OK;
}

// ---------------------------------------------------------------------------------------------------
// setup on the device (sk_setup.cuh)
// ---------------------------------------------------------------------------------------------------
static int to_dev_geom(const sk_density_geometry_t& in, SkDevGeom& out)
{
    if (in.geometry < SK_GEOM_SHELL || in.geometry > SK_GEOM_SPIRAL_EXPDISK)
        return fail(SK_ERR_UNSUPPORTED, "geometry kind without a device density function");
    out.kind = in.geometry;
    out.pad = 0;
    out.number = in.number;
    out.mass = in.mass;
    memcpy(out.p, in.p, sizeof out.p);
    return SK_OK;
}
// exclusive scan of n int32 on the engine's stream; sums = scratch for ceil(n/1024) block sums, total = device int32
static int exclusive_scan(sk_engine* e, const int32_t* in, int32_t* out, int n, int32_t* sums, int32_t* total)
{
    const int nb = (n + SK_SCAN_BLOCK - 1) / SK_SCAN_BLOCK;
    sk_scan_block_kernel<<<nb, SK_SCAN_BLOCK, 0, e->stream>>>(in, out, sums, n);
    sk_scan_sums_kernel<<<1, SK_SCAN_BLOCK, 0, e->stream>>>(sums, nb, total);
    sk_scan_add_kernel<<<nb, SK_SCAN_BLOCK, 0, e->stream>>>(out, sums, n);
    CK(cudaGetLastError());
    return SK_OK;
}
namespace
{
    // device arrays of the growing node list
    struct TreeArrays {
        uint4* coord = nullptr;
        int32_t* first = nullptr;
        size_t cap = 0;
        ~TreeArrays()
        {
            dev_free(coord);
            dev_free(first);
        }
    };
}
static int grow_tree(sk_engine* e, TreeArrays& A, size_t need, size_t used)
{
    if (need <= A.cap) return SK_OK;
    size_t cap = std::max<size_t>(need, 2 * A.cap);
    uint4* c = nullptr;
    int32_t* f = nullptr;
    CK(dev_malloc(&c, cap * sizeof(uint4)));
    cudaError_t err = dev_malloc(&f, cap * sizeof(int32_t));
    if (err != cudaSuccess)
    {
        dev_free(c);
        return fail(SK_ERR_CUDA, std::string("octree node list: ") + cudaGetErrorString(err));
    }
    if (used)
    {
        cudaMemcpyAsync(c, A.coord, used * sizeof(uint4), cudaMemcpyDeviceToDevice, e->stream);
        cudaMemcpyAsync(f, A.first, used * sizeof(int32_t), cudaMemcpyDeviceToDevice, e->stream);
        cudaStreamSynchronize(e->stream);
    }
    dev_free(A.coord);
    dev_free(A.first);
    A.coord = c;
    A.first = f;
    A.cap = cap;
    return SK_OK;
}

// Tail shared by sk_engine_build_octree and sk_engine_set_grid_octree: the node list is on the device (first_child and
// lattice coordinates at a level `shift` finer than the deepest one reached); numbers the cells (leaves in node order,
// TreeSpatialGrid.cpp:40-48), takes ownership of copies of the node arrays and builds the neighbour links.
static int number_cells_and_finish(sk_engine* e, const double extent[6], int nn, int maxlev, int shift, uint4* d_coord,
                                   const int32_t* d_first_in, std::vector<void*>& scratch, int32_t* d_total, int* num_cells)
{
    int32_t *d_leaf, *d_cellrank, *d_sums2, *d_child, *d_first, *d_nodeofcell;
    if (int rc = dalloc_zero(scratch, (size_t)nn, &d_leaf)) return rc;
    if (int rc = dalloc_zero(scratch, (size_t)nn, &d_cellrank)) return rc;
    if (int rc = dalloc_zero(scratch, (size_t)nn / SK_SCAN_BLOCK + 2, &d_sums2)) return rc;
    sk_tree_leaf_flags_kernel<<<(nn + 255) / 256, 256, 0, e->stream>>>(d_coord, d_first_in, nn, shift, d_leaf);
    CK(cudaGetLastError());
    if (int rc = exclusive_scan(e, d_leaf, d_cellrank, nn, d_sums2, d_total)) return rc;
    int32_t nc = 0;
    CK(cudaMemcpyAsync(&nc, d_total, sizeof nc, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    drop_grid(e);  // from here on a failure leaves the engine without a grid (grid_kind 0), not with dangling tables
    if (int rc = dalloc_zero(e->grid_allocs, (size_t)nn, &d_child)) return rc;
    if (int rc = dalloc_zero(e->grid_allocs, (size_t)nn, &d_first)) return rc;
    if (int rc = dalloc_zero(scratch, (size_t)nc, &d_nodeofcell)) return rc;
    CK(cudaMemcpyAsync(d_first, d_first_in, (size_t)nn * sizeof(int32_t), cudaMemcpyDeviceToDevice, e->stream));
    sk_tree_number_cells_kernel<<<(nn + 255) / 256, 256, 0, e->stream>>>(d_first_in, d_cellrank, nn, d_child, d_nodeofcell);
    CK(cudaGetLastError());
    *num_cells = nc;
    return finish_octree(e, extent, nn, nc, maxlev, d_first, d_child, reinterpret_cast<const uint32_t*>(d_coord), d_nodeofcell);
}

// sk_engine_set_grid_octree: only the caller's first_child array crosses the bus; levels, lattice coordinates, the checks of
// the node list (parent before child, one parent per node, every node reachable, depth) and the cell numbering run on the
// device, one pass per level
static int set_grid_octree_device(sk_engine* e, const double extent[6], int nn, const int32_t* first_child)
{
    PhaseTimer pt("set_grid_octree");
    std::vector<void*> scratch;
    struct Guard {
        std::vector<void*>& v;
        ~Guard() { free_group(v); }
    } guard{scratch};
    int32_t *d_first, *d_parent, *d_total;
    int* d_status;
    uint4* d_coord;
    if (int rc = upload(scratch, first_child, (size_t)nn, &d_first)) return rc;
    pt.lap("upload first_child");
    if (int rc = dalloc_zero(scratch, (size_t)nn, &d_parent)) return rc;
    if (int rc = dalloc_zero(scratch, (size_t)nn, &d_coord)) return rc;
    if (int rc = dalloc_zero(scratch, 2, &d_status)) return rc;
    if (int rc = dalloc_zero(scratch, 1, &d_total)) return rc;
    const unsigned blocks = (unsigned)((nn + 255) / 256);
    sk_tree_init_kernel<<<blocks, 256, 0, e->stream>>>(d_coord, d_parent, nn);
    for (unsigned L = 0; L <= SK_MAX_TREE_LEVEL; ++L)
        sk_tree_propagate_kernel<<<blocks, 256, 0, e->stream>>>(d_first, nn, L, d_coord, d_parent, d_status);
    sk_tree_check_kernel<<<blocks, 256, 0, e->stream>>>(d_coord, nn, d_status);
    CK(cudaGetLastError());
    int status[2] = {0, 0};
    CK(cudaMemcpyAsync(status, d_status, sizeof status, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    pt.lap("levels and coordinates");
    if (status[0] & 1) return fail(SK_ERR_INVALID, "octree child index out of range or not after its parent");
    if (status[0] & 2) return fail(SK_ERR_INVALID, "octree node has two parents");
    if (status[0] & 4) return fail(SK_ERR_UNSUPPORTED, "octree deeper than 15 levels");
    if (status[0] & 8) return fail(SK_ERR_INVALID, "octree node list is not parent-before-child");
    const int maxlev = status[1];
    int nc = 0;
    int rc = number_cells_and_finish(e, extent, nn, maxlev, SK_MAX_TREE_LEVEL - maxlev, d_coord, d_first, scratch, d_total, &nc);
    pt.lap("cells and links");
    return rc;
}

extern "C" int sk_engine_build_octree(sk_engine_t* e, const double extent[6], const sk_tree_policy_t* policy, int32_t num_media,
                                      const sk_density_geometry_t* media, uint64_t* num_nodes, uint64_t* num_cells)
{
    if (!e || !extent || !policy || !media) return fail(SK_ERR_INVALID, "null argument");
    if (num_media < 1 || num_media > SK_SETUP_MAX_MEDIA) return fail(SK_ERR_UNSUPPORTED, "1..8 media components");
    if (policy->min_level < 0 || policy->max_level < policy->min_level) return fail(SK_ERR_INVALID, "bad tree levels");
    if (policy->max_level > SK_MAX_TREE_LEVEL) return fail(SK_ERR_UNSUPPORTED, "octree deeper than 15 levels");
    if (policy->num_samples < 1) return fail(SK_ERR_INVALID, "numDensitySamples must be positive");
    if (!(extent[3] > extent[0] && extent[4] > extent[1] && extent[5] > extent[2])) return fail(SK_ERR_INVALID, "empty extent");
    if (int rc_bind = bind(e)) return rc_bind;
    SkDevGeomSet G;
    memset(&G, 0, sizeof G);
    G.n = num_media;
    double dust_mass = 0.;
    for (int h = 0; h < num_media; ++h)
    {
        if (int rc = to_dev_geom(media[h], G.g[h])) return rc;
        dust_mass += media[h].mass;  // DensityTreePolicy::_dustMass, DensityTreePolicy.cpp:84-86
    }
    if (!(dust_mass > 0.)) return fail(SK_ERR_INVALID, "the media hold no dust mass");

    const int maxL = policy->max_level, NL = 1 << maxL;
    std::vector<double> T[3];
    midpoint_tables(extent, NL, T);
    std::vector<void*> scratch;
    struct Guard {
        std::vector<void*>& v;
        ~Guard() { free_group(v); }
    } guard{scratch};
    double *dx, *dy, *dz;
    if (int rc = upload(scratch, T[0].data(), T[0].size(), &dx)) return rc;
    if (int rc = upload(scratch, T[1].data(), T[1].size(), &dy)) return rc;
    if (int rc = upload(scratch, T[2].data(), T[2].size(), &dz)) return rc;
    int32_t* d_total;
    if (int rc = dalloc_zero(scratch, 1, &d_total)) return rc;
    SkTreeBuild B;
    B.xv = dx;
    B.yv = dy;
    B.zv = dz;
    B.N = NL;
    B.min_level = policy->min_level;
    B.max_level = policy->max_level;
    B.num_samples = policy->num_samples;
    B.max_fraction = policy->max_dust_fraction;
    B.max_tau = policy->max_dust_optical_depth;
    B.max_dispersion = policy->max_dust_density_dispersion;
    B.kappa = policy->dust_kappa;
    B.dust_mass = dust_mass;
    B.seed = (uint32_t)e->cfg.seed;

    TreeArrays A;
    if (int rc = grow_tree(e, A, 1 << 16, 0)) return rc;
    const uint4 root = make_uint4(0, 0, 0, 0);
    CK(cudaMemcpyAsync(A.coord, &root, sizeof root, cudaMemcpyHostToDevice, e->stream));
    // level-wise scratch, grown with the widest level
    int32_t *d_divide = nullptr, *d_rank = nullptr, *d_sums = nullptr;
    size_t level_cap = 0;
    std::vector<void*> level_scratch;
    Guard guard2{level_scratch};
    size_t lbeg = 0, lend = 1;
    int level = 0, maxlev = 0;
    while (lend != lbeg)  // DensityTreePolicy::constructTree, DensityTreePolicy.cpp:260-303
    {
        const int n = (int)(lend - lbeg);
        if ((size_t)n > level_cap)
        {
            free_group(level_scratch);
            level_cap = std::max<size_t>(2 * (size_t)n, 1 << 16);
            if (int rc = dalloc_zero(level_scratch, level_cap, &d_divide)) return rc;
            if (int rc = dalloc_zero(level_scratch, level_cap, &d_rank)) return rc;
            if (int rc = dalloc_zero(level_scratch, level_cap / SK_SCAN_BLOCK + 2, &d_sums)) return rc;
        }
        sk_tree_evaluate_kernel<<<(n + 127) / 128, 128, 0, e->stream>>>(B, G, A.coord, (int)lbeg, (int)lend, d_divide);
        CK(cudaGetLastError());
        if (int rc = exclusive_scan(e, d_divide, d_rank, n, d_sums, d_total)) return rc;
        int32_t ndiv = 0;
        CK(cudaMemcpyAsync(&ndiv, d_total, sizeof ndiv, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        const size_t nn_new = lend + 8 * (size_t)ndiv;
        if (nn_new > (size_t)SK_LINK_INDEX_MASK) return fail(SK_ERR_UNSUPPORTED, "octree with more than 2^26 nodes");
        if (int rc = grow_tree(e, A, nn_new, lend)) return rc;
        sk_tree_subdivide_kernel<<<(n + 127) / 128, 128, 0, e->stream>>>(NL, A.coord, A.first, (int)lbeg, (int)lend, d_divide,
                                                                        d_rank);
        CK(cudaGetLastError());
        if (ndiv) maxlev = level + 1;
        level++;
        lbeg = lend;
        lend = nn_new;
    }
    const int nn = (int)lend;
    int nc = 0;
    if (int rc = number_cells_and_finish(e, extent, nn, maxlev, maxL - maxlev, A.coord, A.first, scratch, d_total, &nc)) return rc;
    if (num_nodes) *num_nodes = (uint64_t)nn;
    if (num_cells) *num_cells = (uint64_t)nc;
    return SK_OK;
}

extern "C" int sk_engine_read_octree(sk_engine_t* e, int32_t* first_child)
{
    if (!e || !first_child) return fail(SK_ERR_INVALID, "null argument");
    if (e->grid_kind != 2 || !e->first_child_dev) return fail(SK_ERR_STATE, "the engine holds no octree");
    if (int rc_bind = bind(e)) return rc_bind;
    CK(cudaMemcpyAsync(first_child, e->first_child_dev, (size_t)e->M.nnodes * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return SK_OK;
}

extern "C" int sk_engine_sample_medium(sk_engine_t* e, const sk_density_geometry_t* medium, int32_t num_samples)
{
    if (!e || !medium) return fail(SK_ERR_INVALID, "null argument");
    if (num_samples < 1) return fail(SK_ERR_INVALID, "numDensitySamples must be positive");
    if (e->grid_kind != 1 && e->grid_kind != 2)
        return fail(e->grid_kind ? SK_ERR_UNSUPPORTED : SK_ERR_STATE, "density sampling needs a Cartesian or octree grid");
    SkDevGeom g;
    if (int rc = to_dev_geom(*medium, g)) return rc;
    if (int rc_bind = bind(e)) return rc_bind;
    free_group(e->medium_allocs);
    const int nc = e->grid_cells;
    double *d_vol, *d_dens = nullptr;
    if (int rc = dalloc_zero(e->medium_allocs, (size_t)nc, &d_vol)) return rc;
    if (e->grid_kind == 1)
    {
        if (int rc = dalloc_zero(e->medium_allocs, (size_t)nc, &d_dens)) return rc;
        sk_sample_medium_kernel<1><<<(nc + 127) / 128, 128, 0, e->stream>>>(e->M, g, num_samples, (uint32_t)e->cfg.seed, nc, nullptr,
                                                                            d_dens, d_vol);
    }
    else
        sk_sample_medium_kernel<2><<<(nc + 127) / 128, 128, 0, e->stream>>>(e->M, g, num_samples, (uint32_t)e->cfg.seed, nc,
                                                                            const_cast<SkCellRec*>(e->M.cells), nullptr, d_vol);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    e->dens_host.clear();
    drop_velocities(e);
    e->M.dens = d_dens;
    e->M.densx = nullptr;
    e->M.nmed = 1;
    e->M.volume = d_vol;
    e->M.ncells = nc;
    return SK_OK;
}

// ParticleMedium sampled on the device (sk_setup.cuh): the search blocks and their particle lists are built on the host
extern "C" int sk_engine_sample_medium_particles(sk_engine_t* e, int32_t num_particles, const double* particles,
                                                 double density_scale, int32_t num_samples)
{
    if (!e || !particles || num_particles < 1 || num_samples < 1) return fail(SK_ERR_INVALID, "bad particle medium");
    if (!e->grid_kind) return fail(SK_ERR_STATE, "set the grid before the medium");
    if (e->grid_kind == 3 && !e->M.vbox) return fail(SK_ERR_STATE, "sampling in Voronoi cells needs their extents");
    if (e->grid_kind == 3 && num_samples == 1)
        return fail(SK_ERR_UNSUPPORTED, "the density at the centroid of a Voronoi cell (numDensitySamples = 1)");
    const int np = num_particles;
    double hsum = 0.;
    for (int m = 0; m < np; ++m)
    {
        if (!(particles[5 * (size_t)m + 3] > 0.)) return fail(SK_ERR_INVALID, "smoothing length must be positive");
        hsum += particles[5 * (size_t)m + 3];
    }
    if (int rc_bind = bind(e)) return rc_bind;
    // search blocks: cubic, about 32768 of them but not much smaller than a typical particle
    const double* ext = e->M.ext;
    const double wx = ext[3] - ext[0], wy = ext[4] - ext[1], wz = ext[5] - ext[2];
    const double bw = std::max(cbrt(wx * wy * wz / 32768.), 0.75 * hsum / np);
    SkSphSearch S;
    memcpy(S.ext, ext, sizeof S.ext);
    S.nbx = std::max(1, std::min(256, (int)ceil(wx / bw)));
    S.nby = std::max(1, std::min(256, (int)ceil(wy / bw)));
    S.nbz = std::max(1, std::min(256, (int)ceil(wz / bw)));
    S.inv[0] = S.nbx / wx;
    S.inv[1] = S.nby / wy;
    S.inv[2] = S.nbz / wz;
    S.np = np;
    const size_t nblocks = (size_t)S.nbx * S.nby * S.nbz;
    auto range = [&](double lo, double hi, int axis, int nb, int& a, int& b) {
        a = (int)floor((lo - ext[axis]) * S.inv[axis]);
        b = (int)floor((hi - ext[axis]) * S.inv[axis]);
        a = std::max(a - 1, 0);          // one block of margin: the device rounds the block of a position on its own
        b = std::min(b + 1, nb - 1);
    };
    std::vector<int32_t> start(nblocks + 1, 0);
    for (int pass = 0; pass < 2; ++pass)
    {
        std::vector<int32_t> fill;
        std::vector<int32_t> list;
        if (pass == 1)
        {
            for (size_t b = 0; b < nblocks; ++b) start[b + 1] += start[b];
            if ((uint64_t)start[nblocks] > 0x7fffffffu) return fail(SK_ERR_UNSUPPORTED, "particle search lists beyond 2^31 entries");
            fill.assign(start.begin(), start.end() - 1);
            list.resize((size_t)start[nblocks]);
        }
        for (int m = 0; m < np; ++m)
        {
            const double* q = particles + 5 * (size_t)m;
            int i0, i1, j0, j1, k0, k1;
            range(q[0] - q[3], q[0] + q[3], 0, S.nbx, i0, i1);
            range(q[1] - q[3], q[1] + q[3], 1, S.nby, j0, j1);
            range(q[2] - q[3], q[2] + q[3], 2, S.nbz, k0, k1);
            for (int i = i0; i <= i1; ++i)
                for (int j = j0; j <= j1; ++j)
                    for (int k = k0; k <= k1; ++k)
                    {
                        const size_t b = ((size_t)i * S.nby + j) * S.nbz + k;
                        if (pass == 0)
                            start[b + 1]++;
                        else
                            list[fill[b]++] = m;   // particles in ascending index within every block
                    }
        }
        if (pass == 1)
        {
            std::vector<void*> scratch;
            struct Guard {
                std::vector<void*>& v;
                ~Guard() { free_group(v); }
            } guard{scratch};
            int32_t *d_start, *d_list;
            double *d_part, *d_dens;
            if (int rc = upload(scratch, start.data(), start.size(), &d_start)) return rc;
            if (int rc = upload(scratch, list.data(), std::max<size_t>(list.size(), 1), &d_list)) return rc;
            if (int rc = upload(scratch, particles, 5 * (size_t)np, &d_part)) return rc;
            S.start = d_start;
            S.list = d_list;
            S.part = d_part;
            const int nc = e->grid_cells;
            // the medium state: densities through a scratch array into the cell records; volumes of the boxes, or of the
            // tessellation this engine built
            free_group(e->medium_allocs);
            drop_velocities(e);
            e->M.densx = nullptr;
            e->M.nmed = 1;
            e->M.dens = nullptr;
            e->M.volume = nullptr;
            double* d_vol = nullptr;
            if (e->grid_kind != 3)
            {
                if (int rc = dalloc_zero(e->medium_allocs, (size_t)nc, &d_vol)) return rc;
            }
            else if (!e->voronoi_volume.empty())
            {
                if (int rc = upload(e->medium_allocs, e->voronoi_volume.data(), (size_t)nc, &d_vol)) return rc;
            }
            if (e->grid_kind == 1)
            {
                if (int rc = dalloc_zero(e->medium_allocs, (size_t)nc, &d_dens)) return rc;
                sk_sample_particles_kernel<1><<<(nc + 127) / 128, 128, 0, e->stream>>>(e->M, S, density_scale, num_samples, (uint32_t)e->cfg.seed,
                                                                                      nc, nullptr, nullptr, d_dens, d_vol);
                e->M.dens = d_dens;
            }
            else
            {
                if (int rc = dalloc_zero(scratch, (size_t)nc, &d_dens)) return rc;
                if (e->grid_kind == 2)
                {
                    sk_sample_particles_kernel<2><<<(nc + 127) / 128, 128, 0, e->stream>>>(e->M, S, density_scale, num_samples,
                                                                                          (uint32_t)e->cfg.seed, nc, nullptr, nullptr, d_dens, d_vol);
                    sk_set_density_kernel<<<(nc + 255) / 256, 256, 0, e->stream>>>(const_cast<SkCellRec*>(e->M.cells), d_dens, nc);
                }
                else
                {
                    sk_sample_particles_kernel<3><<<(nc + 127) / 128, 128, 0, e->stream>>>(e->M, S, density_scale, num_samples,
                                                                                          (uint32_t)e->cfg.seed, nc, nullptr,
                                                                                          const_cast<double4*>(e->M.vrec), d_dens, d_vol);
                    sk_set_density_voronoi_kernel<<<(nc + 255) / 256, 256, 0, e->stream>>>(const_cast<double4*>(e->M.vrec), d_dens, nc);
                }
            }
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(e->stream));
            e->dens_host.clear();
            e->M.volume = d_vol;
            e->M.ncells = nc;
        }
    }
    return SK_OK;
}

extern "C" int sk_engine_read_medium(sk_engine_t* e, double* number_density, double* volume)
{
    if (!e) return fail(SK_ERR_INVALID, "null argument");
    if (!e->M.ncells) return fail(SK_ERR_STATE, "the engine holds no medium state");
    if (int rc_bind = bind(e)) return rc_bind;
    const int nc = e->M.ncells;
    if (number_density)
    {
        if (e->M.dens)
            CK(cudaMemcpyAsync(number_density, e->M.dens, (size_t)nc * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
        else
        {
            std::vector<void*> scratch;
            double* d;
            int rc = dalloc_zero(scratch, (size_t)nc, &d);
            if (!rc)
            {
                sk_gather_density_kernel<<<(nc + 255) / 256, 256, 0, e->stream>>>(e->M.cells, e->M.vrec, nc, d);
                cudaError_t err = cudaGetLastError();
                if (err == cudaSuccess)
                    err = cudaMemcpyAsync(number_density, d, (size_t)nc * sizeof(double), cudaMemcpyDeviceToHost, e->stream);
                if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
                if (err != cudaSuccess) rc = fail(SK_ERR_CUDA, std::string("read medium: ") + cudaGetErrorString(err));
            }
            free_group(scratch);
            if (rc) return rc;
        }
    }
    if (volume)
    {
        if (!e->M.volume) return fail(SK_ERR_STATE, "the engine holds no cell volumes");
        CK(cudaMemcpyAsync(volume, e->M.volume, (size_t)nc * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    }
    CK(cudaStreamSynchronize(e->stream));
    return SK_OK;
}

extern "C" int sk_engine_set_dustmix(sk_engine_t* e, const sk_dustmix_t* mix)
{
    return sk_engine_set_dustmixes(e, 1, mix);
}

extern "C" int sk_engine_set_dustmixes(sk_engine_t* e, int32_t num_media, const sk_dustmix_t* mixes)
{
    if (!e || !mixes || num_media < 1 || mixes[0].num_lambda < 2) return fail(SK_ERR_INVALID, "bad dust mix");
    if (num_media > SK_MAX_MEDIA)
        return fail(SK_ERR_UNSUPPORTED, "between 1 and " + std::to_string(SK_MAX_MEDIA) + " medium components are supported");
    const int n = mixes[0].num_lambda;
    for (int h = 0; h < num_media; ++h)
    {
        if (!mixes[h].lambda_border || !mixes[h].sigma_abs || !mixes[h].sigma_sca || !mixes[h].asymmpar)
            return fail(SK_ERR_INVALID, "bad dust mix");
        if (mixes[h].num_lambda != n || memcmp(mixes[h].lambda_border, mixes[0].lambda_border, (size_t)n * sizeof(double)))
            return fail(SK_ERR_INVALID, "the dust mixes of one simulation share one wavelength grid (DustMix.cpp:52-98)");
    }
    if (int rc_bind = bind(e)) return rc_bind;
    free_group(e->dust_allocs);
    // one table set per component, [h*n + i]
    std::vector<double> abs((size_t)num_media * n), sca((size_t)num_media * n), ext((size_t)num_media * n), gp((size_t)num_media * n);
    for (int h = 0; h < num_media; ++h)
        for (int i = 0; i < n; ++i)
        {
            const size_t k = (size_t)h * n + i;
            abs[k] = mixes[h].sigma_abs[i];
            sca[k] = mixes[h].sigma_sca[i];
            ext[k] = mixes[h].sigma_abs[i] + mixes[h].sigma_sca[i];  // DustMix.cpp:160-163
            gp[k] = mixes[h].asymmpar[i];
        }
    double *a, *b, *c, *d, *g;
    if (int rc = upload(e->dust_allocs, mixes[0].lambda_border, (size_t)n, &a)) return rc;
    if (int rc = upload(e->dust_allocs, abs.data(), abs.size(), &b)) return rc;
    if (int rc = upload(e->dust_allocs, sca.data(), sca.size(), &c)) return rc;
    if (int rc = upload(e->dust_allocs, ext.data(), ext.size(), &d)) return rc;
    if (int rc = upload(e->dust_allocs, gp.data(), gp.size(), &g)) return rc;
    e->M.nlam = n;
    e->M.lam_border = a;
    e->M.sig_abs = b;
    e->M.sig_sca = c;
    e->M.sig_ext = d;
    e->M.gpar = g;
    e->num_mixes = num_media;
    e->dust_lam_border.assign(mixes[0].lambda_border, mixes[0].lambda_border + n);
    e->dust_sig_abs = abs;
    return SK_OK;
}

extern "C" int sk_engine_set_wavelength_grids(sk_engine_t* e, int32_t n, const sk_wavelength_grid_t* grids,
                                              int32_t rf_grid)
{
    if (!e || n < 0 || (n && !grids) || rf_grid >= n) return fail(SK_ERR_INVALID, "bad wavelength grids");
    if (rf_grid >= 0 && e->M.ncells <= 0) return fail(SK_ERR_STATE, "set the medium before a radiation field grid");
    if (int rc_bind = bind(e)) return rc_bind;
    free_group(e->wlg_allocs);
    free_group(e->rf_allocs);
    std::vector<SkDevWlg> dev(n ? n : 1);
    e->wlg_host.assign(grids, grids + n);
    e->wlg_lambda.clear();
    e->wlg_border_lo.clear();
    e->wlg_border_hi.clear();
    for (int i = 0; i < n; ++i)
    {
        double *b, *l, *d;
        int32_t* el;
        if (int rc = upload(e->wlg_allocs, grids[i].borders, (size_t)grids[i].num_borders, &b)) return rc;
        if (int rc = upload(e->wlg_allocs, grids[i].ell, (size_t)grids[i].num_borders + 1, &el)) return rc;
        if (int rc = upload(e->wlg_allocs, grids[i].lambda, (size_t)grids[i].num_bins, &l)) return rc;
        if (int rc = upload(e->wlg_allocs, grids[i].dlambda, (size_t)grids[i].num_bins, &d)) return rc;
        dev[i].num_bins = grids[i].num_bins;
        dev[i].num_borders = grids[i].num_borders;
        dev[i].borders = b;
        dev[i].ell = el;
        dev[i].lambda = l;
        dev[i].dlambda = d;
        e->wlg_lambda.emplace_back(grids[i].lambda, grids[i].lambda + grids[i].num_bins);
        e->wlg_border_lo.push_back(grids[i].borders[0]);
        e->wlg_border_hi.push_back(grids[i].borders[grids[i].num_borders - 1]);
    }
    SkDevWlg* dw;
    if (int rc = upload(e->wlg_allocs, dev.data(), dev.size(), &dw)) return rc;
    e->M.wlg = dw;
    e->M.nwlg = n;
    e->M.rf_grid = rf_grid;
    e->M.nrf = 0;
    e->M.rf1 = e->M.rf2 = e->M.rf2c = nullptr;
    if (rf_grid >= 0)
    {
        e->M.nrf = grids[rf_grid].num_bins;
        size_t cnt = (size_t)e->M.ncells * e->M.nrf;
        if (int rc = dalloc_zero(e->rf_allocs, cnt, &e->M.rf1)) return rc;
        if (int rc = dalloc_zero(e->rf_allocs, cnt, &e->M.rf2)) return rc;
        if (int rc = dalloc_zero(e->rf_allocs, cnt, &e->M.rf2c)) return rc;
    }
    return SK_OK;
}

// SourceSystem::setupSelfAfter, SourceSystem.cpp:14-41
extern "C" int sk_engine_set_sources(sk_engine_t* e, int32_t n, const sk_source_t* sources, double source_bias)
{
    if (!e || n < 1 || !sources) return fail(SK_ERR_INVALID, "bad sources");
    if (int rc_bind = bind(e)) return rc_bind;
    free_group(e->src_allocs);
    double L = 0.;
    for (int h = 0; h < n; ++h) L += sources[h].luminosity;
    e->Ltot = L;
    e->Lv.assign(n, 0.);
    e->Wv.assign(n, 0.);
    if (L)
    {
        double wLsum = 0., wsum = 0.;
        for (int h = 0; h < n; ++h)
        {
            e->Lv[h] = sources[h].luminosity / L;
            wLsum += sources[h].source_weight * e->Lv[h];
            wsum += sources[h].source_weight;
        }
        for (int h = 0; h < n; ++h)
            e->Wv[h] = (1 - source_bias) * (sources[h].source_weight * e->Lv[h]) / wLsum
                       + source_bias * sources[h].source_weight / wsum;
    }
    std::vector<SkDevSource> dev(n);
    bool moving = false;
    for (int h = 0; h < n; ++h)
    {
        const sk_source_t& s = sources[h];
        SkDevSource& d = dev[h];
        memset(&d, 0, sizeof d);
        if (s.kind != SK_SRC_POINT && s.kind != SK_SRC_GEOMETRIC) return fail(SK_ERR_UNSUPPORTED, "unknown source kind");
        if (s.kind == SK_SRC_GEOMETRIC && (s.geometry < SK_GEOM_SHELL || s.geometry > SK_GEOM_SPIRAL_EXPDISK))
            return fail(SK_ERR_UNSUPPORTED, "geometry has no device sampler");
        if (s.sed_n < 2 || !s.sed_lambda || !s.sed_p || !s.sed_P) return fail(SK_ERR_INVALID, "source without SED tables");
        d.kind = s.kind;
        d.geometry = s.geometry;
        d.sed_kind = s.sed_kind;
        d.bias_kind = s.bias_kind;
        memcpy(d.position, s.position, sizeof d.position);
        memcpy(d.gp, s.geom_params, sizeof d.gp);
        d.geom_table_n = s.geom_table_n;
        d.sed_n = s.sed_n;
        d.oligo_n = s.oligo_n;
        double* p;
        if (s.geom_table_n)
        {
            if (int rc = upload(e->src_allocs, s.geom_table_x, (size_t)s.geom_table_n, &p)) return rc;
            d.geom_table_x = p;
            if (int rc = upload(e->src_allocs, s.geom_table_P, (size_t)s.geom_table_n, &p)) return rc;
            d.geom_table_P = p;
        }
        if (int rc = upload(e->src_allocs, s.sed_lambda, (size_t)s.sed_n, &p)) return rc;
        d.sed_lambda = p;
        if (int rc = upload(e->src_allocs, s.sed_p, (size_t)s.sed_n, &p)) return rc;
        d.sed_p = p;
        if (int rc = upload(e->src_allocs, s.sed_P, (size_t)s.sed_n, &p)) return rc;
        d.sed_P = p;
        if (s.oligo_n)
        {
            if (int rc = upload(e->src_allocs, s.oligo_lambda, (size_t)s.oligo_n, &p)) return rc;
            d.oligo_lambda = p;
        }
        d.sed_temperature = s.sed_temperature;
        d.sed_norm = s.sed_norm;
        d.wavelength_bias = s.wavelength_bias;
        d.bias_min = s.bias_min;
        d.bias_max = s.bias_max;
        d.oligo_probability = s.oligo_probability;
        d.Lw = L ? e->Lv[h] / e->Wv[h] : 0.;
        if (s.velocity_kind < SK_VEL_NONE || s.velocity_kind > SK_VEL_CYLINDRICAL) return fail(SK_ERR_UNSUPPORTED, "unknown source velocity kind");
        d.velocity_kind = s.velocity_kind;
        memcpy(d.velocity, s.velocity, sizeof d.velocity);
        if (s.velocity_kind != SK_VEL_NONE) moving = true;
    }
    e->src_moving = moving;
    e->M.kin = (e->vel_given || moving) ? 1 : 0;
    SkDevSource* ds;
    if (int rc = upload(e->src_allocs, dev.data(), dev.size(), &ds)) return rc;
    unsigned long long* iv;
    if (int rc = dalloc_zero(e->src_allocs, (size_t)n + 1, &iv)) return rc;
    e->M.src = ds;
    e->M.Iv = iv;
    e->M.nsrc = n;
    e->npackets = 0;
    return SK_OK;
}

// DistantInstrument::setupSelfBefore (DistantInstrument.cpp:39-50), FrameInstrument::setupSelfBefore
// (FrameInstrument.cpp:12-32), FluxRecorder::finalizeConfiguration (FluxRecorder.cpp:185-300)
extern "C" int sk_engine_set_instruments(sk_engine_t* e, int32_t n, const sk_instrument_t* instruments,
                                         int32_t has_medium_emission)
{
    if (!e || n < 0 || (n && !instruments)) return fail(SK_ERR_INVALID, "bad instruments");
    if (n > SK_MAX_INSTR) return fail(SK_ERR_UNSUPPORTED, "more than 8 instruments");
    if (int rc_bind = bind(e)) return rc_bind;
    free_group(e->instr_allocs);
    e->instr.assign(n, HostInstr());
    size_t det = 0, stat = 0;
    int num_pix_lists = 0, num_sed_lists = 0;
    for (int i = 0; i < n; ++i)
    {
        HostInstr& q = e->instr[i];
        q.d = instruments[i];
        const sk_instrument_t& d = q.d;
        if (d.wavelength_grid < 0 || d.wavelength_grid >= e->M.nwlg)
            return fail(SK_ERR_INVALID, "instrument wavelength grid index out of range");
        if (d.num_scattering_levels > SK_MAX_LEVELS) return fail(SK_ERR_UNSUPPORTED, "too many scattering levels");
        if (d.kind < SK_INSTR_SED || d.kind > SK_INSTR_FULL) return fail(SK_ERR_UNSUPPORTED, "unknown instrument kind");
        if (!(d.redshift > -1.)) return fail(SK_ERR_INVALID, "instrument redshift must exceed -1");
        q.include_sed = d.kind == SK_INSTR_SED || d.kind == SK_INSTR_FULL;
        q.include_ifu = d.kind == SK_INSTR_FRAME || d.kind == SK_INSTR_FULL;
        q.nl = e->wlg_host[d.wavelength_grid].num_bins;
        q.npix = q.include_ifu ? (size_t)d.num_pixels_x * d.num_pixels_y : 0;
        q.record_total_only = !d.record_components;
        size_t lensed = q.include_sed ? (size_t)q.nl : 0, lenifu = q.include_ifu ? q.npix * q.nl : 0;
        for (int c = 0; c < SK_NUM_COMP; ++c)
        {
            bool need;
            if (q.record_total_only)
                need = c == SK_COMP_TOTAL;
            else if (c == SK_COMP_TRANSPARENT || c == SK_COMP_PRIMARY_DIRECT || c == SK_COMP_PRIMARY_SCATTERED)
                need = true;
            else if (c == SK_COMP_SECONDARY_DIRECT || c == SK_COMP_SECONDARY_SCATTERED || c == SK_COMP_SECONDARY_TRANSPARENT)
                need = has_medium_emission != 0;
            else if (c >= SK_COMP_PRIMARY_SCATTERED_LEVEL)
                need = (c - SK_COMP_PRIMARY_SCATTERED_LEVEL) < d.num_scattering_levels;
            else
                need = false;
            q.sed_off[c] = q.ifu_off[c] = -1;
            if (need && lensed)
            {
                q.sed_off[c] = (long long)det;
                det += lensed;
            }
            if (need && lenifu)
            {
                q.ifu_off[c] = (long long)det;
                det += lenifu;
            }
        }
        q.sed_slot = (d.record_statistics && lensed) ? num_sed_lists++ : -1;
        q.pix_slot = -1;
        if (d.record_statistics && lenifu)
        {
            if (lenifu > (size_t)0x7fffffff) return fail(SK_ERR_UNSUPPORTED, "per-pixel statistics on a frame with more than 2^31 pixel bins");
            q.pix_slot = num_pix_lists++;
        }
        for (int k = 0; k < 5; ++k)
        {
            q.wifu_off[k] = -1;
            if (q.pix_slot >= 0)
            {
                q.wifu_off[k] = (long long)stat;
                stat += lenifu;
            }
        }
        for (int k = 0; k < 5; ++k)
        {
            q.wsed_off[k] = -1;
            if (d.record_statistics && lensed)
            {
                q.wsed_off[k] = (long long)stat;
                stat += lensed;
            }
        }
    }
    if (int rc = dalloc_zero(e->instr_allocs, det, &e->det_block)) return rc;
    if (int rc = dalloc_zero(e->instr_allocs, stat, &e->stat_block)) return rc;
    e->det_count = det;
    e->stat_count = stat;
    std::vector<SkDevInstr> dev(n ? n : 1);
    for (int i = 0; i < n; ++i)
    {
        const HostInstr& q = e->instr[i];
        const sk_instrument_t& d = q.d;
        SkDevInstr& v = dev[i];
        memset(&v, 0, sizeof v);
        v.kind = d.kind;
        v.wlg = d.wavelength_grid;
        v.nx = d.num_pixels_x;
        v.ny = d.num_pixels_y;
        v.num_levels = d.num_scattering_levels;
        v.record_total_only = q.record_total_only;
        v.record_stats = d.record_statistics;
        v.include_sed = q.include_sed;
        v.include_ifu = q.include_ifu;
        v.nl = q.nl;
        v.npix = q.npix;
        v.costheta = cos(d.inclination);
        v.sintheta = sin(d.inclination);
        v.cosphi = cos(d.azimuth);
        v.sinphi = sin(d.azimuth);
        v.cosomega = cos(d.roll);
        v.sinomega = sin(d.roll);
        // Direction(inclination, azimuth), Direction.cpp:10-35
        {
            const double eps = 1e-8;
            double th = d.inclination, ph = d.azimuth;
            if (th <= eps)
            {
                v.kobs[0] = 0;
                v.kobs[1] = 0;
                v.kobs[2] = 1;
            }
            else if (th >= M_PI - eps)
            {
                v.kobs[0] = 0;
                v.kobs[1] = 0;
                v.kobs[2] = -1;
            }
            else
            {
                double st = sin(th);
                v.kobs[0] = st * cos(ph);
                v.kobs[1] = st * sin(ph);
                v.kobs[2] = cos(th);
            }
        }
        v.radius2 = d.radius * d.radius;
        v.zp1 = 1. + d.redshift;
        if (q.include_ifu)
        {
            v.xpmin = d.center_x - 0.5 * d.field_of_view_x;
            v.xpsiz = d.field_of_view_x / d.num_pixels_x;
            v.ypmin = d.center_y - 0.5 * d.field_of_view_y;
            v.ypsiz = d.field_of_view_y / d.num_pixels_y;
        }
        v.same_as_preceding = 0;
        if (i > 0)
        {
            const sk_instrument_t& p = e->instr[i - 1].d;  // DistantInstrument.cpp:54-62
            if (d.distance == p.distance && d.inclination == p.inclination && d.azimuth == p.azimuth && d.roll == p.roll)
                v.same_as_preceding = 1;
        }
        for (int c = 0; c < SK_NUM_COMP; ++c)
        {
            v.sed[c] = q.sed_off[c] >= 0 ? e->det_block + q.sed_off[c] : nullptr;
            v.ifu[c] = q.ifu_off[c] >= 0 ? e->det_block + q.ifu_off[c] : nullptr;
        }
        for (int k = 0; k < 5; ++k) v.wsed[k] = q.wsed_off[k] >= 0 ? e->stat_block + q.wsed_off[k] : nullptr;
        for (int k = 0; k < 5; ++k) v.wifu[k] = q.wifu_off[k] >= 0 ? e->stat_block + q.wifu_off[k] : nullptr;
        v.pix_slot = q.pix_slot;
        v.sed_slot = q.sed_slot >= 0 ? num_pix_lists + q.sed_slot : -1;  // (the lists of SED bins follow the pixel lists)
    }
    e->instr_same_observer.assign(n, 0);
    e->instr_kobs.assign(n, std::array<double, 3>{0., 0., 1.});
    for (int i = 0; i < n; ++i)
    {
        e->instr_same_observer[i] = dev[i].same_as_preceding;
        e->instr_kobs[i] = {dev[i].kobs[0], dev[i].kobs[1], dev[i].kobs[2]};
    }
    SkDevInstr* di;
    if (int rc = upload(e->instr_allocs, dev.data(), dev.size(), &di)) return rc;
    e->M.instr = di;
    e->M.ninstr = n;
    e->num_pix_lists = num_pix_lists;
    e->num_sed_lists = num_sed_lists;
    e->M.pix_base_d = SK_BANK_FIELDS_D(n);
    e->M.pix_base_i = SK_BANK_FIELDS_I(n);
    return SK_OK;
}

// nearest-grid-point lookup of the dust mix: DustMix::indexForLambda = NR::locateClip(_lambdav, lambda), DustMix.cpp:276-279
static int dust_index_for_lambda(const sk_engine* e, double lambda)
{
    const std::vector<double>& b = e->dust_lam_border;
    int n = (int)b.size();
    if (lambda < b[0]) return 0;
    int jl = -1, ju = n - 1;
    while (ju - jl > 1)
    {
        int jm = (ju + jl) >> 1;
        if (lambda < b[jm])
            ju = jm;
        else
            jl = jm;
    }
    return jl;
}

extern "C" int sk_engine_set_secondary(sk_engine_t* e, const sk_secondary_t* sec)
{
    return sk_engine_set_secondary_media(e, 1, sec);
}

extern "C" int sk_engine_set_secondary_media(sk_engine_t* e, int32_t num_media, const sk_secondary_t* sec)
{
    if (!e || !sec || num_media < 1) return fail(SK_ERR_INVALID, "null argument");
    if (num_media != e->num_mixes) return fail(SK_ERR_INVALID, "one set of emission tables per dust mix is needed");
    if (e->M.rf_grid < 0) return fail(SK_ERR_STATE, "dust emission needs a radiation field grid");
    if (e->grid_kind == 3 && !e->M.vbox)
        return fail(SK_ERR_STATE, "dust emission from a Voronoi grid needs the cell extents (sk_engine_set_voronoi_extents)");
    if (!e->M.volume) return fail(SK_ERR_STATE, "dust emission needs the cell volumes (sk_engine_set_medium)");
    if (!e->M.nlam) return fail(SK_ERR_STATE, "set the dust mix before the secondary emission tables");
    if (sec->emission_grid < 0 || sec->emission_grid >= e->M.nwlg) return fail(SK_ERR_INVALID, "bad emission grid index");
    for (int h = 0; h < num_media; ++h)
    {
        if (sec[h].num_temperatures < 2 || !sec[h].temperature || !sec[h].planck_abs || !sec[h].rf_sigma_abs || !sec[h].em_sigma_abs)
            return fail(SK_ERR_INVALID, "missing emission calculator tables");
        if (sec[h].emission_grid != sec->emission_grid || sec[h].num_temperatures != sec->num_temperatures
            || memcmp(sec[h].temperature, sec->temperature, (size_t)sec->num_temperatures * sizeof(double)))
            return fail(SK_ERR_INVALID, "the emission calculators of the dust components must share their grids");
    }
    if (int rc_bind = bind(e)) return rc_bind;
    free_group(e->sec_allocs);
    e->sec = *sec;
    const sk_wavelength_grid_t& g = e->wlg_host[sec->emission_grid];
    const std::vector<double>& glam = e->wlg_lambda[sec->emission_grid];
    const int n = g.num_bins, nem = n + 2, nrf = e->M.nrf, nc = e->M.ncells;
    // DisjointWavelengthGrid::extlambdav, DisjointWavelengthGrid.cpp:346-356
    std::vector<double> ext(nem);
    ext[0] = e->wlg_border_lo[sec->emission_grid];
    for (int ell = 0; ell < n; ++ell) ext[ell + 1] = glam[ell];
    ext[nem - 1] = e->wlg_border_hi[sec->emission_grid];
    const int nT = sec->num_temperatures, nlam = e->M.nlam;
    std::vector<double> kabs((size_t)num_media * nrf), emsig((size_t)num_media * nem), rfsig((size_t)num_media * nrf),
        planckabs((size_t)num_media * nT);
    const std::vector<double>& rflam = e->wlg_lambda[e->M.rf_grid];
    for (int h = 0; h < num_media; ++h)
    {
        for (int ell = 0; ell < nrf; ++ell)
            kabs[(size_t)h * nrf + ell] = e->dust_sig_abs[(size_t)h * nlam + dust_index_for_lambda(e, rflam[ell])];
        std::copy(sec[h].em_sigma_abs, sec[h].em_sigma_abs + nem, emsig.begin() + (size_t)h * nem);
        std::copy(sec[h].rf_sigma_abs, sec[h].rf_sigma_abs + nrf, rfsig.begin() + (size_t)h * nrf);
        std::copy(sec[h].planck_abs, sec[h].planck_abs + nT, planckabs.begin() + (size_t)h * nT);
    }
    double *a, *b, *c, *d, *f, *k, *cmb;
    std::vector<double> cmbv((size_t)std::max(nrf, 1), 0.);
    if (sec->rf_cmb) std::copy(sec->rf_cmb, sec->rf_cmb + nrf, cmbv.begin());
    if (int rc = upload(e->sec_allocs, cmbv.data(), cmbv.size(), &cmb)) return rc;
    e->M.sec_cmb = cmb;
    if (int rc = upload(e->sec_allocs, ext.data(), (size_t)nem, &a)) return rc;
    if (int rc = upload(e->sec_allocs, emsig.data(), emsig.size(), &b)) return rc;
    if (int rc = upload(e->sec_allocs, rfsig.data(), rfsig.size(), &c)) return rc;
    if (int rc = upload(e->sec_allocs, sec->temperature, (size_t)nT, &d)) return rc;
    if (int rc = upload(e->sec_allocs, planckabs.data(), planckabs.size(), &f)) return rc;
    if (int rc = upload(e->sec_allocs, kabs.data(), kabs.size(), &k)) return rc;
    e->M.sec_nem = nem;
    e->M.sec_nT = sec->num_temperatures;
    e->M.sec_lambda = a;
    e->M.sec_emsig = b;
    e->M.sec_rfsig = c;
    e->M.sec_T = d;
    e->M.sec_planckabs = f;
    e->M.sec_kabs_rf = k;
    if (int rc = dalloc_zero(e->sec_allocs, (size_t)nc * nem, &e->M.sec_pv)) return rc;
    if (int rc = dalloc_zero(e->sec_allocs, (size_t)nc * nem, &e->M.sec_Pv)) return rc;
    if (int rc = dalloc_zero(e->sec_allocs, (size_t)nc, &e->M.sec_Lv)) return rc;
    if (int rc = dalloc_zero(e->sec_allocs, (size_t)nc, &e->M.sec_ws)) return rc;
    if (int rc = dalloc_zero(e->sec_allocs, (size_t)nc + 1, &e->M.sec_Iv)) return rc;
    e->M.sec_xi = sec->wavelength_bias;
    e->M.sec_bias_min = sec->bias_min;
    e->M.sec_bias_max = sec->bias_max;
    e->has_secondary = true;
    e->sec_num_media = num_media;
    e->secondary_ready = false;
    return SK_OK;
}

extern "C" int sk_engine_clear_instruments(sk_engine_t* e)
{
    if (!e) return fail(SK_ERR_INVALID, "null engine");
    if (int rc_bind = bind(e)) return rc_bind;
    if (e->det_count) CK(cudaMemsetAsync(e->det_block, 0, e->det_count * sizeof(double), e->stream));
    if (e->stat_count) CK(cudaMemsetAsync(e->stat_block, 0, e->stat_count * sizeof(double), e->stream));
    return SK_OK;
}

extern "C" int sk_engine_cuda_stream(sk_engine_t* e, void** stream)
{
    if (!e || !stream) return fail(SK_ERR_INVALID, "null argument");
    *stream = (void*)e->stream;
    return SK_OK;
}

extern "C" int sk_engine_clear_rf(sk_engine_t* e, int32_t primary)
{
    if (!e) return fail(SK_ERR_INVALID, "null engine");
    if (int rc_bind = bind(e)) return rc_bind;
    size_t bytes = (size_t)e->M.ncells * e->M.nrf * sizeof(double);
    if (!bytes) return SK_OK;
    if (primary)
    {
        CK(cudaMemsetAsync(e->M.rf1, 0, bytes, e->stream));
        CK(cudaMemsetAsync(e->M.rf2, 0, bytes, e->stream));
    }
    else
        CK(cudaMemsetAsync(e->M.rf2c, 0, bytes, e->stream));
    return SK_OK;
}

// SourceSystem::prepareForLaunch, SourceSystem.cpp:75-97
extern "C" int sk_engine_prepare_primary(sk_engine_t* e, uint64_t num_packets)
{
    if (!e || !e->M.nsrc) return fail(SK_ERR_STATE, "no sources");
    if (!e->Ltot)
        return fail(SK_ERR_INVALID, "Cannot launch primary source photon packets when total luminosity is zero");
    if (!num_packets) return fail(SK_ERR_INVALID, "zero packets");
    if (int rc_bind = bind(e)) return rc_bind;
    int Ns = e->M.nsrc;
    std::vector<unsigned long long> Iv(Ns + 1);
    Iv[0] = 0;
    double W = 0.;
    for (int h = 1; h != Ns; ++h)
    {
        W += e->Wv[h - 1];
        unsigned long long idx = (unsigned long long)std::round(W * (double)num_packets);
        Iv[h] = std::min<unsigned long long>(idx, num_packets);
    }
    Iv[Ns] = num_packets;
    CK(cudaMemcpyAsync(const_cast<unsigned long long*>(e->M.Iv), Iv.data(), Iv.size() * sizeof(unsigned long long),
                       cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->M.Lpp = e->Ltot / (double)num_packets;
    e->npackets = num_packets;
    return SK_OK;
}

// SecondarySourceSystem::prepareForLaunch (SecondarySourceSystem.cpp:84-126) for the single DustSecondarySource:
// prepareLuminosities / preparePacketMap (DustSecondarySource.cpp:26-146).  The per-cell luminosities and spectra are
// computed on the device; the cumulative weights are summed sequentially on the host in the reference's order so that the
// history-index map _Iv is the reference's (it rounds W*N, DustSecondarySource.cpp:136-145).
extern "C" int sk_engine_prepare_secondary(sk_engine_t* e, uint64_t num_packets, double* luminosity)
{
    if (!e || !luminosity) return fail(SK_ERR_INVALID, "null argument");
    if (!e->has_secondary) return fail(SK_ERR_STATE, "call sk_engine_set_secondary first");
    if (!num_packets) return fail(SK_ERR_INVALID, "zero packets");
    if (int rc_bind = bind(e)) return rc_bind;
    const int M = e->M.ncells;
    const unsigned blocks = (unsigned)((M + 127) / 128);
    if (e->grid_kind == 1)
        sk_dust_luminosity_kernel<1><<<blocks, 128, 0, e->stream>>>(e->M);
    else if (e->grid_kind == 2)
        sk_dust_luminosity_kernel<2><<<blocks, 128, 0, e->stream>>>(e->M);
    else
        sk_dust_luminosity_kernel<3><<<blocks, 128, 0, e->stream>>>(e->M);
    CK(cudaGetLastError());
    if (e->sec_pin_cells != (size_t)M)
    {
        if (e->sec_pin) cudaFreeHost(e->sec_pin);
        e->sec_pin = nullptr;
        e->sec_pin_cells = 0;
        CK(cudaMallocHost(&e->sec_pin, (3 * (size_t)M + 2) * sizeof(double)));
        e->sec_pin_cells = (size_t)M;
    }
    double* Lv = e->sec_pin;                                                         // [M]
    double* ws = e->sec_pin + (size_t)M;                                              // [M]
    unsigned long long* Iv = reinterpret_cast<unsigned long long*>(e->sec_pin + 2 * (size_t)M);  // [M + 1]
    CK(cudaMemcpyAsync(Lv, e->M.sec_Lv, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    double L = 0.;
    for (int m = 0; m < M; ++m) L += Lv[m];
    *luminosity = L;
    e->secondary_ready = false;
    if (!L) return SK_OK;
    double wsum = 0.;
    for (int m = 0; m < M; ++m)
    {
        Lv[m] /= L;
        wsum += Lv[m] > 0 ? 1. : 0.;
    }
    const double xi = e->sec.spatial_bias;
    Iv[0] = 0;
    double W = 0.;
    for (int m = 0; m < M; ++m)
    {
        double w = (Lv[m] > 0 ? 1. : 0.) / wsum;
        double Wm = (1 - xi) * Lv[m] + xi * w;
        ws[m] = Wm > 0 ? Lv[m] / Wm : 0.;
        if (m + 1 != M)
        {
            W += Wm;
            unsigned long long idx = (unsigned long long)std::round(W * (double)num_packets);
            Iv[m + 1] = std::min<unsigned long long>(idx, num_packets);
        }
    }
    Iv[M] = num_packets;
    CK(cudaMemcpyAsync(e->M.sec_Lv, Lv, (size_t)M * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->M.sec_ws, ws, (size_t)M * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->M.sec_Iv, Iv, ((size_t)M + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, e->stream));
    if (e->grid_kind == 1)
        sk_emission_spectrum_kernel<1><<<blocks, 128, 0, e->stream>>>(e->M);
    else if (e->grid_kind == 2)
        sk_emission_spectrum_kernel<2><<<blocks, 128, 0, e->stream>>>(e->M);
    else
        sk_emission_spectrum_kernel<3><<<blocks, 128, 0, e->stream>>>(e->M);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    e->M.sec_Lpp = L / (double)num_packets;  // SecondarySourceSystem.cpp:119
    e->secondary_ready = true;
    return SK_OK;
}

// ---------------------------------------------------------------------------------------------------
// the segment driver: performLifeCycle(firstIndex, numIndices, primary, peel, store) over the bank
// ---------------------------------------------------------------------------------------------------
static size_t bank_capacity_limit(int nlists)
{
    // in-flight packets per GPU; ~230 B of state each, 3.9 GB at 2^24 (measured on cfg2, 1e8 packets: 2^22 462 ms, 2^23 442 ms,
    // 2^24 432 ms, 2^25 435 ms per step: fewer rounds, a shorter drain).  Per-history pixel / bin lists add 392 B per slot and
    // list plus their pool of chunks: those runs keep 2^23.  SK_BANK overrides (tests use small banks to exercise refill).
    const char* s = getenv("SK_BANK");
    long long v = s ? atoll(s) : 0;
    return v > 0 ? (size_t)v : (size_t)1 << (nlists > 0 ? 23 : 24);
}

static int ensure_bank(sk_engine* e, uint64_t count)
{
    const int nlists = e->num_pix_lists + (e->M.kin ? e->num_sed_lists : 0);
    size_t cap = std::min<uint64_t>(count, bank_capacity_limit(nlists));
    cap = std::max<size_t>((cap + 255) / 256 * 256, 256);
    int nd = SK_BANK_FIELDS_D(e->M.ninstr) + nlists * SK_PIX_K;
    int ni = SK_BANK_FIELDS_I(e->M.ninstr) + nlists * SK_PIX_INTS;
    e->M.kin_base_d = nd;  // the extra per-packet fields of a run with kinematics follow the pixel lists
    e->M.kin_base_i = ni;
    if (e->M.kin)
    {
        nd += SK_KD_COUNT;
        ni += SK_KI_COUNT;
    }
    e->bank.n = (int32_t)cap;
    if ((size_t)e->bank.cap >= cap && e->bank_fields_d == nd && e->bank_fields_i == ni) return SK_OK;
    dev_free(e->bank.d);
    dev_free(e->bank.i);
    dev_free(e->bank.list);
    dev_free(e->bank.free_list);
    dev_free(e->bank.pool_lell);
    dev_free(e->bank.pool_w);
    dev_free(e->bank.pool_next);
    dev_free(e->bank.pool_free);
    e->bank.d = nullptr;
    e->bank.i = nullptr;
    e->bank.list = nullptr;
    e->bank.free_list = nullptr;
    e->bank.pool_lell = e->bank.pool_next = e->bank.pool_free = nullptr;
    e->bank.pool_w = nullptr;
    e->bank.cap = 0;
    e->pool_chunks = 0;
    if (nlists)
    {
        // continuation chunks of the per-history pixel lists: two per slot and list, at least 16384 (388 bytes each;
        // SK_PIX_POOL overrides).  A history takes one for every SK_PIX_C distinct frame pixels beyond the first SK_PIX_K.
        const char* ps = getenv("SK_PIX_POOL");
        const long long want = ps ? atoll(ps) : 0;
        const size_t nch = want > 0 ? (size_t)want : std::max<size_t>(16384, 2 * cap * nlists);
        CK(dev_malloc(&e->bank.pool_lell, nch * SK_PIX_C * sizeof(int32_t)));
        CK(dev_malloc(&e->bank.pool_w, nch * SK_PIX_C * sizeof(double)));
        CK(dev_malloc(&e->bank.pool_next, nch * sizeof(int32_t)));
        CK(dev_malloc(&e->bank.pool_free, nch * sizeof(int32_t)));
        if (!e->bank.pool_ctl) CK(dev_malloc(&e->bank.pool_ctl, 2 * sizeof(int)));
        e->pool_chunks = (int)nch;
    }
    CK(dev_malloc(&e->bank.d, cap * nd * sizeof(double)));
    CK(dev_malloc(&e->bank.i, cap * ni * sizeof(int32_t)));
    CK(dev_malloc(&e->bank.list, cap * sizeof(int32_t)));
    CK(dev_malloc(&e->bank.free_list, cap * sizeof(int32_t)));
    if (!e->bank.ctl) CK(dev_malloc(&e->bank.ctl, SK_CTL_WORDS * sizeof(unsigned int)));
    if (!e->ctl_host) CK(cudaMallocHost(&e->ctl_host, (SK_CTL_WORDS + 2) * sizeof(unsigned int)));
    if (!e->ev_ctl) CK(cudaEventCreateWithFlags(&e->ev_ctl, cudaEventDisableTiming));
    e->bank.cap = (int32_t)cap;
    e->bank_fields_d = nd;
    e->bank_fields_i = ni;
    return SK_OK;
}

template <int GRID, int MODE, bool STORE, bool SMEMT, bool MULTI, bool KIN = false>
static int launch_trace_impl(sk_engine* e, const SkRunArgs& A, const SkObsDir& dir)
{
    auto kern = sk_wf_trace<GRID, MODE, STORE, SMEMT, MULTI, KIN>;
    // occupancy of this instantiation on this engine's device for its shared-memory footprint (cached in the engine:
    // engines on different devices run from different host threads).  The staged tables never exceed 48 KB (set_tables),
    // the default limit of dynamic shared memory, so no per-device function attribute has to be set.
    const size_t smem = SMEMT ? e->smem_bytes : 0;
    if (e->occupancy_smem != e->smem_bytes)
    {
        e->occupancy.clear();
        e->occupancy_smem = e->smem_bytes;
    }
    int per_sm = 0;
    auto it = e->occupancy.find((const void*)kern);
    if (it != e->occupancy.end())
        per_sm = it->second;
    else
    {
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SK_TRACE_BLOCK, smem));
        if (per_sm < 1) per_sm = 1;
        e->occupancy[(const void*)kern] = per_sm;
    }
    // persistent grid: every SM filled to the occupancy this kernel gets, but no more warps than chunks of rays
    unsigned long long chunks = ((unsigned long long)e->bank.n + SK_CHUNK - 1) / SK_CHUNK;
    unsigned long long blocks = (chunks + (SK_TRACE_BLOCK / 32) - 1) / (SK_TRACE_BLOCK / 32);
    unsigned grid = (unsigned)std::min<unsigned long long>((unsigned long long)e->num_sms * per_sm, std::max<unsigned long long>(blocks, 1));
    CK(cudaMemsetAsync(&e->bank.ctl[SK_CTL_CURSOR], 0, sizeof(unsigned int), e->stream));
    if (int rc = stage_begin(e, MODE == 0 ? SK_STAGE_TRACE_FORWARD : MODE == 1 ? SK_STAGE_TRACE_INTERACTION : SK_STAGE_TRACE_PEEL))
        return rc;
    kern<<<grid, SK_TRACE_BLOCK, smem, e->stream>>>(e->M, A, e->bank, dir);
    CK(cudaGetLastError());
    return stage_end(e);
}

template <int GRID, int MODE, bool STORE>
static int launch_trace(sk_engine* e, const SkRunArgs& A, const SkObsDir& dir)
{
    // only the Cartesian grid looks borders up while it walks (TMA-staged tables); the octree walks in lattice coordinates
    // several medium components with their own mixes: the instantiation that sums the opacities
    // kinematics: the instantiation that looks the sections up in every cell at the wavelength the cell perceives
    if (e->M.kin)
    {
        if (GRID == 1 && e->M.lattice_in_smem) return launch_trace_impl<GRID, MODE, STORE, GRID == 1, true, true>(e, A, dir);
        return launch_trace_impl<GRID, MODE, STORE, false, true, true>(e, A, dir);
    }
    if (e->M.nmed > 1)
    {
        if (GRID == 1 && e->M.lattice_in_smem) return launch_trace_impl<GRID, MODE, STORE, GRID == 1, true>(e, A, dir);
        return launch_trace_impl<GRID, MODE, STORE, false, true>(e, A, dir);
    }
    if (GRID == 1 && e->M.lattice_in_smem) return launch_trace_impl<GRID, MODE, STORE, GRID == 1, false>(e, A, dir);
    return launch_trace_impl<GRID, MODE, STORE, false, false>(e, A, dir);
}

template <int GRID>
static int run_bank(sk_engine* e, const SkRunArgs& A)
{
    const SkDevModel& M = e->M;
    const SkBank& K = e->bank;
    unsigned eblocks = (unsigned)((K.n + SK_EVENT_BLOCK - 1) / SK_EVENT_BLOCK);
    // observer groups: consecutive instruments that share the observer (Instrument.hpp:107) share one peel-off ray
    std::vector<std::pair<int, int>> groups;
    if (A.peel)
        for (int j0 = 0; j0 < (int)e->instr.size();)
        {
            int j1 = j0 + 1;
            while (j1 < (int)e->instr.size() && e->instr_same_observer[j1]) j1++;
            groups.emplace_back(j0, j1);
            j0 = j1;
        }
    // detect is a persistent grid-stride kernel with the SED arrays of the observer group in shared memory
    auto launch_detect = [&](int j0, int j1, int last) -> int {
        const unsigned dblocks = std::min<unsigned>(eblocks, (unsigned)e->num_sms * 8u);
        int nl = 0;
        for (int j = j0; j < j1; ++j)
            if (e->instr[j].include_sed) nl = std::max(nl, e->instr[j].nl);
        size_t smem = (size_t)(j1 - j0) * SK_NUM_COMP * nl * sizeof(double);
        if (smem > 40 * 1024)  // too many bins for shared memory: add straight to the global arrays
        {
            nl = 0;
            smem = 0;
        }
        if (int rc = stage_begin(e, SK_STAGE_DETECT)) return rc;
        if (M.kin)
            sk_wf_detect<true><<<dblocks, SK_EVENT_BLOCK, smem, e->stream>>>(M, A, K, j0, j1, last, nl);
        else
            sk_wf_detect<false><<<dblocks, SK_EVENT_BLOCK, smem, e->stream>>>(M, A, K, j0, j1, last, nl);
        CK(cudaGetLastError());
        return stage_end(e);
    };
    SkObsDir nodir;
    nodir.d.set(0., 0., 1., nullptr);
    nodir.px = nodir.py = 0.;
    nodir.pz = 1.;
    CK(cudaMemsetAsync(K.i + (size_t)I_STATE * K.cap, 0, (size_t)K.n * sizeof(int32_t), e->stream));
    for (unsigned long long round = 0;; ++round)
    {
        CK(cudaMemsetAsync(K.ctl, 0, SK_CTL_WORDS * sizeof(unsigned int), e->stream));
        const int g0a = groups.empty() ? 0 : groups[0].first, g0b = groups.empty() ? 0 : groups[0].second;
        if (int rc = stage_begin(e, SK_STAGE_ADVANCE)) return rc;
        if (M.nmed > 1 || M.kin)
            sk_wf_advance<GRID, true><<<eblocks, SK_EVENT_BLOCK, 0, e->stream>>>(M, A, K, g0a, g0b);
        else
            sk_wf_advance<GRID, false><<<eblocks, SK_EVENT_BLOCK, 0, e->stream>>>(M, A, K, g0a, g0b);
        CK(cudaGetLastError());
        if (int rc = stage_end(e)) return rc;
        if (int rc = stage_begin(e, SK_STAGE_LAUNCH)) return rc;
        sk_wf_launch<GRID><<<eblocks, SK_EVENT_BLOCK, 0, e->stream>>>(M, A, K, g0a, g0b);
        CK(cudaGetLastError());
        if (int rc = stage_end(e)) return rc;
        CK(cudaMemcpyAsync(e->ctl_host, K.ctl, SK_CTL_WORDS * sizeof(unsigned int), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaMemcpyAsync(e->ctl_host + SK_CTL_WORDS, A.work_counter, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                           e->stream));
        CK(cudaEventRecord(e->ev_ctl, e->stream));
        // the rest of the round is enqueued before the census is looked at, so the device never waits for the host
        if (groups.empty())
        {
            if (int rc = launch_detect(0, 0, 1)) return rc;
        }
        for (size_t gi = 0; gi < groups.size(); ++gi)
        {
            const int j0 = groups[gi].first, j1 = groups[gi].second;
            if (gi > 0)
            {
                CK(cudaMemsetAsync(&K.ctl[SK_CTL_NLIST], 0, sizeof(unsigned int), e->stream));
                if (int rc = stage_begin(e, SK_STAGE_PEEL_SETUP)) return rc;
                sk_wf_peel_setup<<<eblocks, SK_EVENT_BLOCK, 0, e->stream>>>(M, A.model, K, j0, j1);
                CK(cudaGetLastError());
                if (int rc = stage_end(e)) return rc;
            }
            SkObsDir obs;
            obs.px = e->instr_kobs[j0][0];
            obs.py = e->instr_kobs[j0][1];
            obs.pz = e->instr_kobs[j0][2];
            obs.d.set(obs.px, obs.py, obs.pz, GRID == 2 ? M.lat_h : nullptr);
            if (int rc = launch_trace<GRID, 2, false>(e, A, obs)) return rc;
            const int last = gi + 1 == groups.size();
            if (last) CK(cudaMemsetAsync(&K.ctl[SK_CTL_NLIST], 0, sizeof(unsigned int), e->stream));
            if (int rc = launch_detect(j0, j1, last)) return rc;
        }
        if (M.force_scattering)
        {
            // forward path, interaction optical depth and the walk to the interaction point in one kernel; when the radiation
            // field is stored the rays are first put in order of their wavelength bin (see SK_RF_INDEX)
            if (A.store && M.nrf > 1 && M.nrf <= SK_SORT_MAX_BINS)
            {
                if (!e->sort_cursor) CK(dev_malloc(&e->sort_cursor, (SK_SORT_MAX_BINS + 1) * sizeof(unsigned int)));
                CK(cudaMemsetAsync(e->sort_cursor, 0, (M.nrf + 1) * sizeof(unsigned int), e->stream));
                if (int rc = stage_begin(e, SK_STAGE_PEEL_SETUP)) return rc;
                const unsigned cblocks = std::min<unsigned>(eblocks, (unsigned)e->num_sms * 8u);
                sk_wf_bin_count<<<cblocks, SK_EVENT_BLOCK, 0, e->stream>>>(K, M.nrf, e->sort_cursor);
                sk_wf_bin_scan<<<1, 32, 0, e->stream>>>(e->sort_cursor, M.nrf);
                const unsigned sblocks = (unsigned)(((size_t)K.n + SK_SORT_TILE - 1) / SK_SORT_TILE);
                sk_wf_bin_scatter<<<sblocks, SK_EVENT_BLOCK, 0, e->stream>>>(K, M.nrf, e->sort_cursor);
                CK(cudaGetLastError());
                e->launches_total += 2;
                if (int rc = stage_end(e)) return rc;
                std::swap(e->bank.list, e->bank.free_list);  // the sorted list is the ray list of the trace ...
                int rc = launch_trace<GRID, 0, true>(e, A, nodir);
                std::swap(e->bank.list, e->bank.free_list);  // ... and the buffers return to their roles
                if (rc) return rc;
            }
            else if (int rc = A.store ? launch_trace<GRID, 0, true>(e, A, nodir) : launch_trace<GRID, 0, false>(e, A, nodir))
                return rc;
        }
        else
        {
            CK(cudaMemsetAsync(&K.ctl[SK_CTL_NLIST], 0, sizeof(unsigned int), e->stream));
            if (int rc = stage_begin(e, SK_STAGE_SAMPLE)) return rc;
            sk_wf_sample<<<eblocks, SK_EVENT_BLOCK, 0, e->stream>>>(M, A, K);
            CK(cudaGetLastError());
            if (int rc = stage_end(e)) return rc;
            if (int rc = launch_trace<GRID, 1, false>(e, A, nodir)) return rc;
        }
        // census of this round: stop when the bank is empty and every history has been handed out
        CK(cudaEventSynchronize(e->ev_ctl));
        unsigned long long dispensed;
        memcpy(&dispensed, e->ctl_host + SK_CTL_WORDS, sizeof dispensed);
        e->rounds_total++;
        const unsigned nlive = e->ctl_host[SK_CTL_NLIVE];
        if (nlive == 0 && dispensed >= A.count) break;
        // draining: pack the survivors into the head of the bank when fewer than half of the slots in use are alive
        if (dispensed >= A.count && (size_t)nlive * 2 < (size_t)K.n && K.n > SK_EVENT_BLOCK)
        {
            const int n_new = (int)std::max<size_t>(((size_t)nlive + SK_EVENT_BLOCK - 1) / SK_EVENT_BLOCK * SK_EVENT_BLOCK, SK_EVENT_BLOCK);
            CK(cudaMemsetAsync(K.ctl, 0, SK_CTL_WORDS * sizeof(unsigned int), e->stream));
            sk_wf_partition<<<eblocks, SK_EVENT_BLOCK, 0, e->stream>>>(K, n_new);
            CK(cudaGetLastError());
            // at most the survivors of the tail move; the kernel reads the actual count from the control word
            const unsigned mblocks = (unsigned)(((size_t)nlive + SK_EVENT_BLOCK - 1) / SK_EVENT_BLOCK);
            sk_wf_move<<<std::max(mblocks, 1u), SK_EVENT_BLOCK, 0, e->stream>>>(K, e->bank_fields_d, e->bank_fields_i);
            CK(cudaGetLastError());
            e->launches_total += 2;
            e->bank.n = n_new;
            eblocks = (unsigned)((K.n + SK_EVENT_BLOCK - 1) / SK_EVENT_BLOCK);
        }
    }
    return SK_OK;
}

extern "C" int sk_engine_set_history_interleave(sk_engine_t* e, uint64_t block, uint32_t num_parts, uint32_t part)
{
    if (!e) return fail(SK_ERR_INVALID, "null engine");
    if (num_parts < 1 || part >= num_parts) return fail(SK_ERR_INVALID, "part must be below num_parts");
    if (num_parts > 1 && (block == 0 || (block & (block - 1)))) return fail(SK_ERR_INVALID, "block must be a power of two");
    e->il_block = block;
    e->il_parts = num_parts;
    e->il_part = part;
    return SK_OK;
}

extern "C" int sk_engine_launch_segment(sk_engine_t* e, uint64_t first, uint64_t count, int32_t primary,
                                        int32_t peel, int32_t store, uint32_t stream_id)
{
    if (!e) return fail(SK_ERR_INVALID, "null engine");
    if (!e->grid_kind || !e->M.ncells || !e->M.nlam) return fail(SK_ERR_STATE, "engine is not fully configured");
    if (primary && !e->M.nsrc) return fail(SK_ERR_STATE, "engine is not fully configured");
    if (primary && !e->npackets) return fail(SK_ERR_STATE, "call sk_engine_prepare_primary first");
    if (!primary && !e->secondary_ready) return fail(SK_ERR_STATE, "call sk_engine_prepare_secondary first");
    if (e->num_mixes != e->M.nmed) return fail(SK_ERR_STATE, "the number of dust mixes does not match the number of medium components");
    if (!primary && e->sec_num_media != e->M.nmed) return fail(SK_ERR_STATE, "emission tables do not match the medium components");
    if (store && e->M.rf_grid < 0) return fail(SK_ERR_STATE, "no radiation field grid configured");
    if (store && !e->M.force_scattering)
        return fail(SK_ERR_INVALID, "storing the radiation field requires forced scattering (Configuration.cpp:476-482)");
    if (int rc_bind = bind(e)) return rc_bind;
    if (!e->num_sms) CK(cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, e->cfg.device));
    // (An L2 persisting window over the cell records was measured to make no difference -- the 30 MB of records stay in
    //  the 126 MB L2 on their own -- and cudaGetDeviceProperties / cudaDeviceSetLimit cost milliseconds per engine.)
    CK(cudaEventRecord(e->ev0, e->stream));
    e->stage_of_pair.clear();
    // this engine's share of [first, first+count): everything, or every il_parts-th block of il_block histories
    unsigned long long share = count;
    if (count && e->il_parts > 1)
    {
        const unsigned long long B = e->il_block, cycle = B * e->il_parts, rem = count % cycle, lo = (unsigned long long)e->il_part * B;
        share = count / cycle * B + (rem > lo ? std::min<unsigned long long>(rem - lo, B) : 0ull);
    }
    if (e->M.kin && !e->M.vel)
    {
        // only the sources move: the walks with kinematics read a velocity per cell all the same
        double4* v;
        if (int rc = dalloc_zero(e->vel_allocs, (size_t)e->M.ncells, &v)) return rc;
        e->M.vel = v;
    }
    if (share)
    {
        if (int rc = ensure_bank(e, share)) return rc;
        if (e->pool_chunks)
        {
            // every chunk is free at the start of a segment (all histories of the previous one have ended)
            sk_pool_reset_kernel<<<(e->pool_chunks + 255) / 256, 256, 0, e->stream>>>(e->bank.pool_free, e->bank.pool_ctl,
                                                                                    e->pool_chunks);
            CK(cudaGetLastError());
        }
        SkRunArgs A;
        A.first = first;
        A.count = share;
        A.il_shift = -1;
        A.il_stride = A.il_offset = 0;
        if (e->il_parts > 1)
        {
            A.il_shift = 0;
            while ((1ull << A.il_shift) < e->il_block) A.il_shift++;
            A.il_stride = e->il_block * e->il_parts;
            A.il_offset = (unsigned long long)e->il_part * e->il_block;
        }
        A.primary = primary;
        A.peel = peel && !e->instr.empty();
        A.store = store;
        A.stream_id = stream_id;
        A.work_counter = e->work_counter;
        CK(cudaMemsetAsync(e->work_counter, 0, sizeof(unsigned long long), e->stream));
        if (!e->model_dev) CK(dev_malloc(&e->model_dev, sizeof(SkDevModel)));
        CK(cudaMemcpyAsync(e->model_dev, &e->M, sizeof(SkDevModel), cudaMemcpyHostToDevice, e->stream));
        A.model = e->model_dev;
        int rc = e->grid_kind == 1 ? run_bank<1>(e, A) : e->grid_kind == 2 ? run_bank<2>(e, A) : run_bank<3>(e, A);
        if (rc) return rc;
        if (e->pool_chunks)
        {
            int ctl[2] = {0, 0};
            CK(cudaMemcpyAsync(ctl, e->bank.pool_ctl, sizeof ctl, cudaMemcpyDeviceToHost, e->stream));
            CK(cudaStreamSynchronize(e->stream));
            e->pixel_overflows += (unsigned long long)ctl[1];
        }
    }
    CK(cudaEventRecord(e->ev1, e->stream));
    e->timing_pending = true;
    return SK_OK;
}

extern "C" int sk_engine_synchronize(sk_engine_t* e)
{
    if (!e) return fail(SK_ERR_INVALID, "null engine");
    if (int rc_bind = bind(e)) return rc_bind;
    CK(cudaStreamSynchronize(e->stream));
    if (e->timing_pending)
    {
        CK(cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1));
        for (int k = 0; k < SK_STAGE_COUNT; ++k) e->stage_ms[k] = 0.f;
        for (size_t i = 0; i < e->stage_of_pair.size(); ++i)
        {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, e->stage_events[2 * i], e->stage_events[2 * i + 1]));
            e->stage_ms[e->stage_of_pair[i]] += ms;
        }
        e->timing_pending = false;
    }
    return SK_OK;
}

extern "C" int sk_engine_run_segment(sk_engine_t* e, uint64_t first, uint64_t count, int32_t primary, int32_t peel,
                                     int32_t store, uint32_t stream_id)
{
    if (int rc = sk_engine_launch_segment(e, first, count, primary, peel, store, stream_id)) return rc;
    return sk_engine_synchronize(e);
}

extern "C" int sk_engine_last_kernel_ms(sk_engine_t* e, float* ms)
{
    if (!e || !ms) return fail(SK_ERR_INVALID, "null argument");
    *ms = e->last_ms;
    return SK_OK;
}

extern "C" int sk_engine_last_stage_ms(sk_engine_t* e, float out[SK_STAGE_COUNT])
{
    if (!e || !out) return fail(SK_ERR_INVALID, "null argument");
    for (int k = 0; k < SK_STAGE_COUNT; ++k) out[k] = e->stage_ms[k];
    return SK_OK;
}

extern "C" int sk_engine_communicate_rf(sk_engine_t* e, int32_t primary)
{
    if (!e) return fail(SK_ERR_INVALID, "null engine");
    if (int rc_bind = bind(e)) return rc_bind;
    if (!primary && e->M.rf2)
    {
        CK(cudaMemcpyAsync(e->M.rf2, e->M.rf2c, (size_t)e->M.ncells * e->M.nrf * sizeof(double),
                           cudaMemcpyDeviceToDevice, e->stream));
        CK(cudaStreamSynchronize(e->stream));
    }
    return SK_OK;
}

extern "C" int sk_engine_absorbed_luminosity(sk_engine_t* e, int32_t primary, double* out)
{
    if (!e || !out) return fail(SK_ERR_INVALID, "null argument");
    if (e->M.rf_grid < 0) return fail(SK_ERR_STATE, "no radiation field");
    if (int rc_bind = bind(e)) return rc_bind;
    // kappa_abs per RF bin: DustMix::sectionAbs(lambda_ell) = _sigmaabsv[indexForLambda(lambda_ell)]
    const std::vector<double>& lam = e->wlg_lambda[e->M.rf_grid];
    if (e->num_mixes != e->M.nmed) return fail(SK_ERR_STATE, "the number of dust mixes does not match the number of medium components");
    std::vector<double> kabs(lam.size() * (size_t)e->M.nmed);
    for (int h = 0; h < e->M.nmed; ++h)
        for (size_t i = 0; i < lam.size(); ++i)
            kabs[(size_t)h * lam.size() + i] = e->dust_sig_abs[(size_t)h * e->M.nlam + dust_index_for_lambda(e, lam[i])];
    double* dk = nullptr;
    CK(dev_malloc(&dk, kabs.size() * sizeof(double)));
    CK(cudaMemcpyAsync(dk, kabs.data(), kabs.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemsetAsync(e->scalar, 0, sizeof(double), e->stream));
    sk_absorbed_kernel<<<296, 256, 0, e->stream>>>(primary ? e->M.rf1 : e->M.rf2, e->M.dens, e->M.cells, e->M.vrec, e->M.densx,
                                                   e->M.nmed, dk, e->M.ncells, e->M.nrf, e->scalar);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, e->scalar, sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    dev_free(dk);
    return SK_OK;
}

static int fetch_doubles(sk_engine* e, const double* dev, size_t n, double* out);
extern "C" int sk_engine_read_rf(sk_engine_t* e, int32_t which, double* out)
{
    if (!e || !out) return fail(SK_ERR_INVALID, "null argument");
    const double* src = which == 0 ? e->M.rf1 : which == 1 ? e->M.rf2 : e->M.rf2c;
    if (!src) return fail(SK_ERR_STATE, "no radiation field");
    if (int rc_bind = bind(e)) return rc_bind;
    // the device keeps the tables wavelength-major (SK_RF_INDEX): transpose into scratch, one transfer
    const size_t len = (size_t)e->M.ncells * e->M.nrf;
    if (e->scratch_len < len)
    {
        dev_free(e->scratch);
        e->scratch = nullptr;
        e->scratch_len = 0;
        CK(dev_malloc(&e->scratch, len * sizeof(double)));
        e->scratch_len = len;
    }
    dim3 grid((unsigned)((e->M.ncells + 31) / 32), (unsigned)((e->M.nrf + 31) / 32));
    sk_rf_transpose_kernel<<<grid, dim3(32, 8), 0, e->stream>>>(src, e->scratch, e->M.ncells, e->M.nrf);
    CK(cudaGetLastError());
    return fetch_doubles(e, e->scratch, len, out);
}

// device -> caller buffer through a pinned staging buffer (pageable destinations otherwise crawl at page-fault speed)
static int fetch_doubles(sk_engine* e, const double* dev, size_t n, double* out)
{
    const size_t bytes = n * sizeof(double);
    bool direct = bytes < ((size_t)1 << 16);
    if (!direct)
    {
        // a page-locked destination (cudaHostAlloc / cudaHostRegister by the caller) takes the DMA directly
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, out) == cudaSuccess)
            direct = attr.type == cudaMemoryTypeHost;
        else
            cudaGetLastError();
    }
    if (direct)
    {
        CK(cudaMemcpyAsync(out, dev, bytes, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        return SK_OK;
    }
    // staging buffer of two chunks, sized to the transfer (page-locking 64 MB for a small read costs more than the read)
    const size_t chunk = std::min<size_t>((size_t)32 << 20, (bytes + 4095) / 4096 * 4096);
    if (e->pinned_bytes < 2 * chunk)
    {
        if (e->pinned) cudaFreeHost(e->pinned);
        e->pinned = nullptr;
        e->pinned_bytes = 0;
        CK(cudaMallocHost(&e->pinned, 2 * chunk));
        e->pinned_bytes = 2 * chunk;
    }
    if (!e->pin_ev[0])
    {
        CK(cudaEventCreateWithFlags(&e->pin_ev[0], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&e->pin_ev[1], cudaEventDisableTiming));
    }
    // double-buffered: the copy of chunk i+1 over PCIe overlaps the host memcpy of chunk i
    size_t nchunks = (bytes + chunk - 1) / chunk;
    for (size_t i = 0; i <= nchunks; ++i)
    {
        if (i < nchunks)
        {
            size_t off = i * chunk, len = std::min(chunk, bytes - off);
            CK(cudaMemcpyAsync((char*)e->pinned + (i & 1) * chunk, (const char*)dev + off, len, cudaMemcpyDeviceToHost,
                               e->stream));
            CK(cudaEventRecord(e->pin_ev[i & 1], e->stream));
        }
        if (i > 0)
        {
            size_t j = i - 1, off = j * chunk, len = std::min(chunk, bytes - off);
            CK(cudaEventSynchronize(e->pin_ev[j & 1]));
            memcpy((char*)out + off, (char*)e->pinned + (j & 1) * chunk, len);
        }
    }
    return SK_OK;
}

static int read_array(sk_engine* e, int instrument, int component, bool ifu, double* out)
{
    if (!e || !out || instrument < 0 || instrument >= (int)e->instr.size() || component < 0 || component >= SK_NUM_COMP)
        return fail(SK_ERR_INVALID, "bad instrument/component");
    if (int rc_bind = bind(e)) return rc_bind;
    const HostInstr& q = e->instr[instrument];
    if (ifu ? !q.include_ifu : !q.include_sed) return fail(SK_ERR_INVALID, "instrument does not record this");
    size_t len = ifu ? q.npix * q.nl : (size_t)q.nl;
    const long long* off = ifu ? q.ifu_off : q.sed_off;
    if (component == SK_COMP_TOTAL && !q.record_total_only)
    {
        // FluxRecorder::calibrateAndWrite: total = direct + scattered (+ secondary), FluxRecorder.cpp:522-526; summed on
        // the device into scratch, one transfer
        if (e->scratch_len < len)
        {
            dev_free(e->scratch);
            e->scratch = nullptr;
            e->scratch_len = 0;
            CK(dev_malloc(&e->scratch, len * sizeof(double)));
            e->scratch_len = len;
        }
        const bool sec = off[SK_COMP_SECONDARY_DIRECT] >= 0;
        unsigned blocks = (unsigned)std::min<size_t>((len + 255) / 256, 4096);
        sk_sum_kernel<<<blocks, 256, 0, e->stream>>>(e->scratch, e->det_block + off[SK_COMP_PRIMARY_DIRECT],
                                                     e->det_block + off[SK_COMP_PRIMARY_SCATTERED],
                                                     sec ? e->det_block + off[SK_COMP_SECONDARY_DIRECT] : nullptr,
                                                     sec ? e->det_block + off[SK_COMP_SECONDARY_SCATTERED] : nullptr, len);
        CK(cudaGetLastError());
        return fetch_doubles(e, e->scratch, len, out);
    }
    if (off[component] < 0) return fail(SK_ERR_INVALID, "component not recorded");
    return fetch_doubles(e, e->det_block + off[component], len, out);
}
extern "C" int sk_engine_read_sed(sk_engine_t* e, int32_t instrument, int32_t component, double* out)
{
    return read_array(e, instrument, component, false, out);
}
extern "C" int sk_engine_read_ifu(sk_engine_t* e, int32_t instrument, int32_t component, double* out)
{
    return read_array(e, instrument, component, true, out);
}
extern "C" int sk_engine_read_sed_stats(sk_engine_t* e, int32_t instrument, int32_t k, double* out)
{
    if (!e || !out || instrument < 0 || instrument >= (int)e->instr.size() || k < 0 || k > 4)
        return fail(SK_ERR_INVALID, "bad instrument/power");
    const HostInstr& q = e->instr[instrument];
    if (q.wsed_off[k] < 0) return fail(SK_ERR_INVALID, "statistics not recorded");
    if (int rc_bind = bind(e)) return rc_bind;
    CK(cudaMemcpyAsync(out, e->stat_block + q.wsed_off[k], (size_t)q.nl * sizeof(double), cudaMemcpyDeviceToHost,
                       e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return SK_OK;
}

extern "C" int sk_engine_read_ifu_stats(sk_engine_t* e, int32_t instrument, int32_t k, double* out)
{
    if (!e || !out || instrument < 0 || instrument >= (int)e->instr.size() || k < 0 || k > 4)
        return fail(SK_ERR_INVALID, "bad instrument/power");
    const HostInstr& q = e->instr[instrument];
    if (q.wifu_off[k] < 0) return fail(SK_ERR_INVALID, "statistics not recorded");
    if (int rc_bind = bind(e)) return rc_bind;
    return fetch_doubles(e, e->stat_block + q.wifu_off[k], q.npix * (size_t)q.nl, out);
}

extern "C" int sk_engine_counters(sk_engine_t* e, sk_counters_t* out, int32_t reset)
{
    if (!e || !out) return fail(SK_ERR_INVALID, "null argument");
    if (int rc_bind = bind(e)) return rc_bind;
    unsigned long long c[16];
    CK(cudaMemcpyAsync(c, e->M.counters, sizeof c, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    memset(out, 0, sizeof *out);
    out->packets = c[0];
    out->forward_paths = c[1];
    out->forward_segments = c[2];
    out->replay_segments = c[3];
    out->peel_paths = c[4];
    out->peel_segments = c[5];
    out->scatterings = c[6];
    out->rf_deposits = c[7];
    out->detections = c[8];
    out->fallbacks = c[9];
    out->kernel_launches = e->launches_total;
    out->rounds = e->rounds_total;
    out->pixel_overflows = e->pixel_overflows;
    if (reset)
    {
        CK(cudaMemsetAsync(e->M.counters, 0, sizeof c, e->stream));
        e->launches_total = e->rounds_total = 0;
        e->pixel_overflows = 0;
    }
    return SK_OK;
}

// ---------------------------------------------------------------------------------------------------
// Measured ceiling of the crossing loop's memory access pattern (bench.py's second roofline): every lane follows a
// chain of dependent 32-byte record fetches (one 256-bit load per step, the next index comes out of the record) through
// a table of `num_records` records linked in a random cycle -- the cell-record gather of the trace kernels without any
// of their arithmetic, at full occupancy.  records/s of this kernel bounds the cell crossings/s of ANY walk that needs
// one dependent record per crossing on this device.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sk_gather_chain_kernel(const SkCellRec* __restrict__ table, int num_records, int steps,
                                                               unsigned long long* sink)
{
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    int m = (int)(((unsigned long long)tid * 2654435761ull) % (unsigned long long)num_records);
    double acc = 0.;
    for (int i = 0; i < steps; ++i)
    {
        int4 a, b;
        sk_ld256(&table[m], a, b);
        acc += __hiloint2double(a.y, a.x);
        m = a.z;
    }
    if (acc == 12345.678) atomicAdd(sink, (unsigned long long)m);  // keeps the chain alive
}
__global__ void sk_gather_init_kernel(SkCellRec* table, int num_records, unsigned long long mult, unsigned long long add)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= num_records) return;
    SkCellRec r;
    r.dens = 1.;
    // an affine map modulo num_records with an odd multiplier coprime to it scatters successive fetches over the table
    r.link[0] = (int)(((unsigned long long)m * mult + add) % (unsigned long long)num_records);
    for (int w = 1; w < 6; ++w) r.link[w] = -1;
    table[m] = r;
}
extern "C" int sk_engine_measure_gather_peak(sk_engine_t* e, int32_t num_records, double* records_per_s)
{
    if (!e || !records_per_s || num_records < 1024) return fail(SK_ERR_INVALID, "bad argument");
    if (int rc_bind = bind(e)) return rc_bind;
    if (!e->num_sms) CK(cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, e->cfg.device));
    SkCellRec* table = nullptr;
    unsigned long long* sink = nullptr;
    CK(dev_malloc(&table, (size_t)num_records * sizeof(SkCellRec)));
    CK(dev_malloc(&sink, sizeof(unsigned long long)));
    unsigned long long mult = 2654435761ull;
    while (std::gcd(mult, (unsigned long long)num_records) != 1) mult += 2;
    sk_gather_init_kernel<<<(num_records + 255) / 256, 256, 0, e->stream>>>(table, num_records, mult, 40503ull);
    CK(cudaGetLastError());
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    const int steps = 512;
    double best = 0.;
    // 8 resident blocks of 256 threads per SM = full occupancy; two warm-up launches, best of five
    const unsigned blocks = (unsigned)e->num_sms * 8u;
    for (int it = 0; it < 7; ++it)
    {
        CK(cudaEventRecord(a, e->stream));
        sk_gather_chain_kernel<<<blocks, 256, 0, e->stream>>>(table, num_records, steps, sink);
        CK(cudaGetLastError());
        CK(cudaEventRecord(b, e->stream));
        CK(cudaEventSynchronize(b));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (it >= 2) best = std::max(best, (double)blocks * 256. * steps / (ms * 1e-3));
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    dev_free(table);
    dev_free(sink);
    *records_per_s = best;
    return SK_OK;
}

extern "C" int sk_engine_device_buffer(sk_engine_t* e, int32_t which, void** device_ptr, uint64_t* num_doubles)
{
    if (!e || !device_ptr || !num_doubles) return fail(SK_ERR_INVALID, "null argument");
    size_t rfcount = (size_t)e->M.ncells * e->M.nrf;
    switch (which)
    {
        case 0: *device_ptr = e->M.rf1; *num_doubles = rfcount; break;
        case 1: *device_ptr = e->M.rf2; *num_doubles = rfcount; break;
        case 2: *device_ptr = e->M.rf2c; *num_doubles = rfcount; break;
        case 3: *device_ptr = e->det_block; *num_doubles = e->det_count; break;
        case 4: *device_ptr = e->stat_block; *num_doubles = e->stat_count; break;
        default: return fail(SK_ERR_INVALID, "unknown buffer");
    }
    return SK_OK;
}
