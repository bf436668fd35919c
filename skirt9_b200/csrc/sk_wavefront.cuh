// sk_wavefront.cuh -- the photon life cycle as a sequence of stage kernels over a bank of in-flight packets:
// MonteCarloSimulation::performLifeCycle (SKIRT/core/MonteCarloSimulation.cpp:538-613) and everything it calls.
//
// Execution model.  The reference runs one history at a time per thread; here a BANK of up to
// `cap` histories (millions) is in flight per GPU, its state held as a structure of arrays in HBM, and the body of the
// reference's forced-scattering loop (.cpp:565-585) is cut at its three grid traversals into stage kernels that each
// advance the whole bank by one step:
//     advance  = [interaction of the previous round: .cpp:724-741,576-580] + [launch into free slots: SourceSystem::launch]
//                + [peel-off set-up towards the first observer: .cpp:617-634 / 784-842]
//     trace<2> = optical depth to the observer for every peel-off ray          (MediumSystem.cpp:1192-1219)
//     detect   = FluxRecorder::detect for the observer group; after the last group the scattering itself
//                (MediumSystem.cpp:796-823) and the list of forward rays
//     trace<0> = forward path to the grid boundary, RF deposits fused in       (MediumSystem.cpp:849-871, .cpp:638-665)
//     sample   = interaction optical depth with path-length biasing            (.cpp:696-722)
//     trace<1> = walk to the interaction point                                 (SpatialGridPath.cpp:164-206)
// Event kernels (advance, detect, sample) are element-wise over the bank: every thread owns one slot, all threads run
// the same short code, state is read and written coalesced.  Trace kernels are persistent: lanes are NOT bound to
// packets; each lane walks one ray cell by cell and, when its ray ends, takes the next unprocessed ray of the stage
// from a global list (chunks of SK_CHUNK rays reserved per warp with one atomic) -> the cell-crossing loop stays
// converged although path lengths vary from 1 to several hundred segments, each kernel is small enough to live in
// the instruction cache, and each gets the register budget / occupancy that suits it.
// Paths are never materialised (the reference stores vector<Segment>, SpatialGridPath.hpp:93-115): the forward path
// is walked once for the total optical depth and re-walked up to the sampled interaction point; both walks use
// identical arithmetic so they see identical segments.
#pragma once
#include "sk_blocks.cuh"

#ifndef SK_EVENT_BLOCK
#define SK_EVENT_BLOCK 256
#endif
#ifndef SK_TRACE_BLOCK
#define SK_TRACE_BLOCK 128
#endif
#ifndef SK_TRACE_MINBLOCKS
#define SK_TRACE_MINBLOCKS 8        // the walks wait on L2 latency: resident warps count for more than a few spilled registers
#endif
#ifndef SK_TRACE_MINBLOCKS_STORE
#define SK_TRACE_MINBLOCKS_STORE 6  // forward walk with radiation-field deposits: exp + logarithmic mean per segment
#endif
#ifndef SK_TRACE_MINBLOCKS_VORONOI
#define SK_TRACE_MINBLOCKS_VORONOI 8  // (measured: 8 blocks with two neighbour records in flight beat 5 blocks with four)
#endif
#ifndef SK_TRACE_MINBLOCKS_PEEL
#define SK_TRACE_MINBLOCKS_PEEL 10  // the peel-off kernel keeps its (shared) direction in parameter space: fewer registers
#endif
#ifndef SK_TRACE_MINBLOCKS_MULTI_LESS
#define SK_TRACE_MINBLOCKS_MULTI_LESS 1  // several components: one resident block fewer pays for the extra lane state (measured)
#endif
#ifndef SK_TRACE_MINBLOCKS_KIN_LESS
#define SK_TRACE_MINBLOCKS_KIN_LESS 1  // kinematics: perceived wavelength, table index and k.v of the next cell in lane registers (measured: 1 beats 2 and 3)
#endif
#ifndef SK_CHUNK
#define SK_CHUNK 64
#endif
#ifndef SK_REFILL_MIN
#define SK_REFILL_MIN 6
#endif
#ifndef SK_REFILL_MIN_FUSED
#define SK_REFILL_MIN_FUSED 10  // forward + interaction walk: two path ends per packet, the service block runs twice as often
#endif
#ifndef SK_REFILL_MIN_SHORT
#define SK_REFILL_MIN_SHORT 12  // walks to the interaction point are short (a dozen cells): refill threshold of their own
#endif

// ---------------------------------------------------------------------------------------------------
// small helpers shared by the stage kernels
// ---------------------------------------------------------------------------------------------------
// Appends `slot` of every thread with `want` to a list whose length is the control word `ctl_word` -- ONE atomic per
// block: all warps of all SMs adding to the same address is what otherwise bounds the element-wise kernels (the L2 atomic
// unit serialises per address).  Every thread of the block must call (uses two block barriers); `nlive`, when not null,
// is a census counter that receives the block's number of `count_live` threads in the same pass.
__device__ __forceinline__ void sk_block_append(int32_t* list, unsigned int* ctl_word, bool want, int slot,
                                                unsigned int* nlive = nullptr, bool count_live = false,
                                                unsigned long long* live_counter = nullptr)
{
    __shared__ unsigned int s_cnt[2][SK_EVENT_BLOCK / 32];
    __shared__ unsigned int s_base;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned mask = __ballot_sync(0xffffffffu, want);
    const unsigned live = nlive ? __ballot_sync(0xffffffffu, count_live) : 0u;
    if (lane == 0)
    {
        s_cnt[0][warp] = __popc(mask);
        s_cnt[1][warp] = __popc(live);
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        unsigned total = 0, tlive = 0;
        for (int w = 0; w < SK_EVENT_BLOCK / 32; ++w)
        {
            const unsigned c = s_cnt[0][w];
            s_cnt[0][w] = total;  // exclusive prefix
            total += c;
            tlive += s_cnt[1][w];
        }
        s_base = total ? atomicAdd(ctl_word, total) : 0u;
        if (nlive && tlive) atomicAdd(nlive, tlive);
        if (live_counter && tlive) atomicAdd(live_counter, (unsigned long long)tlive);
    }
    __syncthreads();
    if (want) list[s_base + s_cnt[0][warp] + __popc(mask & ((1u << lane) - 1u))] = slot;
    __syncthreads();  // the shared words are reused by the next call
}

// Reserves one index of a 64-bit counter for every thread with `want` (the history dispenser), one atomic per block.
__device__ __forceinline__ unsigned long long sk_block_reserve(unsigned long long* counter, bool want)
{
    __shared__ unsigned int r_cnt[SK_EVENT_BLOCK / 32];
    __shared__ unsigned long long r_base;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned mask = __ballot_sync(0xffffffffu, want);
    if (lane == 0) r_cnt[warp] = __popc(mask);
    __syncthreads();
    if (threadIdx.x == 0)
    {
        unsigned total = 0;
        for (int w = 0; w < SK_EVENT_BLOCK / 32; ++w)
        {
            const unsigned c = r_cnt[w];
            r_cnt[w] = total;
            total += c;
        }
        r_base = total ? atomicAdd(counter, (unsigned long long)total) : 0ull;
    }
    __syncthreads();
    const unsigned long long h = r_base + r_cnt[warp] + __popc(mask & ((1u << lane) - 1u));
    __syncthreads();
    return h;
}

__device__ __forceinline__ void sk_flush_counters(const SkDevModel& M, SkLocalCounters& cnt)
{
    unsigned int* c = reinterpret_cast<unsigned int*>(&cnt);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(SkLocalCounters) / sizeof(unsigned int)); ++i)
    {
        unsigned int v = __reduce_add_sync(0xffffffffu, c[i]);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&M.counters[i], (unsigned long long)v);
    }
}

__device__ __forceinline__ void sk_rng_load(SkRng& g, const SkDevModel& M, const SkRunArgs& A, const SkBank& K, int slot)
{
    sk_rng_init(g, M.seed, A.stream_id,
                ((unsigned long long)(uint32_t)K.I(I_HHI, slot) << 32) | (uint32_t)K.I(I_HLO, slot),
                (uint32_t)K.I(I_DRAW, slot));
}

// (to_sed: the entry is a wavelength bin of the SED statistics instead of a pixel and bin of the frame statistics)
__device__ __forceinline__ void sk_record_pixel_stats(const SkDevInstr& q, int lell, double w, bool to_sed = false)
{
    double wn = 1.;
    for (int kk = 0; kk <= 4; ++kk)
    {
        atomicAdd(to_sed ? &q.wsed[kk][lell] : &q.wifu[kk][lell], wn);
        wn *= w;
    }
}
// A detection's contribution to the history's pixel list (ContributionList::addContribution, FluxRecorder.hpp:331, with the
// grouping of FluxRecorder.cpp:996-999 done on insertion): the first SK_PIX_K distinct (pixel, bin) entries live in the
// history's bank slot, further ones in chunks of SK_PIX_C entries chained from a pool (newest chunk first).  Chunks are only
// taken here (detect kernel) and only given back by sk_finish_history (advance kernel), never in the same kernel, so one
// atomic counter is the whole stack discipline.  Should the pool run dry the contribution is recorded on its own
// (counted in sk_counters_t::pixel_overflows; the pool is sized so that this does not happen).
#define SK_PIX_INTS (SK_PIX_K + 2)
// With kinematics the same lists, keyed by wavelength bin, hold the history's contributions to the SED statistics (to_sed,
// list q.sed_slot): its peel-off packets then differ in wavelength (FluxRecorder.cpp:962-986).
__device__ __forceinline__ void sk_add_pixel_contribution(const SkDevModel& M, const SkBank& K, const SkDevInstr& q,
                                                          int slot, int lell, double w, bool to_sed = false)
{
    const int ls = to_sed ? q.sed_slot : q.pix_slot;
    const int fi = M.pix_base_i + ls * SK_PIX_INTS, fd = M.pix_base_d + ls * SK_PIX_K;
    const int total = K.I(fi + SK_PIX_K, slot);
    const int n = min(total, SK_PIX_K);
    for (int i = 0; i < n; ++i)
        if (K.I(fi + i, slot) == lell)
        {
            K.D(fd + i, slot) += w;
            return;
        }
    if (total < SK_PIX_K)
    {
        K.I(fi + total, slot) = lell;
        K.D(fd + total, slot) = w;
        K.I(fi + SK_PIX_K, slot) = total + 1;
        return;
    }
    int head = K.I(fi + SK_PIX_K + 1, slot);
    const int rem = total - SK_PIX_K;  // entries in the chain; the newest chunk holds ((rem - 1) % C) + 1 of them
    int cnt = rem ? ((rem - 1) % SK_PIX_C) + 1 : 0;
    for (int c = head; c >= 0 && rem; c = K.pool_next[c], cnt = SK_PIX_C)
        for (int j = 0; j < cnt; ++j)
            if (K.pool_lell[(size_t)c * SK_PIX_C + j] == lell)
            {
                K.pool_w[(size_t)c * SK_PIX_C + j] += w;
                return;
            }
    const int at = rem % SK_PIX_C;
    if (at == 0)
    {
        const int top = atomicSub(&K.pool_ctl[0], 1) - 1;
        if (top < 0)
        {
            atomicAdd(&K.pool_ctl[0], 1);
            atomicAdd(&K.pool_ctl[1], 1);
            sk_record_pixel_stats(q, lell, w, to_sed);
            return;
        }
        const int c = K.pool_free[top];
        K.pool_next[c] = head;
        head = c;
        K.I(fi + SK_PIX_K + 1, slot) = c;
    }
    K.pool_lell[(size_t)head * SK_PIX_C + at] = lell;
    K.pool_w[(size_t)head * SK_PIX_C + at] = w;
    K.I(fi + SK_PIX_K, slot) = total + 1;
}

// Ends a history: FluxRecorder::recordContributions for the SED arrays (FluxRecorder.cpp:962-986); frees the slot.
__device__ __forceinline__ void sk_finish_history(const SkDevModel& M, const SkBank& K, int slot, bool kin = false)
{
    for (int j = 0; j < M.ninstr; ++j)
    {
        const SkDevInstr& q = M.instr[j];
        if (!q.record_stats) continue;
        if (kin && q.sed_slot >= 0) continue;  // (its SED bins are in a list, below)
        int ell = K.I(I_HELL0 + j, slot);
        if (ell >= 0)
        {
            double w = K.D(D_HISTW0 + j, slot);
            double wn = 1.;
            for (int kk = 0; kk <= 4; ++kk)
            {
                atomicAdd(&q.wsed[kk][ell], wn);
                wn *= w;
            }
        }
    }
    // the same per frame pixel (FluxRecorder.cpp:990-1013): the history's list holds one entry per pixel and bin
    for (int jj = 0; jj < (kin ? 2 : 1) * M.ninstr; ++jj)
    {
        const bool to_sed = jj >= M.ninstr;  // second pass (kinematics): the lists of SED bins
        const SkDevInstr& q = M.instr[to_sed ? jj - M.ninstr : jj];
        const int ls = to_sed ? q.sed_slot : q.pix_slot;
        if (ls < 0) continue;
        const int fi = M.pix_base_i + ls * SK_PIX_INTS, fd = M.pix_base_d + ls * SK_PIX_K;
        const int total = K.I(fi + SK_PIX_K, slot);
        const int n = min(total, SK_PIX_K);
        for (int i = 0; i < n; ++i) sk_record_pixel_stats(q, K.I(fi + i, slot), K.D(fd + i, slot), to_sed);
        const int rem = total - SK_PIX_K;
        if (rem > 0)
        {
            int cnt = ((rem - 1) % SK_PIX_C) + 1;
            for (int c = K.I(fi + SK_PIX_K + 1, slot); c >= 0; cnt = SK_PIX_C)
            {
                for (int i = 0; i < cnt; ++i)
                    sk_record_pixel_stats(q, K.pool_lell[(size_t)c * SK_PIX_C + i], K.pool_w[(size_t)c * SK_PIX_C + i], to_sed);
                const int next = K.pool_next[c];
                K.pool_free[atomicAdd(&K.pool_ctl[0], 1)] = c;  // back on the stack of free chunks
                c = next;
            }
        }
    }
    K.I(I_STATE, slot) = 0;
}

// ---------------------------------------------------------------------------------------------------
// Forced scattering: the interaction optical depth, MonteCarloSimulation::simulateForcedPropagation (.cpp:696-722) with
// Random::exponCutoff (Random.cpp:105-117) and the path-length bias xi.  The reference draws when the forward path is
// complete; here the work is cut in three so that only a few instructions run in the trace kernel, where the lanes that
// have just finished a path are a small part of their warp:
//   sk_predraw_interaction  (launch / detect kernels, full warps) draws the deviates at the place in the history's random
//                           sequence where the reference draws them -- nothing else draws in between -- and packs them
//                           into one number: -u when the biased (linear) branch was chosen, +u otherwise
//   sk_interaction_depth    (trace kernel, at the end of the forward path) turns it into tau_int for the path's tau_path
//   sk_wf_advance           applies the bias weight p/q to the packet
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double sk_predraw_interaction(SkRng& g, double xi)
{
    if (xi == 0.) return sk_uniform(g);
    const bool biased = sk_uniform(g) < xi;
    const double u = sk_uniform(g);
    return biased ? -u : u;
}
// Random::exponCutoff's rejection loop (x > xmax can only come from rounding when u is within an ulp of 1): further
// deviates of the history, drawn here.  Arguments by value so that the kernel's parameter copies never have their address
// taken.
__device__ __noinline__ double sk_redraw_interaction(uint32_t seed, uint32_t stream_id, int32_t* bank_i, int cap, int slot,
                                                     double xmax)
{
    SkBank K;
    K.i = bank_i;
    K.cap = cap;
    SkRng g;
    sk_rng_init(g, seed, stream_id, ((unsigned long long)(uint32_t)K.I(I_HHI, slot) << 32) | (uint32_t)K.I(I_HLO, slot),
                (uint32_t)K.I(I_DRAW, slot));
    double x;
    do
        x = -log(1.0 - sk_uniform(g) * (1.0 - exp(-xmax)));
    while (x > xmax);
    K.I(I_DRAW, slot) = (int)g.draw;
    return x;
}
__device__ __forceinline__ double sk_interaction_depth(double u, double taupath, bool& redraw)
{
    redraw = false;
    if (u < 0.) return -u * taupath;        // uniform * taupath (.cpp:712)
    if (taupath < 1e-10) return u * taupath;  // Random.cpp:109-110
    const double x = -log(1.0 - u * (1.0 - exp(-taupath)));
    redraw = x > taupath;
    return x;
}

// ---------------------------------------------------------------------------------------------------
// Trace kernels: walk all rays of the list with dynamic lane refill.
//   MODE 0  forced scattering: the forward path to the boundary -- MediumSystem::setExtinctionOpticalDepths
//           (MediumSystem.cpp:849-871), with STORE fused with MonteCarloSimulation::storeRadiationField (.cpp:638-665) --
//           then, in the same lane, the interaction optical depth (simulateForcedPropagation, .cpp:696-722) and the walk
//           to the interaction point, SpatialGridPath::findInteractionPoint (SpatialGridPath.cpp:164-206): paths are
//           not stored, the second walk repeats the first with identical arithmetic while the cell records are cache-hot
//   MODE 1  non-forced scattering: walk to the interaction point drawn by sk_wf_sample,
//           MediumSystem::setInteractionPointUsingExtinction (MediumSystem.cpp:978-1010)
//   MODE 2  optical depth to the observer: MediumSystem::getExtinctionOpticalDepth (MediumSystem.cpp:1192-1219)
// Structure: a compact inner loop that only crosses cells, and an outer service block (store the results of finished
// rays, load new rays) that is entered when at least SK_REFILL_MIN lanes are idle, so that its cost is shared.
// ---------------------------------------------------------------------------------------------------
// MULTI: several medium components with their own mixes (MediumSystem.cpp:874-885, 1012-1040, 1222-1240): the opacity of a
// cell is the sum of sigma_h n_h; the sections of the components beyond the first live in lane registers and their
// densities are fetched from densx once the cell is known.  A separate instantiation, so that the single-medium kernels
// keep their registers and instruction count.
// KIN (with MULTI): kinematics, the per-cell look-ups at the perceived wavelength; its own instantiation for the same reason.
template <int GRID, int MODE, bool STORE, bool TABLES_IN_SMEM, bool MULTI, bool KIN>
__global__ void __launch_bounds__(SK_TRACE_BLOCK, (GRID == 3   ? SK_TRACE_MINBLOCKS_VORONOI
                                                   : MODE == 2 ? SK_TRACE_MINBLOCKS_PEEL
                                                   : STORE     ? SK_TRACE_MINBLOCKS_STORE
                                                               : SK_TRACE_MINBLOCKS)
                                                      - (KIN ? SK_TRACE_MINBLOCKS_KIN_LESS : MULTI ? SK_TRACE_MINBLOCKS_MULTI_LESS : 0))
    sk_wf_trace(const SkDevModel M, const SkRunArgs A, const SkBank K, const SkObsDir obs)
{
    extern __shared__ __align__(16) double smem[];
    __shared__ unsigned long long tma_bar;
    SkSmemTables T;
    if (TABLES_IN_SMEM)
    {
        // stage the per-axis border tables of the Cartesian grid in shared memory with three TMA bulk copies that complete
        // on one mbarrier (the tables are padded to 16-byte multiples, SK_TABLE_PAD); the instantiation is separate from
        // the global-memory one so that the lookups in the crossing loop compile to LDS
        const int n0 = SK_TABLE_PAD(M.nx + 1), n1 = SK_TABLE_PAD(M.ny + 1), n2 = SK_TABLE_PAD(M.nz + 1);
        if (threadIdx.x == 0) sk_mbar_init(&tma_bar, 1);
        __syncthreads();
        if (threadIdx.x == 0)
        {
            sk_mbar_arrive_expect_tx(&tma_bar, (unsigned)((n0 + n1 + n2) * sizeof(double)));
            sk_tma_load_bulk(smem, M.xv, (unsigned)(n0 * sizeof(double)), &tma_bar);
            sk_tma_load_bulk(smem + n0, M.yv, (unsigned)(n1 * sizeof(double)), &tma_bar);
            sk_tma_load_bulk(smem + n0 + n1, M.zv, (unsigned)(n2 * sizeof(double)), &tma_bar);
        }
        sk_mbar_wait(&tma_bar, 0);
        T.X = smem;
        T.Y = smem + n0;
        T.Z = smem + n0 + n1;
    }
    else
    {
        T.X = M.xv;
        T.Y = M.yv;
        T.Z = M.zv;
    }
    const SkDevModel* __restrict__ Mg = A.model;
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const bool forced = M.force_scattering != 0;
    const int n = (int)K.ctl[SK_CTL_NLIST];
    SkLocalCounters cnt;
    memset(&cnt, 0, sizeof cnt);

    int chunk_pos = 0, chunk_end = 0;  // the warp's reserved range of the list
    bool exhausted = false;            // the global list has been handed out completely
    // lane state: ACTIVE = walking a ray; PENDING = its ray has ended, results not yet written; REPLAY (MODE 0) = the lane is
    // on the second walk of its packet, to the interaction point; FOUND = that walk ended at an interaction point
    // FRONT: an empty segment in front of the grid; NOSCAT: explicit absorption in a medium that does not scatter at this
    // wavelength (the walk then accumulates the extinction for the radiation field and reports a zero path optical depth)
    enum { ACTIVE = 1, PENDING = 2, REPLAY = 4, FOUND = 8, FRONT = 16, NOSCAT = 32 };
    unsigned ls = 0;
    int slot = 0;
    SkDir kray;  // direction of the lane's ray; all peel-off rays (MODE 2) share the observer's direction, which stays in
    kray.set(0., 0., 1., nullptr);  // the kernel's parameter space: no per-lane registers, uniform sign tests
    const SkDir& k = MODE == 2 ? obs.d : kray;
    SkStepper<GRID> st;  // while a lane is not ACTIVE its stepper keeps the cell in which the walk ended
    st.cm = -1;
    double tau = 0, s = 0, limit = 0, section = 0;
    // MULTI: the second component's extinction section and its density in cell mpre live in lane registers; the density is
    // requested as soon as the walk knows its next cell, so that it travels together with the cell record.  Further
    // components (rare) are read from the tables at each crossing.
    double sec1 = 0., dn1 = 0.;
    int mpre = -1, ilam_ray = 0;
    // MULTI with explicit absorption (MediumSystem.cpp:937-955, 1112-1150): the walk is in scattering optical depth and the
    // absorption optical depth is accumulated next to it -- the ratio of the two varies from cell to cell
    const bool multi_explicit = MULTI && MODE != 2 && M.explicit_absorption;
    double seca0 = 0., seca1 = 0., taua = 0., taua_end = 0.;
    // MULTI with kinematics (sk_engine_set_velocities): the cell perceives lam_ray / (1 - k.v_m/c) and the sections are those
    // of that wavelength (MediumSystem.cpp:888-900, 958-972, 1242-1258); ilam_ray follows it from cell to cell, and the sections
    // in the lane registers are reloaded when it changes.  pk = the ray's direction in physical coordinates.
    constexpr bool kin = MULTI && KIN;
    const bool more_media = MULTI && M.nmed > 1;
    double lam_ray = 0., lamp = 0., pkx = 0., pky = 0., pkz = 0.;
    double kv = 0.;  // k . v of cell mpre, requested like the second component's density as soon as the next cell is known
    int rf_lo = 0;
#define SK_FETCH_DENSX()                                                                          \
    if (more_media || kin)                                                                        \
    {                                                                                             \
        mpre = st.m();                                                                            \
        if (more_media && mpre >= 0) dn1 = __ldg(&M.densx[(size_t)mpre]);                         \
        if (kin && mpre >= 0)                                                                     \
        {                                                                                         \
            const double4 v_ = sk_ld_rec(&M.vel[mpre]);                                           \
            kv = pkx * v_.x + pky * v_.y + pkz * v_.z;                                            \
        }                                                                                         \
    }
#define SK_LOAD_SECTIONS()                                                                        \
    {                                                                                             \
        const double* __restrict__ sg_ = multi_explicit ? M.sig_sca : M.sig_ext;                 \
        section = sg_[ilam_ray];                                                                  \
        sec1 = more_media ? sg_[M.nlam + ilam_ray] : 0.;                                          \
        if (multi_explicit)                                                                       \
        {                                                                                         \
            seca0 = M.sig_abs[ilam_ray];                                                          \
            seca1 = more_media ? M.sig_abs[M.nlam + ilam_ray] : 0.;                               \
        }                                                                                         \
    }
    int nseg = 0;
    // MODE 0 + STORE extras: luminosity of the packet, extinction factor at the start of the current segment, and the
    // column of the radiation field table for the packet's wavelength bin (null: outside the grid, .cpp:643-644)
    double lum = 0, extBeg = 1, extfac = 1;  // extfac: extinction over interaction optical depth (1 unless explicit absorption)
    double* rf = nullptr;
    double s_int = 0;  // result of a walk to the interaction point

    while (true)
    {
        // ---------------- service block: results of finished rays out, new rays in
        bool start = false;  // the lane begins a walk in this pass: a new ray, or the second walk of its packet
        if (ls & PENDING)
        {
            ls &= ~PENDING;
            if (MODE == 0 && !(ls & REPLAY))
            {
                if (ls & NOSCAT) tau = 0.;
                K.D(D_TAUPATH, slot) = tau;
                cnt.fwd_paths++;
                cnt.fwd_segs += nseg;
                if (STORE && rf && !kin) cnt.rf += nseg - ((ls & FRONT) ? 1 : 0);  // one deposit per segment inside the grid
                // no extinction along the path: the packet cannot scatter; sk_wf_advance terminates it (.cpp:702-706)
                if (tau > 0.)
                {
                    bool redraw;
                    limit = sk_interaction_depth(K.D(D_TAUINT, slot), tau, redraw);
                    if (redraw) limit = sk_redraw_interaction(M.seed, A.stream_id, K.i, K.cap, slot, tau);
                    K.D(D_TAUINT, slot) = limit;
                    ls |= REPLAY;
                    start = true;
                }
            }
            else if (MODE == 0 || MODE == 1)
            {
                // the interaction cell = the cell in which the walk stopped.  A walk that left the grid is at or beyond the
                // exit optical depth of its last segment: the interaction point is the end of the path, in the last cell
                // crossed (SpatialGridPath.cpp:199-205; the stepper holds -2 - m, and sk_wf_advance looks the lattice
                // position up again: level -1); -1: the path missed the grid.
                SkCellPos hit = st.cell(M);
                if (hit.m < -1)
                {
                    hit.m = -2 - hit.m;
                    hit.lev = -1;
                }
                // a walk that stopped inside a cell holds the optical depth at the cell's far wall in s_int: linear
                // interpolation over the segment (SpatialGridPath.cpp:188-194)
                if (multi_explicit)  // SpatialGridPath.cpp:185-195 / MediumSystem.cpp:1143-1147: interpolated like the distance
                    K.D(D_TAUABS, slot) = st.m() >= 0 ? sk_interp_linlin(limit, tau, s_int, taua, taua_end) : taua;
                if (st.m() >= 0) s_int = sk_interp_linlin(limit, tau, s_int, s, s + st.ds());
                K.D(D_SINT, slot) = s_int;
                K.I(I_MINT, slot) = hit.m;
                K.I(I_MIX, slot) = hit.ix;
                K.I(I_MIY, slot) = hit.iy;
                K.I(I_MIZ, slot) = hit.iz;
                K.I(I_MLEV, slot) = hit.lev;
                if (MODE == 0)
                    cnt.replay_segs += nseg;
                else
                {
                    if (ls & FOUND) K.I(I_STATE, slot) |= SK_ST_FOUND;
                    cnt.fwd_paths++;
                    cnt.fwd_segs += nseg;
                }
            }
            else
            {
                K.D(D_PTAU, slot) = tau;
                cnt.peel_paths++;
                cnt.peel_segs += nseg;
            }
        }
        {
            const unsigned idle = __ballot_sync(0xffffffffu, !(ls & ACTIVE) && !start);
            if (chunk_pos >= chunk_end && !exhausted)
            {
                unsigned b = 0;
                if (lane == 0) b = atomicAdd(&K.ctl[SK_CTL_CURSOR], (unsigned)SK_CHUNK);
                b = __shfl_sync(0xffffffffu, b, 0);
                if ((int)b >= n)
                    exhausted = true;
                else
                {
                    chunk_pos = (int)b;
                    chunk_end = min((int)b + SK_CHUNK, n);
                }
            }
            const int idx = chunk_pos + __popc(idle & lt_mask);
            const int take = min(__popc(idle), chunk_end - chunk_pos);
            if (!(ls & ACTIVE) && !start && idx < chunk_end)
            {
                slot = K.list[idx];
                start = true;
                ls = 0;
                if (MODE != 2)
                    kray.load(K.D(D_KX, slot), K.D(D_KY, slot), K.D(D_KZ, slot), K.D(D_IKX, slot), K.D(D_IKY, slot),
                              K.D(D_IKZ, slot), GRID == 2 ? M.lat_invh : nullptr);
                section = K.D(D_SIGEXT, slot);
                if (MULTI)
                {
                    ilam_ray = K.I(I_ILAM, slot);
                    if (kin)
                    {
                        // a peel-off ray travels at the wavelength of the peel-off packet (sk_peel_setup_values)
                        lam_ray = MODE == 2 ? K.D(M.kin_base_d + SK_KD_PLAMBDA, slot) : K.D(D_LAMBDA, slot);
                        if (MODE == 2) ilam_ray = sk_locate_clip_hint(M.lam_border, M.nlam, lam_ray, ilam_ray);
                        pkx = MODE == 2 ? obs.px : K.D(D_KX, slot);
                        pky = MODE == 2 ? obs.py : K.D(D_KY, slot);
                        pkz = MODE == 2 ? obs.pz : K.D(D_KZ, slot);
                    }
                    SK_LOAD_SECTIONS();
                }
                if (!MULTI && MODE != 2 && M.explicit_absorption)
                {
                    // the walk to the interaction point is in scattering optical depth (MediumSystem.cpp:905-934, 1075-1110);
                    // the extinction that attenuates the radiation field deposits is a fixed multiple of it
                    const double sca = M.sig_sca[K.I(I_ILAM, slot)];
                    if (sca > 0.)
                    {
                        if (STORE) extfac = section / sca;
                        section = sca;
                    }
                    else
                    {
                        if (STORE) extfac = 1.;
                        ls |= NOSCAT;
                    }
                }
                if (MODE == 0 && STORE)
                {
                    const int rf_ell = K.I(I_RFELL, slot);
                    rf = rf_ell >= 0 ? (A.primary ? M.rf1 : M.rf2c) + SK_RF_INDEX(M, 0, rf_ell) : nullptr;
                    lum = K.D(D_W, slot) / K.D(D_LAMBDA, slot);
                    if (kin)
                    {
                        // the bin follows the perceived wavelength from cell to cell (MonteCarloSimulation.cpp:667-691):
                        // rf = the whole table, lum = the weight W (the perceived luminosity is W / lambda_perceived)
                        rf = A.primary ? M.rf1 : M.rf2c;
                        lum = K.D(D_W, slot);
                    }
                }
                if (MODE == 1) limit = K.D(D_TAUINT, slot);
                if (MODE == 2) limit = K.D(D_LIMIT, slot);
            }
            if (start)
            {
                double rx = K.D(D_RX, slot), ry = K.D(D_RY, slot), rz = K.D(D_RZ, slot);
                SkCellPos p{K.I(I_M, slot), K.I(I_IX, slot), K.I(I_IY, slot), K.I(I_IZ, slot), K.I(I_LEV, slot)};
                ls |= ACTIVE;
                tau = 0.;
                s = 0.;
                nseg = 0;
                s_int = 0.;
                if (MULTI) taua = 0.;
                if (MODE == 0 && STORE) extBeg = 1.;
                if (p.m < 0)
                {
                    // the path starts outside (or exactly on the border of) the grid: PathSegmentGenerator::moveInside
                    double cumds = 0.;
                    double tx = rx, ty = ry, tz = rz;
                    const double dkx = MODE == 2 ? obs.px : K.D(D_KX, slot), dky = MODE == 2 ? obs.py : K.D(D_KY, slot),
                                 dkz = MODE == 2 ? obs.pz : K.D(D_KZ, slot);
                    p.m = -1;
                    if (sk_move_inside(tx, ty, tz, dkx, dky, dkz, Mg->ext, M.eps, cumds))
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
                            ls |= FRONT;
                        }
                    }
                    if (p.m < 0)
                    {
                        // the path misses the grid: no segments
                        ls = (ls & ~ACTIVE) | PENDING;
                        s_int = s;
                    }
                }
                st.begin(M, rx, ry, rz, p);
                SK_FETCH_DENSX();
            }
            if (take > 0) chunk_pos += take;
            if (!__any_sync(0xffffffffu, (ls & (ACTIVE | PENDING)) != 0))
            {
                if (exhausted) break;
                continue;  // the chunk ran out exactly here: reserve the next one
            }
        }
        // ---------------- inner loop: cross cells until enough lanes have finished their ray
        int want_idle = (exhausted && chunk_pos >= chunk_end) ? 32 : (MODE == 1 ? SK_REFILL_MIN_SHORT : MODE == 0 ? SK_REFILL_MIN_FUSED : SK_REFILL_MIN);
        asm volatile("" : "+r"(want_idle));  // keeps the threshold in a register instead of re-deriving it every crossing
        do
        {
            if (ls & ACTIVE)
            {
                int m;
                double dens, ds;
                st.template exit<MODE == 2>(M, Mg, T, cnt, k, m, dens, ds);
                bool done = false;
                double kappa = section * dens;  // opacity of the cell at the ray's wavelength
                if (MULTI && m >= 0)
                {
                    if (kin)
                    {
                        if (m != mpre) SK_FETCH_DENSX();  // (the Voronoi walk may re-locate its cell inside exit())
                        lamp = lam_ray / (1 - kv / SK_C_LIGHT);  // sk_perceived
                        const int il = sk_locate_clip_hint(M.lam_border, M.nlam, lamp, ilam_ray);
                        if (il != ilam_ray)
                        {
                            ilam_ray = il;
                            SK_LOAD_SECTIONS();
                        }
                        kappa = section * dens;
                    }
                    if (more_media && m != mpre) SK_FETCH_DENSX();  // (the Voronoi walk may re-locate its cell inside exit())
                    kappa = __fma_rn(sec1, dn1, kappa);
                    const double* __restrict__ sigx = multi_explicit ? M.sig_sca : M.sig_ext;
#pragma unroll 1
                    for (int h = 2; h < M.nmed; ++h)  // (third and fourth component: rare, kept out of the unrolled code)
                        kappa = __fma_rn(__ldg(&sigx[h * M.nlam + ilam_ray]),
                                         __ldg(&M.densx[(size_t)(h - 1) * (size_t)M.ncells + (size_t)m]), kappa);
                    if (multi_explicit)
                    {
                        double kappa_abs = __fma_rn(seca1, dn1, seca0 * dens);
#pragma unroll 1
                        for (int h = 2; h < M.nmed; ++h)
                            kappa_abs = __fma_rn(__ldg(&M.sig_abs[h * M.nlam + ilam_ray]),
                                                 __ldg(&M.densx[(size_t)(h - 1) * (size_t)M.ncells + (size_t)m]), kappa_abs);
                        taua_end = __fma_rn(kappa_abs, ds > 0. ? ds : 0., taua);  // absorption optical depth at the far wall
                    }
                }
                if (MODE == 0)
                {
                    if (ds > 0.)  // SpatialGridPath::addSegment, SpatialGridPath.cpp:41-48
                    {
                        nseg++;
                        const double tau1 = __fma_rn(kappa, ds, tau);
                        if (ls & REPLAY)
                        {
                            if (limit < tau1)
                            {
                                s_int = tau1;  // interaction inside this segment: interpolated by the service block
                                done = true;
                            }
                        }
                        else if (STORE && rf)
                        {
                            // the logarithms of the extinction factors at both ends of the segment are -tau and -tau1
                            const double lnBeg = multi_explicit ? -(tau + taua) : -(tau * extfac),
                                         lnEnd = multi_explicit ? -(tau1 + taua_end) : -(tau1 * extfac);
                            const double extEnd = exp(lnEnd);
                            const double extMean = sk_lnmean4(extEnd, extBeg, lnEnd, lnBeg);
                            const double Lds = lum * extMean * ds;
                            if (kin)
                            {
                                if (m >= 0)
                                {
                                    const SkDevWlg& wg = M.wlg[M.rf_grid];
                                    rf_lo = sk_wlg_upper_hint(wg, lamp, rf_lo);
                                    const int ell = wg.ell[rf_lo];
                                    if (ell >= 0)
                                    {
                                        atomicAdd(&rf[SK_RF_INDEX(M, m, ell)], (lum / lamp) * extMean * ds);
                                        cnt.rf++;
                                    }
                                }
                            }
                            else
#ifdef SK_RF_AGGREGATE
                            // (experiment, DESIGN.md section 9: lanes of the warp that deposit into the same cell and bin
                            //  are summed with shuffles and one lane issues the atomic)
                            {
                                const unsigned am = __activemask();
                                const unsigned peers = __match_any_sync(am, (unsigned long long)(rf + m));
                                const int maxcnt = __reduce_max_sync(am, __popc(peers));
                                double sum = Lds;
                                if (maxcnt > 1)
                                {
                                    sum = 0.;
                                    unsigned rem = peers;
                                    for (int kk = 0; kk < maxcnt; ++kk)
                                    {
                                        const int src = rem ? __ffs(rem) - 1 : (int)lane;
                                        const double t = __shfl_sync(am, Lds, src);
                                        if (rem) sum += t;
                                        rem &= rem - 1;
                                    }
                                }
                                if (lane == (unsigned)(__ffs(peers) - 1)) atomicAdd(&rf[m], sum);
                            }
#else
                            atomicAdd(&rf[m], Lds);  // MediumSystem.cpp:1294-1300
#endif
                            extBeg = extEnd;
                        }
                        if (!done)
                        {
                            s += ds;
                            tau = tau1;
                            if (MULTI) taua = taua_end;
                        }
                    }
                }
                else if (MODE == 1)
                {
                    if (forced ? (ds > 0.) : (ds >= 0.))  // ds < 0: the generator ended without a segment (Voronoi)
                    {
                        nseg++;
                        const double tau1 = (ls & NOSCAT) ? 0. : __fma_rn(kappa, ds, tau);
                        if (limit < tau1)
                        {
                            s_int = tau1;  // interaction inside this segment: interpolated by the service block
                            ls |= FOUND;
                            done = true;
                        }
                        else
                        {
                            s += ds;
                            tau = tau1;
                            if (MULTI) taua = taua_end;
                        }
                    }
                }
                else if (ds >= 0.)
                {
                    nseg++;
                    tau = __fma_rn(kappa, ds, tau);
                    if (tau >= limit)
                    {
                        tau = INFINITY;  // MediumSystem.cpp:1215
                        done = true;
                    }
                }
                if (!done)
                {
                    if (GRID != 3 || st.m() >= 0) st.move(M, Mg, T, cnt, k);
                    SK_FETCH_DENSX();
                    if (st.m() < 0)
                    {
                        // the path has left the grid (non-forced: no interaction)
                        done = true;
                        if (MODE == 1 || (MODE == 0 && (ls & REPLAY)))
                        {
                            s_int = s;
                            st.cm = -2 - m;  // the last cell crossed
                        }
                    }
                }
                if (done) ls ^= (ACTIVE | PENDING);  // ACTIVE -> PENDING
            }
        } while (__popc(__ballot_sync(0xffffffffu, !(ls & ACTIVE))) < want_idle);
    }
    sk_flush_counters(M, cnt);
#undef SK_LOAD_SECTIONS
#undef SK_FETCH_DENSX
}

// ---------------------------------------------------------------------------------------------------
// Event kernels: one thread per slot of the bank
// ---------------------------------------------------------------------------------------------------
// Peel-off set-up towards the observer group [j0, j1): the weight of the peel-off packet and whether it needs an
// optical depth at all.  MonteCarloSimulation::peelOffEmission (.cpp:617-634) / peelOffScattering (.cpp:784-842,
// consolidated branch), DustMix::peeloffScattering HG branch (DustMix.cpp:430-445), MediumSystem::peelOffScattering
// (MediumSystem.cpp:734-767) with the single-medium weight 1, PhotonPacket::launchScatteringPeelOff / launchEmissionPeelOff
// (PhotonPacket.cpp:66-103).  Returns true when the slot joins the peel-off ray list.
// Several medium components (cold paths, kept out of line so that the single-medium kernels keep their registers):
// MediumSystem::weightsForScattering (MediumSystem.cpp:697-730) -- the scattering opacities of the components in the
// interaction cell, normalised -- and the weighted sum of the components' phase functions towards the observer
// (MediumSystem::peelOffScattering, MediumSystem.cpp:745-754; DustMix::peeloffScattering HG branch, DustMix.cpp:430-445)
__device__ __noinline__ double sk_peel_intensity_media(const SkDevModel* __restrict__ Mg, int mint, int ilam, double costheta)
{
    const SkDevModel& M = *Mg;
    double wv[SK_MAX_MEDIA], sum = 0.;
    for (int h = 0; h < M.nmed; ++h)
    {
        wv[h] = mint >= 0 ? sk_component_density_any(M, mint, h) * M.sig_sca[h * M.nlam + ilam] : 0.;
        sum += wv[h];
    }
    double I = 0.;
    for (int h = 0; h < M.nmed; ++h)
    {
        const double w = sum > 0. ? wv[h] / sum : 0.;
        if (w > 0.)
        {
            double gp = M.gpar[h * M.nlam + ilam];
            double value = fabs(gp) > 0.95 ? sk_mean_hg(gp, costheta) : sk_value_hg(gp, costheta);
            I += value * w;
        }
    }
    return I;
}
// MediumSystem::albedoForScattering (MediumSystem.cpp:678-693): sum k_sca / sum k_ext over the components
__device__ __noinline__ double sk_albedo_media(const SkDevModel* __restrict__ Mg, int m, int ilam)
{
    const SkDevModel& M = *Mg;
    const double ksca = sk_opacity_sum(M, M.sig_sca, ilam, m), kext = sk_opacity_sum(M, M.sig_ext, ilam, m);
    return kext > 0. ? ksca / kext : 0.;
}
// MediumSystem::simulateScattering (MediumSystem.cpp:805-818): the scattering component, from NR::cdf over the scattering
// opacities in the interaction cell and NR::locateClip of the deviate u
__device__ __noinline__ int sk_scattering_component(const SkDevModel* __restrict__ Mg, int mint, int ilam, double u)
{
    const SkDevModel& M = *Mg;
    double Xv[SK_MAX_MEDIA + 1];
    Xv[0] = 0.;
    for (int h = 0; h < M.nmed; ++h) Xv[h + 1] = Xv[h] + sk_component_density_any(M, mint, h) * M.sig_sca[h * M.nlam + ilam];
    const double norm = Xv[M.nmed];
    int hsel = 0;
    for (int h = 1; h < M.nmed; ++h)
        if (u >= Xv[h] / norm) hsel = h;
    return hsel;
}
template <bool MULTI>
__device__ __forceinline__ bool sk_peel_setup_values(const SkDevModel& M, const SkDevModel* __restrict__ Mg, const SkBank& K, int slot, bool scattering,
                                                     int j0, int j1, double W, double lambda, double x, double y, double z,
                                                     double kx, double ky, double kz, int ilam, int mint)
{
    const SkDevInstr& q0 = M.instr[j0];
    const double ox = q0.kobs[0], oy = q0.kobs[1], oz = q0.kobs[2];
    double peelW;
    if (MULTI && M.kin)
    {
        // kinematics: `lambda` is the wavelength the interaction cell perceives (scattering) and the peel-off packet leaves
        // Doppler-shifted for the direction towards the observer by the bulk velocity of the cell
        // (PhotonPacket::launchScatteringPeelOff, PhotonPacket.cpp:89-103), or by the velocity of the emitter from the rest
        // wavelength at emission (launchEmissionPeelOff, PhotonPacket.cpp:66-85)
        const int kb = M.kin_base_d;
        if (scattering)
        {
            const double4 vm = M.vel[mint >= 0 ? mint : 0];
            lambda = sk_shifted_emission(lambda, ox, oy, oz, mint >= 0 ? vm.x : 0., mint >= 0 ? vm.y : 0., mint >= 0 ? vm.z : 0.);
        }
        else
            lambda = sk_shifted_emission(K.D(kb + SK_KD_LAMBDA0, slot), ox, oy, oz, K.D(kb + SK_KD_VSX, slot),
                                         K.D(kb + SK_KD_VSY, slot), K.D(kb + SK_KD_VSZ, slot));
        K.D(kb + SK_KD_PLAMBDA, slot) = lambda;
    }
    if (scattering)
    {
        double costheta = kx * ox + ky * oy + kz * oz;
        double I = 0.;
        if (!MULTI)
        {
            double gp = M.gpar[ilam];
            double value = fabs(gp) > 0.95 ? sk_mean_hg(gp, costheta) : sk_value_hg(gp, costheta);
            I += value * 1.;
        }
        else
            I = sk_peel_intensity_media(Mg, mint, ilam, costheta);
        peelW = W * I;
    }
    else
        peelW = W;  // isotropic emission
    K.D(D_PEELW, slot) = peelW;
    bool need = false;
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
            K.D(D_PTAU, slot) = INFINITY;  // MediumSystem.cpp:1196
            need = false;
        }
        else
            K.D(D_LIMIT, slot) = log(L) + 745;  // MediumSystem.cpp:1199
    }
    return need;
}
__device__ __forceinline__ bool sk_peel_setup(const SkDevModel& M, const SkDevModel* __restrict__ Mg, const SkBank& K, int slot, int st,
                                              int j0, int j1)
{
    if (M.kin && (st & SK_ST_SCATTER))  // the wavelength the interaction cell perceives and its index (sk_wf_advance)
        return sk_peel_setup_values<true>(M, Mg, K, slot, true, j0, j1, K.D(D_W, slot), K.D(M.kin_base_d + SK_KD_LAMP, slot),
                                          K.D(D_RX, slot), K.D(D_RY, slot), K.D(D_RZ, slot), K.D(D_KX, slot), K.D(D_KY, slot),
                                          K.D(D_KZ, slot), K.I(M.kin_base_i + SK_KI_ILAMP, slot), K.I(I_MINT, slot));
    if (M.nmed > 1 || M.kin)
        return sk_peel_setup_values<true>(M, Mg, K, slot, (st & SK_ST_SCATTER) != 0, j0, j1, K.D(D_W, slot), K.D(D_LAMBDA, slot),
                                          K.D(D_RX, slot), K.D(D_RY, slot), K.D(D_RZ, slot), K.D(D_KX, slot), K.D(D_KY, slot),
                                          K.D(D_KZ, slot), K.I(I_ILAM, slot), (st & SK_ST_SCATTER) ? K.I(I_MINT, slot) : -1);
    return sk_peel_setup_values<false>(M, Mg, K, slot, (st & SK_ST_SCATTER) != 0, j0, j1, K.D(D_W, slot), K.D(D_LAMBDA, slot),
                                K.D(D_RX, slot), K.D(D_RY, slot), K.D(D_RZ, slot), K.D(D_KX, slot), K.D(D_KY, slot),
                                K.D(D_KZ, slot), K.I(I_ILAM, slot), (st & SK_ST_SCATTER) ? K.I(I_MINT, slot) : -1);
}

// advance: the interaction that ends the previous round and, for the surviving packets, the peel-off set-up towards
// the first observer group; free slots are collected for the launch kernel.
// (five resident blocks of 256 threads = 48 registers for the single-medium instantiation: what the compiler chose on its own
//  for the kernels profiled in DESIGN.md section 9, and measured level with the alternatives there; pinned so that unrelated
//  changes of the model structure do not flip the choice)
#ifndef SK_ADVANCE_MINBLOCKS
#define SK_ADVANCE_MINBLOCKS (MULTI ? 4 : 5)
#endif
#define SK_ADVANCE_BOUNDS __launch_bounds__(SK_EVENT_BLOCK, SK_ADVANCE_MINBLOCKS)
template <int GRID, bool MULTI>
__global__ void SK_ADVANCE_BOUNDS sk_wf_advance(const SkDevModel M, const SkRunArgs A, const SkBank K,
                                                                 const int j0, const int j1)
{
    const SkSmemTables T{M.xv, M.yv, M.zv};
    const SkDevModel* __restrict__ Mg = A.model;
    const bool forced = M.force_scattering != 0;
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = slot < K.n;
    int st = valid ? K.I(I_STATE, slot) : 0;

    // Every field the interaction needs is requested before the first use, for live and dead slots alike: the kernel is a
    // streaming pass over the bank, and its speed is the number of loads in flight, not the instruction count.
    const int sl = valid ? slot : 0;
    const int m = K.I(I_MINT, sl), ilam = K.I(I_ILAM, sl), nscatt = K.I(I_NSCATT, sl);
    const int mix = K.I(I_MIX, sl), miy = K.I(I_MIY, sl), miz = K.I(I_MIZ, sl), mlev = K.I(I_MLEV, sl);
    const double sigext = K.D(D_SIGEXT, sl), W0 = K.D(D_W, sl), taupath = K.D(D_TAUPATH, sl), sint = K.D(D_SINT, sl);
    const double tauint = K.D(D_TAUINT, sl);
    const double kx = K.D(D_KX, sl), ky = K.D(D_KY, sl), kz = K.D(D_KZ, sl);
    const double rx0 = K.D(D_RX, sl), ry0 = K.D(D_RY, sl), rz0 = K.D(D_RZ, sl);
    const double lambda = K.D(D_LAMBDA, sl), lthr = K.D(D_LTHR, sl);
    bool survivor = false;
    double W = W0, x = rx0, y = ry0, z = rz0;
    double lamp = lambda;  // kinematics: wavelength perceived by the interaction cell, and its index in the dust tables
    int ilamp = ilam;

    // ---- the interaction: albedo weight, move, termination test (.cpp:724-741, 576-580)
    if (st & SK_ST_LIVE)
    {
        bool alive = true;
        if (!forced && !(st & SK_ST_FOUND)) alive = false;  // escaped, MonteCarloSimulation.cpp:594
        // forced scattering without extinction along the path: the packet cannot scatter, terminate it (.cpp:702-706); the
        // forward trace has found taupath = 0 (no segments, or only empty cells) and drawn no interaction point
        if (forced && !(taupath > 0.)) alive = false;
        if (alive)
        {
            // MediumSystem::albedoForScattering, MediumSystem.cpp:678-693
            double albedo = 0.;
            if (MULTI && M.kin)
            {
                // the wavelength the interaction cell perceives (MediumSystem::perceivedWavelengthForScattering,
                // MediumSystem.cpp:667-674): albedo, peel-off weights and the scattering itself use it
                const double4 vm = M.vel[m >= 0 ? m : 0];
                lamp = m >= 0 ? sk_perceived(lambda, kx, ky, kz, vm.x, vm.y, vm.z) : lambda;
                ilamp = sk_locate_clip_hint(M.lam_border, M.nlam, lamp, ilam);
            }
            if (m >= 0)
            {
                double dn = sk_cell_density<GRID>(M, m);
                double ksca = dn * M.sig_sca[ilam];
                double kext = dn * sigext;
                albedo = kext > 0. ? ksca / kext : 0.;
                if (MULTI) albedo = sk_albedo_media(Mg, m, ilamp);
            }
            if (forced)
            {
                // the path-length bias weight p/q of the interaction optical depth the trace kernel has used (.cpp:714-718),
                // then the escape fraction and the albedo (.cpp:729-733)
                const double em = expm1(-taupath);
                const double xi = M.path_length_bias;
                if (xi != 0.)
                {
                    const double pw = -exp(-tauint) / em;
                    const double qw = (1.0 - xi) * pw + xi / taupath;
                    W *= pw / qw;
                }
                // explicit absorption: the absorption optical depth up to the interaction point, a fixed multiple of the
                // scattering optical depth for one medium with constant sections (SpatialGridPath.cpp:185-195)
                if (MULTI)  // (several components: the absorption optical depth accumulated by the walk)
                    W *= -em * (M.explicit_absorption ? exp(-K.D(D_TAUABS, sl)) : albedo);
                else
                    W *= -em * (M.explicit_absorption ? exp(-(tauint * M.sig_abs[ilam] / M.sig_sca[ilam])) : albedo);
            }
            else if (MULTI)
                W *= M.explicit_absorption ? exp(-K.D(D_TAUABS, sl)) : albedo;
            else
                W *= M.explicit_absorption ? exp(-(tauint * M.sig_abs[ilam] / M.sig_sca[ilam])) : albedo;  // .cpp:757-773
            x = rx0 + sint * kx;  // PhotonPacket::propagate, PhotonPacket.cpp:107-111
            y = ry0 + sint * ky;
            z = rz0 + sint * kz;
            K.D(D_W, slot) = W;
            K.D(D_RX, slot) = x;
            K.D(D_RY, slot) = y;
            K.D(D_RZ, slot) = z;
            double L = W / lambda;
            if (forced)
            {
                if (L <= 0 || (L <= lthr && nscatt >= M.min_scatt_events)) alive = false;
            }
            else if (L <= 0)
                alive = false;
            if (alive)
            {
                // the next paths start in the interaction cell unless rounding moved the point out of it
                SkCellPos c{m, mix, miy, miz, mlev};
                bool inside = m >= 0 && sk_box_strictly_inside(M.ext, x, y, z);
                if (GRID == 3)
                {
                    if (inside) c.m = sk_voronoi_walk(Mg, x, y, z, m);  // a new path looks its cell up (.cpp:1076)
                }
                else if (inside && (mlev < 0 || !sk_cell_contains<GRID>(M, T, c, x, y, z)))
                {
                    SkCellPos c2;
                    sk_locate<GRID>(Mg, T, x, y, z, c2);
                    c = c2;
                }
                if (!inside) c.m = -1;
                K.I(I_M, slot) = c.m;
                K.I(I_IX, slot) = c.ix;
                K.I(I_IY, slot) = c.iy;
                K.I(I_IZ, slot) = c.iz;
                K.I(I_LEV, slot) = c.lev;
                st = SK_ST_LIVE | SK_ST_SCATTER;
                K.I(I_STATE, slot) = st;
                survivor = true;
                if (MULTI && M.kin)
                {
                    K.D(M.kin_base_d + SK_KD_LAMP, slot) = lamp;
                    K.I(M.kin_base_i + SK_KI_ILAMP, slot) = ilamp;
                }
            }
        }
        if (!alive)
        {
            sk_finish_history(M, K, slot, MULTI && M.kin);
            st = 0;
        }
    }

    // ---- free slots go to the launch kernel through a compact list (the launch code then runs with full warps); the
    //      census of the survivors rides along
    const bool live = (st & SK_ST_LIVE) != 0;
    sk_block_append(K.free_list, &K.ctl[SK_CTL_NFREE], valid && !live, slot, &K.ctl[SK_CTL_NLIVE], live);
    if (A.peel && j1 > j0)
    {
        bool need = survivor && sk_peel_setup_values<MULTI>(M, Mg, K, slot, true, j0, j1, W, lamp, x, y, z, kx, ky, kz, ilamp, m);
        sk_block_append(K.list, &K.ctl[SK_CTL_NLIST], need, slot);
    }
}

// launch: a new history into every free slot collected by `advance` (SourceSystem::launch, SourceSystem.cpp:101-113 /
// SecondarySourceSystem::launch, SecondarySourceSystem.cpp:130-142), then its emission peel-off set-up
// (MonteCarloSimulation::peelOffEmission, .cpp:617-634).  One thread per entry of the free list.
#ifndef SK_LAUNCH_MINBLOCKS
#define SK_LAUNCH_MINBLOCKS 4
#endif
template <int GRID>
__global__ void __launch_bounds__(SK_EVENT_BLOCK, SK_LAUNCH_MINBLOCKS) sk_wf_launch(const SkDevModel M, const SkRunArgs A, const SkBank K,
                                                                const int j0, const int j1)
{
    const SkSmemTables T{M.xv, M.yv, M.zv};
    const SkDevModel* __restrict__ Mg = A.model;
    const unsigned nfree = K.ctl[SK_CTL_NFREE];
    if (blockIdx.x * blockDim.x >= nfree) return;  // the whole block is beyond the list
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    const bool want = idx < nfree;
    const int slot = want ? K.free_list[idx] : 0;
    bool live = false;
    {
        const unsigned long long h = sk_block_reserve(A.work_counter, want);
        if (want && h < A.count)
        {
            const unsigned long long history = A.first + sk_history_of(A, h);
            SkRng g;
            sk_rng_init(g, M.seed, A.stream_id, history, 0);
            SkLaunch pp;
            if (A.primary)
                sk_launch_primary(Mg, g, history, pp);
            else
                sk_launch_secondary<GRID>(Mg, T, g, history, pp);
            const double lambda0 = pp.lambda;
            if (M.kin)
            {
                // PhotonPacket::launch with a velocity interface (PhotonPacket.cpp:33): the wavelength is Doppler-shifted for
                // the launch direction, the weight keeps the rest wavelength
                pp.lambda = sk_shifted_emission(lambda0, pp.kx, pp.ky, pp.kz, pp.vx, pp.vy, pp.vz);
                pp.ilam = sk_locate_clip_hint(M.lam_border, M.nlam, pp.lambda, pp.ilam);
            }
            if (pp.W / pp.lambda > 0)  // MonteCarloSimulation.cpp:553
            {
                SkCellPos c;
                c.m = -1;
                c.ix = c.iy = c.iz = c.lev = 0;
                if (sk_box_strictly_inside(M.ext, pp.rx, pp.ry, pp.rz)) sk_locate<GRID>(Mg, T, pp.rx, pp.ry, pp.rz, c);
                K.D(D_RX, slot) = pp.rx;
                K.D(D_RY, slot) = pp.ry;
                K.D(D_RZ, slot) = pp.rz;
                K.D(D_KX, slot) = pp.kx;
                K.D(D_KY, slot) = pp.ky;
                K.D(D_KZ, slot) = pp.kz;
                K.D(D_IKX, slot) = sk_dir_recip(pp.kx, GRID == 2 ? M.lat_h[0] : 1.);
                K.D(D_IKY, slot) = sk_dir_recip(pp.ky, GRID == 2 ? M.lat_h[1] : 1.);
                K.D(D_IKZ, slot) = sk_dir_recip(pp.kz, GRID == 2 ? M.lat_h[2] : 1.);
                if (M.force_scattering) K.D(D_TAUINT, slot) = sk_predraw_interaction(g, M.path_length_bias);
                K.D(D_LAMBDA, slot) = pp.lambda;
                K.D(D_W, slot) = pp.W;
                K.D(D_LTHR, slot) = (pp.W / pp.lambda) / M.min_weight_reduction;  // .cpp:563
                K.D(D_SIGEXT, slot) = M.sig_ext[pp.ilam];
                K.I(I_HLO, slot) = (int)(uint32_t)history;
                K.I(I_HHI, slot) = (int)(uint32_t)(history >> 32);
                K.I(I_DRAW, slot) = (int)g.draw;
                K.I(I_NSCATT, slot) = 0;
                K.I(I_ILAM, slot) = pp.ilam;
                // the packet's bin of the radiation field grid (MonteCarloSimulation.cpp:643): fixed for its whole life
                K.I(I_RFELL, slot) = M.rf_grid >= 0 ? sk_wlg_bin(M.wlg[M.rf_grid], pp.lambda) : -1;
                K.I(I_M, slot) = c.m;
                K.I(I_IX, slot) = c.ix;
                K.I(I_IY, slot) = c.iy;
                K.I(I_IZ, slot) = c.iz;
                K.I(I_LEV, slot) = c.lev;
                for (int j = 0; j < M.ninstr; ++j)
                {
                    K.D(D_HISTW0 + j, slot) = 0.;
                    K.I(I_HELL0 + j, slot) = -1;
                    const int ps = M.instr[j].pix_slot;
                    if (ps >= 0)
                    {
                        K.I(M.pix_base_i + ps * SK_PIX_INTS + SK_PIX_K, slot) = 0;
                        K.I(M.pix_base_i + ps * SK_PIX_INTS + SK_PIX_K + 1, slot) = -1;
                    }
                    const int ss = M.kin ? M.instr[j].sed_slot : -1;
                    if (ss >= 0)
                    {
                        K.I(M.pix_base_i + ss * SK_PIX_INTS + SK_PIX_K, slot) = 0;
                        K.I(M.pix_base_i + ss * SK_PIX_INTS + SK_PIX_K + 1, slot) = -1;
                    }
                }
                if (M.kin)
                {
                    const int kb = M.kin_base_d;
                    K.D(kb + SK_KD_LAMBDA0, slot) = lambda0;
                    K.D(kb + SK_KD_VSX, slot) = pp.vx;
                    K.D(kb + SK_KD_VSY, slot) = pp.vy;
                    K.D(kb + SK_KD_VSZ, slot) = pp.vz;
                }
                K.I(I_STATE, slot) = SK_ST_LIVE;
                live = true;
            }
        }
    }
    {
        const bool need = A.peel && j1 > j0 && live && sk_peel_setup(M, Mg, K, slot, SK_ST_LIVE, j0, j1);
        // (the launched packets are counted -- sk_counters_t::packets -- in the same pass)
        sk_block_append(K.list, &K.ctl[SK_CTL_NLIST], need, slot, &K.ctl[SK_CTL_NLIVE], live, &M.counters[0]);
    }
}

// peel-off set-up towards a further observer group
__global__ void __launch_bounds__(SK_EVENT_BLOCK) sk_wf_peel_setup(const SkDevModel M, const SkDevModel* __restrict__ Mg, const SkBank K,
                                                                    const int j0, const int j1)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const int st = slot < K.n ? K.I(I_STATE, slot) : 0;
    bool need = (st & SK_ST_LIVE) && sk_peel_setup(M, Mg, K, slot, st, j0, j1);
    sk_block_append(K.list, &K.ctl[SK_CTL_NLIST], need, slot);
}

// detect: FluxRecorder::detect (FluxRecorder.cpp:304-468) for the observer group [j0, j1); when `last`, also the
// pending scattering events -- MediumSystem::simulateScattering (MediumSystem.cpp:796-823) + DustMix::performScattering
// HG branch (DustMix.cpp:496-511) -- and the list of forward rays.  Persistent grid-stride kernel: each block keeps the
// SED arrays of the group in shared memory (nl_stride > 0) and adds them to the global arrays once at its end.
// (KIN: the instantiation for runs with kinematics, so that the other keeps its registers)
template <bool KIN>
__global__ void __launch_bounds__(SK_EVENT_BLOCK) sk_wf_detect(const SkDevModel M, const SkRunArgs A, const SkBank K,
                                                                const int j0, const int j1, const int last,
                                                                const int nl_stride)
{
    extern __shared__ double sed_sm[];
    const int per_instr = SK_NUM_COMP * nl_stride;
    for (int i = threadIdx.x; i < (j1 - j0) * per_instr; i += blockDim.x) sed_sm[i] = 0.;
    __syncthreads();
    SkLocalCounters cnt;
    memset(&cnt, 0, sizeof cnt);
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < K.n; slot += gridDim.x * blockDim.x)
    {
        int st = K.I(I_STATE, slot);
        const bool live = (st & SK_ST_LIVE) != 0;
        if (j1 > j0)
        {
            // (every lane of the warp walks the instruments, dead slots with `live` false: the frame tallies are combined
            //  over the warp, sk_record_ifu_warp)
            double lambda = 1., x = 0., y = 0., z = 0., L = 0.;
            int nscatt = 0;
            if (live)
            {
                lambda = KIN ? K.D(M.kin_base_d + SK_KD_PLAMBDA, slot) : K.D(D_LAMBDA, slot);  // of the peel-off packet
                x = K.D(D_RX, slot);
                y = K.D(D_RY, slot);
                z = K.D(D_RZ, slot);
                L = K.D(D_PEELW, slot) / lambda;
                nscatt = (st & SK_ST_SCATTER) ? K.I(I_NSCATT, slot) + 1 : 0;
            }
            for (int j = j0; j < j1; ++j)
            {
                const SkDevInstr& q = M.instr[j];
                int l = -1, ell = -1;
                const bool hit = live && sk_detect_geometry(M, q, x, y, z, lambda, l, ell);
                double Lext = 0.;
                if (hit)
                {
                    Lext = L * exp(-K.D(D_PTAU, slot));
                    cnt.det++;
                    if (q.include_sed)
                        sk_record_sed(q, ell, L, Lext, nscatt, A.primary != 0, nl_stride ? sed_sm + (j - j0) * per_instr : nullptr,
                                      nl_stride);
                    if (KIN && q.sed_slot >= 0)
                        sk_add_pixel_contribution(M, K, q, slot, ell, Lext, true);
                    else if (q.record_stats && q.include_sed)
                    {
                        K.D(D_HISTW0 + j, slot) += Lext;  // FluxRecorder.cpp:457-466
                        K.I(I_HELL0 + j, slot) = ell;
                    }
                    if (q.pix_slot >= 0 && l >= 0) sk_add_pixel_contribution(M, K, q, slot, l + ell * (int)q.npix, Lext);
                }
                if (q.include_ifu) sk_record_ifu_warp(q, hit && l >= 0, l, ell, L, Lext, nscatt, A.primary != 0);
            }
        }
        if (last)
        {
            if (live && (st & SK_ST_SCATTER))
            {
                SkRng g;
                sk_rng_load(g, M, A, K, slot);
                int hsel = 0;
                // (kinematics: the index of the wavelength the interaction cell perceives, MediumSystem.cpp:802)
                const int ilam_s = KIN ? K.I(M.kin_base_i + SK_KI_ILAMP, slot) : K.I(I_ILAM, slot);
                if (M.nmed > 1) hsel = sk_scattering_component(A.model, K.I(I_MINT, slot), ilam_s, sk_uniform(g));
                double gp = M.gpar[hsel * M.nlam + ilam_s];
                double kx = K.D(D_KX, slot), ky = K.D(D_KY, slot), kz = K.D(D_KZ, slot);
                if (fabs(gp) < 1e-6)
                    sk_random_direction(g, kx, ky, kz);
                else
                {
                    double f = ((1.0 - gp) * (1.0 + gp)) / (1.0 - gp + 2.0 * gp * sk_uniform(g));
                    double costheta = (1.0 + gp * gp - f * f) / (2.0 * gp);
                    sk_random_direction_about(g, kx, ky, kz, costheta);
                }
                K.D(D_KX, slot) = kx;
                K.D(D_KY, slot) = ky;
                K.D(D_KZ, slot) = kz;
                const bool lattice = M.grid_kind == 2;  // the octree is walked in lattice coordinates
                K.D(D_IKX, slot) = sk_dir_recip(kx, lattice ? M.lat_h[0] : 1.);
                K.D(D_IKY, slot) = sk_dir_recip(ky, lattice ? M.lat_h[1] : 1.);
                K.D(D_IKZ, slot) = sk_dir_recip(kz, lattice ? M.lat_h[2] : 1.);
                if (KIN)
                {
                    // PhotonPacket::scatter(bfk, bfv, lambda), PhotonPacket.cpp:115-122: the packet leaves at the perceived
                    // wavelength, shifted by the bulk velocity of the cell for its new direction
                    const int mint = K.I(I_MINT, slot);
                    const double4 vm = M.vel[mint >= 0 ? mint : 0];
                    const double lnew = sk_shifted_emission(K.D(M.kin_base_d + SK_KD_LAMP, slot), kx, ky, kz, mint >= 0 ? vm.x : 0.,
                                                            mint >= 0 ? vm.y : 0., mint >= 0 ? vm.z : 0.);
                    const int inew = sk_locate_clip_hint(M.lam_border, M.nlam, lnew, ilam_s);
                    K.D(D_LAMBDA, slot) = lnew;
                    K.I(I_ILAM, slot) = inew;
                    K.D(D_SIGEXT, slot) = M.sig_ext[inew];
                    if (M.rf_grid >= 0) K.I(I_RFELL, slot) = sk_wlg_bin(M.wlg[M.rf_grid], lnew);
                }
                // the deviates of the interaction that follows this scattering (simulateForcedPropagation is the next to draw)
                if (M.force_scattering) K.D(D_TAUINT, slot) = sk_predraw_interaction(g, M.path_length_bias);
                K.I(I_DRAW, slot) = (int)g.draw;
                K.I(I_NSCATT, slot) += 1;
                K.I(I_STATE, slot) = st & ~SK_ST_SCATTER;
                cnt.scatt++;
            }
            if (M.force_scattering) sk_block_append(K.list, &K.ctl[SK_CTL_NLIST], live, slot);
        }
    }
    sk_flush_counters(M, cnt);
    if (nl_stride)
    {
        __syncthreads();
        for (int i = threadIdx.x; i < (j1 - j0) * per_instr; i += blockDim.x)
        {
            const double v = sed_sm[i];
            if (v != 0.)
            {
                const SkDevInstr& q = M.instr[j0 + i / per_instr];
                const int c = (i % per_instr) / nl_stride, ell = i % nl_stride;
                atomicAdd(&q.sed[c][ell], v);
            }
        }
    }
}

// sample: the interaction optical depth of non-forced propagation -- Random::expon for simulateNonForcedPropagation
// (.cpp:749) -- and the list of rays to walk to the interaction point.  (With forced scattering the forward trace kernel
// draws the interaction optical depth itself, sk_sample_interaction.)
__global__ void __launch_bounds__(SK_EVENT_BLOCK) sk_wf_sample(const SkDevModel M, const SkRunArgs A, const SkBank K)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = slot < K.n;
    const int st = valid ? K.I(I_STATE, slot) : 0;
    const bool live = (st & SK_ST_LIVE) != 0;
    if (live)
    {
        SkRng g;
        sk_rng_load(g, M, A, K, slot);
        K.D(D_TAUINT, slot) = -log(sk_uniform(g));
        K.I(I_DRAW, slot) = (int)g.draw;
        K.I(I_STATE, slot) = st & ~SK_ST_FOUND;
    }
    sk_block_append(K.list, &K.ctl[SK_CTL_NLIST], live, slot);
}


// ---------------------------------------------------------------------------------------------------
// Forward rays in order of their radiation-field wavelength bin (segments that store the radiation field): a counting
// sort of the ray list by I_RFELL (bin -1 = outside the grid first), list -> free_list (unused at this point of the round).
// count: histogram of the bins; scan: exclusive prefix sum into the cursors; scatter: every block reserves a range per bin
// for its tile of the list with one atomic per bin, its threads take ranks inside from shared-memory atomics.
// ---------------------------------------------------------------------------------------------------
#define SK_SORT_MAX_BINS 1024
#define SK_SORT_TILE 2048
__global__ void __launch_bounds__(SK_EVENT_BLOCK) sk_wf_bin_count(const SkBank K, const int nbins, unsigned int* counts)
{
    __shared__ unsigned int h[SK_SORT_MAX_BINS + 1];
    for (int i = threadIdx.x; i <= nbins; i += blockDim.x) h[i] = 0;
    __syncthreads();
    const unsigned n = K.ctl[SK_CTL_NLIST];
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicAdd(&h[K.I(I_RFELL, K.list[i]) + 1], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i <= nbins; i += blockDim.x)
        if (h[i]) atomicAdd(&counts[i], h[i]);
}
__global__ void sk_wf_bin_scan(unsigned int* counts, const int nbins)
{
    if (threadIdx.x || blockIdx.x) return;
    unsigned int run = 0;
    for (int i = 0; i <= nbins; ++i)
    {
        const unsigned int c = counts[i];
        counts[i] = run;  // becomes the cursor of bin i
        run += c;
    }
}
__global__ void __launch_bounds__(SK_EVENT_BLOCK) sk_wf_bin_scatter(const SkBank K, const int nbins, unsigned int* cursor)
{
    __shared__ unsigned int h[SK_SORT_MAX_BINS + 1];
    __shared__ unsigned int base[SK_SORT_MAX_BINS + 1];
    const unsigned n = K.ctl[SK_CTL_NLIST];
    const unsigned t0 = blockIdx.x * SK_SORT_TILE;
    if (t0 >= n) return;
    const unsigned t1 = min(t0 + SK_SORT_TILE, n);
    for (int i = threadIdx.x; i <= nbins; i += blockDim.x) h[i] = 0;
    __syncthreads();
    int slot[SK_SORT_TILE / SK_EVENT_BLOCK], bin[SK_SORT_TILE / SK_EVENT_BLOCK];
    unsigned rank[SK_SORT_TILE / SK_EVENT_BLOCK];
#pragma unroll
    for (int j = 0; j < SK_SORT_TILE / SK_EVENT_BLOCK; ++j)
    {
        const unsigned i = t0 + j * SK_EVENT_BLOCK + threadIdx.x;
        bin[j] = -1;
        if (i < t1)
        {
            slot[j] = K.list[i];
            bin[j] = K.I(I_RFELL, slot[j]) + 1;
            rank[j] = atomicAdd(&h[bin[j]], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= nbins; i += blockDim.x) base[i] = h[i] ? atomicAdd(&cursor[i], h[i]) : 0u;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SK_SORT_TILE / SK_EVENT_BLOCK; ++j)
        if (bin[j] >= 0) K.free_list[base[bin[j]] + rank[j]] = slot[j];
}

// ---------------------------------------------------------------------------------------------------
// Draining: once every history of the segment has been handed out the bank empties geometrically (a packet survives a
// round with probability ~3/4) over several dozen rounds.  So that the element-wise kernels and the trace grids shrink with
// it, the live packets are packed into the head of the bank whenever fewer than half of the slots in use are alive:
// partition lists the live slots of the tail [n_new, n) and the dead slots of the head [0, n_new), move copies the former
// into the latter (all fields), after which the bank is [0, n_new).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SK_EVENT_BLOCK) sk_wf_partition(const SkBank K, const int n_new)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = slot < K.n;
    const bool live = valid && (K.I(I_STATE, slot) & SK_ST_LIVE);
    sk_block_append(K.list, &K.ctl[SK_CTL_NLIST], live && slot >= n_new, slot);       // sources
    sk_block_append(K.free_list, &K.ctl[SK_CTL_NFREE], valid && !live && slot < n_new, slot);  // destinations
}
__global__ void __launch_bounds__(SK_EVENT_BLOCK) sk_wf_move(const SkBank K, const int nd, const int ni)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K.ctl[SK_CTL_NLIST]) return;
    const int src = K.list[i], dst = K.free_list[i];
    for (int f = 0; f < nd; ++f) K.D(f, dst) = K.D(f, src);
    for (int f = 0; f < ni; ++f) K.I(f, dst) = K.I(f, src);
    K.I(I_STATE, src) = 0;
}
