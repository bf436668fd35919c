// sk_blocks.cuh -- building blocks of the photon life cycle on the device: the pieces of
// MonteCarloSimulation::performLifeCycle (SKIRT/core/MonteCarloSimulation.cpp:538-613) and of everything it calls
// that the stage kernels of sk_wavefront.cuh are assembled from: cell location and cell crossing for each grid type,
// the source samplers, the instrument projection and the detector tallies.
#pragma once
#include "sk_device.cuh"

// Per-packet state of the in-flight bank: field-major arrays of `cap` entries (structure of arrays in HBM),
// the device counterpart of PhotonPacket (SKIRT/utils/PhotonPacket.hpp:337-363).
enum {
    D_RX, D_RY, D_RZ, D_KX, D_KY, D_KZ, D_IKX, D_IKY, D_IKZ, D_LAMBDA, D_W, D_LTHR, D_SIGEXT, D_TAUPATH, D_TAUINT, D_SINT,
    D_PEELW, D_PTAU, D_LIMIT, D_HISTW0, SK_ND = D_HISTW0 + SK_MAX_INSTR
};
enum {
    I_HLO, I_HHI, I_DRAW, I_NSCATT, I_STATE, I_ILAM, I_M, I_IX, I_IY, I_IZ, I_LEV, I_MINT, I_MIX, I_MIY, I_MIZ,
    I_MLEV, I_RFELL, I_HELL0, SK_NI = I_HELL0 + SK_MAX_INSTR
};
// the absorption optical depth at the interaction point (explicit absorption with several medium components) travels from the
// forward trace to the advance kernel of the next round in the field of the peel-off limit, which is idle in between
#define D_TAUABS D_LIMIT
// extra fields of a bank with kinematics (SkDevModel::kin), relative to kin_base_d / kin_base_i: the wavelength of the current
// peel-off packet, the wavelength the interaction cell perceives and its index in the dust tables, the rest wavelength at
// emission and the velocity of the emitter (PhotonPacket::_lambda0, _bvi: needed for the emission peel-offs)
enum { SK_KD_PLAMBDA, SK_KD_LAMP, SK_KD_LAMBDA0, SK_KD_VSX, SK_KD_VSY, SK_KD_VSZ, SK_KD_COUNT };
enum { SK_KI_ILAMP, SK_KI_COUNT };
// I_STATE bits
#define SK_ST_LIVE 1
#define SK_ST_SCATTER 2    // a scattering event is pending (peel-off of kind "scattering", then new direction)
#define SK_ST_FOUND 4      // non-forced propagation: interaction point found

struct SkBank {
    double* d;        // [D_HISTW0 + ninstr][cap]
    int32_t* i;       // [I_HELL0 + ninstr][cap]
    int32_t* list;    // [cap] slots of the rays of the trace stage being prepared / consumed
    int32_t* free_list;  // [cap] free slots found by the advance kernel, filled by the launch kernel
    unsigned int* ctl;  // control words: see SK_CTL_*
    int32_t cap;      // allocated slots = stride of the field-major arrays
    int32_t n;        // slots in use [0, n): the whole bank while histories are handed out, shrinking while it drains
    // pool of list chunks for histories that reach more than SK_PIX_K distinct frame pixels (per-pixel statistics): chunk c
    // holds entries pool_lell/pool_w[c*SK_PIX_C ..), pool_next[c] chains to the next older chunk of the same history;
    // pool_free is a stack of free chunks, pool_ctl = {stack height, contributions recorded early because the pool ran dry}
    int32_t* pool_lell;
    double* pool_w;
    int32_t* pool_next;
    int32_t* pool_free;
    int* pool_ctl;
    __device__ __forceinline__ double& D(int f, int s) const { return d[(size_t)f * cap + s]; }
    __device__ __forceinline__ int32_t& I(int f, int s) const { return i[(size_t)f * cap + s]; }
};
enum { SK_CTL_NLIST, SK_CTL_CURSOR, SK_CTL_NLIVE, SK_CTL_NFREE, SK_CTL_WORDS = 4 };
#define SK_BANK_FIELDS_D(ninstr) (D_HISTW0 + ((ninstr) > 0 ? (ninstr) : 1))
#define SK_BANK_FIELDS_I(ninstr) (I_HELL0 + ((ninstr) > 0 ? (ninstr) : 1))

struct SkLocalCounters {
    unsigned int packets, fwd_paths, fwd_segs, replay_segs, peel_paths, peel_segs, scatt, rf, det, fallbacks;
};

// border tables are allocated and staged in multiples of 16 bytes (TMA bulk copies): entries rounded up to even
#define SK_TABLE_PAD(n) (((n) + 1) & ~1)
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
// integer lattice: a node is (ix,iy,iz,level) in units of the finest level, so the child that holds the lattice point
// (fx,fy,fz) is given by one bit of each coordinate per level.  `fc` is the node id of the first child of the node
// (at level `lev`) to descend from, or -(cell+1) when that node is a leaf already.
__device__ __forceinline__ void sk_tree_descend(const int32_t* __restrict__ node_child, int maxlevel, int fc, int fx, int fy,
                                                int fz, int lev, SkCellPos& out)
{
    while (fc >= 0)
    {
        const int b = maxlevel - lev - 1;
        const int l = ((fx >> b) & 1) | (((fy >> b) & 1) << 1) | (((fz >> b) & 1) << 2);
        lev++;
        fc = __ldg(&node_child[fc + l]);
    }
    const int mask = ~((1 << (maxlevel - lev)) - 1);
    out.m = -(fc + 1);
    out.ix = fx & mask;
    out.iy = fy & mask;
    out.iz = fz & mask;
    out.lev = lev;
}
// Lattice coordinate of a position along one axis of the octree domain, u = (x - min) / pitch, as a multiplication by the
// reciprocal pitch.  A result within 1e-9 of an integer is that integer: a position that lies exactly on a cell border (a
// point source at the centre of the grid) then stays exactly on it whatever the rounding of the reciprocal, and so in the
// cell the reference's comparisons with the recursive midpoints put it in; for any other position the shift (1e-9 of the
// finest cell) is far below eps.  Every kernel converts with this one function, so they all agree on the cell.
__device__ __forceinline__ double sk_lat_coord(double x, double xmin, double invh)
{
    const double u = (x - xmin) * invh;
    const double r = rint(u);
    return fabs(u - r) < 1e-9 ? r : u;
}
// the finest lattice cell that holds lattice coordinate u (the domain box is closed: u == N belongs to the last cell)
__device__ __forceinline__ int sk_lat_index(double u, int N)
{
    const int i = __double2int_rd(u);
    return i < 0 ? 0 : (i >= N ? N - 1 : i);
}

// one 32-byte Voronoi cell record {site; density} through the read-only path (one 256-bit load, one sector)
__device__ __forceinline__ double4 sk_ld_rec(const double4* p)
{
    unsigned long long q0, q1, q2, q3;
    asm("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(q0), "=l"(q1), "=l"(q2), "=l"(q3) : "l"(p));
    return make_double4(__longlong_as_double(q0), __longlong_as_double(q1), __longlong_as_double(q2), __longlong_as_double(q3));
}

// VoronoiMeshSnapshot::cellIndex (VoronoiMeshSnapshot.cpp:1006-1040): the cell whose site is nearest to the position, found
// by walking the neighbour graph from `hint` (or from the block table when hint < 0) to ever closer sites; see
// oracle/sk_oracle.c voronoi_walk for why the walk ends in the cell that contains the position.
__device__ __noinline__ int sk_voronoi_walk(const SkDevModel* __restrict__ Mg, double x, double y, double z, int hint)
{
    const SkDevModel& M = *Mg;
    int m = hint;
    if (m < 0)
    {
        const int nb = M.vnb;
        int i = (int)((x - M.ext[0]) / (M.ext[3] - M.ext[0]) * nb);
        int j = (int)((y - M.ext[1]) / (M.ext[4] - M.ext[1]) * nb);
        int k = (int)((z - M.ext[2]) / (M.ext[5] - M.ext[2]) * nb);
        i = i < 0 ? 0 : i >= nb ? nb - 1 : i;
        j = j < 0 ? 0 : j >= nb ? nb - 1 : j;
        k = k < 0 ? 0 : k >= nb ? nb - 1 : k;
        m = M.vblock[((size_t)i * nb + j) * nb + k];
    }
    double4 s = sk_ld_rec(&M.vrec[m]);
    double dx = x - s.x, dy = y - s.y, dz = z - s.z;
    double d = dx * dx + dy * dy + dz * dz;
    while (true)
    {
        int best = -1;
        double dbest = d;
        const long long i1 = M.vnbr_off[m + 1];
        for (long long i = M.vnbr_off[m]; i < i1; ++i)
        {
            const int mi = __ldg(&M.vnbr[i]);
            if (mi < 0) continue;
            const double4 t = sk_ld_rec(&M.vrec[mi]);
            double ex = x - t.x, ey = y - t.y, ez = z - t.z;
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
    if (GRID == 3)
    {
        out.ix = out.iy = out.iz = out.lev = 0;
        out.m = sk_voronoi_walk(Mg, x, y, z, -1);
    }
    else if (GRID == 1)
    {
        out.ix = sk_locate_clip(T.X, Mg->nx + 1, x);  // CartesianSpatialGrid.cpp:105-107
        out.iy = sk_locate_clip(T.Y, Mg->ny + 1, y);
        out.iz = sk_locate_clip(T.Z, Mg->nz + 1, z);
        out.lev = 0;
        out.m = out.iz + Mg->nz * out.iy + Mg->nz * Mg->ny * out.ix;
    }
    else
    {
        const int N = Mg->nx;
        const int fx = sk_lat_index(sk_lat_coord(x, Mg->ext[0], Mg->lat_invh[0]), N);
        const int fy = sk_lat_index(sk_lat_coord(y, Mg->ext[1], Mg->lat_invh[1]), N);
        const int fz = sk_lat_index(sk_lat_coord(z, Mg->ext[2], Mg->lat_invh[2]), N);
        sk_tree_descend(Mg->node_child, Mg->maxlevel, __ldg(&Mg->node_child[0]), fx, fy, fz, 0, out);
    }
}

// ---------------------------------------------------------------------------------------------------
// Ray direction in traversal coordinates.  The ray carries the reciprocals of its direction components (0 where the
// reference treats the component as zero, fabs(k) <= 1e-15), so an exit distance costs a multiplication instead of a
// division per axis, and the sign / zero / grazing tests of the crossing loop as bits of one integer.
//   Cartesian, Voronoi: k = the unit direction.
//   Octree: k = direction in lattice units per unit of path length, k_a / pitch_a: the ray is walked in the lattice
//   coordinates u = (r - min) / pitch, in which the borders of cell (ix, iy, iz, level) are the integers ix and
//   ix + 2^(maxlevel - level); path lengths (ds, eps) stay physical.
// ---------------------------------------------------------------------------------------------------
#define SK_DIR_NEG 1u     // bits 0..2: component < 0 (the ray leaves through the lower wall of that axis)
#define SK_DIR_ZERO 8u    // bits 3..5: fabs(component) <= 1e-15: no exit through the walls of that axis
#define SK_DIR_GRAZE 64u  // bits 6..8: !(fabs(component) > 1e-3): the eps advance across such a wall may be lost to rounding
struct SkDir {
    double kx, ky, kz;
    double ikx, iky, ikz;
    double bigx, bigy, bigz;  // observer directions only (set()): DBL_MAX for an axis without exit walls, else 0
    unsigned flags;
    // (x, y, z) = the unit direction; h = pitch of the finest octree lattice per axis, or null for physical coordinates
    __host__ __device__ __forceinline__ void set(double x, double y, double z, const double* h)
    {
        unsigned f = 0;
        if (x < 0.0) f |= SK_DIR_NEG;
        if (y < 0.0) f |= SK_DIR_NEG << 1;
        if (z < 0.0) f |= SK_DIR_NEG << 2;
        if (!(fabs(x) > 1e-15)) f |= SK_DIR_ZERO;
        if (!(fabs(y) > 1e-15)) f |= SK_DIR_ZERO << 1;
        if (!(fabs(z) > 1e-15)) f |= SK_DIR_ZERO << 2;
        if (!(fabs(x) > 1e-3)) f |= SK_DIR_GRAZE;
        if (!(fabs(y) > 1e-3)) f |= SK_DIR_GRAZE << 1;
        if (!(fabs(z) > 1e-3)) f |= SK_DIR_GRAZE << 2;
        flags = f;
        kx = h ? x / h[0] : x;
        ky = h ? y / h[1] : y;
        kz = h ? z / h[2] : z;
        ikx = (f & SK_DIR_ZERO) ? 0. : (h ? h[0] : 1.0) / x;
        iky = (f & (SK_DIR_ZERO << 1)) ? 0. : (h ? h[1] : 1.0) / y;
        ikz = (f & (SK_DIR_ZERO << 2)) ? 0. : (h ? h[2] : 1.0) / z;
        bigx = (f & SK_DIR_ZERO) ? DBL_MAX : 0.;
        bigy = (f & (SK_DIR_ZERO << 1)) ? DBL_MAX : 0.;
        bigz = (f & (SK_DIR_ZERO << 2)) ? DBL_MAX : 0.;
    }
    // the same from the unit direction and the reciprocals an event kernel has stored in the bank (sk_dir_recip):
    // invh = reciprocal pitch of the octree lattice per axis, or null for physical coordinates
    __device__ __forceinline__ void load(double x, double y, double z, double ix, double iy, double iz, const double* invh)
    {
        unsigned f = 0;
        if (x < 0.0) f |= SK_DIR_NEG;
        if (y < 0.0) f |= SK_DIR_NEG << 1;
        if (z < 0.0) f |= SK_DIR_NEG << 2;
        if (fabs(ix) > 1e290) f |= SK_DIR_ZERO;
        if (fabs(iy) > 1e290) f |= SK_DIR_ZERO << 1;
        if (fabs(iz) > 1e290) f |= SK_DIR_ZERO << 2;
        if (!(fabs(x) > 1e-3)) f |= SK_DIR_GRAZE;
        if (!(fabs(y) > 1e-3)) f |= SK_DIR_GRAZE << 1;
        if (!(fabs(z) > 1e-3)) f |= SK_DIR_GRAZE << 2;
        flags = f;
        kx = invh ? x * invh[0] : x;
        ky = invh ? y * invh[1] : y;
        kz = invh ? z * invh[2] : z;
        ikx = ix;
        iky = iy;
        ikz = iz;
        bigx = bigy = bigz = 0.;
    }
};
// Reciprocal of a direction component of a packet's own (random) direction in traversal coordinates (h = lattice pitch of
// the axis for the octree, 1 otherwise), computed once per direction by the event kernel that creates it.  Where the
// reference treats the component as zero (fabs(k) <= 1e-15: no exit through the walls of that axis) the reciprocal is
// +-1e300 with the sign of the component: the exit distance (wall - position) * 1e300 of that axis then loses every
// comparison, with no test in the crossing loop.  (The observer's direction, where exact zeros are the rule for a face-on
// instrument, uses exact offsets instead: SkDir::set.)
#define SK_RECIP_ZERO 1e300
__device__ __forceinline__ double sk_dir_recip(double k, double h)
{
    return (fabs(k) > 1e-15) ? h / k : (k < 0.0 ? -SK_RECIP_ZERO : SK_RECIP_ZERO);
}
// the observer's direction of a peel-off trace: traversal form + the unit vector itself (for entering the grid)
struct SkObsDir {
    SkDir d;
    double px, py, pz;
};

// one 32-byte record with a single 256-bit load through the read-only path (LDG.E.256 on sm_100a): half the L1 tag
// look-ups of two 128-bit loads, and the L1 data pipe is what bounds the crossing loop
__device__ __forceinline__ void sk_ld256(const void* p, int4& a, int4& b)
{
    unsigned long long q0, q1, q2, q3;
    asm("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(q0), "=l"(q1), "=l"(q2), "=l"(q3) : "l"(p));
    a = make_int4((int)(unsigned)q0, (int)(q0 >> 32), (int)(unsigned)q1, (int)(q1 >> 32));
    b = make_int4((int)(unsigned)q2, (int)(q2 >> 32), (int)(unsigned)q3, (int)(q3 >> 32));
}

// ---------------------------------------------------------------------------------------------------
// Steppers: one lane's walk through the grid, cut in two halves so that a walk can stop INSIDE a cell (the interaction
// point) without having stepped out of it:
//     begin(r, cell)   start in `cell` at position r (the caller has located the cell)
//     exit(...)        the segment through the current cell: its index m, density and length ds; does not move.
//                      ds < 0 (Voronoi only): the generator ended without a further segment, m() < 0 afterwards
//     move(...)        step to the next cell; m() < 0 afterwards means the path has left the grid
//     cell()           the current cell
//   Cartesian: CartesianSpatialGrid::MySegmentGenerator::next, CartesianSpatialGrid.cpp:95-162
//   Octree:    TreeSpatialGrid::MySegmentGenerator::next, TreeSpatialGrid.cpp:140-216, with TreeNode::neighbor
//              (TreeNode.cpp:103-112) served by the per-cell links
//   Voronoi:   VoronoiMeshSnapshot::MySegmentGenerator::next, VoronoiMeshSnapshot.cpp:1087-1179
// ---------------------------------------------------------------------------------------------------
template <int GRID>
struct SkStepper;

// ---- Cartesian grid
template <>
struct SkStepper<1> {
    double rx, ry, rz;
    int cm, ix, iy, iz;
    double ds_;
    int axis;
    __device__ __forceinline__ int m() const { return cm; }
    __device__ __forceinline__ double ds() const { return ds_; }
    __device__ __forceinline__ SkCellPos cell(const SkDevModel&) const { return SkCellPos{cm, ix, iy, iz, 0}; }
    __device__ __forceinline__ void begin(const SkDevModel&, double x, double y, double z, const SkCellPos& p)
    {
        rx = x;
        ry = y;
        rz = z;
        cm = p.m;
        ix = p.ix;
        iy = p.iy;
        iz = p.iz;
    }
    template <bool OBSERVER>
    __device__ __forceinline__ void exit(const SkDevModel& M, const SkDevModel* __restrict__, const SkSmemTables& T,
                                         SkLocalCounters&, const SkDir& k, int& m_out, double& dens_out, double& ds_out)
    {
        const double dens = __ldg(&M.dens[cm]);
        const double xE = T.X[ix + ((k.kx < 0.0) ? 0 : 1)];
        const double yE = T.Y[iy + ((k.ky < 0.0) ? 0 : 1)];
        const double zE = T.Z[iz + ((k.kz < 0.0) ? 0 : 1)];
        const double dsx = (k.flags & SK_DIR_ZERO) ? DBL_MAX : (xE - rx) * k.ikx;
        const double dsy = (k.flags & (SK_DIR_ZERO << 1)) ? DBL_MAX : (yE - ry) * k.iky;
        const double dsz = (k.flags & (SK_DIR_ZERO << 2)) ? DBL_MAX : (zE - rz) * k.ikz;
        if (dsx <= dsy && dsx <= dsz)
        {
            axis = 0;
            ds_ = dsx;
        }
        else if (dsy < dsx && dsy <= dsz)
        {
            axis = 1;
            ds_ = dsy;
        }
        else
        {
            axis = 2;
            ds_ = dsz;
        }
        m_out = cm;
        dens_out = dens;
        ds_out = ds_;
    }
    __device__ __forceinline__ void move(const SkDevModel& M, const SkDevModel* __restrict__, const SkSmemTables& T,
                                         SkLocalCounters&, const SkDir& k)
    {
        bool outside;
        if (axis == 0)
        {
            rx = T.X[ix + ((k.kx < 0.0) ? 0 : 1)];
            ry += k.ky * ds_;
            rz += k.kz * ds_;
            ix += (k.kx < 0.0) ? -1 : 1;
            outside = (ix >= M.nx || ix < 0);
        }
        else if (axis == 1)
        {
            ry = T.Y[iy + ((k.ky < 0.0) ? 0 : 1)];
            rx += k.kx * ds_;
            rz += k.kz * ds_;
            iy += (k.ky < 0.0) ? -1 : 1;
            outside = (iy >= M.ny || iy < 0);
        }
        else
        {
            rz = T.Z[iz + ((k.kz < 0.0) ? 0 : 1)];
            rx += k.kx * ds_;
            ry += k.ky * ds_;
            iz += (k.kz < 0.0) ? -1 : 1;
            outside = (iz >= M.nz || iz < 0);
        }
        cm = outside ? -1 : iz + M.nz * iy + M.nz * M.ny * ix;
    }
};

// ---- Octree, walked in lattice coordinates.
// Cold path: the full neighbour search of TreeSpatialGrid.cpp:190-207 for the situations the fast path does not decide
// itself (near-ties between exit walls, grazing directions, no neighbour across the exit wall although the position is
// still in the domain): the leaf that contains the new position (TreeNode::neighbor + root()->leafChild fall-back), the
// nextafter escape when stuck in the same cell, and termination.
__device__ __noinline__ void sk_tree_step_rare(const SkDevModel* __restrict__ Mg, double& ux, double& uy, double& uz,
                                               unsigned dirflags, int m_old, SkCellPos& q)
{
    const int N = Mg->nx;
    const double top = (double)N;
    for (int attempt = 0; attempt < 2; ++attempt)
    {
        if (!(ux >= 0. && ux <= top && uy >= 0. && uy <= top && uz >= 0. && uz <= top))
        {
            q.m = -1;
            return;
        }
        sk_tree_descend(Mg->node_child, Mg->maxlevel, __ldg(&Mg->node_child[0]), sk_lat_index(ux, N), sk_lat_index(uy, N),
                        sk_lat_index(uz, N), 0, q);
        if (q.m != m_old) return;
        if (attempt == 0)
        {
            // PathSegmentGenerator::propagateToNextAfter, PathSegmentGenerator.hpp:148-153
            ux = nextafter(ux, (dirflags & SK_DIR_NEG) ? -DBL_MAX : DBL_MAX);
            uy = nextafter(uy, (dirflags & (SK_DIR_NEG << 1)) ? -DBL_MAX : DBL_MAX);
            uz = nextafter(uz, (dirflags & (SK_DIR_NEG << 2)) ? -DBL_MAX : DBL_MAX);
        }
    }
    q.m = -1;  // still in the same cell: the path ends (TreeSpatialGrid.cpp:205-207)
}

template <>
struct SkStepper<2> {
    double ux, uy, uz;    // position in lattice coordinates
    int cm, ix, iy, iz;   // current cell: index and the lattice coordinates of its lower corner
    int sh;               // log2 of its size in lattice units = maxlevel - level
    double ds_;
    int link;             // the link across the exit wall
    int how;              // bits 0..1 exit axis, bit 2: the next cell needs the full search
    __device__ __forceinline__ int m() const { return cm; }
    __device__ __forceinline__ double ds() const { return ds_; }
    __device__ __forceinline__ void begin(const SkDevModel& M, double x, double y, double z, const SkCellPos& p)
    {
        ux = sk_lat_coord(x, M.ext[0], M.lat_invh[0]);
        uy = sk_lat_coord(y, M.ext[1], M.lat_invh[1]);
        uz = sk_lat_coord(z, M.ext[2], M.lat_invh[2]);
        cm = p.m;
        ix = p.ix;
        iy = p.iy;
        iz = p.iz;
        sh = M.maxlevel - p.lev;
    }
    __device__ __forceinline__ SkCellPos cell(const SkDevModel& M) const { return SkCellPos{cm, ix, iy, iz, M.maxlevel - sh}; }
    template <bool OBSERVER>
    __device__ __forceinline__ void exit(const SkDevModel& M, const SkDevModel* __restrict__, const SkSmemTables&,
                                         SkLocalCounters&, const SkDir& k, int& m_out, double& dens_out, double& ds_out)
    {
        // one 32-byte sector: density + the six neighbour links of the current cell
        int4 a, b;
        sk_ld256(&M.cells[cm], a, b);
        const unsigned f = k.flags;
        const int size = 1 << sh;
        const bool nx = f & SK_DIR_NEG, ny = f & (SK_DIR_NEG << 1), nz = f & (SK_DIR_NEG << 2);
        // the exit walls are lattice planes: integers, exactly representable
        const double xnext = (double)(ix + (nx ? 0 : size));
        const double ynext = (double)(iy + (ny ? 0 : size));
        const double znext = (double)(iz + (nz ? 0 : size));
        // no exit through the walls of an axis the direction has no component along: the observer's direction adds
        // DBL_MAX to (wall - position) * 0, a packet's own direction multiplies by +-1e300 (sk_dir_recip)
        const double dsx = OBSERVER ? __fma_rn(xnext - ux, k.ikx, k.bigx) : (xnext - ux) * k.ikx;
        const double dsy = OBSERVER ? __fma_rn(ynext - uy, k.iky, k.bigy) : (ynext - uy) * k.iky;
        const double dsz = OBSERVER ? __fma_rn(znext - uz, k.ikz, k.bigz) : (znext - uz) * k.ikz;
        // exit wall: x if dsx<=dsy && dsx<=dsz, else y if dsy<=dsx && dsy<=dsz, else z (TreeSpatialGrid.cpp:160-178); once x
        // is ruled out, y wins exactly when dsy<=dsz
        const bool cyz = dsy <= dsz;
        const double mn = cyz ? dsy : dsz;
        const bool takex = dsx <= mn;
        const double ds = takex ? dsx : mn;
        const int lx = nx ? a.z : a.w, ly = ny ? b.x : b.y, lz = nz ? b.z : b.w;
        const int lyz = cyz ? ly : lz;
        link = takex ? lx : lyz;
        // The link decides the next cell unless the new position may also have crossed a second wall (the exit distance
        // of another wall is within a few eps of the nearest) or the direction grazes the exit wall (the eps advance may
        // be lost to rounding); those cases take the reference's full search.
        const double near = ds + M.eps4;
        const bool bx = dsx <= near, by = dsy <= near, bz = dsz <= near;
        const bool tie = (bx && by) || (bx && bz) || (by && bz);
        const unsigned gyz = cyz ? (SK_DIR_GRAZE << 1) : (SK_DIR_GRAZE << 2);
        const unsigned gbit = takex ? SK_DIR_GRAZE : gyz;
        const bool rare = tie || (f & gbit);
        how = (takex ? 0 : (cyz ? 1 : 2)) | (rare ? 4 : 0);
        ds_ = ds;
        m_out = cm;
        dens_out = __hiloint2double(a.y, a.x);
        ds_out = ds;
    }
    __device__ __forceinline__ void move(const SkDevModel& M, const SkDevModel* __restrict__ Mg, const SkSmemTables&,
                                         SkLocalCounters& cnt, const SkDir& k)
    {
        const double adv = ds_ + M.eps;
        ux = __fma_rn(k.kx, adv, ux);
        uy = __fma_rn(k.ky, adv, uy);
        uz = __fma_rn(k.kz, adv, uz);
        if (!(how & 4))
        {
            // The common way for a path to end: no neighbour across the exit wall, which then is a wall of the domain, and
            // the position has been advanced eps beyond it (the direction does not graze the wall, the exit is not a
            // near-tie): TreeNode::neighbor() and root()->leafChild() both come back empty (TreeSpatialGrid.cpp:190-205).
            if (link < 0)
            {
                cm = -1;
                return;
            }
            // step to the lattice point just across the exit wall, then align to the neighbour's level
            const unsigned f = k.flags;
            const int size = 1 << sh;
            const int axis = how & 3;
            ix += axis == 0 ? ((f & SK_DIR_NEG) ? -1 : size) : 0;
            iy += axis == 1 ? ((f & (SK_DIR_NEG << 1)) ? -1 : size) : 0;
            iz += axis == 2 ? ((f & (SK_DIR_NEG << 2)) ? -1 : size) : 0;
            const int nlev = (link >> SK_LINK_LEVEL_SHIFT) & 15;
            int nsh = M.maxlevel - nlev;
            int fc = link & SK_LINK_INDEX_MASK;
            if (link & SK_LINK_INTERNAL)
            {
                // the neighbour is subdivided: descend to the leaf that holds the new position; along the exit axis the
                // lattice coordinate is the one just across the wall (exact), across it the position decides (it lies
                // inside the neighbour node: no clamping)
                const int fx = axis == 0 ? ix : __double2int_rd(ux);
                const int fy = axis == 1 ? iy : __double2int_rd(uy);
                const int fz = axis == 2 ? iz : __double2int_rd(uz);
                do
                {
                    nsh--;
                    const int l = ((fx >> nsh) & 1) | (((fy >> nsh) & 1) << 1) | (((fz >> nsh) & 1) << 2);
                    fc = __ldg(&M.node_child[fc + l]);
                } while (fc >= 0);
                fc = -(fc + 1);
                ix = fx;
                iy = fy;
                iz = fz;
            }
            const int mask = -1 << nsh;  // a leaf at the same, a coarser or (after the descent) a finer level
            cm = fc;
            ix &= mask;
            iy &= mask;
            iz &= mask;
            sh = nsh;
            return;
        }
        const double top = (double)M.nx;
        if (link < 0 && !(ux >= 0. && ux <= top && uy >= 0. && uy <= top && uz >= 0. && uz <= top))
        {
            cm = -1;
            return;
        }
        // (copies confine the address-taken variables, which live in local memory, to this cold branch)
        cnt.fallbacks++;
        double tx = ux, ty = uy, tz = uz;
        SkCellPos q;
        sk_tree_step_rare(Mg, tx, ty, tz, k.flags, cm, q);
        ux = tx;
        uy = ty;
        uz = tz;
        cm = q.m;
        ix = q.ix;
        iy = q.iy;
        iz = q.iz;
        sh = M.maxlevel - q.lev;
    }
};

#ifndef SK_VOR_UNROLL
#define SK_VOR_UNROLL 3  // neighbour records requested together (measured on 2e5 cells, 1e7 packets: 1: 407 ms, 2: 324 ms, 3: 313 ms,
                         // 4: 366 ms per segment; the index-list layout of round 1 with 2: 352 ms)
#endif
// ---- Voronoi mesh: the exit point is the nearest intersection with the bisecting planes towards the neighbouring sites
// and with the domain walls
template <>
struct SkStepper<3> {
    double rx, ry, rz;
    int cm, mq;
    double sq;
    __device__ __forceinline__ int m() const { return cm; }
    __device__ __forceinline__ double ds() const { return sq; }
    __device__ __forceinline__ void begin(const SkDevModel&, double x, double y, double z, const SkCellPos& p)
    {
        rx = x;
        ry = y;
        rz = z;
        cm = p.m;
    }
    __device__ __forceinline__ SkCellPos cell(const SkDevModel&) const { return SkCellPos{cm, 0, 0, 0, 0}; }
    // The exit distance towards neighbour i is the quotient s_i = num_i / den_i with den_i = n.k > 0 (bisecting plane,
    // n = p_i - p_r; VoronoiMeshSnapshot.cpp:1108-1129) or |k_axis| (domain wall, .cpp:1134-1143).  The reference divides
    // for every neighbour and keeps the smallest positive quotient; here the candidates are compared by cross-multiplication
    // (s_i < s_q <=> num_i den_q < num_q den_i, all den > 0) and only the winner is divided -- the same operands, so the
    // same ds; the ~15 divisions per crossing were two thirds of the loop's instructions.  The candidates of a cell are one
    // contiguous run of 32-byte records {site of the neighbour; its index} (M.vnrec) instead of an index list plus as many
    // scattered site records, and SK_VOR_UNROLL of them are requested before the first is used, so that their L2 round
    // trips overlap: the loop is bound by the chain of dependent loads per crossing (offsets -> records -> next cell), not by
    // its arithmetic.  (Rejected, measured: plane records {n; |n|^2/2} relative to the cell's site -- 12 instead of 27 fp64
    // instructions per neighbour, rounding-level differences only, parity tests green -- were SLOWER, 394 vs 313 ms, because
    // the winner's index then is one more dependent load per crossing.)
    template <bool OBSERVER>
    __device__ __forceinline__ void exit(const SkDevModel& M, const SkDevModel* __restrict__ Mg, const SkSmemTables&,
                                         SkLocalCounters& cnt, const SkDir& k, int& m_out, double& dens_out, double& ds_out)
    {
        while (true)
        {
            const double4 pr = sk_ld_rec(&M.vrec[cm]);
            const int NO_INDEX = -99;
            double numq = DBL_MAX, denq = 1.0;
            mq = NO_INDEX;
            const int i1 = (int)__ldg(&M.vnbr_off[cm + 1]);
            for (int i = (int)__ldg(&M.vnbr_off[cm]); i < i1; i += SK_VOR_UNROLL)
            {
                int mi[SK_VOR_UNROLL];
                double4 pj[SK_VOR_UNROLL];
#pragma unroll
                for (int u = 0; u < SK_VOR_UNROLL; ++u) pj[u] = sk_ld_rec(&M.vnrec[min(i + u, i1 - 1)]);  // (a repeated last entry never wins again)
#pragma unroll
                for (int u = 0; u < SK_VOR_UNROLL; ++u) mi[u] = (int)__double_as_longlong(pj[u].w);
#pragma unroll
                for (int u = 0; u < SK_VOR_UNROLL; ++u)
                {
                    double num, den;
                    if (mi[u] >= 0)
                    {
                        const double4 pi = pj[u];
                        const double nx = pi.x - pr.x, ny = pi.y - pr.y, nz = pi.z - pr.z;
                        den = nx * k.kx + ny * k.ky + nz * k.kz;
                        const double px = 0.5 * (pi.x + pr.x), py = 0.5 * (pi.y + pr.y), pz = 0.5 * (pi.z + pr.z);
                        num = nx * (px - rx) + ny * (py - ry) + nz * (pz - rz);
                    }
                    else
                    {
                        const int axis = (-mi[u] - 1) >> 1;
                        const double wall = M.ext[axis + (((-mi[u] - 1) & 1) ? 3 : 0)];
                        num = wall - (axis == 0 ? rx : axis == 1 ? ry : rz);
                        den = axis == 0 ? k.kx : axis == 1 ? k.ky : k.kz;
                        if (den < 0.)
                        {
                            num = -num;
                            den = -den;
                        }
                    }
                    if (den > 0. && num > 0. && num * denq < numq * den)
                    {
                        numq = num;
                        denq = den;
                        mq = mi[u];
                    }
                }
            }
            if (mq != NO_INDEX)
            {
                sq = numq / denq;
                m_out = cm;
                dens_out = pr.w;
                ds_out = sq;
                return;
            }
            // no exit point (rare): nudge the position, look the cell up again (.cpp:1156-1167)
            cnt.fallbacks++;
            rx += k.kx * M.eps;
            ry += k.ky * M.eps;
            rz += k.kz * M.eps;
            cm = sk_box_contains(M.ext, rx, ry, rz) ? sk_voronoi_walk(Mg, rx, ry, rz, cm) : -1;
            if (cm < 0)
            {
                // outside the domain: the path ends without a further segment
                m_out = -1;
                dens_out = 0.;
                ds_out = -1.;
                return;
            }
        }
    }
    __device__ __forceinline__ void move(const SkDevModel& M, const SkDevModel* __restrict__, const SkSmemTables&,
                                         SkLocalCounters&, const SkDir& k)
    {
        const double adv = sq + M.eps;
        rx += k.kx * adv;
        ry += k.ky * adv;
        rz += k.kz * adv;
        cm = mq < 0 ? -1 : mq;
    }
};
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
            double m = p[5], tanp = p[6], Rz = p[7], phiz = p[8], w = p[9], N = p[10], cn = p[11];
            double c = 1.0 + (cn - 1.0) * w;
            double phi, t;
            const double gamma = log(R / Rz) / tanp + phiz + 0.5 * M_PI / m;  // the same value in every trial
            do
            {
                phi = 2.0 * M_PI * sk_uniform(g);
                double perturbation = (1.0 - w) + w * cn * sk_pow_even(sin(0.5 * m * (gamma - phi)), 2 * N);
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
    double vx, vy, vz;  // velocity of the emitter (kinematics): the source at the launch position, or the emitting dust cell
};

// the bulk velocity of a source at the launch position: PointSource velocityX/Y/Z, or GeometricSource::velocityMagnitude()
// times the vector field (GeometricSource.cpp:73-79; RadialVectorField.cpp:19-37, CylindricalVectorField.cpp:19-38)
__device__ __forceinline__ void sk_source_velocity(const SkDevSource& s, double x, double y, double z, double& vx, double& vy,
                                                   double& vz)
{
    vx = vy = vz = 0.;
    if (s.velocity_kind == SK_VEL_CONSTANT)
    {
        vx = s.velocity[0];
        vy = s.velocity[1];
        vz = s.velocity[2];
    }
    else if (s.velocity_kind == SK_VEL_RADIAL || s.velocity_kind == SK_VEL_CYLINDRICAL)
    {
        const double mag = s.velocity[0], unity = s.velocity[1], expon = s.velocity[2];
        double ux, uy, uz;
        if (s.velocity_kind == SK_VEL_RADIAL)
        {
            ux = x;
            uy = y;
            uz = z;
        }
        else
        {
            ux = -y;
            uy = x;
            uz = 0.;
        }
        const double rr = sqrt(ux * ux + uy * uy + uz * uz);
        if (rr == 0.) return;
        ux /= rr;
        uy /= rr;
        uz /= rr;
        double f = 1.;
        if (unity > 0.)
            if ((expon > 0. && rr < unity) || (expon < 0. && rr > unity)) f = pow(rr / unity, expon);
        vx = mag * (f * ux);
        vy = mag * (f * uy);
        vz = mag * (f * uz);
    }
}

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
    sk_source_velocity(s, pp.rx, pp.ry, pp.rz, pp.vx, pp.vy, pp.vz);
}

// DustSecondarySource::launch (DustSecondarySource.cpp:511-581): implemented in sk_secondary.cuh
template <int GRID>
__device__ __noinline__ void sk_launch_secondary(const SkDevModel* __restrict__ Mg, const SkSmemTables& T, SkRng& g,
                                    unsigned long long history, SkLaunch& pp);

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
    ell = sk_wlg_bin(M.wlg[q.wlg], lambda * q.zp1);  // the packet's redshifted wavelength, FluxRecorder.cpp:309-313
    return ell >= 0;
}

// Second half of FluxRecorder::detect (FluxRecorder.cpp:320-433): component routing and the tallies.
// Component routing by origin and number of scatterings (FluxRecorder.cpp:345-380): the array that takes the extincted
// luminosity, and -- or -1 -- the transparent array and the array of the individual scattering level.
__device__ __forceinline__ void sk_route(const SkDevInstr& q, int nscatt, bool primary_origin, int& c_ext, int& c_tr, int& c_lev)
{
    c_tr = -1;
    c_lev = -1;
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
}
// The SED arrays.  `sed_sm` (optional) is the block's shared-memory copy [SK_NUM_COMP][nl_stride] of this instrument's SED
// arrays: all packets of a round hit the same few hundred SED bins, so they are combined per block before they reach the
// L2 atomics.
__device__ __forceinline__ void sk_record_sed(const SkDevInstr& q, int ell, double L, double Lext, int nscatt,
                                              bool primary_origin, double* sed_sm, int nl_stride)
{
    int c_ext, c_tr, c_lev;
    sk_route(q, nscatt, primary_origin, c_ext, c_tr, c_lev);
    if (sed_sm)
    {
        atomicAdd(&sed_sm[c_ext * nl_stride + ell], Lext);
        if (c_tr >= 0) atomicAdd(&sed_sm[c_tr * nl_stride + ell], L);
        if (c_lev >= 0) atomicAdd(&sed_sm[c_lev * nl_stride + ell], Lext);
    }
    else
    {
        atomicAdd(&q.sed[c_ext][ell], Lext);  // LockFree::add, LockFree.hpp:23-37 -> native fp64 RED
        if (c_tr >= 0) atomicAdd(&q.sed[c_tr][ell], L);
        if (c_lev >= 0) atomicAdd(&q.sed[c_lev][ell], Lext);
    }
}
// The frame arrays, index l + ell * Npix (FluxRecorder.cpp:433).  Called by ALL lanes of a warp (`hit` false for lanes
// without a detection): lanes whose detections land on the same pixel, bin and component set -- the image of a point
// source, a bright knot -- are summed over the warp with shuffles first, and one lane issues the atomics for them
// (same-address atomics serialise in the L2).
__device__ __forceinline__ void sk_record_ifu_warp(const SkDevInstr& q, bool hit, int l, int ell, double L, double Lext,
                                                   int nscatt, bool primary_origin)
{
    const unsigned lane = threadIdx.x & 31;
    const size_t index = hit ? (size_t)l + (size_t)ell * q.npix : 0;
    // lanes with the same pixel-bin and the same number of scatterings route to the same arrays
    const unsigned long long key = hit ? ((index << 8) | (unsigned long long)min(nscatt, 255)) + 1ull : 0ull;
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int maxcnt = __reduce_max_sync(0xffffffffu, hit ? __popc(peers) : 0);
    double sumL = L, sumLext = Lext;
    if (maxcnt > 1)
    {
        sumL = 0.;
        sumLext = 0.;
        unsigned rem = hit ? peers : 0u;
        for (int k = 0; k < maxcnt; ++k)
        {
            const int src = rem ? __ffs(rem) - 1 : (int)lane;
            const double a = __shfl_sync(0xffffffffu, L, src), b = __shfl_sync(0xffffffffu, Lext, src);
            if (rem)
            {
                sumL += a;
                sumLext += b;
            }
            rem &= rem - 1;
        }
    }
    if (hit && lane == (unsigned)(__ffs(peers) - 1))
    {
        int c_ext, c_tr, c_lev;
        sk_route(q, nscatt, primary_origin, c_ext, c_tr, c_lev);
        atomicAdd(&q.ifu[c_ext][index], sumLext);
        if (c_tr >= 0) atomicAdd(&q.ifu[c_tr][index], sumL);
        if (c_lev >= 0) atomicAdd(&q.ifu[c_lev][index], sumLext);
    }
}

// ---------------------------------------------------------------------------------------------------
// per-grid accessors used by the event kernels
// ---------------------------------------------------------------------------------------------------
template <int GRID>
__device__ __forceinline__ double sk_cell_density(const SkDevModel& M, int m)
{
    return GRID == 3 ? M.vrec[m].w : GRID == 2 ? M.cells[m].dens : M.dens[m];
}
// number density of medium component h in cell m (MediumState::numberDensity(m,h))
template <int GRID>
__device__ __forceinline__ double sk_component_density(const SkDevModel& M, int m, int h)
{
    return h == 0 ? sk_cell_density<GRID>(M, m) : M.densx[(size_t)(h - 1) * (size_t)M.ncells + (size_t)m];
}
// the same without the grid as a template parameter (kernels that serve every grid kind)
__device__ __forceinline__ double sk_component_density_any(const SkDevModel& M, int m, int h)
{
    if (h > 0) return M.densx[(size_t)(h - 1) * (size_t)M.ncells + (size_t)m];
    return M.grid_kind == 3 ? M.vrec[m].w : M.grid_kind == 2 ? M.cells[m].dens : M.dens[m];
}
// MediumSystem::opacitySca / opacityExt in cell m for sections that are constant in space (MediumSystem.cpp:632-662): the
// sum over the components of n_h sigma_h, sigma from the table sig[h*nlam + ilam]
__device__ __forceinline__ double sk_opacity_sum(const SkDevModel& M, const double* __restrict__ sig, int ilam, int m)
{
    double result = 0.;
    for (int h = 0; h < M.nmed; ++h) result += sk_component_density_any(M, m, h) * sig[h * M.nlam + ilam];
    return result;
}
// true when (x,y,z) lies in the half-open box of cell c (octree: in lattice coordinates, like the walk itself)
template <int GRID>
__device__ __forceinline__ bool sk_cell_contains(const SkDevModel& M, const SkSmemTables& T, const SkCellPos& c, double x,
                                                 double y, double z)
{
    if (GRID == 2)
    {
        const int size = 1 << (M.maxlevel - c.lev);
        const double ux = sk_lat_coord(x, M.ext[0], M.lat_invh[0]), uy = sk_lat_coord(y, M.ext[1], M.lat_invh[1]),
                     uz = sk_lat_coord(z, M.ext[2], M.lat_invh[2]);
        return ux >= (double)c.ix && ux < (double)(c.ix + size) && uy >= (double)c.iy && uy < (double)(c.iy + size)
               && uz >= (double)c.iz && uz < (double)(c.iz + size);
    }
    return x >= T.X[c.ix] && x < T.X[c.ix + 1] && y >= T.Y[c.iy] && y < T.Y[c.iy + 1] && z >= T.Z[c.iz] && z < T.Z[c.iz + 1];
}
