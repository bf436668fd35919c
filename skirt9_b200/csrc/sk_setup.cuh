// sk_setup.cuh -- setup on the device: octree construction by the density policy and sampling of the
// medium state, the two steps of the reference's setup that scale with the number of cells.
//
//   DensityTreePolicy::constructTree / needsSubdivide     SKIRT/core/DensityTreePolicy.cpp:116-309
//   OctTreeNode::createChildren                           SKIRT/core/OctTreeNode.cpp:22-35
//   MediumSystem::setupSelfAfter (cell loop)              SKIRT/core/MediumSystem.cpp:286-330, PropertySampler :46-106
//   Geometry::density of the supported geometries         ShellGeometry.cpp:30-36, ExpDiskGeometry.cpp:32-42,
//                                                         RingGeometry.cpp:39-43, SpiralStructureGeometryDecorator.cpp:24-29,71-75
//
// The tree grows level by level like the reference's node list: one thread evaluates one node of the level (its samples
// in sequence, so that the sum has the order of the reference's loop), an exclusive scan of the flags ranks the nodes to
// divide, and the children are appended in that order -- the breadth-first node numbering of constructTree.  Node boxes
// are never stored: a node is its lattice coordinate at the finest level the policy allows, and its borders are entries
// of the per-axis midpoint tables (the same doubles the reference obtains from Box::center recursively).
#pragma once
#include "sk_device.cuh"

#define SK_SETUP_MAX_MEDIA 8
#define SK_STREAM_TREE 0x54524545u  // "TREE"
#define SK_STREAM_CELL 0x43454c4cu  // "CELL"

struct SkDevGeom {
    int32_t kind, pad;
    double number, mass;
    double p[SK_DENSITY_MAX_PARAMS];
};
struct SkDevGeomSet {
    int32_t n, pad;
    SkDevGeom g[SK_SETUP_MAX_MEDIA];
};

// Geometry::density(Position) of a normalised geometry
__device__ __forceinline__ double sk_geom_density(const SkDevGeom& g, double x, double y, double z)
{
    const double* p = g.p;
    switch (g.kind)
    {
        case SK_GEOM_SHELL:
        {
            // SpheGeometry::density(Position) -> ShellGeometry::density(r), ShellGeometry.cpp:30-36
            double r = sqrt(x * x + y * y + z * z);
            if (r < p[0] || r > p[1]) return 0.0;
            return p[3] * pow(r, -p[2]);
        }
        case SK_GEOM_EXPDISK:
        case SK_GEOM_SPIRAL_EXPDISK:
        {
            // ExpDiskGeometry::density(R,z), ExpDiskGeometry.cpp:32-42
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
            if (g.kind == SK_GEOM_EXPDISK) return rho;
            // SpiralStructureGeometryDecorator::density / perturbation, .cpp:24-29,71-75; Vec::cylindrical: phi = atan2(y,x)
            double phi = atan2(y, x);
            double m = p[6], tanp = p[7], R0 = p[8], phi0 = p[9], w = p[10], N = p[11], cn = p[12];
            double gamma = log(R / R0) / tanp + phi0 + 0.5 * M_PI / m;
            double perturbation = (1.0 - w) + w * cn * sk_pow_even(sin(0.5 * m * (gamma - phi)), 2 * N);
            return rho * perturbation;
        }
        case SK_GEOM_RING:
        {
            // RingGeometry::density(R,z), RingGeometry.cpp:39-43
            double R = sqrt(x * x + y * y);
            double u = (R - p[0]) / (M_SQRT2 * p[1]);
            return p[3] * exp(-u * u) * exp(-fabs(z) / p[2]);
        }
    }
    return 0.0;
}

// ---------------------------------------------------------------------------------------------------
// exclusive scan of int32 flags (three passes; block = 1024 items)
// ---------------------------------------------------------------------------------------------------
#define SK_SCAN_BLOCK 1024
__global__ void __launch_bounds__(SK_SCAN_BLOCK) sk_scan_block_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                                      int32_t* __restrict__ block_sum, int n)
{
    __shared__ int32_t warp_sum[32];
    const int i = blockIdx.x * SK_SCAN_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t v = i < n ? in[i] : 0;
    int32_t incl = v;
    for (int o = 1; o < 32; o <<= 1)
    {
        int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0)
    {
        int32_t w = warp_sum[lane];
        int32_t wi = w;
        for (int o = 1; o < 32; o <<= 1)
        {
            int32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        warp_sum[lane] = wi - w;  // exclusive over warps
        if (lane == 31) block_sum[blockIdx.x] = wi;
    }
    __syncthreads();
    if (i < n) out[i] = warp_sum[warp] + incl - v;
}
// one block: exclusive scan of the block sums in place, total to *total
__global__ void __launch_bounds__(SK_SCAN_BLOCK) sk_scan_sums_kernel(int32_t* __restrict__ block_sum, int nblocks,
                                                                     int32_t* __restrict__ total)
{
    __shared__ int32_t warp_sum[32];
    __shared__ int32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nblocks; base += SK_SCAN_BLOCK)
    {
        const int i = base + threadIdx.x;
        int32_t v = i < nblocks ? block_sum[i] : 0;
        int32_t incl = v;
        for (int o = 1; o < 32; o <<= 1)
        {
            int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        if (warp == 0)
        {
            int32_t w = warp_sum[lane];
            int32_t wi = w;
            for (int o = 1; o < 32; o <<= 1)
            {
                int32_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_sum[lane] = wi - w;
        }
        __syncthreads();
        const int32_t c = carry;
        if (i < nblocks) block_sum[i] = c + warp_sum[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == SK_SCAN_BLOCK - 1) carry = c + warp_sum[warp] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(SK_SCAN_BLOCK) sk_scan_add_kernel(int32_t* __restrict__ out,
                                                                    const int32_t* __restrict__ block_sum, int n)
{
    const int i = blockIdx.x * SK_SCAN_BLOCK + threadIdx.x;
    if (i < n) out[i] += block_sum[blockIdx.x];
}

// ---------------------------------------------------------------------------------------------------
// octree construction
// ---------------------------------------------------------------------------------------------------
struct SkTreeBuild {
    const double *xv, *yv, *zv;  // midpoint tables at the policy's finest level, N+1 entries each
    int N;                       // 1 << max_level
    int min_level, max_level, num_samples;
    double max_fraction, max_tau, max_dispersion, kappa, dust_mass;
    uint32_t seed;
};

// DensityTreePolicy::needsSubdivide for the nodes [lbeg, lend) of one level, DensityTreePolicy.cpp:116-227
__global__ void __launch_bounds__(128) sk_tree_evaluate_kernel(SkTreeBuild B, SkDevGeomSet G, const uint4* __restrict__ coord,
                                                               int lbeg, int lend, int32_t* __restrict__ divide)
{
    const int l = lbeg + blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= lend) return;
    const uint4 c = coord[l];
    const int lev = (int)c.w;
    int flag = 0;
    if (lev < B.min_level)
        flag = 1;
    else if (lev < B.max_level)
    {
        const int size = B.N >> lev;
        const double xmin = B.xv[c.x], xmax = B.xv[c.x + size];
        const double ymin = B.yv[c.y], ymax = B.yv[c.y + size];
        const double zmin = B.zv[c.z], zmax = B.zv[c.z + size];
        SkRng g;
        sk_rng_init(g, B.seed, SK_STREAM_TREE, (unsigned long long)l, 0);
        double rhosum = 0., rhomin = DBL_MAX, rhomax = 0.;
        for (int i = 0; i != B.num_samples; ++i)
        {
            // Random::position(Box), Random.cpp:168-176 with Box::fracPos
            double ux = sk_uniform(g);
            double uy = sk_uniform(g);
            double uz = sk_uniform(g);
            double x = xmin + ux * (xmax - xmin);
            double y = ymin + uy * (ymax - ymin);
            double z = zmin + uz * (zmax - zmin);
            double rhoi = 0.;
            for (int h = 0; h < G.n; ++h) rhoi += G.g[h].mass * sk_geom_density(G.g[h], x, y, z);
            rhosum += rhoi;
            if (rhoi < rhomin) rhomin = rhoi;
            if (rhoi > rhomax) rhomax = rhoi;
        }
        const double rho = rhosum / B.num_samples;
        const double dx = xmax - xmin, dy = ymax - ymin, dz = zmax - zmin;
        const double V = dx * dy * dz;  // Box::volume
        const double M = rho * V;
        if (B.max_fraction > 0. && M / B.dust_mass > B.max_fraction) flag = 1;
        if (B.max_tau > 0. && B.kappa * rho * sqrt(dx * dx + dy * dy + dz * dz) > B.max_tau) flag = 1;
        if (B.max_dispersion > 0.)
        {
            double q = rhomax > 0 ? (rhomax - rhomin) / rhomax : 0.;
            if (q > B.max_dispersion) flag = 1;
        }
    }
    divide[l - lbeg] = flag;
}

// TreeNode::subdivide -> OctTreeNode::createChildren (OctTreeNode.cpp:22-35) for the flagged nodes of the level; the
// children of the r-th flagged node become nodes lend + 8r .. lend + 8r + 7 (child c: bit 0 = upper x half, bit 1 = y,
// bit 2 = z)
__global__ void __launch_bounds__(128) sk_tree_subdivide_kernel(int N, uint4* __restrict__ coord, int32_t* __restrict__ first_child,
                                                                int lbeg, int lend, const int32_t* __restrict__ divide,
                                                                const int32_t* __restrict__ rank)
{
    const int l = lbeg + blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= lend) return;
    if (!divide[l - lbeg])
    {
        first_child[l] = -1;
        return;
    }
    const int fc = lend + 8 * rank[l - lbeg];
    first_child[l] = fc;
    const uint4 c = coord[l];
    const unsigned half = (unsigned)(N >> ((int)c.w + 1));
#pragma unroll
    for (int ch = 0; ch < 8; ++ch)
        coord[fc + ch] = make_uint4(c.x + ((ch & 1) ? half : 0u), c.y + ((ch & 2) ? half : 0u), c.z + ((ch & 4) ? half : 0u),
                                    c.w + 1u);
}

// A node list handed in by the caller (sk_engine_set_grid_octree): levels and lattice coordinates of the children of the
// nodes of level L, one pass per level.  Coordinates are top-aligned on the 2^SK_MAX_TREE_LEVEL lattice (child bit c of a
// node of level L sits at bit SK_MAX_TREE_LEVEL-1-L) and are shifted down once the depth of the tree is known.
// flags: bit 0 child index out of range or not after its parent, bit 1 node with two parents, bit 2 deeper than
// SK_MAX_TREE_LEVEL, bit 3 node not reachable from the root; status[1] = deepest level
#define SK_TREE_UNKNOWN_LEVEL 0xffffffffu
__global__ void sk_tree_propagate_kernel(const int32_t* __restrict__ first_child, int nn, unsigned L, uint4* __restrict__ coord,
                                         int32_t* __restrict__ parent, int* __restrict__ status)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nn) return;
    const uint4 c = coord[l];
    if (c.w != L) return;
    const int fc = first_child[l];
    if (fc < 0) return;
    if (fc <= l || fc + 8 > nn)
    {
        atomicOr(&status[0], 1);
        return;
    }
    if (L >= SK_MAX_TREE_LEVEL)
    {
        atomicOr(&status[0], 4);
        return;
    }
    const unsigned bit = 1u << (SK_MAX_TREE_LEVEL - 1 - L);
    for (int ch = 0; ch < 8; ++ch)
    {
        if (atomicCAS(&parent[fc + ch], -1, l) != -1)
        {
            atomicOr(&status[0], 2);
            continue;
        }
        coord[fc + ch] = make_uint4(c.x | ((ch & 1) ? bit : 0u), c.y | ((ch & 2) ? bit : 0u), c.z | ((ch & 4) ? bit : 0u), L + 1);
    }
    atomicMax(&status[1], (int)L + 1);
}
__global__ void sk_tree_init_kernel(uint4* __restrict__ coord, int32_t* __restrict__ parent, int nn)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nn) return;
    coord[l] = make_uint4(0, 0, 0, l == 0 ? 0u : SK_TREE_UNKNOWN_LEVEL);
    parent[l] = -1;
}
__global__ void sk_tree_check_kernel(const uint4* __restrict__ coord, int nn, int* __restrict__ status)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < nn && coord[l].w == SK_TREE_UNKNOWN_LEVEL) atomicOr(&status[0], 8);
}

// after the last level: rescale the lattice coordinates to the deepest level actually reached, flag the leaves
__global__ void sk_tree_leaf_flags_kernel(uint4* __restrict__ coord, const int32_t* __restrict__ first_child, int nn, int shift,
                                          int32_t* __restrict__ leaf)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nn) return;
    if (shift)
    {
        uint4 c = coord[l];
        c.x >>= shift;
        c.y >>= shift;
        c.z >>= shift;
        coord[l] = c;
    }
    leaf[l] = first_child[l] < 0 ? 1 : 0;
}
// TreeSpatialGrid::_cellindexv / _idv (TreeSpatialGrid.cpp:40-48): cells are the leaves in node order
__global__ void sk_tree_number_cells_kernel(const int32_t* __restrict__ first_child, const int32_t* __restrict__ cell_rank, int nn,
                                            int32_t* __restrict__ node_child, int32_t* __restrict__ node_of_cell)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nn) return;
    const int fc = first_child[l];
    if (fc >= 0)
        node_child[l] = fc;
    else
    {
        const int m = cell_rank[l];
        node_child[l] = -(m + 1);
        node_of_cell[m] = l;
    }
}

// ---------------------------------------------------------------------------------------------------
// medium state: volume and sampled number density per cell, MediumSystem.cpp:286-330 with PropertySampler :80-106
// ---------------------------------------------------------------------------------------------------
template <int GRID>
__global__ void __launch_bounds__(128) sk_sample_medium_kernel(SkDevModel M, SkDevGeom g, int num_samples, uint32_t seed, int ncells,
                                                               SkCellRec* __restrict__ cells, double* __restrict__ dens,
                                                               double* __restrict__ volume)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= ncells) return;
    int ix, iy, iz, size;
    if (GRID == 1)
    {
        // m = k + Nz*j + Nz*Ny*i, CartesianSpatialGrid.cpp:210-213
        iz = m % M.nz;
        iy = (m / M.nz) % M.ny;
        ix = m / (M.nz * M.ny);
        size = 1;
    }
    else
    {
        const uint4 c = reinterpret_cast<const uint4*>(M.cell_coord)[m];
        ix = (int)c.x;
        iy = (int)c.y;
        iz = (int)c.z;
        size = 1 << (M.maxlevel - (int)c.w);
    }
    const double xmin = M.xv[ix], xmax = M.xv[ix + size];
    const double ymin = M.yv[iy], ymax = M.yv[iy + size];
    const double zmin = M.zv[iz], zmax = M.zv[iz + size];
    double n;
    if (num_samples == 1)
    {
        // SpatialGrid::centralPositionInCell = Box::center, Box.hpp:135
        n = g.number * sk_geom_density(g, 0.5 * (xmin + xmax), 0.5 * (ymin + ymax), 0.5 * (zmin + zmax));
    }
    else
    {
        SkRng r;
        sk_rng_init(r, seed, SK_STREAM_CELL, (unsigned long long)m, 0);
        double sum = 0.;
        for (int i = 0; i != num_samples; ++i)
        {
            double ux = sk_uniform(r);
            double uy = sk_uniform(r);
            double uz = sk_uniform(r);
            double x = xmin + ux * (xmax - xmin);
            double y = ymin + uy * (ymax - ymin);
            double z = zmin + uz * (zmax - zmin);
            sum += g.number * sk_geom_density(g, x, y, z);
        }
        n = sum / num_samples;
    }
    volume[m] = (xmax - xmin) * (ymax - ymin) * (zmax - zmin);
    if (GRID == 1)
        dens[m] = n;
    else
        cells[m].dens = n;
}
__global__ void sk_gather_density_kernel(const SkCellRec* __restrict__ cells, const double4* __restrict__ vrec, int ncells,
                                         double* __restrict__ out)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < ncells) out[m] = cells ? cells[m].dens : vrec[m].w;
}

// ---------------------------------------------------------------------------------------------------
// Voronoi tessellation on the device: VoronoiMeshSnapshot::buildMesh (VoronoiMeshSnapshot.cpp:491-730),
// which the reference delegates to the vendored voro++ (container::compute_cell per site, then neighbors / volume / vertices).
// One thread builds one cell the way voro++ does: the domain box around the site is cut with the bisecting planes towards
// the other sites, visiting the blocks of a uniform search grid outwards, until every unvisited site is farther than twice
// the cell's largest vertex distance.  The cell is a list of vertices, each the meeting point of three planes (simple cells:
// sites in general position), in coordinates relative to the site; per-thread state lives in local memory (a cell has a few
// dozen planes and vertices).  Same operations in the same order as oracle/sk_oracle.c (vc_*), so lists, volumes and boxes
// are bit-identical.  Output per cell: the neighbours whose plane carries a face, in the order they cut the cell, then the
// walls -1..-6 that do; volume; enclosing box.
// ---------------------------------------------------------------------------------------------------
#define SK_VC_MAXP 96   // planes kept per cell: 6 walls + the bisectors that have cut it so far (unused ones are dropped when full)
#define SK_VC_MAXT 192  // vertices
#define SK_VC_MAXNB 64  // faces of a finished cell (slots of the temporary neighbour table)
struct SkVCell {
    int np, nt;
    double pn[SK_VC_MAXP][3], pd[SK_VC_MAXP];
    int pid[SK_VC_MAXP];
    unsigned char ta[SK_VC_MAXT], tb[SK_VC_MAXT], tc[SK_VC_MAXT];
    double vx[SK_VC_MAXT], vy[SK_VC_MAXT], vz[SK_VC_MAXT];
    double rmax2;
};
struct SkVoronoiBuild {
    double ext[6];
    double w;            // width of the cubic search blocks
    int gx, gy, gz, n;
    const double* sites; // [3n]
    const int32_t* start; // [blocks+1] first entry of each block in `order`
    const int32_t* order; // site indices block by block, ascending inside a block
    const int32_t* blk;   // block of every site
    int32_t* nbr;         // [n * SK_VC_MAXNB]
    int32_t* count;       // [n]
    double* volume;       // [n]
    double* box;          // [6n]
    int* error;           // first failure: -2 buffers, -3 degenerate sites, -4 coinciding sites, -5 too many faces
};

__device__ __forceinline__ bool sk_vc_vertex(const SkVCell& c, int a, int b, int d, double& x, double& y, double& z)
{
    const double* A = c.pn[a];
    const double* B = c.pn[b];
    const double* C = c.pn[d];
    const double bcx = B[1] * C[2] - B[2] * C[1], bcy = B[2] * C[0] - B[0] * C[2], bcz = B[0] * C[1] - B[1] * C[0];
    const double det = A[0] * bcx + A[1] * bcy + A[2] * bcz;
    if (det == 0.) return false;
    const double cax = C[1] * A[2] - C[2] * A[1], cay = C[2] * A[0] - C[0] * A[2], caz = C[0] * A[1] - C[1] * A[0];
    const double abx = A[1] * B[2] - A[2] * B[1], aby = A[2] * B[0] - A[0] * B[2], abz = A[0] * B[1] - A[1] * B[0];
    const double da = c.pd[a], db = c.pd[b], dd = c.pd[d];
    x = (da * bcx + db * cax + dd * abx) / det;
    y = (da * bcy + db * cay + dd * aby) / det;
    z = (da * bcz + db * caz + dd * abz) / det;
    return true;
}
__device__ __forceinline__ void sk_vc_rmax(SkVCell& c)
{
    double r = 0.;
    for (int t = 0; t < c.nt; ++t)
    {
        const double q = c.vx[t] * c.vx[t] + c.vy[t] * c.vy[t] + c.vz[t] * c.vz[t];
        if (q > r) r = q;
    }
    c.rmax2 = r;
}
__device__ __noinline__ int sk_vc_clip(SkVCell& c, double nx, double ny, double nz, double d, int id)
{
    unsigned char out[SK_VC_MAXT];
    int nout = 0;
    for (int t = 0; t < c.nt; ++t)
    {
        out[t] = (nx * c.vx[t] + ny * c.vy[t] + nz * c.vz[t] - d) > 0.;
        nout += out[t];
    }
    if (!nout) return 0;
    if (nout == c.nt) return -1;
    if (c.np >= SK_VC_MAXP)
    {
        unsigned char map[SK_VC_MAXP], keep[SK_VC_MAXP];
        for (int j = 0; j < c.np; ++j) keep[j] = j < 6;
        for (int t = 0; t < c.nt; ++t) keep[c.ta[t]] = keep[c.tb[t]] = keep[c.tc[t]] = 1;
        int np = 0;
        for (int j = 0; j < c.np; ++j)
        {
            map[j] = (unsigned char)np;
            if (!keep[j]) continue;
            c.pn[np][0] = c.pn[j][0];
            c.pn[np][1] = c.pn[j][1];
            c.pn[np][2] = c.pn[j][2];
            c.pd[np] = c.pd[j];
            c.pid[np] = c.pid[j];
            np++;
        }
        c.np = np;
        for (int t = 0; t < c.nt; ++t)
        {
            c.ta[t] = map[c.ta[t]];
            c.tb[t] = map[c.tb[t]];
            c.tc[t] = map[c.tc[t]];
        }
        if (c.np >= SK_VC_MAXP) return -2;
    }
    const int P = c.np++;
    c.pn[P][0] = nx;
    c.pn[P][1] = ny;
    c.pn[P][2] = nz;
    c.pd[P] = d;
    c.pid[P] = id;
    // the edges (pairs of planes) of the removed vertices that lead to a kept vertex: those that occur once among them
    unsigned char ea[3 * SK_VC_MAXT], eb[3 * SK_VC_MAXT], eo[3 * SK_VC_MAXT];
    int ne = 0;
    for (int t = 0; t < c.nt; ++t)
    {
        if (!out[t]) continue;
        const unsigned char pa[3] = {c.ta[t], c.ta[t], c.tb[t]}, pb[3] = {c.tb[t], c.tc[t], c.tc[t]};
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
    int nt = 0;
    for (int t = 0; t < c.nt; ++t)
        if (!out[t])
        {
            c.ta[nt] = c.ta[t];
            c.tb[nt] = c.tb[t];
            c.tc[nt] = c.tc[t];
            c.vx[nt] = c.vx[t];
            c.vy[nt] = c.vy[t];
            c.vz[nt] = c.vz[t];
            nt++;
        }
    for (int q = 0; q < ne; ++q)
    {
        if (eo[q] > 2) return -3;
        if (eo[q] != 1) continue;
        if (nt >= SK_VC_MAXT) return -2;
        c.ta[nt] = ea[q];
        c.tb[nt] = eb[q];
        c.tc[nt] = (unsigned char)P;
        if (!sk_vc_vertex(c, ea[q], eb[q], P, c.vx[nt], c.vy[nt], c.vz[nt])) return -3;
        nt++;
    }
    c.nt = nt;
    sk_vc_rmax(c);
    return 0;
}
__device__ __noinline__ int sk_vc_measure(const SkVCell& c, double& volume, double box[6], unsigned char* used)
{
    for (int j = 0; j < c.np; ++j) used[j] = 0;
    box[0] = box[1] = box[2] = DBL_MAX;
    box[3] = box[4] = box[5] = -DBL_MAX;
    for (int t = 0; t < c.nt; ++t)
    {
        used[c.ta[t]] = used[c.tb[t]] = used[c.tc[t]] = 1;
        if (c.vx[t] < box[0]) box[0] = c.vx[t];
        if (c.vy[t] < box[1]) box[1] = c.vy[t];
        if (c.vz[t] < box[2]) box[2] = c.vz[t];
        if (c.vx[t] > box[3]) box[3] = c.vx[t];
        if (c.vy[t] > box[4]) box[4] = c.vy[t];
        if (c.vz[t] > box[5]) box[5] = c.vz[t];
    }
    double V = 0.;
    for (int f = 0; f < c.np; ++f)
    {
        if (!used[f]) continue;
        int t0 = -1;
        for (int t = 0; t < c.nt && t0 < 0; ++t)
            if (c.ta[t] == f || c.tb[t] == f || c.tc[t] == f) t0 = t;
        int o1 = c.ta[t0] == f ? c.tb[t0] : c.ta[t0];
        int via = c.tc[t0] == f ? c.tb[t0] : c.tc[t0];
        if (via == o1) via = c.tc[t0];
        int cur = t0, steps = 0;
        double sum = 0.;
        double px = 0., py = 0., pz = 0.;
        while (true)
        {
            int next = -1;
            for (int t = 0; t < c.nt; ++t)
            {
                if (t == cur) continue;
                const int a = c.ta[t], b = c.tb[t], d = c.tc[t];
                if ((a == f || b == f || d == f) && (a == via || b == via || d == via)) next = t;
            }
            if (next < 0 || ++steps > c.nt) return -3;
            const int a = c.ta[next], b = c.tb[next], d = c.tc[next];
            const int other = (a != f && a != via) ? a : (b != f && b != via) ? b : d;
            if (next == t0) break;
            if (steps >= 2)
            {
                const double ax = c.vx[t0], ay = c.vy[t0], az = c.vz[t0];
                const double bx = c.vx[next], by = c.vy[next], bz = c.vz[next];
                sum += ax * (py * bz - pz * by) + ay * (pz * bx - px * bz) + az * (px * by - py * bx);
            }
            px = c.vx[next];
            py = c.vy[next];
            pz = c.vz[next];
            via = other;
            cur = next;
        }
        V += fabs(sum);
    }
    volume = V / 6.;
    return 0;
}

__global__ void __launch_bounds__(64) sk_voronoi_build_kernel(const SkVoronoiBuild B)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= B.n) return;
    SkVCell c;
    const double p[3] = {B.sites[3 * (size_t)m], B.sites[3 * (size_t)m + 1], B.sites[3 * (size_t)m + 2]};
    c.np = 6;
    c.nt = 0;
    for (int w = 0; w < 6; ++w)
    {
        const int axis = w >> 1, upper = w & 1;
        c.pn[w][0] = c.pn[w][1] = c.pn[w][2] = 0.;
        c.pn[w][axis] = upper ? 1. : -1.;
        c.pd[w] = upper ? B.ext[axis + 3] - p[axis] : -(B.ext[axis] - p[axis]);
        c.pid[w] = -(w + 1);
    }
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            for (int k = 0; k < 2; ++k)
            {
                const int t = c.nt++;
                c.ta[t] = (unsigned char)i;
                c.tb[t] = (unsigned char)(2 + j);
                c.tc[t] = (unsigned char)(4 + k);
                sk_vc_vertex(c, i, 2 + j, 4 + k, c.vx[t], c.vy[t], c.vz[t]);
            }
    sk_vc_rmax(c);
    const int gx = B.gx, gy = B.gy, gz = B.gz;
    const int b0 = B.blk[m];
    const int bi = b0 / (gy * gz), bj = (b0 / gz) % gy, bk = b0 % gz;
    const int smax = gx > gy ? (gx > gz ? gx : gz) : (gy > gz ? gy : gz);
    int rc = 0;
    for (int s = 0; s <= smax && !rc; ++s)
    {
        // every site in a block at Chebyshev distance s or more is at least (s-1) w away
        if (s >= 2 && (double)(s - 1) * B.w * (double)(s - 1) * B.w >= 4. * c.rmax2) break;
        for (int i = bi - s; i <= bi + s && !rc; ++i)
        {
            if (i < 0 || i >= gx) continue;
            for (int j = bj - s; j <= bj + s && !rc; ++j)
            {
                if (j < 0 || j >= gy) continue;
                const bool shell = (i == bi - s || i == bi + s || j == bj - s || j == bj + s);
                for (int k = bk - s; k <= bk + s && !rc; k += (shell || s == 0) ? 1 : 2 * s)
                {
                    if (k < 0 || k >= gz) continue;
                    const size_t b = ((size_t)i * gy + j) * gz + k;
                    const int q1 = B.start[b + 1];
                    for (int q = B.start[b]; q < q1 && !rc; ++q)
                    {
                        const int mi = B.order[q];
                        if (mi == m) continue;
                        const double nx = B.sites[3 * (size_t)mi] - p[0], ny = B.sites[3 * (size_t)mi + 1] - p[1],
                                     nz = B.sites[3 * (size_t)mi + 2] - p[2];
                        const double n2 = nx * nx + ny * ny + nz * nz;
                        if (n2 >= 4. * c.rmax2) continue;  // its bisector lies beyond the farthest vertex
                        rc = n2 == 0. ? -4 : sk_vc_clip(c, nx, ny, nz, 0.5 * n2, mi);
                    }
                }
            }
        }
    }
    unsigned char used[SK_VC_MAXP];
    double b6[6], vol = 0.;
    if (!rc) rc = sk_vc_measure(c, vol, b6, used);
    int cnt = 0;
    if (!rc)
    {
        int32_t* out = B.nbr + (size_t)m * SK_VC_MAXNB;
        for (int j = 6; j < c.np && !rc; ++j)
            if (used[j])
            {
                if (cnt >= SK_VC_MAXNB) rc = -5;
                else out[cnt++] = c.pid[j];
            }
        for (int j = 0; j < 6 && !rc; ++j)
            if (used[j])
            {
                if (cnt >= SK_VC_MAXNB) rc = -5;
                else out[cnt++] = c.pid[j];
            }
    }
    if (rc)
    {
        atomicCAS(B.error, 0, rc);
        B.count[m] = 0;
        return;
    }
    B.count[m] = cnt;
    B.volume[m] = vol;
    for (int a = 0; a < 3; ++a)
    {
        B.box[6 * (size_t)m + a] = p[a] + b6[a];
        B.box[6 * (size_t)m + a + 3] = p[a] + b6[a + 3];
    }
}
// packs the per-cell neighbour slots into the lists the grid takes (offsets from an exclusive scan of the counts)
__global__ void sk_voronoi_compact_kernel(const int32_t* __restrict__ nbr, const int32_t* __restrict__ count,
                                          const int32_t* __restrict__ offset, int n, int32_t* __restrict__ out)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    const int c = count[m], o = offset[m];
    for (int i = 0; i < c; ++i) out[o + i] = nbr[(size_t)m * SK_VC_MAXNB + i];
}

// ---------------------------------------------------------------------------------------------------
// ParticleMedium: the smoothed-particle density of ParticleSnapshot::density(Position) (ParticleSnapshot.cpp:248-258) with the
// CubicSplineSmoothingKernel (CubicSplineSmoothingKernel.cpp:39-47), sampled per cell by the cell loop of
// MediumSystem::setupSelfAfter (MediumSystem.cpp:286-330, PropertySampler :46-106).  Like the reference's BoxSearch
// (BoxSearch.cpp:201-207) the sum runs over the particles listed for the search block that holds the position, in ascending
// particle index; particles that do not reach the position contribute an exact zero, so the value is that of the sum over all
// particles in ascending index.  One thread per cell, its samples in sequence.
// ---------------------------------------------------------------------------------------------------
struct SkSphSearch {
    double ext[6];
    double inv[3];         // blocks per unit length
    int nbx, nby, nbz, np;
    const int32_t* start;  // [blocks+1]
    const int32_t* list;   // particle indices per block, ascending
    const double* part;    // [5 np] x y z h M
};
__device__ __forceinline__ double sk_sph_kernel_density(double u)
{
    if (u < 0.0 || u >= 1.0)
        return 0.0;
    else if (u < 0.5)
        return 8.0 / M_PI * (1.0 - 6.0 * u * u * (1.0 - u));
    else
        return 8.0 / M_PI * 2.0 * (1.0 - u) * (1.0 - u) * (1.0 - u);
}
__device__ __forceinline__ double sk_sph_density(const SkSphSearch& S, double x, double y, double z)
{
    int i = (int)((x - S.ext[0]) * S.inv[0]), j = (int)((y - S.ext[1]) * S.inv[1]), k = (int)((z - S.ext[2]) * S.inv[2]);
    i = i < 0 ? 0 : i >= S.nbx ? S.nbx - 1 : i;
    j = j < 0 ? 0 : j >= S.nby ? S.nby - 1 : j;
    k = k < 0 ? 0 : k >= S.nbz ? S.nbz - 1 : k;
    const size_t b = ((size_t)i * S.nby + j) * S.nbz + k;
    const int q1 = S.start[b + 1];
    double sum = 0.;
    for (int q = S.start[b]; q < q1; ++q)
    {
        const double* p = S.part + 5 * (size_t)__ldg(&S.list[q]);
        const double dx = x - __ldg(p), dy = y - __ldg(p + 1), dz = z - __ldg(p + 2);
        const double h = __ldg(p + 3);
        const double r2 = dx * dx + dy * dy + dz * dz;
        if (r2 >= h * h) continue;
        const double u = sqrt(r2) / h;
        sum += sk_sph_kernel_density(u) * (__ldg(p + 4) / (h * h * h));  // Particle::density(), ParticleSnapshot.cpp:45
    }
    return sum > 0. ? sum : 0.;
}
template <int GRID>
__global__ void __launch_bounds__(128) sk_sample_particles_kernel(SkDevModel M, SkSphSearch S, double scale, int num_samples, uint32_t seed,
                                                                  int ncells, SkCellRec* __restrict__ cells, double4* __restrict__ vrec,
                                                                  double* __restrict__ dens, double* __restrict__ volume)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= ncells) return;
    double b0, b1, b2, b3, b4, b5;
    if (GRID == 3)
    {
        const double* bx = M.vbox + 6 * (size_t)m;
        b0 = bx[0];
        b1 = bx[1];
        b2 = bx[2];
        b3 = bx[3];
        b4 = bx[4];
        b5 = bx[5];
    }
    else
    {
        int ix, iy, iz, size;
        if (GRID == 1)
        {
            iz = m % M.nz;
            iy = (m / M.nz) % M.ny;
            ix = m / (M.nz * M.ny);
            size = 1;
        }
        else
        {
            const uint4 c = reinterpret_cast<const uint4*>(M.cell_coord)[m];
            ix = (int)c.x;
            iy = (int)c.y;
            iz = (int)c.z;
            size = 1 << (M.maxlevel - (int)c.w);
        }
        b0 = M.xv[ix];
        b3 = M.xv[ix + size];
        b1 = M.yv[iy];
        b4 = M.yv[iy + size];
        b2 = M.zv[iz];
        b5 = M.zv[iz + size];
    }
    double n;
    if (num_samples == 1)
        n = sk_sph_density(S, 0.5 * (b0 + b3), 0.5 * (b1 + b4), 0.5 * (b2 + b5)) * scale;
    else
    {
        SkRng r;
        sk_rng_init(r, seed, SK_STREAM_CELL, (unsigned long long)m, 0);
        double sum = 0.;
        for (int i = 0; i != num_samples; ++i)
        {
            double x, y, z;
            if (GRID == 3)
            {
                // VoronoiMeshSnapshot::generatePosition(m), VoronoiMeshSnapshot.cpp:976-989
                const double4 sm = vrec[m];
                const long long i0 = M.vnbr_off[m], i1 = M.vnbr_off[m + 1];
                bool found = false;
                x = sm.x;
                y = sm.y;
                z = sm.z;
                for (int it = 0; it < 10000 && !found; ++it)
                {
                    const double ux = sk_uniform(r), uy = sk_uniform(r), uz = sk_uniform(r);
                    const double px = b0 + ux * (b3 - b0);
                    const double py = b1 + uy * (b4 - b1);
                    const double pz = b2 + uz * (b5 - b2);
                    const double dx = px - sm.x, dy = py - sm.y, dz = pz - sm.z;
                    const double target = dx * dx + dy * dy + dz * dz;
                    found = true;
                    for (long long q = i0; q < i1; ++q)
                    {
                        const int id = __ldg(&M.vnbr[q]);
                        if (id < 0) continue;
                        const double4 t = vrec[id];
                        const double ex = px - t.x, ey = py - t.y, ez = pz - t.z;
                        if (ex * ex + ey * ey + ez * ez < target)
                        {
                            found = false;
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
                const double ux = sk_uniform(r);
                const double uy = sk_uniform(r);
                const double uz = sk_uniform(r);
                x = b0 + ux * (b3 - b0);
                y = b1 + uy * (b4 - b1);
                z = b2 + uz * (b5 - b2);
            }
            sum += sk_sph_density(S, x, y, z) * scale;
        }
        n = sum / num_samples;
    }
    if (GRID == 1)
        dens[m] = n;
    else
        dens[m] = n;  // written into the cell records by the caller's kernel (octree: SkCellRec::dens, Voronoi: vrec.w)
    if (GRID != 3) volume[m] = (b3 - b0) * (b4 - b1) * (b5 - b2);
}
