// nlist_kernels.cu -- HOOMD-layout full neighbour list built on the GPU from a cell list.
//
// Row "next #1" of SURVEY.md 8(f): the step immediately before the pair-force path. Produces
// exactly the arrays the force kernels (and HOOMD's NeighborListGPU) use: n_neigh[i] valid
// entries of row i starting at nlist[head_list[i]], every j != i with
// |minImage(r_i - r_j)|^2 < r_list(type_i, type_j)^2, full storage.
// Orthorhombic and triclinic boxes; rows are ordered by stencil cell (z, y, x from -1 to +1) and
// by particle index inside a cell, so the build is deterministic and, for spatially sorted
// particles, consecutive entries of a row are consecutive indices (coalesced position gathers
// later). Also here: the device-side displacement check and the Morton (SFC) particle order.
#include "../../include/azp_b200.h"
#include "azp_core.cuh"

#include <cub/device/device_radix_sort.cuh>

#include <cmath>
#include <cstdlib>

namespace azp
    {
// Scratch of the builder (sort keys, cub temporaries): stream-ordered allocations from a PRIVATE
// memory pool whose release threshold is unlimited. The default pool gives freed memory back to
// the driver at the next synchronisation (threshold 0), so every rebuild paid the physical
// allocation again -- measured: 28-46 ms per rebuild at N = 32,000, where the kernels take
// 0.2 ms. A private pool leaves the host application's allocator policy alone.
static cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t st)
    {
    static cudaMemPool_t pools[64] = {};
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess)
        return err;
    if (dev < 0 || dev >= 64)
        return cudaMallocAsync(p, bytes, st);
    if (!pools[dev])
        {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool = nullptr;
        err = cudaMemPoolCreate(&pool, &props);
        if (err != cudaSuccess)
            {
            cudaGetLastError();
            return cudaMallocAsync(p, bytes, st);
            }
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        pools[dev] = pool;
        }
    return cudaMallocFromPoolAsync(p, bytes, pools[dev], st);
    }
    } // namespace azp

namespace azp
    {
struct CellGrid
    {
    unsigned int dim[3];
    int reach[3]; // stencil half-width per axis: 1, or 0 when the axis has a single cell
    };

template<class S> struct NlistArgs
    {
    const S* pos;
    unsigned int N;
    unsigned int ntypes;
    BoxDim<S> box;
    const S* rlistsq;
    unsigned int* n_neigh;
    const uint64_t* head_list;
    unsigned int* nlist;
    unsigned int* cell_of;
    unsigned int* cell_start;
    unsigned int* cell_order;
    S* cell_pos;     // positions in cell order (contiguous per cell)
    S* pos_at_build; // optional copy of pos for the displacement check
    CellGrid grid;
    S r_list_max; // largest list cutoff: bounds the stencil of the fine-grid sweep
    bool fine;    // half-width cells (see nlist_rows_fine)
    unsigned int row_offset;
    unsigned int n_rows;
    const unsigned int* capacity; // FILL only: reuse of the previous build's row capacities
    unsigned int* overflow;
    };

// Cell of a position: fractional coordinates of the (possibly triclinic) box -- HOOMD's
// BoxDim::makeFraction -- periodic axes wrapped, non-periodic axes clamped (a particle slightly
// outside the box on a wall-confined axis stays in the boundary cell, next to its neighbours).
template<class S> AZP_D void cell_coords(const BoxDim<S>& b, const CellGrid& g, S x, S y, S z, int c[3])
    {
    const S p[3] = {x - b.xy * y - (b.xz - b.xy * b.yz) * z, y - b.yz * z, z};
#pragma unroll
    for (int d = 0; d < 3; ++d)
        {
        S f = p[d] * b.Linv[d] + S(0.5);
        if (b.periodic[d])
            f -= floor(f);
        int ci = (int)floor(f * S(g.dim[d]));
        ci = max(0, min((int)g.dim[d] - 1, ci));
        c[d] = ci;
        }
    }

// Where a particle is binned. The sweeps take the periodic image of a pair from how the candidate's
// cell was reached, which is only right if cell and position of every particle agree. Per axis:
// the fractional coordinate f = q / L + 1/2 as ONE fused multiply-add of the RAW position (an
// explicit fma: every kernel gets the same bits whatever the compiler contracts elsewhere), the
// whole box lengths k = floor(f) the particle is away from the box (0 on a non-periodic axis),
// the wrapped coordinate fw = f - k that fixes the cell, and the position with the same k lattice
// vectors taken off. Cell and position then agree by construction -- also for a particle exactly
// on a box face, where the rounding of 1 / L decides on which side f falls (fp32, L = 9:
// -4.5 * Linv + 0.5 = -4e-9 -> k = -1, the particle is binned in the LAST cell at +4.5), for a
// float64 coordinate that rounded onto the upper face in fp32, and for a particle a step outside
// the box. Deriving the cell from the shifted position instead would round a second time and can
// land on the other side. A particle with k = 0 on all axes keeps its position bit for bit.
template<class S> struct Binned
    {
    Vec4<S> p; // in-box image
    S fw[3];   // wrapped fractional coordinates, [0, 1] (1 only by rounding)
    };

template<class S> AZP_D Binned<S> bin_particle(const BoxDim<S>& b, const Vec4<S>& p)
    {
    const S q[3] = {p.x - b.xy * p.y - (b.xz - b.xy * b.yz) * p.z, p.y - b.yz * p.z, p.z};
    Binned<S> o;
    S k[3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
        {
        const S f = fma(q[d], b.Linv[d], S(0.5));
        k[d] = b.periodic[d] ? floor(f) : S(0);
        o.fw[d] = f - k[d];
        }
    o.p = p;
    if (k[0] != S(0) || k[1] != S(0) || k[2] != S(0))
        {
        // lattice vectors (Lx, 0, 0), (xy Ly, Ly, 0), (xz Lz, yz Lz, Lz)
        o.p.x -= k[0] * b.L[0] + k[1] * b.xy * b.L[1] + k[2] * b.xz * b.L[2];
        o.p.y -= k[1] * b.L[1] + k[2] * b.yz * b.L[2];
        o.p.z -= k[2] * b.L[2];
        }
    return o;
    }

// cell of wrapped fractional coordinates (a non-periodic axis clamps: a particle slightly outside
// the box on a wall-confined axis stays in the boundary cell, next to its neighbours)
template<class S> AZP_D void grid_cell(const CellGrid& g, const S fw[3], int c[3])
    {
#pragma unroll
    for (int d = 0; d < 3; ++d)
        c[d] = max(0, min((int)g.dim[d] - 1, (int)floor(fw[d] * S(g.dim[d]))));
    }

template<class S> __global__ void nlist_assign_cells(const NlistArgs<S> a, unsigned int* iota)
    {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N)
        return;
    const Vec4<S> p = load4(a.pos, i);
    const Binned<S> bp = bin_particle(a.box, p);
    int c[3];
    grid_cell(a.grid, bp.fw, c);
    a.cell_of[i] = ((unsigned int)c[2] * a.grid.dim[1] + (unsigned int)c[1]) * a.grid.dim[0] + (unsigned int)c[0];
    iota[i] = i;
    if (a.pos_at_build)
        store4(a.pos_at_build, i, p.x, p.y, p.z, p.w);
    }

// cell_start[c] = first position in the sorted key array whose key is >= c (c = 0 .. ncells)
__global__ void nlist_cell_starts(const unsigned int* sorted_cells, unsigned int N, unsigned int ncells, unsigned int* cell_start)
    {
    const unsigned int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > ncells)
        return;
    unsigned int lo = 0, hi = N;
    while (lo < hi)
        {
        const unsigned int mid = (lo + hi) >> 1;
        if (sorted_cells[mid] < c)
            lo = mid + 1;
        else
            hi = mid;
        }
    cell_start[c] = lo;
    }

// cell_pos[s] = binned position of particle cell_order[s]: the candidates of a cell become one
// contiguous run of 16-byte elements (a broadcast / coalesced load in the row kernel instead of
// index -> gather)
template<class S> __global__ void nlist_sorted_positions(const NlistArgs<S> a)
    {
    const unsigned int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.N)
        return;
    const Vec4<S> p = bin_particle(a.box, load4(a.pos, a.cell_order[s])).p;
    store4(a.cell_pos, s, p.x, p.y, p.z, p.w);
    }

// The rows. A group of TPP lanes owns a row; it walks the 27 stencil cells in a fixed order
// (z, y, x from -1 to +1) and, inside a cell, the candidates in cell order (= particle index),
// TPP at a time: one 16-byte load of the cell-sorted position, the displacement, the cutoff test
// of the type pair, and an ordered append (ballot + prefix popcount inside the group), so the
// entries of a row come out in the same deterministic order whatever TPP is.
//
// Minimum image without rint(): the periodic image of a candidate is fixed by how its stencil
// cell was reached -- wrapped below 0 or above dim - 1 on an axis means image -1 / +1 on that
// axis -- so the group applies BoxDim::minImage's own update sequence (z, then y, then x, with
// the tilt factors) with that image number instead of rint(d / L). For every pair closer than
// r_list (<= cell width <= L / 3) the two agree bit for bit; farther candidates fail the cutoff
// either way. An axis with a single cell (box shorter than 3 r_list) keeps rint().
// ORTHO: no tilt factors and at least three cells on every periodic axis -- the image update
// collapses to one multiply-add per axis with the image number of the stencil cell (identical
// bits: the tilt terms of the general sequence subtract exact zeros).
template<class S, bool FILL, unsigned int TPP, bool ORTHO> __global__ void __launch_bounds__(128) nlist_rows(const NlistArgs<S> a)
    {
    // squared list cutoffs of the type pairs (<= 8 types) staged in shared memory
    __shared__ S s_rlistsq[64];
    const bool small_table = a.ntypes <= 8u;
    if (small_table)
        {
        for (unsigned int t = threadIdx.x; t < a.ntypes * a.ntypes; t += blockDim.x)
            s_rlistsq[t] = a.rlistsq[t];
        __syncthreads();
        }
    const unsigned int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int lane = gtid & (TPP - 1u);
    const unsigned int r = gtid / TPP;
    const bool active = r < a.n_rows;
    const unsigned int i = active ? r + a.row_offset : 0u;
    const Binned<S> bi = bin_particle(a.box, load4(a.pos, i));
    const Vec4<S> pi = bi.p;
    const unsigned int ti = min(scalar_as_uint(pi.w), a.ntypes - 1u);
    int c[3];
    grid_cell(a.grid, bi.fw, c);
    unsigned int count = 0;
    unsigned int* row = (FILL && active) ? a.nlist + a.head_list[r] : nullptr;
    const unsigned int cap = (FILL && a.capacity && active) ? a.capacity[r] : 0xffffffffu;
    const unsigned int group_shift = (threadIdx.x & 31u) & ~(TPP - 1u);
    const unsigned int group_mask = TPP == 32u ? 0xffffffffu : ((1u << TPP) - 1u);
    const int dx_ = (int)a.grid.dim[0], dy_ = (int)a.grid.dim[1], dz_ = (int)a.grid.dim[2];
    const BoxDim<S>& b = a.box;
    for (int oz = -a.grid.reach[2]; oz <= a.grid.reach[2]; ++oz)
        for (int oy = -a.grid.reach[1]; oy <= a.grid.reach[1]; ++oy)
            for (int ox = -a.grid.reach[0]; ox <= a.grid.reach[0]; ++ox)
                {
                int cx = c[0] + ox, cy = c[1] + oy, cz = c[2] + oz;
                // image of the candidates of this cell relative to particle i (r_i - r_j is
                // brought back by -img * L): reached by wrapping below 0 -> the candidates sit
                // one box length above -> r_i - r_j is about -L -> img = -1
                S imgx = S(0), imgy = S(0), imgz = S(0);
                bool skip = !active;
                if (cx < 0 || cx >= dx_)
                    {
                    skip = skip || !b.periodic[0];
                    imgx = cx < 0 ? S(-1) : S(1);
                    cx = cx < 0 ? cx + dx_ : cx - dx_;
                    }
                if (cy < 0 || cy >= dy_)
                    {
                    skip = skip || !b.periodic[1];
                    imgy = cy < 0 ? S(-1) : S(1);
                    cy = cy < 0 ? cy + dy_ : cy - dy_;
                    }
                if (cz < 0 || cz >= dz_)
                    {
                    skip = skip || !b.periodic[2];
                    imgz = cz < 0 ? S(-1) : S(1);
                    cz = cz < 0 ? cz + dz_ : cz - dz_;
                    }
                unsigned int s0 = 0, s1 = 0;
                if (!skip)
                    {
                    const unsigned int cell = ((unsigned int)cz * a.grid.dim[1] + (unsigned int)cy) * a.grid.dim[0] + (unsigned int)cx;
                    s0 = a.cell_start[cell], s1 = a.cell_start[cell + 1];
                    }
                // warp-uniform trip count (the groups of a warp may look at different cells)
                const unsigned int trips = __reduce_max_sync(0xffffffffu, (s1 - s0 + TPP - 1u) / TPP);
                for (unsigned int k = 0; k < trips; ++k)
                    {
                    const unsigned int s = s0 + k * TPP + lane;
                    bool pass = false;
                    unsigned int j = 0;
                    if (s < s1)
                        {
                        const Vec4<S> pj = load4(a.cell_pos, s);
                        S x = pi.x - pj.x, y = pi.y - pj.y, z = pi.z - pj.z;
                        if (ORTHO)
                            {
                            x -= b.L[0] * imgx;
                            y -= b.L[1] * imgy;
                            z -= b.L[2] * imgz;
                            }
                        // BoxDim::minImage's sequence with the image numbers of this cell
                        else
                            {
                        if (b.periodic[2])
                            {
                            const S img = a.grid.reach[2] ? imgz : rint_small(z * b.Linv[2]);
                            z -= b.L[2] * img;
                            y -= b.L[2] * b.yz * img;
                            x -= b.L[2] * b.xz * img;
                            }
                        if (b.periodic[1])
                            {
                            const S img = a.grid.reach[1] ? imgy : rint_small(y * b.Linv[1]);
                            y -= b.L[1] * img;
                            x -= b.L[1] * b.xy * img;
                            }
                        if (b.periodic[0])
                            {
                            const S img = a.grid.reach[0] ? imgx : rint_small(x * b.Linv[0]);
                            x -= b.L[0] * img;
                            }
                            }
                        const S rsq = x * x + y * y + z * z;
                        const unsigned int tj = min(scalar_as_uint(pj.w), a.ntypes - 1u);
                        const unsigned int tp = index2d(a.ntypes, ti, tj);
                        if (rsq < (small_table ? s_rlistsq[tp] : a.rlistsq[tp]))
                            {
                            j = a.cell_order[s];
                            pass = j != i;
                            }
                        }
                    const unsigned int votes = (__ballot_sync(0xffffffffu, pass) >> group_shift) & group_mask;
                    if (pass)
                        {
                        const unsigned int slot = count + __popc(votes & ((1u << lane) - 1u));
                        if (FILL && slot < cap)
                            row[slot] = j;
                        }
                    count += __popc(votes);
                    }
                }
    if (!active || lane != 0)
        return;
    if (!FILL)
        a.n_neigh[r] = count;
    else if (a.capacity)
        {
        // capacity reuse: the fill is also the count; a row that needed more slots than it has
        // is truncated and reported (the caller rebuilds with a count pass)
        a.n_neigh[r] = min(count, cap);
        if (count > cap)
            atomicMax(a.overflow, 1u);
        }
    }

template<class S> static BoxDim<S> nlist_box(const azp_box& ab)
    {
    BoxDim<S> b;
    for (int d = 0; d < 3; ++d)
        {
        b.L[d] = S(ab.L[d]);
        b.Linv[d] = S(1.0) / b.L[d];
        b.periodic[d] = ab.periodic[d];
        }
    b.xy = S(ab.tilt[0]);
    b.xz = S(ab.tilt[1]);
    b.yz = S(ab.tilt[2]);
    b.flags = 0;
    return b;
    }

template<class S> static NlistArgs<S> convert(const azp_nlist_args& a)
    {
    NlistArgs<S> k;
    k.pos = static_cast<const S*>(a.d_pos);
    k.N = a.N;
    k.ntypes = a.ntypes;
    k.box = nlist_box<S>(a.box);
    for (int d = 0; d < 3; ++d)
        {
        k.grid.dim[d] = a.cell_dim[d];
        k.grid.reach[d] = a.cell_dim[d] >= 3 ? 1 : 0;
        }
    // fine grid (azp_nlist_cell_dim chose half-width cells): orthorhombic, fully periodic, at
    // least five cells per axis, cells narrower than the list cutoff but two of them span it
    k.r_list_max = S(a.r_list_max);
    k.fine = a.box.tilt[0] == 0.0 && a.box.tilt[1] == 0.0 && a.box.tilt[2] == 0.0;
    for (int d = 0; d < 3; ++d)
        {
        const double w = a.box.L[d] / double(a.cell_dim[d]);
        k.fine = k.fine && a.box.periodic[d] && a.cell_dim[d] >= 5 && w < a.r_list_max && 2.0 * w >= a.r_list_max;
        }
    if (k.fine)
        for (int d = 0; d < 3; ++d)
            k.grid.reach[d] = 2;
    k.rlistsq = static_cast<const S*>(a.d_rlistsq);
    k.n_neigh = a.d_n_neigh;
    k.head_list = a.d_head_list;
    k.nlist = a.d_nlist;
    k.cell_of = a.d_cell_of;
    k.cell_start = a.d_cell_start;
    k.cell_order = a.d_cell_order;
    k.cell_pos = static_cast<S*>(a.d_cell_pos);
    k.pos_at_build = static_cast<S*>(a.d_pos_at_build);
    k.row_offset = a.n_rows ? a.row_offset : 0u;
    k.n_rows = a.n_rows ? a.n_rows : a.N;
    k.capacity = a.d_capacity;
    k.overflow = a.d_overflow;
    return k;
    }

static bool valid_common(const azp_nlist_args* a)
    {
    if (!a || !a->d_pos || !a->d_rlistsq || !a->d_cell_of || !a->d_cell_start || !a->d_cell_order || !a->d_cell_pos)
        return false;
    for (int d = 0; d < 3; ++d)
        if (a->cell_dim[d] == 0 || a->cell_dim[d] == 2)
            return false;
    return a->ntypes > 0;
    }

template<class S> static int bin(const azp_nlist_args* a, cudaStream_t st)
    {
    if (!valid_common(a))
        return (int)cudaErrorInvalidValue;
    if (a->N == 0)
        return 0;
    const NlistArgs<S> k = convert<S>(*a);
    const unsigned int ncells = a->cell_dim[0] * a->cell_dim[1] * a->cell_dim[2];
    unsigned int *iota = nullptr, *sorted_cells = nullptr;
    void* temp = nullptr;
    size_t temp_bytes = 0;
    cudaError_t err = scratch_alloc((void**)&iota, sizeof(unsigned int) * a->N, st);
    if (err != cudaSuccess)
        return (int)err;
    err = scratch_alloc((void**)&sorted_cells, sizeof(unsigned int) * a->N, st);
    if (err != cudaSuccess)
        {
        cudaFreeAsync(iota, st);
        return (int)err;
        }
    const unsigned int block = 256;
    nlist_assign_cells<S><<<(a->N + block - 1) / block, block, 0, st>>>(k, iota);
    int bits = 1;
    while ((1ull << bits) < (unsigned long long)ncells + 1 && bits < 32)
        ++bits;
    cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, a->d_cell_of, sorted_cells, iota, a->d_cell_order, (int)a->N, 0, bits, st);
    err = scratch_alloc(&temp, temp_bytes, st);
    if (err == cudaSuccess)
        {
        cub::DeviceRadixSort::SortPairs(temp, temp_bytes, a->d_cell_of, sorted_cells, iota, a->d_cell_order, (int)a->N, 0, bits, st);
        nlist_cell_starts<<<(ncells + 1 + block - 1) / block, block, 0, st>>>(sorted_cells, a->N, ncells, a->d_cell_start);
        nlist_sorted_positions<S><<<(a->N + block - 1) / block, block, 0, st>>>(k);
        err = cudaGetLastError();
        cudaFreeAsync(temp, st);
        }
    cudaFreeAsync(sorted_cells, st);
    cudaFreeAsync(iota, st);
    return (int)err;
    }

// Fine-grid sweep: cells HALF the list cutoff wide (orthorhombic, fully periodic boxes with at
// least five such cells per axis). The 5 x 5 x 5 stencil is pruned per particle: for every (z, y)
// offset the distance from the particle to the slab of cells bounds what is left of r_list^2 for
// x, and only the cells of that x interval are swept -- as ONE contiguous run of the cell-sorted
// positions (cells that differ only in x are neighbours in memory), plus one more run when the
// interval wraps around the box. About a third of the candidates of the 27 full-width cells
// remain. The bounds carry a slack of 1e-5 L (fp32) for the rounding of the fractional
// coordinates, so no pair inside r_list can be pruned; the exact test on every candidate, the
// image update, the ordered append and the capacity logic are those of nlist_rows. The 25 (z, y)
// offsets are walked by every group of the warp together (the trip counts are warp-uniform).
template<class S, bool FILL, unsigned int TPP> __global__ void __launch_bounds__(128) nlist_rows_fine(const NlistArgs<S> a)
    {
    __shared__ S s_rlistsq[64];
    const bool small_table = a.ntypes <= 8u;
    if (small_table)
        {
        for (unsigned int t = threadIdx.x; t < a.ntypes * a.ntypes; t += blockDim.x)
            s_rlistsq[t] = a.rlistsq[t];
        __syncthreads();
        }
    const unsigned int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int lane = gtid & (TPP - 1u);
    const unsigned int r = gtid / TPP;
    const bool active = r < a.n_rows;
    const unsigned int i = active ? r + a.row_offset : 0u;
    const Binned<S> bi = bin_particle(a.box, load4(a.pos, i));
    const Vec4<S> pi = bi.p;
    const unsigned int ti = min(scalar_as_uint(pi.w), a.ntypes - 1u);
    const BoxDim<S>& b = a.box;
    const int dim[3] = {(int)a.grid.dim[0], (int)a.grid.dim[1], (int)a.grid.dim[2]};
    // grid coordinates of particle i: cell c (grid_cell's arithmetic), position g inside the grid
    S g[3], w[3], slack[3];
    int c[3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
        {
        g[d] = bi.fw[d] * S(dim[d]);
        c[d] = max(0, min(dim[d] - 1, (int)floor(g[d])));
        w[d] = b.L[d] / S(dim[d]);
        slack[d] = b.L[d] * (sizeof(S) == 4 ? S(1e-5) : S(1e-12));
        }
    const S rmax = a.r_list_max;
    const S rmaxsq = rmax * rmax;
    unsigned int count = 0;
    unsigned int* row = (FILL && active) ? a.nlist + a.head_list[r] : nullptr;
    const unsigned int cap = (FILL && a.capacity && active) ? a.capacity[r] : 0xffffffffu;
    const unsigned int group_shift = (threadIdx.x & 31u) & ~(TPP - 1u);
    const unsigned int group_mask = TPP == 32u ? 0xffffffffu : ((1u << TPP) - 1u);
    for (int oz = -2; oz <= 2; ++oz)
        {
        // lower bound of |z_i - z_j| for any j in the slab of cells c_z + o_z
        S dz = oz == 0 ? S(0) : (oz > 0 ? (S(c[2] + oz) - g[2]) : (g[2] - S(c[2] + oz + 1))) * w[2] - slack[2];
        dz = fmax(dz, S(0));
        int cz = c[2] + oz;
        S imgz = S(0);
        if (cz < 0)
            cz += dim[2], imgz = S(-1);
        else if (cz >= dim[2])
            cz -= dim[2], imgz = S(1);
        for (int oy = -2; oy <= 2; ++oy)
            {
            S dy = oy == 0 ? S(0) : (oy > 0 ? (S(c[1] + oy) - g[1]) : (g[1] - S(c[1] + oy + 1))) * w[1] - slack[1];
            dy = fmax(dy, S(0));
            int cy = c[1] + oy;
            S imgy = S(0);
            if (cy < 0)
                cy += dim[1], imgy = S(-1);
            else if (cy >= dim[1])
                cy -= dim[1], imgy = S(1);
            const S remsq = rmaxsq - dz * dz - dy * dy;
            const bool reach_row = active && remsq >= S(0);
            // x interval of cells that can hold a neighbour, within the 5-cell stencil
            int xlo = 1, xhi = 0; // empty
            if (reach_row)
                {
                const S rx = (::sqrt(remsq) + slack[0]) / w[0];
                xlo = max(c[0] - 2, (int)floor(g[0] - rx));
                xhi = min(c[0] + 2, (int)floor(g[0] + rx));
                }
            const unsigned int line = ((unsigned int)cz * a.grid.dim[1] + (unsigned int)cy) * a.grid.dim[0];
            // two runs: the cells inside [0, dim) and the cells that wrapped (below 0 OR above
            // dim - 1: with at least five cells per axis a 5-cell interval cannot do both)
#pragma unroll
            for (int seg = 0; seg < 2; ++seg)
                {
                int a0, a1;
                S imgx = S(0);
                if (seg == 0)
                    a0 = max(xlo, 0), a1 = min(xhi, dim[0] - 1);
                else if (xlo < 0)
                    a0 = xlo + dim[0], a1 = min(xhi, -1) + dim[0], imgx = S(-1);
                else
                    a0 = max(xlo, dim[0]) - dim[0], a1 = xhi - dim[0], imgx = S(1);
                unsigned int s0 = 0, s1 = 0;
                if (reach_row && a0 <= a1)
                    s0 = a.cell_start[line + (unsigned int)a0], s1 = a.cell_start[line + (unsigned int)a1 + 1u];
                const unsigned int trips = __reduce_max_sync(0xffffffffu, (s1 - s0 + TPP - 1u) / TPP);
                for (unsigned int k = 0; k < trips; ++k)
                    {
                    const unsigned int s = s0 + k * TPP + lane;
                    bool pass = false;
                    unsigned int j = 0;
                    if (s < s1)
                        {
                        const Vec4<S> pj = load4(a.cell_pos, s);
                        S x = pi.x - pj.x, y = pi.y - pj.y, z = pi.z - pj.z;
                        x -= b.L[0] * imgx;
                        y -= b.L[1] * imgy;
                        z -= b.L[2] * imgz;
                        const S rsq = x * x + y * y + z * z;
                        const unsigned int tj = min(scalar_as_uint(pj.w), a.ntypes - 1u);
                        const unsigned int tp = index2d(a.ntypes, ti, tj);
                        if (rsq < (small_table ? s_rlistsq[tp] : a.rlistsq[tp]))
                            {
                            j = a.cell_order[s];
                            pass = j != i;
                            }
                        }
                    const unsigned int votes = (__ballot_sync(0xffffffffu, pass) >> group_shift) & group_mask;
                    if (pass)
                        {
                        const unsigned int slot = count + __popc(votes & ((1u << lane) - 1u));
                        if (FILL && slot < cap)
                            row[slot] = j;
                        }
                    count += __popc(votes);
                    }
                }
            }
        }
    if (!active || lane != 0)
        return;
    if (!FILL)
        a.n_neigh[r] = count;
    else if (a.capacity)
        {
        a.n_neigh[r] = min(count, cap);
        if (count > cap)
            atomicMax(a.overflow, 1u);
        }
    }

template<class S, bool FILL, unsigned int TPP> static int rows_tpp(const NlistArgs<S>& k, cudaStream_t st)
    {
    const unsigned int block = 128;
    const unsigned long long threads = (unsigned long long)k.n_rows * TPP;
    const unsigned int grid = (unsigned int)((threads + block - 1) / block);
    if (k.fine)
        {
        nlist_rows_fine<S, FILL, TPP><<<grid, block, 0, st>>>(k);
        return (int)cudaGetLastError();
        }
    // a non-periodic axis has image 0 for every stencil cell that is not skipped, and an axis
    // with a single cell (reach 0) needs rint(): both keep the general sequence
    bool ortho = k.box.xy == S(0) && k.box.xz == S(0) && k.box.yz == S(0);
    for (int d = 0; d < 3; ++d)
        ortho = ortho && k.box.periodic[d] && k.grid.reach[d] == 1;
    if (ortho)
        nlist_rows<S, FILL, TPP, true><<<grid, block, 0, st>>>(k);
    else
        nlist_rows<S, FILL, TPP, false><<<grid, block, 0, st>>>(k);
    return (int)cudaGetLastError();
    }

template<class S, bool FILL> static int rows(const azp_nlist_args* a, cudaStream_t st)
    {
    if (!valid_common(a) || !a->d_n_neigh)
        return (int)cudaErrorInvalidValue;
    if (FILL && (!a->d_head_list || !a->d_nlist))
        return (int)cudaErrorInvalidValue;
    if (FILL && a->d_capacity && !a->d_overflow)
        return (int)cudaErrorInvalidValue;
    if (a->N == 0)
        return 0;
    const NlistArgs<S> k = convert<S>(*a);
    if ((unsigned long long)k.row_offset + k.n_rows > a->N)
        return (int)cudaErrorInvalidValue;
    // lanes per row (a run of candidates is swept TPP at a time). Measured on B200 with the
    // fine-grid sweep (rebuild reusing the capacities): C2, N = 1 M: 2 lanes 1.71 ms, 4 lanes 1.78,
    // 8 lanes 2.14, 16 lanes 2.91; C4, N = 8 M: 2 lanes 4.54 ms, 4 lanes 6.42 -- the sweep is bound
    // by the ordered append, so the fewest lanes that still pair up the 16-byte candidate loads win
    unsigned int tpp = a->threads_per_row;
    if (tpp == 0)
        tpp = 2u;
    switch (tpp)
        {
    case 1:
        return rows_tpp<S, FILL, 1>(k, st);
    case 2:
        return rows_tpp<S, FILL, 2>(k, st);
    case 4:
        return rows_tpp<S, FILL, 4>(k, st);
    case 8:
        return rows_tpp<S, FILL, 8>(k, st);
    case 16:
        return rows_tpp<S, FILL, 16>(k, st);
    case 32:
        return rows_tpp<S, FILL, 32>(k, st);
    default:
        return (int)cudaErrorInvalidValue;
        }
    }

// Displacement check on the device (HOOMD NeighborList::distanceCheck): *flag is raised when any
// particle has moved farther than sqrt(maxsq) from its position at the last build (minimum
// image, so a particle that was wrapped across a face does not count as moved).
template<class S> __global__ void nlist_moved(const S* __restrict__ pos, const S* __restrict__ pos_at_build, const BoxDim<S> box, const S maxsq, const unsigned int N, unsigned int* flag)
    {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool moved = false;
    if (i < N)
        {
        const Vec4<S> p = load4(pos, i);
        const Vec4<S> q = load4(pos_at_build, i);
        S x = p.x - q.x, y = p.y - q.y, z = p.z - q.z;
        min_image_general(box, x, y, z);
        moved = x * x + y * y + z * z > maxsq;
        }
    if (__any_sync(0xffffffffu, moved) && (threadIdx.x & 31u) == 0)
        atomicMax(flag, 1u);
    }

// Morton (Z-order) key of a position: 10 bits per axis of the fractional coordinate
template<class S> __global__ void sfc_keys(const S* __restrict__ pos, const BoxDim<S> box, const unsigned int N, unsigned int* keys, unsigned int* iota)
    {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N)
        return;
    const Vec4<S> p = load4(pos, i);
    CellGrid g;
    g.dim[0] = g.dim[1] = g.dim[2] = 1024u;
    int c[3];
    cell_coords(box, g, p.x, p.y, p.z, c);
    unsigned int key = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d)
        {
        unsigned int v = (unsigned int)c[d] & 0x3ffu;
        v = (v | (v << 16)) & 0x030000ffu;
        v = (v | (v << 8)) & 0x0300f00fu;
        v = (v | (v << 4)) & 0x030c30c3u;
        v = (v | (v << 2)) & 0x09249249u;
        key |= v << d;
        }
    keys[i] = key;
    iota[i] = i;
    }

template<class S> static int sfc_order(const void* d_pos, const azp_box* box, unsigned int N, unsigned int* d_order, cudaStream_t st)
    {
    if (!d_pos || !box || !d_order)
        return (int)cudaErrorInvalidValue;
    if (N == 0)
        return 0;
    unsigned int *keys = nullptr, *keys_out = nullptr, *iota = nullptr;
    void* temp = nullptr;
    size_t temp_bytes = 0;
    cudaError_t err = scratch_alloc((void**)&keys, 3 * sizeof(unsigned int) * (size_t)N, st);
    if (err != cudaSuccess)
        return (int)err;
    keys_out = keys + N;
    iota = keys + 2 * (size_t)N;
    const unsigned int block = 256;
    sfc_keys<S><<<(N + block - 1) / block, block, 0, st>>>(static_cast<const S*>(d_pos), nlist_box<S>(*box), N, keys, iota);
    cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, keys, keys_out, iota, d_order, (int)N, 0, 30, st);
    err = scratch_alloc(&temp, temp_bytes, st);
    if (err == cudaSuccess)
        {
        cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys_out, iota, d_order, (int)N, 0, 30, st);
        err = cudaGetLastError();
        cudaFreeAsync(temp, st);
        }
    cudaFreeAsync(keys, st);
    return (int)err;
    }
    } // namespace azp

namespace azp
    {
// d_dst[k] = d_src[d_idx[k]], rows of `chunks` 16-byte pieces; thread t moves piece t % chunks of
// row t / chunks, so the stores are fully coalesced and the loads are 16-byte gathers.
__global__ void gather_rows_kernel(const uint4* __restrict__ src, const long long* __restrict__ idx, unsigned long long n, unsigned int chunks, uint4* __restrict__ dst)
    {
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * chunks)
        return;
    const unsigned long long k = t / chunks;
    const unsigned int c = (unsigned int)(t - k * chunks);
    dst[t] = __ldg(src + (unsigned long long)idx[k] * chunks + c);
    }
    } // namespace azp

namespace azp
    {
// *(row at dst_addr[k]) = src[idx[k]]: the halo PUSH. The destination addresses may be peer-mapped
// (another GPU's ghost region reached over NVLink); consecutive k of one peer are consecutive
// addresses, so the 16-byte stores coalesce into full NVLink write packets.
__global__ void push_rows_kernel(const uint4* __restrict__ src, const long long* __restrict__ idx, const unsigned long long* __restrict__ dst_addr, unsigned long long n, unsigned int chunks)
    {
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * chunks)
        return;
    const unsigned long long k = t / chunks;
    const unsigned int c = (unsigned int)(t - k * chunks);
    uint4* dst = reinterpret_cast<uint4*>(dst_addr[k]) + c;
    *dst = __ldg(src + (unsigned long long)idx[k] * chunks + c);
    }
    } // namespace azp

extern "C"
    {
    int azp_push_rows(const void* d_src, const int64_t* d_idx, const uint64_t* d_dst_addr, uint64_t n, uint32_t row_bytes, void* stream)
        {
        if (n == 0)
            return 0;
        if (!d_src || !d_idx || !d_dst_addr || row_bytes == 0 || row_bytes % 16 != 0)
            return (int)cudaErrorInvalidValue;
        const unsigned int chunks = row_bytes / 16;
        const unsigned long long threads = n * chunks;
        const unsigned int block = 256;
        azp::push_rows_kernel<<<(unsigned int)((threads + block - 1) / block), block, 0, (cudaStream_t)stream>>>(
            static_cast<const uint4*>(d_src), reinterpret_cast<const long long*>(d_idx),
            reinterpret_cast<const unsigned long long*>(d_dst_addr), n, chunks);
        return (int)cudaGetLastError();
        }

    int azp_gather_rows(const void* d_src, const int64_t* d_idx, uint64_t n, uint32_t row_bytes, void* d_dst, void* stream)
        {
        if (n == 0)
            return 0;
        if (!d_src || !d_idx || !d_dst || row_bytes == 0 || row_bytes % 16 != 0)
            return (int)cudaErrorInvalidValue;
        const unsigned int chunks = row_bytes / 16;
        const unsigned long long threads = n * chunks;
        const unsigned int block = 256;
        azp::gather_rows_kernel<<<(unsigned int)((threads + block - 1) / block), block, 0, (cudaStream_t)stream>>>(
            static_cast<const uint4*>(d_src), reinterpret_cast<const long long*>(d_idx), n, chunks, static_cast<uint4*>(d_dst));
        return (int)cudaGetLastError();
        }

    // Largest grid with cells no smaller than r_list_max. An axis that cannot hold three such
    // cells gets a single cell (the stencil then covers it once and minimum image does the rest).
    // Triclinic boxes: the cell width is measured between the lattice planes (HOOMD's
    // BoxDim::getNearestPlaneDistance).
    int azp_nlist_cell_dim(const azp_box* box, double r_list_max, uint32_t dim[3])
        {
        if (!box || !dim || !(r_list_max > 0))
            return (int)cudaErrorInvalidValue;
        const double xy = box->tilt[0], xz = box->tilt[1], yz = box->tilt[2];
        const double t = xy * yz - xz;
        const double plane[3] = {box->L[0] / sqrt(1.0 + xy * xy + t * t), box->L[1] / sqrt(1.0 + yz * yz), box->L[2]};
        for (int d = 0; d < 3; ++d)
            {
            const double n = plane[d] / r_list_max;
            uint32_t c = n >= 3.0 ? (uint32_t)n : 1u;
            if (c > 1024u)
                c = 1024u;
            dim[d] = c;
            }
        // One word of cell_start and one search thread per cell: a dilute system in a huge box
        // gets at most 2^26 cells (256 MB), with cells wider than the cutoff -- the 27-cell
        // sweep only needs them no narrower. The BASELINE configurations have 0.26 M (C2) to
        // 26 M (C5) half-width cells.
        const unsigned long long max_cells = 1ull << 26;
        auto total = [](const uint32_t* d) { return (unsigned long long)d[0] * d[1] * d[2]; };
        for (int pass = 0; pass < 8 && total(dim) > max_cells; ++pass)
            {
            const double shrink = 0.999 * cbrt(double(max_cells) / double(total(dim)));
            for (int d = 0; d < 3; ++d)
                if (dim[d] >= 3u)
                    {
                    const uint32_t c = (uint32_t)(dim[d] * shrink);
                    dim[d] = c >= 3u ? c : 3u;
                    }
            }
        // Half-width cells for the fine-grid sweep (nlist_rows_fine): orthorhombic, fully periodic
        // boxes that hold at least five such cells per axis. (A grid capped at 1024 cells keeps
        // cells at least half the cutoff wide; if it made them as wide as the cutoff the sweep
        // falls back to the 27-cell one by itself.)
        const bool ortho = xy == 0.0 && xz == 0.0 && yz == 0.0 && box->periodic[0] && box->periodic[1] && box->periodic[2];
        if (ortho && getenv("AZP_NLIST_COARSE") == nullptr)
            {
            uint32_t fine[3];
            bool ok = true;
            for (int d = 0; d < 3; ++d)
                {
                const double n = 2.0 * box->L[d] / r_list_max;
                fine[d] = n > 1024.0 ? 1024u : (uint32_t)n;
                const double w = box->L[d] / fine[d];
                ok = ok && fine[d] >= 5u && w < r_list_max && 2.0 * w >= r_list_max;
                }
            if (ok && total(fine) <= max_cells)
                for (int d = 0; d < 3; ++d)
                    dim[d] = fine[d];
            }
        return 0;
        }
    int azp_nlist_moved_f32(const void* d_pos, const void* d_pos_at_build, const azp_box* box, double max_dist, uint32_t N, uint32_t* d_flag, void* st)
        {
        if (!d_pos || !d_pos_at_build || !box || !d_flag)
            return (int)cudaErrorInvalidValue;
        if (N == 0)
            return 0;
        azp::BoxDim<float> b = azp::nlist_box<float>(*box);
        b.flags = (box->tilt[0] != 0 || box->tilt[1] != 0 || box->tilt[2] != 0) ? 1 : 0;
        azp::nlist_moved<float><<<(N + 255u) / 256u, 256, 0, (cudaStream_t)st>>>(static_cast<const float*>(d_pos), static_cast<const float*>(d_pos_at_build), b, float(max_dist * max_dist), N, d_flag);
        return (int)cudaGetLastError();
        }
    int azp_nlist_moved_f64(const void* d_pos, const void* d_pos_at_build, const azp_box* box, double max_dist, uint32_t N, uint32_t* d_flag, void* st)
        {
        if (!d_pos || !d_pos_at_build || !box || !d_flag)
            return (int)cudaErrorInvalidValue;
        if (N == 0)
            return 0;
        azp::BoxDim<double> b = azp::nlist_box<double>(*box);
        b.flags = (box->tilt[0] != 0 || box->tilt[1] != 0 || box->tilt[2] != 0) ? 1 : 0;
        azp::nlist_moved<double><<<(N + 255u) / 256u, 256, 0, (cudaStream_t)st>>>(static_cast<const double*>(d_pos), static_cast<const double*>(d_pos_at_build), b, max_dist * max_dist, N, d_flag);
        return (int)cudaGetLastError();
        }
    int azp_sfc_order_f32(const void* d_pos, const azp_box* box, uint32_t N, uint32_t* d_order, void* st)
        {
        return azp::sfc_order<float>(d_pos, box, N, d_order, (cudaStream_t)st);
        }
    int azp_sfc_order_f64(const void* d_pos, const azp_box* box, uint32_t N, uint32_t* d_order, void* st)
        {
        return azp::sfc_order<double>(d_pos, box, N, d_order, (cudaStream_t)st);
        }
    int azp_nlist_bin_f32(const azp_nlist_args* a, void* st)
        {
        return azp::bin<float>(a, (cudaStream_t)st);
        }
    int azp_nlist_bin_f64(const azp_nlist_args* a, void* st)
        {
        return azp::bin<double>(a, (cudaStream_t)st);
        }
    int azp_nlist_count_f32(const azp_nlist_args* a, void* st)
        {
        return azp::rows<float, false>(a, (cudaStream_t)st);
        }
    int azp_nlist_count_f64(const azp_nlist_args* a, void* st)
        {
        return azp::rows<double, false>(a, (cudaStream_t)st);
        }
    int azp_nlist_fill_f32(const azp_nlist_args* a, void* st)
        {
        return azp::rows<float, true>(a, (cudaStream_t)st);
        }
    int azp_nlist_fill_f64(const azp_nlist_args* a, void* st)
        {
        return azp::rows<double, true>(a, (cudaStream_t)st);
        }
    }
