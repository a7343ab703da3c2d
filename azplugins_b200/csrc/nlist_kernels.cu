// nlist_kernels.cu -- HOOMD-layout full neighbour list built on the GPU from a cell list.
//
// Row "next #1" of SURVEY.md 8(f): the step immediately before the pair-force path. Produces
// exactly the arrays the force kernels (and HOOMD's NeighborListGPU) use: n_neigh[i] valid
// entries of row i starting at nlist[head_list[i]], every j != i with
// |minImage(r_i - r_j)|^2 < r_list(type_i, type_j)^2, full storage.
// Orthorhombic boxes; rows are ordered by stencil cell (z, y, x from -1 to +1) and by particle
// index inside a cell, so the build is deterministic and, for spatially sorted particles,
// consecutive entries of a row are consecutive indices (coalesced position gathers later).
#include "../../include/azp_b200.h"
#include "azp_core.cuh"

#include <cub/device/device_radix_sort.cuh>

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
    CellGrid grid;
    unsigned int row_offset;
    unsigned int n_rows;
    const unsigned int* capacity; // FILL only: reuse of the previous build's row capacities
    unsigned int* overflow;
    };

template<class S> AZP_D void cell_coords(const BoxDim<S>& b, const CellGrid& g, S x, S y, S z, int c[3])
    {
    const S p[3] = {x, y, z};
#pragma unroll
    for (int d = 0; d < 3; ++d)
        {
        S f = p[d] * b.Linv[d] + S(0.5);
        f -= floor(f);
        int ci = (int)(f * S(g.dim[d]));
        ci = max(0, min((int)g.dim[d] - 1, ci));
        c[d] = ci;
        }
    }

template<class S> __global__ void nlist_assign_cells(const NlistArgs<S> a, unsigned int* iota)
    {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N)
        return;
    const Vec4<S> p = load4(a.pos, i);
    int c[3];
    cell_coords(a.box, a.grid, p.x, p.y, p.z, c);
    a.cell_of[i] = ((unsigned int)c[2] * a.grid.dim[1] + (unsigned int)c[1]) * a.grid.dim[0] + (unsigned int)c[0];
    iota[i] = i;
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

template<class S, bool FILL> __global__ void __launch_bounds__(128) nlist_rows(const NlistArgs<S> a)
    {
    const unsigned int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n_rows)
        return;
    const unsigned int i = r + a.row_offset;
    const Vec4<S> pi = load4(a.pos, i);
    const unsigned int ti = scalar_as_uint(pi.w);
    int c[3];
    cell_coords(a.box, a.grid, pi.x, pi.y, pi.z, c);
    unsigned int count = 0;
    unsigned int* row = FILL ? a.nlist + a.head_list[r] : nullptr;
    const unsigned int cap = (FILL && a.capacity) ? a.capacity[r] : 0xffffffffu;
    for (int oz = -a.grid.reach[2]; oz <= a.grid.reach[2]; ++oz)
        for (int oy = -a.grid.reach[1]; oy <= a.grid.reach[1]; ++oy)
            for (int ox = -a.grid.reach[0]; ox <= a.grid.reach[0]; ++ox)
                {
                int cx = c[0] + ox, cy = c[1] + oy, cz = c[2] + oz;
                const int dx_ = (int)a.grid.dim[0], dy_ = (int)a.grid.dim[1], dz_ = (int)a.grid.dim[2];
                if (cx < 0 || cx >= dx_)
                    {
                    if (!a.box.periodic[0])
                        continue;
                    cx = (cx + dx_) % dx_;
                    }
                if (cy < 0 || cy >= dy_)
                    {
                    if (!a.box.periodic[1])
                        continue;
                    cy = (cy + dy_) % dy_;
                    }
                if (cz < 0 || cz >= dz_)
                    {
                    if (!a.box.periodic[2])
                        continue;
                    cz = (cz + dz_) % dz_;
                    }
                const unsigned int cell = ((unsigned int)cz * a.grid.dim[1] + (unsigned int)cy) * a.grid.dim[0] + (unsigned int)cx;
                const unsigned int s0 = a.cell_start[cell], s1 = a.cell_start[cell + 1];
                for (unsigned int s = s0; s < s1; ++s)
                    {
                    const unsigned int j = a.cell_order[s];
                    if (j == i)
                        continue;
                    const Vec4<S> pj = load4(a.pos, j);
                    S dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                    min_image_general(a.box, dx, dy, dz);
                    const S rsq = dx * dx + dy * dy + dz * dz;
                    if (rsq < a.rlistsq[index2d(a.ntypes, ti, scalar_as_uint(pj.w))])
                        {
                        if (FILL && count < cap)
                            row[count] = j;
                        ++count;
                        }
                    }
                }
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

template<class S> static NlistArgs<S> convert(const azp_nlist_args& a)
    {
    NlistArgs<S> k;
    k.pos = static_cast<const S*>(a.d_pos);
    k.N = a.N;
    k.ntypes = a.ntypes;
    for (int d = 0; d < 3; ++d)
        {
        k.box.L[d] = S(a.box.L[d]);
        k.box.Linv[d] = S(1.0) / k.box.L[d];
        k.box.periodic[d] = a.box.periodic[d];
        k.grid.dim[d] = a.cell_dim[d];
        k.grid.reach[d] = a.cell_dim[d] >= 3 ? 1 : 0;
        }
    k.box.xy = k.box.xz = k.box.yz = S(0);
    k.box.flags = 0;
    k.rlistsq = static_cast<const S*>(a.d_rlistsq);
    k.n_neigh = a.d_n_neigh;
    k.head_list = a.d_head_list;
    k.nlist = a.d_nlist;
    k.cell_of = a.d_cell_of;
    k.cell_start = a.d_cell_start;
    k.cell_order = a.d_cell_order;
    k.row_offset = a.n_rows ? a.row_offset : 0u;
    k.n_rows = a.n_rows ? a.n_rows : a.N;
    k.capacity = a.d_capacity;
    k.overflow = a.d_overflow;
    return k;
    }

static bool valid_common(const azp_nlist_args* a)
    {
    if (!a || !a->d_pos || !a->d_rlistsq || !a->d_cell_of || !a->d_cell_start || !a->d_cell_order)
        return false;
    if (a->box.tilt[0] != 0 || a->box.tilt[1] != 0 || a->box.tilt[2] != 0)
        return false; // orthorhombic boxes only
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
    cudaError_t err = cudaMallocAsync(&iota, sizeof(unsigned int) * a->N, st);
    if (err != cudaSuccess)
        return (int)err;
    err = cudaMallocAsync(&sorted_cells, sizeof(unsigned int) * a->N, st);
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
    err = cudaMallocAsync(&temp, temp_bytes, st);
    if (err == cudaSuccess)
        {
        cub::DeviceRadixSort::SortPairs(temp, temp_bytes, a->d_cell_of, sorted_cells, iota, a->d_cell_order, (int)a->N, 0, bits, st);
        nlist_cell_starts<<<(ncells + 1 + block - 1) / block, block, 0, st>>>(sorted_cells, a->N, ncells, a->d_cell_start);
        err = cudaGetLastError();
        cudaFreeAsync(temp, st);
        }
    cudaFreeAsync(sorted_cells, st);
    cudaFreeAsync(iota, st);
    return (int)err;
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
    const unsigned int block = 128;
    nlist_rows<S, FILL><<<(k.n_rows + block - 1) / block, block, 0, st>>>(k);
    return (int)cudaGetLastError();
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
    int azp_nlist_cell_dim(const azp_box* box, double r_list_max, uint32_t dim[3])
        {
        if (!box || !dim || !(r_list_max > 0))
            return (int)cudaErrorInvalidValue;
        for (int d = 0; d < 3; ++d)
            {
            const double n = box->L[d] / r_list_max;
            uint32_t c = n >= 3.0 ? (uint32_t)n : 1u;
            if (c > 1024u)
                c = 1024u;
            dim[d] = c;
            }
        return 0;
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
