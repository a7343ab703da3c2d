// TEST INFRASTRUCTURE (oracle). Not part of the product path.
//
// CPU cell-list neighbour search producing HOOMD's NeighborList arrays (SURVEY Appendix A.2 and
// 8(a) a15): `n_neigh[i]` valid entries of row i starting at `nlist[head_list[i]]`, every pair with
// r < r_list(type_i, type_j) at build time, `full` (both directions) or `half` (j > i only)
// storage. Used by the tests to feed identical lists to the oracle and the CUDA path and to
// check the product's GPU builder; rows are sorted by j so the result is deterministic.
#ifndef AZP_ORACLE_NLIST_CPU_H_
#define AZP_ORACLE_NLIST_CPU_H_

#include "driver_loops.h"
#include <algorithm>
#include <vector>

namespace azp_oracle
    {
template<class S> struct CellGrid
    {
    int dim[3];
    bool use_cells;
    std::vector<unsigned int> start, order;
    };

template<class S>
inline void build_grid(CellGrid<S>& g, unsigned int N, const S* pos, const Box<S>& b, S rmax)
    {
    g.use_cells = (b.xy == 0 && b.xz == 0 && b.yz == 0);
    for (int d = 0; d < 3; ++d)
        {
        g.dim[d] = int(b.L[d] / rmax);
        if (g.dim[d] < 3 || !b.periodic[d])
            g.use_cells = g.use_cells && (g.dim[d] >= 3) && b.periodic[d];
        }
    if (!g.use_cells)
        return;
    const size_t nc = size_t(g.dim[0]) * g.dim[1] * g.dim[2];
    std::vector<unsigned int> cell(N);
    g.start.assign(nc + 1, 0);
    for (unsigned int i = 0; i < N; ++i)
        {
        int c[3];
        for (int d = 0; d < 3; ++d)
            {
            S f = pos[4 * size_t(i) + d] / b.L[d] + S(0.5);
            f -= std::floor(f);
            c[d] = std::min(g.dim[d] - 1, std::max(0, int(f * g.dim[d])));
            }
        cell[i] = (unsigned int)((size_t(c[2]) * g.dim[1] + c[1]) * g.dim[0] + c[0]);
        g.start[cell[i] + 1]++;
        }
    for (size_t c = 0; c < nc; ++c)
        g.start[c + 1] += g.start[c];
    g.order.resize(N);
    std::vector<unsigned int> fill(g.start.begin(), g.start.end() - 1);
    for (unsigned int i = 0; i < N; ++i)
        g.order[fill[cell[i]]++] = i;
    }

// pass = 0: count into n_neigh; pass = 1: write rows at head_list
template<class S>
inline void nlist_pass(int pass,
                       unsigned int N,
                       const S* pos,
                       const Box<S>& b,
                       unsigned int ntypes,
                       const S* rlistsq,
                       int half,
                       unsigned int* n_neigh,
                       const uint64_t* head_list,
                       unsigned int* nlist,
                       int nthreads,
                       unsigned int n_rows)
    {
    S rmax = 0;
    for (unsigned int t = 0; t < ntypes * ntypes; ++t)
        rmax = std::max(rmax, S(std::sqrt(rlistsq[t])));
    CellGrid<S> g;
    build_grid(g, N, pos, b, rmax);
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads > 0 ? nthreads : 1)
    for (long long ii = 0; ii < (long long)(n_rows ? n_rows : N); ++ii)
        {
        const unsigned int i = (unsigned int)ii;
        const S* pi = pos + 4 * size_t(i);
        const unsigned int ti = type_of(pi);
        std::vector<unsigned int> row;
        auto consider = [&](unsigned int j)
        {
            if (j == i || (half && j < i))
                return;
            const S* pj = pos + 4 * size_t(j);
            S dx = pi[0] - pj[0], dy = pi[1] - pj[1], dz = pi[2] - pj[2];
            min_image_rint(b, dx, dy, dz);
            const S rsq = dx * dx + dy * dy + dz * dz;
            if (rsq < rlistsq[type_of(pj) * ntypes + ti])
                row.push_back(j);
        };
        if (g.use_cells)
            {
            int c[3];
            for (int d = 0; d < 3; ++d)
                {
                S f = pi[d] / b.L[d] + S(0.5);
                f -= std::floor(f);
                c[d] = std::min(g.dim[d] - 1, std::max(0, int(f * g.dim[d])));
                }
            for (int oz = -1; oz <= 1; ++oz)
                for (int oy = -1; oy <= 1; ++oy)
                    for (int ox = -1; ox <= 1; ++ox)
                        {
                        const int cx = (c[0] + ox + g.dim[0]) % g.dim[0];
                        const int cy = (c[1] + oy + g.dim[1]) % g.dim[1];
                        const int cz = (c[2] + oz + g.dim[2]) % g.dim[2];
                        const size_t cc = (size_t(cz) * g.dim[1] + cy) * g.dim[0] + cx;
                        for (unsigned int s = g.start[cc]; s < g.start[cc + 1]; ++s)
                            consider(g.order[s]);
                        }
            }
        else
            for (unsigned int j = 0; j < N; ++j)
                consider(j);
        std::sort(row.begin(), row.end());
        if (pass == 0)
            n_neigh[i] = (unsigned int)row.size();
        else
            std::copy(row.begin(), row.end(), nlist + head_list[i]);
        }
    }
    } // namespace azp_oracle
#endif
