// launch.cuh -- host launch layer: argument conversion, launch-parameter choice, dispatch over
// the (xplor, virial, single-type) instantiations. This is the B200 counterpart of the
// shift_mode x compute_virial x threads_per_particle switch inside HOOMD's
// gpu_compute_pair_forces<E> driver; threads_per_particle is
// a runtime argument of the kernels here, so only 8 (iso) / 4 (dpd, aniso) variants are compiled.
#ifndef AZP_LAUNCH_CUH_
#define AZP_LAUNCH_CUH_

#include "../../include/azp_b200.h"
#include "pair_kernels.cuh"

#include <map>
#include <mutex>
#include <utility>

namespace azp
    {
// ---- long-row deferral scratch: one small queue per (device, stream), allocated on first use --
constexpr unsigned int kLongRowThreshold = 512; // rows longer than this go to the second pass
constexpr unsigned int kLongRowCapacity = 1u << 16;
constexpr unsigned int kRowsPerGroup = AZP_ROWS_PER_GROUP;
constexpr double kShortRow = 64.0; // mean row capacity up to which rows count as short
struct LongRowScratch
    {
    unsigned int* queue = nullptr;
    unsigned int* count = nullptr;
    };
// defined once in capi.cu
cudaError_t long_row_scratch(cudaStream_t stream, LongRowScratch& out);

template<class S> inline BoxDim<S> convert_box(const azp_box& b)
    {
    BoxDim<S> o;
    for (int d = 0; d < 3; ++d)
        {
        o.L[d] = S(b.L[d]);
        o.Linv[d] = S(1.0) / o.L[d];
        o.periodic[d] = b.periodic[d];
        }
    o.xy = S(b.tilt[0]);
    o.xz = S(b.tilt[1]);
    o.yz = S(b.tilt[2]);
    const bool tilted = (o.xy != S(0)) || (o.xz != S(0)) || (o.yz != S(0));
    const bool periodic3 = b.periodic[0] && b.periodic[1] && b.periodic[2];
    o.flags = (!tilted && periodic3) ? 2 : (tilted ? 1 : 0);
    return o;
    }

template<class S> inline KernelArgs<S> convert_args(const azp_pair_args& a)
    {
    KernelArgs<S> k;
    k.force = static_cast<S*>(a.d_force);
    k.virial = static_cast<S*>(a.d_virial);
    k.torque = static_cast<S*>(a.d_torque);
    k.virial_pitch = (size_t)a.virial_pitch;
    k.pos = static_cast<const S*>(a.d_pos);
    k.vel = static_cast<const S*>(a.d_vel);
    k.orientation = static_cast<const S*>(a.d_orientation);
    k.tag = a.d_tag;
    k.n_neigh = a.d_n_neigh;
    k.nlist = a.d_nlist;
    k.head_list = a.d_head_list;
    k.rcutsq = static_cast<const S*>(a.d_rcutsq);
    k.ronsq = static_cast<const S*>(a.d_ronsq);
    k.row_ids = a.d_row_ids;
    k.box = convert_box<S>(a.box);
    k.N = a.N;
    k.ntypes = a.ntypes;
    k.shift_mode = a.shift_mode;
    k.row_offset = a.row_offset;
    k.n_row_ids = a.n_row_ids;
    k.seed = a.seed & 0xffffu;
    k.timestep = (unsigned int)(a.timestep & 0xffffffffull);
    k.deltaT = S(a.deltaT);
    k.T = S(a.T);
    k.long_queue = nullptr;
    k.long_count = nullptr;
    k.long_capacity = 0;
    k.long_threshold = kLongRowThreshold;
    k.force_b = nullptr;
    k.virial_b = nullptr;
    k.rcutsq_b = nullptr;
    k.params_b = nullptr;
    k.shift_mode_b = 0;
    return k;
    }

struct LaunchShape
    {
    unsigned int block;
    unsigned int tpp_log2;
    unsigned int grid;
    };

inline bool is_pow2(unsigned int x)
    {
    return x && !(x & (x - 1));
    }

// Launch parameters: honour the caller's (block_size, threads_per_particle) -- HOOMD's autotuner
// dimensions -- or choose them from the mean row capacity and the number of rows.
inline cudaError_t choose_shape(const azp_pair_args& a, LaunchShape& s, unsigned int block_limit, bool multirow = false)
    {
    unsigned int block = a.block_size ? a.block_size : 128u;
    if (block % 32u != 0 || block > block_limit)
        return cudaErrorInvalidValue;
    unsigned int tpp = a.threads_per_particle;
    if (tpp == 0)
        {
        const unsigned int rows = a.N ? a.N : 1u;
        const double mean_cap = a.size_neigh_list ? double(a.size_neigh_list) / rows : 64.0;
        // Measured on B200: the kernels are instruction-issue bound, so the fewest
        // lanes per row win (least reduction / prologue work per neighbour) as long as the grid
        // fills the machine: one lane per row up to ~160 neighbours, then split long rows, then
        // split further until there are about five CTAs per SM.
        tpp = 1;
        while (tpp < 32u && mean_cap / tpp > 160.0)
            tpp <<= 1;
        const unsigned long long nslots = a.d_row_ids ? a.n_row_ids : a.N;
        while (tpp < 32u && nslots * tpp / block < 148ull * 5ull)
            tpp <<= 1;
        }
    if (!is_pow2(tpp) || tpp > 32u)
        return cudaErrorInvalidValue;
    unsigned int lg = 0;
    while ((1u << lg) < tpp)
        ++lg;
    const unsigned long long nslots = a.d_row_ids ? a.n_row_ids : a.N;
    const unsigned long long threads = nslots * tpp;
    // rows per group of tpp lanes (row_kernel's main pass walks slot, slot + G, ...): one for
    // long rows; short rows (mean capacity <= kShortRow) take kRowsPerGroup rows per group with
    // the next row's metadata and list line prefetched -- never fewer CTAs than fill the machine
    unsigned long long grid = (threads + block - 1) / block;
    const double mean_cap_rows = a.size_neigh_list && a.N ? double(a.size_neigh_list) / a.N : 1e9;
    if (multirow && kRowsPerGroup > 1 && mean_cap_rows <= kShortRow)
        {
        const unsigned long long g = (grid + kRowsPerGroup - 1) / kRowsPerGroup;
        grid = g > 148ull * 16ull ? g : (grid < 148ull * 16ull ? grid : 148ull * 16ull);
        }
    if (grid > 0x7fffffffull)
        return cudaErrorInvalidValue;
    s.block = block;
    s.tpp_log2 = lg;
    s.grid = (unsigned int)grid;
    return cudaSuccess;
    }

inline cudaError_t check_common(const azp_pair_args* a, const void* d_params)
    {
    if (!a || !d_params)
        return cudaErrorInvalidValue;
    if (a->ntypes == 0)
        return cudaErrorInvalidValue;
    if (a->N == 0 || (a->d_row_ids && a->n_row_ids == 0))
        return cudaSuccess; // nothing to do; caught by the callers before launching
    if (!a->d_force || !a->d_pos || !a->d_n_neigh || !a->d_nlist || !a->d_head_list || !a->d_rcutsq)
        return cudaErrorInvalidValue;
    if (a->compute_virial > 1 || (a->compute_virial && !a->d_virial))
        return cudaErrorInvalidValue;
    return cudaSuccess;
    }

inline bool nothing_to_do(const azp_pair_args* a)
    {
    return a->N == 0 || (a->d_row_ids && a->n_row_ids == 0);
    }

template<class K> inline cudaError_t ensure_smem(K kernel, size_t bytes)
    {
    if (bytes > 227u * 1024u)
        return cudaErrorInvalidValue; // type-pair table does not fit in shared memory
    if (bytes > 48u * 1024u)
        return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return cudaSuccess;
    }

// Launch the main pass and, when the caller announced long rows (n_max) and the rows are split
// over fewer than 32 lanes, the warp-per-row second pass over the deferred rows.
template<class Fam>
inline cudaError_t launch_rows(KernelArgs<typename Fam::S> k, const void* d_params, const LaunchShape& s, unsigned int n_max, cudaStream_t stream)
    {
    typedef typename Fam::E E;
    const size_t ntp = Fam::NTM == 1 ? 1 : size_t(k.ntypes) * k.ntypes;
    const typename E::param_type* params = static_cast<const typename E::param_type*>(d_params);
    const bool defer = n_max > kLongRowThreshold && s.tpp_log2 < 5;
    cudaError_t err;
    if (defer)
        {
        LongRowScratch scratch;
        err = long_row_scratch(stream, scratch);
        if (err != cudaSuccess)
            return err;
        err = cudaMemsetAsync(scratch.count, 0, sizeof(unsigned int), stream);
        if (err != cudaSuccess)
            return err;
        k.long_queue = scratch.queue;
        k.long_count = scratch.count;
        k.long_capacity = kLongRowCapacity;
        }
    const size_t smem = Fam::smem_bytes(ntp, s.block);
#ifndef AZP_NO_ONE_LANE
    // (the staged one-lane kernels are built for blocks of at most 256 threads)
    if (s.tpp_log2 == 0 && s.block <= row_kernel_max_threads<Fam, true>())
        {
        auto kernel = row_kernel<Fam, false, true>;
        // the staged neighbour-list ring (ListStage) sits behind the family's own tables
        const size_t smem1 = smem + ((AZP_STAGE_LIST != 0 && Fam::PIPE == 2) ? ListStage::bytes(s.block) : 0);
        err = ensure_smem(kernel, smem1);
        if (err != cudaSuccess)
            return err;
        kernel<<<s.grid, s.block, smem1, stream>>>(k, params, 0u);
        }
    else
#endif
        {
        auto kernel = row_kernel<Fam, false, false>;
        err = ensure_smem(kernel, smem);
        if (err != cudaSuccess)
            return err;
        kernel<<<s.grid, s.block, smem, stream>>>(k, params, s.tpp_log2);
        }
    err = cudaGetLastError();
    if (err != cudaSuccess || !defer)
        return err;
    auto long_kernel = row_kernel<Fam, true>;
    const unsigned int long_block = 128;
    const size_t long_smem = Fam::smem_bytes(ntp, long_block);
    err = ensure_smem(long_kernel, long_smem);
    if (err != cudaSuccess)
        return err;
    long_kernel<<<148 * 4, long_block, long_smem, stream>>>(k, params, 5u);
    return cudaGetLastError();
    }

template<class E, class S, bool XPLOR, bool VIRIAL, int NTM>
inline cudaError_t launch_pair_variant(const KernelArgs<S>& k, const void* d_params, const LaunchShape& s, unsigned int n_max, cudaStream_t stream)
    {
    return launch_rows<IsoFamily<E, S, XPLOR, VIRIAL, NTM>>(k, d_params, s, n_max, stream);
    }

template<class E, class S> cudaError_t launch_pair(const azp_pair_args* a, const void* d_params, cudaStream_t stream)
    {
    cudaError_t err = check_common(a, d_params);
    if (err != cudaSuccess)
        return err;
    if (a->shift_mode > 2 || (a->shift_mode == 2 && !a->d_ronsq))
        return cudaErrorInvalidValue;
    if (nothing_to_do(a))
        return cudaSuccess;
    LaunchShape s;
    err = choose_shape(*a, s, max_block<S>());
    if (err != cudaSuccess)
        return err;
    const KernelArgs<S> k = convert_args<S>(*a);
    const bool xplor = a->shift_mode == 2, vir = a->compute_virial != 0;
    const int ntm = a->ntypes == 1 ? 1 : (a->ntypes == 2 && IsoTraits<E>::register_tables ? 2 : 0);
#define AZP_CASE(X, V, T)           \
    if (xplor == X && vir == V && ntm == T) \
        return launch_pair_variant<E, S, X, V, T>(k, d_params, s, a->n_max, stream);
    AZP_CASE(false, false, 0)
    AZP_CASE(false, false, 1)
    AZP_CASE(false, false, 2)
    AZP_CASE(false, true, 0)
    AZP_CASE(false, true, 1)
    AZP_CASE(false, true, 2)
    AZP_CASE(true, false, 0)
    AZP_CASE(true, false, 1)
    AZP_CASE(true, false, 2)
    AZP_CASE(true, true, 0)
    AZP_CASE(true, true, 1)
    AZP_CASE(true, true, 2)
#undef AZP_CASE
    return cudaErrorInvalidValue;
    }

// Two isotropic potentials over one sweep of the list (FusedIsoFamily). `a` and `b` must describe
// the same rows, particles and list; outputs, cutoffs, parameters and shift mode (none / shift)
// are per potential. Launch shape and row selection are taken from `a`.
template<class EA, class EB, class S>
cudaError_t launch_pair_fused(const azp_pair_args* a, const void* d_params_a, const azp_pair_args* b, const void* d_params_b, cudaStream_t stream)
    {
    cudaError_t err = check_common(a, d_params_a);
    if (err == cudaSuccess)
        err = check_common(b, d_params_b);
    if (err != cudaSuccess)
        return err;
    if (a->shift_mode > 1 || b->shift_mode > 1)
        return cudaErrorNotSupported; // xplor: evaluate the two potentials separately
    if (a->d_pos != b->d_pos || a->d_n_neigh != b->d_n_neigh || a->d_nlist != b->d_nlist
        || a->d_head_list != b->d_head_list || a->N != b->N || a->ntypes != b->ntypes
        || a->row_offset != b->row_offset || a->d_row_ids != b->d_row_ids || a->n_row_ids != b->n_row_ids
        || a->compute_virial != b->compute_virial || a->virial_pitch != b->virial_pitch)
        return cudaErrorInvalidValue;
    if (nothing_to_do(a))
        return cudaSuccess;
    LaunchShape s;
    err = choose_shape(*a, s, max_block<S>());
    if (err != cudaSuccess)
        return err;
    KernelArgs<S> k = convert_args<S>(*a);
    k.force_b = static_cast<S*>(b->d_force);
    k.virial_b = static_cast<S*>(b->d_virial);
    k.rcutsq_b = static_cast<const S*>(b->d_rcutsq);
    k.params_b = d_params_b;
    k.shift_mode_b = b->shift_mode;
    const bool vir = a->compute_virial != 0;
    const int ntm = a->ntypes == 1 ? 1 : (a->ntypes == 2 ? 2 : 0);
#define AZP_CASE(V, T)        \
    if (vir == V && ntm == T) \
        return launch_rows<FusedIsoFamily<EA, EB, S, V, T>>(k, d_params_a, s, a->n_max, stream);
    AZP_CASE(false, 0)
    AZP_CASE(false, 1)
    AZP_CASE(false, 2)
    AZP_CASE(true, 0)
    AZP_CASE(true, 1)
    AZP_CASE(true, 2)
#undef AZP_CASE
    return cudaErrorInvalidValue;
    }

template<class E, class S, bool VIRIAL, int NTM>
inline cudaError_t launch_dpd_variant(const KernelArgs<S>& k, const void* d_params, const LaunchShape& s, unsigned int n_max, cudaStream_t stream)
    {
    return launch_rows<DpdFamily<E, S, VIRIAL, NTM>>(k, d_params, s, n_max, stream);
    }

template<class E, class S> cudaError_t launch_dpd(const azp_pair_args* a, const void* d_params, cudaStream_t stream)
    {
    cudaError_t err = check_common(a, d_params);
    if (err != cudaSuccess)
        return err;
    if (nothing_to_do(a))
        return cudaSuccess;
    if (!a->d_vel || !a->d_tag)
        return cudaErrorInvalidValue;
    LaunchShape s;
    err = choose_shape(*a, s, max_block<S>(), kRowsPerGroup > 1);
    if (err != cudaSuccess)
        return err;
    const KernelArgs<S> k = convert_args<S>(*a);
    const bool vir = a->compute_virial != 0;
    const int ntm = a->ntypes == 1 ? 1 : (a->ntypes == 2 ? 2 : 0);
    if (vir)
        {
        if (ntm == 1)
            return launch_dpd_variant<E, S, true, 1>(k, d_params, s, a->n_max, stream);
        if (ntm == 2)
            return launch_dpd_variant<E, S, true, 2>(k, d_params, s, a->n_max, stream);
        return launch_dpd_variant<E, S, true, 0>(k, d_params, s, a->n_max, stream);
        }
    if (ntm == 1)
        return launch_dpd_variant<E, S, false, 1>(k, d_params, s, a->n_max, stream);
    if (ntm == 2)
        return launch_dpd_variant<E, S, false, 2>(k, d_params, s, a->n_max, stream);
    return launch_dpd_variant<E, S, false, 0>(k, d_params, s, a->n_max, stream);
    }

template<class E, class S, bool VIRIAL, int NTM>
inline cudaError_t launch_aniso_variant(const KernelArgs<S>& k, const void* d_params, const LaunchShape& s, unsigned int n_max, cudaStream_t stream)
    {
    return launch_rows<AnisoFamily<E, S, VIRIAL, NTM>>(k, d_params, s, n_max, stream);
    }

template<class E, class S> cudaError_t launch_aniso(const azp_pair_args* a, const void* d_params, cudaStream_t stream)
    {
    cudaError_t err = check_common(a, d_params);
    if (err != cudaSuccess)
        return err;
    if (a->shift_mode > 1)
        return cudaErrorInvalidValue; // AnisotropicPair accepts none / shift only
    if (nothing_to_do(a))
        return cudaSuccess;
    if (!a->d_orientation || !a->d_torque)
        return cudaErrorInvalidValue;
    LaunchShape s;
    err = choose_shape(*a, s, max_block<S>(), kRowsPerGroup > 1);
    if (err != cudaSuccess)
        return err;
    const KernelArgs<S> k = convert_args<S>(*a);
    const bool vir = a->compute_virial != 0;
    const int ntm = a->ntypes == 1 ? 1 : (a->ntypes == 2 ? 2 : 0);
    if (vir)
        {
        if (ntm == 1)
            return launch_aniso_variant<E, S, true, 1>(k, d_params, s, a->n_max, stream);
        if (ntm == 2)
            return launch_aniso_variant<E, S, true, 2>(k, d_params, s, a->n_max, stream);
        return launch_aniso_variant<E, S, true, 0>(k, d_params, s, a->n_max, stream);
        }
    if (ntm == 1)
        return launch_aniso_variant<E, S, false, 1>(k, d_params, s, a->n_max, stream);
    if (ntm == 2)
        return launch_aniso_variant<E, S, false, 2>(k, d_params, s, a->n_max, stream);
    return launch_aniso_variant<E, S, false, 0>(k, d_params, s, a->n_max, stream);
    }
    } // namespace azp

#endif
