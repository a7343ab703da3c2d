// pair_kernels.cuh -- the sm_100a neighbour-list pair-force kernels.
//
// Replaces the kernel bodies HOOMD-blue supplies to azplugins through
//   gpu_compute_pair_forces<E>        (reference src/PotentialPairGPUKernel.cu.inc:25-28)
//   gpu_compute_dpd_forces<E>         (reference src/PotentialPairDPDThermoGPUKernel.cu.inc:21-24)
//   gpu_compute_pair_aniso_forces<E>  (reference src/AnisoPotentialPairGPUKernel.cu.inc:21-25)
// with a from-scratch design for B200 (SURVEY.md 3.2-3.4 give the per-pair sequence kept here).
//
// One kernel skeleton (row_kernel) drives three "families" (IsoFamily, DpdFamily, AnisoFamily)
// that differ in what they gather per neighbour and what they accumulate:
//
//   * one row (particle) per group of `tpp` consecutive lanes, tpp a runtime power of two <= 32;
//   * the row is consumed as aligned uint4 vectors of four neighbour indices: one 16-byte
//     coalesced nlist load, then four independent 16-byte position gathers in flight per lane
//     (ld.global.nc, L1-resident for spatially sorted particles) before any math. The first and
//     last partial vectors of a row go through a guarded scalar path, so rows may start at any
//     head_list offset and nothing outside [head, head + n_neigh) is ever read;
//   * per type-pair constants (Evaluator::cache_type: parameters, derived constants, energy at
//     r_cut) and the effective r_cut^2 are built once per CTA in shared memory; a pair whose
//     potential is switched off gets r_cut^2 = 0, so ONE compare per neighbour implements the
//     reference's "rsq < rcutsq && parameter != 0" test. Single-type systems keep the constants
//     in registers for the whole kernel, two-type systems per row (TypeLookup, NTM = 1 / 2 / 0);
//   * minimum image costs 3 full-rate instructions per axis (magic-number rint, no FRND) and is
//     skipped for a whole warp when every row of the warp is farther than the largest cutoff
//     from all periodic faces: for such rows wrapping can only change pairs that fail the
//     cutoff test either way, so the result is bit-identical;
//   * forces accumulate with explicit FMAs (identical bits with or without the virial);
//     force/energy/virial/torque are reduced over the tpp lanes with xor-shuffles and written by
//     lane 0 as one 16-byte store (+ one for the torque, + 6 virial scalars).
// There is no tensor-core work here: the path is a gather-bound stencil.
#ifndef AZP_PAIR_KERNELS_CUH_
#define AZP_PAIR_KERNELS_CUH_

#include "azp_core.cuh"
#include "azp_philox.cuh"

#include <type_traits>

// rows per group of lanes in the main pass of the short-row families (DPD, anisotropic); 1 = one
// row per group (default: measured, the multi-row walk with next-row prefetch gains 2-4 % on
// C4/C5 but its loop state costs every kernel 20-30 registers when compiled in)
#ifndef AZP_ROWS_PER_GROUP
#define AZP_ROWS_PER_GROUP 1
#endif
#ifndef AZP_TRIP_WRAP
#define AZP_TRIP_WRAP 0
#endif
// queued neighbours a lane evaluates per heavy round of the deferred-accept families (1 or 2)
// (measured, C4 4 M / C5 4 M: DPD 1 -> 2: 0.685 -> 0.748 ms; two-patch Morse with the queue
// 1 -> 2: 0.452 -> 0.439 ms, in-place evaluation 0.454 ms)
#ifndef AZP_HEAVY_UNROLL
#define AZP_HEAVY_UNROLL 1
#endif
#ifndef AZP_ANISO_HEAVY_UNROLL
#define AZP_ANISO_HEAVY_UNROLL 2
#endif
// lines (128 B) ahead of a lane's cursor that the pipelined loop requests into L2; 0 = off
#ifndef AZP_NLIST_LINE_PREFETCH
#define AZP_NLIST_LINE_PREFETCH 0
#endif
// Neighbour-list staging of the one-lane-per-row pipelined loop (ListStage below): 0 = off (each
// lane loads its row 16 bytes per trip through L1), 1 = TMA bulk copies (cp.async.bulk +
// mbarrier), 2 = cp.async (LDGSTS) 16-byte copies
#ifndef AZP_STAGE_LIST
#define AZP_STAGE_LIST 0
#endif

namespace azp
    {
template<class S> struct KernelArgs
    {
    S* force;
    S* virial;
    S* torque;
    size_t virial_pitch;
    const S* pos;
    const S* vel;
    const S* orientation;
    const unsigned int* tag;
    const unsigned int* n_neigh;
    const unsigned int* nlist;
    const uint64_t* head_list;
    const S* rcutsq;
    const S* ronsq;
    const unsigned int* row_ids;
    BoxDim<S> box;
    unsigned int N;
    unsigned int ntypes;
    unsigned int shift_mode;
    unsigned int row_offset;
    unsigned int n_row_ids;
    unsigned int seed;
    unsigned int timestep;
    S deltaT;
    S T;
    // long-row deferral (see row_kernel): rows longer than long_threshold are queued by the main
    // pass and evaluated by a second pass with a whole warp per row
    unsigned int* long_queue;
    unsigned int* long_count;
    unsigned int long_capacity;
    unsigned int long_threshold;
    // second potential of a fused two-potential pass (FusedIsoFamily): own outputs and tables,
    // same particles and list
    S* force_b;
    S* virial_b;
    const S* rcutsq_b;
    const void* params_b;
    unsigned int shift_mode_b;
    };

// largest block_size accepted: 512 threads (<= 128 registers) for fp32; the fp64 variants hold
// twice the register state (software pipeline + accumulators), so they are compiled for blocks of
// at most 256 threads (<= 255 registers) instead of spilling.
constexpr unsigned int kMaxBlock = 512;
template<class S> constexpr unsigned int max_block()
    {
    return sizeof(S) == 8 ? 256u : kMaxBlock;
    }

template<class S> AZP_D S shfl_xor(S v, unsigned int o)
    {
    return __shfl_xor_sync(0xffffffffu, v, o);
    }

// xplor smoothing coefficients per type pair
template<class S> struct XplorEntry
    {
    S ronsq;
    S denom_inv; // 1 / (rcutsq - ronsq)^3
    };

// ---------------------------------------------------------------------------------------------
// Per-row geometry shared by the families: particle i, its type, and the displacement rule.
// ---------------------------------------------------------------------------------------------
template<class S> struct RowGeometry
    {
    Vec4<S> pi;
    unsigned int ti;
    bool skip_wrap;
    S Lx, Ly, Lz, ix, iy, iz;

    AZP_D void displacement(const BoxDim<S>& b, const Vec4<S>& pj, S& dx, S& dy, S& dz) const
        {
        dx = pi.x - pj.x;
        dy = pi.y - pj.y;
        dz = pi.z - pj.z;
        if (!skip_wrap)
            wrap(b, dx, dy, dz);
        }
    AZP_D void wrap(const BoxDim<S>& b, S& dx, S& dy, S& dz) const
        {
        if (b.flags == 2)
            min_image_ortho(Lx, Ly, Lz, ix, iy, iz, dx, dy, dz);
        else
            min_image_general(b, dx, dy, dz);
        }
    // the same with the skip_wrap decision taken by the caller (once per trip of four neighbours)
    template<bool WRAP> AZP_D void displacement_t(const BoxDim<S>& b, const Vec4<S>& pj, S& dx, S& dy, S& dz) const
        {
        dx = pi.x - pj.x;
        dy = pi.y - pj.y;
        dz = pi.z - pj.z;
        if (WRAP)
            wrap(b, dx, dy, dz);
        }
    };

// Virial accumulators (compiled out when VIRIAL is false).
template<class S> struct Virial6
    {
    S xx = S(0), xy = S(0), xz = S(0), yy = S(0), yz = S(0), zz = S(0);
    AZP_D void reduce(unsigned int o)
        {
        xx += shfl_xor(xx, o);
        xy += shfl_xor(xy, o);
        xz += shfl_xor(xz, o);
        yy += shfl_xor(yy, o);
        yz += shfl_xor(yz, o);
        zz += shfl_xor(zz, o);
        }
    AZP_D void store(S* virial, size_t pitch, unsigned int row) const
        {
        S* v = virial + row;
        v[0] = S(0.5) * xx;
        v[pitch] = S(0.5) * xy;
        v[2 * pitch] = S(0.5) * xz;
        v[3 * pitch] = S(0.5) * yy;
        v[4 * pitch] = S(0.5) * yz;
        v[5 * pitch] = S(0.5) * zz;
        }
    };

// Shared-memory type-pair table common to the families: cache[ntp], rcutsq_eff[ntp], rc_max.
// The table is addressed through the `extern __shared__` array itself (never through a stored
// generic pointer), so every access compiles to an LDS with a known address space.
extern __shared__ __align__(16) unsigned char azp_smem[];

template<class E, class S> struct PairTable
    {
    typedef typename E::cache_type Cache;
    unsigned int base_off;   // byte offset of the cache array (0 unless a second table follows)
    unsigned int rcutsq_off; // byte offset of rcutsq[]

    AZP_HD static size_t bytes(size_t ntp)
        {
        return ntp * sizeof(Cache) + (ntp + 1) * sizeof(S);
        }
    AZP_D void carve(unsigned int ntp, unsigned int base = 0u)
        {
        base_off = base;
        rcutsq_off = base + ntp * (unsigned int)sizeof(Cache);
        }
    AZP_D Cache& cache(unsigned int t) const
        {
        return reinterpret_cast<Cache*>(azp_smem + base_off)[t];
        }
    // effective r_cut^2: 0 for pairs whose potential is switched off
    AZP_D S& rcutsq(unsigned int t) const
        {
        return reinterpret_cast<S*>(azp_smem + rcutsq_off)[t];
        }
    AZP_D unsigned int end_off(unsigned int ntp) const
        {
        return rcutsq_off + (ntp + 1) * (unsigned int)sizeof(S);
        }
    // sqrt(max rcutsq), for the interior-warp test; stored after rcutsq[ntp - 1]
    AZP_D void finish(unsigned int ntp)
        {
        if (threadIdx.x == 0)
            {
            S m = S(0);
            for (unsigned int t = 0; t < ntp; ++t)
                m = fmax(m, rcutsq(t));
            rcutsq(ntp) = ::sqrt(m);
            }
        }
    };

// Word-wise select between two POD structs held in registers (compiles to one SEL per live word).
template<class C> AZP_D C select_words(bool pick_b, const C& a, const C& b)
    {
    static_assert(sizeof(C) % 4 == 0, "cache_type must be a multiple of 4 bytes");
    union U
        {
        C c;
        uint32_t w[sizeof(C) / 4];
        AZP_D U() { }
        };
    U ua, ub, uo;
    ua.c = a;
    ub.c = b;
#pragma unroll
    for (unsigned int k = 0; k < sizeof(C) / 4; ++k)
        uo.w[k] = pick_b ? ub.w[k] : ua.w[k];
    return uo.c;
    }

// Per-lane queue of accepted neighbours ("deferred accept") for the families whose
// per-pair work is heavy and sits behind the cutoff test (DPD thermostat: Philox + pow; two-patch
// Morse: quaternion rotations + three exponentials). Only 36 % (C4) / 51 % (C5) of the entries of
// a row pass the test, and with one row per lane a warp would run the heavy path for nearly every
// entry with that fraction of its lanes active. Instead the cheap scan (gather, minimum image,
// cutoff test) pushes the index of an accepted neighbour into a small queue in shared memory, and
// the heavy path runs in warp-wide rounds that pop one index per lane -- only when some lane's
// queue could overflow on the next trip, and at the end of the row. A round gathers the position
// again (an L1 hit: the scan has just loaded it) and recomputes the displacement; keeping only
// the 4-byte index keeps the queue at 4 KB per 128-thread CTA, so the unified L1 stays a cache
// for the position gathers (a 24-byte slot cut the L1 hit rate from 74 % to 22 %).
// Layout: slot s of thread t at word s * blockDim.x + t (conflict-free).
struct AcceptQueue
    {
#ifndef AZP_QUEUE_SLOTS
#define AZP_QUEUE_SLOTS 8
#endif
    static constexpr unsigned int Q = AZP_QUEUE_SLOTS; // slots per lane
    static constexpr unsigned int ROOM = 4; // a trip pushes at most 4 entries per lane
    unsigned int n = 0;
    unsigned int off; // byte offset into azp_smem

    AZP_HD static size_t bytes(size_t block)
        {
        return Q * block * sizeof(unsigned int) + 16;
        }
    AZP_D void carve(unsigned int byte_off)
        {
        off = (byte_off + 15u) & ~15u;
        n = 0;
        }
    AZP_D bool needs_drain() const
        {
        return n + ROOM > Q;
        }
    AZP_D void push(unsigned int j)
        {
        reinterpret_cast<unsigned int*>(azp_smem + off)[n * blockDim.x + threadIdx.x] = j;
        ++n;
        }
    AZP_D unsigned int pop()
        {
        --n;
        return reinterpret_cast<const unsigned int*>(azp_smem + off)[n * blockDim.x + threadIdx.x];
        }
    };

// Staging of the neighbour-list stream through shared memory (AZP_STAGE_LIST). With one lane per
// row a lane's 16-byte index load is a fresh HBM miss every other trip, and the bytes a warp
// keeps in flight (32 x 16) bound the stream by latency. Here every lane copies
// the next CH index vectors of ITS OWN row into a private slot of a two-stage ring --
// asynchronously, without registers: as one TMA bulk copy (cp.async.bulk, completion on the
// warp's mbarrier of that stage) or as CH cp.async copies -- two chunks (8 trips) ahead of the
// math, and reads the indices of a trip back with one LDS.128. Slots are PITCH = 16 CH + 16
// bytes apart, so the 32 lanes of a warp read 32 different bank groups (20 words mod 32 walks
// the multiples of 4). No lane ever reads another lane's slot: no cross-lane exchange of row
// pointers, and the cp.async variant needs no barrier at all. Only full vectors inside the row
// are staged; the partial vectors at the row ends keep their guarded scalar path.
struct ListStage
    {
    static constexpr unsigned int CH = 4;                 // index vectors per row and chunk
    static constexpr unsigned int PITCH = 16u * CH + 16u; // bytes per row slot
    static constexpr unsigned int STAGES = 2;
    static constexpr unsigned int WARP_BYTES = STAGES * 32u * PITCH;
    unsigned int slots; // byte offset (azp_smem) of this lane's slot of stage 0
    unsigned int bars;  // byte offset of this warp's mbarriers (one per stage)

    AZP_HD static size_t bytes(size_t block)
        {
        return (block / 32u) * (WARP_BYTES + 8u * STAGES) + 128u;
        }
    // `off`: first free byte of the dynamic shared memory
    AZP_D void carve(unsigned int off)
        {
        const unsigned int warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, nwarps = blockDim.x >> 5;
        const unsigned int base = (off + 127u) & ~127u;
        slots = base + warp * WARP_BYTES + lane * PITCH;
        bars = base + nwarps * WARP_BYTES + warp * 8u * STAGES;
        }
    AZP_D unsigned int slot(unsigned int stage) const
        {
        return slots + stage * 32u * PITCH;
        }
    static AZP_D unsigned int shared_addr(unsigned int off)
        {
        return (unsigned int)__cvta_generic_to_shared(azp_smem + off);
        }
    AZP_D void init() const
        {
#if AZP_STAGE_LIST == 1
        if ((threadIdx.x & 31u) == 0u)
            for (unsigned int st = 0; st < STAGES; ++st)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(shared_addr(bars + 8u * st)) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
#endif
        }
    // copy vectors [first, first + count) of the lane's row (count <= CH) into its slot of `stage`
    AZP_D void issue(unsigned int stage, const uint4* src, unsigned int count) const
        {
#if AZP_STAGE_LIST == 1
        const unsigned int bar = shared_addr(bars + 8u * stage);
        if (count > 0u)
            {
            const unsigned int nbytes = 16u * count;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nbytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(shared_addr(slot(stage))),
                         "l"(src), "r"(nbytes), "r"(bar)
                         : "memory");
            }
        else
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
#else
        const unsigned int dst = shared_addr(slot(stage));
#pragma unroll
        for (unsigned int c = 0; c < CH; ++c)
            if (c < count)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * c), "l"(src + c) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
#endif
        }
    // chunk `k` has landed (chunks are issued in order, at most one younger chunk is in flight)
    AZP_D void wait(unsigned int k) const
        {
#if AZP_STAGE_LIST == 1
        const unsigned int bar = shared_addr(bars + 8u * (k % STAGES));
        const unsigned int parity = (k / STAGES) & 1u;
        asm volatile("{\n"
                     ".reg .pred p;\n"
                     "AZP_STAGE_WAIT:\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                     "@!p bra AZP_STAGE_WAIT;\n"
                     "}" ::"r"(bar),
                     "r"(parity)
                     : "memory");
#else
        asm volatile("cp.async.wait_group 1;" ::: "memory");
#endif
        }
    AZP_D uint4 read(unsigned int stage, unsigned int t) const
        {
        return *reinterpret_cast<const uint4*>(azp_smem + slot(stage) + 16u * t);
        }
    };

// How a family finds the constants of the (type_i, type_j) pair of a neighbour. NTM is the
// "type mode" template flag of the kernels:
//   1  single-type system: one set of constants, in registers for the whole kernel;
//   2  two types: the two candidate sets of the row (type_j = 0 / 1 given type_i) are loaded into
//      registers once per row and a neighbour picks one with a predicate select -- no shared
//      memory traffic in the neighbour loop (the L1/LSU data pipe is this kernel's scarce
//      resource);
//   0  general: tables in shared memory, indexed per neighbour with Index2D.
template<class E, class S, int NTM> struct TypeLookup
    {
    typedef typename E::cache_type Cache;
    PairTable<E, S> tab;
    Cache cA, cB;
    S rcA, rcB;
    unsigned int ti, ntypes;

    AZP_D void after_stage()
        {
        if (NTM == 1)
            {
            cA = tab.cache(0);
            rcA = tab.rcutsq(0);
            }
        }
    AZP_D void begin_row(unsigned int ti_, unsigned int ntypes_)
        {
        ti = ti_;
        ntypes = ntypes_;
        if (NTM == 2)
            {
            const unsigned int t = ti_ & 1u; // defensive: ids outside {0,1} cannot index past the table
            cA = tab.cache(index2d(2u, t, 0u));
            cB = tab.cache(index2d(2u, t, 1u));
            rcA = tab.rcutsq(index2d(2u, t, 0u));
            rcB = tab.rcutsq(index2d(2u, t, 1u));
            }
        }
    // true when the potential is switched off for every partner type of this row's particle
    // (all effective cutoffs are zero): the row then contributes nothing and is not streamed
    AZP_D bool row_disabled() const
        {
        if (NTM == 1)
            return !(rcA > S(0));
        if (NTM == 2)
            return !(rcA > S(0)) && !(rcB > S(0));
        bool off = true;
        for (unsigned int tj = 0; tj < ntypes; ++tj)
            off = off && !(tab.rcutsq(index2d(ntypes, ti, tj)) > S(0));
        return off;
        }
    AZP_D unsigned int pair_index(unsigned int tj) const
        {
        return NTM == 1 ? 0u : index2d(NTM == 2 ? 2u : ntypes, ti, tj);
        }
    AZP_D S rcutsq(unsigned int tj) const
        {
        if (NTM == 1)
            return rcA;
        if (NTM == 2)
            return tj ? rcB : rcA;
        return tab.rcutsq(index2d(ntypes, ti, tj));
        }
    // k * rcutsq with the product hoisted out of the neighbour loop for NTM = 1 / 2
    AZP_D S rcutsq_scaled(unsigned int tj, S k) const
        {
        if (NTM == 1)
            return rcA * k;
        if (NTM == 2)
            return tj ? rcB * k : rcA * k;
        return tab.rcutsq(index2d(ntypes, ti, tj)) * k;
        }
    AZP_D Cache cache(unsigned int tj) const
        {
        if (NTM == 1)
            return cA;
        if (NTM == 2)
            return select_words(tj != 0u, cA, cB);
        return tab.cache(index2d(ntypes, ti, tj));
        }
    };

// Rare-form deferral ("form split"). Some evaluators choose between arithmetically different
// FORMS by the type pair -- Colloid: point-point, sphere-point, sphere-sphere (reference
// src/PairEvaluatorColloid.h:233-269). In a two-type system a row sees two forms: the common one
// (solvent-solvent for a solvent row) and a rare, costlier one (the one or two colloids within
// reach). Evaluated in place, ONE lane with a rare neighbour makes its whole warp run the rare
// form as well -- for C3 that is 38 % of all neighbour slots. Instead the head of a rare
// neighbour only pushes its index into the lane's AcceptQueue, the in-place evaluation always
// uses the row's common form (which also drops the per-neighbour select of the constants), and
// the queued neighbours are evaluated when the row is done (or the queue is full): the lanes of
// a warp that hold rare entries run the rare form together, a few times per row instead of a
// few dozen. An evaluator opts in with `static constexpr bool kSplitForms = true`,
// `static int form(const cache_type&)` (lower = cheaper = evaluated in place), `kHeavyForm`
// (forms from this one on are ALWAYS deferred, so their code stays out of the neighbour loop and
// its register allocation) and `evalPairLight` (evalPair restricted to the forms below it).
template<class E, class = void> struct SplitForms : std::false_type
    {
    };
template<class E> struct SplitForms<E, typename std::enable_if<E::kSplitForms>::type> : std::true_type
    {
    };

template<class E, class S, bool ENABLED> struct FormSplit
    {
    AcceptQueue queue;
    bool defer_a, defer_b; // pairs with a neighbour of type 0 / 1 are deferred
    // Called after TypeLookup::begin_row has loaded the row's two candidate sets. A deferred
    // partner's in-place constants are replaced by the other partner's and its in-place cutoff
    // by zero, so the neighbour loop needs no extra test or select: a rare neighbour is simply
    // "outside"; drain_one() reloads the true constants from the shared-memory table.
    template<class TL> AZP_D void begin_row(TL& types)
        {
        queue.n = 0;
        const bool live_a = types.rcA > S(0), live_b = types.rcB > S(0);
        const int fa = E::form(types.cA), fb = E::form(types.cB);
        // defer the costlier form when both partners are live and their forms differ, and any
        // form the evaluator keeps out of the in-place code altogether (form >= kHeavyForm)
        defer_a = live_a && ((live_b && fa > fb) || fa >= E::kHeavyForm);
        defer_b = live_b && ((live_a && fb > fa) || fb >= E::kHeavyForm);
        if (defer_a)
            {
            types.cA = types.cB;
            types.rcA = S(0);
            }
        if (defer_b)
            {
            types.cB = types.cA;
            types.rcB = S(0);
            }
        }
    AZP_D bool rare(unsigned int tj) const
        {
        return tj ? defer_b : defer_a;
        }
    AZP_D bool any() const
        {
        return defer_a || defer_b;
        }
    };
template<class E, class S> struct FormSplit<E, S, false>
    {
    template<class TL> AZP_D void begin_row(TL&) { }
    AZP_D bool rare(unsigned int) const
        {
        return false;
        }
    AZP_D bool any() const
        {
        return false;
        }
    };

template<class E, class S> AZP_D void carve_split(FormSplit<E, S, true>& sp, unsigned int& off)
    {
    sp.queue.carve(off);
    off = sp.queue.off + (unsigned int)AcceptQueue::bytes(blockDim.x);
    }
template<class E, class S> AZP_D void carve_split(FormSplit<E, S, false>&, unsigned int&) { }
template<class E, class S> AZP_D void push_split(FormSplit<E, S, true>& sp, unsigned int j)
    {
    sp.queue.push(j);
    }
template<class E, class S> AZP_D void push_split(FormSplit<E, S, false>&, unsigned int) { }
template<class E, class S> AZP_D unsigned int pop_split(FormSplit<E, S, true>& sp)
    {
    return sp.queue.pop();
    }
template<class E, class S> AZP_D unsigned int pop_split(FormSplit<E, S, false>&)
    {
    return 0u;
    }
template<class E, class S> AZP_D bool split_pending(const FormSplit<E, S, true>& sp)
    {
    return sp.queue.n > 0u;
    }
template<class E, class S> AZP_D bool split_pending(const FormSplit<E, S, false>&)
    {
    return false;
    }
// fewer than `room` free slots left
template<class E, class S> AZP_D bool split_short(const FormSplit<E, S, true>& sp, unsigned int room)
    {
    return sp.queue.n + room > AcceptQueue::Q;
    }
template<class E, class S> AZP_D bool split_short(const FormSplit<E, S, false>&, unsigned int)
    {
    return false;
    }
// the in-place evaluation of a split family only ever sees the forms below E::kHeavyForm
template<bool LIGHT, class E, class S> AZP_D auto eval_in_place(E& eval, S& f, S& e) -> typename std::enable_if<LIGHT>::type
    {
    eval.evalPairLight(f, e);
    }
template<bool LIGHT, class E, class S> AZP_D auto eval_in_place(E& eval, S& f, S& e) -> typename std::enable_if<!LIGHT>::type
    {
    eval.evalPair(f, e, false);
    }

// =============================================================================================
// Isotropic family: F_i = sum dx * force_divr, E_i = 1/2 sum U, W_i = 1/2 sum dx_a dx_b force_divr
// =============================================================================================
// Per-evaluator tuning traits of the isotropic family (defaults; an evaluator header may
// specialise them): software-pipeline depth of the neighbour loop (2: index vector two trips and
// gathers one trip ahead; 0: load, gather, compute in program order with the smallest register
// footprint) and whether a two-type system keeps its candidate constants in registers.
template<class E> struct IsoTraits
    {
    static constexpr int pipe = 2;
    static constexpr bool register_tables = true;
    static constexpr int one_lane_cap = 0; // register cap of the one-lane main pass (OneLaneCap)
    };

template<class E_, class S_, bool XPLOR, bool VIRIAL, int NTM_> struct IsoFamily
    {
    typedef E_ E;
    typedef S_ S;
    typedef typename E::cache_type Cache;
    static constexpr int NTM = NTM_;
    static constexpr int PIPE = IsoTraits<E>::pipe;
    static constexpr bool QUEUE = false;
    static constexpr bool MULTIROW = false;
    // rare-form deferral (FormSplit): two-type systems of an evaluator with several forms
    static constexpr bool SPLIT = SplitForms<E>::value && NTM == 2 && !XPLOR;

    TypeLookup<E, S, NTM> types;
    FormSplit<E, S, SPLIT> split;
    unsigned int xplor_off;
    S fx = S(0), fy = S(0), fz = S(0), pe = S(0);
    Virial6<S> w;

    AZP_HD static size_t smem_bytes(size_t ntp, size_t block)
        {
        return PairTable<E, S>::bytes(ntp) + (XPLOR ? ntp * sizeof(XplorEntry<S>) : 0) + 32
               + (SPLIT ? AcceptQueue::bytes(block) : 0);
        }

    AZP_D XplorEntry<S>& xplor(unsigned int t) const
        {
        return reinterpret_cast<XplorEntry<S>*>(azp_smem + xplor_off)[t];
        }

    AZP_D void stage(const KernelArgs<S>& a, const typename E::param_type* params, unsigned int ntp)
        {
        PairTable<E, S>& tab = types.tab;
        tab.carve(ntp);
        xplor_off = (tab.end_off(ntp) + 15u) & ~15u;
        unsigned int split_off = xplor_off; // SPLIT excludes XPLOR: the queue takes the table's place
        carve_split(split, split_off);
        for (unsigned int t = threadIdx.x; t < ntp; t += blockDim.x)
            {
            const S rc = a.rcutsq[t];
            S ron = S(0);
            if (XPLOR)
                ron = a.ronsq[t];
            const bool energy_shift = (a.shift_mode == 1) || (XPLOR && ron > rc);
            const Cache c = E::make_cache(params[t], rc, energy_shift);
            tab.cache(t) = c;
            tab.rcutsq(t) = E::disabled(c) ? S(0) : rc;
            if (XPLOR)
                {
                const S d = rc - ron;
                xplor(t).ronsq = ron;
                xplor(t).denom_inv = S(1.0) / (d * d * d);
                }
            }
        __syncthreads();
        tab.finish(ntp);
        __syncthreads();
        types.after_stage();
        }

    AZP_D void begin_row(const KernelArgs<S>& a, unsigned int, unsigned int ti)
        {
        types.begin_row(ti, a.ntypes);
        split.begin_row(types);
        }

    // What is left of a neighbour once its gathered position has been consumed: the pipelined
    // loop computes the heads of a trip, re-issues the gathers into the same registers, then
    // runs the bodies (no register rotation between trips).
    struct Head
        {
        S dx, dy, dz;
        unsigned int tj;
        };
    AZP_D Head head(const KernelArgs<S>& a, const RowGeometry<S>& g, unsigned int j, const Vec4<S>& pj)
        {
        Head h;
        g.displacement(a.box, pj, h.dx, h.dy, h.dz);
        h.tj = NTM == 1 ? 0u : scalar_as_uint(pj.w);
        if (SPLIT && split.rare(h.tj))
            push_split(split, j); // evaluated by drain_one() once the row is done
        return h;
        }
    template<bool WRAP> AZP_D Head head_t(const KernelArgs<S>& a, const RowGeometry<S>& g, unsigned int j, const Vec4<S>& pj)
        {
        Head h;
        g.template displacement_t<WRAP>(a.box, pj, h.dx, h.dy, h.dz);
        h.tj = NTM == 1 ? 0u : scalar_as_uint(pj.w);
        if (SPLIT && split.rare(h.tj))
            push_split(split, j);
        return h;
        }
    AZP_D void pair(const KernelArgs<S>& a, const RowGeometry<S>& g, unsigned int j, const Vec4<S>& pj)
        {
        body(a, head(a, g, j, pj));
        if (SPLIT)
            drain_if_short(a, g, 1u);
        }
    // FormSplit: evaluate the queued rare-form neighbours of this lane. Lanes of a warp that hold
    // entries run the loop together; a lane's queue cannot overflow because process_row drains
    // whenever fewer than a trip's worth of slots is left.
    AZP_D void drain_if_short(const KernelArgs<S>& a, const RowGeometry<S>& g, unsigned int room)
        {
        if (split_short(split, room))
            drain(a, g);
        }
    AZP_D void drain(const KernelArgs<S>& a, const RowGeometry<S>& g)
        {
        while (split_pending(split))
            {
            const unsigned int j = pop_split(split);
            const Vec4<S> pj = load4(a.pos, j);
            Head h;
            g.displacement(a.box, pj, h.dx, h.dy, h.dz);
            h.tj = scalar_as_uint(pj.w);
            evaluate<false, true>(h);
            }
        }
    // a row whose live partners are all deferred still has to be streamed
    AZP_D bool row_off() const
        {
        return types.row_disabled() && !split.any();
        }
    AZP_D void body(const KernelArgs<S>&, const Head& h)
        {
        evaluate<SPLIT, false>(h);
        }
    // INPLACE_OF_SPLIT: the in-place evaluation of a family with rare-form deferral (rare
    // neighbours arrive with rcutsq = 0): only the forms below E::kHeavyForm are compiled in
    // TABLE: constants from the shared-memory table (the drain of a split row, whose register
    // copies hold the in-place constants) instead of the row's register copies
    template<bool INPLACE_OF_SPLIT, bool TABLE> AZP_D void evaluate(const Head& h)
        {
        const S dx = h.dx, dy = h.dy, dz = h.dz;
        const unsigned int tj = h.tj;
        const unsigned int t_tab = TABLE ? index2d(2u, types.ti & 1u, tj & 1u) : 0u;
        const S rsq = fma(dz, dz, fma(dy, dy, dx * dx));
        const S rcutsq = TABLE ? types.tab.rcutsq(t_tab) : types.rcutsq(tj);
        const bool inside = rsq < rcutsq;
        S force_divr = S(0), pair_eng = S(0);
        if (XPLOR)
            {
            // rare mode: keep the branch (the smoothing reads another table)
            if (inside)
                {
                const Cache c = types.cache(tj);
                E eval(rsq, rcutsq, c);
                eval.evalPair(force_divr, pair_eng, false);
                const XplorEntry<S> x = xplor(types.pair_index(tj));
                if (rsq >= x.ronsq)
                    {
                    const S m = rsq - rcutsq;
                    const S s = m * m * (rcutsq + S(2.0) * rsq - S(3.0) * x.ronsq) * x.denom_inv;
                    const S ds = S(12.0) * (rsq - x.ronsq) * m * x.denom_inv;
                    const S old_eng = pair_eng;
                    pair_eng = old_eng * s;
                    force_divr = s * force_divr - ds * old_eng;
                    }
                }
            }
        else
            {
            // Warp-level test instead of a per-lane branch. In a dense fluid some lane of the
            // warp is always inside the cutoff, so a divergent branch around the evaluator is
            // never skipped and only costs BSSY/BRA/BSYNC: when any lane needs the evaluator,
            // every lane runs it and rejected lanes are zeroed by a select (which also discards
            // any inf/NaN produced for an out-of-range rsq). When NO lane is inside -- potentials
            // that are switched off for most type pairs, e.g. Hertz acting only between colloids
            // -- the whole warp skips the evaluation and the accumulation with one uniform branch.
            // Evaluators opt in with kWarpVote (the vote costs two issue slots per neighbour,
            // which the cheap dense-fluid potentials, PLJ and Yukawa, do not get back).
            // (not in the in-place loop of a split row: its forms are the light ones and in a
            // dense fluid the vote never skips)
            if (E::kWarpVote && !INPLACE_OF_SPLIT && __ballot_sync(__activemask(), inside) == 0u)
                return;
            const Cache c = TABLE ? types.tab.cache(t_tab) : types.cache(tj);
            S f = S(0), e = S(0);
            E eval(rsq, rcutsq, c);
            eval_in_place<INPLACE_OF_SPLIT>(eval, f, e);
            force_divr = inside ? f : S(0);
            pair_eng = inside ? e : S(0);
            }
        fx = fma(dx, force_divr, fx);
        fy = fma(dy, force_divr, fy);
        fz = fma(dz, force_divr, fz);
        pe += pair_eng;
        if (VIRIAL)
            {
            const S vx = dx * force_divr, vy = dy * force_divr, vz = dz * force_divr;
            w.xx = fma(dx, vx, w.xx);
            w.xy = fma(dx, vy, w.xy);
            w.xz = fma(dx, vz, w.xz);
            w.yy = fma(dy, vy, w.yy);
            w.yz = fma(dy, vz, w.yz);
            w.zz = fma(dz, vz, w.zz);
            }
        }

    AZP_D void reset()
        {
        fx = fy = fz = pe = S(0);
        w = Virial6<S>();
        }

    AZP_D void finish(const KernelArgs<S>& a, unsigned int row, bool writer, unsigned int tpp)
        {
        for (unsigned int o = tpp >> 1; o > 0; o >>= 1)
            {
            fx += shfl_xor(fx, o);
            fy += shfl_xor(fy, o);
            fz += shfl_xor(fz, o);
            pe += shfl_xor(pe, o);
            if (VIRIAL)
                w.reduce(o);
            }
        if (writer)
            {
            store4(a.force, row, fx, fy, fz, S(0.5) * pe);
            if (VIRIAL)
                w.store(a.virial, a.virial_pitch, row);
            }
        }
    };

// =============================================================================================
// Fused two-potential isotropic family: two evaluators over ONE sweep of
// the neighbour list -- the reference's documented use is pair.Colloid plus pair.Hertz on the
// same nlist (src/pair.py:66-76), two ForceComputes that each stream the whole list. One gather,
// one displacement and one r^2 per neighbour feed both evaluators; each potential keeps its own
// type-pair table (own r_cut^2: a pair may be inside one cutoff and outside the other), its own
// accumulators and its own outputs, so the read-outs are those of two separate launches, bit for
// bit. Modes none / shift (xplor falls back to separate launches).
// =============================================================================================
template<class EA, class EB, class S_, bool VIRIAL, int NTM_> struct FusedIsoFamily
    {
    typedef EA E; // the params pointer of the launch is potential A's
    typedef S_ S;
    static constexpr int NTM = NTM_;
    static constexpr int PIPE = 2;
    static constexpr bool QUEUE = false;
    static constexpr bool MULTIROW = false;
    static constexpr bool SPLIT = false;

    TypeLookup<EA, S, NTM> types; // potential A (row_kernel reads rc_max through it: see stage)
    TypeLookup<EB, S, NTM> types_b;
    S fx = S(0), fy = S(0), fz = S(0), pe = S(0);
    S gx = S(0), gy = S(0), gz = S(0), qe = S(0);
    Virial6<S> w, w_b;

    AZP_HD static size_t smem_bytes(size_t ntp, size_t)
        {
        return PairTable<EA, S>::bytes(ntp) + 16 + PairTable<EB, S>::bytes(ntp) + 32;
        }

    AZP_D void stage(const KernelArgs<S>& a, const typename EA::param_type* params, unsigned int ntp)
        {
        PairTable<EA, S>& ta = types.tab;
        PairTable<EB, S>& tb = types_b.tab;
        ta.carve(ntp);
        tb.carve(ntp, (ta.end_off(ntp) + 15u) & ~15u);
        const typename EB::param_type* params_b = static_cast<const typename EB::param_type*>(a.params_b);
        for (unsigned int t = threadIdx.x; t < ntp; t += blockDim.x)
            {
            const S rca = a.rcutsq[t], rcb = a.rcutsq_b[t];
            const typename EA::cache_type ca = EA::make_cache(params[t], rca, a.shift_mode == 1);
            const typename EB::cache_type cb = EB::make_cache(params_b[t], rcb, a.shift_mode_b == 1);
            ta.cache(t) = ca;
            ta.rcutsq(t) = EA::disabled(ca) ? S(0) : rca;
            tb.cache(t) = cb;
            tb.rcutsq(t) = EB::disabled(cb) ? S(0) : rcb;
            }
        __syncthreads();
        ta.finish(ntp);
        tb.finish(ntp);
        __syncthreads();
        // the interior-warp test of process_row reads rc_max of table A: make it the larger one
        if (threadIdx.x == 0)
            ta.rcutsq(ntp) = fmax(ta.rcutsq(ntp), tb.rcutsq(ntp));
        __syncthreads();
        types.after_stage();
        types_b.after_stage();
        }

    AZP_D void begin_row(const KernelArgs<S>& a, unsigned int, unsigned int ti)
        {
        types.begin_row(ti, a.ntypes);
        types_b.begin_row(ti, a.ntypes);
        }
    // row_kernel skips a row only when BOTH potentials are switched off for its particle type
    struct BothOff
        {
        };

    struct Head
        {
        S dx, dy, dz;
        unsigned int tj;
        };
    AZP_D Head head(const KernelArgs<S>& a, const RowGeometry<S>& g, unsigned int, const Vec4<S>& pj) const
        {
        Head h;
        g.displacement(a.box, pj, h.dx, h.dy, h.dz);
        h.tj = NTM == 1 ? 0u : scalar_as_uint(pj.w);
        return h;
        }
    template<bool WRAP> AZP_D Head head_t(const KernelArgs<S>& a, const RowGeometry<S>& g, unsigned int, const Vec4<S>& pj) const
        {
        Head h;
        g.template displacement_t<WRAP>(a.box, pj, h.dx, h.dy, h.dz);
        h.tj = NTM == 1 ? 0u : scalar_as_uint(pj.w);
        return h;
        }
    AZP_D void pair(const KernelArgs<S>& a, const RowGeometry<S>& g, unsigned int j, const Vec4<S>& pj)
        {
        body(a, head(a, g, j, pj));
        }
    template<class EX, class TL>
    AZP_D void one(const TL& tl, S rsq, unsigned int tj, S dx, S dy, S dz, S& ax, S& ay, S& az, S& ae, Virial6<S>& aw)
        {
        const S rcutsq = tl.rcutsq(tj);
        const bool inside = rsq < rcutsq;
        if (EX::kWarpVote && __ballot_sync(__activemask(), inside) == 0u)
            return;
        const typename EX::cache_type c = tl.cache(tj);
        S f = S(0), e = S(0);
        EX eval(rsq, rcutsq, c);
        eval.evalPair(f, e, false);
        const S force_divr = inside ? f : S(0);
        const S pair_eng = inside ? e : S(0);
        ax = fma(dx, force_divr, ax);
        ay = fma(dy, force_divr, ay);
        az = fma(dz, force_divr, az);
        ae += pair_eng;
        if (VIRIAL)
            {
            const S vx = dx * force_divr, vy = dy * force_divr, vz = dz * force_divr;
            aw.xx = fma(dx, vx, aw.xx);
            aw.xy = fma(dx, vy, aw.xy);
            aw.xz = fma(dx, vz, aw.xz);
            aw.yy = fma(dy, vy, aw.yy);
            aw.yz = fma(dy, vz, aw.yz);
            aw.zz = fma(dz, vz, aw.zz);
            }
        }
    AZP_D void body(const KernelArgs<S>&, const Head& h)
        {
        const S rsq = fma(h.dz, h.dz, fma(h.dy, h.dy, h.dx * h.dx));
        one<EA>(types, rsq, h.tj, h.dx, h.dy, h.dz, fx, fy, fz, pe, w);
        one<EB>(types_b, rsq, h.tj, h.dx, h.dy, h.dz, gx, gy, gz, qe, w_b);
        }

    AZP_D void reset()
        {
        fx = fy = fz = pe = S(0);
        gx = gy = gz = qe = S(0);
        w = Virial6<S>();
        w_b = Virial6<S>();
        }

    AZP_D void finish(const KernelArgs<S>& a, unsigned int row, bool writer, unsigned int tpp)
        {
        for (unsigned int o = tpp >> 1; o > 0; o >>= 1)
            {
            fx += shfl_xor(fx, o);
            fy += shfl_xor(fy, o);
            fz += shfl_xor(fz, o);
            pe += shfl_xor(pe, o);
            gx += shfl_xor(gx, o);
            gy += shfl_xor(gy, o);
            gz += shfl_xor(gz, o);
            qe += shfl_xor(qe, o);
            if (VIRIAL)
                {
                w.reduce(o);
                w_b.reduce(o);
                }
            }
        if (writer)
            {
            store4(a.force, row, fx, fy, fz, S(0.5) * pe);
            store4(a.force_b, row, gx, gy, gz, S(0.5) * qe);
            if (VIRIAL)
                {
                w.store(a.virial, a.virial_pitch, row);
                w_b.store(a.virial_b, a.virial_pitch, row);
                }
            }
        }
    };

// =============================================================================================
// DPD thermostat family: also gathers vel_j and tag_j; force from force_divr (conservative +
// drag + random), virial from the conservative part only.
// =============================================================================================
template<class E_, class S_, bool VIRIAL, int NTM_> struct DpdFamily
    {
    typedef E_ E;
    typedef S_ S;
    typedef typename E::cache_type Cache;
    static constexpr int NTM = NTM_;
    static constexpr int PIPE = 0;
    static constexpr bool QUEUE = true;
    static constexpr bool MULTIROW = AZP_ROWS_PER_GROUP > 1;
    static constexpr bool SPLIT = false;

    TypeLookup<E, S, NTM> types;
    AcceptQueue queue;
    Vec4<S> vi;
    unsigned int tag_i;
    S fx = S(0), fy = S(0), fz = S(0), pe = S(0);
    Virial6<S> w;

    AZP_HD static size_t smem_bytes(size_t ntp, size_t block)
        {
        return PairTable<E, S>::bytes(ntp) + 16 + AcceptQueue::bytes(block);
        }

    AZP_D void stage(const KernelArgs<S>& a, const typename E::param_type* params, unsigned int ntp)
        {
        PairTable<E, S>& tab = types.tab;
        tab.carve(ntp);
        queue.carve(tab.end_off(ntp));
        for (unsigned int t = threadIdx.x; t < ntp; t += blockDim.x)
            {
            const S rc = a.rcutsq[t];
            tab.cache(t) = E::make_cache_thermo(params[t], rc, a.deltaT, a.T);
            tab.rcutsq(t) = rc;
            }
        __syncthreads();
        tab.finish(ntp);
        __syncthreads();
        types.after_stage();
        }

    AZP_D void begin_row(const KernelArgs<S>& a, unsigned int i, unsigned int ti)
        {
        types.begin_row(ti, a.ntypes);
        vi = load4(a.vel, i);
        tag_i = __ldg(a.tag + i);
        }

    template<class C>
    AZP_D void accept(const KernelArgs<S>& a, const C& c, const Vec4<S>& vj, unsigned int tag_j, S rsq, S rcutsq, S dx, S dy, S dz)
        {
        const S rdotv = dx * (vi.x - vj.x) + dy * (vi.y - vj.y) + dz * (vi.z - vj.z);
        S force_divr = S(0), force_divr_cons = S(0), pair_eng = S(0);
        E eval(rsq, rcutsq, c);
        eval.set_seed_ij_timestep((uint16_t)a.seed, tag_i, tag_j, a.timestep);
        eval.setDeltaT(a.deltaT);
        eval.setRDotV(rdotv);
        eval.setT(a.T);
        eval.evalThermoPair(force_divr, force_divr_cons, pair_eng, false);
        fx = fma(dx, force_divr, fx);
        fy = fma(dy, force_divr, fy);
        fz = fma(dz, force_divr, fz);
        pe += pair_eng;
        if (VIRIAL)
            {
            const S vx = dx * force_divr_cons, vy = dy * force_divr_cons, vz = dz * force_divr_cons;
            w.xx = fma(dx, vx, w.xx);
            w.xy = fma(dx, vy, w.xy);
            w.xz = fma(dx, vz, w.xz);
            w.yy = fma(dy, vy, w.yy);
            w.yz = fma(dy, vz, w.yz);
            w.zz = fma(dz, vz, w.zz);
            }
        }

    // cheap part, every entry: minimum image + cutoff test; accepted entries are queued
    AZP_D void scan(const KernelArgs<S>& a, const RowGeometry<S>& g, unsigned int j, const Vec4<S>& pj)
        {
        S dx, dy, dz;
        g.displacement(a.box, pj, dx, dy, dz);
        const S rsq = fma(dz, dz, fma(dy, dy, dx * dx));
        const unsigned int tj = NTM == 1 ? 0u : scalar_as_uint(pj.w);
        // the scan's rsq is accumulated with FMAs; the decisive test (heavy) uses the
        // reference's rounding of rsq, so the scan admits a margin of a few ulp
        if (rsq < types.rcutsq_scaled(tj, S(1.0) + S(8) * (sizeof(S) == 4 ? S(1.1920929e-7) : S(2.220446049250313e-16))))
            {
            queue.push(j);
#ifdef AZP_DPD_PREFETCH
            // the heavy round gathers vel[j] and tag[j] behind the pop: start them now (no register)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(a.vel + 4 * (size_t)j));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(a.tag + j));
#endif
            }
        }

    // heavy part: up to AZP_HEAVY_UNROLL queued neighbours per lane and round (lanes with an
    // empty queue idle). Velocity and tag are only needed for accepted pairs (about a third of
    // the list at buffer 0.4), so they are gathered here, behind the scan's cutoff test --
    // together with the position, and for all neighbours of the round before any of them is
    // evaluated, so that a round exposes ONE gather latency (the few entries inside the scan's
    // ulp margin but outside the cutoff load them in vain).
    AZP_D void heavy_one(const KernelArgs<S>& a, const RowGeometry<S>& g, const Vec4<S>& pj, const Vec4<S>& vj, unsigned int tag_j)
        {
        S dx, dy, dz;
        g.displacement(a.box, pj, dx, dy, dz);
        // rsq = dot(dx, dx) rounded like the reference's host loop (no FMA): for s < 2 the
        // weight is not continuous in the last ulp below the cutoff (eval_dpd.cuh), so
        // both the cutoff decision and r must see the reference's rsq
        const S rsq = ref::dot3(dx, dy, dz, dx, dy, dz);
        const unsigned int tj = NTM == 1 ? 0u : scalar_as_uint(pj.w);
        const S rcutsq = types.rcutsq(tj);
        if (rsq < rcutsq)
            {
            const Cache c = types.cache(tj);
            accept(a, c, vj, tag_j, rsq, rcutsq, dx, dy, dz);
            }
        }
    AZP_D void heavy(const KernelArgs<S>& a, const RowGeometry<S>& g)
        {
#ifdef AZP_DIAG_NO_HEAVY
        // timing diagnosis only (wrong forces): what does the scan + list stream cost alone?
        if (queue.n > 0u)
            fx += __uint_as_float(queue.pop());
        return;
#endif
        const unsigned int pending = queue.n;
        if (pending > 0u)
            {
            const unsigned int j0 = queue.pop();
            const bool two = AZP_HEAVY_UNROLL > 1 && pending > 1u;
            const unsigned int j1 = two ? queue.pop() : j0;
            const Vec4<S> p0 = load4(a.pos, j0);
            const Vec4<S> v0 = load4(a.vel, j0);
            const unsigned int t0 = __ldg(a.tag + j0);
            if (AZP_HEAVY_UNROLL > 1)
                {
                const Vec4<S> p1 = load4(a.pos, j1);
                const Vec4<S> v1 = load4(a.vel, j1);
                const unsigned int t1 = __ldg(a.tag + j1);
                heavy_one(a, g, p0, v0, t0);
                if (two)
                    heavy_one(a, g, p1, v1, t1);
                }
            else
                heavy_one(a, g, p0, v0, t0);
            }
        }

    AZP_D void reset()
        {
        fx = fy = fz = pe = S(0);
        w = Virial6<S>();
        }

    AZP_D void finish(const KernelArgs<S>& a, unsigned int row, bool writer, unsigned int tpp)
        {
        for (unsigned int o = tpp >> 1; o > 0; o >>= 1)
            {
            fx += shfl_xor(fx, o);
            fy += shfl_xor(fy, o);
            fz += shfl_xor(fz, o);
            pe += shfl_xor(pe, o);
            if (VIRIAL)
                w.reduce(o);
            }
        if (writer)
            {
            store4(a.force, row, fx, fy, fz, S(0.5) * pe);
            if (VIRIAL)
                w.store(a.virial, a.virial_pitch, row);
            }
        }
    };

// =============================================================================================
// Anisotropic family: gathers orientation_j; vector force, torque on i, energy,
// virial 1/2 dx_a F_b.
// =============================================================================================
template<class E_, class S_, bool VIRIAL, int NTM_> struct AnisoFamily
    {
    typedef E_ E;
    typedef S_ S;
    typedef typename E::cache_type Cache;
    static constexpr int NTM = NTM_;
    // measured on C5: the software pipeline of the isotropic family (PIPE = 2) needs 99
    // registers here and is 50 % slower (0.69 vs 0.46 ms per 4 M particles) than hiding the
    // latency with occupancy (58 registers, 8 CTAs per SM)
    static constexpr int PIPE = 0;
    // Deferred accept (AcceptQueue) for the two-patch Morse family: off. Round 1 measured it
    // 6 % slower than evaluating in place (-11 % instructions, but one exposed gather latency
    // per heavy round). Round 2, with all gathers of a round issued together and two queued
    // neighbours per lane and round (AZP_ANISO_HEAVY_UNROLL): 0.454 -> 0.439 ms at 4 M
    // particles but 1.746 -> 1.733 ms (0.7 %) at the full 16 M, and a row's sum then depends on
    // when its warp drains (row ranges are no longer bit-identical to the full launch), so the
    // in-place evaluation stays.
#ifndef AZP_ANISO_QUEUE
#define AZP_ANISO_QUEUE 0
#endif
    static constexpr bool QUEUE = AZP_ANISO_QUEUE != 0;
    static constexpr bool MULTIROW = AZP_ROWS_PER_GROUP > 1;
    static constexpr bool SPLIT = false;

    TypeLookup<E, S, NTM> types;
    AcceptQueue queue;
    typename E::row_type ri; // what the evaluator keeps of orientation_i (two-patch Morse: the director)
    S fx = S(0), fy = S(0), fz = S(0), pe = S(0);
    S tx = S(0), ty = S(0), tz = S(0);
    Virial6<S> w;

    AZP_HD static size_t smem_bytes(size_t ntp, size_t block)
        {
        return PairTable<E, S>::bytes(ntp) + 16 + (QUEUE ? AcceptQueue::bytes(block) : 0);
        }

    AZP_D void stage(const KernelArgs<S>& a, const typename E::param_type* params, unsigned int ntp)
        {
        PairTable<E, S>& tab = types.tab;
        tab.carve(ntp);
        queue.carve(tab.end_off(ntp));
        for (unsigned int t = threadIdx.x; t < ntp; t += blockDim.x)
            {
            const S rc = a.rcutsq[t];
            tab.cache(t) = E::make_cache(params[t], rc, a.shift_mode == 1);
            tab.rcutsq(t) = rc;
            }
        __syncthreads();
        tab.finish(ntp);
        __syncthreads();
        types.after_stage();
        }

    AZP_D void begin_row(const KernelArgs<S>& a, unsigned int i, unsigned int ti)
        {
        types.begin_row(ti, a.ntypes);
        ri = E::make_row(load4(a.orientation, i));
        }

    template<class C>
    AZP_D void accept(const KernelArgs<S>& a, const C& c, unsigned int j, S rsq, S rcutsq, const Vec3<S>& dr)
        {
        // the orientation gather and the evaluator body are skipped for the (about half)
        // rejected entries
        accept(c, load4(a.orientation, j), rsq, rcutsq, dr);
        }
    template<class C> AZP_D void accept(const C& c, const Vec4<S>& qj, S rsq, S rcutsq, const Vec3<S>& dr)
        {
        Vec3<S> force {S(0), S(0), S(0)}, torque_i {S(0), S(0), S(0)}, torque_j {S(0), S(0), S(0)};
        S pair_eng = S(0);
        E eval(typename E::FromRow(), dr, ri, qj, rcutsq, c);
        eval.evaluatePair(rsq, force, pair_eng, false, torque_i, torque_j);
        fx += force.x;
        fy += force.y;
        fz += force.z;
        tx += torque_i.x;
        ty += torque_i.y;
        tz += torque_i.z;
        pe += pair_eng;
        if (VIRIAL)
            {
            w.xx = fma(dr.x, force.x, w.xx);
            w.xy = fma(dr.y, force.x, w.xy);
            w.xz = fma(dr.z, force.x, w.xz);
            w.yy = fma(dr.y, force.y, w.yy);
            w.yz = fma(dr.z, force.y, w.yz);
            w.zz = fma(dr.z, force.z, w.zz);
            }
        }

    AZP_D void pair(const KernelArgs<S>& a, const RowGeometry<S>& g, unsigned int j, const Vec4<S>& pj)
        {
        Vec3<S> dr;
        g.displacement(a.box, pj, dr.x, dr.y, dr.z);
        const S rsq = fma(dr.z, dr.z, fma(dr.y, dr.y, dr.x * dr.x));
        // the reference evaluator rejects only rsq > rcutsq (strictly), so accept rsq <= rcutsq
        const unsigned int tj = NTM == 1 ? 0u : scalar_as_uint(pj.w);
        const S rcutsq = types.rcutsq(tj);
        if (rsq <= rcutsq)
            {
#ifdef AZP_DIAG_NO_HEAVY
            fx += dr.x; // timing diagnosis only (wrong forces)
            return;
#endif
            const Cache c = types.cache(tj);
            accept(a, c, j, rsq, rcutsq, dr);
            }
        }

    // deferred-accept variant (AZP_ANISO_QUEUE): see AcceptQueue
    AZP_D void scan(const KernelArgs<S>& a, const RowGeometry<S>& g, unsigned int j, const Vec4<S>& pj)
        {
        Vec3<S> dr;
        g.displacement(a.box, pj, dr.x, dr.y, dr.z);
        const S rsq = fma(dr.z, dr.z, fma(dr.y, dr.y, dr.x * dr.x));
        const unsigned int tj = NTM == 1 ? 0u : scalar_as_uint(pj.w);
        if (rsq <= types.rcutsq(tj))
            {
            queue.push(j);
#ifdef AZP_ANISO_PREFETCH
            asm volatile("prefetch.global.L1 [%0];" ::"l"(a.orientation + 4 * (size_t)j));
#endif
            }
        }
    AZP_D void heavy_one(const KernelArgs<S>& a, const RowGeometry<S>& g, const Vec4<S>& pj, const Vec4<S>& qj)
        {
        Vec3<S> dr;
        g.displacement(a.box, pj, dr.x, dr.y, dr.z);
        const S rsq = fma(dr.z, dr.z, fma(dr.y, dr.y, dr.x * dr.x));
        const unsigned int tj = NTM == 1 ? 0u : scalar_as_uint(pj.w);
        accept(types.cache(tj), qj, rsq, types.rcutsq(tj), dr);
        }
    AZP_D void heavy(const KernelArgs<S>& a, const RowGeometry<S>& g)
        {
        // all gathers of the round (position + orientation of up to AZP_ANISO_HEAVY_UNROLL queued
        // neighbours) are issued together: one exposed latency per round; the scan applied
        // pair()'s own cutoff test to the same displacement, so it is not repeated
        const unsigned int pending = queue.n;
        if (pending > 0u)
            {
            const unsigned int j0 = queue.pop();
            const bool two = AZP_ANISO_HEAVY_UNROLL > 1 && pending > 1u;
            const unsigned int j1 = two ? queue.pop() : j0;
            const Vec4<S> p0 = load4(a.pos, j0);
            const Vec4<S> q0 = load4(a.orientation, j0);
            if (AZP_ANISO_HEAVY_UNROLL > 1)
                {
                const Vec4<S> p1 = load4(a.pos, j1);
                const Vec4<S> q1 = load4(a.orientation, j1);
                heavy_one(a, g, p0, q0);
                if (two)
                    heavy_one(a, g, p1, q1);
                }
            else
                heavy_one(a, g, p0, q0);
            }
        }

    AZP_D void reset()
        {
        fx = fy = fz = pe = S(0);
        tx = ty = tz = S(0);
        w = Virial6<S>();
        }

    AZP_D void finish(const KernelArgs<S>& a, unsigned int row, bool writer, unsigned int tpp)
        {
        for (unsigned int o = tpp >> 1; o > 0; o >>= 1)
            {
            fx += shfl_xor(fx, o);
            fy += shfl_xor(fy, o);
            fz += shfl_xor(fz, o);
            pe += shfl_xor(pe, o);
            tx += shfl_xor(tx, o);
            ty += shfl_xor(ty, o);
            tz += shfl_xor(tz, o);
            if (VIRIAL)
                w.reduce(o);
            }
        if (writer)
            {
            store4(a.force, row, fx, fy, fz, S(0.5) * pe);
            store4(a.torque, row, tx, ty, tz, S(0));
            if (VIRIAL)
                w.store(a.virial, a.virial_pitch, row);
            }
        }
    };

// =============================================================================================
// The kernel skeleton
// =============================================================================================
// Compile-time dispatch between the two family interfaces (pair: evaluate in place; scan/heavy:
// deferred accept). A family defines only its own; the unused branch is never instantiated.
template<class Fam, class S>
AZP_D auto pair_dispatch(Fam& fam, const KernelArgs<S>& a, const RowGeometry<S>& g, unsigned int j, const Vec4<S>& pj)
    -> typename std::enable_if<!Fam::QUEUE>::type
    {
    fam.pair(a, g, j, pj);
    }
template<class Fam, class S>
AZP_D auto pair_dispatch(Fam&, const KernelArgs<S>&, const RowGeometry<S>&, unsigned int, const Vec4<S>&)
    -> typename std::enable_if<Fam::QUEUE>::type
    {
    }
struct NoHead
    {
    };
// a row is not streamed when its particle type has every partner switched off -- in BOTH
// potentials for a fused family (detected by its BothOff tag)
template<class Fam> AZP_D auto row_disabled_dispatch(const Fam& fam) -> decltype(typename Fam::BothOff(), bool())
    {
    return fam.types.row_disabled() && fam.types_b.row_disabled();
    }
template<class Fam> AZP_D auto row_disabled_dispatch(const Fam& fam) -> decltype(fam.row_off())
    {
    return fam.row_off();
    }
template<class Fam, class... X> AZP_D bool row_disabled_dispatch(const Fam& fam, X...)
    {
    return fam.types.row_disabled();
    }
template<class Fam, class S>
AZP_D auto head_dispatch(Fam& fam, const KernelArgs<S>& a, const RowGeometry<S>& g, unsigned int j, const Vec4<S>& pj)
    -> typename std::enable_if<Fam::PIPE == 2, typename Fam::Head>::type
    {
    return fam.head(a, g, j, pj);
    }
template<class Fam, class S>
AZP_D auto head_dispatch(Fam&, const KernelArgs<S>&, const RowGeometry<S>&, unsigned int, const Vec4<S>&)
    -> typename std::enable_if<Fam::PIPE != 2, NoHead>::type
    {
    return NoHead();
    }
template<bool WRAP, class Fam, class S, class H>
AZP_D auto heads_dispatch(Fam& fam, const KernelArgs<S>& a, const RowGeometry<S>& g, const uint4& j, const Vec4<S>& p0, const Vec4<S>& p1,
                          const Vec4<S>& p2, const Vec4<S>& p3, H& h0, H& h1, H& h2, H& h3) -> typename std::enable_if<Fam::PIPE == 2>::type
    {
    h0 = fam.template head_t<WRAP>(a, g, j.x, p0);
    h1 = fam.template head_t<WRAP>(a, g, j.y, p1);
    h2 = fam.template head_t<WRAP>(a, g, j.z, p2);
    h3 = fam.template head_t<WRAP>(a, g, j.w, p3);
    }
template<bool WRAP, class Fam, class S, class H>
AZP_D auto heads_dispatch(Fam&, const KernelArgs<S>&, const RowGeometry<S>&, const uint4&, const Vec4<S>&, const Vec4<S>&, const Vec4<S>&,
                          const Vec4<S>&, H&, H&, H&, H&) -> typename std::enable_if<Fam::PIPE != 2>::type
    {
    }
template<class Fam, class S, class H>
AZP_D auto body_dispatch(Fam& fam, const KernelArgs<S>& a, const H& h) -> typename std::enable_if<Fam::PIPE == 2>::type
    {
    fam.body(a, h);
    }
template<class Fam, class S, class H>
AZP_D auto body_dispatch(Fam&, const KernelArgs<S>&, const H&) -> typename std::enable_if<Fam::PIPE != 2>::type
    {
    }
template<class Fam, class S>
AZP_D auto scan_dispatch(Fam& fam, const KernelArgs<S>& a, const RowGeometry<S>& g, unsigned int j, const Vec4<S>& pj)
    -> typename std::enable_if<Fam::QUEUE>::type
    {
    fam.scan(a, g, j, pj);
    }
template<class Fam, class S>
AZP_D auto scan_dispatch(Fam&, const KernelArgs<S>&, const RowGeometry<S>&, unsigned int, const Vec4<S>&)
    -> typename std::enable_if<!Fam::QUEUE>::type
    {
    }
template<class Fam, class S>
AZP_D auto heavy_dispatch(Fam& fam, const KernelArgs<S>& a, const RowGeometry<S>& g) -> typename std::enable_if<Fam::QUEUE>::type
    {
    fam.heavy(a, g);
    }
template<class Fam, class S>
AZP_D auto heavy_dispatch(Fam&, const KernelArgs<S>&, const RowGeometry<S>&) -> typename std::enable_if<!Fam::QUEUE>::type
    {
    }
template<class Fam> AZP_D auto queue_needs_drain(const Fam& fam) -> typename std::enable_if<Fam::QUEUE, bool>::type
    {
    return fam.queue.needs_drain();
    }
template<class Fam> AZP_D auto queue_needs_drain(const Fam&) -> typename std::enable_if<!Fam::QUEUE, bool>::type
    {
    return false;
    }
template<class Fam> AZP_D auto queue_pending(const Fam& fam) -> typename std::enable_if<Fam::QUEUE, bool>::type
    {
    return fam.queue.n > 0u;
    }
template<class Fam> AZP_D auto queue_pending(const Fam&) -> typename std::enable_if<!Fam::QUEUE, bool>::type
    {
    return false;
    }

// rare-form deferral (FormSplit) of the isotropic family: drain when a trip could overflow the
// lane's queue, and when the row is done
template<class Fam, class S>
AZP_D auto split_trip_dispatch(Fam& fam, const KernelArgs<S>& a, const RowGeometry<S>& g) -> typename std::enable_if<Fam::SPLIT>::type
    {
    fam.drain_if_short(a, g, AcceptQueue::ROOM);
    }
template<class Fam, class S>
AZP_D auto split_trip_dispatch(Fam&, const KernelArgs<S>&, const RowGeometry<S>&) -> typename std::enable_if<!Fam::SPLIT>::type
    {
    }
template<class Fam, class S>
AZP_D auto split_finish_dispatch(Fam& fam, const KernelArgs<S>& a, const RowGeometry<S>& g) -> typename std::enable_if<Fam::SPLIT>::type
    {
    fam.drain(a, g);
    }
template<class Fam, class S>
AZP_D auto split_finish_dispatch(Fam&, const KernelArgs<S>&, const RowGeometry<S>&) -> typename std::enable_if<!Fam::SPLIT>::type
    {
    }

// One row for one group of `tpp` lanes: geometry, neighbour stream, reduction, store.
template<class Fam, bool STAGED = false>
AZP_D void process_row(Fam& fam,
                       const KernelArgs<typename Fam::S>& a,
                       const unsigned int ntp,
                       const unsigned int row,
                       unsigned int n,
                       const uint64_t head,
                       const bool active,
                       const unsigned int lane,
                       const unsigned int tpp,
                       const ListStage* ls = nullptr)
    {
    typedef typename Fam::S S;
    const unsigned int i = active ? row + a.row_offset : 0u;

    RowGeometry<S> g;
    g.pi = load4(a.pos, i);
    g.ti = scalar_as_uint(g.pi.w);
    g.Lx = a.box.L[0], g.Ly = a.box.L[1], g.Lz = a.box.L[2];
    g.ix = a.box.Linv[0], g.iy = a.box.Linv[1], g.iz = a.box.Linv[2];
        {
        // the warp may skip the minimum-image wrap when the box is orthorhombic and fully
        // periodic and every active row of the warp is farther than rc_max from all faces
        const S rc_max = fam.types.tab.rcutsq(ntp);
        const S m = S(0.49999);
        const bool inside = (fabs(g.pi.x) + rc_max < m * g.Lx) && (fabs(g.pi.y) + rc_max < m * g.Ly)
                            && (fabs(g.pi.z) + rc_max < m * g.Lz);
        g.skip_wrap = (a.box.flags == 2) && __all_sync(0xffffffffu, inside || !active);
        }
    fam.begin_row(a, i, g.ti);
    if (row_disabled_dispatch(fam))
        n = 0; // e.g. Hertz acting only between colloids: solvent rows just write zeros

    // ---- the row as aligned uint4 vectors of neighbour indices ------------------------------
    // `pre` = entries between the 16-byte boundary below the row start and the row start; the
    // valid entries of `base` are [pre, end). Full vectors are v in [v_begin, v_end); the (at
    // most 3 + 3) entries in front of / behind them are handled by a guarded scalar epilogue, so
    // rows may start at any head_list offset and nothing outside the row is ever read.
    const unsigned int* rowp = a.nlist + head;
    const unsigned int pre = (unsigned int)((reinterpret_cast<uintptr_t>(rowp) >> 2) & 3u);
    const unsigned int* base = rowp - pre;
    const unsigned int end = pre + n;
    const uint4* base4 = reinterpret_cast<const uint4*>(base);
    const unsigned int v_begin = (pre + 3u) >> 2;
    const unsigned int v_end = end >> 2;

    // Software pipeline. PIPE = 2: the index vector is loaded two trips ahead (it streams from
    // HBM: ~1 us) and the four position gathers one trip ahead (L1/L2), so a lane always has one
    // nlist load and four gathers in flight while it does the math of the current vector --
    // best for the cheap isotropic evaluators. A trip first consumes the four gathered positions
    // (heads: displacement + type, 4 registers each), re-issues the gathers into the same
    // registers, then runs the four bodies as one branch-free block: no register rotation
    // (15 MOVs per trip less than copying the positions aside; 0.292 -> 0.280 ms on C2). PIPE = 0: load, gather, compute in program order
    // with the smallest register footprint -- best for the heavy DPD / anisotropic evaluators,
    // which hide latency with occupancy instead (measured).
    unsigned int v = v_begin + lane;
    if (Fam::PIPE == 2 && STAGED)
        {
        // One lane per row, neighbour list staged through shared memory (ListStage). The loop is
        // warp-uniform: it runs for the longest row of the warp, lanes whose row is done idle.
        // It walks the chunks of CH index vectors; the CH trips of a chunk are unrolled, so the
        // ring stage and the slot offsets are compile-time constants inside a trip. Trip t
        // consumes the positions gathered during the previous trip, reads the indices of the
        // next trip from the ring and gathers, then runs the bodies. The last trip of a chunk
        // waits for the next chunk (issued two chunks = 2 CH trips earlier) and refills the
        // stage it has just finished with the chunk after that.
        constexpr unsigned int CH = ListStage::CH;
        const unsigned int full = 0xffffffffu;
        unsigned int rem = v_begin < v_end ? v_end - v_begin : 0u; // vectors of this lane not yet consumed
        const unsigned int nchunks = (__reduce_max_sync(full, rem) + CH - 1u) / CH;
        const uint4* src = base4 + v_begin; // first vector of the chunk being consumed
        uint4 j_cur = make_uint4(0u, 0u, 0u, 0u);
        Vec4<S> p0, p1, p2, p3;
        if (nchunks > 0u)
            {
            ls->issue(0u, src, min(rem, CH));
            ls->issue(1u, src + CH, rem > CH ? min(rem - CH, CH) : 0u);
            ls->wait(0u);
            if (rem > 0u)
                {
                j_cur = ls->read(0u, 0u);
                p0 = load4(a.pos, j_cur.x);
                p1 = load4(a.pos, j_cur.y);
                p2 = load4(a.pos, j_cur.z);
                p3 = load4(a.pos, j_cur.w);
                }
            }
        for (unsigned int k = 0; k < nchunks; ++k)
            {
            const unsigned int stage = k & 1u;
#pragma unroll
            for (unsigned int t = 0; t < CH; ++t)
                {
                const bool act = t < rem;
                decltype(head_dispatch(fam, a, g, 0u, p0)) h0, h1, h2, h3;
                if (act)
                    {
                    h0 = head_dispatch(fam, a, g, j_cur.x, p0);
                    h1 = head_dispatch(fam, a, g, j_cur.y, p1);
                    h2 = head_dispatch(fam, a, g, j_cur.z, p2);
                    h3 = head_dispatch(fam, a, g, j_cur.w, p3);
                    }
                if (t + 1u < CH)
                    {
                    if (t + 1u < rem)
                        {
                        j_cur = ls->read(stage, t + 1u);
                        p0 = load4(a.pos, j_cur.x);
                        p1 = load4(a.pos, j_cur.y);
                        p2 = load4(a.pos, j_cur.z);
                        p3 = load4(a.pos, j_cur.w);
                        }
                    }
                else
                    {
                    const bool more = k + 1u < nchunks; // warp-uniform
                    if (more)
                        ls->wait(k + 1u);
                    if (rem > CH)
                        {
                        j_cur = ls->read(stage ^ 1u, 0u);
                        p0 = load4(a.pos, j_cur.x);
                        p1 = load4(a.pos, j_cur.y);
                        p2 = load4(a.pos, j_cur.z);
                        p3 = load4(a.pos, j_cur.w);
                        }
                    if (more)
                        {
                        __syncwarp();
                        ls->issue(stage, src + 2u * CH, rem > 2u * CH ? min(rem - 2u * CH, CH) : 0u);
                        }
                    }
                if (act)
                    {
                    body_dispatch(fam, a, h0);
                    body_dispatch(fam, a, h1);
                    body_dispatch(fam, a, h2);
                    body_dispatch(fam, a, h3);
                    split_trip_dispatch(fam, a, g);
                    }
                }
            rem = rem > CH ? rem - CH : 0u;
            src += CH;
            }
#if AZP_STAGE_LIST == 2
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
        }
    else if (Fam::PIPE == 2)
        {
        // pointer + countdown form: the loop carries one 64-bit cursor and one trip counter
        // (instead of v, v_end and the lane stride, which ptxas re-derived every trip)
        const uint4* cur4 = base4 + v;
        unsigned int trips = v < v_end ? (v_end - v + tpp - 1u) / tpp : 0u;
        uint4 j_cur = make_uint4(0u, 0u, 0u, 0u), j_nxt = make_uint4(0u, 0u, 0u, 0u);
        Vec4<S> p0, p1, p2, p3;
        if (trips > 0u)
            {
            j_cur = load_index4(cur4);
            p0 = load4(a.pos, j_cur.x);
            p1 = load4(a.pos, j_cur.y);
            p2 = load4(a.pos, j_cur.z);
            p3 = load4(a.pos, j_cur.w);
            if (trips > 1u)
                j_nxt = load_index4(cur4 + tpp);
            }
        while (trips > 0u)
            {
#if AZP_TRIP_WRAP
            // one (warp-uniform) skip_wrap branch per trip instead of one per neighbour
            decltype(head_dispatch(fam, a, g, 0u, p0)) h0, h1, h2, h3;
            if (g.skip_wrap)
                heads_dispatch<false>(fam, a, g, j_cur, p0, p1, p2, p3, h0, h1, h2, h3);
            else
                heads_dispatch<true>(fam, a, g, j_cur, p0, p1, p2, p3, h0, h1, h2, h3);
#else
            const auto h0 = head_dispatch(fam, a, g, j_cur.x, p0);
            const auto h1 = head_dispatch(fam, a, g, j_cur.y, p1);
            const auto h2 = head_dispatch(fam, a, g, j_cur.z, p2);
            const auto h3 = head_dispatch(fam, a, g, j_cur.w, p3);
#endif
            if (trips > 1u)
                {
                j_cur = j_nxt;
                p0 = load4(a.pos, j_cur.x);
                p1 = load4(a.pos, j_cur.y);
                p2 = load4(a.pos, j_cur.z);
                p3 = load4(a.pos, j_cur.w);
                if (trips > 2u)
                    j_nxt = load_index4(cur4 + 2u * tpp);
                }
            body_dispatch(fam, a, h0);
            body_dispatch(fam, a, h1);
            body_dispatch(fam, a, h2);
            body_dispatch(fam, a, h3);
            split_trip_dispatch(fam, a, g);
#if AZP_NLIST_LINE_PREFETCH
            // at the start of a 128-byte line of the row: request the line AZP_NLIST_LINE_PREFETCH
            // lines ahead (one lane-private request per 8 trips)
            if ((reinterpret_cast<uintptr_t>(cur4) & 127u) == 0u && trips > 8u * AZP_NLIST_LINE_PREFETCH)
                prefetch_l2(cur4 + 8u * AZP_NLIST_LINE_PREFETCH);
#endif
            cur4 += tpp;
            --trips;
            }
        }
    else if (!Fam::QUEUE)
        {
        for (; v < v_end; v += tpp)
            {
            const uint4 j = load_index4(base4 + v);
            const Vec4<S> q0 = load4(a.pos, j.x);
            const Vec4<S> q1 = load4(a.pos, j.y);
            const Vec4<S> q2 = load4(a.pos, j.z);
            const Vec4<S> q3 = load4(a.pos, j.w);
            pair_dispatch(fam, a, g, j.x, q0);
            pair_dispatch(fam, a, g, j.y, q1);
            pair_dispatch(fam, a, g, j.z, q2);
            pair_dispatch(fam, a, g, j.w, q3);
            split_trip_dispatch(fam, a, g);
            }
        }

    // leftovers in front of and behind the full vectors
    const unsigned int head_end = min(4u * v_begin, end);              // [pre, head_end)
    const unsigned int tail_begin = max(4u * v_end, head_end);         // [tail_begin, end)
    const unsigned int n_left = (head_end - pre) + (end - tail_begin); // <= 6
    if (!Fam::QUEUE)
        {
        for (unsigned int q = lane; q < n_left; q += tpp)
            {
            const unsigned int idx = q < head_end - pre ? pre + q : tail_begin + (q - (head_end - pre));
            const unsigned int j = __ldg(base + idx);
            const Vec4<S> pj = load4(a.pos, j);
            pair_dispatch(fam, a, g, j, pj);
            }
        }
    else
        {
        // Deferred accept (see AcceptQueue): the loops are warp-uniform (every lane of the warp
        // is here: row_kernel calls process_row unconditionally), the scan is per lane, and a
        // warp-wide heavy round runs whenever some lane could not take another trip.
        const unsigned int full = 0xffffffffu;
        bool more = v < v_end;
        // the index vector (the HBM stream: the longest latency of a trip) is loaded one trip
        // ahead; 4 registers, C4 0.378 -> 0.360 ms per 2 M particles. (The same prefetch in the
        // in-place loop of the anisotropic family makes ptxas use 83 registers instead of 58
        // and is 20 % slower, so that loop loads its indices just in time; prefetching the
        // position gathers across the heavy rounds as well costs 96 registers and 24 %.)
        uint4 j_nxt = make_uint4(0u, 0u, 0u, 0u);
        if (more)
            j_nxt = load_index4(base4 + v);
        while (__any_sync(full, more))
            {
            if (more)
                {
                const uint4 j = j_nxt;
                if (v + tpp < v_end)
                    j_nxt = load_index4(base4 + v + tpp);
                const Vec4<S> q0 = load4(a.pos, j.x);
                const Vec4<S> q1 = load4(a.pos, j.y);
                const Vec4<S> q2 = load4(a.pos, j.z);
                const Vec4<S> q3 = load4(a.pos, j.w);
                scan_dispatch(fam, a, g, j.x, q0);
                scan_dispatch(fam, a, g, j.y, q1);
                scan_dispatch(fam, a, g, j.z, q2);
                scan_dispatch(fam, a, g, j.w, q3);
                v += tpp;
                more = v < v_end;
                }
            while (__any_sync(full, queue_needs_drain(fam)))
                heavy_dispatch(fam, a, g);
            }
        unsigned int q = lane;
        while (__any_sync(full, q < n_left))
            {
            if (q < n_left)
                {
                const unsigned int idx = q < head_end - pre ? pre + q : tail_begin + (q - (head_end - pre));
                const unsigned int j = __ldg(base + idx);
                const Vec4<S> pj = load4(a.pos, j);
                scan_dispatch(fam, a, g, j, pj);
                q += tpp;
                }
            while (__any_sync(full, queue_needs_drain(fam)))
                heavy_dispatch(fam, a, g);
            }
        while (__any_sync(full, queue_pending(fam)))
            heavy_dispatch(fam, a, g);
        }

    split_finish_dispatch(fam, a, g);
    fam.finish(a, row, active && lane == 0, tpp);
    }

// LONGPASS = false: the main pass, one row per group of tpp lanes. When the caller announced
// rows longer than long_threshold (azp_pair_args.n_max) and tpp < 32, such rows are not
// evaluated here: lane 0 of the group queues the row and the group moves on, so one 1,800-entry
// colloid row no longer holds a 2-lane group (and its CTA slot) for the duration of ~15 ordinary
// rows. LONGPASS = true: the second pass, launched with tpp = 32 on a fixed grid; every warp
// takes queued rows until the queue is empty. If the queue overflows, the main pass evaluates
// the row in place, so correctness never depends on the queue capacity.
// ONE_LANE: the main pass compiled for threads_per_particle = 1 (the launch shape of every dense
// configuration): the lane stride of the neighbour loop, the shuffle reductions and the group
// broadcasts fold away at compile time instead of being re-derived from tpp_log2 every trip.
// Register cap of the one-lane main pass, per evaluator (IsoTraits<E>::one_lane_cap): 0 = none
// (launch bounds of the generic kernel), 80 = built for blocks of at most 256 threads with three
// of them resident, 96 / 72 = blocks of at most 128 threads with five / seven resident. A capped
// kernel only takes the block sizes it was built for (launch.cuh routes the others to the
// generic one).
template<class Fam> struct OneLaneCap
    {
    static constexpr int value = 0;
    };
template<class E, class S, bool X, bool V, int N> struct OneLaneCap<IsoFamily<E, S, X, V, N>>
    {
    static constexpr int value = sizeof(S) == 4 ? IsoTraits<E>::one_lane_cap : 0;
    };
template<class Fam, bool ONE_LANE> constexpr unsigned int row_kernel_max_threads()
    {
    return !ONE_LANE || OneLaneCap<Fam>::value == 0 ? max_block<typename Fam::S>() : (OneLaneCap<Fam>::value == 80 ? 256u : 128u);
    }
template<class Fam, bool ONE_LANE> constexpr unsigned int row_kernel_min_blocks()
    {
    return !ONE_LANE || OneLaneCap<Fam>::value == 0 ? 0u // 0 = unspecified
                                                       : (OneLaneCap<Fam>::value == 72 ? 7u : (OneLaneCap<Fam>::value == 96 ? 5u : 3u));
    }
#define AZP_ROW_KERNEL_BOUNDS(Fam, ONE_LANE) __launch_bounds__(row_kernel_max_threads<Fam, ONE_LANE>(), row_kernel_min_blocks<Fam, ONE_LANE>())
template<class Fam, bool LONGPASS, bool ONE_LANE = false>
__global__ void AZP_ROW_KERNEL_BOUNDS(Fam, ONE_LANE)
    row_kernel(const __grid_constant__ KernelArgs<typename Fam::S> a,
               const typename Fam::E::param_type* __restrict__ params,
               const unsigned int tpp_log2_arg)
    {
    const unsigned int ntp = Fam::NTM == 1 ? 1u : a.ntypes * a.ntypes;
    Fam fam;
    fam.stage(a, params, ntp);

    const unsigned int tpp_log2 = ONE_LANE ? 0u : tpp_log2_arg;
    const unsigned int tpp = ONE_LANE ? 1u : (1u << tpp_log2);
    const unsigned int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int lane = ONE_LANE ? 0u : (gtid & (tpp - 1u));
    if (!LONGPASS && !Fam::MULTIROW)
        {
        // one row per group of tpp lanes (the launch layer sizes the grid accordingly)
        const unsigned int slot = ONE_LANE ? gtid : (gtid >> tpp_log2);
        const unsigned int nslots = a.row_ids ? a.n_row_ids : a.N;
        bool active = slot < nslots;
        unsigned int row = 0, n = 0;
        uint64_t head = 0;
        if (active)
            {
            row = a.row_ids ? __ldg(a.row_ids + slot) : slot;
            n = __ldg(a.n_neigh + row);
            head = __ldg(a.head_list + row);
            if (a.long_queue && n > a.long_threshold)
                {
                // every lane of the group takes the same decision: lane 0 reserves the slot and
                // broadcasts the outcome through the group's shuffle
                unsigned int pos = 0xffffffffu;
                if (lane == 0)
                    {
                    pos = atomicAdd(a.long_count, 1u);
                    if (pos < a.long_capacity)
                        a.long_queue[pos] = row;
                    }
                pos = __shfl_sync(__activemask(), pos, 0, tpp);
                if (pos < a.long_capacity)
                    {
                    active = false; // deferred: the second pass writes this row
                    n = 0;
                    }
                }
            }
        constexpr bool STAGED = AZP_STAGE_LIST != 0 && ONE_LANE && Fam::PIPE == 2;
        if (STAGED)
            {
            ListStage ls;
            ls.carve((unsigned int)Fam::smem_bytes(ntp, blockDim.x));
            ls.init();
            process_row<Fam, STAGED>(fam, a, ntp, row, n, head, active, lane, tpp, &ls);
            }
        else
            process_row<Fam, false>(fam, a, ntp, row, n, head, active, lane, tpp);
        }
    else if (!LONGPASS)
        {
        // A group of tpp lanes takes the rows slot, slot + G, slot + 2 G, ... (G = groups in the
        // grid); the launch layer sizes the grid so that this is `rows_per_group` rows
        // (launch.cuh, kRowsPerGroup). With more than one row per group the (n_neigh, head_list)
        // pair of the next row is loaded, and the first 128-byte line of its neighbour list
        // prefetched into L1, while the current row is evaluated: a short row (C4: 35 entries,
        // C5: 20) is otherwise two dependent cold misses (head, then the list line) in front of a
        // few hundred cycles of work.
        const unsigned int nslots = a.row_ids ? a.n_row_ids : a.N;
        const unsigned int groups = (gridDim.x * blockDim.x) >> tpp_log2;
        const unsigned int per_warp = 32u >> tpp_log2;
        const unsigned int in_warp = (gtid & 31u) >> tpp_log2;
        unsigned int s0 = (gtid >> 5) * per_warp;
        // metadata of the next two rows: (n, head) of row k + 2 is requested while row k is
        // evaluated, so the prefetch of row k + 1's list line never waits for its address
        unsigned int row_1 = 0, n_1 = 0, row_2 = 0, n_2 = 0;
        uint64_t head_1 = 0, head_2 = 0;
        bool act_1 = false, act_2 = false;
        auto fetch = [&](unsigned int slot0, unsigned int& row, unsigned int& n, uint64_t& head, bool& active)
            {
            const unsigned int slot = slot0 + in_warp;
            active = slot0 < nslots && slot < nslots;
            row = 0, n = 0, head = 0;
            if (active)
                {
                row = a.row_ids ? __ldg(a.row_ids + slot) : slot;
                n = __ldg(a.n_neigh + row);
                head = __ldg(a.head_list + row);
                }
            };
        fetch(s0, row_1, n_1, head_1, act_1);
        if (s0 + groups < nslots)
            fetch(s0 + groups, row_2, n_2, head_2, act_2);
        bool first = true;
        for (; s0 < nslots; s0 += groups)
            {
            unsigned int row = row_1, n = n_1;
            uint64_t head = head_1;
            bool active = act_1;
            row_1 = row_2, n_1 = n_2, head_1 = head_2, act_1 = act_2;
            act_2 = false;
            if (s0 + groups < nslots)
                {
                if (s0 + 2u * groups < nslots)
                    fetch(s0 + 2u * groups, row_2, n_2, head_2, act_2);
                if (act_1 && lane == 0)
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(a.nlist + head_1));
                }
            if (active && a.long_queue && n > a.long_threshold)
                {
                // every lane of the group takes the same decision: lane 0 reserves the slot and
                // broadcasts the outcome through the group's shuffle
                unsigned int pos = 0xffffffffu;
                if (lane == 0)
                    {
                    pos = atomicAdd(a.long_count, 1u);
                    if (pos < a.long_capacity)
                        a.long_queue[pos] = row;
                    }
                pos = __shfl_sync(__activemask(), pos, 0, tpp);
                if (pos < a.long_capacity)
                    {
                    active = false; // deferred: the second pass writes this row
                    n = 0;
                    }
                }
            if (!first)
                fam.reset();
            first = false;
            process_row(fam, a, ntp, row, n, head, active, lane, tpp);
            }
        }
    else
        {
        const unsigned int nslots = min(*a.long_count, a.long_capacity);
        const unsigned int groups = (gridDim.x * blockDim.x) >> tpp_log2;
        const unsigned int per_warp = 32u >> tpp_log2;
        // warp-uniform trip count: the groups of a warp walk the queue together
        for (unsigned int s0 = ((gtid >> 5) * per_warp); s0 < nslots; s0 += groups)
            {
            const unsigned int slot = s0 + ((gtid & 31u) >> tpp_log2);
            const bool active = slot < nslots;
            unsigned int row = 0, n = 0;
            uint64_t head = 0;
            if (active)
                {
                row = a.long_queue[slot];
                n = __ldg(a.n_neigh + row);
                head = __ldg(a.head_list + row);
                }
            fam.reset();
            process_row(fam, a, ntp, row, n, head, active, lane, tpp);
            }
        }
    }
    } // namespace azp

#endif
