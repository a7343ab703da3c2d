/* azp_b200.h -- C ABI of the B200-native azplugins pair-force path.
 *
 * This is the drop-in boundary. Each entry point replaces one symbol that HOOMD-blue's host
 * ForceCompute classes call and that azplugins instantiates (all paths relative to the
 * reference tree):
 *
 *   azp_pair_forces_f32/_f64   <- hoomd::md::kernel::gpu_compute_pair_forces<E>(pair_args_t,
 *                                 const E::param_type*), instantiated at
 *                                 src/PotentialPairGPUKernel.cu.inc:25-28 for
 *                                 E = PairEvaluator{Colloid,ExpandedYukawa,Hertz,
 *                                 PerturbedLennardJones} (src/CMakeLists.txt:45-50) and, as
 *                                 "PotentialPairConservativeGeneralWeight", for the DPD evaluator's
 *                                 evalForceAndEnergy (src/export_PotentialPairDPDThermo.cc.inc:33-35)
 *   azp_dpd_forces_f32/_f64    <- gpu_compute_dpd_forces<E>(dpd_pair_args_t, const param_type*),
 *                                 src/PotentialPairDPDThermoGPUKernel.cu.inc:21-24
 *   azp_aniso_forces_f32/_f64  <- gpu_compute_pair_aniso_forces<E>(a_pair_args_t,
 *                                 const param_type*, const shape_type*),
 *                                 src/AnisoPotentialPairGPUKernel.cu.inc:21-25
 *   azp_param_size/pack/unpack <- E::param_type(pybind11::dict) / asDict() / toPython(), e.g.
 *                                 src/PairEvaluatorPerturbedLennardJones.h:33-54
 *
 * The C++ shim include/azp_hoomd_shim.h re-declares the three HOOMD templates on top of
 * these functions; INTEGRATION.md shows the binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - all d_* pointers are DEVICE pointers owned by the caller (HOOMD GlobalArrays); the library
 *     allocates nothing that outlives a call except a small per-device scratch for long rows;
 *   - f32 entry points read/write float / float4 arrays, f64 double / double4 (HOOMD `Scalar`
 *     for HOOMD_LONGREAL_SIZE 32 / 64);
 *   - array layouts are HOOMD's: pos = (x,y,z,type bit-cast),
 *     vel = (vx,vy,vz,mass), orientation = quaternion (s,x,y,z), force = (fx,fy,fz,energy),
 *     torque = (tx,ty,tz,0), virial[k*pitch + i] with k = xx,xy,xz,yy,yz,zz;
 *     n_neigh/nlist/head_list are NeighborListGPU's arrays in `full` storage mode;
 *   - outputs are overwritten for every row (zeros for empty rows);
 *   - every function returns a cudaError_t value as int (0 = success, 1 = cudaErrorInvalidValue
 *     for bad arguments) and never throws; kernels are enqueued on `stream` and the call does
 *     not synchronise;
 *   - there is NO CPU fallback: without a CUDA device the compute entry points fail.
 */
#ifndef AZP_B200_H_
#define AZP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
    {
#endif

#define AZP_B200_ABI_VERSION 1

/* evaluator ids */
enum azp_evaluator
    {
    AZP_EV_PERTURBED_LENNARD_JONES = 0, /* src/PairEvaluatorPerturbedLennardJones.h */
    AZP_EV_EXPANDED_YUKAWA = 1,         /* src/PairEvaluatorExpandedYukawa.h */
    AZP_EV_COLLOID = 2,                 /* src/PairEvaluatorColloid.h */
    AZP_EV_HERTZ = 3,                   /* src/PairEvaluatorHertz.h */
    AZP_EV_DPD_GENERAL_WEIGHT = 4,      /* src/DPDPairEvaluatorGeneralWeight.h */
    AZP_EV_TWO_PATCH_MORSE = 5,         /* src/AnisoPairEvaluatorTwoPatchMorse.h */
    AZP_EV_COUNT = 6
    };

/* energy shift modes of md::PotentialPair */
enum azp_shift_mode
    {
    AZP_SHIFT_NONE = 0,
    AZP_SHIFT_SHIFT = 1,
    AZP_SHIFT_XPLOR = 2
    };

/* HOOMD BoxDim flattened. Lengths are converted to Scalar inside the call and Linv = 1/L is
 * formed in Scalar precision, as BoxDim does. The box is centred on the origin. */
typedef struct azp_box
    {
    double L[3];
    double tilt[3];      /* xy, xz, yz */
    int32_t periodic[3]; /* per direction */
    int32_t _pad;
    } azp_box;

/* Union of HOOMD's pair_args_t / dpd_pair_args_t / a_pair_args_t. Fields a
 * family does not use may be NULL / 0. */
typedef struct azp_pair_args
    {
    /* outputs */
    void* d_force;         /* Scalar4[N] */
    void* d_virial;        /* Scalar[6 * virial_pitch]; may be NULL when compute_virial == 0 */
    void* d_torque;        /* Scalar4[N], aniso only */
    uint64_t virial_pitch;
    /* particle data, indexed by global particle index (local rows first, then ghosts) */
    const void* d_pos;         /* Scalar4[>= row_offset + N, + ghosts] */
    const void* d_vel;         /* Scalar4, DPD only */
    const void* d_orientation; /* Scalar4, aniso only */
    const uint32_t* d_tag;     /* DPD only */
    /* neighbour list (full storage), indexed by row */
    const uint32_t* d_n_neigh;
    const uint32_t* d_nlist;
    const uint64_t* d_head_list;
    uint64_t size_neigh_list;
    /* per type-pair tables, Index2D(ntypes)(i,j) = j*ntypes + i */
    const void* d_rcutsq; /* Scalar[ntypes^2] */
    const void* d_ronsq;  /* Scalar[ntypes^2], isotropic xplor only */
    azp_box box;
    uint32_t N;      /* rows to compute */
    uint32_t ntypes;
    uint32_t shift_mode;     /* azp_shift_mode; DPD ignores it, aniso accepts none/shift */
    uint32_t compute_virial; /* 0/1 */
    /* launch parameters (HOOMD's Autotuner<2> dimensions); 0 = let the library choose.
     * block_size: multiple of 32, <= 512 (f32) / <= 256 (f64) */
    uint32_t block_size;
    uint32_t threads_per_particle; /* power of two <= 32 */
    /* DPD thermostat */
    uint32_t seed;     /* uint16 range */
    /* largest row capacity of the list (HOOMD pair_args_t::n_max); 0 = unknown. When it exceeds
     * 512 the library defers rows longer than that to a second warp-per-row pass (row-length
     * skew, e.g. colloids in solvent). Results do not depend on it. */
    uint32_t n_max;
    uint64_t timestep; /* truncated to 32 bits like the reference evaluator does */
    double deltaT;
    double T;
    /* Extensions for the multi-GPU particle-slice scheduler (0 / NULL = HOOMD behaviour):
     * row r describes particle i = row_offset + r; outputs, n_neigh and head_list are indexed
     * by r, gathered particle data by global index. d_row_ids, when set, lists the n_row_ids
     * rows to compute (e.g. interior rows while the position exchange is in flight) and rows
     * not listed are left untouched. */
    uint32_t row_offset;
    uint32_t n_row_ids;
    const uint32_t* d_row_ids;
    } azp_pair_args;

/* library / device */
int azp_abi_version(void);
const char* azp_error_string(int code);
const char* azp_evaluator_name(int evaluator); /* reference getName() strings */

/* param_type staging: sizes match the reference structs (fp32/fp64): PLJ, Yukawa, Colloid, DPD
 * 16/32, Hertz 4/8, TwoPatchMorse 24/48. `fields` order:
 *   PLJ {epsilon, sigma, attraction_scale_factor}   Yukawa {epsilon, kappa, delta}
 *   Colloid {A, a_1, a_2, sigma}   Hertz {epsilon}   DPD {A, gamma, s}
 *   TwoPatchMorse {M_d, M_r, r_eq, omega, alpha, repulsion(0/1)}
 * pack reproduces the roundings of the reference constructors; unpack those of asDict(). */
int azp_param_num_fields(int evaluator);
int azp_param_size(int evaluator, int scalar_bits);
int azp_param_pack(int evaluator, int scalar_bits, const double* fields, void* out);
int azp_param_unpack(int evaluator, int scalar_bits, const void* in, double* fields);

/* force kernels; `stream` is a cudaStream_t (NULL = default stream) */
int azp_pair_forces_f32(int evaluator, const azp_pair_args* args, const void* d_params, void* stream);
int azp_pair_forces_f64(int evaluator, const azp_pair_args* args, const void* d_params, void* stream);
int azp_dpd_forces_f32(int evaluator, const azp_pair_args* args, const void* d_params, void* stream);
int azp_dpd_forces_f64(int evaluator, const azp_pair_args* args, const void* d_params, void* stream);
int azp_aniso_forces_f32(int evaluator, const azp_pair_args* args, const void* d_params,
                         const void* d_shape_params, void* stream);
int azp_aniso_forces_f64(int evaluator, const azp_pair_args* args, const void* d_params,
                         const void* d_shape_params, void* stream);

/* Fused two-potential pass: two isotropic potentials evaluated over ONE
 * sweep of the neighbour list -- the reference's documented case is pair.Colloid + pair.Hertz on
 * one nlist (src/pair.py:66-76), which HOOMD runs as two ForceComputes that each stream the
 * list. `a` and `b` are the argument structs the two separate calls would get: they must name
 * the same particles, rows and list (d_pos, d_n_neigh, d_nlist, d_head_list, N, ntypes,
 * row_offset, d_row_ids, compute_virial, virial_pitch equal); outputs, d_rcutsq, shift_mode
 * (none / shift) and parameters are per potential; launch parameters are taken from `a`. The
 * outputs equal those of the two separate calls bit for bit. Supported pair: (AZP_EV_COLLOID,
 * AZP_EV_HERTZ); anything else, or xplor, returns cudaErrorNotSupported (801) and the caller
 * launches the potentials separately. */
int azp_pair_forces_fused_f32(int evaluator_a, const azp_pair_args* a, const void* d_params_a,
                              int evaluator_b, const azp_pair_args* b, const void* d_params_b,
                              void* stream);
int azp_pair_forces_fused_f64(int evaluator_a, const azp_pair_args* a, const void* d_params_a,
                              int evaluator_b, const azp_pair_args* b, const void* d_params_b,
                              void* stream);

/* Launch autotuner (HOOMD Autotuner<2> equivalent): times the (block_size, threads_per_particle)
 * candidates on the given arguments with CUDA events, writes the fastest pair and its time.
 * family: 0 pair, 1 dpd, 2 aniso. Synchronises the stream. */
int azp_autotune(int family, int evaluator, int scalar_bits, const azp_pair_args* args,
                 const void* d_params, void* stream, uint32_t* best_block, uint32_t* best_tpp,
                 float* best_ms);

/* Halo packing for the multi-GPU particle-slice scheduler: d_dst[k] = d_src[d_idx[k]] for rows of
 * row_bytes (a multiple of 16: Scalar4 = 16 or 32). One coalesced 16-byte store per thread. */
int azp_gather_rows(const void* d_src, const int64_t* d_idx, uint64_t n, uint32_t row_bytes,
                    void* d_dst, void* stream);

/* Halo push over peer memory: row k of the send list goes to the address d_dst_addr[k], which may
 * be a peer-mapped address of another GPU's ghost region (NVLink P2P stores; the caller orders the
 * push against the readers with barriers). row_bytes as above. */
int azp_push_rows(const void* d_src, const int64_t* d_idx, const uint64_t* d_dst_addr, uint64_t n,
                  uint32_t row_bytes, void* stream);

/* External harmonic barriers. Replaces
 *   gpu::compute_harmonic_barrier<BarrierEvaluatorT>(d_force, d_virial, d_pos, d_params,
 *       global_box, evaluator, N, ntypes, block_size)        reference src/HarmonicBarrierGPU.cuh:100-132
 * for BarrierEvaluatorT = PlanarBarrierEvaluator (src/PlanarBarrierEvaluator.h:31-63: plane normal
 * to +y at y = location) and SphericalBarrierEvaluator (src/SphericalBarrierEvaluator.h:30-67:
 * sphere of radius location about the origin). d_params is the reference's Scalar2[ntypes]
 * {k, offset} (src/HarmonicBarrier.h:120-122); positions are wrapped into the box first (:86-88);
 * the virial array is zeroed in the same pass (:129). Bit-identical to the reference CPU class. */
enum azp_barrier_geometry
    {
    AZP_BARRIER_PLANAR = 0,
    AZP_BARRIER_SPHERICAL = 1
    };
typedef struct azp_barrier_args
    {
    void* d_force;        /* Scalar4[N], overwritten */
    void* d_virial;       /* Scalar[6 * virial_pitch], zeroed; may be NULL */
    uint64_t virial_pitch;
    const void* d_pos;    /* Scalar4[N] */
    const void* d_params; /* Scalar2[ntypes] {k, offset} */
    azp_box box;          /* global box */
    double location;      /* the Variant's value at this timestep: H (planar) or R (spherical) */
    uint32_t N;
    uint32_t ntypes;
    int32_t geometry;     /* azp_barrier_geometry */
    uint32_t block_size;  /* 0 = library default (256); multiple of 32, <= 256 */
    } azp_barrier_args;
int azp_harmonic_barrier_f32(const azp_barrier_args* args, void* stream);
int azp_harmonic_barrier_f64(const azp_barrier_args* args, void* stream);
/* BarrierEvaluator::valid(global_box): 1 when the barrier lies inside the box
 * (src/HarmonicBarrier.h:126-130 throws "Barrier position is invalid" otherwise). */
int azp_harmonic_barrier_valid(int geometry, int scalar_bits, double location, const azp_box* box);

/* Wall potentials wall.Colloid / wall.LJ93. Replaces
 *   hoomd::md::kernel::gpu_compute_potential_external_forces<EvaluatorWalls<E>>(args, d_params, d_field)
 * as instantiated by reference src/PotentialExternalWallGPUKernel.cu.inc:13-34 for
 * E = WallEvaluatorColloid (src/WallEvaluatorColloid.h) and WallEvaluatorLJ93
 * (src/WallEvaluatorLJ93.h). Layouts (S = float | double, all fields S unless noted):
 *   d_params[ntypes]: Colloid {c_1 = A sigma^6 / 7560, c_2 = A / 6, a, rcutsq, rextrap}
 *                     LJ93    {sigma_3, A, rcutsq, rextrap}         (azp_wall_param_size bytes each)
 *   d_walls: { uint32 n_spheres, n_cylinders, n_planes, pad;
 *              spheres[20]   {r, origin[3], int32 inside, int32 open};
 *              cylinders[20] {r, origin[3], axis[3] (unit), int32 inside, int32 open};
 *              planes[60]    {origin[3], normal[3] (unit), int32 open, int32 pad} }  (azp_walls_size bytes)
 * rextrap > 0 selects HOOMD's extrapolated mode (linear continuation closer than rextrap to a wall
 * or behind it). Energy is always shifted at r_cut, the virial is F_a * pos_b, as HOOMD's wall
 * loop does. */
enum azp_wall_evaluator
    {
    AZP_WALL_COLLOID = 0,
    AZP_WALL_LJ93 = 1
    };
#define AZP_MAX_SPHERE_WALLS 20
#define AZP_MAX_CYLINDER_WALLS 20
#define AZP_MAX_PLANE_WALLS 60
typedef struct azp_wall_args
    {
    void* d_force;        /* Scalar4[N], overwritten */
    void* d_virial;       /* Scalar[6 * virial_pitch], overwritten; may be NULL */
    uint64_t virial_pitch;
    const void* d_pos;    /* Scalar4[N] */
    const void* d_params; /* per type, see above */
    const void* d_walls;  /* wall list, see above */
    uint32_t N;
    uint32_t ntypes;
    uint32_t block_size;  /* 0 = library default (256); multiple of 32, <= 256 */
    uint32_t _pad;
    } azp_wall_args;
int azp_wall_forces_f32(int evaluator, const azp_wall_args* args, void* stream);
int azp_wall_forces_f64(int evaluator, const azp_wall_args* args, void* stream);
int azp_wall_param_size(int evaluator, int scalar_bits);
int azp_walls_size(int scalar_bits);

/* Velocity-Verlet (NVE) steps either side of the force path: what HOOMD's
 * md.methods.ConstantVolume does around the force computes the reference plugs in (usage:
 * reference src/pytest/test_pair.py:325-327). HOOMD is not in the reference tree: restated as
 *   step one: v += a dt/2; x += v dt; wrap into the (orthorhombic) box, counting images
 *   step two: F = sum of d_forces[0..n_forces); a = F/m (m = vel.w); v += a dt/2
 * IEEE arithmetic without FMA contraction (a numpy restatement is bit-exact). */
#define AZP_MD_MAX_FORCES 8
typedef struct azp_md_args
    {
    void* d_pos;       /* Scalar4[N] (step one) */
    void* d_vel;       /* Scalar4[N] (vx, vy, vz, mass) */
    void* d_accel;     /* Scalar4[N] (ax, ay, az, 0): read by step one, written by step two */
    int32_t* d_image;  /* int32[3 N] or NULL (step one) */
    void* d_net_force; /* Scalar4[N] or NULL: sum of the forces, energy in .w (step two) */
    const void* d_forces[AZP_MD_MAX_FORCES]; /* Scalar4[N] each (step two) */
    uint32_t n_forces;
    uint32_t N;
    azp_box box;
    double dt;
    } azp_md_args;
int azp_nve_step_one_f32(const azp_md_args* args, void* stream);
int azp_nve_step_one_f64(const azp_md_args* args, void* stream);
int azp_nve_step_two_f32(const azp_md_args* args, void* stream);
int azp_nve_step_two_f64(const azp_md_args* args, void* stream);

/* Langevin thermostat (hoomd.md.methods.Langevin; BASELINE.json configs[0] runs the
 * PerturbedLennardJones fluid under it). Step one is azp_nve_step_one; this step two adds the
 * Brownian force per particle before the second half kick:
 *   F_bd = coeff (r_x, r_y, r_z) - gamma v;  coeff = sqrt(6 gamma kT / dt);  r ~ Uniform(-1, 1)
 *   a = (sum of d_forces + F_bd) / m;  v += a dt / 2
 * gamma per particle type (type id bit-cast in pos.w; args->d_pos must be set). Random numbers:
 * HOOMD's RandomGenerator(Seed(rng_id, timestep, seed), Counter(tag)), three draws (Philox4x32-10,
 * SURVEY.md Appendix B). HOOMD is not in the reference tree: rng_id (24 = RNGIdentifier::
 * TwoStepLangevin as recalled from HOOMD v7.0.1) is a caller-supplied field, parity with HOOMD's
 * stream is unpinned. IEEE arithmetic without FMA contraction. */
typedef struct azp_langevin_args
    {
    const uint32_t* d_tag; /* u32[N] */
    const void* d_gamma;   /* Scalar[ntypes] */
    uint32_t ntypes;
    uint32_t seed;         /* 16 bits used */
    uint64_t timestep;
    double kT;
    uint32_t rng_id;       /* 8 bits used */
    uint32_t noiseless;    /* != 0: drag only (HOOMD's tally / noiseless_t) */
    } azp_langevin_args;
int azp_langevin_step_two_f32(const azp_md_args* args, const azp_langevin_args* langevin, void* stream);
int azp_langevin_step_two_f64(const azp_md_args* args, const azp_langevin_args* langevin, void* stream);

/* Uniform(-1,1) value the DPD evaluator draws for a pair (host side; same code as the kernel).
 * Exposes the RNG keying of src/DPDPairEvaluatorGeneralWeight.h:213-233 for parity tests. */
double azp_dpd_alpha(int scalar_bits, uint32_t seed, uint32_t tag_i, uint32_t tag_j, uint64_t timestep);
void azp_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* HOOMD-layout full neighbour list on the GPU (row "next" of SURVEY.md 8(f), used to feed the
 * path): cell-list build of n_neigh / nlist / head_list with r_list = r_cut + buffer per type
 * pair. Two calls: azp_nlist_count fills d_n_neigh; the caller turns per-type maxima into
 * head_list (prefix sum of Nmax[type]); azp_nlist_fill writes the rows. d_cell_* are scratch
 * arrays sized by azp_nlist_scratch_sizes. */
typedef struct azp_nlist_args
    {
    const void* d_pos; /* Scalar4[N] */
    uint32_t N;
    uint32_t ntypes;
    azp_box box;
    const void* d_rlistsq; /* Scalar[ntypes^2] */
    double r_list_max;
    uint32_t* d_n_neigh;         /* [N] out (count) / in (fill) */
    const uint64_t* d_head_list; /* [N] in (fill) */
    uint32_t* d_nlist;           /* out (fill) */
    uint32_t* d_cell_of;         /* [N] scratch */
    uint32_t* d_cell_start;      /* [ncells + 1] scratch */
    uint32_t* d_cell_order;      /* [N] scratch */
    uint32_t cell_dim[3];        /* from azp_nlist_cell_dim */
    /* rows to build: particles [row_offset, row_offset + n_rows) against all N particles;
     * d_n_neigh / d_head_list are indexed by row - row_offset. n_rows = 0 means all N rows.
     * (Used by the multi-GPU scheduler: a rank builds only the rows of its particle slice.) */
    uint32_t row_offset;
    uint32_t n_rows;
    uint32_t _pad;
    /* Optional (azp_nlist_fill only): skip the count pass by reusing the row capacities of the
     * previous build, as HOOMD does. d_capacity[row] = slots of the row; the fill then writes at
     * most that many entries, stores the true count in d_n_neigh[row] and raises *d_overflow
     * (uint32, zeroed by the caller) when a row needed more -- the caller then falls back to
     * count + fill. NULL = exact fill after the count pass; d_n_neigh is then not written. */
    const uint32_t* d_capacity;
    uint32_t* d_overflow;
    void* d_cell_pos;         /* Scalar4[N] scratch: positions in cell order (written by bin) */
    void* d_pos_at_build;     /* Scalar4[N] or NULL: bin copies d_pos here for azp_nlist_moved */
    uint32_t threads_per_row; /* lanes per row in count / fill: 0 = library choice, else 1..32 (2^k) */
    uint32_t _pad2;
    } azp_nlist_args;

int azp_nlist_cell_dim(const azp_box* box, double r_list_max, uint32_t dim[3]);
int azp_nlist_bin_f32(const azp_nlist_args* args, void* stream);
int azp_nlist_bin_f64(const azp_nlist_args* args, void* stream);
int azp_nlist_count_f32(const azp_nlist_args* args, void* stream);
int azp_nlist_count_f64(const azp_nlist_args* args, void* stream);
int azp_nlist_fill_f32(const azp_nlist_args* args, void* stream);
int azp_nlist_fill_f64(const azp_nlist_args* args, void* stream);
/* Displacement check on the device (HOOMD NeighborList::distanceCheck): *d_flag (zeroed by the
 * caller) is raised when some particle is farther than max_dist (= buffer / 2) from its position
 * at the last build, minimum image applied. */
int azp_nlist_moved_f32(const void* d_pos, const void* d_pos_at_build, const azp_box* box, double max_dist, uint32_t N, uint32_t* d_flag, void* stream);
int azp_nlist_moved_f64(const void* d_pos, const void* d_pos_at_build, const azp_box* box, double max_dist, uint32_t N, uint32_t* d_flag, void* stream);
/* Space-filling-curve order of the particles (what HOOMD's SFCPackTuner applies to ParticleData so
 * that neighbours in space are neighbours in memory): d_order[k] = index of the particle that
 * comes k-th along a 30-bit Morton curve through the box. The caller permutes its arrays. */
int azp_sfc_order_f32(const void* d_pos, const azp_box* box, uint32_t N, uint32_t* d_order, void* stream);
int azp_sfc_order_f64(const void* d_pos, const azp_box* box, uint32_t N, uint32_t* d_order, void* stream);

#ifdef __cplusplus
    }
#endif

#endif /* AZP_B200_H_ */
