"""Velocity-Verlet steps: bit-exact against a numpy restatement of the
two streaming kernels, and energy / momentum conservation of a Lennard-Jones-like fluid driven by
the pair kernels with neighbour-list rebuilds."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _numpy_step_one(pos, vel, accel, image, L, dt):
    S = pos.dtype.type
    half_dt = S(0.5) * S(dt)
    v = vel.copy()
    v[:, :3] = vel[:, :3] + accel[:, :3] * half_dt
    x = pos[:, :3] + v[:, :3] * S(dt)
    img = image.copy()
    for d in range(3):
        lo = -S(L[d]) / S(2.0)
        hi = lo + S(L[d])
        up, down = x[:, d] >= hi, x[:, d] < lo
        x[up, d] = x[up, d] - S(L[d])
        x[down & ~up, d] = x[down & ~up, d] + S(L[d])
        img[up, d] += 1
        img[down & ~up, d] -= 1
    p = pos.copy()
    p[:, :3] = x
    return p, v, img


def _numpy_step_two(vel, forces, dt):
    S = vel.dtype.type
    F = np.zeros_like(forces[0])
    for f in forces:
        F = F + f
    minv = S(1.0) / vel[:, 3]
    a = np.zeros_like(vel)
    a[:, :3] = F[:, :3] * minv[:, None]
    v = vel.copy()
    v[:, :3] = vel[:, :3] + a[:, :3] * (S(0.5) * S(dt))
    return v, a, F


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_steps_bit_exact_against_numpy(dtype):
    import azplugins_b200 as az

    rng = np.random.default_rng(4)
    N, L, dt = 50001, (9.0, 11.0, 13.0), 0.005
    xyz = rng.uniform(-0.5, 0.5, (N, 3)) * np.array(L)
    xyz[:200] = np.sign(xyz[:200]) * (0.5 * np.array(L) - 1e-4)  # about to cross a face
    vel = rng.standard_normal((N, 3)) * 3.0
    state = az.State(az.Box(*L), ["A"], xyz, velocity=vel, mass=rng.uniform(0.5, 2.0, N), dtype=dtype)
    bar = az.external.SphericalHarmonicBarrier(location=3.0)
    bar.params["A"] = dict(k=20.0, offset=0.0)
    pla = az.external.PlanarHarmonicBarrier(location=1.0)
    pla.params["A"] = dict(k=5.0, offset=0.5)
    ig = az.md.Integrator(dt=dt, forces=[bar, pla]).attach(state)
    ig.accel.copy_(torch.from_numpy(rng.standard_normal((N, 4)).astype(dtype)))
    p0, v0, a0 = state.pos.cpu().numpy(), state.vel.cpu().numpy(), ig.accel.cpu().numpy()
    args = ig._args()
    ig._call("azp_nve_step_one", args)
    p1, v1, img1 = _numpy_step_one(p0, v0, a0, np.zeros((N, 3), np.int32), L, dt)
    assert np.array_equal(state.pos.cpu().numpy(), p1)
    assert np.array_equal(state.vel.cpu().numpy(), v1)
    assert np.array_equal(ig.image.cpu().numpy(), img1)
    assert np.abs(img1).sum() > 0
    bar.compute()
    pla.compute()
    ig._call("azp_nve_step_two", args)
    v2, a2, F = _numpy_step_two(v1, [bar._force.cpu().numpy(), pla._force.cpu().numpy()], dt)
    assert np.array_equal(state.vel.cpu().numpy(), v2)
    assert np.array_equal(ig.accel.cpu().numpy()[:, :3], a2[:, :3])
    assert np.array_equal(ig.net_force.cpu().numpy(), F)


def test_nve_conserves_energy_and_momentum():
    import azplugins_b200 as az
    from azplugins_b200 import synth

    rng = np.random.default_rng(9)
    N = 8000
    xyz, L = synth.jittered_lattice(N, 0.8, rng, jitter=0.05)  # gentle start: no close contacts
    v = rng.standard_normal((N, 3))
    v -= v.mean(axis=0)
    state = az.State(az.Box.cube(L), ["A"], xyz, velocity=v, dtype=np.float64)
    nl = az.nlist.Cell(buffer=0.4)
    plj = az.pair.PerturbedLennardJones(nlist=nl, default_r_cut=3.0, mode="shift")
    plj.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0, attraction_scale_factor=0.5)
    ig = az.md.Integrator(dt=0.002, forces=[plj]).attach(state)
    ig.run(100)  # melt the lattice
    e0 = ig.kinetic_energy() + ig.potential_energy()
    p0 = ig.momentum()
    builds0 = nl.num_builds
    ig.run(400)
    e1 = ig.kinetic_energy() + ig.potential_energy()
    assert nl.num_builds > builds0, "the neighbour list was never rebuilt"
    assert abs(e1 - e0) < 2e-4 * abs(ig.kinetic_energy()), (e0, e1, ig.kinetic_energy())
    assert np.abs(ig.momentum() - p0).max() < 1e-8 * N
    assert state.timestep == 500


def _lj_system(dtype, delay):
    import azplugins_b200 as az
    from azplugins_b200 import synth

    rng = np.random.default_rng(12)
    N = 6000
    xyz, L = synth.jittered_lattice(N, 0.8, rng, jitter=0.05)
    v = rng.standard_normal((N, 3))
    v -= v.mean(axis=0)
    state = az.State(az.Box.cube(L), ["A"], xyz, velocity=v, dtype=dtype)
    nl = az.nlist.Cell(buffer=0.4, rebuild_check_delay=delay)
    plj = az.pair.PerturbedLennardJones(nlist=nl, default_r_cut=3.0, mode="shift")
    plj.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0, attraction_scale_factor=0.5)
    bar = az.external.SphericalHarmonicBarrier(location=0.45 * L)
    bar.params["A"] = dict(k=10.0, offset=0.0)
    return state, nl, az.md.Integrator(dt=0.002, forces=[plj, bar]).attach(state)


def test_graph_replay_equals_eager_steps():
    """CUDA-graph replay of the un-checked steps (rebuild_check_delay) gives the same trajectory,
    bit for bit, as launching every step, with the same neighbour-list rebuilds."""
    runs = {}
    for graph in (False, True):
        state, nl, ig = _lj_system(np.float32, delay=6)
        ig.run(150, graph=graph)
        torch.cuda.synchronize()
        runs[graph] = (state.pos.clone(), state.vel.clone(), nl.num_builds, state.timestep)
    assert runs[False][2] == runs[True][2] and runs[True][2] >= 3
    assert runs[False][3] == runs[True][3] == 150
    assert torch.equal(runs[False][0], runs[True][0])
    assert torch.equal(runs[False][1], runs[True][1])


def test_graph_mode_rejects_time_dependent_forces():
    import azplugins_b200 as az

    state, nl, ig = _lj_system(np.float32, delay=4)
    ig.forces[1].location = lambda t: 5.0  # a variant: may depend on the time step
    with pytest.raises(ValueError):
        ig.run(3, graph=True)


def test_dpd_temperature_reference_test():
    """The reference's only test of the DPD thermostat, src/pytest/test_pair_dpd.py:13-46,
    reproduced step for step: 10^3 simple-cubic lattice (a = 0.6, box 6 x 6 x 6), momenta
    thermalised at kT = 1.5, DPDGeneralWeight(kT = 1.5, r_cut = 1) with A = 0, gamma = 4.5,
    s = 0.5 (random + drag part only), Cell(buffer = 0.4), NVE at dt = 0.01; 10 steps, then the
    kinetic temperature averaged over 100 steps must be 1.5 within 10 %. This is the one
    reference-held check of the random force (amplitude, weight, Philox stream statistics)."""
    import azplugins_b200 as az

    n, a, kT = 10, 0.6, 1.5
    g = (np.arange(n) + 0.5) * a - 0.5 * n * a
    xyz = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    rng = np.random.default_rng(12)
    vel = rng.standard_normal(xyz.shape) * np.sqrt(kT)
    vel -= vel.mean(axis=0)  # thermalize_particle_momenta removes the centre-of-mass momentum
    N = len(xyz)
    ndof = 3 * N - 3
    vel *= np.sqrt(kT * ndof / (vel ** 2).sum())  # ... and rescales to kT exactly
    for dtype in (np.float32, np.float64):
        # the State is created at another dt on purpose: the integrator's dt must reach the force
        state = az.State(az.Box.cube(n * a), ["A"], xyz, velocity=vel, dtype=dtype, seed=7, dt=0.005)
        cell = az.nlist.Cell(buffer=0.4)
        dpd = az.pair.DPDGeneralWeight(nlist=cell, kT=kT, default_r_cut=1.0)
        dpd.params[("A", "A")] = dict(A=0.0, gamma=4.5, s=0.5)
        ig = az.md.Integrator(dt=0.01, forces=[dpd], methods=[az.md.ConstantVolume()]).attach(state)
        assert state.dt == 0.01
        ig.run(10)
        samples = np.zeros(100)
        for k in range(100):
            samples[k] = 2.0 * ig.kinetic_energy() / ndof
            ig.run(1)
        avg = samples.mean()
        print("DPD thermostat <kT> =", avg, np.dtype(dtype).name)
        assert avg == pytest.approx(1.5, 0.1)
        # momentum is conserved by the pairwise random and drag forces
        assert np.abs(ig.momentum()).max() < 1e-2 * np.sqrt(N * kT)


def _philox_u32(ctr0, tags, key0, key1):
    """First output word of Philox4x32-10 for counters {ctr0, 0, 0, tag} (numpy, uint64 maths)."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    mask = np.uint64(0xFFFFFFFF)
    c = [np.full(tags.shape, ctr0, np.uint64), np.zeros(tags.shape, np.uint64),
         np.zeros(tags.shape, np.uint64), tags.astype(np.uint64)]
    k0, k1 = key0, key1
    for _ in range(10):
        p0 = np.uint64(M0) * c[0]
        p1 = np.uint64(M1) * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0), p1 & mask,
             (p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1), p0 & mask]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c[0].astype(np.uint64), c[1].astype(np.uint64)


def _numpy_langevin_two(vel, forces, tags, gamma, kT, dt, seed, timestep, rng_id):
    S = vel.dtype.type
    F = np.zeros_like(forces[0])
    for f in forces:
        F = F + f
    key0 = ((rng_id & 0xFF) << 24) | ((seed & 0xFFFF) << 8) | ((timestep >> 32) & 0xFF)
    key1 = timestep & 0xFFFFFFFF
    r = np.zeros((len(vel), 3), dtype=vel.dtype)
    for d in range(3):
        w0, w1 = _philox_u32(d, tags, key0, key1)
        if vel.dtype == np.float32:
            u = w0.astype(np.float32) * S(2.3283064365386963e-10) + S(1.1641532182693481e-10)
        else:
            u = ((w0 << np.uint64(32)) | w1).astype(np.float64) * S(5.421010862427522e-20) + S(2.710505431213761e-20)
        r[:, d] = S(-1) + S(2) * u
    coeff = np.sqrt(S(6) * gamma * S(kT) / S(dt))
    bd = r * coeff[:, None] - gamma[:, None] * vel[:, :3]
    minv = S(1.0) / vel[:, 3]
    a = np.zeros_like(vel)
    a[:, :3] = (F[:, :3] + bd) * minv[:, None]
    v = vel.copy()
    v[:, :3] = vel[:, :3] + a[:, :3] * (S(0.5) * S(dt))
    return v, a, F


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_langevin_step_two_bit_exact_against_numpy(dtype):
    """azp_langevin_step_two_*: drag + uniform random force keyed by (rng id, time step, seed, tag)
    exactly as the numpy restatement (Philox4x32-10 in uint64 arithmetic), two particle types with
    different gamma, a time step above 2^32."""
    import azplugins_b200 as az

    rng = np.random.default_rng(21)
    N, L, dt, kT, seed = 30011, (9.0, 11.0, 13.0), 0.004, 1.3, 0xBEEF
    xyz = rng.uniform(-0.49, 0.49, (N, 3)) * np.array(L)
    typeid = rng.integers(0, 2, N)
    state = az.State(az.Box(*L), ["A", "B"], xyz, typeid=typeid, velocity=rng.standard_normal((N, 3)),
                     mass=rng.uniform(0.5, 2.0, N), dtype=dtype)
    state.tag.copy_(torch.from_numpy(rng.permutation(N).astype(np.int32)).to(state.tag.dtype))
    bar = az.external.SphericalHarmonicBarrier(location=3.0)
    bar.params["A"] = dict(k=20.0, offset=0.0)
    bar.params["B"] = dict(k=10.0, offset=0.1)
    method = az.md.Langevin(kT=kT, default_gamma=1.5, seed=seed)
    method.gamma["B"] = 0.25
    ig = az.md.Integrator(dt=dt, forces=[bar], methods=[method]).attach(state)
    state.timestep = (3 << 32) + 12345
    bar.compute()
    v0 = state.vel.cpu().numpy()
    ig._step_two(ig._args())
    gamma = np.where(typeid == 0, 1.5, 0.25).astype(dtype)
    tags = state.tag.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    v1, a1, F = _numpy_langevin_two(v0, [bar._force.cpu().numpy()], tags, gamma, kT, dt, seed,
                                    state.timestep, az.md.Langevin.RNG_ID)
    assert np.array_equal(state.vel.cpu().numpy(), v1)
    assert np.array_equal(ig.accel.cpu().numpy()[:, :3], a1[:, :3])
    assert np.array_equal(ig.net_force.cpu().numpy(), F)
    # noiseless: drag only
    method.noiseless = True
    state.vel.copy_(torch.from_numpy(v0))
    ig._step_two(ig._args())
    minv = dtype(1.0) / v0[:, 3]
    a_drag = (F[:, :3] + (-gamma[:, None] * v0[:, :3])) * minv[:, None]
    assert np.array_equal(ig.accel.cpu().numpy()[:, :3], a_drag)


def test_langevin_config1_reaches_kT():
    """BASELINE.json configs[0] in small: PerturbedLennardJones fluid (rho = 0.8, r_cut = 3) under
    Langevin NVT. The kinetic temperature relaxes to kT from a cold start and stays there; the
    random force of two steps differs, that of the same (step, seed) does not."""
    import azplugins_b200 as az
    from azplugins_b200 import synth

    rng = np.random.default_rng(3)
    N, kT = 8000, 1.2
    xyz, L = synth.jittered_lattice(N, 0.8, rng, jitter=0.05)
    state = az.State(az.Box.cube(L), ["A"], xyz, velocity=np.zeros((N, 3)), dtype=np.float32)
    nl = az.nlist.Cell(buffer=0.4)
    plj = az.pair.PerturbedLennardJones(nlist=nl, default_r_cut=3.0, mode="shift")
    plj.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0, attraction_scale_factor=0.5)
    ig = az.md.Integrator(dt=0.005, forces=[plj], methods=[az.md.Langevin(kT=kT, default_gamma=1.0, seed=5)])
    ig.attach(state)
    ig.run(1500)  # ~7.5 relaxation times
    samples = []
    for _ in range(100):
        ig.run(5)
        samples.append(2.0 * ig.kinetic_energy() / (3 * N))
    avg = float(np.mean(samples))
    print("Langevin <kT> =", avg)
    assert avg == pytest.approx(kT, rel=0.03)
    assert nl.num_builds > 3
    with pytest.raises(ValueError):
        ig.run(2, graph=True)
