"""CPU model of the fine-grid sweep of the neighbour-list builder (csrc/nlist_kernels.cu,
nlist_rows_fine): the same grid arithmetic in numpy float32 -- cell of a particle, per-(z, y)
slab bounds, x interval, the two runs of a wrapped interval -- must offer every pair inside
r_list as a candidate, also in the smallest boxes the fine grid accepts (five cells per axis),
for particles on cell faces and box faces, and never offer a candidate twice."""

import numpy as np
import pytest

F = np.float32


def grid_dims(L, r):
    """azp_nlist_cell_dim's fine branch."""
    dims = []
    for d in range(3):
        n = 2.0 * L[d] / r
        c = 1024 if n > 1024.0 else int(n)
        w = L[d] / c
        if not (c >= 5 and w < r and 2.0 * w >= r):
            return None
        dims.append(c)
    return dims


def fma32(a, b, c):
    """One fused multiply-add in float32: the float32 product is exact in float64."""
    return (np.asarray(a, dtype=F).astype(np.float64) * np.asarray(b, dtype=F).astype(np.float64) + np.float64(c)).astype(F)


def bin_particle(p, L):
    """bin_particle of the kernel file: in-box image and wrapped fractional coordinates, both from
    ONE fused multiply-add of the raw position."""
    Linv = (F(1.0) / L.astype(F)).astype(F)
    f = fma32(p, np.broadcast_to(Linv, p.shape), 0.5)
    k = np.floor(f)
    fw = (f - k).astype(F)
    o = (p.astype(F) - (k * L.astype(F)).astype(F)).astype(F)
    return o, fw


def grid_coords(p, L, dim):
    """g (position inside the grid) and c (cell) of RAW positions p, float32 like the kernel."""
    _, fw = bin_particle(p, L)
    g = (fw * np.array(dim, dtype=F)).astype(F)
    c = np.clip(np.floor(g).astype(np.int64), 0, np.array(dim) - 1)
    return g, c


def binned_position(p, L):
    return bin_particle(p, L)[0]


def candidates(i, pos, L, dim, r, cells):
    """Candidate particles of row i with their image vectors, as nlist_rows_fine walks them."""
    g, c = grid_coords(pos[i:i + 1], L, dim)
    g, c = g[0], c[0]
    w = (L.astype(F) / np.array(dim, dtype=F)).astype(F)
    slack = (L.astype(F) * F(1e-5)).astype(F)
    rsq = F(r) * F(r)
    out = []
    for oz in range(-2, 3):
        dz = F(0) if oz == 0 else ((F(c[2] + oz) - g[2]) if oz > 0 else (g[2] - F(c[2] + oz + 1))) * w[2] - slack[2]
        dz = max(dz, F(0))
        cz, imgz = c[2] + oz, 0
        if cz < 0:
            cz, imgz = cz + dim[2], -1
        elif cz >= dim[2]:
            cz, imgz = cz - dim[2], 1
        for oy in range(-2, 3):
            dy = F(0) if oy == 0 else ((F(c[1] + oy) - g[1]) if oy > 0 else (g[1] - F(c[1] + oy + 1))) * w[1] - slack[1]
            dy = max(dy, F(0))
            cy, imgy = c[1] + oy, 0
            if cy < 0:
                cy, imgy = cy + dim[1], -1
            elif cy >= dim[1]:
                cy, imgy = cy - dim[1], 1
            remsq = rsq - dz * dz - dy * dy
            if remsq < 0:
                continue
            rx = (np.sqrt(F(remsq)) + slack[0]) / w[0]
            xlo = max(c[0] - 2, int(np.floor(g[0] - rx)))
            xhi = min(c[0] + 2, int(np.floor(g[0] + rx)))
            for seg in range(2):
                if seg == 0:
                    a0, a1, imgx = max(xlo, 0), min(xhi, dim[0] - 1), 0
                elif xlo < 0:
                    a0, a1, imgx = xlo + dim[0], min(xhi, -1) + dim[0], -1
                else:
                    a0, a1, imgx = max(xlo, dim[0]) - dim[0], xhi - dim[0], 1
                for cx in range(a0, a1 + 1):
                    for j in cells.get((cx, cy, cz), ()):
                        out.append((j, (imgx, imgy, imgz)))
    return out


@pytest.mark.parametrize("ncell,seed", [(5, 0), (5, 1), (6, 2), (7, 3), (11, 4)])
def test_fine_grid_offers_every_pair_inside_r_list_exactly_once(ncell, seed):
    rng = np.random.default_rng(seed)
    r = 1.7
    # box lengths that give exactly ncell (.. ncell + 2) half-width cells per axis
    L = np.array([0.5 * r * (ncell + 0.3), 0.5 * r * (ncell + 1.6), 0.5 * r * (ncell + 2.9)])
    dim = grid_dims(L, r)
    assert dim is not None and dim[0] == ncell
    n = 600
    pos = rng.uniform(-0.5, 0.5, size=(n, 3)) * L
    # adversarial points: on box faces, on cell faces, in the corners
    w = L / np.array(dim)
    pos[:40] = (np.floor(rng.uniform(0, 1, size=(40, 3)) * dim) * w - 0.5 * L)
    pos[40:60] = 0.5 * L.astype(F).astype(np.float64) * rng.choice([-1.0, 1.0], size=(20, 3))  # ON the box faces
    pos[60:70, 0] += L[0]  # one box length outside: binned and swept with their in-box image
    pos = pos.astype(F).astype(np.float64)
    posb = binned_position(pos, L).astype(np.float64)
    assert (np.abs(posb) <= 0.5 * L * (1 + 1e-6)).all() and (posb[60:70, 0] != pos[60:70, 0]).all()
    g, c = grid_coords(pos, L, dim)
    # cell and binned position agree (what makes the image of the stencil the right one)
    lo = c * w - 0.5 * L
    assert ((posb >= lo - 1e-5 * L) & (posb <= lo + w + 1e-5 * L)).all()
    cells = {}
    for j in range(n):
        cells.setdefault(tuple(int(x) for x in c[j]), []).append(j)
    for i in range(n):
        cand = candidates(i, pos, L, dim, r, cells)
        ids = [j for j, _ in cand]
        assert len(ids) == len(set(ids)), "a candidate was offered twice"
        img = dict(cand)
        d = pos[i] - pos
        d -= L * np.round(d / L)
        need = np.nonzero(((d ** 2).sum(axis=1) < r * r) & (np.arange(n) != i))[0]
        missing = [j for j in need if j not in img]
        assert not missing, (i, missing)
        # the image taken from the stencil brings the displacement of the BINNED positions inside
        # r_list like minImage does for the raw ones (a particle on the upper box face is binned
        # in cell 0 and swept one box length lower)
        for j in need:
            dd = posb[i] - posb[j] - L * np.array(img[j])
            assert (dd ** 2).sum() < r * r * (1 + 1e-5), (i, j)


def test_fine_grid_is_refused_when_the_box_is_too_small():
    assert grid_dims(np.array([4.0, 20.0, 20.0]), 1.7) is None  # 4 cells along x
    assert grid_dims(np.array([20.0, 20.0, 20.0]), 1.7) is not None


def test_box_face_particles_of_the_gpu_test_are_binned_where_they_sit():
    """The configuration of test_builder_with_particles_on_and_outside_the_box_faces (L = 9,
    r_list = 2.9, six cells per axis; fp32 1 / 9 rounds up, so -4.5 * Linv + 0.5 is -4e-9 under a
    fused multiply-add): the cell of every particle holds its binned position, and the sweep
    offers every pair inside r_list once with the image that brings the BINNED displacement
    inside. Deriving the row's cell from the already shifted position (a second rounding) put 56
    of these particles in the cell across the box."""
    rng = np.random.default_rng(5)
    Lc, r, n = 9.0, 2.9, 1500
    L = np.array([Lc] * 3)
    dim = grid_dims(L, r)
    assert dim == [6, 6, 6]
    xyz = rng.uniform(-0.5, 0.5, size=(n, 3)) * Lc
    xyz[:30] = 0.5 * np.float64(np.float32(Lc)) * rng.choice([-1.0, 1.0], size=(30, 3))
    xyz[30:60] = np.where(rng.random((30, 3)) < 0.5, 0.5 * np.float64(np.float32(Lc)), xyz[30:60])
    xyz[60:90, 1] += Lc
    xyz[90:120, 2] -= Lc
    pos = xyz.astype(F).astype(np.float64)
    posb = binned_position(pos, L).astype(np.float64)
    g, c = grid_coords(pos, L, dim)
    w = L / np.array(dim)
    lo = c * w - 0.5 * L
    assert ((posb >= lo - 1e-5 * L) & (posb <= lo + w + 1e-5 * L)).all()
    # the old derivation: cell of the shifted position
    _, c_again = grid_coords(posb, L, dim)
    assert (c_again != c).any()
    cells = {}
    for j in range(n):
        cells.setdefault(tuple(int(x) for x in c[j]), []).append(j)
    for i in list(range(150)) + list(range(150, n, 9)):
        cand = candidates(i, pos, L, dim, r, cells)
        ids = [j for j, _ in cand]
        assert len(ids) == len(set(ids))
        img = dict(cand)
        d = pos[i] - pos
        d -= L * np.round(d / L)
        need = np.nonzero(((d ** 2).sum(axis=1) < r * r) & (np.arange(n) != i))[0]
        assert all(j in img for j in need), i
        for j in need:
            dd = posb[i] - posb[j] - L * np.array(img[j])
            assert (dd ** 2).sum() < r * r * (1 + 1e-5), (i, j)


def _library_dims(L, r, periodic=(True, True, True), tilt=(0.0, 0.0, 0.0)):
    import ctypes

    import azplugins_b200 as az
    from azplugins_b200 import _lib

    c = az.Box(L[0], L[1], L[2], *tilt, periodic=periodic).to_c()
    d = (ctypes.c_uint32 * 3)()
    assert _lib.lib.azp_nlist_cell_dim(ctypes.byref(c), float(r), d) == 0
    return [int(x) for x in d]


def test_library_grid_choice_matches_the_model_and_caps_the_cell_count():
    """azp_nlist_cell_dim (host code, no GPU needed): half-width cells where the model says so,
    full-width cells otherwise, never 2 cells on an axis, and at most 2^26 cells in all (a dilute
    system in a huge box must not allocate gigabytes of cell_start)."""
    for L, r in [((126.0, 126.0, 126.0), 3.9), ((298.7, 298.7, 298.7), 2.0), ((9.0, 9.0, 9.0), 2.9),
                 ((20.0, 31.0, 47.0), 1.7), ((8.0, 8.0, 8.0), 2.9), ((7.0, 30.0, 30.0), 2.9)]:
        fine = grid_dims(np.array(L), r)
        got = _library_dims(L, r)
        if fine is not None:
            assert got == fine, (L, r)
        else:
            assert got == [int(x / r) if x / r >= 3.0 else 1 for x in L], (L, r)
    # BASELINE sizes stay on the fine grid
    assert _library_dims((126.0,) * 3, 3.9) == [64, 64, 64]
    assert _library_dims((298.7,) * 3, 2.0) == [298, 298, 298]
    # too many half-width cells -> full-width cells; too many of those -> wider cells
    assert _library_dims((1000.0,) * 3, 3.0) == [333, 333, 333]
    for L, r in [((2000.0,) * 3, 2.0), ((3000.0, 3000.0, 900.0), 1.0), ((5000.0, 5000.0, 10.0), 2.0)]:
        d = _library_dims(L, r)
        assert d[0] * d[1] * d[2] <= 1 << 26 and all(x == 1 or x >= 3 for x in d)
        assert all(x == 1 or Lx / x >= r for x, Lx in zip(d, L))  # no narrower than the cutoff
    # a slab (non-periodic z) and a tilted box never take the fine grid
    assert _library_dims((30.0, 30.0, 30.0), 2.9, periodic=(True, True, False)) == [10, 10, 10]
    assert _library_dims((30.0, 30.0, 30.0), 2.9, tilt=(0.1, 0.0, 0.0))[2] == 10
