"""Simulation box -- the subset of ``hoomd.Box`` the pair path needs (centred, optional tilt)."""

from . import _lib


class Box:
    def __init__(self, Lx, Ly=None, Lz=None, xy=0.0, xz=0.0, yz=0.0, periodic=(True, True, True)):
        self.Lx = float(Lx)
        self.Ly = float(Lx if Ly is None else Ly)
        self.Lz = float(Lx if Lz is None else Lz)
        self.xy, self.xz, self.yz = float(xy), float(xz), float(yz)
        self.periodic = tuple(bool(p) for p in periodic)

    @classmethod
    def cube(cls, L):
        return cls(L, L, L)

    @property
    def L(self):
        return (self.Lx, self.Ly, self.Lz)

    @property
    def volume(self):
        return self.Lx * self.Ly * self.Lz

    def nearest_plane_distance(self):
        """Distances between opposite faces (``hoomd.Box``'s nearest plane distances): what bounds
        the cutoff under the minimum-image convention in a tilted box."""
        t = self.xy * self.yz - self.xz
        return (self.Lx / (1.0 + self.xy ** 2 + t ** 2) ** 0.5, self.Ly / (1.0 + self.yz ** 2) ** 0.5, self.Lz)

    def to_c(self):
        b = _lib.AzpBox()
        for d, v in enumerate(self.L):
            b.L[d] = v
        for d, v in enumerate((self.xy, self.xz, self.yz)):
            b.tilt[d] = v
        for d, v in enumerate(self.periodic):
            b.periodic[d] = int(v)
        return b

    def __repr__(self):
        return "Box(Lx=%g, Ly=%g, Lz=%g, xy=%g, xz=%g, yz=%g)" % (
            self.Lx, self.Ly, self.Lz, self.xy, self.xz, self.yz)
