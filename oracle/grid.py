"""Grid / field containers for the NumPy oracle (test infrastructure only).

Restates the parts of Oceananigans' ``RectilinearGrid`` / ``Field`` the hot
path relies on [OCN-mem: Oceananigans v0.10x Grids/rectilinear_grid.jl,
Fields/field.jl, BoundaryConditions/fill_halo_regions*.jl]:

* 1-based interior indices ``1..N`` with ``H`` halo cells either side;
* a ``Face`` location along a ``Bounded`` axis has ``N + 1`` interior points
  (reference relies on it at src/SpatialFilters/SpatialFilters.jl:152-162);
* spacings ``Δᶜ[i] = xᶠ[i+1] - xᶠ[i]`` (at centres) and ``Δᶠ[i] = xᶜ[i] - xᶜ[i-1]``
  (at faces); uniform axes use the scalar ``L / N``;
* ``Flat`` axes behave as one periodic cell of unit width (δ = 0, ℑ = identity,
  Δ = 1), which reproduces Oceananigans' Flat operators.

Index arrays passed to ``field[i, j, k]`` are NumPy integer arrays that
broadcast against each other (open grids), so one "call" of a kernel function
evaluates every cell at once while the code keeps the per-cell form of the
Julia source.
"""
from __future__ import annotations

import numpy as np

C, F = "c", "f"  # Center / Face


class Off1D:
    """1-D array addressed with 1-based (halo-aware) indices."""

    def __init__(self, data, H):
        self.data = np.asarray(data)
        self.H = H

    def __getitem__(self, idx):
        return self.data[np.asarray(idx) + (self.H - 1)]


class OGrid:
    """Rectilinear grid.  ``topology`` is a 3-tuple of 'P' (Periodic), 'B' (Bounded), 'F' (Flat)."""

    def __init__(self, size, x=None, y=None, z=None, extent=None, halo=(3, 3, 3),
                 topology=("P", "P", "B"), dtype=np.float64):
        self.dtype = np.dtype(dtype)
        topology = tuple(t[0].upper() for t in topology)
        self.user_topology = topology
        size = list(size) if np.iterable(size) else [size]
        halo = list(halo) if np.iterable(halo) else [halo]
        # expand Flat axes: Oceananigans omits them from `size`/`halo`
        nflat = sum(t == "F" for t in topology)
        if len(size) == 3 - nflat and nflat:
            it = iter(size)
            size = [1 if t == "F" else next(it) for t in topology]
        if len(halo) == 3 - nflat and nflat:
            it = iter(halo)
            halo = [1 if t == "F" else next(it) for t in topology]
        coords = [x, y, z]
        if extent is not None:
            ext = list(extent) if np.iterable(extent) else [extent]
            if len(ext) == 3 - nflat and nflat:
                it = iter(ext)
                ext = [1 if t == "F" else next(it) for t in topology]
            coords = [(0.0, float(e)) for e in ext]
            # Oceananigans' default z for extent=... is (-Lz, 0)
            coords[2] = (-float(ext[2]), 0.0)
        self.N = tuple(int(n) for n in size)
        self.H = tuple(int(max(h, 1)) for h in halo)
        # Flat behaves like one periodic unit cell
        self.topology = tuple("P" if t == "F" else t for t in topology)
        self.flat = tuple(t == "F" for t in topology)
        self.L = [0.0, 0.0, 0.0]
        self.uniform = [True, True, True]
        self.faces = [None] * 3   # Off1D over faces  (index 1-H .. N+1+H)
        self.centers = [None] * 3
        self.dc = [None] * 3      # spacing at centres
        self.df = [None] * 3      # spacing at faces
        for d in range(3):
            self._build_axis(d, coords[d])

    # -- construction ---------------------------------------------------------
    def _build_axis(self, d, spec):
        N, H, topo = self.N[d], self.H[d], self.topology[d]
        ft = self.dtype.type
        if self.flat[d] or spec is None:
            spec = (0.0, 1.0)
        if callable(spec):
            f_int = np.array([spec(k) for k in range(1, N + 2)], dtype=np.float64)
            uniform = False
        elif isinstance(spec, tuple) and len(spec) == 2:
            x0, x1 = float(spec[0]), float(spec[1])
            f_int = x0 + (x1 - x0) * np.arange(N + 1) / N
            uniform = True
        else:
            f_int = np.asarray(spec, dtype=np.float64)
            assert f_int.size == N + 1
            uniform = False
        L = f_int[-1] - f_int[0]
        dc_int = np.diff(f_int)
        if uniform:
            dc_int = np.full(N, L / N)
        # extend spacings into halos [OCN-mem: generate_coordinate]: periodic wraps,
        # bounded repeats the boundary cell's spacing
        if topo == "P":
            lo = np.array([dc_int[(-(m + 1)) % N] for m in range(H)][::-1])
            hi = np.array([dc_int[m % N] for m in range(H + 1)])
        else:
            lo = np.full(H, dc_int[0])
            hi = np.full(H + 1, dc_int[-1])
        dc = np.concatenate([lo, dc_int, hi])          # centres 1-H .. N+H+1
        faces = np.empty(dc.size + 1)
        faces[H] = f_int[0]
        for m in range(H, dc.size):
            faces[m + 1] = faces[m] + dc[m]
        for m in range(H, 0, -1):
            faces[m - 1] = faces[m] - dc[m - 1]
        if uniform:  # exact nodes as Oceananigans computes them for regular axes
            faces = f_int[0] + (L / N) * (np.arange(faces.size) - H)
        faces[H:H + N + 1] = f_int if not uniform else faces[H:H + N + 1]
        centers = 0.5 * (faces[:-1] + faces[1:])
        df = np.empty(centers.size)
        df[1:] = np.diff(centers)
        df[0] = df[1]
        if uniform:
            df[:] = L / N
        self.L[d] = ft(L)
        self.uniform[d] = uniform
        self.faces[d] = Off1D(faces.astype(self.dtype), H)
        self.centers[d] = Off1D(centers.astype(self.dtype), H)
        self.dc[d] = Off1D(dc.astype(self.dtype), H)
        self.df[d] = Off1D(df.astype(self.dtype), H)

    # -- sizes ---------------------------------------------------------------
    def extent_of(self, loc):
        """Interior extent of a field at ``loc`` (N, or N+1 for Face x Bounded, 1 for reduced)."""
        out = []
        for d in range(3):
            if loc[d] is None:
                out.append(1)
            elif loc[d] == F and self.topology[d] == "B":
                out.append(self.N[d] + 1)
            else:
                out.append(self.N[d])
        return tuple(out)

    def indices(self, loc):
        """Open-grid 1-based index arrays covering the interior of a field at ``loc``."""
        nx, ny, nz = self.extent_of(loc)
        i = np.arange(1, nx + 1).reshape(-1, 1, 1)
        j = np.arange(1, ny + 1).reshape(1, -1, 1)
        k = np.arange(1, nz + 1).reshape(1, 1, -1)
        return i, j, k

    # -- metrics (Oceananigans Operators/spacings_and_areas_and_volumes.jl) ----
    def spacing(self, d, l, idx):
        return (self.dc[d] if l == C else self.df[d])[idx]

    def node(self, d, l, idx):
        return (self.centers[d] if l == C else self.faces[d])[idx]

    def Dx(self, loc, i, j, k):
        return self.spacing(0, loc[0], i)

    def Dy(self, loc, i, j, k):
        return self.spacing(1, loc[1], j)

    def Dz(self, loc, i, j, k):
        return self.spacing(2, loc[2], k)

    def Ax(self, loc, i, j, k):
        return self.Dy(loc, i, j, k) * self.Dz(loc, i, j, k)

    def Ay(self, loc, i, j, k):
        return self.Dx(loc, i, j, k) * self.Dz(loc, i, j, k)

    def Az(self, loc, i, j, k):
        return self.Dx(loc, i, j, k) * self.Dy(loc, i, j, k)

    def V(self, loc, i, j, k):
        return self.Dx(loc, i, j, k) * self.Dy(loc, i, j, k) * self.Dz(loc, i, j, k)


class OField:
    """Field with halos.  ``loc`` entries are 'c', 'f' or None (reduced axis)."""

    def __init__(self, grid: OGrid, loc=(C, C, C), data=None):
        self.grid = grid
        self.loc = tuple(loc)
        ext = grid.extent_of(self.loc)
        self.ext = ext
        self.H = tuple(0 if self.loc[d] is None else grid.H[d] for d in range(3))
        shape = tuple(ext[d] + 2 * self.H[d] for d in range(3))
        self.data = np.zeros(shape, dtype=grid.dtype) if data is None else data
        assert self.data.shape == shape

    @property
    def interior(self):
        s = tuple(slice(self.H[d], self.H[d] + self.ext[d]) for d in range(3))
        return self.data[s]

    def __getitem__(self, ijk):
        i, j, k = ijk
        idx = []
        for d, a in enumerate((i, j, k)):
            a = np.asarray(a)
            idx.append(np.zeros_like(a) if self.loc[d] is None else a + (self.H[d] - 1))
        return self.data[tuple(idx)]

    def set(self, f):
        """``set!``: f is a number, an array of the interior shape, or f(x, y, z)."""
        g = self.grid
        if callable(f):
            i, j, k = g.indices(self.loc)
            x = g.node(0, self.loc[0] or C, i)
            y = g.node(1, self.loc[1] or C, j)
            z = g.node(2, self.loc[2] or C, k)
            vals = np.vectorize(f, otypes=[np.float64])(*np.broadcast_arrays(x, y, z))
            self.interior[...] = vals
        else:
            self.interior[...] = f
        return self

    # -- fill_halo_regions! ---------------------------------------------------
    def fill_halo(self, top_gradient=None, bottom_gradient=None, impenetrable=True):
        """Default-BC halo fill [OCN-mem]:
        Periodic -> every halo layer wraps; Bounded, Center-located (or tangential) -> the first
        halo layer mirrors the boundary cell (no-flux); Bounded, Face-located (wall-normal) ->
        boundary faces 1 and N+1 are set to 0 (impenetrable), deeper halos untouched.
        ``top_gradient`` / ``bottom_gradient`` apply a GradientBoundaryCondition in z.
        """
        g = self.grid
        for d in range(3):
            if self.loc[d] is None:
                continue
            H, n = self.H[d], self.ext[d]
            a = np.moveaxis(self.data, d, 0)
            if g.topology[d] == "P":
                for m in range(H):
                    a[H - 1 - m] = a[H + n - 1 - m]
                    a[H + n + m] = a[H + m]
            elif self.loc[d] == C:
                a[H - 1] = a[H]
                a[H + n] = a[H + n - 1]
                if d == 2 and bottom_gradient is not None:
                    a[H - 1] = a[H] - bottom_gradient * g.df[2][1]
                if d == 2 and top_gradient is not None:
                    a[H + n] = a[H + n - 1] + top_gradient * g.df[2][g.N[2] + 1]
            elif impenetrable:
                a[H] = 0
                a[H + n - 1] = 0
        return self


class ConstantField:
    """A ``Number`` used where a field is expected (e.g. ``U = 0``)."""

    def __init__(self, value):
        self.value = value

    def __getitem__(self, ijk):
        return self.value


class ZeroField(ConstantField):
    def __init__(self):
        super().__init__(0.0)


class DiffField:
    """Lazy ``a - b`` (Oceananigans BinaryOperation), e.g. ``u - U`` perturbations
    (src/TurbulentKineticEnergyEquation.jl:327-330)."""

    def __init__(self, a, b):
        self.a, self.b = a, b
        self.grid = getattr(a, "grid", None)
        self.loc = getattr(a, "loc", None)

    def __getitem__(self, ijk):
        return self.a[ijk] - self.b[ijk]


def average_dims(field: OField, dims):
    """``Field(Average(field, dims=dims))`` on a regularly spaced grid: the plain mean over the listed
    (1-based) dims, returned as a reduced field whose remaining halos are filled with default BCs."""
    loc = tuple(None if (d + 1) in dims else field.loc[d] for d in range(3))
    out = OField(field.grid, loc)
    out.interior[...] = field.interior.mean(axis=tuple(d - 1 for d in dims), keepdims=True)
    return out.fill_halo(impenetrable=False)


def materialize(grid, loc, values):
    """``Field(op)``: store interior values and fill halos with default BCs."""
    out = OField(grid, loc)
    out.interior[...] = values
    return out.fill_halo(impenetrable=False)
