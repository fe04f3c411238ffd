"""RectilinearGrid: the host-side description of the staggered C-grid the library works on.

This is plumbing for the Python host (the Julia host passes Oceananigans' own grid through the same
`ocb_grid` descriptor).  Semantics follow Oceananigans' RectilinearGrid as the reference
uses it (test/test_utils.jl:12-27): 1-based interior indices with `halo` cells either side, a Face
location on a Bounded axis has N+1 interior points, `extent=(Lx,Ly,Lz)` puts z in (-Lz, 0), Flat axes
behave as one periodic unit cell.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

Periodic, Bounded, Flat = "Periodic", "Bounded", "Flat"
Center, Face = "c", "f"


def _norm_topo(t):
    t = str(getattr(t, "__name__", t))
    return {"p": Periodic, "b": Bounded, "f": Flat}[t[0].lower()]


class RectilinearGrid:
    def __init__(self, size, x=None, y=None, z=None, extent=None, halo=(3, 3, 3),
                 topology=(Periodic, Periodic, Bounded), dtype=torch.float64, device="cuda"):
        topology = tuple(_norm_topo(t) for t in topology)
        self.user_topology = topology
        nflat = sum(t == Flat for t in topology)

        def expand(val, fill):
            val = list(val) if np.iterable(val) else [val]
            if len(val) == 3 - nflat and nflat:
                it = iter(val)
                val = [fill if t == Flat else next(it) for t in topology]
            if len(val) != 3:
                raise ValueError(f"expected {3 - nflat} or 3 entries, got {len(val)}")
            return val

        size = expand(size, 1)
        halo = expand(halo, 1)
        coords = [x, y, z]
        if extent is not None:
            ext = expand(extent, 1.0)
            coords = [(0.0, float(ext[0])), (0.0, float(ext[1])), (-float(ext[2]), 0.0)]
        self.N = tuple(int(n) for n in size)
        self.H = tuple(max(int(h), 1) for h in halo)
        self.flat = tuple(t == Flat for t in topology)
        self.topology = tuple(Periodic if t == Flat else t for t in topology)
        self.dtype = dtype
        self.np_dtype = np.float64 if dtype == torch.float64 else np.float32
        self.device = torch.device(device)
        self.L, self.uniform, self.spacing0 = [0.0] * 3, [True] * 3, [1.0] * 3
        self._faces, self._centers, self._dc, self._df = [None] * 3, [None] * 3, [None] * 3, [None] * 3
        for d in range(3):
            self._build_axis(d, coords[d])
        self._dev = {}
        self._desc = None

    # ---------------------------------------------------------------------------------------------------
    def _build_axis(self, d, spec):
        N, H, topo = self.N[d], self.H[d], self.topology[d]
        if self.flat[d] or spec is None:
            spec = (0.0, 1.0)
        if callable(spec):
            f_int = np.array([spec(k) for k in range(1, N + 2)], dtype=np.float64)
            uniform = False
        elif isinstance(spec, tuple) and len(spec) == 2:
            f_int = float(spec[0]) + (float(spec[1]) - float(spec[0])) * np.arange(N + 1) / N
            uniform = True
        else:
            f_int = np.asarray(spec, dtype=np.float64)
            if f_int.size != N + 1:
                raise ValueError(f"axis {d + 1}: need {N + 1} face coordinates, got {f_int.size}")
            uniform = False
        L = float(f_int[-1] - f_int[0])
        dc_int = np.full(N, L / N) if uniform else np.diff(f_int)
        if topo == Periodic:       # spacings wrap into the halos
            lo = np.array([dc_int[(-(m + 1)) % N] for m in range(H)][::-1])
            hi = np.array([dc_int[m % N] for m in range(H + 1)])
        else:                      # bounded: the boundary cell's spacing is repeated
            lo, hi = np.full(H, dc_int[0]), np.full(H + 1, dc_int[-1])
        dc = np.concatenate([lo, dc_int, hi])               # centres 1-H .. N+H+1
        faces = np.empty(dc.size + 1)
        faces[H] = f_int[0]
        for m in range(H, dc.size):
            faces[m + 1] = faces[m] + dc[m]
        for m in range(H, 0, -1):
            faces[m - 1] = faces[m] - dc[m - 1]
        if uniform:
            faces = f_int[0] + (L / N) * (np.arange(faces.size) - H)
        else:
            faces[H:H + N + 1] = f_int
        centers = 0.5 * (faces[:-1] + faces[1:])
        df = np.empty(centers.size)
        df[1:] = np.diff(centers)
        df[0] = df[1]
        if uniform:
            df[:] = L / N
        self.L[d], self.uniform[d], self.spacing0[d] = L, uniform, L / N
        # index m of these arrays <-> 1-based index m + 1 - H
        self._faces[d], self._centers[d] = faces.astype(self.np_dtype), centers.astype(self.np_dtype)
        self._dc[d], self._df[d] = dc.astype(self.np_dtype), df.astype(self.np_dtype)

    # ---- host metric access (1-based indices) --------------------------------------------------------------
    def nodes(self, d, loc, idx):
        a = self._centers[d] if loc == Center else self._faces[d]
        return a[np.asarray(idx) + self.H[d] - 1]

    def spacings(self, d, loc, idx):
        a = self._dc[d] if loc == Center else self._df[d]
        return a[np.asarray(idx) + self.H[d] - 1]

    def extent_of(self, loc):
        out = []
        for d in range(3):
            if loc[d] is None:
                out.append(1)
            elif loc[d] == Face and self.topology[d] == Bounded:
                out.append(self.N[d] + 1)
            else:
                out.append(self.N[d])
        return tuple(out)

    def total_volume(self, loc):
        """Σ V over the interior of a field at `loc` (the denominator of `Average`)."""
        tot = 1.0
        for d in range(3):
            n = self.extent_of(loc)[d]
            tot *= float(np.sum(self.spacings(d, loc[d], np.arange(1, n + 1)).astype(np.float64)))
        return tot

    @property
    def z_bottom(self):
        return float(self._faces[2][self.H[2]])

    # ---- device copies and the C descriptor ----------------------------------------------------------------
    def device_array(self, name):
        if name not in self._dev:
            d = {"x": 0, "y": 1, "z": 2}[name[-1]]
            src = {"dc": self._dc, "df": self._df, "c": self._centers, "f": self._faces}[name[:-1]][d]
            self._dev[name] = torch.from_numpy(np.ascontiguousarray(src)).to(self.device)
        return self._dev[name]

    def descriptor(self) -> _lib.ocb_grid:
        if self._desc is None:
            g = _lib.ocb_grid()
            g.dtype = _lib.OCB_F64 if self.dtype == torch.float64 else _lib.OCB_F32
            for d in range(3):
                g.N[d], g.H[d] = self.N[d], self.H[d]
                g.topo[d] = _lib.OCB_PERIODIC if self.topology[d] == Periodic else _lib.OCB_BOUNDED
                g.L[d], g.d[d] = self.L[d], self.spacing0[d]
            g.z_bottom = self.z_bottom
            if not self.uniform[2]:
                # Δzᵃᵃᶜ / zᵃᵃᶜ: Nz + 2Hz entries; Δzᵃᵃᶠ / zᵃᵃᶠ: Nz + 2Hz + 1 (entry m <-> k = m + 1 - Hz)
                g.dzc = self.device_array("dcz").data_ptr()
                g.dzf = self.device_array("dfz").data_ptr()
                g.zc = self.device_array("cz").data_ptr()
                g.zf = self.device_array("fz").data_ptr()
            self._desc = g
        return self._desc

    def sweep_supported(self):
        """The stencil sweep takes regularly spaced x and y (z may be stretched)."""
        return self.uniform[0] and self.uniform[1]

    def __repr__(self):
        return (f"{self.N[0]}×{self.N[1]}×{self.N[2]} RectilinearGrid{{{str(self.dtype).split('.')[-1]}, "
                f"{', '.join(self.user_topology)}}} on {self.device} with {self.H[0]}×{self.H[1]}×{self.H[2]} halo")
