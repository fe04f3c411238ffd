"""What both arms of bench.py (`--impl b200` and `--impl reference`) print about the workload: the same metric, unit and
`config` object, so that the driver's `same_config` comparison holds.  Implementation details (kernel variant, transport,
per-GPU slab) live outside `config`."""
from __future__ import annotations

METRIC = "Gcells/s, fused KE + tracer-variance budget sweep (Float64)"
UNIT = "Gcells/s"
DEFAULT_SIZE = (1024, 1024, 1024)


def workload_config(size, world, fields=False):
    """BASELINE.json configs[2]: the `size` grid as a whole, slab-decomposed along y over `world` GPUs (strong scaling)."""
    nx, ny, nz = size
    return {
        "workload": f"config 3{'B' if fields else 'A'}: {nx}x{ny}x{nz} Float64 (whole job), fused KE budget "
                    "{K, ADV(Centered-2), PR, u_i b_i, eps} + tracer variance {chi, 2c div F}, 7 volume integrals"
                    + (" + 7 output fields" if fields else ""),
        "grid": [nx, ny, nz], "halo": 3, "topology": "Periodic,Periodic,Bounded",
        "decomposition": (f"y-slabs x{world}: {nx}x{ny // world}x{nz} per GPU, halo exchange + global integral all-reduce every step"
                          if world > 1 else "single GPU"),
        "l2": "inputs (5 fields x %.2f GB per GPU) exceed the 126 MB L2; no flush needed" % (nx * (ny // world) * nz * 8 / 1e9),
    }
