"""oceanostics.jl_b200 -- B200-native (sm_100a) evaluation of the Oceanostics diagnostic hot path.

The directory name carries a dot, so it is imported through `ocb_loader.py` at the repo root
(`import ocb_loader; ocb = ocb_loader.load()`), which registers this package as `oceanostics_b200`.
"""
from . import _lib
from ._lib import OcbError
from .grids import RectilinearGrid, Periodic, Bounded, Flat, Center, Face
from .fields import (Field, CenterField, XFaceField, YFaceField, ZFaceField, ZeroField, ConstantField, BinaryOperation,
                     average_xy)
from .operations import (KernelFunctionOperation, ComputedField, ReducedField, Integral, Average, FusedDiagnostics, Closure, SmagorinskyLilly,
                         SweepOperands, SweepPlan, compute, compute_at)
from .models import (NonhydrostaticModel, HydrostaticFreeSurfaceModel, FPlane, ConstantCartesianCoriolis, BuoyancyTracer,
                     Forcing, get_coriolis_frequency_components)
from . import KineticEnergyEquation, TurbulentKineticEnergyEquation, TracerVarianceEquation, FlowDiagnostics, SpatialFilters, BackgroundPotentialEnergyEquation
from . import FilteredKineticEnergyEquation, SubFilterKineticEnergyEquation, AvailablePotentialEnergyEquation
from . import UMomentumEquation, VMomentumEquation, WMomentumEquation, TracerEquation
from .SpatialFilters import BoxFilter, GaussianFilter

ScalarDiffusivity = Closure
from . import output_writers
from .output_writers import NetCDFWriter, JLD2Writer, TimeInterval, IterationInterval, AveragedTimeInterval
