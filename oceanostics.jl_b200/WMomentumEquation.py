"""WMomentumEquation: the terms of the w-momentum equation one by one (reference: src/WMomentumEquation.jl) -- Advection,
BuoyancyAcceleration, CoriolisAcceleration, PressureGradient, ViscousDissipation, TotalViscousDissipation, Forcing, Tendency and
their W-prefixed aliases.  SURVEY.md section 8f row 4 (pass-through modules)."""
from ._momentum_terms import make_module as _make

globals().update(_make("w"))
