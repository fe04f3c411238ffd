"""UMomentumEquation: the terms of the u-momentum equation one by one (reference: src/UMomentumEquation.jl) -- Advection,
BuoyancyAcceleration, CoriolisAcceleration, PressureGradient, ViscousDissipation, TotalViscousDissipation, Forcing, Tendency and
their U-prefixed aliases.  SURVEY.md section 8f row 4 (pass-through modules)."""
from ._momentum_terms import make_module as _make

globals().update(_make("u"))
