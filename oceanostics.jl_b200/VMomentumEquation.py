"""VMomentumEquation: the terms of the v-momentum equation one by one (reference: src/VMomentumEquation.jl) -- Advection,
BuoyancyAcceleration, CoriolisAcceleration, PressureGradient, ViscousDissipation, TotalViscousDissipation, Forcing, Tendency and
their V-prefixed aliases.  SURVEY.md section 8f row 4 (pass-through modules)."""
from ._momentum_terms import make_module as _make

globals().update(_make("v"))
