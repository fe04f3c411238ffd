# OceanosticsB200.jl -- the Julia host layer of liboceanostics_b200.so (north_star bullet 1).
#
# Loaded next to Oceanostics, it adds `compute!` methods at the dispatch point Oceanostics itself overrides for its
# filters and the sorted reference height (src/SpatialFilters/SpatialFilters.jl:309,347-378, box_filter.jl:229-230,
# gaussian_filter.jl:533-534, src/BackgroundPotentialEnergyEquation.jl:254-262,610-619): a `Field` whose operand is a
# `CustomKFO{typeof(f)}` (src/Oceanostics.jl:5) for a supported kernel function `f`, and a `Field` whose operand is
# Oceananigans' `Integral` / `Average` reduction over such an operation.  Constructors, `Field(op)`, `compute!`,
# OutputWriters: unchanged.  Every method `ccall`s the C ABI of include/oceanostics_b200.h; nothing is evaluated in
# Julia and there is no CPU route (a CPU-architecture grid keeps Oceananigans' own `compute!`).
#
# Julia is not part of the build image, so this file cannot be executed there; tests/test_julia_layer.py parses it and
# checks the struct mirrors (field order, types, array lengths), the diagnostic table and the bound symbols against the
# header, so the two cannot drift apart.
module OceanosticsB200

using CUDA
using Oceananigans
using Oceananigans: fields
using Oceananigans.Architectures: GPU, architecture
using Oceananigans.Grids: RectilinearGrid, Periodic, Bounded, Flat, Center, Face, topology, halo_size, znode, xspacings, yspacings
using Oceananigans.Fields: Field, ZeroField, AbstractField, location, compute_at!, set_status!, offset_index, instantiated_location
using Oceananigans.AbstractOperations: KernelFunctionOperation, BinaryOperation
using Oceananigans.BoundaryConditions: fill_halo_regions!
using Oceananigans.TurbulenceClosures: ScalarDiffusivity, ThreeDimensionalFormulation
using Oceanostics
using Oceanostics: CustomKFO, get_coriolis_frequency_components
import Oceananigans.Fields: compute!

const lib = get(ENV, "OCEANOSTICS_B200_LIB", "liboceanostics_b200.so")

# ---- constants of include/oceanostics_b200.h ---------------------------------------------------------------------------
const OCB_ABI_VERSION = 2
const OCB_F32, OCB_F64 = Int32(0), Int32(1)
const OCB_PERIODIC, OCB_BOUNDED = Int32(0), Int32(1)
const OCB_CENTER, OCB_FACE, OCB_NOTHING = Int32(0), Int32(1), Int32(2)
const OCB_FIELD_ARRAY, OCB_FIELD_ZERO, OCB_FIELD_CONSTANT = Int32(0), Int32(1), Int32(2)
const OCB_MAX_CLOSURE_FIELDS = 2
const OCB_N_AUX = 7
const OCB_MAX_REQ = 16
const OCB_MAX_FILTER_WIDTH = 64
const OCB_MAX_RANKS = 16
const OCB_RING_SYNC_WORDS = 1024
const OCB_REQ_PERTURBATION = Int32(1)
const OCB_REQ_TENSOR_DIM1, OCB_REQ_TENSOR_DIM2, OCB_REQ_TENSOR_DIM3 = Int32(1 << 4), Int32(1 << 5), Int32(1 << 6)
const OCB_REQ_COLLOCATED = Int32(1 << 7)
const OCB_REQ_CARRIED_HEIGHT = Int32(1 << 8)
const OCB_REQ_HYDROSTATIC_SPLIT = Int32(1 << 9)
const OCB_REQ_CENTERED4 = Int32(1 << 10)
const OCB_REQ_TERM_SHIFT = 12
const OCB_TERM_TENDENCY, OCB_TERM_ADVECTION, OCB_TERM_BUOYANCY, OCB_TERM_CORIOLIS, OCB_TERM_PRESSURE, OCB_TERM_DIFFUSION, OCB_TERM_FORCING = Int32(0), Int32(1), Int32(2), Int32(3), Int32(4), Int32(5), Int32(6)
term_flag(code) = Int32(code) << OCB_REQ_TERM_SHIFT
const OCB_FILTER_BOX, OCB_FILTER_GAUSSIAN, OCB_FILTER_GAUSSIAN_STRETCHED = Int32(0), Int32(1), Int32(2)
const OCB_POLICY_PERIODIC, OCB_POLICY_SHRINK, OCB_POLICY_EDGE, OCB_POLICY_CONSTANT = Int32(0), Int32(1), Int32(2), Int32(3)
const OCB_BPE_THREE_DIMENSIONAL_SORT, OCB_BPE_HEAVISIDE_INTEGRAL = Int32(0), Int32(1)
const OCB_OP_COPY, OCB_OP_ADD, OCB_OP_SUB, OCB_OP_MUL = Int32(0), Int32(1), Int32(2), Int32(3)

# OCB_DIAG_* in the order of the header's enum (0-based ids)
const DIAG_NAMES = (:KE, :ADV, :PR, :WB, :EPS, :EPS_ISO, :TKE, :SPX, :SPY, :SPZ, :SP, :CHI, :CDIFF, :SMOD, :OMOD, :Q,
                    :S11, :S22, :S33, :S12, :S13, :S23, :O12, :O13, :O23, :T11, :T22, :T33, :T11C, :T22C, :T33C,
                    :T12, :T13, :T23, :RI, :RO, :EPV, :TWPV, :DEPV, :VF11, :VF22, :VF33, :VF12, :VF13, :VF23,
                    :FEPS, :SFEPS, :SFKE, :PI, :BPE, :APE, :UPSILON, :APE_EPS, :KE_STRESS, :KE_FORCING, :C_TENDENCY,
                    :KE_TENDENCY, :NU_SMAG, :U_TERM, :V_TERM, :W_TERM, :C_TERM, :QX, :QY, :QZ)
diag_id(name::Symbol) = Int32(findfirst(==(name), DIAG_NAMES) - 1)

# ---- mirrors of the C structs (same field order and types; isbits, so `Ref{T}` passes them by pointer) -----------------
struct OcbGrid
    dtype::Int32
    N::NTuple{3,Int32}
    H::NTuple{3,Int32}
    topo::NTuple{3,Int32}
    L::NTuple{3,Float64}
    d::NTuple{3,Float64}
    z_bottom::Float64
    dzc::CuPtr{Cvoid}
    dzf::CuPtr{Cvoid}
    zc::CuPtr{Cvoid}
    zf::CuPtr{Cvoid}
end

struct OcbField
    data::CuPtr{Cvoid}
    kind::Int32
    loc::NTuple{3,Int32}
    halo::NTuple{3,Int32}
    size::NTuple{3,Int32}
    stride::NTuple{3,Int64}
    value::Float64
end

struct OcbClosure
    nu_const::Float64
    n_nu_fields::Int32
    nu_fields::NTuple{2,OcbField}
    kappa_const::Float64
    n_kappa_fields::Int32
    kappa_fields::NTuple{2,OcbField}
    kappa_scale::NTuple{2,Float64}
end

struct OcbRequest
    diag::Int32
    flags::Int32
    out::CuPtr{Cvoid}
    out_halo::NTuple{3,Int32}
    out_stride::NTuple{3,Int64}
    window_lo::NTuple{3,Int32}
    window_hi::NTuple{3,Int32}
    out_offset::NTuple{3,Int64}
    integral_slot::Int32
end

struct OcbSweepInputs
    u::OcbField
    v::OcbField
    w::OcbField
    b::OcbField
    p::OcbField
    U::OcbField
    V::OcbField
    W::OcbField
    B::OcbField
    aux::NTuple{7,OcbField}
    closure::OcbClosure
    f::NTuple{3,Float64}
    dir::NTuple{3,Float64}
    f_dir::Float64
    ghat::NTuple{3,Float64}
    ro_bg::NTuple{6,Float64}
    les::NTuple{4,Float64}
end

struct OcbFilterPass
    dim::Int32
    kind::Int32
    width::Int32
    policy::Int32
    extent::Int32
    left::Float64
    right::Float64
    weights::NTuple{129,Float64}
    sigma::Float64
    period::Float64
    nodes::CuPtr{Cvoid}
    spacings::CuPtr{Cvoid}
end

struct OcbRing
    rank::Int32
    nranks::Int32
    sync::NTuple{16,CuPtr{Cvoid}}
    timeout_seconds::Float64
end

# ---- status codes -> the exceptions the reference raises -----------------------------------------------------------------
last_error() = unsafe_string(ccall((:ocb_last_error, lib), Cstring, ()))
function check(status::Integer)
    status == 0 && return nothing
    status == 1 && throw(ArgumentError(last_error()))          # OCB_ERR_INVALID: src/Oceanostics.jl:114-120, SpatialFilters.jl:234-271
    status == 2 && throw(UnsupportedByLibrary(last_error()))   # OCB_ERR_UNSUPPORTED: the caller falls back to Oceananigans' compute!
    error("liboceanostics_b200 status $status: " * last_error())
end
struct UnsupportedByLibrary <: Exception
    msg::String
end
function __init__()
    v = ccall((:ocb_abi_version, lib), Cint, ())
    v == OCB_ABI_VERSION || error("liboceanostics_b200.so has ABI version $v; this module binds version $OCB_ABI_VERSION")
    ccall((:ocb_device_count, lib), Cint, ()) > 0 || @warn "no CUDA device: every diagnostic keeps Oceananigans' own compute!"
end
stream_handle() = Ptr{Cvoid}(UInt(CUDA.stream().handle))

# ---- Oceananigans objects -> descriptors ------------------------------------------------------------------------------------
loccode(::Center) = OCB_CENTER
loccode(::Face) = OCB_FACE
loccode(::Nothing) = OCB_NOTHING
topocode(::Type{Periodic}) = OCB_PERIODIC
topocode(::Type{Flat}) = OCB_PERIODIC                     # Flat axes are passed as Periodic with N = 1
topocode(::Type{Bounded}) = OCB_BOUNDED
devptr(a::CuArray) = reinterpret(CuPtr{Cvoid}, pointer(a))
devptr(a::SubArray) = devptr(parent(a))
devptr(a::Base.ReshapedArray) = devptr(parent(a))
devptr(a::OffsetArrays.OffsetArray) = devptr(parent(a))
import OffsetArrays
const NULLPTR = CuPtr{Cvoid}(0)
i32(t) = map(Int32, t)
i64(t) = map(Int64, t)

"A `RectilinearGrid` with regularly spaced x and y (z may be stretched) as the library sees it."
function ocb_grid(grid::RectilinearGrid)
    architecture(grid) isa GPU || throw(UnsupportedByLibrary("the grid lives on the CPU: Oceananigans evaluates it"))
    FT = eltype(grid)
    dx, dy = grid.Δxᶜᵃᵃ, grid.Δyᵃᶜᵃ
    (dx isa Number && dy isa Number) || throw(UnsupportedByLibrary("the stencil sweep needs regularly spaced x and y axes"))
    zreg = grid.z.Δᵃᵃᶜ isa Number
    N, H = size(grid), halo_size(grid)
    return OcbGrid(FT === Float64 ? OCB_F64 : OCB_F32, i32(N), i32(H), map(topocode, topology(grid)),
                   (Float64(grid.Lx), Float64(grid.Ly), Float64(grid.Lz)),
                   (Float64(dx), Float64(dy), zreg ? Float64(grid.z.Δᵃᵃᶜ) : 0.0), Float64(znode(1, grid, Face())),
                   zreg ? NULLPTR : devptr(grid.z.Δᵃᵃᶜ), zreg ? NULLPTR : devptr(grid.z.Δᵃᵃᶠ),
                   zreg ? NULLPTR : devptr(grid.z.cᵃᵃᶜ), zreg ? NULLPTR : devptr(grid.z.cᵃᵃᶠ))
end

const ZERO_FIELD = OcbField(NULLPTR, OCB_FIELD_ZERO, (0, 0, 0), (0, 0, 0), (0, 0, 0), (0, 0, 0), 0.0)
ocb_field(::ZeroField) = ZERO_FIELD
ocb_field(::Nothing) = ZERO_FIELD
ocb_field(x::Number) = x == 0 ? ZERO_FIELD : OcbField(NULLPTR, OCB_FIELD_CONSTANT, (0, 0, 0), (0, 0, 0), (0, 0, 0), (0, 0, 0), Float64(x))
function ocb_field(f::Field)
    p = parent(f)
    loc = map(loccode, instantiated_location(f))
    H = map((h, l) -> l == OCB_NOTHING ? Int32(0) : Int32(h), halo_size(f.grid), loc)   # reduced axes carry no halo
    st = map((s, l) -> l == OCB_NOTHING ? Int64(0) : Int64(s), strides(p), loc)         # ... and broadcast
    return OcbField(devptr(p), OCB_FIELD_ARRAY, loc, H, i32(size(p)), st, 0.0)
end
# anything else (a lazy operation) is materialised first: Field(arg), computed with the rest of the arguments
ocb_field(op) = throw(UnsupportedByLibrary("operand of type $(typeof(op)) must be materialised with Field(...)"))

"`u - U`, the lazy perturbation Oceanostics' constructors build (src/TurbulentKineticEnergyEquation.jl:327-330): (field, mean)."
perturbation_split(x::BinaryOperation) = x.op === (-) ? (x.a, x.b) : throw(UnsupportedByLibrary("only `u - U` is evaluated lazily"))
perturbation_split(x) = (x, nothing)

# closures: isotropic scalar diffusivities and eddy-viscosity closures whose ν_e / κ_e live in closure_fields; tuples add
function ocb_closure(closure, closure_fields, tracer_id::Int = 1)
    nu, ka = Ref(0.0), Ref(0.0)
    nuf, kaf, kas = OcbField[], OcbField[], Float64[]
    add!(c::ScalarDiffusivity{<:Any,ThreeDimensionalFormulation}, K) = begin
        c.ν isa Number || throw(UnsupportedByLibrary("function / field viscosities of ScalarDiffusivity are not supported"))
        nu[] += c.ν
        κ = c.κ isa NamedTuple ? c.κ[tracer_id] : c.κ
        ka[] += κ
    end
    add!(c, K) = begin      # SmagorinskyLilly, AnisotropicMinimumDissipation, DynamicSmagorinsky: K = (νₑ = Field, κₑ = (b = Field, ...))
        hasproperty(K, :νₑ) || throw(ArgumentError("closure $(nameof(typeof(c))) is not an isotropic scalar-diffusivity closure"))
        push!(nuf, ocb_field(K.νₑ))
        if hasproperty(K, :κₑ)
            push!(kaf, ocb_field(K.κₑ[tracer_id])); push!(kas, 1.0)
        elseif hasproperty(c, :Pr)                                 # κ = νₑ / Pr
            Pr = c.Pr isa NamedTuple ? c.Pr[tracer_id] : c.Pr
            push!(kaf, ocb_field(K.νₑ)); push!(kas, 1.0 / Pr)
        end
    end
    if closure isa Tuple
        foreach((c, K) -> add!(c, K), closure, closure_fields)
    elseif closure !== nothing
        add!(closure, closure_fields)
    end
    length(nuf) <= OCB_MAX_CLOSURE_FIELDS && length(kaf) <= OCB_MAX_CLOSURE_FIELDS ||
        throw(UnsupportedByLibrary("at most $OCB_MAX_CLOSURE_FIELDS eddy-viscosity fields per closure tuple"))
    pad(v, fill) = ntuple(i -> i <= length(v) ? v[i] : fill, OCB_MAX_CLOSURE_FIELDS)
    return OcbClosure(nu[], Int32(length(nuf)), pad(nuf, ZERO_FIELD), ka[], Int32(length(kaf)), pad(kaf, ZERO_FIELD), pad(kas, 0.0))
end
const NO_CLOSURE = OcbClosure(0.0, Int32(0), (ZERO_FIELD, ZERO_FIELD), 0.0, Int32(0), (ZERO_FIELD, ZERO_FIELD), (0.0, 0.0))

# ---- operand slots of one diagnostic -------------------------------------------------------------------------------------
"Mutable builder of `OcbSweepInputs`; `fields` collects every Field argument for the `compute_at!` refresh."
mutable struct Slots
    u; v; w; b; p; U; V; W; B
    aux::Vector{Any}
    closure::OcbClosure
    f::NTuple{3,Float64}; dir::NTuple{3,Float64}; f_dir::Float64; ghat::NTuple{3,Float64}; ro_bg::NTuple{6,Float64}; les::NTuple{4,Float64}
    flags::Int32
end
Slots() = Slots(nothing, nothing, nothing, nothing, nothing, nothing, nothing, nothing, nothing, Any[nothing for _ in 1:OCB_N_AUX],
                NO_CLOSURE, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), 0.0, (0.0, 0.0, 1.0), ntuple(_ -> 0.0, 6), (0.0, 0.0, 0.0, 0.0), Int32(0))
function velocities!(s::Slots, u, v, w)
    (s.u, U), (s.v, V), (s.w, W) = perturbation_split(u), perturbation_split(v), perturbation_split(w)
    if U !== nothing || V !== nothing || W !== nothing
        s.U, s.V, s.W = U, V, W
        s.flags |= OCB_REQ_PERTURBATION
    end
    return s
end
function tracer!(s::Slots, c)
    s.b, B = perturbation_split(c)
    if B !== nothing
        s.B = B; s.flags |= OCB_REQ_PERTURBATION
    end
    return s
end
descriptor(s::Slots) = OcbSweepInputs(ocb_field(s.u), ocb_field(s.v), ocb_field(s.w), ocb_field(s.b), ocb_field(s.p), ocb_field(s.U),
                                      ocb_field(s.V), ocb_field(s.W), ocb_field(s.B), ntuple(a -> ocb_field(s.aux[a]), OCB_N_AUX), s.closure,
                                      s.f, s.dir, s.f_dir, s.ghat, s.ro_bg, s.les)
field_arguments(s::Slots) = filter(x -> x isa Field, Any[s.u, s.v, s.w, s.b, s.p, s.U, s.V, s.W, s.B, s.aux...])
ghat(buoyancy) = buoyancy === nothing ? (0.0, 0.0, 1.0) : map(x -> -Float64(x), Tuple(buoyancy.gravity_unit_vector))   # ĝ = -gravity unit vector
tracer_index(::Val{id}) where id = id

# ---- the diagnostic table: kernel function (or kernel type) => (OCB_DIAG name, arguments -> Slots) ------------------------
# One row per kernel function of SURVEY.md section 8(a)/(f); `args` is `op.arguments` exactly as the reference's constructor
# passes them (file:line of the constructor in the comment).
const KEq, TKEq, TVEq, FD = Oceanostics.KineticEnergyEquation, Oceanostics.TurbulentKineticEnergyEquation, Oceanostics.TracerVarianceEquation, Oceanostics.FlowDiagnostics
const FKEq, SFKEq = Oceanostics.FilteredKineticEnergyEquation, Oceanostics.SubFilterKineticEnergyEquation
const BPEq, APEq = Oceanostics.BackgroundPotentialEnergyEquation, Oceanostics.AvailablePotentialEnergyEquation
const TC = Oceananigans.TurbulenceClosures

vel3(a) = velocities!(Slots(), a[1], a[2], a[3])
velnt(v) = velocities!(Slots(), v.u, v.v, v.w)
function with_closure(s::Slots, closure, K, id = 1)
    s.closure = ocb_closure(closure, K, id)
    return s
end
# viscous_flux_* / ∂ⱼ_τᵢⱼ argument list (closure, closure_fields, clock, model_fields, buoyancy)
flux_slots(a) = with_closure(velnt(a[4]), a[1], a[2])
# c∇_dot_qᶜ / χ argument list (closure, closure_fields, Val(id), c, clock, model_fields, buoyancy)
tracer_flux_slots(a) = with_closure(tracer!(velnt(a[6]), a[4]), a[1], a[2], tracer_index(a[3]))
function sp_slots(a)          # (u′, v′, w′, U, V, W): u′ is a Field or the lazy u - U (TurbulentKineticEnergyEquation.jl:147-159)
    s = vel3(a)
    s.U === nothing && (s.U = a[4]; s.V = a[5]; s.W = a[6])
    return s
end
function tensor_pair(a, I, J)  # StrainRate / Vorticity / Stress kernels take (uᵢ,) or (uᵢ, uⱼ)
    s = Slots()
    setproperty!(s, (:u, :v, :w)[I], a[1])
    I != J && setproperty!(s, (:u, :v, :w)[J], a[2])
    return s
end

const DIAG = Dict{Any,Tuple{Symbol,Function}}(
    KEq.kinetic_energy_ccc                       => (:KE,      vel3),                                     # KineticEnergyEquation.jl:45
    KEq.uᵢ∂ⱼuⱼuᵢᶜᶜᶜ                             => (:ADV,     a -> advected!(velnt(a[1]), a[2])),            # :186 (velocities, advection)
    KEq.uᵢ∂ᵢpᶜᶜᶜ                                => (:PR,      a -> (s = velnt(a[1]); s.p = materialized(a[2]); s)),  # :357 (velocities, pressure)
    KEq.uᵢbᵢᶜᶜᶜ                                 => (:WB,      a -> (s = velnt(a[1]); s.b = a[3].b; s.ghat = ghat(a[2]); s)),  # :421
    KEq.viscous_dissipation_rate_ccc             => (:EPS,     a -> with_closure(velnt(a[2]), a[3].closure, a[1])),  # :491 (closure_fields, fields, parameters)
    KEq.isotropic_viscous_dissipation_rate_ccc   => (:EPS_ISO, a -> with_closure(vel3(a), a[4].closure, a[4].closure_fields)),  # :545
    TKEq.turbulent_kinetic_energy_ccc            => (:TKE,     a -> (s = vel3(a); s.U, s.V, s.W = a[4], a[5], a[6]; s.flags |= OCB_REQ_PERTURBATION; s)),  # TurbulentKineticEnergyEquation.jl:56
    TKEq.shear_production_rate_x_ccc             => (:SPX,     sp_slots),                                 # :151
    TKEq.shear_production_rate_y_ccc             => (:SPY,     sp_slots),                                 # :213
    TKEq.shear_production_rate_z_ccc             => (:SPZ,     sp_slots),                                 # :274
    TKEq.shear_production_rate_ccc               => (:SP,      sp_slots),                                 # :323
    TVEq.tracer_variance_dissipation_rate_ccc    => (:CHI,     tracer_flux_slots),                        # TracerVarianceEquation.jl:218
    TVEq.c∇_dot_qᶜ                              => (:CDIFF,   tracer_flux_slots),                        # :140
    FD.strain_rate_tensor_modulus_ccc            => (:SMOD,    vel3),                                     # FlowDiagnostics.jl:477
    FD.vorticity_tensor_modulus_ccc              => (:OMOD,    vel3),                                     # :623
    FD.Q_velocity_gradient_tensor_invariant_ccc  => (:Q,       vel3),                                     # :951
    FD.StrainRateTensorKernel{1,1}               => (:S11,     a -> tensor_pair(a, 1, 1)),                # :562-567
    FD.StrainRateTensorKernel{2,2}               => (:S22,     a -> tensor_pair(a, 2, 2)),
    FD.StrainRateTensorKernel{3,3}               => (:S33,     a -> tensor_pair(a, 3, 3)),
    FD.StrainRateTensorKernel{1,2}               => (:S12,     a -> tensor_pair(a, 1, 2)),
    FD.StrainRateTensorKernel{1,3}               => (:S13,     a -> tensor_pair(a, 1, 3)),
    FD.StrainRateTensorKernel{2,3}               => (:S23,     a -> tensor_pair(a, 2, 3)),
    FD.VorticityTensorKernel{1,2}                => (:O12,     a -> tensor_pair(a, 1, 2)),                # :699-701
    FD.VorticityTensorKernel{1,3}                => (:O13,     a -> tensor_pair(a, 1, 3)),
    FD.VorticityTensorKernel{2,3}                => (:O23,     a -> tensor_pair(a, 2, 3)),
    FD.StressTensorKernel{1,1,false}             => (:T11,     a -> tensor_pair(a, 1, 1)),                # :834-847
    FD.StressTensorKernel{2,2,false}             => (:T22,     a -> tensor_pair(a, 2, 2)),
    FD.StressTensorKernel{3,3,false}             => (:T33,     a -> tensor_pair(a, 3, 3)),
    FD.StressTensorKernel{1,1,true}              => (:T11C,    a -> tensor_pair(a, 1, 1)),
    FD.StressTensorKernel{2,2,true}              => (:T22C,    a -> tensor_pair(a, 2, 2)),
    FD.StressTensorKernel{3,3,true}              => (:T33C,    a -> tensor_pair(a, 3, 3)),
    FD.StressTensorKernel{1,2,false}             => (:T12,     a -> tensor_pair(a, 1, 2)),
    FD.StressTensorKernel{1,3,false}             => (:T13,     a -> tensor_pair(a, 1, 3)),
    FD.StressTensorKernel{2,3,false}             => (:T23,     a -> tensor_pair(a, 2, 3)),
    FD.richardson_number_ccf                     => (:RI,      a -> (s = vel3(a); s.b = a[4]; s.dir = map(Float64, Tuple(a[5])); s)),  # :118
    FD.rossby_number_fff                         => (:RO,      a -> (s = vel3(a); q = a[4]; s.f = (q.fx, q.fy, q.fz);
                                                                      s.ro_bg = (q.dWdy_bg, q.dVdz_bg, q.dUdz_bg, q.dWdx_bg, q.dUdy_bg, q.dVdx_bg); s)),  # :194
    FD.ertel_potential_vorticity_fff             => (:EPV,     a -> (s = vel3(a); s.b = a[4]; s.f = (a[5], a[6], a[7]); s)),  # :346
    FD.potential_vorticity_in_thermal_wind_fff   => (:TWPV,    a -> (s = Slots(); s.u, s.v, s.b = a[1], a[2], a[3]; s.f = (0.0, 0.0, a[4]); s)),  # :343
    FD.directional_ertel_potential_vorticity_fff => (:DEPV,    a -> (s = vel3(a); s.b = a[4]; q = a[5]; s.dir = (q.dir_x, q.dir_y, q.dir_z); s.f_dir = q.f_dir; s)),  # :423
    TC.viscous_flux_ux                           => (:VF11,    flux_slots),                               # FilteredKineticEnergyEquation.jl:377-387
    TC.viscous_flux_vy                           => (:VF22,    flux_slots),
    TC.viscous_flux_wz                           => (:VF33,    flux_slots),
    TC.viscous_flux_uy                           => (:VF12,    flux_slots),                               # = viscous_flux_vx
    TC.viscous_flux_vx                           => (:VF12,    flux_slots),
    TC.viscous_flux_uz                           => (:VF13,    flux_slots),                               # = viscous_flux_wx
    TC.viscous_flux_wx                           => (:VF13,    flux_slots),
    TC.viscous_flux_vz                           => (:VF23,    flux_slots),                               # = viscous_flux_wy
    TC.viscous_flux_wy                           => (:VF23,    flux_slots),
    FKEq.filtered_kinetic_energy_ccc             => (:KE,      vel3),                                     # :101 (the same kernel on ū, v̄, w̄)
    FKEq.filtered_dissipation_rate_ccc           => (:FEPS,    a -> (s = velnt(a[1]); ff = a[2];          # :388 (fv, ff): aux 0..5 = τ̄₁₁ τ̄₂₂ τ̄₃₃ τ̄₁₂ τ̄₁₃ τ̄₂₃
                                                                      s.aux[1:6] = [ff.τ̄₁₁, ff.τ̄₂₂, ff.τ̄₃₃, ff.τ̄₁₂, ff.τ̄₁₃, ff.τ̄₂₃]; s)),
    SFKEq.subfilter_kinetic_energy_ccc           => (:SFKE,    a -> (s = velocities!(Slots(), a[2], a[3], a[4]); s.aux[7] = a[1]; s)),  # SubFilterKineticEnergyEquation.jl:81
    BPEq.minus_bz✶_ccc                           => (:BPE,     a -> (s = Slots(); s.b = a[1]; s.aux[7] = a[2]; s)),                      # BackgroundPotentialEnergyEquation.jl:906
    APEq.local_ape_ccc                           => (:APE,     a -> (s = Slots(); s.aux[6] = a[1]; s.b = a[2]; s.aux[7] = a[end];        # AvailablePotentialEnergyEquation.jl:110-111
                                                                      length(a) == 4 && (s.aux[4] = a[3]; s.flags |= OCB_REQ_CARRIED_HEIGHT); s)),
    APEq.upsilon_ccc                             => (:UPSILON, a -> (s = Slots(); s.aux[7] = a[1]; s)),                                   # :190
    APEq.ape_dissipation_rate_ccc                => (:APE_EPS, a -> (s = tracer_flux_slots(a[2:end]); s.aux[5] = a[1]; s)),               # :298 (Υ, flux arguments...)
    KEq.uᵢ∂ⱼ_τᵢⱼᶜᶜᶜ                             => (:KE_STRESS,  flux_slots),                            # KineticEnergyEquation.jl:246
    KEq.uᵢFᵤᵢᶜᶜᶜ                                => (:KE_FORCING, a -> (s = velnt(a[3]); F = a[1]; clock = a[2];                          # :299 (forcing, clock, fields)
                                                                         s.aux[1:3] = [forcing_field(F.u, a[3].u, clock, a[3]), forcing_field(F.v, a[3].v, clock, a[3]),
                                                                                       forcing_field(F.w, a[3].w, clock, a[3])]; s)),
    TVEq.c∂ₜcᶜᶜᶜ                                => (:C_TENDENCY, tendency_slots_tracer),                 # TracerVarianceEquation.jl:88
    KEq.uᵢGᵢᶜᶜᶜ                                 => (:KE_TENDENCY, tendency_slots_momentum),              # KineticEnergyEquation.jl:142
)
# ---- pass-through modules (UMomentumEquation.jl, VMomentumEquation.jl, WMomentumEquation.jl, TracerEquation.jl): one term each --
const UEq, VEq, WEq, TrEq = Oceanostics.UMomentumEquation, Oceanostics.VMomentumEquation, Oceanostics.WMomentumEquation, Oceanostics.TracerEquation
const NHM, ADVM, BF, COR = Oceananigans.Models.NonhydrostaticModels, Oceananigans.Advection, Oceananigans.BuoyancyFormulations, Oceananigans.Coriolis
term!(s::Slots, code) = (s.flags |= term_flag(code); s)
adv_slots(a) = term!(advected!(velnt(a[2]), a[1]), OCB_TERM_ADVECTION)                                    # (advection, velocities, uᵢ)
buoy_slots(a) = (s = Slots(); s.b = a[2].b; s.ghat = ghat(a[1]); term!(s, OCB_TERM_BUOYANCY))             # (buoyancy, tracers)
cor_slots(a) = (s = velnt(a[2]); s.f = a[1] === nothing ? (0.0, 0.0, 0.0) : map(Float64, get_coriolis_frequency_components(a[1])); term!(s, OCB_TERM_CORIOLIS))
pgrad_slots(a) = (s = Slots(); s.p = a[1]; term!(s, OCB_TERM_PRESSURE))                                   # (hydrostatic_pressure,)
visc_slots(a) = term!(flux_slots(a), OCB_TERM_DIFFUSION)                                                  # (closure, diffusivities, clock, model_fields, buoyancy)
force_slots(axis) = a -> (s = Slots(); s.aux[axis] = forcing_field(a[1], getproperty(a[3], (:u, :v, :w)[axis]), a[2], a[3]); term!(s, OCB_TERM_FORCING))
function velocity_tendency_slots(axis)      # (advection, coriolis, stokes_drift, closure, immersed_bc, buoyancy, background_fields, velocities, tracers, aux, K, pHY′, clock, forcing)
    return a -> begin
        advection, coriolis, stokes, closure, ibc, buoyancy, background, velocities, tracers, aux, K, pHY, clock, forcing = a
        (ibc === nothing && stokes === nothing && all(isnothing, values(background.velocities))) ||
            throw(UnsupportedByLibrary("velocity tendency with immersed boundaries / Stokes drift / background fields"))
        s = advected!(with_closure(velnt(velocities), closure, K), advection)
        s.f = coriolis === nothing ? (0.0, 0.0, 0.0) : map(Float64, get_coriolis_frequency_components(coriolis))
        buoyancy === nothing || (s.b = tracers.b; s.ghat = ghat(buoyancy))
        pHY === nothing || (s.p = pHY; s.flags |= OCB_REQ_HYDROSTATIC_SPLIT)
        s.aux[axis] = forcing_field(forcing, getproperty(velocities, (:u, :v, :w)[axis]), clock, merge(velocities, tracers))
        term!(s, OCB_TERM_TENDENCY)
    end
end
for (row, key) in ((:U_TERM, 1), (:V_TERM, 2), (:W_TERM, 3))
    Eq = (UEq, VEq, WEq)[key]
    DIAG[(ADVM.div_𝐯u, ADVM.div_𝐯v, ADVM.div_𝐯w)[key]] = (row, adv_slots)                                    # UMomentumEquation.jl:99-101
    DIAG[(BF.x_dot_g_bᶠᶜᶜ, BF.y_dot_g_bᶜᶠᶜ, BF.z_dot_g_bᶜᶜᶠ)[key]] = (row, buoy_slots)                      # :138
    DIAG[(COR.x_f_cross_U, COR.y_f_cross_U, COR.z_f_cross_U)[key]] = (row, cor_slots)                          # :171
    key < 3 && (DIAG[(UEq.hydrostatic_pressure_gradient_x, VEq.hydrostatic_pressure_gradient_y)[key]] = (row, pgrad_slots))   # :204
    DIAG[(TC.∂ⱼ_τ₁ⱼ, TC.∂ⱼ_τ₂ⱼ, TC.∂ⱼ_τ₃ⱼ)[key]] = (row, visc_slots)                                         # :285
    DIAG[(UEq.forcing_fcc, VEq.forcing_cfc, WEq.forcing_ccf)[key]] = (row, force_slots(key))                   # :453
    DIAG[(NHM.u_velocity_tendency, NHM.v_velocity_tendency, NHM.w_velocity_tendency)[key]] = (row, velocity_tendency_slots(key))   # :499
end
DIAG[ADVM.div_Uc] = (:C_TERM, a -> term!(advected!(tracer!(velnt(a[2]), a[3]), a[1]), OCB_TERM_ADVECTION))                          # TracerEquation.jl:72
DIAG[TC.∇_dot_qᶜ] = (:C_TERM, a -> term!(with_closure(tracer!(velnt(a[6]), a[4]), a[1], a[2], tracer_index(a[3])), OCB_TERM_DIFFUSION))  # :114
DIAG[TrEq.forcing_ccc] = (:C_TERM, a -> (s = Slots(); s.aux[1] = forcing_field(a[1], a[3].b, a[2], a[3]); term!(s, OCB_TERM_FORCING)))   # :324
tracer_flux_row(a) = with_closure(tracer!(Slots(), a[4]), a[1], a[2], tracer_index(a[3]))                 # (closure, closure_fields, Val(id), c, clock, fields, buoyancy)
DIAG[TC.diffusive_flux_x] = (:QX, tracer_flux_row)                                                        # :223
DIAG[TC.diffusive_flux_y] = (:QY, tracer_flux_row)                                                        # :257
DIAG[TC.diffusive_flux_z] = (:QZ, tracer_flux_row)                                                        # :291

# SFEPS, PI: the reference wraps an operation TREE in a one-argument kernel (`εˢ[i,j,k]`, `Πᵏ[i,j,k]`:
# SubFilterKineticEnergyEquation.jl:143-147, FilteredKineticEnergyEquation.jl:246-257); `tree_slots` walks that tree's leaves.
DIAG[SFKEq.subfilter_ke_dissipation_rate_ccc] = (:SFEPS, a -> subfilter_dissipation_slots(a[1]))
DIAG[FKEq.cross_scale_ke_flux_ccc] = (:PI, a -> cross_scale_flux_slots(a[1]))

"Centered(order = 2 | 4) is what the library restates (the reference's examples all run Centered(order = 4)); any other scheme keeps Oceananigans' own `compute!` (UnsupportedByLibrary)."
function advection_supported(adv)
    (adv isa Oceananigans.Advection.Centered && Oceananigans.Advection.required_halo_size_x(adv) <= 2) ||
        throw(UnsupportedByLibrary("advection scheme $(summary(adv)): only Centered(order = 2 | 4) is evaluated by the library"))
    return Oceananigans.Advection.required_halo_size_x(adv) == 2 ? OCB_REQ_CENTERED4 : Int32(0)      # the request flag
end
advected!(s::Slots, adv) = (s.flags |= advection_supported(adv); s)

"A lazy argument (`sum(model.pressures)`, `u - U` handed to a diagnostic that takes no perturbation flag) is materialised once and refreshed with the other arguments."
materialized(x::Field) = x
materialized(x) = get!(() -> Field(x), MATERIALIZED, objectid(x))
const MATERIALIZED = IdDict{UInt,Any}()

"`model.forcing.u` evaluated at the points of the field it forces (Oceananigans runs the functor inside its kernels): a Field the library contracts with the velocity."
function forcing_field(F, forced::Field, clock, model_fields)
    F === nothing && return nothing
    F isa Number && return F
    LX, LY, LZ = location(forced)
    return materialized(KernelFunctionOperation{LX,LY,LZ}((i, j, k, grid, F, clock, mf) -> F(i, j, k, grid, clock, mf), forced.grid, F, clock, model_fields))
end

# c∂ₜcᶜᶜᶜ (TracerVarianceEquation.jl:70-88) and uᵢGᵢᶜᶜᶜ (KineticEnergyEquation.jl:120-142): NonhydrostaticModel without
# background fields, immersed boundaries, Stokes drift or biogeochemistry; Centered(order = 2 | 4) advection
function tendency_slots_tracer(a)
    id, name, advection, closure, immersed, buoyancy, bgc, background, velocities, tracers, aux, K, clock, forcing = a
    (immersed === nothing && bgc === nothing && all(isnothing, values(background.tracers))) ||
        throw(UnsupportedByLibrary("tracer tendency with immersed boundaries / biogeochemistry / background fields"))
    c = tracers[tracer_index(name isa Val ? id : id)]
    s = advected!(with_closure(tracer!(velnt(velocities), c), closure, K, tracer_index(id)), advection)
    s.aux[1] = forcing_field(forcing, c, clock, merge(velocities, tracers))
    return s
end
function tendency_slots_momentum(a)
    advection, coriolis, stokes, closure, ibu, ibv, ibw, buoyancy, background, velocities, tracers, aux, K, pHY, clock, forcing = a
    (ibu === nothing && ibv === nothing && ibw === nothing && stokes === nothing && all(isnothing, values(background.velocities))) ||
        throw(UnsupportedByLibrary("momentum tendency with immersed boundaries / Stokes drift / background fields"))
    s = advected!(with_closure(velnt(velocities), closure, K), advection)
    s.f = coriolis === nothing ? (0.0, 0.0, 0.0) : map(Float64, get_coriolis_frequency_components(coriolis))
    if buoyancy !== nothing
        s.b = tracers.b; s.ghat = ghat(buoyancy)
    end
    if pHY !== nothing                                                           # hydrostatic pressure anomaly: ∇ₕ pHY′ in u, v; no buoyancy in w
        s.p = pHY; s.flags |= OCB_REQ_HYDROSTATIC_SPLIT
    end
    mf = merge(velocities, tracers)
    s.aux[1:3] = [forcing_field(forcing.u, velocities.u, clock, mf), forcing_field(forcing.v, velocities.v, clock, mf),
                  forcing_field(forcing.w, velocities.w, clock, mf)]
    return s
end

# εˢ = Field(filter(Field(ε))) - εˡ  and  Πᵏ = -Σ w to_center(τⁱʲ) to_center(S̄ⁱʲ): the leaves of the wrapped trees
function subfilter_dissipation_slots(tree::BinaryOperation)      # (filtered ε Field) - (εˡ KFO)
    s = DIAG[FKEq.filtered_dissipation_rate_ccc][2](tree.b.arguments)
    s.aux[7] = tree.a
    return s
end
function cross_scale_flux_slots(tree)
    leaves = collect_cross_scale_leaves(tree)                    # (ū, v̄, w̄, filtered uⁱuʲ Fields by component, dims, collocated)
    s = velocities!(Slots(), leaves.u, leaves.v, leaves.w)
    for (slot, key) in enumerate((:τ₁₁, :τ₂₂, :τ₃₃, :τ₁₂, :τ₁₃, :τ₂₃))
        haskey(leaves.filtered_flux, key) && (s.aux[slot] = leaves.filtered_flux[key])
    end
    1 in leaves.dims && (s.flags |= OCB_REQ_TENSOR_DIM1)
    2 in leaves.dims && (s.flags |= OCB_REQ_TENSOR_DIM2)
    3 in leaves.dims && (s.flags |= OCB_REQ_TENSOR_DIM3)
    leaves.collocated && (s.flags |= OCB_REQ_COLLOCATED)
    return s
end
"Walk -Σ weight * to_center(Field(filter(Field(uⁱuʲ))) - ūⁱūʲ) * to_center(S̄ⁱʲ) down to its Fields (FilteredKineticEnergyEquation.jl:42-48,176-188)."
function collect_cross_scale_leaves(tree)
    fluxes, vel, dims = Dict{Symbol,Any}(), Dict{Symbol,Any}(), Set{Int}()
    collocated = Ref(false)
    visit(x::Field) = nothing
    function visit(x::BinaryOperation)
        if x.op === (-) && x.a isa Field && x.b isa KernelFunctionOperation && x.b.kernel_function isa FD.StressTensorKernel
            I, J, C = typeof(x.b.kernel_function).parameters
            fluxes[Symbol("τ", "₁₂₃"[nextind("₁₂₃", 0, I)], "₁₂₃"[nextind("₁₂₃", 0, J)])] = x.a
            push!(dims, I, J); collocated[] |= C
            vel[(:u, :v, :w)[I]] = x.b.arguments[1]
            I != J && (vel[(:u, :v, :w)[J]] = x.b.arguments[2])
        else
            visit(x.a); visit(x.b)
        end
    end
    visit(x::KernelFunctionOperation) = foreach(visit, x.arguments)
    visit(x) = hasproperty(x, :args) ? foreach(visit, x.args) : hasproperty(x, :arg) ? visit(x.arg) : hasproperty(x, :operand) ? visit(x.operand) : nothing
    visit(tree)
    return (u = get(vel, :u, nothing), v = get(vel, :v, nothing), w = get(vel, :w, nothing), filtered_flux = fluxes, dims = dims, collocated = collocated[])
end

# ---- plans -----------------------------------------------------------------------------------------------------------
mutable struct Plan
    handle::Ptr{Cvoid}
    integrals::CuVector{Float64}
    slots::Slots
    n_slots::Int
    function Plan(handle, integrals, slots, n)
        p = new(handle, integrals, slots, n)
        finalizer(q -> ccall((:ocb_plan_destroy, lib), Cint, (Ptr{Cvoid},), q.handle), p)
        return p
    end
end
const PLANS = IdDict{Any,Plan}()

kernel_key(op::KernelFunctionOperation) = op.kernel_function isa Function ? op.kernel_function : typeof(op.kernel_function)
supported(op) = op isa KernelFunctionOperation && haskey(DIAG, kernel_key(op)) && op.grid isa RectilinearGrid && architecture(op.grid) isa GPU

"The destination's index window, 1-based inclusive (SpatialFilters.jl:322-330: `KernelParameters(size(dest), offset_index.(dest.indices))`)."
function window(dest::Field)
    lo = ntuple(d -> dest.indices[d] isa Colon ? Int32(0) : Int32(first(dest.indices[d])), 3)
    hi = ntuple(d -> dest.indices[d] isa Colon ? Int32(0) : Int32(last(dest.indices[d])), 3)
    off = ntuple(d -> dest.indices[d] isa Colon ? Int64(1) : Int64(first(dest.indices[d])), 3)
    return lo, hi, off
end
function request(op, dest::Union{Field,Nothing}, slot::Integer, flags::Int32, reduction_window = nothing)
    id = diag_id(DIAG[kernel_key(op)][1])
    if dest === nothing
        lo, hi = reduction_window === nothing ? ((Int32(0), Int32(0), Int32(0)), (Int32(0), Int32(0), Int32(0))) : reduction_window
        return OcbRequest(id, flags, NULLPTR, (0, 0, 0), (0, 0, 0), lo, hi, (1, 1, 1), Int32(slot))
    end
    lo, hi, off = window(dest)
    p = parent(dest)
    return OcbRequest(id, flags, devptr(p), i32(halo_size(dest.grid)), i64(strides(p)), lo, hi, off, Int32(slot))
end

"One plan for `members` = [(operation, destination Field or nothing, integral slot or -1, reduction window)]; their operands must agree."
function create_plan(members)
    op1 = members[1][1]
    slots = merged_slots([DIAG[kernel_key(m[1])][2](m[1].arguments) for m in members])
    reqs = [request(m[1], m[2], m[3], slot_flags(m[1]), length(m) > 3 ? m[4] : nothing) for m in members]
    n_slots = 1 + maximum(m[3] for m in members)
    handle = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ocb_plan_create, lib), Cint, (Ref{OcbGrid}, Ref{OcbSweepInputs}, Ptr{OcbRequest}, Int32, Ref{Ptr{Cvoid}}),
                ocb_grid(op1.grid), descriptor(slots), reqs, Int32(length(reqs)), handle))
    return Plan(handle[], CUDA.zeros(Float64, max(n_slots, 1)), slots, n_slots)
end
slot_flags(op) = DIAG[kernel_key(op)][2](op.arguments).flags
"Union of the operand sets of sibling diagnostics; a conflict (two different `u`) means they cannot share a sweep."
function merged_slots(all)
    out = all[1]
    for s in all[2:end], name in (:u, :v, :w, :b, :p, :U, :V, :W, :B)
        a, b = getproperty(out, name), getproperty(s, name)
        (a === nothing || b === nothing || a === b) || throw(UnsupportedByLibrary("sibling diagnostics read different `$name` operands"))
        a === nothing && setproperty!(out, name, b)
    end
    for s in all[2:end]
        for a in 1:OCB_N_AUX
            (out.aux[a] === nothing || s.aux[a] === nothing || out.aux[a] === s.aux[a]) || throw(UnsupportedByLibrary("sibling diagnostics read different aux[$a]"))
            out.aux[a] === nothing && (out.aux[a] = s.aux[a])
        end
        s.closure == NO_CLOSURE || (out.closure = s.closure)
        s.f == (0.0, 0.0, 0.0) || (out.f = s.f)
        s.ro_bg == ntuple(_ -> 0.0, 6) || (out.ro_bg = s.ro_bg)
        s.dir == (0.0, 0.0, 1.0) || (out.dir = s.dir)
        s.ghat == (0.0, 0.0, 1.0) || (out.ghat = s.ghat)
        s.f_dir == 0.0 || (out.f_dir = s.f_dir)
    end
    return out
end
function execute!(plan::Plan, time)
    foreach(a -> compute_at!(a, time), field_arguments(plan.slots))                       # (1) refresh every Field argument
    check(ccall((:ocb_plan_execute, lib), Cint, (Ptr{Cvoid}, CuPtr{Float64}, Ptr{Cvoid}), plan.handle, plan.integrals, stream_handle()))
end
plan_info(plan::Plan) = begin
    n, v, b = Ref{Int32}(0), Ref{Int32}(0), Ref{Float64}(0.0)
    check(ccall((:ocb_plan_info, lib), Cint, (Ptr{Cvoid}, Ref{Int32}, Ref{Int32}, Ref{Float64}), plan.handle, n, v, b))
    (launches = n[], variant = v[], algorithmic_bytes = b[])
end

"`fill_halo_regions!(comp)` on the device (default boundary conditions: periodic wrap, no-flux mirror, impenetrable faces)."
function fill_halo!(f::Field; impenetrable = false)
    check(ccall((:ocb_fill_halo, lib), Cint, (Ref{OcbGrid}, Ref{OcbField}, Int32, Float64, Float64, Ptr{Cvoid}),
                ocb_grid(f.grid), ocb_field(f), Int32(impenetrable), NaN, NaN, stream_handle()))
    return f
end

# ---- compute! for Field(op): the five obligations of the reference's own overrides (SpatialFilters.jl:347-378) ------------
const ComputedKFOField = Field{<:Any,<:Any,<:Any,<:CustomKFO}
function compute!(comp::ComputedKFOField, time = nothing)
    group = get(GROUP_OF, comp, nothing)
    group === nothing || return (compute_group!(group, time); comp)                          # delegation to the sibling registry
    supported(comp.operand) || return invoke(compute!, Tuple{Field{<:Any,<:Any,<:Any,<:KernelFunctionOperation},Any}, comp, time)
    try
        plan = get!(() -> create_plan([(comp.operand, comp, -1)]), PLANS, comp)
        execute!(plan, time)                                                                 # (1) arguments, (2) window-aware write
    catch e
        e isa UnsupportedByLibrary || rethrow()
        return invoke(compute!, Tuple{Field{<:Any,<:Any,<:Any,<:KernelFunctionOperation},Any}, comp, time)   # Oceananigans' GPU kernel, never a CPU route
    end
    fill_halo_regions!(comp)                                                                 # (3)
    set_status!(comp.status, time)                                                           # (4)
    return comp                                                                              # (5)
end

# ---- compute! for Field(Integral(op)) / Field(Average(op)): the reduction fused into the sweep (SURVEY.md section 3.3) ------
# Oceananigans: Integral(op; dims = :) = Reduction(sum!, op * volume; dims); the operand of the reduced Field is that Reduction.
const ReducedKFOField = Field{Nothing,Nothing,Nothing,<:Oceananigans.AbstractOperations.Reduction}
"The KFO under `Integral(op)` (`op * Vᶜᶜᶜ` metric product) or `Average(op)`, or nothing."
function reduced_operand(red)
    x = red.operand
    kind = red isa Oceananigans.AbstractOperations.Average ? :average : :integral
    x isa BinaryOperation && x.op === (*) && x.a isa KernelFunctionOperation && (x = x.a)    # op * volume
    x isa KernelFunctionOperation && red.dims === (1, 2, 3) ? (x, kind) : nothing
end
function compute!(comp::ReducedKFOField, time = nothing)
    found = reduced_operand(comp.operand)
    (found === nothing || !supported(found[1])) && return invoke(compute!, Tuple{Field{<:Any,<:Any,<:Any,<:Oceananigans.AbstractOperations.Reduction},Any}, comp, time)
    op, kind = found
    group = get(GROUP_OF, comp, nothing)
    if group !== nothing
        compute_group!(group, time)
    else
        plan = get!(() -> create_plan([(op, nothing, 0)]), PLANS, comp)
        execute!(plan, time)
        store_reduction!(comp, plan.integrals, 1, kind, op)
    end
    set_status!(comp.status, time)
    return comp
end
function store_reduction!(comp, integrals, slot, kind, op)
    dst = view(parent(comp), :)                                                              # the 1 x 1 x 1 reduced parent
    copyto!(dst, 1, integrals, slot, 1)
    kind === :average && (dst ./= total_volume(op))                                          # Average = Σ f V / Σ V
    return comp
end
total_volume(op) = sum(Oceananigans.AbstractOperations.volume_operation(op))  # Σ V at the operation's location (host-side scalar, cached by Oceananigans)

# ---- the sibling registry: one sweep for every member (delegation pattern of BackgroundPotentialEnergyEquation.jl:246-262) --
mutable struct FusedDiagnostics
    members::Vector{Any}              # computed Fields and reduced Fields
    kinds::Vector{Any}                # (op, :field | :integral | :average, slot)
    plan::Plan
    time::Any
    computed::Bool
end
const GROUP_OF = IdDict{Any,FusedDiagnostics}()
"`fuse!(fields...)`: Field(op) and Field(Integral(op)) / Field(Average(op)) members evaluated by ONE sweep over u, v, w, b, p."
function fuse!(fields...)
    length(fields) <= OCB_MAX_REQ || throw(ArgumentError("at most $OCB_MAX_REQ diagnostics per fused sweep"))
    kinds, items, slot = Any[], Any[], 0
    for f in fields
        if f isa ComputedKFOField
            supported(f.operand) || throw(ArgumentError("$(summary(f.operand)) is not evaluated by liboceanostics_b200"))
            push!(kinds, (f.operand, :field, -1)); push!(items, (f.operand, f, -1))
        else
            found = reduced_operand(f.operand)
            found === nothing && throw(ArgumentError("fuse! takes Field(op), Field(Integral(op)) or Field(Average(op))"))
            push!(kinds, (found[1], found[2], slot)); push!(items, (found[1], nothing, slot)); slot += 1
        end
    end
    group = FusedDiagnostics(collect(Any, fields), kinds, create_plan(items), nothing, false)
    foreach(f -> GROUP_OF[f] = group, fields)
    return group
end
function compute_group!(g::FusedDiagnostics, time)
    # Oceananigans' compute_at! rule: always at t = 0, else only when the time moved
    g.computed && time !== nothing && time != zero(time) && time == g.time && return g
    execute!(g.plan, time)
    for (f, (op, kind, slot)) in zip(g.members, g.kinds)
        kind === :field ? fill_halo_regions!(f) : store_reduction!(f, g.plan.integrals, slot + 1, kind, op)
        set_status!(f.status, time)
    end
    g.time, g.computed = time, true
    return g
end

# ---- K2: the staged separable filters (replaces the body of _compute_staged_filter!, SpatialFilters.jl:347-378) ----------
const SF = Oceanostics.SpatialFilters
policy_code(::SF.PeriodicBoundary) = (OCB_POLICY_PERIODIC, 0.0, 0.0)
policy_code(::SF.ShrinkBoundary) = (OCB_POLICY_SHRINK, 0.0, 0.0)
policy_code(::SF.EdgeBoundary) = (OCB_POLICY_EDGE, 0.0, 0.0)
policy_code(p::SF.ConstantBoundary) = (OCB_POLICY_CONSTANT, Float64(p.left), Float64(p.right))
policy_code(p::SF.SizedBoundary) = policy_code(p.policy)
function filter_pass(kern, d::Int, width::Int, policy, extent::Int, ψ)
    code, left, right = policy_code(policy)
    w = zeros(Float64, 2 * OCB_MAX_FILTER_WIDTH + 1)
    kind, sigma, period, nodes, spacings = OCB_FILTER_BOX, 0.0, 0.0, NULLPTR, NULLPTR
    if kern isa SF.GaussianFilterKernel
        kind = OCB_FILTER_GAUSSIAN
        w[1:2width+1] .= kern.weights                                                          # gaussian_weights (gaussian_filter.jl:312)
    elseif kern isa SF.StretchedGaussianFilterKernel
        kind, sigma, period = OCB_FILTER_GAUSSIAN_STRETCHED, Float64(kern.σ), Float64(kern.L)
        nodes, spacings = devptr(kern.nodes), devptr(kern.spacings)
    end
    return OcbFilterPass(Int32(d), kind, Int32(width), code, Int32(extent), left, right, Tuple(w), sigma, period, nodes, spacings)
end
"`compute!` of a staged (2-D / 3-D) or single-pass filter Field: arguments first, window-aware destination, halos, status."
function compute_filter!(comp::Field, time)
    op = comp.operand
    ψ, passes = SF.filter_passes(op)                                                           # [(kernel, dim, width, policy)] in staging order
    compute_at!(ψ, time)                                                                       # (1) SpatialFilters.jl:354-356
    ψf = ψ isa Field ? ψ : materialized(ψ)
    extent = i32(size(interior(ψf)))
    descs = [filter_pass(k, d, w, pol, Int(extent[d]), ψf) for (k, d, w, pol) in passes]
    lo, hi, _ = window(comp)
    FT = eltype(comp.grid) === Float64 ? OCB_F64 : OCB_F32
    check(ccall((:ocb_filter_execute, lib), Cint,
                (Int32, Ref{OcbField}, Ref{NTuple{3,Int32}}, Ptr{OcbFilterPass}, Int32, Ref{OcbField}, Ref{NTuple{3,Int32}}, Ref{NTuple{3,Int32}}, Ptr{Cvoid}),
                FT, ocb_field(ψf), extent, descs, Int32(length(descs)), ocb_field(comp), lo, hi, stream_handle()))   # (2)
    fill_halo_regions!(comp)                                                                   # (3) :375
    set_status!(comp.status, time)                                                             # (4) :376
    return comp
end
for T in (:(SF._BoxFilter2D), :(SF._BoxFilter3D), :(SF._GaussianFilter2D), :(SF._GaussianFilter3D))   # box_filter.jl:216-230, gaussian_filter.jl:520-534
    @eval compute!(comp::Field{<:Any,<:Any,<:Any,<:$T}, time = nothing) =
        architecture(comp.grid) isa GPU ? compute_filter!(comp, time) : SF._compute_staged_filter!(comp, time)
end

# ---- K3: the sorted reference height (replaces the body of compute!(::SortedReferenceHeightField), BPE :610-619) -----------
method_code(::BPEq.ThreeDimensionalSort) = OCB_BPE_THREE_DIMENSIONAL_SORT
method_code(::BPEq.HeavisideIntegral) = OCB_BPE_HEAVISIDE_INTEGRAL
function compute!(z✶::BPEq.SortedReferenceHeightField, time = nothing)
    s = z✶.operand                                                                             # SortedReferenceState
    architecture(z✶.grid) isa GPU || return invoke(compute!, Tuple{Field,Any}, z✶, time)
    compute_at!(s.buoyancy, time)                                                              # :612
    b = s.buoyancy isa Field ? s.buoyancy : materialized(s.buoyancy)
    if s.method isa BPEq.ProfileLookup
        prof = BPEq.lookup_profile(s.method, s)                                                # validated (b✶, z✶) pair (:528-547)
        check(ccall((:ocb_bpe_profile_lookup, lib), Cint, (Ref{OcbGrid}, Ref{OcbField}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Ref{OcbField}, Ref{OcbField}, Ptr{Cvoid}),
                    ocb_grid(z✶.grid), ocb_field(b), devptr(prof.b), devptr(prof.z), Int64(length(prof.b)), ocb_field(z✶), ocb_field(s.reference_potential), stream_handle()))
    else
        check(ccall((:ocb_bpe_reference_height, lib), Cint, (Ref{OcbGrid}, Ref{OcbField}, Int32, Ref{OcbField}, Ref{OcbField}, CuPtr{Int64}, CuPtr{Cvoid}, Ptr{Cvoid}),
                    ocb_grid(z✶.grid), ocb_field(b), method_code(s.method), ocb_field(z✶), ocb_field(s.reference_potential),
                    pointer(s.perm), devptr(s.workspace), stream_handle()))                    # perm as sortperm! returns it (:757), b✶ in the workspace
    end
    fill_halo_regions!(z✶); fill_halo_regions!(s.reference_potential)                          # :615
    set_status!(z✶.status, time)                                                               # :616
    return z✶
end
"`VerticalSort` (BPE :584-597): a model-grid ccc field read in sorted order onto the 1 x 1 x N column."
gather_sorted!(dst::CuArray, src::Field, perm::CuVector{Int64}) =
    check(ccall((:ocb_gather_sorted, lib), Cint, (Int32, Ref{OcbField}, CuPtr{Int64}, CuPtr{Cvoid}, Int64, Ptr{Cvoid}),
                eltype(src.grid) === Float64 ? OCB_F64 : OCB_F32, ocb_field(src), pointer(perm), devptr(dst), Int64(length(perm)), stream_handle()))
"The radix sort alone: sorted keys and the 0-based permutation."
sort_pairs!(keys_out::CuVector{T}, perm::CuVector{UInt32}, keys_in::CuVector{T}) where T <: Union{Float32,Float64} =
    check(ccall((:ocb_sort_pairs, lib), Cint, (Int32, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{UInt32}, Int64, Ptr{Cvoid}),
                T === Float64 ? OCB_F64 : OCB_F32, devptr(keys_in), devptr(keys_out), pointer(perm), Int64(length(keys_in)), stream_handle()))
"`Field(a ∘ b)` / `@at loc a` with each operand interpolated to the destination's location."
pointwise!(dst::Field, op::Int32, a, b = nothing; alpha = 1.0) =
    check(ccall((:ocb_pointwise, lib), Cint, (Ref{OcbGrid}, Int32, Float64, Ref{OcbField}, Ref{OcbField}, Ref{OcbField}, Ptr{Cvoid}),
                ocb_grid(dst.grid), op, alpha, ocb_field(a), ocb_field(b), ocb_field(dst), stream_handle()))

# ---- y-slab decomposition on one NVLink box (Oceananigans Distributed with Partition(y = R), one rank per GPU) -------------
# MPI (or NCCL) only carries the 64-byte CUDA IPC handles at set-up; the data path is peer stores + flag words.
ipc_export(a::CuArray) = begin
    h, off = Vector{UInt8}(undef, 64), Ref{Int64}(0)
    check(ccall((:ocb_ipc_export, lib), Cint, (CuPtr{Cvoid}, Ptr{UInt8}, Ref{Int64}), devptr(a), h, off)); (h, off[])
end
ipc_import(h::Vector{UInt8}) = begin
    base = Ref{CuPtr{Cvoid}}(NULLPTR)
    check(ccall((:ocb_ipc_import, lib), Cint, (Ptr{UInt8}, Ref{CuPtr{Cvoid}}), h, base)); base[]
end
ipc_close(base::CuPtr{Cvoid}) = check(ccall((:ocb_ipc_close, lib), Cint, (CuPtr{Cvoid},), base))
"Store `h` edge rows of every field into the ring neighbours' halo rows and publish `step` in their flag words."
halo_push!(fields::Vector{<:Field}, h, peer_arrays::Vector{CuPtr{Cvoid}}, peer_ny::Vector{Int32}, lo_flag, hi_flag, step) =
    check(ccall((:ocb_halo_push_y_many, lib), Cint, (Int32, Ptr{OcbField}, Int32, Int32, Ptr{CuPtr{Cvoid}}, Ptr{Int32}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Ptr{Cvoid}),
                eltype(fields[1].grid) === Float64 ? OCB_F64 : OCB_F32, map(ocb_field, fields), Int32(length(fields)), Int32(h), peer_arrays, peer_ny,
                lo_flag, hi_flag, Int64(step), stream_handle()))
halo_wait!(flags::CuVector{Int64}, step; timeout = 10.0) =
    check(ccall((:ocb_halo_wait, lib), Cint, (CuPtr{Cvoid}, Int64, Float64, Ptr{Cvoid}), devptr(flags), Int64(step), Float64(timeout), stream_handle()))
"Pack / unpack transport for MPI send/recv (CUDA-aware MPI or NCCL on the Julia side)."
halo_pack!(buffer::CuArray, f::Field, side, h; unpack = false) =
    check(ccall((unpack ? :ocb_halo_unpack_y : :ocb_halo_pack_y, lib), Cint, (Int32, Ref{OcbField}, Int32, Int32, CuPtr{Cvoid}, Ptr{Cvoid}),
                eltype(f.grid) === Float64 ? OCB_F64 : OCB_F32, ocb_field(f), Int32(side), Int32(h), devptr(buffer), stream_handle()))
halo_pack_many!(buffers::Vector{CuPtr{Cvoid}}, fields::Vector{<:Field}, h; unpack = false) =
    check(ccall((:ocb_halo_pack_y_many, lib), Cint, (Int32, Ptr{OcbField}, Int32, Int32, Ptr{CuPtr{Cvoid}}, Int32, Ptr{Cvoid}),
                eltype(fields[1].grid) === Float64 ? OCB_F64 : OCB_F32, map(ocb_field, fields), Int32(length(fields)), Int32(h), buffers, Int32(unpack), stream_handle()))
"One slab's budget step over the peer-memory ring: the gated class-sum sweep + the in-kernel all-reduce; `plan.integrals` then holds the GLOBAL integrals."
function ring_step!(plan::Plan, ring::OcbRing, step)
    ccall((:ocb_plan_ring_capable, lib), Cint, (Ptr{Cvoid},), plan.handle) == 1 || throw(UnsupportedByLibrary("plan is not ring capable"))
    check(ccall((:ocb_plan_execute_ring, lib), Cint, (Ptr{Cvoid}, CuPtr{Float64}, Ref{OcbRing}, Int64, Ptr{Cvoid}), plan.handle, plan.integrals, ring, Int64(step), stream_handle()))
end

# ---- on-device LES closure field (the step before the path: Oceananigans' compute_diffusivities! for SmagorinskyLilly) ------
"νₑ = (C Δᶠ)² √(2 |S|² ς(N²/|S|²)) written into `νₑ` by the library (OCB_DIAG_NU_SMAG); `closure` carries C, Cb and Pr."
function compute_smagorinsky_viscosity!(νₑ::Field, closure, velocities, b)
    s = velnt(velocities); s.b = b
    s.les = (Float64(closure.C), Float64(something(closure.Cb, 0.0)), Float64(closure.Pr isa NamedTuple ? first(closure.Pr) : closure.Pr), 0.0)
    req = OcbRequest(diag_id(:NU_SMAG), Int32(0), devptr(parent(νₑ)), i32(halo_size(νₑ.grid)), i64(strides(parent(νₑ))), (0, 0, 0), (0, 0, 0), (1, 1, 1), Int32(-1))
    handle = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ocb_plan_create, lib), Cint, (Ref{OcbGrid}, Ref{OcbSweepInputs}, Ref{OcbRequest}, Int32, Ref{Ptr{Cvoid}}), ocb_grid(νₑ.grid), descriptor(s), req, Int32(1), handle))
    plan = Plan(handle[], CUDA.zeros(Float64, 1), s, 0)
    execute!(plan, nothing)
    return fill_halo_regions!(νₑ)
end

end # module
