# TerrariumB200.jl -- Julia host glue: Terrarium's `initialize` / `timestep!` / `run!` on top of
# libterrarium_b200.so (include/terrarium_b200.h) through `ccall`.
#
# STATUS: UNTESTED SOURCE.  Neither `julia` nor Terrarium's dependencies exist in the build image or
# on the GPU boxes, so this file has never been executed; the tested caller of the same C ABI is
# the Python ctypes mirror (terrarium.jl_b200/integrator.py).  The struct layouts below must match
# include/terrarium_b200.h field for field (TRM_ABI_VERSION = 4).
#
# What it replaces in the reference (paths relative to the Terrarium.jl root):
#   initialize(model, timestepper, inputs...)            src/timesteppers/model_integrator.jl:145-161
#   timestep!(integrator, ::ForwardEuler | ::Heun, dt)   src/timesteppers/forward_euler.jl:19-31, heun.jl:37-71
#   run!(integrator; steps, period, dt)                  src/timesteppers/model_integrator.jl:72-88
#   interior(field) / set!(field, value)                 src/state_variables.jl:476-489, src/initializers.jl:23-27
# Drop-in use: wrap the host grid in `B200Grid(...)`; the reference's `initialize` / `timestep!` / `run!` then dispatch here.
module TerrariumB200

using Terrarium
using Terrarium: SoilModel, LandModel, ForwardEuler, Heun, RichardsEq, NoFlow, ConstantSoilHydraulics,
                 UnsatKVanGenuchten, UnsatKLinear, default_dt, convert_dt, get_steps
import FreezeCurves: VanGenuchten, BrooksCorey

const LIB = get(ENV, "TERRARIUM_B200_LIB", joinpath(@__DIR__, "..", "csrc", "libterrarium_b200.so"))

const TRM_ABI_VERSION = Int32(4)
const TRM_BC_NSLOTS = 8
@enum FieldId::Cint internal_energy=0 temperature=1 liquid_water_fraction=2 saturation_water_ice=3 pressure_head=4 hydraulic_conductivity=5 surface_excess_water=6 water_table=7 ground_temperature=8 skin_temperature=9 ground_heat_flux=10 surface_shortwave_up=11 surface_longwave_up=12 surface_net_radiation=13 sensible_heat_flux=14 latent_heat_flux=15 evaporation_ground=16 infiltration=17 surface_runoff=18 carbon_vegetation=21 vegetation_area_fraction=22 canopy_water=23 balanced_leaf_area_index=24 leaf_area_index=25 phenology_factor=26 canopy_water_conductance=27 leaf_to_air_co2_ratio=28 net_assimilation=29 leaf_respiration=30 gross_primary_production=31 autotrophic_respiration=32 net_primary_production=33 soil_moisture_limiting_factor=34 canopy_water_interception=35 canopy_water_removal=36 saturation_canopy_water=37 rainfall_ground=38 evaporation_canopy=39 transpiration=40 plant_available_water=41 root_fraction=42

struct TrmParams                      # trm_params
    mineral_porosity::Cdouble; organic_porosity::Cdouble; rho_soc::Cdouble; rho_org::Cdouble
    kappa::NTuple{5, Cdouble}; heatcap::NTuple{5, Cdouble}
    rho_w::Cdouble; Lsl::Cdouble; Llg::Cdouble; rho_a::Cdouble; c_a::Cdouble; Tref::Cdouble; sigma::Cdouble; eps_mw::Cdouble
    K_sat::Cdouble; vg_alpha::Cdouble; vg_n::Cdouble; bc_psis::Cdouble; bc_lambda::Cdouble; theta_res::Cdouble
    impedance::Cdouble; vwc_forcing::Cdouble
    albedo::Cdouble; emissivity::Cdouble; kappa_skin::Cdouble; C_h::Cdouble; min_windspeed::Cdouble; tau_r::Cdouble; evap_beta::Cdouble
    # vegetated LandModel (40 doubles, header order: field_capacity .. C_can)
    vegetation::NTuple{40, Cdouble}
end
struct TrmBC; kind::Int32; input::Int32; end       # trm_bc
struct TrmConfig                      # trm_config
    abi_version::Int32; dtype::Int32; ncol::Int64; col0::Int64; nz::Int32; device::Int32
    model::Int32; timestepper::Int32; hydrology::Int32; swrc::Int32; unsat_k::Int32; sat_halo::Int32; skin::Int32; math::Int32
    vegetation::Int32                 # trm_vegetation: 0 = nothing (bare ground), 1 = VegetationCarbon
    ground_resistance::Int32          # trm_ground_resistance: 0 = constant factor, 1 = SoilMoistureResistanceFactor
    albedo_kind::Int32                # 0 = ConstantAlbedo, 1 = PrescribedAlbedo (inputs :albedo, :emissivity)
    radiative::Int32                  # 0 = DiagnosedRadiativeFluxes, 1 = PrescribedRadiativeFluxes
    turbulent::Int32                  # 0 = DiagnosedTurbulentFluxes, 1 = PrescribedTurbulentFluxes
    reserved0::Int32
    z_faces::Ptr{Cdouble}
    params::TrmParams
    bc::NTuple{TRM_BC_NSLOTS, TrmBC}
end

struct B200Error <: Exception; status::Cint; msg::String; end
function check(status::Cint, what)
    status == 0 && return nothing
    throw(B200Error(status, "trm_$what failed ($status): " * unsafe_string(ccall((:trm_last_error, LIB), Cstring, ()))))
end

"""Integrator whose state lives in the B200 library; mirrors `Terrarium.ModelIntegrator`."""
mutable struct B200Integrator{NF, Model, TS}
    handle::Ptr{Cvoid}
    model::Model
    timestepper::TS
    ncol::Int
    nz::Int
    # function valued boundary conditions f(x, t) (examples/simulations/soil_heat_global.jl:72-93): evaluated on the host before
    # every step at the x-nodes of the columns (input slot => function)
    callbacks::Vector{Pair{Int, Any}}
    xnodes::Vector{NF}
end
B200Integrator{NF, M, TS}(handle, model, ts, ncol, nz) where {NF, M, TS} = B200Integrator{NF, M, TS}(handle, model, ts, ncol, nz, Pair{Int, Any}[], NF[])
Base.eltype(::B200Integrator{NF}) where {NF} = NF

dtype_code(::Type{Float32}) = Int32(0)
dtype_code(::Type{Float64}) = Int32(1)

# residual water content of a FreezeCurves SWRC (`swrc.vol.θres`, FreezeCurves.jl SoilWaterVolume); 0 when the type has none
swrc_theta_res(swrc) = hasproperty(swrc, :vol) && hasproperty(swrc.vol, :θres) ? Float64(ustrip(swrc.vol.θres)) : 0.0
# user VWC forcing (src/processes/soil/hydrology/soil_hydrology.jl:38-48): the library evaluates a constant source / sink in
# every cell. `nothing` -> 0; a discrete-form Forcing whose parameters carry a `.value` (the form of
# test/soil/soil_hydrology_tests.jl:191-233) or a plain number -> that value; anything else cannot run inside the kernel.
function vwc_forcing_value(f)
    isnothing(f) && return 0.0
    p = hasproperty(f, :parameters) ? f.parameters : f
    p isa Number && return Float64(p)
    hasproperty(p, :value) && return Float64(p.value)
    error("TerrariumB200: only constant VWC forcings (a number, or Forcing(parameters = (value = ...), discrete_form = true)) are supported on the B200 path")
end

function params_of(model)
    soil, c = model.soil, model.constants
    k, h = soil.energy.thermal_properties.conductivities, soil.energy.thermal_properties.heat_capacities
    hp = soil.hydrology.hydraulic_properties
    vg = hp.swrc isa VanGenuchten ? hp.swrc : VanGenuchten()
    bc = hp.swrc isa BrooksCorey ? hp.swrc : BrooksCorey()
    land = model isa LandModel
    seb = land ? model.surface_energy_balance : nothing
    TrmParams(
        soil.strat.porosity.mineral_porosity, soil.strat.porosity.organic_porosity, soil.biogeochem.ρ_soc, soil.biogeochem.ρ_org,
        (k.water, k.ice, k.air, k.mineral, k.organic), (h.water, h.ice, h.air, h.mineral, h.organic),
        c.ρw, c.Lsl, c.Llg, c.ρₐ, c.cₐ, c.Tref, c.σ, c.ε,
        hp.sat_hydraulic_cond, ustrip(vg.α), vg.n, ustrip(bc.ψₛ), bc.λ, swrc_theta_res(hp.swrc),
        hp.unsat_hydraulic_cond isa UnsatKVanGenuchten ? hp.unsat_hydraulic_cond.impedance : 7.0, vwc_forcing_value(soil.hydrology.vwc_forcing),
        (land && seb.albedo isa Terrarium.ConstantAlbedo) ? seb.albedo.albedo : 0.3, (land && seb.albedo isa Terrarium.ConstantAlbedo) ? seb.albedo.emissivity : 0.97, land ? seb.skin_temperature.κₛ : 2.0,
        land ? model.atmosphere.aerodynamics.C_h : 1.2e-3, land ? model.atmosphere.min_windspeed : 0.01,
        land ? model.surface_hydrology.surface_runoff.τ_r : 3600.0, 1.0,
        vegetation_params(model, hp, c))
end

# parameters of VegetationCarbon + PALADYN canopy hydrology in the order of trm_params (reference defaults when absent)
function vegetation_params(model, hp, c)
    veg = model isa LandModel && !isnothing(model.vegetation) ? model.vegetation : Terrarium.VegetationCarbon(Float64)
    ph, sc, ar, cd, vd, rd = veg.photosynthesis, veg.stomatal_conductance, veg.autotrophic_respiration, veg.carbon_dynamics,
                             veg.vegetation_dynamics, veg.root_distribution
    vegetated = model isa LandModel && !isnothing(model.vegetation)
    ci = vegetated ? model.surface_hydrology.canopy_interception : Terrarium.PALADYNCanopyInterception(Float64)
    et = vegetated ? model.surface_hydrology.evapotranspiration : Terrarium.PALADYNCanopyEvapotranspiration(Float64)
    Cdouble.((hp.field_capacity, hp.wilting_point, c.C_mass,
              ph.τ25, ph.Kc25, ph.Ko25, ph.q10_τ, ph.q10_Kc, ph.q10_Ko, ph.α_leaf, ph.α_a, ph.α_C3, ph.cq, ph.k_ext,
              ph.T_CO2_high, ph.T_CO2_low, ph.T_photos_high, ph.T_photos_low, ph.θ_r,
              sc.g₁, sc.g_min, ar.cn_sapwood, ar.cn_root, ar.aws,
              cd.SLA, cd.awl, cd.LAI_min, cd.LAI_max, cd.γL, cd.γR, cd.γS,
              vd.ν_seed, vd.γv_min, rd.a, rd.b,
              ci.α_int, ci.k_ext, ci.w_can_max, ci.τ_w, et.C_can))
end

"""
    initialize_b200(model, timestepper; device = 0, math = :faithful, col0 = 0, ncol = nothing)

`Terrarium.initialize` for the B200 library.  Initial conditions are taken from a CPU `initialize` of the same
model (the reference's own initializer code path) and uploaded with `trm_set_field`.
"""
function initialize_b200(model::Union{SoilModel{NF}, LandModel{NF}}, timestepper; device = 0, math = :faithful,
                         boundary_conditions = (;), initializers = (;)) where {NF}
    # CPU reference state at t0 (initial interior values only: boundary conditions act from the first step on and may hold
    # device-resident sources the CPU path does not know)
    ref = Terrarium.initialize(model, timestepper; initializers)
    grid = Terrarium.get_field_grid(Terrarium.get_grid(model))
    zf = collect(Float64, Terrarium.znodes(grid, Terrarium.Face()))
    nz, ncol = length(zf) - 1, size(grid, 1)
    hyd = model.soil.hydrology
    bcs, bc_sources = translate_bcs(boundary_conditions)
    cfg = Ref(TrmConfig(TRM_ABI_VERSION, dtype_code(NF), ncol, 0, nz, device,
        model isa LandModel ? 1 : 0, timestepper isa Heun ? 1 : 0, hyd.vertical_flow isa RichardsEq ? 1 : 0,
        hyd.hydraulic_properties.swrc isa VanGenuchten ? 0 : 1, hyd.hydraulic_properties.unsat_hydraulic_cond isa UnsatKVanGenuchten ? 1 : 0,
        0, 0, math === :fast ? 1 : 0, (model isa LandModel && !isnothing(model.vegetation)) ? 1 : 0,
        (model isa LandModel && model.surface_hydrology.evapotranspiration.ground_resistance isa Terrarium.SoilMoistureResistanceFactor) ? 1 : 0,
        (model isa LandModel && model.surface_energy_balance.albedo isa Terrarium.PrescribedAlbedo) ? 1 : 0,
        (model isa LandModel && model.surface_energy_balance.radiative_fluxes isa Terrarium.PrescribedRadiativeFluxes) ? 1 : 0,
        (model isa LandModel && model.surface_energy_balance.turbulent_fluxes isa Terrarium.PrescribedTurbulentFluxes) ? 1 : 0, 0,
        pointer(zf), params_of(model), bcs))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve zf check(ccall((:trm_create, LIB), Cint, (Ref{TrmConfig}, Ref{Ptr{Cvoid}}), cfg, h), "create")
    integ = B200Integrator{NF, typeof(model), typeof(timestepper)}(h[], model, timestepper, ncol, nz)
    finalizer(i -> ccall((:trm_destroy, LIB), Cint, (Ptr{Cvoid},), i.handle), integ)
    for (input, source) in bc_sources                       # device-resident boundary values (user input slots 0..7)
        set_input!(integ, input, source)
    end
    check(ccall((:trm_reset, LIB), Cint, (Ptr{Cvoid},), integ.handle), "reset")   # reset!(integrator.state), model_integrator.jl:98
    names = (model isa LandModel && !isnothing(model.vegetation)) ?
        (:temperature, :saturation_water_ice, :carbon_vegetation, :vegetation_area_fraction, :canopy_water) : (:temperature, :saturation_water_ice)
    for name in names
        set_field!(integ, name, permutedims(Array(Terrarium.interior(getproperty(ref.state, name)))[:, 1, :]))   # [layer, column]
    end
    check(ccall((:trm_initialize, LIB), Cint, (Ptr{Cvoid},), integ.handle), "initialize")
    return integ
end

function set_field!(integ::B200Integrator{NF}, name::Symbol, values::AbstractArray) where {NF}
    v = Array{NF}(values)
    GC.@preserve v check(ccall((:trm_set_field, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64), integ.handle,
                               Cint(getproperty(TerrariumB200, name)), pointer(v), length(v)), "set_field")
end

"""`interior(field)` as a host array `[layer, column]` (layer 1 = bottom cell) or `[column]` for 2-D fields."""
function get_field(integ::B200Integrator{NF}, name::Symbol) where {NF}
    rows = name in (:internal_energy, :temperature, :liquid_water_fraction, :saturation_water_ice, :pressure_head,
                    :plant_available_water, :root_fraction) ? integ.nz :
           name === :hydraulic_conductivity ? integ.nz + 1 : 1
    out = Array{NF}(undef, integ.ncol, rows)
    check(ccall((:trm_get_field, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64), integ.handle,
                Cint(getproperty(TerrariumB200, name)), out, length(out)), "get_field")
    return rows == 1 ? vec(out) : permutedims(out)
end

"""Zero-copy view of a device field: `(CuPtr, leading_dimension, rows)` for `unsafe_wrap(CuArray, ...)`."""
function field_ptr(integ::B200Integrator, name::Symbol)
    p, ld, rows = Ref{Ptr{Cvoid}}(C_NULL), Ref{Int64}(0), Ref{Int32}(0)
    check(ccall((:trm_field_ptr, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{Ptr{Cvoid}}, Ref{Int64}, Ref{Int32}), integ.handle,
                Cint(getproperty(TerrariumB200, name)), p, ld, rows), "field_ptr")
    return p[], ld[], rows[]
end

"""
    bind_host_io!(integ, input, field; nslots = 4) -> (ring_in, ring_out)

Per-step exchange with a host-side coupler (the flow of examples/simulations/speedy_dry_land.jl:45-68) through page-locked
host memory that the stage kernel reads and writes directly (`trm_bind_host_io`): the step from iteration `k` reads input
`input` from `ring_in[:, k % nslots + 1]` and stores the 2-D field `field` of the new state in `ring_out[:, k % nslots + 1]`.
`wait_step(integ, k)` blocks until the step that produced iteration `k` has completed.
"""
function bind_host_io!(integ::B200Integrator{NF}, input::Integer, field::Symbol; nslots = 4) where {NF}
    nbytes = nslots * integ.ncol * sizeof(NF)
    pin, pout = Ref{Ptr{Cvoid}}(C_NULL), Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:trm_host_alloc_ex, LIB), Cint, (Int64, Int32, Ref{Ptr{Cvoid}}), nbytes, 1, pin), "host_alloc_ex")   # write-combined
    check(ccall((:trm_host_alloc_ex, LIB), Cint, (Int64, Int32, Ref{Ptr{Cvoid}}), nbytes, 0, pout), "host_alloc_ex")
    check(ccall((:trm_bind_host_io, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int32), integ.handle,
                Cint(input), pin[], Cint(getproperty(TerrariumB200, field)), pout[], Int32(nslots)), "bind_host_io")
    return unsafe_wrap(Array, Ptr{NF}(pin[]), (integ.ncol, nslots)), unsafe_wrap(Array, Ptr{NF}(pout[]), (integ.ncol, nslots))
end
wait_step(integ::B200Integrator, iteration::Integer) =
    check(ccall((:trm_host_io_wait, LIB), Cint, (Ptr{Cvoid}, Int64), integ.handle, Int64(iteration)), "host_io_wait")
step_async!(integ::B200Integrator, Δt, nsteps = 1) =
    check(ccall((:trm_step_async, LIB), Cint, (Ptr{Cvoid}, Cdouble, Int64), integ.handle, convert_dt(Δt), Int64(nsteps)), "step_async")

function Terrarium.timestep!(integ::B200Integrator, Δt = default_dt(integ.timestepper); finalize = true)
    isempty(integ.callbacks) || push_callbacks!(integ, current_time(integ), convert_dt(Δt))
    check(ccall((:trm_step, LIB), Cint, (Ptr{Cvoid}, Cdouble, Int64), integ.handle, convert_dt(Δt), 1), "step")
    finalize && check(ccall((:trm_compute_auxiliary, LIB), Cint, (Ptr{Cvoid},), integ.handle), "compute_auxiliary")
    return nothing
end

function Terrarium.run!(integ::B200Integrator; steps = nothing, period = nothing, Δt = default_dt(integ.timestepper))
    Δt = convert_dt(Δt)
    n = get_steps(steps, period, Δt)
    if isempty(integ.callbacks)
        check(ccall((:trm_step, LIB), Cint, (Ptr{Cvoid}, Cdouble, Int64), integ.handle, Δt, n), "step")      # n fused stage launches (small domains: one launch)
    else
        for _ in 1:n                                                                                        # host functions: step by step
            push_callbacks!(integ, current_time(integ), Δt)
            check(ccall((:trm_step, LIB), Cint, (Ptr{Cvoid}, Cdouble, Int64), integ.handle, Δt, 1), "step")
        end
    end
    check(ccall((:trm_compute_auxiliary, LIB), Cint, (Ptr{Cvoid},), integ.handle), "compute_auxiliary")
    return integ
end

# ---- multi-GPU: one Julia process per GPU (Distributed / MPI.jl), one integrator per process over its column range; the
# columns never exchange data, NCCL only reduces the global diagnostics (SURVEY.md 8e) ------------------------------------
"""128-byte NCCL unique id created by rank 0; distribute it to the other ranks (e.g. `MPI.Bcast!`, `Distributed.remotecall`)."""
function nccl_unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:trm_nccl_get_unique_id, LIB), Cint, (Ptr{Cvoid},), id), "nccl_get_unique_id")
    return id
end
"""Create the handle's NCCL communicator (`ncclCommInitRank`); collective over all ranks."""
init_comm!(integ::B200Integrator, nranks::Integer, rank::Integer, id::Vector{UInt8}) =
    check(ccall((:trm_nccl_comm_init, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Cvoid}), integ.handle, nranks, rank, id), "nccl_comm_init")
"""Use a communicator NCCL.jl already owns: `adopt_comm!(integ, comm.handle)`."""
adopt_comm!(integ::B200Integrator, comm::Ptr{Cvoid}) =
    check(ccall((:trm_nccl_comm_adopt, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), integ.handle, comm), "nccl_comm_adopt")
"""Global diagnostics (energy, water, extrema, NaN count over ALL ranks' columns); collective."""
function global_diagnostics(integ::B200Integrator)
    d = Ref(TrmDiag(0, 0, 0, 0, 0, 0, 0, 0))
    check(ccall((:trm_diagnostics_allreduce, LIB), Cint, (Ptr{Cvoid}, Ref{TrmDiag}), integ.handle, d), "diagnostics_allreduce")
    return d[]
end

function Terrarium.current_time(integ::B200Integrator)
    t, it = Ref{Cdouble}(0), Ref{Int64}(0)
    check(ccall((:trm_get_clock, LIB), Cint, (Ptr{Cvoid}, Ref{Cdouble}, Ref{Int64}), integ.handle, t, it), "get_clock")
    return t[]
end

# ---- boundary conditions --------------------------------------------------------------------------------------------
# The reference passes `boundary_conditions = (temperature = (top = ValueBoundaryCondition(v),), ...)`
# (src/models/soil/soil_model_bcs.jl:6-40). A closure cannot run inside the library's kernels, so `v` must be one of the
# device-resident sources below (or a number / per-column vector); a Julia function is rejected with an explanation.

"""`clamp(mean[c] + amp[c] * sin(2π t / period - phase[c]), lo, hi)`: the device form of the periodic boundary
condition of examples/simulations/soil_heat_global.jl:72-93."""
Base.@kwdef struct Sinusoid{M, A, P}
    mean::M
    amp::A
    phase::P = 0.0
    period::Float64 = 86400.0
    lo::Float64 = -Inf
    hi::Float64 = Inf
end

"""Snapshots `values[column, it]` at `times[it]` (seconds), interpolated like `FieldTimeSeries[Time(t)]`
(src/input_output/input_sources.jl:165-171); `raster = true` selects the update rule of `RasterInputSource`
(ext/TerrariumRastersExt/TerrariumRastersExt.jl:96-121)."""
struct TimeSeries{V <: AbstractMatrix}
    times::Vector{Float64}
    values::V
    raster::Bool
end
TimeSeries(times, values; raster = false) = TimeSeries(collect(Float64, times), values, raster)

const BC_SLOT = Dict((:temperature, :top) => 0, (:temperature, :bottom) => 1, (:internal_energy, :top) => 2,
                     (:internal_energy, :bottom) => 3, (:saturation_water_ice, :top) => 4, (:saturation_water_ice, :bottom) => 5,
                     (:pressure_head, :top) => 6, (:pressure_head, :bottom) => 7)

bc_kind(bc) = bc_kind(bc.classification)
bc_kind(::Terrarium.Oceananigans.BoundaryConditions.Value) = Int32(1)
bc_kind(::Terrarium.Oceananigans.BoundaryConditions.Gradient) = Int32(2)
bc_kind(::Terrarium.Oceananigans.BoundaryConditions.Flux) = Int32(3)

function translate_bcs(boundary_conditions)
    slots = [TrmBC(0, 0) for _ in 1:TRM_BC_NSLOTS]
    sources = Pair{Int, Any}[]
    for (field, sides) in pairs(boundary_conditions), (side, bc) in pairs(sides)
        slot = get(BC_SLOT, (field, side)) do
            throw(ArgumentError("no boundary condition slot for $field / $side in the B200 library"))
        end
        isnothing(bc.condition) && continue                  # NoFluxBoundaryCondition: the default
        # (a Julia function cannot run inside the library's kernels: it is evaluated on the host before every step, see
        #  set_input!(integ, id, ::Function); device-resident forms -- Sinusoid, TimeSeries, vectors -- avoid that round trip)
        input = length(sources)                              # TRM_IN_USER0 + k
        input < 8 || throw(ArgumentError("at most 8 user boundary inputs"))
        slots[slot + 1] = TrmBC(bc_kind(bc), Int32(input))
        push!(sources, input => bc.condition)
    end
    return Tuple(slots), sources
end

# ---- inputs / forcing -----------------------------------------------------------------------------------------------
const INPUT_ID = Dict(:air_temperature => 8, :air_pressure => 9, :windspeed => 10, :specific_humidity => 11, :rainfall => 12,
                      :snowfall => 13, :surface_shortwave_down => 14, :surface_longwave_down => 15, :daytime_length => 16,
                      :CO2 => 17, :skin_temperature => 18, :SAI => 19, :daily_leaf_respiration => 20,
                      :albedo => 21, :emissivity => 22)
# inputs of the prescribed flux schemes (they carry the names of the fields the diagnosed schemes write): set_input!(integ, PRESCRIBED_INPUT_ID[name], v)
const PRESCRIBED_INPUT_ID = Dict(:surface_shortwave_up => 23, :surface_longwave_up => 24, :sensible_heat_flux => 25, :latent_heat_flux => 26)
input_id(id::Integer) = Cint(id)
input_id(name::Symbol) = Cint(INPUT_ID[name])

percolumn(::Type{NF}, v::Number, n) where {NF} = fill(NF(v), n)
percolumn(::Type{NF}, v::AbstractVector, n) where {NF} = (length(v) == n || throw(DimensionMismatch("expected $n columns")); Array{NF}(v))

"""`set_input!(integ, name_or_slot, source)`: number, per-column vector, `Sinusoid` or `TimeSeries`
(replaces `InputSource`s, src/input_output/input_sources.jl:81-171)."""
function set_input!(integ::B200Integrator, id, v::Number)
    check(ccall((:trm_set_input_const, LIB), Cint, (Ptr{Cvoid}, Cint, Cdouble), integ.handle, input_id(id), v), "set_input_const")
end
function set_input!(integ::B200Integrator{NF}, id, v::AbstractVector) where {NF}
    a = percolumn(NF, v, integ.ncol)
    GC.@preserve a check(ccall((:trm_set_input_field, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), integ.handle, input_id(id), pointer(a)), "set_input_field")
end
function set_input!(integ::B200Integrator{NF}, id, f::Function) where {NF}
    filter!(p -> first(p) != Int(input_id(id)), integ.callbacks)
    push!(integ.callbacks, Int(input_id(id)) => f)
    if isempty(integ.xnodes)   # x-nodes of the columns (SURVEY.md Appendix B.7): ColumnRingGrid 1 + (i - 1/2)(Nc - 1)/Nc, ColumnGrid (i - 1/2)/Nc
        grid = Terrarium.get_field_grid(Terrarium.get_grid(integ.model))
        integ.xnodes = collect(NF, Terrarium.Oceananigans.Grids.xnodes(grid, Center()))
    end
    push_callbacks!(integ, current_time(integ), nothing)
end
# values of every function valued boundary condition for the step that starts at `t`: f(x, t), and under Heun also
# f(x, t + Δt), which the second stage reads where the reference re-evaluates the function at the stage clock (heun.jl:53)
function push_callbacks!(integ::B200Integrator{NF}, t, Δt) where {NF}
    for (id, f) in integ.callbacks
        v0 = NF[f(x, NF(t)) for x in integ.xnodes]
        if isnothing(Δt) || !(integ.timestepper isa Heun)
            GC.@preserve v0 check(ccall((:trm_set_input_field, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), integ.handle, Cint(id), pointer(v0)), "set_input_field")
        else
            v1 = NF[f(x, NF(t) + NF(Δt)) for x in integ.xnodes]
            GC.@preserve v0 v1 check(ccall((:trm_set_input_field_pair, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}), integ.handle,
                                           Cint(id), pointer(v0), pointer(v1)), "set_input_field_pair")
        end
    end
end
function set_input!(integ::B200Integrator{NF}, id, s::Sinusoid) where {NF}
    m, a, p = percolumn(NF, s.mean, integ.ncol), percolumn(NF, s.amp, integ.ncol), percolumn(NF, s.phase, integ.ncol)
    GC.@preserve m a p check(ccall((:trm_set_input_sinusoid, LIB), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Cdouble),
        integ.handle, input_id(id), pointer(m), pointer(a), pointer(p), s.period, s.lo, s.hi), "set_input_sinusoid")
end
function set_input!(integ::B200Integrator{NF}, id, ts::TimeSeries) where {NF}
    size(ts.values) == (integ.ncol, length(ts.times)) || throw(DimensionMismatch("values must be [column, time]"))
    v, t = Array{NF}(ts.values), ts.times                     # column fastest = the library's [nt][ncol]
    GC.@preserve v t if ts.raster
        check(ccall((:trm_set_input_raster, LIB), Cint, (Ptr{Cvoid}, Cint, Int32, Ptr{Cdouble}, Ptr{Cvoid}),
                    integ.handle, input_id(id), length(t), pointer(t), pointer(v)), "set_input_raster")
    else
        check(ccall((:trm_set_input_table, LIB), Cint, (Ptr{Cvoid}, Cint, Int32, Ptr{Cdouble}, Ptr{Cvoid}),
                    integ.handle, input_id(id), length(t), pointer(t), pointer(v)), "set_input_table")
    end
end

"""Input variable as of the last `update_inputs!` (reading `state.<input>` in the reference)."""
function get_input(integ::B200Integrator{NF}, id) where {NF}
    out = Vector{NF}(undef, integ.ncol)
    check(ccall((:trm_get_input, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64), integ.handle, input_id(id), out, length(out)), "get_input")
    return out
end

"""Device pointer of a per-column input for in-place coupling (examples/simulations/speedy_dry_land.jl:45-68)."""
function input_ptr(integ::B200Integrator, id)
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:trm_input_ptr, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{Ptr{Cvoid}}), integ.handle, input_id(id), p), "input_ptr")
    return p[]
end

# ---- ColumnRingGrid conversions (src/grids/column_ring_grid.jl:102-149) ------------------------------------------------
"""Hand the mask of a `ColumnRingGrid` to the library: `ring_index = findall(mask)` (0-based on the C side)."""
function set_ring_mask!(integ::B200Integrator, mask::AbstractVector{Bool})
    idx = Int64.(findall(mask) .- 1)
    length(idx) == integ.ncol || throw(DimensionMismatch("mask selects $(length(idx)) points, the integrator has $(integ.ncol) columns"))
    GC.@preserve idx check(ccall((:trm_set_ring_index, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64), integ.handle, pointer(idx), length(mask)), "set_ring_index")
    return length(mask)
end

field_rows(integ::B200Integrator, name::Symbol) =
    name in (:internal_energy, :temperature, :liquid_water_fraction, :saturation_water_ice, :pressure_head,
             :plant_available_water, :root_fraction) ? integ.nz : name === :hydraulic_conductivity ? integ.nz + 1 : 1

"""`RingGrids.Field(field, grid; fill_value)` data: `[ring point, layer]` with `fill_value` at the ocean points."""
function get_field_ring(integ::B200Integrator{NF}, name::Symbol, nring::Integer; fill_value = NaN) where {NF}
    out = Array{NF}(undef, nring, field_rows(integ, name))
    check(ccall((:trm_get_field_ring, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64, Cdouble), integ.handle,
                Cint(getproperty(TerrariumB200, name)), out, length(out), fill_value), "get_field_ring")
    return out
end
function set_field_ring!(integ::B200Integrator{NF}, name::Symbol, ring::AbstractArray) where {NF}
    v = Array{NF}(ring)
    GC.@preserve v check(ccall((:trm_set_field_ring, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64), integ.handle,
                               Cint(getproperty(TerrariumB200, name)), pointer(v), length(v)), "set_field_ring")
end

# ---- diagnostics, clock, synchronisation ---------------------------------------------------------------------------------
struct TrmDiag                        # trm_diag
    energy::Cdouble; water::Cdouble; t_min::Cdouble; t_max::Cdouble; sat_min::Cdouble; sat_max::Cdouble; nan_count::Cdouble; ncol::Cdouble
end
"""Budgets and extrema of the columns owned by this handle (a multi-GPU caller reduces them, e.g. with NCCL.jl)."""
function diagnostics(integ::B200Integrator)
    d = Ref(TrmDiag(0, 0, 0, 0, 0, 0, 0, 0))
    check(ccall((:trm_diagnostics, LIB), Cint, (Ptr{Cvoid}, Ref{TrmDiag}), integ.handle, d), "diagnostics")
    return d[]
end
synchronize(integ::B200Integrator) = check(ccall((:trm_sync, LIB), Cint, (Ptr{Cvoid},), integ.handle), "sync")
set_clock!(integ::B200Integrator, time, iteration = 0) =
    check(ccall((:trm_set_clock, LIB), Cint, (Ptr{Cvoid}, Cdouble, Int64), integ.handle, time, iteration), "set_clock")
compute_tendencies!(integ::B200Integrator) = check(ccall((:trm_compute_tendencies, LIB), Cint, (Ptr{Cvoid},), integ.handle), "compute_tendencies")

"""Stream-ordered stepping without a host synchronisation (`Simulation` between two scheduled events)."""
step_async!(integ::B200Integrator, Δt, n) =
    check(ccall((:trm_step_async, LIB), Cint, (Ptr{Cvoid}, Cdouble, Int64), integ.handle, convert_dt(Δt), n), "step_async")

# ---- time-averaged output (Oceananigans AveragedTimeInterval) ---------------------------------------------------------
accumulate!(integ::B200Integrator, name::Symbol, weight) =
    check(ccall((:trm_accumulate, LIB), Cint, (Ptr{Cvoid}, Cint, Cdouble), integ.handle, Cint(getproperty(TerrariumB200, name)), weight), "accumulate")
function get_accumulated(integ::B200Integrator{NF}, name::Symbol; scale = 1.0, reset = true) where {NF}
    rows = field_rows(integ, name)
    out = Array{NF}(undef, integ.ncol, rows)
    check(ccall((:trm_get_accumulated, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64, Cdouble, Int32), integ.handle,
                Cint(getproperty(TerrariumB200, name)), out, length(out), scale, reset ? 1 : 0), "get_accumulated")
    return rows == 1 ? vec(out) : permutedims(out)
end

# ---- drop-in dispatch: the reference's own calls, selected by the grid type ----------------------------------------------------
# `grid = B200Grid(ColumnGrid(CPU(), Float64, ExponentialSpacing(N = 30), 10_000_000); device = 0)` is the only line a
# user script changes: `SoilModel(grid)` / `LandModel(grid; ...)`, `initialize(model, timestepper, inputs...; boundary_conditions,
# initializers)`, `timestep!`, `run!`, `integrator.state.<name>` and `interior(...)` then go to the library. The wrapped
# host grid (a CPU `ColumnGrid` / `ColumnRingGrid`) supplies coordinates, the mask and the reference's initializer code path.

struct B200Arch
    device::Int
end

struct B200Grid{NF, G <: Terrarium.AbstractLandGrid{NF}} <: Terrarium.AbstractLandGrid{NF, B200Arch}
    host::G
    device::Int
    math::Symbol
end
B200Grid(host::Terrarium.AbstractLandGrid{NF}; device = 0, math = :faithful) where {NF} = B200Grid{NF, typeof(host)}(host, device, math)
Terrarium.get_field_grid(grid::B200Grid) = Terrarium.get_field_grid(grid.host)
Terrarium.Oceananigans.Architectures.architecture(grid::B200Grid) = B200Arch(grid.device)
Base.eltype(::B200Grid{NF}) where {NF} = NF
Base.show(io::IO, grid::B200Grid) = print(io, "B200Grid(device = $(grid.device), math = $(grid.math)) over ", grid.host)

# the same model on the wrapped host grid (all reference model structs are @kwdef with a `grid` field)
function host_model(model::M) where {M <: Terrarium.AbstractModel}
    fields = (; (f => getfield(model, f) for f in fieldnames(M))...)
    return M.name.wrapper(; fields..., grid = model.grid.host)
end

"""State access of a `B200Integrator`: `integ.state.temperature` is the host copy `[column, 1, layer]` of the device
field (the shape `interior(field)` has in the reference); inputs are read as of the last `update_inputs!`."""
struct B200State{I}
    integrator::I
end
function Base.getproperty(state::B200State, name::Symbol)
    name === :integrator && return getfield(state, :integrator)
    integ = getfield(state, :integrator)
    if name === :inputs
        return state
    elseif haskey(INPUT_ID, name) && !isdefined(TerrariumB200, name)
        return get_input(integ, name)
    end
    a = get_field(integ, name)                        # [layer, column] or [column]
    return a isa AbstractVector ? reshape(a, :, 1, 1) : reshape(permutedims(a), size(a, 2), 1, size(a, 1))
end
Terrarium.Oceananigans.interior(a::Array) = a          # `interior(integ.state.temperature)` as in the examples

function Base.getproperty(integ::B200Integrator, name::Symbol)
    name === :state && return B200State(integ)
    name === :clock && return Terrarium.Oceananigans.TimeSteppers.Clock(; time = Terrarium.current_time(integ), iteration = iteration(integ))
    name === :grid && return Terrarium.get_field_grid(getfield(integ, :model).grid)   # (as ModelIntegrator does for the output writers)
    return getfield(integ, name)
end

# ---- Oceananigans model interface (src/timesteppers/model_integrator.jl:39-66): `Simulation(integ; Δt, stop_time)`, its
# callbacks and its output writers drive a B200Integrator like a ModelIntegrator. Output writers take functions of the
# model next to Fields: `output_functions(integ, names)` hands them the host copy of the named fields, shaped like
# `interior(field)`, fetched from the library when the writer's schedule fires --------------------------------------------
function iteration(integ::B200Integrator)
    it = Ref{Int64}(0)
    check(ccall((:trm_get_clock, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Int64}), integ.handle, C_NULL, it), "get_clock")
    return Int(it[])
end
Base.time(integ::B200Integrator) = Terrarium.current_time(integ)
Base.eltype(::B200Integrator{NF}) where {NF} = NF
Terrarium.Oceananigans.Solvers.iteration(integ::B200Integrator) = iteration(integ)
Terrarium.Oceananigans.Architectures.architecture(integ::B200Integrator) = Terrarium.Oceananigans.Architectures.architecture(getfield(integ, :model).grid)
Terrarium.Oceananigans.TimeSteppers.update_state!(integ::B200Integrator; compute_tendencies = true) =
    check(ccall((:trm_compute_auxiliary, LIB), Cint, (Ptr{Cvoid},), integ.handle), "compute_auxiliary")
Terrarium.Oceananigans.TimeSteppers.time_step!(integ::B200Integrator, Δt; kwargs...) = Terrarium.timestep!(integ, Δt)
Terrarium.Oceananigans.Simulations.timestepper(integ::B200Integrator) = getfield(integ, :timestepper)
"""`JLD2Writer(integ, output_functions(integ, (:temperature, :saturation_water_ice)); filename, schedule)`."""
output_functions(integ::B200Integrator, names) = NamedTuple{Tuple(names)}(Tuple((m -> getproperty(m.state, n)) for n in names))

function Terrarium.initialize(model::Terrarium.AbstractModel{NF, <:B200Grid}, timestepper::Terrarium.AbstractTimeStepper,
                              inputs::Terrarium.InputSource...; boundary_conditions = (;), initializers = (;), kwargs...) where {NF}
    grid = model.grid
    integ = initialize_b200(host_model(model), timestepper; device = grid.device, math = grid.math, boundary_conditions, initializers)
    for source in inputs
        upload_input!(integ, source)
    end
    if grid.host isa Terrarium.ColumnRingGrid
        set_ring_mask!(integ, vec(Array(grid.host.mask)))
    end
    return integ
end

# InputSource{NF, name}: static fields are copied once, FieldTimeSeries become device tables (input_sources.jl:81-171)
function upload_input!(integ::B200Integrator, source::Terrarium.FieldInputSource{NF, name}) where {NF, name}
    set_input!(integ, name, vec(Array(Terrarium.interior(source.field))))
end
function upload_input!(integ::B200Integrator, source::Terrarium.FieldTimeSeriesInputSource{NF, name}) where {NF, name}
    fts = source.fts
    values = reduce(hcat, [vec(Array(Terrarium.interior(fts[i]))) for i in 1:length(fts.times)])   # [column, time]
    set_input!(integ, name, TimeSeries(fts.times, values))
end
upload_input!(::B200Integrator, source) =
    throw(ArgumentError("$(typeof(source)) has no device-resident form; pass the raster as TerrariumB200.TimeSeries(times, values; raster = true)"))

function __init__()
    v = ccall((:trm_abi_version, LIB), Cint, ())
    v == TRM_ABI_VERSION || error("libterrarium_b200 has ABI version $v, this module was written for $TRM_ABI_VERSION")
end

end # module
