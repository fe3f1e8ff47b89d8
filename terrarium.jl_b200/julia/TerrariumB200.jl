# TerrariumB200.jl -- Julia host glue: Terrarium's `initialize` / `timestep!` / `run!` on top of
# libterrarium_b200.so (include/terrarium_b200.h) through `ccall`.
#
# STATUS: UNTESTED SOURCE.  Neither `julia` nor Terrarium's dependencies exist in the build image or
# on the GPU boxes, so this file has never been executed; the tested caller of the same C ABI is
# the Python ctypes mirror (terrarium.jl_b200/integrator.py).  The struct layouts below must match
# include/terrarium_b200.h field for field (TRM_ABI_VERSION = 2).
#
# What it replaces in the reference (paths relative to the Terrarium.jl root):
#   initialize(model, timestepper, inputs...)            src/timesteppers/model_integrator.jl:145-161
#   timestep!(integrator, ::ForwardEuler | ::Heun, dt)   src/timesteppers/forward_euler.jl:19-31, heun.jl:37-71
#   run!(integrator; steps, period, dt)                  src/timesteppers/model_integrator.jl:72-88
#   interior(field) / set!(field, value)                 src/state_variables.jl:476-489, src/initializers.jl:23-27
module TerrariumB200

using Terrarium
using Terrarium: SoilModel, LandModel, ForwardEuler, Heun, RichardsEq, NoFlow, ConstantSoilHydraulics,
                 UnsatKVanGenuchten, UnsatKLinear, default_dt, convert_dt, get_steps
import FreezeCurves: VanGenuchten, BrooksCorey

const LIB = get(ENV, "TERRARIUM_B200_LIB", joinpath(@__DIR__, "..", "csrc", "libterrarium_b200.so"))

const TRM_ABI_VERSION = Int32(2)
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
end
Base.eltype(::B200Integrator{NF}) where {NF} = NF

dtype_code(::Type{Float32}) = Int32(0)
dtype_code(::Type{Float64}) = Int32(1)

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
        hp.sat_hydraulic_cond, ustrip(vg.α), vg.n, ustrip(bc.ψₛ), bc.λ, 0.0,
        hp.unsat_hydraulic_cond isa UnsatKVanGenuchten ? hp.unsat_hydraulic_cond.impedance : 7.0, 0.0,
        land ? seb.albedo.albedo : 0.3, land ? seb.albedo.emissivity : 0.97, land ? seb.skin_temperature.κₛ : 2.0,
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
    ref = Terrarium.initialize(model, timestepper; boundary_conditions, initializers)   # CPU reference state at t0
    grid = Terrarium.get_field_grid(Terrarium.get_grid(model))
    zf = collect(Float64, Terrarium.znodes(grid, Terrarium.Face()))
    nz, ncol = length(zf) - 1, size(grid, 1)
    hyd = model.soil.hydrology
    bcs = ntuple(_ -> TrmBC(0, 0), TRM_BC_NSLOTS)   # TODO(julia side): translate `boundary_conditions` into slots + inputs
    cfg = Ref(TrmConfig(TRM_ABI_VERSION, dtype_code(NF), ncol, 0, nz, device,
        model isa LandModel ? 1 : 0, timestepper isa Heun ? 1 : 0, hyd.vertical_flow isa RichardsEq ? 1 : 0,
        hyd.hydraulic_properties.swrc isa VanGenuchten ? 0 : 1, hyd.hydraulic_properties.unsat_hydraulic_cond isa UnsatKVanGenuchten ? 1 : 0,
        0, 0, math === :fast ? 1 : 0, (model isa LandModel && !isnothing(model.vegetation)) ? 1 : 0,
        (model isa LandModel && model.surface_hydrology.evapotranspiration.ground_resistance isa Terrarium.SoilMoistureResistanceFactor) ? 1 : 0,
        pointer(zf), params_of(model), bcs))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve zf check(ccall((:trm_create, LIB), Cint, (Ref{TrmConfig}, Ref{Ptr{Cvoid}}), cfg, h), "create")
    integ = B200Integrator{NF, typeof(model), typeof(timestepper)}(h[], model, timestepper, ncol, nz)
    finalizer(i -> ccall((:trm_destroy, LIB), Cint, (Ptr{Cvoid},), i.handle), integ)
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

function Terrarium.timestep!(integ::B200Integrator, Δt = default_dt(integ.timestepper); finalize = true)
    check(ccall((:trm_step, LIB), Cint, (Ptr{Cvoid}, Cdouble, Int64), integ.handle, convert_dt(Δt), 1), "step")
    finalize && check(ccall((:trm_compute_auxiliary, LIB), Cint, (Ptr{Cvoid},), integ.handle), "compute_auxiliary")
    return nothing
end

function Terrarium.run!(integ::B200Integrator; steps = nothing, period = nothing, Δt = default_dt(integ.timestepper))
    Δt = convert_dt(Δt)
    n = get_steps(steps, period, Δt)
    check(ccall((:trm_step, LIB), Cint, (Ptr{Cvoid}, Cdouble, Int64), integ.handle, Δt, n), "step")          # n fused stage launches
    check(ccall((:trm_compute_auxiliary, LIB), Cint, (Ptr{Cvoid},), integ.handle), "compute_auxiliary")
    return integ
end

function Terrarium.current_time(integ::B200Integrator)
    t, it = Ref{Cdouble}(0), Ref{Int64}(0)
    check(ccall((:trm_get_clock, LIB), Cint, (Ptr{Cvoid}, Ref{Cdouble}, Ref{Int64}), integ.handle, t, it), "get_clock")
    return t[]
end

end # module
