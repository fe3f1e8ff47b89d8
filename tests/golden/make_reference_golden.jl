# make_reference_golden.jl -- golden vectors from the UNMODIFIED reference (Terrarium.jl), for tests/test_reference_golden.py.
#
#     julia --project=<a Terrarium.jl checkout or environment> tests/golden/make_reference_golden.jl [outdir]
#
# STATUS: never executed by the authors of this repository (no `julia` in the build image or on the GPU boxes). It only uses the
# reference's public API as its own tests and examples use it (citations per case). A maintainer with a Terrarium install runs it
# once; it writes `tests/golden/reference/*.npy` (plain NumPy files, no Julia package beyond Terrarium needed). The consumer test
# replays every case through the C ABI (oracle and CUDA path) and compares at the tolerance of BASELINE.json (Float64, 1e-9).
#
# Cases
#   cfg1a / cfg1b  BASELINE config 1: quick-start column (README.md:85-95 ; examples/simulations/soil_heat_column.jl:13-43)
#   cfg2           BASELINE config 2: heat-only SoilModel, per-column periodic surface temperature
#                  (examples/simulations/soil_heat_global.jl:51-93, on a ColumnGrid so that no mask file is needed)
#   cfg3e / cfg3h  BASELINE config 3: soil energy + Richards (test/soil/soil_hydrology_tests.jl:125-189), ForwardEuler / Heun
#   cfg4b / cfg4v  BASELINE config 4: LandModel bare ground / with VegetationCarbon, Heun (test/coupled_models/land_model_tests.jl:6-71)
#   sem_*          the semantics that live in un-vendored dependencies (SURVEY.md Appendix B), one by one:
#                  Value-BC halo (test/boundary_conditions.jl:16-19), NoFlow saturation halo for constant / function initializers
#                  (examples/simulations/soil_heat_column.jl:18-22), inv(swrc) values (soil_hydraulic_closures.jl:95-118), top
#                  Flux-BC magnitude (src/timesteppers/abstract_timestepper.jl:69)
using Terrarium
using Terrarium: fill_halo_regions!, compute_auxiliary!, compute_tendencies!

const OUT = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "reference")
mkpath(OUT)

# ---- NumPy .npy writer (format 1.0, little endian, Fortran order) ----
function write_npy(name::AbstractString, A::AbstractArray{T}) where {T <: Union{Float64, Float32, Int64}}
    descr = T === Float64 ? "<f8" : T === Float32 ? "<f4" : "<i8"
    shape = join(size(A), ", ") * (ndims(A) == 1 ? "," : "")
    hdr = "{'descr': '$descr', 'fortran_order': True, 'shape': ($shape), }"
    hdr *= " "^(64 - mod(10 + length(hdr) + 1, 64)) * "\n"
    open(joinpath(OUT, name * ".npy"), "w") do io
        write(io, 0x93); write(io, "NUMPY"); write(io, 0x01); write(io, 0x00); write(io, UInt16(length(hdr))); write(io, hdr)
        write(io, Array(A))
    end
    return nothing
end
write_npy(name, x::Number) = write_npy(name, [Float64(x)])

# interior of a Field as [column, layer] (3-D) or [column] (2-D)
function host(f)
    A = Array(interior(f))
    return size(A, 3) == 1 ? vec(A) : A[:, 1, :]
end
function dump(prefix, state, names)
    for n in names
        write_npy("$(prefix)_$(n)", Float64.(host(getproperty(state, n))))
    end
end
zfaces(grid) = collect(Float64, znodes(Terrarium.get_field_grid(grid), Face()))

const SOIL_OUT = (:temperature, :internal_energy, :saturation_water_ice, :liquid_water_fraction)
const RICH_OUT = (SOIL_OUT..., :pressure_head, :water_table, :surface_excess_water, :hydraulic_conductivity)
const LAND_OUT = (RICH_OUT..., :skin_temperature, :ground_heat_flux, :sensible_heat_flux, :latent_heat_flux, :surface_net_radiation,
                  :evaporation_ground, :infiltration, :surface_runoff)
const VEG_OUT = (LAND_OUT..., :carbon_vegetation, :vegetation_area_fraction, :canopy_water, :net_assimilation, :transpiration)

# deterministic per-column parameters (written out; the consumer reads them back)
percol(Nc) = (lat = collect(range(-1.5, 1.5, length = Nc)), lon = mod.(2.399963229728653 .* (1:Nc), 2π))
mean_annual_temperature(lat) = 20 - abs(40 * sin(lat))            # examples/simulations/soil_heat_global.jl:51

richards_soil(NF) = let swrc = VanGenuchten(α = 2.0, n = 2.0)      # test/soil/soil_hydrology_tests.jl:127-130
    hp = ConstantSoilHydraulics(NF; swrc, unsat_hydraulic_cond = UnsatKVanGenuchten(NF))
    SoilEnergyWaterCarbon(NF; hydrology = SoilHydrology(NF, RichardsEq(); hydraulic_properties = hp))
end

# ---------------------------------------------------------------------------------------------------------------------
# cfg1: quick-start column, 1000 steps of 300 s, Float64
# ---------------------------------------------------------------------------------------------------------------------
let NF = Float64
    grid = ColumnGrid(CPU(), NF, ExponentialSpacing(N = 10))
    write_npy("cfg1_zfaces", zfaces(grid))
    # (a) README variant: DefaultInitializer (T = 0, sat = 0), constant surface temperature 1 degC
    integ = initialize(SoilModel(grid), ForwardEuler(NF); boundary_conditions = PrescribedSurfaceTemperature(:T_ub, 1.0))
    run!(integ; steps = 1000, Δt = 300.0)
    dump("cfg1a", integ.state, SOIL_OUT)
    # (b) freeze-thaw variant, examples/simulations/soil_heat_column.jl:13-43
    initializer = SoilInitializer(NF, energy = QuasiThermalSteadyState(NF, T₀ = -1.0), hydrology = ConstantSaturation(NF, sat = 1.0))
    integ = initialize(SoilModel(grid; initializer), ForwardEuler(NF); boundary_conditions = PrescribedSurfaceTemperature(:T_ub, 1.0))
    dump("cfg1b_initial", integ.state, SOIL_OUT)
    run!(integ; steps = 1000, Δt = 300.0)
    dump("cfg1b", integ.state, SOIL_OUT)
end

# ---------------------------------------------------------------------------------------------------------------------
# cfg2 / cfg3: 24 columns x 30 layers, per-column surface temperature T0 + 10 sin(2 pi t / day - lon) handed over as an input
# variable every step (the coupling flow of examples/simulations/speedy_dry_land.jl:54-57), 1000 steps
# ---------------------------------------------------------------------------------------------------------------------
function run_soil(prefix, soil_of, stepper_of, Δt, richards; Nc = 24, nsteps = 1000)
    NF = Float64
    grid = ColumnGrid(CPU(), NF, ExponentialSpacing(Δz_min = 0.05, Δz_max = 100.0, N = 30), Nc)
    (; lat, lon) = percol(Nc)
    T0 = mean_annual_temperature.(lat)
    write_npy("$(prefix)_zfaces", zfaces(grid)); write_npy("$(prefix)_lon", lon); write_npy("$(prefix)_T0", T0)
    model = isnothing(soil_of) ? SoilModel(grid) : SoilModel(grid; soil = soil_of(NF))
    T_ub = Field(grid, XY())                                      # src/grids/grid_utils.jl:48-68
    source = InputSource(grid, T_ub; name = :T_ub)
    col(x) = clamp(round(Int, x * Nc + 0.5), 1, Nc)               # x-node of column i on a ColumnGrid is (i - 1/2) / Nc
    initializers = richards ?
        (temperature = (x, z) -> T0[col(x)] - 0.05 * z, saturation_water_ice = (x, z) -> min(1, 0.5 - 0.1 * z)) :
        (temperature = (x, z) -> T0[col(x)] - 0.05 * z, saturation_water_ice = (x, z) -> 1.0)
    integ = initialize(model, stepper_of(NF), source; boundary_conditions = PrescribedSurfaceTemperature(:T_ub), initializers)
    dump("$(prefix)_initial", integ.state, richards ? RICH_OUT[1:7] : SOIL_OUT)
    for i in 0:(nsteps - 1)
        t = i * Δt
        set!(integ.state.inputs.T_ub, reshape(T0 .+ 10.0 .* sin.(2π * t / 86400.0 .- lon), Nc, 1))
        timestep!(integ, Δt; finalize = false)
    end
    compute_auxiliary!(integ.state, integ.model)
    dump(prefix, integ.state, richards ? RICH_OUT : SOIL_OUT)
    write_npy("$(prefix)_time", integ.clock.time)
end
run_soil("cfg2", nothing, ForwardEuler, 300.0, false)
run_soil("cfg3e", richards_soil, ForwardEuler, 60.0, true)
run_soil("cfg3h", richards_soil, Heun, 60.0, true)

# ---------------------------------------------------------------------------------------------------------------------
# cfg4: LandModel, Heun, synthetic atmosphere (BASELINE.md section 5) set as input fields every step
# ---------------------------------------------------------------------------------------------------------------------
function run_land(prefix, vegetated, windspeed, nsteps; Nc = 16, Δt = 60.0)
    NF = Float64
    grid = ColumnGrid(CPU(), NF, ExponentialSpacing(Δz_min = 0.05, Δz_max = 100.0, N = 30), Nc)
    (; lat, lon) = percol(Nc)
    T0 = mean_annual_temperature.(lat)
    write_npy("$(prefix)_zfaces", zfaces(grid)); write_npy("$(prefix)_lon", lon); write_npy("$(prefix)_T0", T0)
    # turnover rates: the as-coded per-year rates act per second (carbon_dynamics.jl:33-43) and empty the carbon pool within a
    # minute; the long run uses rates that keep the state in a physical range (same values as tests/test_vegetation.py)
    vegetation = VegetationCarbon(NF; carbon_dynamics = PALADYNCarbonDynamics(NF; γL = 1.0e-9, γR = 1.0e-9, γS = 1.0e-10),
                                  vegetation_dynamics = PALADYNVegetationDynamics(NF; γv_min = 1.0e-8))
    land = vegetated ? LandModel(grid; soil = richards_soil(NF), vegetation) :
                       LandModel(grid; soil = richards_soil(NF), vegetation = nothing)
    col(x) = clamp(round(Int, x * Nc + 0.5), 1, Nc)
    base = (temperature = (x, z) -> T0[col(x)] - 0.05 * z, saturation_water_ice = (x, z) -> min(1, 0.5 - 0.1 * z),
            skin_temperature = (x,) -> T0[col(x)])
    initializers = vegetated ? merge(base, (carbon_vegetation = 10.0, vegetation_area_fraction = 0.5)) : base
    integ = initialize(land, Heun(NF); initializers)
    st = integ.state
    for (name, value) in (:surface_longwave_down => 300.0, :specific_humidity => 0.005, :air_pressure => 101325.0, :windspeed => windspeed)
        set!(getproperty(st.inputs, name), value)
    end
    vegetated && (set!(st.inputs.CO2, 400.0); set!(st.inputs.SAI, 0.5))
    for i in 0:(nsteps - 1)
        t = i * Δt
        set!(st.inputs.air_temperature, reshape(T0 .+ 8.0 .* sin.(2π * t / 86400.0 .- lon), Nc, 1))
        set!(st.inputs.surface_shortwave_down, reshape(max.(0.0, 600.0 .* sin.(2π * t / 86400.0 .- lon)), Nc, 1))
        set!(st.inputs.rainfall, mod(t, 86400.0) < 6 * 3600.0 ? 2.0e-8 : 0.0)
        timestep!(integ, Δt; finalize = false)
    end
    compute_auxiliary!(integ.state, integ.model)
    dump(prefix, integ.state, vegetated ? VEG_OUT : LAND_OUT)
end
run_land("cfg4b", false, 0.5, 1000)
run_land("cfg4b_v3", false, 3.0, 60)        # the BASELINE wind speed for the first simulated hour (diverges later, DESIGN.md section 7)
run_land("cfg4v", true, 0.5, 1000)

# ---------------------------------------------------------------------------------------------------------------------
# sem_*: un-vendored semantics, one by one
# ---------------------------------------------------------------------------------------------------------------------
let NF = Float64
    # (1) Value-BC halo: interior 0.5, boundary value 1.0 -> halo value (test/boundary_conditions.jl:16-19 expects 1.5),
    #     on a stretched grid as well (halo spacing rule)
    for (tag, spacing) in (("uniform", UniformSpacing(N = 10)), ("stretched", ExponentialSpacing(N = 10)))
        grid = ColumnGrid(CPU(), NF, spacing)
        integ = initialize(SoilModel(grid), ForwardEuler(NF); boundary_conditions = PrescribedSurfaceTemperature(:T_ub, 1.0))
        T = integ.state.temperature
        set!(T, 0.5)
        fill_halo_regions!(integ.state)
        Nz = size(interior(T), 3)
        write_npy("sem_value_bc_halo_$(tag)", [T[1, 1, Nz], T[1, 1, Nz + 1], T[1, 1, 0], T[1, 1, 1]])
    end
    # (2) NoFlow: what does the z-halo of the auxiliary saturation field hold after a constant / a function initializer, and
    #     what top-layer energy tendency follows (the top-face conductivity averages the halo cell)
    for (tag, init) in (("constant", 1.0), ("function", (x, z) -> 1.0))
        grid = ColumnGrid(CPU(), NF, ExponentialSpacing(N = 10))
        integ = initialize(SoilModel(grid), ForwardEuler(NF); boundary_conditions = PrescribedSurfaceTemperature(:T_ub, 1.0),
                           initializers = (temperature = -1.0, saturation_water_ice = init))
        s = integ.state.saturation_water_ice
        Nz = size(interior(s), 3)
        Terrarium.update_state!(integ.state, integ.model, integ.inputs)
        write_npy("sem_noflow_sat_halo_$(tag)", [s[1, 1, Nz + 1], s[1, 1, 0], integ.state.tendencies.internal_energy[1, 1, Nz],
                                                 integ.state.tendencies.internal_energy[1, 1, 1]])
        timestep!(integ, 300.0)
        dump("sem_noflow_sat_halo_$(tag)_step1", integ.state, SOIL_OUT)
    end
    # (3) inverse retention curves at 20 water contents, theta_sat = 0.49 (soil_hydraulic_closures.jl:115-118)
    θ = collect(range(0.02, 0.49, length = 20))
    for (tag, swrc) in (("vangenuchten_a2_n2", VanGenuchten(α = 2.0, n = 2.0)), ("vangenuchten_default", VanGenuchten()), ("brookscorey_default", BrooksCorey()))
        ψ = [ustrip(inv(swrc)(θi; θsat = 0.49)) for θi in θ]
        write_npy("sem_swrc_inverse_$(tag)", hcat(θ, Float64.(ψ)))
        write_npy("sem_swrc_forward_$(tag)", hcat(Float64.(ψ), [ustrip(swrc(ψi; θsat = 0.49)) for ψi in ψ]))
    end
    # (4) magnitude and sign of a top Flux BC on the internal energy: one ForwardEuler step of a thawed, uniform column
    #     (no conduction: T uniform) with GroundHeatFlux = 10 W/m2 -> dU_top = ? (compute_z_bcs!, abstract_timestepper.jl:69)
    grid = ColumnGrid(CPU(), NF, UniformSpacing(Δz = 0.1, N = 10))
    integ = initialize(SoilModel(grid), ForwardEuler(NF); boundary_conditions = GroundHeatFlux(10.0),
                       initializers = (temperature = 5.0, saturation_water_ice = (x, z) -> 1.0))
    U0 = copy(host(integ.state.internal_energy))
    timestep!(integ, 60.0)
    write_npy("sem_flux_bc_energy_top", hcat(vec(U0), vec(host(integ.state.internal_energy))))
    # same for the infiltration flux on the saturation of a Richards column (InfiltrationFlux, soil_model_bcs.jl:29)
    model = SoilModel(grid; soil = richards_soil(NF))
    integ = initialize(model, ForwardEuler(NF); boundary_conditions = InfiltrationFlux(-1.0e-8),
                       initializers = (temperature = 5.0, saturation_water_ice = (x, z) -> 0.5))
    s0 = copy(host(integ.state.saturation_water_ice))
    timestep!(integ, 60.0)
    write_npy("sem_flux_bc_saturation_top", hcat(vec(s0), vec(host(integ.state.saturation_water_ice))))
end

println("wrote golden vectors of the reference to ", OUT)
