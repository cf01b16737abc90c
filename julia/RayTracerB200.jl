# RayTracerB200.jl -- thin `ccall` front of librt_sssp.so that re-creates the exported interface of
# RayTracer.jl (src/RayTracer.jl:24-34) for the shortest-path-method hot path.  No CUDA.jl, no kernels in
# Julia: every call lands in the C ABI declared in include/rt_sssp.h.
#
#     include("julia/RayTracerB200.jl"); using .RayTracerB200
#     gr, G, halo = init_annulus(180, 50; spacing = 1)
#     source      = closest_point(gr, 0.0, R; system = :polar)
#     Vp          = interpolate_velocity(gr.r, LinearInterpolation(profile.r, profile.Vp))
#     D           = bfm(G, halo, source, gr, Vp)          # D.dist :: Vector{Float64}, D.prev :: Vector{Int64}
#     path        = recontruct_path(D.prev, source, receiver)
#
# NOTE: Julia is not installed in the build image, so this file is the binding a maintainer would add; the
# same entry points are exercised by the Python ctypes front (raytracer.jl_b200/api.py) in tests/ and bench.py.
module RayTracerB200

using SparseArrays

export Grid2D, BellmanFordMoore, R, init_annulus, closest_point, interpolate_velocity, bfm, recontruct_path,
       LinearInterpolation, bfm_batch, bfm_batch_multi, set_device, bfm_gpu, interpolate!, symrcm, nodal_degree,
       dual_velocity, SparseAdjencyList, sparse_adjacency_list, travel_times, set_schedule!, Grid3D, grid, coordinates, BFM,
       AbstractSPM, Dijkstra, RadiusStepping, dijkstra, radius_stepping, Point, connectivity, CartesianIndex,
       polardistance3D, set_option!, GridPartition, partition_grid, directions, bfm_continue, bfm_multiphase,
       comm_unique_id, comm_init, comm_shard, comm_destroy, bfm_batch_sharded

const R = 6371.0                                   # src/utils.jl:2
const LIB = get(ENV, "RT_SSSP_LIB", joinpath(@__DIR__, "..", "raytracer.jl_b200", "librt_sssp.so"))

struct RtStats                                     # mirrors rt_stats (include/rt_sssp.h)
    sweeps::Int64
    relaxed_edges::Int64
    vertex_updates::Int64
    graph_edges::Int64
    kernel_ms::Float64
    relax_ms::Float64
    relax_launches::Int64
    total_launches::Int64
    prev_ms::Float64
    screened_edges::Int64
    exact_edges::Int64
end

function check(rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:rt_last_error, LIB), Cstring, ()))
    rc == 3 && throw(BoundsError())                # interpolation outside the knots, as Interpolations.jl throws
    error("rt_sssp error $rc: $msg")
end

mutable struct MeshHandle
    ptr::Ptr{Cvoid}
    function MeshHandle(p::Ptr{Cvoid})
        h = new(p)
        finalizer(x -> (x.ptr != C_NULL && ccall((:rt_mesh_free, LIB), Cint, (Ptr{Cvoid},), x.ptr); x.ptr = C_NULL), h)
        return h
    end
end

# Grid2D of src/GridAnnulus.jl:9-21 plus the native handle
mutable struct Grid2D
    x::Vector{Float64}
    z::Vector{Float64}
    θ::Vector{Float64}
    r::Vector{Float64}
    e2n::Dict{Int,Vector{Int64}}
    nθ::Int64
    nr::Int64
    nel::Int64
    nnods::Int64
    neighbours::Vector{Vector{Int64}}
    element_type::Dict{Int,Symbol}
    handle::Union{Nothing,MeshHandle}
end
Base.length(gr::Grid2D) = gr.nnods

abstract type AbstractSPM end                      # src/SSSP/ssspm.jl:1
for algorithm in (:BellmanFordMoore, :Dijkstra, :RadiusStepping)   # src/SSSP/ssspm.jl:3-10
    @eval begin
        struct $(algorithm){T,M} <: AbstractSPM
            prev::T
            dist::M
        end
    end
end
Base.getindex(spm::AbstractSPM) = spm.prev         # src/SSSP/ssspm.jl:12

struct LinearInterpolation                         # stand-in for Interpolations.LinearInterpolation (README.md:32)
    knots::Vector{Float64}
    values::Vector{Float64}
end

# init_annulus(nθ, nr; spacing) -- src/GridAnnulus.jl:57-70, built by CUDA kernels
function init_annulus(nθ::Int64, nr::Int64; spacing = 20)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rt_annulus_build, LIB), Cint, (Int64, Int64, Cdouble, Ref{Ptr{Cvoid}}), nθ, nr, Float64(spacing), out))
    h = MeshHandle(out[])
    sz = zeros(Int64, 8)
    check(ccall((:rt_mesh_sizes, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}), h.ptr, sz))
    n, nel, se, nnz, hr, sn = sz[1], sz[2], sz[3], sz[4], sz[5], sz[6]
    x, z, θ, r = zeros(n), zeros(n), zeros(n), zeros(n)
    e2n_off, e2n_idx = zeros(Int64, nel + 1), zeros(Int64, se)
    colptr, rowval = zeros(Int64, n + 1), zeros(Int64, nnz)
    halo = zeros(Int64, hr, 2)                     # column-major (2H x 2) exactly as the ABI writes it
    nbr_off, nbr_idx, el_type = zeros(Int64, nel + 1), zeros(Int64, sn), zeros(Int8, nel)
    check(ccall((:rt_mesh_export, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Ptr{Int64},
                 Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int8}),
                h.ptr, x, z, θ, r, e2n_off, e2n_idx, colptr, rowval, halo, nbr_off, nbr_idx, el_type))
    e2n = Dict{Int,Vector{Int64}}(i => e2n_idx[(e2n_off[i] + 1):e2n_off[i + 1]] for i in 1:nel)
    neighbours = [nbr_idx[(nbr_off[i] + 1):nbr_off[i + 1]] for i in 1:nel]
    etype = Dict{Int,Symbol}(i => (el_type[i] == 0 ? :Quad : :Tri) for i in 1:nel)
    gr = Grid2D(x, z, θ, r, e2n, sz[7], sz[8], nel, n, neighbours, etype, h)
    G = SparseMatrixCSC{Bool,Int64}(nel, n, colptr, rowval, fill(true, nnz))
    return gr, G, halo
end

# adopt a graph built by the reference itself (gr without a native handle)
function mesh_handle(G::SparseMatrixCSC{Bool,Int64}, halo::Matrix, gr)
    hasproperty(gr, :handle) && gr.handle !== nothing && return gr.handle
    nel = gr.nel
    e2n_off = zeros(Int64, nel + 1)
    for i in 1:nel
        e2n_off[i + 1] = e2n_off[i] + length(gr.e2n[i])
    end
    e2n_idx = reduce(vcat, (gr.e2n[i] for i in 1:nel))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    halo64 = Matrix{Int64}(halo)
    check(ccall((:rt_mesh_from_arrays, LIB), Cint,
                (Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int64, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
                G.n, nel, e2n_off, e2n_idx, G.colptr, G.rowval, halo64, size(halo64, 1), gr.x, gr.z, gr.θ, gr.r, out))
    h = MeshHandle(out[])
    hasproperty(gr, :handle) && (gr.handle = h)
    return h
end

# interpolate_velocity(r, interpolant) -- src/utils.jl:38-44 (buffer kw: src/ShortestPath.jl:74-90)
function interpolate_velocity(r::AbstractArray, itp::LinearInterpolation; buffer = nothing)
    V = similar(r, Float64)
    rr = Vector{Float64}(vec(r))
    check(ccall((:rt_interp_velocity, LIB), Cint,
                (Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Cdouble, Ptr{Float64}),
                itp.knots, itp.values, length(itp.knots), rr, length(rr), buffer === nothing ? -1.0 : Float64(buffer), V))
    return V
end

# closest_point(gr, px, pz; system) -- src/GridAnnulus.jl:823-840
function closest_point(gr::Grid2D, px, pz; system = :cartesian)
    out = zeros(Int64, 1)
    check(ccall((:rt_closest_point, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Cint, Ptr{Int64}),
                gr.handle.ptr, [Float64(px)], [Float64(pz)], 1, system == :cartesian ? 0 : 1, out))
    return out[1]
end

# batched form (receiver sweeps): one device pass per query, any number of queries
function closest_point(gr::Grid2D, px::AbstractVector, pz::AbstractVector; system = :cartesian)
    out = zeros(Int64, length(px))
    check(ccall((:rt_closest_point, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Cint, Ptr{Int64}),
                gr.handle.ptr, Vector{Float64}(px), Vector{Float64}(pz), length(px), system == :cartesian ? 0 : 1, out))
    return out
end

# generic solver / mesh option (include/rt_sssp.h: "schedule", "canonical_prev", "weight3d", "delta", ...)
function set_option!(gr, key::AbstractString, value::Real)
    check(ccall((:rt_set_option, LIB), Cint, (Ptr{Cvoid}, Cstring, Cdouble), gr.handle.ptr, key, Float64(value)))
    return gr
end

# bfm(G, halo, source, gr, U) -- src/SSSP/bfm.jl:1-52
function bfm(G::SparseMatrixCSC{Bool,M}, halo::Matrix, source::Integer, gr, U::AbstractArray{T}) where {M,T}
    D, st = bfm_batch(G, halo, Int64[source], gr, U)
    println("Converged in $(st.sweeps + 1) iterations")          # bfm.jl:49 (it starts at 1)
    return BellmanFordMoore(D.prev[:, 1], D.dist[:, 1])
end

# bfm_gpu(G, halo, source, gr, U) -- src/SSSP/bfm_gpu.jl:212-250: same call, Float32 arithmetic (precision = 32);
# D.dist::Vector{Float32}, D.prev::Vector{Int32} as in the reference's device path
function bfm_gpu(G::SparseMatrixCSC{Bool,M}, halo::Matrix, source::Int, gr, U::AbstractArray{T}) where {M,T}
    D, st = bfm_batch(G, halo, Int64[source], gr, U; precision = 32)
    println("Converged in $(st.sweeps + 1) iterations")
    return BellmanFordMoore(Int32.(D.prev[:, 1]), Float32.(D.dist[:, 1]))   # exact: the values are Float32 numbers
end

# batch API: many earthquakes on one mesh; tables are n x nsrc (column per source)
function bfm_batch(G::SparseMatrixCSC{Bool,Int64}, halo::Matrix, sources::Vector{Int64}, gr, U::AbstractArray;
                   precision::Integer = 64)
    h = mesh_handle(G, halo, gr)
    n, ns = G.n, length(sources)
    dist = Matrix{Float64}(undef, n, ns)
    prev = Matrix{Int64}(undef, n, ns)
    st = Ref(RtStats(0, 0, 0, 0, 0.0, 0.0, 0, 0, 0.0, 0, 0))
    check(ccall((:rt_bfm_solve, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}, Int64, Cint, Ptr{Float64}, Ptr{Int64}, Ref{RtStats}),
                h.ptr, Vector{Float64}(U), sources, ns, precision, dist, prev, st))
    return BellmanFordMoore(prev, dist), st[]
end

# many earthquakes over several GPUs from one Julia process: `grids` are replicas of the same mesh, one per device
# (build each after `set_device(d)`); replica d solves a contiguous block of `sources` on its own host thread
set_device(d::Integer) = check(ccall((:rt_set_device, LIB), Cint, (Cint,), d))
function bfm_batch_multi(grids::Vector, sources::Vector{Int64}, U::AbstractArray; precision::Integer = 64)
    hs = Ptr{Cvoid}[g.handle.ptr for g in grids]
    n, ns = length(U), length(sources)
    dist = Matrix{Float64}(undef, n, ns)
    prev = Matrix{Int64}(undef, n, ns)
    st = Ref(RtStats(0, 0, 0, 0, 0.0, 0.0, 0, 0, 0.0, 0, 0))
    check(ccall((:rt_bfm_solve_multi, LIB), Cint,
                (Ptr{Ptr{Cvoid}}, Cint, Ptr{Float64}, Ptr{Int64}, Int64, Cint, Ptr{Float64}, Ptr{Int64}, Ref{RtStats}),
                hs, length(hs), Vector{Float64}(U), sources, ns, precision, dist, prev, st))
    return BellmanFordMoore(prev, dist), st[]
end

# many earthquakes with ONE PROCESS PER GPU (mpirun / Distributed.jl): rank 0 makes the 128-byte id with comm_unique_id()
# and ships it to the other ranks by whatever the driver already has; every rank then calls
#     set_device(local_rank); gr, G, halo = init_annulus(...); c = comm_init(id, rank, world)
#     D, st = bfm_batch_sharded(c, G, halo, sources, gr, U)      # the full [n x nsrc] tables on every rank
# rank r solves the contiguous block comm_shard(length(sources), r, world); the tables are gathered with one
# ncclAllGather each over NVLink inside the library (rt_bfm_solve_sharded).
mutable struct Comm
    ptr::Ptr{Cvoid}
end
function comm_unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:rt_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id))
    return id
end
function comm_init(id::Vector{UInt8}, rank::Integer, world::Integer)
    c = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rt_comm_init, LIB), Cint, (Ptr{UInt8}, Cint, Cint, Ref{Ptr{Cvoid}}), id, rank, world, c))
    return Comm(c[])
end
function comm_shard(nsrc::Integer, rank::Integer, world::Integer)
    first, count = Ref{Int64}(0), Ref{Int64}(0)
    check(ccall((:rt_comm_shard, LIB), Cint, (Int64, Cint, Cint, Ref{Int64}, Ref{Int64}), nsrc, rank, world, first, count))
    return (first[] + 1):(first[] + count[])
end
comm_destroy(c::Comm) = (c.ptr != C_NULL && ccall((:rt_comm_destroy, LIB), Cint, (Ptr{Cvoid},), c.ptr); c.ptr = C_NULL; nothing)
function bfm_batch_sharded(c::Comm, G::SparseMatrixCSC{Bool,Int64}, halo::Matrix, sources::Vector{Int64}, gr, U::AbstractArray;
                           precision::Integer = 64)
    h = mesh_handle(G, halo, gr)
    n, ns = G.n, length(sources)
    dist = Matrix{Float64}(undef, n, ns)
    prev = Matrix{Int64}(undef, n, ns)
    st = Ref(RtStats(0, 0, 0, 0, 0.0, 0.0, 0, 0, 0.0, 0, 0))
    check(ccall((:rt_bfm_solve_sharded_host, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}, Int64, Cint, Ptr{Float64}, Ptr{Int64}, Ref{RtStats}),
                c.ptr, h.ptr, Vector{Float64}(U), sources, ns, precision, dist, prev, st))
    return BellmanFordMoore(prev, dist), st[]
end

# dual_velocity(r, interpolant; buffer) -- src/utils.jl:51-66
function dual_velocity(r::AbstractArray, itp::LinearInterpolation; buffer = 1)
    V = Matrix{Float64}(undef, length(r), 2)
    check(ccall((:rt_dual_velocity, LIB), Cint,
                (Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Cdouble, Ptr{Float64}),
                itp.knots, itp.values, length(itp.knots), Vector{Float64}(vec(r)), length(r), Float64(buffer), V))
    return V
end

# bfm with U::Matrix -- the dual-velocity relax _relax!(..., U::Matrix) src/SSSP/bfm.jl:113-159
function bfm(G::SparseMatrixCSC{Bool,Int64}, halo::Matrix, source::Integer, gr, U::Matrix{Float64})
    h = mesh_handle(G, halo, gr)
    n = G.n
    dist = Vector{Float64}(undef, n)
    prev = Vector{Int64}(undef, n)
    st = Ref(RtStats(0, 0, 0, 0, 0.0, 0.0, 0, 0, 0.0, 0, 0))
    check(ccall((:rt_bfm_solve_dual, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}, Int64, Ptr{Float64}, Ptr{Int64}, Ref{RtStats}),
                h.ptr, U, Int64[source], 1, dist, prev, st))
    println("Converged in $(st[].sweeps + 1) iterations")
    return BellmanFordMoore(prev, dist)
end

# interpolate!(V, gr) -- src/Interpolations/interpolation.jl:5-18 (cell-wise bilinear / barycentric), in place
function interpolate!(V::Vector{Float64}, gr::Grid2D)
    et = Int8[gr.element_type[i] == :Quad ? 0 : 1 for i in 1:gr.nel]
    check(ccall((:rt_interpolate_cells, LIB), Cint, (Ptr{Cvoid}, Ptr{Int8}, Ptr{Float64}), gr.handle.ptr, et, V))
    return V
end

# symrcm(nodal_incidence(gr), degrees) -- src/SSSP/rcm.jl:2-46
function symrcm(gr::Grid2D)
    prm = zeros(Int64, gr.nnods)
    check(ccall((:rt_rcm, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}), gr.handle.ptr, prm))
    return prm
end

# nodal_degree(nodal_incidence(gr)) -- src/topology/topology.jl:70-77
function nodal_degree(gr::Grid2D)
    deg = zeros(Int64, gr.nnods)
    check(ccall((:rt_nodal_adjacency, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int64),
                gr.handle.ptr, deg, C_NULL, C_NULL, 0))
    return deg
end

# sparse_adjacency_list(nodal_incidence(gr)) -- src/topology/topology.jl:88-111 (Int32 CSR: list, deg, idx = 1 + cumsum)
struct SparseAdjencyList{T}
    list::Vector{T}
    deg::Vector{T}
    idx::Vector{T}
end
function sparse_adjacency_list(gr::Grid2D)
    n = gr.nnods
    deg, off = zeros(Int64, n), zeros(Int64, n + 1)
    check(ccall((:rt_nodal_adjacency, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int64),
                gr.handle.ptr, deg, off, C_NULL, 0))
    list = zeros(Int64, off[end])
    check(ccall((:rt_nodal_adjacency, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int64),
                gr.handle.ptr, C_NULL, C_NULL, list, length(list)))
    return SparseAdjencyList(Int32.(list), Int32.(deg), Int32.(off[1:n] .+ 1))
end

# travel_times(D, gr, receivers; isave, flname) -- src/utils.jl:4-15; the gather runs on the device (rt_travel_times), for
# a batch result (dist :: Matrix n x nsrc) it returns nrec x nsrc.  The CSV is written without DataFrames / CSV.jl.
function travel_times(D, gr, receivers; isave = false, flname = "")
    dist = D.dist
    n, ns = size(dist, 1), size(dist, 2)
    rec = Vector{Int64}(receivers)
    out = Matrix{Float64}(undef, length(rec), ns)
    check(ccall((:rt_travel_times, LIB), Cint, (Ptr{Float64}, Int64, Int64, Ptr{Int64}, Int64, Ptr{Float64}),
                Array{Float64}(dist), n, ns, rec, length(rec), out))
    travel_time = ns == 1 ? vec(out) : out
    if isave
        θ = rad2deg.(gr.θ[rec])
        open(joinpath(pwd(), flname), "w") do io
            println(io, "degree,travel_time")
            for (a, b) in zip(θ, travel_time)
                println(io, a, ",", b)
            end
        end
    end
    return travel_time
end

# solver schedule of a mesh: :jacobi = the reference's sweeps (default), Symbol("near-far") = work-efficient ordering
function set_schedule!(gr, schedule::Symbol)
    check(ccall((:rt_set_option, LIB), Cint, (Ptr{Cvoid}, Cstring, Cdouble), gr.handle.ptr, "schedule",
                schedule == :jacobi ? 0.0 : 1.0))
    return gr
end

# ---- 3-D structured grid: grid(c0, c1, nnods) src/StructuredGrid.jl:35-45, nodal_incidence :177-223 (implicit here),
# BFM(G, source, gr, U, fw) src/Dijsktra.jl:294-343 with the weight of src/SSSP/weights.jl:20
mutable struct Grid3D
    c0::NTuple{3,Float64}
    c1::NTuple{3,Float64}
    nnods::NTuple{3,Int64}
    neighbour_levels::Int
    handle::MeshHandle
end
Base.length(gr::Grid3D) = prod(gr.nnods)

function grid(c0, c1, nnods; neighbour_levels = 1, coord_system = :cartesian)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    a, b, nn = Float64[c0...], Float64[c1...], Int64[nnods...]
    check(ccall((:rt_grid3d_build, LIB), Cint, (Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Cint, Cint, Ref{Ptr{Cvoid}}),
                a, b, nn, neighbour_levels, coord_system == :cartesian ? 0 : 1, out))
    return Grid3D((a...,), (b...,), (nn...,), neighbour_levels, MeshHandle(out[]))
end

# Cartesian node coordinates (x fastest), what gr[I] of the reference evaluates lazily
function coordinates(gr::Grid3D)
    n = length(gr)
    X, Y, Z = zeros(n), zeros(n), zeros(n)
    check(ccall((:rt_grid3d_export, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                gr.handle.ptr, X, Y, Z))
    return X, Y, Z
end

function BFM(gr::Grid3D, source::Integer, U::AbstractVector; precision::Integer = 64)
    n = length(gr)
    dist = Vector{Float64}(undef, n)
    prev = Vector{Int64}(undef, n)
    st = Ref(RtStats(0, 0, 0, 0, 0.0, 0.0, 0, 0, 0.0, 0, 0))
    check(ccall((:rt_bfm_solve, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}, Int64, Cint, Ptr{Float64}, Ptr{Int64}, Ref{RtStats}),
                gr.handle.ptr, Vector{Float64}(U), Int64[source], 1, precision, dist, prev, st))
    return BellmanFordMoore(prev, dist)
end

# recontruct_path(prev, source, receiver) -- src/SSSP/ssspm.jl:30-40 (the misspelling is the reference's API)
function recontruct_path(prev::Vector, source, receiver)
    p64 = Vector{Int64}(prev)
    off = zeros(Int64, 2)
    rc = Int64[receiver]
    check(ccall((:rt_reconstruct_paths, LIB), Cint,
                (Ptr{Int64}, Int64, Int64, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Int64}, Int64),
                p64, length(p64), source, rc, 1, off, C_NULL, 0))
    path = zeros(Int64, off[2])
    check(ccall((:rt_reconstruct_paths, LIB), Cint,
                (Ptr{Int64}, Int64, Int64, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Int64}, Int64),
                p64, length(p64), source, rc, 1, off, path, length(path)))
    return Vector{Int}(path)
end
# recontruct_path(D, source, receiver) -- src/SSSP/ssspm.jl:14-28, the method on the result structs: the chase runs
# `while ipath ∉ path` (until a node repeats), then `source` is appended; an unset predecessor throws like the reference
function recontruct_path(D::AbstractSPM, source, receiver)
    p64 = Vector{Int64}(D.prev)
    off = zeros(Int64, 2)
    rc = Int64[receiver]
    st = ccall((:rt_reconstruct_paths_guarded, LIB), Cint,
               (Ptr{Int64}, Int64, Int64, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Int64}, Int64),
               p64, length(p64), source, rc, 1, off, C_NULL, 0)
    st == 4 && throw(BoundsError(D.prev, 0))
    check(st)
    path = zeros(Int64, off[2])
    check(ccall((:rt_reconstruct_paths_guarded, LIB), Cint,
                (Ptr{Int64}, Int64, Int64, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Int64}, Int64),
                p64, length(p64), source, rc, 1, off, path, length(path)))
    return Vector{Int}(path)
end

# ---- the rest of the 3-D grid surface: src/StructuredGrid.jl:27-31, 57-104, 121-168, 245-270
struct Point{T}
    x::T
    y::T
    z::T
end
function axes3(gr::Grid3D)
    x, y, z = zeros(gr.nnods[1]), zeros(gr.nnods[2]), zeros(gr.nnods[3])
    check(ccall((:rt_grid3d_axes, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), gr.handle.ptr, x, y, z))
    return x, y, z
end
function Base.getproperty(gr::Grid3D, s::Symbol)   # gr.x, gr.y, gr.z, gr.nels, gr.nxny like the reference's Grid
    s === :x && return axes3(gr)[1]
    s === :y && return axes3(gr)[2]
    s === :z && return axes3(gr)[3]
    s === :nels && return getfield(gr, :nnods) .- 1
    s === :nxny && return getfield(gr, :nnods)[1] * getfield(gr, :nnods)[2]
    return getfield(gr, s)
end
function Base.getindex(gr::Grid3D, I::Int)         # :77-81 (linear index, x fastest)
    xyz = zeros(3)
    st = ccall((:rt_grid3d_points, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Float64}, Ptr{Int64}),
               gr.handle.ptr, Int64[I], 1, xyz, C_NULL)
    st == 1 && throw(BoundsError(gr, I))
    check(st)
    return Point(xyz[1], xyz[2], xyz[3])
end
function Base.getindex(gr::Grid3D, I::Int, J::Int, K::Int)   # :57-62
    @assert I <= gr.nnods[1]
    @assert J <= gr.nnods[2]
    @assert K <= gr.nnods[3]
    x, y, z = axes3(gr)
    return Point(x[I], y[J], z[K])
end
function CartesianIndex(gr::Grid3D, I::Int)        # :90-96
    ijk = zeros(Int64, 3)
    check(ccall((:rt_grid3d_points, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Float64}, Ptr{Int64}),
                gr.handle.ptr, Int64[I], 1, C_NULL, ijk))
    return (ijk[1], ijk[2], ijk[3])
end
function connectivity(gr::Grid3D)                  # :121-142 -> Vector{NTuple{8,Int64}}
    nel = prod(gr.nels)
    e2n = Matrix{Int64}(undef, 8, nel)
    nel > 0 && check(ccall((:rt_grid3d_connectivity, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}), gr.handle.ptr, 1, nel, e2n))
    return [ntuple(k -> e2n[k, i], 8) for i in 1:nel]
end
function connectivity(gr::Grid3D, iel::Int)        # :146-168
    e = zeros(Int64, 8)
    check(ccall((:rt_grid3d_connectivity, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}), gr.handle.ptr, iel, 1, e))
    return ntuple(k -> e[k], 8)
end
function closest_point(gr::Grid3D, x::T, y::T, z::T) where {T}   # :257-270 (raw axis coordinates)
    out = zeros(Int64, 1)
    check(ccall((:rt_closest_point3d, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int64}),
                gr.handle.ptr, [Float64(x)], [Float64(y)], [Float64(z)], 1, out))
    return out[1]
end
function closest_point(gr::Grid3D, x::AbstractVector, y::AbstractVector, z::AbstractVector)   # receiver sweeps
    out = zeros(Int64, length(x))
    check(ccall((:rt_closest_point3d, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int64}),
                gr.handle.ptr, Vector{Float64}(x), Vector{Float64}(y), Vector{Float64}(z), length(x), out))
    return out
end
function polardistance3D(p1::Point, p2::Point)     # :245-249
    out = zeros(1)
    check(ccall((:rt_polardistance3d, LIB), Cint, (Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}),
                Float64[p1.x, p1.y, p1.z], Float64[p2.x, p2.y, p2.z], 1, out))
    return out[1]
end

# ---- alternative solvers behind the same structs: dijkstra(G, source, gr, U) src/SSSP/dijkstra.jl:68-136 and
# radius_stepping(Gsp, source, gr, U) src/SSSP/radius_stepping.jl:7-46 on the star-0 node graph nodal_incidence(gr)
# (the graph argument is implied by gr: pass `nothing` or the reference's own container)
function sssp_nodal(gr::Grid2D, source::Integer, U::AbstractVector, algorithm::Integer)
    n = gr.nnods
    dist, prev = Vector{Float64}(undef, n), Vector{Int64}(undef, n)
    st = Ref(RtStats(0, 0, 0, 0, 0.0, 0.0, 0, 0, 0.0, 0, 0))
    check(ccall((:rt_sssp_nodal, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Cint, Ptr{Float64}, Ptr{Int64}, Ref{RtStats}),
                gr.handle.ptr, Vector{Float64}(U), source, algorithm, dist, prev, st))
    println("Converged in $(st[].sweeps + 1) iterations")
    return prev, dist
end
dijkstra(G, source::Int, gr::Grid2D, U::Vector) = Dijkstra(sssp_nodal(gr, source, U, 0)...)
radius_stepping(Gsp, source::Int, gr::Grid2D, U::Vector) = RadiusStepping(sssp_nodal(gr, source, U, 1)...)

# ---- multiphase: partition_grid / GridPartition src/topology/topology.jl:150-206, bfm_multiphase
# src/SSSP/bfm_multiphase.jl:30-156 (an unfinished draft in the reference: see rt_bfm_continue in include/rt_sssp.h)
struct GridPartition
    id::Vector{String}
    code::Vector{Int32}                            # k > 0 = "Layer_k", -k = "Boundary_k"
    rboundaries::NTuple{7,Float64}
    layers::NTuple{8,String}
    boundaries::NTuple{7,String}
    nlayers::Int
    nboundaries::Int
    iterator::Dict{Int,Tuple}
end
function partition_grid(gr::Grid2D)
    code = zeros(Int32, gr.nnods)
    check(ccall((:rt_partition_grid, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}), gr.handle.ptr, code))
    rl = (R - 20.0, R - 35.0, R - 210.0, R - 410.0, R - 660.0, R - 2740.0, R - 2891.5)
    layers = ntuple(i -> "Layer_$i", 8)
    boundaries = ntuple(i -> "Boundary_$i", 7)
    nl, nmax = 8, 15
    it = Dict{Int,Tuple}()
    it[1] = it[nmax] = (layers[1], boundaries[1])
    for i in 2:(nl - 1)
        it[i] = (layers[i], boundaries[i - 1], boundaries[i])
        it[nmax - i + 1] = (layers[i], boundaries[i - 1], boundaries[i])
    end
    it[nl] = (layers[end], boundaries[end])
    id = [c > 0 ? "Layer_$c" : "Boundary_$(-c)" for c in code]
    return GridPartition(id, code, rl, layers, boundaries, 8, 7, it)
end
function directions(nlayers)                       # bfm_multiphase.jl:2-14
    nmax = 2nlayers - 1
    d = Dict{Int,NTuple{2,Symbol}}()
    d[1] = d[nmax] = (:above, :above)
    for i in 2:(nlayers - 1)
        d[i] = d[nmax - i + 1] = (:below, :above)
    end
    d[nlayers] = (:below, :below)
    return d
end
function bfm_continue(G::SparseMatrixCSC{Bool,Int64}, halo::Matrix, gr, U::Vector{Float64}, allowed, seeds::Vector{Int64},
                      dist::Vector{Float64}, prev::Vector{Int64})
    h = mesh_handle(G, halo, gr)
    d, p = copy(dist), copy(prev)
    al = allowed === nothing ? C_NULL : Vector{UInt8}(allowed)
    st = Ref(RtStats(0, 0, 0, 0, 0.0, 0.0, 0, 0, 0.0, 0, 0))
    check(ccall((:rt_bfm_continue, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{UInt8}, Ptr{Int64}, Int64, Ptr{Float64}, Ptr{Int64}, Ref{RtStats}),
                h.ptr, U, al, seeds, length(seeds), d, p, st))
    return BellmanFordMoore(p, d), st[]
end
function bfm_multiphase(G::SparseMatrixCSC{Bool,Int64}, halo::Matrix, source::Int, gr, U::Vector{Float64},
                        partition::GridPartition, interpolant::LinearInterpolation; nphases = 3, buffer_zone = 1.0)
    n = G.n
    U = copy(U)
    rdir = directions(partition.nlayers)
    rb = Dict(a => b for (a, b) in zip(partition.boundaries, partition.rboundaries))
    code_of(nm) = (s = split(nm, "_"); s[1] == "Layer" ? parse(Int, s[2]) : -parse(Int, s[2]))
    bnodes = Dict(a => findall(partition.code .== code_of(a)) for a in partition.boundaries)
    dist = fill(typemax(Float64), n); dist[source] = 0.0
    prev = zeros(Int64, n)
    for i in 1:size(halo, 1)                        # init_halo_path! (bfm.jl:64-70)
        prev[halo[i, 2]] = halo[i, 1]; prev[halo[i, 1]] = halo[i, 2]
    end
    D = BellmanFordMoore(prev, dist)
    for i in 1:nphases
        level = partition.iterator[i]
        for (k, b) in enumerate(level[2:end])       # boundary_velocity! (:16-28)
            rq = rdir[i][k] == :above ? rb[b] - buffer_zone : rb[b] + buffer_zone
            U[bnodes[b]] .= interpolate_velocity([rq], interpolant)[1]
        end
        seeds = i == 1 ? Int64[source] : Int64[j for j in bnodes[level[2]] if isfinite(D.dist[j])]
        allowed = UInt8[partition.id[j] in level for j in 1:n]
        D, st = bfm_continue(G, halo, gr, U, allowed, seeds, D.dist, D.prev)
        println("Level $i converged in $(st.sweeps) iterations")
    end
    return D
end

end # module
