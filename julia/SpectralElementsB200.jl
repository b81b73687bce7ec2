#
# SpectralElementsB200.jl -- drop-in GPU methods for SpectralElements.jl's matrix-free hot path.
#
# `using SpectralElements, SpectralElementsB200` adds methods with the reference's own signatures
# (ABu, lapl, hlmz, mass, gatherScatter, mask, pcg, pcg!, opLHS, solve!) that ccall libsemb.so
# (include/semb.h).  Host code stays in Julia; the library owns every device buffer; there is no
# CUDA.jl kernel generation and no CPU fallback.
#
# NOTE: Julia is not installed in the build container, so this file is written against the C ABI and
# syntax-reviewed, but has NOT been executed.  tests/ drive the same ABI through Python ctypes.
#
# The methods below have the reference's exact signatures and REPLACE its CPU methods (that is the drop-in);
# Julia >= 1.10 refuses method overwriting while precompiling a package, so this module opts out.
__precompile__(false)

module SpectralElementsB200

using SpectralElements
using SpectralElements: Mesh, Diffusion, ConvectionDiffusion
import SpectralElements: ABu, lapl, hlmz, mass, gatherScatter, mask, pcg, pcg!, opLHS, solve!, grad, advect,
                         laplace, evolve!, step!, fixU!

const libsemb = get(ENV, "LIBSEMB", joinpath(@__DIR__, "..", "spectralelements.jl_b200", "lib", "libsemb.so"))

struct SembError <: Exception
    code::Cint
    msg::String
end

function check(rc::Cint)
    rc < 0 && throw(SembError(rc, unsafe_string(ccall((:semb_last_error, libsemb), Cstring, ()))))
    return rc
end

# ---- context (one per process / GPU) ------------------------------------------------------------
const CTX = Ref{Ptr{Cvoid}}(C_NULL)

function context(device::Integer = parse(Int, get(ENV, "SEMB_DEVICE", "0")))
    if CTX[] == C_NULL
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:semb_init, libsemb), Cint, (Cint, Ref{Ptr{Cvoid}}), device, h))
        CTX[] = h[]
        atexit(() -> ccall((:semb_finalize, libsemb), Cint, (Ptr{Cvoid},), CTX[]))
    end
    return CTX[]
end

"""Join the NCCL communicator: rank 0 creates the id, the caller broadcasts it (e.g. MPI.Bcast!)."""
function comm_unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:semb_comm_unique_id, libsemb), Cint, (Ptr{UInt8},), id))
    return id
end
comm_init(nranks, rank, id::Vector{UInt8}) =
    check(ccall((:semb_comm_init, libsemb), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), context(), nranks, rank, id))

# ---- device mesh cache: Mesh is immutable (mesh.jl:25), so objectid keys a semb_mesh* -------------
const MESHES = Dict{UInt,Ptr{Cvoid}}()

function devmesh(msh::Mesh)
    get!(MESHES, objectid(msh)) do
        h = Ref{Ptr{Cvoid}}(C_NULL)
        # the Julia Mesh already holds the operator arrays: hand them over as they are
        check(ccall((:semb_mesh_create_arrays, libsemb), Cint,
                    (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Cint, Cint,
                     Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                     Ref{Ptr{Cvoid}}),
                    context(), msh.nr, msh.ns, msh.Ex, msh.Ey, msh.ifperiodic[1], msh.ifperiodic[2],
                    msh.Dr, msh.Ds, msh.G11, msh.G12, msh.G22, msh.B, h))
        # metric terms for grad / advect / the Stokes split (enum semb_mesh_array: JAC = 2, JACI = 3, RX = 4, RY = 5,
        # SX = 6, SY = 7, BI = 9; approxHlmzInv and the Stokes operators need Bi)
        for (which, a) in ((2, msh.Jac), (3, msh.Jaci), (4, msh.rx), (5, msh.ry), (6, msh.sx), (7, msh.sy), (9, msh.Bi))
            check(ccall((:semb_mesh_set, libsemb), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), h[], which, a))
        end
        h[]
    end
end

f64(a::Array{Float64}) = a
f64(a::AbstractArray) = convert(Array{Float64}, a)   # BitMatrix masks (mesh.jl:171) are widened here
coef(c::Number) = (Ptr{Float64}(C_NULL), Float64(c), nothing)
coef(c::AbstractArray) = (a = f64(c); (pointer(a), 0.0, a))

# ---- operators: arrays in, fresh Array{Float64,2} out ------------------------------------------------
# lapl(u,msh), lapl.jl:26-36
function lapl(u::Array, msh::Mesh)
    out = similar(u, Float64)
    check(ccall((:semb_lapl_host, libsemb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), devmesh(msh), f64(u), out))
    return out
end

# hlmz(u,ν,k,msh), hlmz.jl:12-19 (ν, k scalar or array); lapl(u,ν,msh), lapl.jl:38-45
function hlmz(u::Array, ν, k, msh::Mesh)
    out = similar(u, Float64)
    (pν, sν, kν) = coef(ν); (pk, sk, kk) = coef(k)
    GC.@preserve kν kk check(ccall((:semb_hlmz_host, libsemb), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cdouble, Ptr{Float64}, Cdouble, Ptr{Float64}),
        devmesh(msh), f64(u), pν, sν, pk, sk, out))
    return out
end
lapl(u::Array, ν::Array, msh::Mesh) = hlmz(u, ν, 0.0, msh)
# the reference's "dealiased" mesh-pair forms ignore msh2 (lapl.jl:47-52, hlmz.jl:22-30, mass.jl:25-30)
lapl(u::Array, msh1::Mesh, msh2::Mesh) = lapl(u, msh1)
hlmz(u::Array, ν, k, msh1::Mesh, msh2::Mesh) = hlmz(u, ν, k, msh1)

# mass(u,msh), mass.jl:12-22
function mass(u::Array, msh::Mesh)
    out = similar(u, Float64)
    check(ccall((:semb_mass_host, libsemb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), devmesh(msh), f64(u), out))
    return out
end

mass(u::Array, msh1::Mesh, msh2::Mesh) = mass(u, msh1)

# gatherScatter(u,msh), gatherScatter.jl:18-21
function gatherScatter(u, msh::Mesh)
    out = similar(u, Float64)
    check(ccall((:semb_gather_scatter_host, libsemb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
                devmesh(msh), f64(u), out))
    return out
end

# ABu(As,Br,u), ABu.jl:9-37 (general rectangular blocks; [] = identity)
function ABu(As::AbstractArray, Br::AbstractArray, u::AbstractArray)
    (m, n) = size(u)
    (ma, na) = length(As) == 0 ? (0, 0) : size(As)
    (mb, nb) = length(Br) == 0 ? (0, 0) : size(Br)
    mo = length(Br) == 0 ? m : Int(m * mb / nb)      # InexactError as in ABu.jl:16
    no = length(As) == 0 ? n : Int(n / na * ma)      # ABu.jl:26
    out = zeros(Float64, mo, no)
    A = length(As) == 0 ? Float64[] : f64(As); B = length(Br) == 0 ? Float64[] : f64(Br)
    check(ccall((:semb_abu_host, libsemb), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Cint, Cint, Ptr{Float64}, Cint, Cint, Ptr{Float64}, Cint, Cint, Ptr{Float64}),
                context(), A, ma, na, B, mb, nb, f64(u), m, n, out))
    return out
end

# mask with the mesh at hand (padded device layout, no extra copy); the reference's 2-argument form follows below
function mask(u::Array, M::Array, msh::Mesh)
    out = similar(u, Float64)
    Mf = length(M) == 0 ? Float64[] : f64(M)
    GC.@preserve Mf check(ccall((:semb_mask_host, libsemb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                                devmesh(msh), f64(u), length(M) == 0 ? Ptr{Float64}(C_NULL) : pointer(Mf), out))
    return out
end

# mask(u,M), mask.jl:10-18 as the reference calls it (no mesh): M .* u on the device, copy(u) for M = []
mask(u::Array, M::Array) = length(M) == 0 ? copy(u) : mul(f64(M), f64(u))
# gatherScatter(u,QQtx,QQty), gatherScatter.jl:8-16: the dense form, through the device ABu
gatherScatter(u, QQtx::AbstractArray, QQty::AbstractArray) = ABu(QQty, QQtx, u)

# grad(u,msh), grad.jl:15-34 ; advect(T,ux,uy,mshV,mshD,Jr,Js), advect.jl:45-64 (Jr, Js are rebuilt by the library)
function devfield(msh::Mesh, a::Union{Array,Nothing} = nothing)
    f = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:semb_field_create, libsemb), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), devmesh(msh), f))
    a === nothing || check(ccall((:semb_field_upload, libsemb), Cint, (Ptr{Cvoid}, Ptr{Float64}), f[], f64(a)))
    return f[]
end
function devget(f::Ptr{Cvoid}, like::Array)
    out = similar(like, Float64)
    check(ccall((:semb_field_download, libsemb), Cint, (Ptr{Cvoid}, Ptr{Float64}), f, out))
    ccall((:semb_field_destroy, libsemb), Cint, (Ptr{Cvoid},), f)
    return out
end
function grad(u::Array, msh::Mesh)
    (fu, fx, fy) = (devfield(msh, u), devfield(msh), devfield(msh))
    check(ccall((:semb_grad, libsemb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), devmesh(msh), fu, fx, fy))
    ccall((:semb_field_destroy, libsemb), Cint, (Ptr{Cvoid},), fu)
    return devget(fx, u), devget(fy, u)
end
function advect(T::Array, ux::Array, uy::Array, mshV::Mesh, mshD::Mesh, Jr, Js)
    (fT, fx, fy, fo) = (devfield(mshV, T), devfield(mshV, ux), devfield(mshV, uy), devfield(mshV))
    check(ccall((:semb_advect, libsemb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                devmesh(mshV), devmesh(mshD), fT, fx, fy, fo))
    for f in (fT, fx, fy); ccall((:semb_field_destroy, libsemb), Cint, (Ptr{Cvoid},), f); end
    return devget(fo, T)
end

# ---- explicit-argument forms on plain arrays (examples/p2d_explicit.jl:183-188, examples/semPS.jl:168-172) ----------
nz(a) = length(a) == 0 ? nothing : f64(a)          # Julia's `[]` -> NULL
ptr(a) = a === nothing ? Ptr{Float64}(C_NULL) : pointer(a)
# laplace(u,Dr,Ds,G11,G12,G22), lapl.jl:70-81 ; laplace(u,Jr,Js,Dr,Ds,G11,G12,G22), lapl.jl:83-103 (dealiased)
function laplace(u::Array, Jr, Js, Dr, Ds, G11, G12, G22)
    out = similar(u, Float64)
    (jr, js) = (nz(Jr), nz(Js))
    (uf, dr, ds, g11, g12, g22) = (f64(u), f64(Dr), f64(Ds), f64(G11), f64(G12), f64(G22))
    GC.@preserve jr js uf dr ds g11 g12 g22 check(ccall((:semb_laplace_host, libsemb), Cint,
        (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Cint, Ptr{Float64}, Cint,
         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        context(), size(u, 1), size(u, 2), dr, size(Dr, 1), ds, size(Ds, 1), ptr(jr), jr === nothing ? 0 : size(Jr, 1),
        ptr(js), js === nothing ? 0 : size(Js, 1), g11, g12, g22, uf, out))
    return out
end
laplace(u::Array, Dr, Ds, G11, G12, G22) = laplace(u, [], [], Dr, Ds, G11, G12, G22)
mul(a::Array, b::Array) = (out = similar(b, Float64); check(ccall((:semb_mul_host, libsemb), Cint,
    (Ptr{Cvoid}, Csize_t, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), context(), length(b), f64(a), f64(b), out)); out)
# lapl(u,M,Jr,Js,QQtx,QQty,Dr,Ds,G11,G12,G22,mult), lapl.jl:54-68 (the mult hook only acts in the reverse pass)
function lapl(u::Array, M, Jr, Js, QQtx, QQty, Dr, Ds, G11, G12, G22, mult)
    Au = ABu(QQty, QQtx, laplace(u, Jr, Js, Dr, Ds, G11, G12, G22))     # gatherScatter.jl:8-16
    return length(M) == 0 ? Au : mul(f64(M), Au)                         # mask.jl:10-18
end
# mass(u,M,B,Jr,Js,QQtx,QQty,mult), mass.jl:32-50
function mass(u::Array, M, B, Jr, Js, QQtx, QQty, mult)
    out = similar(u, Float64)
    (jr, js, b, uf) = (nz(Jr), nz(Js), nz(B), f64(u))
    GC.@preserve jr js b uf check(ccall((:semb_mass_explicit_host, libsemb), Cint,
        (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Cint, Cint, Ptr{Float64}, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        context(), size(u, 1), size(u, 2), ptr(jr), jr === nothing ? 0 : size(Jr, 1), jr === nothing ? 0 : size(Jr, 2),
        ptr(js), js === nothing ? 0 : size(Js, 1), js === nothing ? 0 : size(Js, 2), ptr(b), uf, out))
    Bu = ABu(QQty, QQtx, out)
    return length(M) == 0 ? Bu : mul(f64(M), Bu)
end

# ---- Stokes split: diver.jl / stokes.jl are not executable as shipped (stokes.jl is not even included,
# SpectralElements.jl:53); these methods bind the reconstruction documented in include/semb.h ----------------------------
bcstr(bc) = String(bc)          # ['D','D','N','N'] -> "DDNN"
mutable struct StokesB200       # Stokes(bcVX,bcVY,mshV,mshD,mshP), stokes.jl:75-108, reduced to the pressure system
    h::Ptr{Cvoid}
    mshV::Mesh
    mshP::Mesh
end
function StokesB200(bcVX, bcVY, mshV::Mesh, mshP::Mesh; b0 = 1.0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:semb_stokes_create, libsemb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cstring, Cstring, Cdouble, Ref{Ptr{Cvoid}}),
                devmesh(mshV), devmesh(mshP), bcstr(bcVX), bcstr(bcVY), b0, h))
    return finalizer(s -> ccall((:semb_stokes_destroy, libsemb), Cint, (Ptr{Cvoid},), s.h), StokesB200(h[], mshV, mshP))
end
function gradᵀ(u::Array, msh::Mesh)          # grad.jl:44-63
    (fu, fx, fy) = (devfield(msh, u), devfield(msh), devfield(msh))
    check(ccall((:semb_gradT, libsemb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), devmesh(msh), fu, fx, fy))
    ccall((:semb_field_destroy, libsemb), Cint, (Ptr{Cvoid},), fu)
    return devget(fx, u), devget(fy, u)
end
function diver(ux::Array, uy::Array, sks::StokesB200)      # diver.jl:17-31
    (fx, fy, fo) = (devfield(sks.mshV, ux), devfield(sks.mshV, uy), devfield(sks.mshP))
    check(ccall((:semb_diver, libsemb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), sks.h, fx, fy, fo))
    for f in (fx, fy); ccall((:semb_field_destroy, libsemb), Cint, (Ptr{Cvoid},), f); end
    return devget(fo, sks.mshP.x)
end
function diverᵀ(pr::Array, sks::StokesB200)                # diver.jl:53-63
    (fp, fx, fy) = (devfield(sks.mshP, pr), devfield(sks.mshV), devfield(sks.mshV))
    check(ccall((:semb_diverT, libsemb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), sks.h, fp, fx, fy))
    ccall((:semb_field_destroy, libsemb), Cint, (Ptr{Cvoid},), fp)
    return devget(fx, sks.mshV.x), devget(fy, sks.mshV.x)
end
function approxHlmzInv(u::Array, b0::Number, mshV::Mesh, bc)   # diver.jl:92-104
    (fu, fo) = (devfield(mshV, u), devfield(mshV))
    check(ccall((:semb_approx_hlmz_inv, libsemb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cstring, Ptr{Cvoid}),
                devmesh(mshV), fu, b0, bcstr(bc), fo))
    ccall((:semb_field_destroy, libsemb), Cint, (Ptr{Cvoid},), fu)
    return devget(fo, u)
end
function opStokesLHS(q::Array, sks::StokesB200)            # stokes.jl:110-121
    (fq, fo) = (devfield(sks.mshP, q), devfield(sks.mshP))
    check(ccall((:semb_stokes_op, libsemb), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), sks.h, fq, fo))
    ccall((:semb_field_destroy, libsemb), Cint, (Ptr{Cvoid},), fq)
    return devget(fo, q)
end
# pressureProject!, stokes.jl:159-177: vx, vy, pr are updated in place; returns the PCG iteration count
function pressureProject!(vx::Array, vy::Array, pr::Array, sks::StokesB200; tol = 1e-8, maxiter = -1)
    (fx, fy, fp) = (devfield(sks.mshV, vx), devfield(sks.mshV, vy), devfield(sks.mshP, pr))
    (it, res) = (Ref{Clonglong}(0), Ref{Cdouble}(0.0))
    check(ccall((:semb_stokes_project, libsemb), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Clonglong, Ref{Clonglong}, Ref{Cdouble}),
                sks.h, fx, fy, fp, tol, maxiter, it, res))
    vx .= devget(fx, vx); vy .= devget(fy, vy); pr .= devget(fp, pr)
    return it[]
end

# ---- device-resident time-step drivers (SURVEY 8f-1/8f-2): u, uh[1..k], ub, ν, f, rhs (+ vx, vy) stay in HBM ----------
# enum semb_diffusion_field_id (include/semb.h)
const DFN_U, DFN_UB, DFN_NU, DFN_F, DFN_RHS, DFN_VX, DFN_VY, DFN_UH0 = 0, 1, 2, 3, 4, 5, 6, 8
const DRIVERS = Dict{UInt,Ptr{Cvoid}}()

# Field keeps only the mask array (mesh.jl:179-185): the 'D'/'N' flags are read back from its four outer lines
bcflags(M) = (m = f64(M); String([m[1, 2] == 0 ? 'D' : 'N', m[end, 2] == 0 ? 'D' : 'N',
                                  m[2, 1] == 0 ? 'D' : 'N', m[2, end] == 0 ? 'D' : 'N']))
function dfnfield(d::Ptr{Cvoid}, which::Integer)
    f = Ref{Ptr{Cvoid}}(C_NULL)      # borrowed handle: owned by the driver, never destroyed here
    check(ccall((:semb_diffusion_field, libsemb), Cint, (Ptr{Cvoid}, Cint, Ref{Ptr{Cvoid}}), d, which, f))
    return f[]
end
dfnput(d::Ptr{Cvoid}, which::Integer, a::Array) =
    check(ccall((:semb_field_upload, libsemb), Cint, (Ptr{Cvoid}, Ptr{Float64}), dfnfield(d, which), f64(a)))
dfnget!(a::Array, d::Ptr{Cvoid}, which::Integer) =
    check(ccall((:semb_field_download, libsemb), Cint, (Ptr{Cvoid}, Ptr{Float64}), dfnfield(d, which), a))

"""The device twin of a `Diffusion` / `ConvectionDiffusion`, created at the first step (so at `istep == 0`) from
the host struct's current state and cached by `objectid`."""
function driver(eq, mshV::Mesh, mshD::Union{Mesh,Nothing} = nothing)
    get!(DRIVERS, objectid(eq)) do
        (fld, ts) = (eq.fld, eq.tstep)
        k = length(fld.uh)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        if mshD === nothing          # Diffusion(bc,msh;Ti,Tf,dt,k), diffusion.jl:20-34
            check(ccall((:semb_diffusion_create, libsemb), Cint,
                        (Ptr{Cvoid}, Cstring, Cdouble, Cdouble, Cdouble, Cint, Ref{Ptr{Cvoid}}),
                        devmesh(mshV), bcflags(fld.M), ts.time[1], ts.Tf[1], ts.dt[1], k, h))
        else                         # ConvectionDiffusion(...), convectionDiffusion.jl:31-56
            check(ccall((:semb_convdiff_create, libsemb), Cint,
                        (Ptr{Cvoid}, Ptr{Cvoid}, Cstring, Cdouble, Cdouble, Cdouble, Cint, Ref{Ptr{Cvoid}}),
                        devmesh(mshV), devmesh(mshD), bcflags(fld.M), ts.time[1], ts.Tf[1], ts.dt[1], k, h))
            dfnput(h[], DFN_VX, eq.vx); dfnput(h[], DFN_VY, eq.vy)
        end
        dfnput(h[], DFN_U, fld.u); dfnput(h[], DFN_UB, fld.ub); dfnput(h[], DFN_NU, eq.ν); dfnput(h[], DFN_F, eq.f)
        for i in 1:k
            dfnput(h[], DFN_UH0 + i - 1, fld.uh[i])
        end
        h[]
    end
end

"""set_precond!(eq, kind): the opM of the step's solve (pcg.jl:37) -- `:reference` is what the reference passes (identity in
diffusion.jl:71, opPrecond in convectionDiffusion.jl:118), `:fdm` the FDM preconditioner of ν*lapl + b0*mass (lapl.jl:105-119;
constant ν only): same solution to the solver tolerance in several times fewer iterations, but not the reference's
iteration counts.  Opt-in."""
set_precond!(dfn::Diffusion, kind::Symbol) =
    check(ccall((:semb_diffusion_set_precond, libsemb), Cint, (Ptr{Cvoid}, Cint), driver(dfn, dfn.msh), kind == :fdm ? 2 : 0))
set_precond!(cdn::ConvectionDiffusion, kind::Symbol) =
    check(ccall((:semb_diffusion_set_precond, libsemb), Cint, (Ptr{Cvoid}, Cint), driver(cdn, cdn.mshV, cdn.mshD), kind == :fdm ? 2 : 0))

# one step of either equation: updateHist! + time/BDF update on the device, the user closures on the host (only
# what a closure other than the no-op fixU! may have changed is uploaded), makeRHS! + solve! on the device
function devstep!(eq, d::Ptr{Cvoid}, msh::Mesh, setBC!, setForcing!, setVisc!; tol = 1e-8, sync_velocity = true)
    (fld, ts) = (eq.fld, eq.tstep)
    (t, n) = (Ref{Cdouble}(0.0), Ref{Clonglong}(0))
    check(ccall((:semb_diffusion_begin_step, libsemb), Cint, (Ptr{Cvoid}, Ref{Cdouble}, Ref{Clonglong}), d, t, n))
    # the host TimeStepper follows (time.jl:70-82), so callbacks and the loop test of simulate! read current values
    check(ccall((:semb_diffusion_state, libsemb), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Clonglong}), d, ts.time, ts.bdfA, ts.bdfB, n))
    ts.istep[1] = n[]
    for (set!, a, which) in ((setBC!, fld.ub, DFN_UB), (setForcing!, eq.f, DFN_F), (setVisc!, eq.ν, DFN_NU))
        set! === fixU! && continue
        set!(a, msh.x, msh.y, ts.time[1])
        dfnput(d, which, a)
    end
    # makeRHS! reads cdn.vx, cdn.vy every step (convectionDiffusion.jl:100-105): a user may have changed them between
    # steps, so the device copies follow the host arrays (2 uploads per step; `sync_velocity = false` skips them when
    # the advecting field is known to be constant)
    if hasproperty(eq, :vx) && sync_velocity
        dfnput(d, DFN_VX, eq.vx); dfnput(d, DFN_VY, eq.vy)
    end
    (it, res) = (Ref{Clonglong}(0), Ref{Cdouble}(0.0))
    rc = check(ccall((:semb_diffusion_finish_step, libsemb), Cint, (Ptr{Cvoid}, Cdouble, Ref{Clonglong}, Ref{Cdouble}),
                     d, tol, it, res))
    rc == 1 && println("warning: res:", res[])           # pcg.jl:39
    SpectralElements.updateHist!(fld)                    # host copy of the history, for callbacks (mesh.jl:199-215)
    dfnget!(fld.u, d, DFN_U)
    return it[]
end

# evolve!(dfn,setBC!,setForcing!,setVisc!), diffusion.jl:81-106
function evolve!(dfn::Diffusion, setBC! = fixU!, setForcing! = fixU!, setVisc! = fixU!)
    devstep!(dfn, driver(dfn, dfn.msh), dfn.msh, setBC!, setForcing!, setVisc!)
    return
end
# step!(cdn), convectionDiffusion.jl:150-157 (updateHist!, updateHist!(tstep), evolve! with the struct's closures)
function step!(cdn::ConvectionDiffusion)
    devstep!(cdn, driver(cdn, cdn.mshV, cdn.mshD), cdn.mshV, cdn.set∂!, cdn.setF!, cdn.setν!)
    return
end
"""Release the device twin of an equation (the library frees its fields)."""
function release!(eq)
    d = pop!(DRIVERS, objectid(eq), C_NULL)
    d == C_NULL || ccall((:semb_diffusion_destroy, libsemb), Cint, (Ptr{Cvoid},), d)
    return
end

# ---- the fused unit and the device-resident Krylov loop ------------------------------------------------
"""opLHS as a callable struct: applying it runs the fused kernel; handing it to pcg runs the whole
loop on the device (an arbitrary Julia closure cannot execute there)."""
struct OpLHS
    msh::Mesh
    ν
    k
    M::Array{Float64}
end
function (op::OpLHS)(u::Array)
    out = similar(u, Float64)
    (pν, sν, kν) = coef(op.ν); (pk, sk, kk) = coef(op.k)
    GC.@preserve kν kk check(ccall((:semb_oplhs_host, libsemb), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cdouble, Ptr{Float64}, Cdouble, Cstring, Ptr{Float64}, Ptr{Float64}),
        devmesh(op.msh), f64(u), pν, sν, pk, sk, C_NULL, op.M, out))
    return out
end
Base.:*(op::OpLHS, u::AbstractArray) = op(u)

# opLHS(u,dfn), diffusion.jl:36-45 / convectionDiffusion.jl:76-85
opLHS(u::Array, dfn::Diffusion) = OpLHS(dfn.msh, dfn.ν, dfn.tstep.bdfB[1], dfn.fld.M)(u)
opLHS(u::Array, cdn::ConvectionDiffusion) = OpLHS(cdn.mshV, cdn.ν, cdn.tstep.bdfB[1], cdn.fld.M)(u)

# struct semb_pcg_opts (include/semb.h)
struct PcgOpts
    nu::Cdouble; nu_arr::Ptr{Cvoid}
    k::Cdouble; k_arr::Ptr{Cvoid}
    bc::Ptr{UInt8}; M_arr::Ptr{Cvoid}
    precond::Cint; prec_b0::Cdouble
    tol::Cdouble; maxiter::Clonglong
    check_every::Cint
end

"""DiagPrecond(msh,b0): opPrecond(u) = u ./ B ./ b0, convectionDiffusion.jl:87-91"""
struct DiagPrecond
    msh::Mesh
    b0::Float64
end

"""FdmPrecond(msh,bc,ν,k): the fast-diagonalisation preconditioner as opM of pcg -- the reference's commented-out
lapl_fdm(b,Bi,Sx,Sy,Sxi,Syi,Di) (lapl.jl:105-119, set-up examples/p2d_explicit.jl:109-141), applied on every element
extended by one node into its neighbours and combined symmetrically (include/semb.h).  ν, k: the constant coefficients
of the operator it approximates; bc: its ['D','N',...] flags.  One per mesh; collective on several ranks."""
struct FdmPrecond
    msh::Mesh
    function FdmPrecond(msh::Mesh, bc, ν::Real = 1.0, k::Real = 0.0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:semb_fdm_create, libsemb), Cint, (Ptr{Cvoid}, Cstring, Cdouble, Cdouble, Ref{Ptr{Cvoid}}),
                    devmesh(msh), bcstr(bc), ν, k, h))
        return new(msh)
    end
end
function (P::FdmPrecond)(r::Array)      # h = opM(r) for a continuous r
    out = similar(r, Float64)
    check(ccall((:semb_fdm_apply_host, libsemb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), devmesh(P.msh), f64(r), out))
    return out
end

# pcg(b,opA;opM,mult,ifv,tol,maxiter), pcg.jl:16-60 -- device-resident.
# mult: the device loop weights its inner products with the mesh's own msh.mult, what every caller in the reference passes
# (diffusion.jl:71, examples/p2d.jl:60).  DEVIATION from the bare default: pcg.jl:18 defaults mult to ones(size(b)); here
# `nothing` means msh.mult, and any other array is an error instead of being silently dropped.
function pcg(b, opA::OpLHS; opM = nothing, mult = nothing, ifv = false, tol = 1e-8, maxiter = length(b))
    (mult === nothing || mult == opA.msh.mult) ||
        throw(ArgumentError("pcg: mult must be msh.mult (or nothing = msh.mult); the device loop has no other weighting"))
    x = zeros(Float64, size(b))
    (pν, sν, kν) = coef(opA.ν); (pk, sk, kk) = coef(opA.k)
    prec = opM isa DiagPrecond ? 1 : (opM isa FdmPrecond ? 2 : 0)
    o = Ref(PcgOpts(sν, C_NULL, sk, C_NULL, C_NULL, C_NULL, prec, prec == 1 ? opM.b0 : 1.0, tol, maxiter, 0))
    it = Ref{Clonglong}(0); res = Ref{Cdouble}(0.0)
    rc = GC.@preserve kν kk check(ccall((:semb_pcg_host, libsemb), Cint,
        (Ptr{Cvoid}, Ref{PcgOpts}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ref{Clonglong}, Ref{Cdouble}),
        devmesh(opA.msh), o, pν, pk, opA.M, f64(b), x, it, res))
    rc == 1 && println("warning: res:", res[])           # pcg.jl:39
    ifv && println("PCG iter: ", it[], ", res: ", res[])   # pcg.jl:57
    return x
end
function pcg!(x, b, opA::OpLHS; kw...)
    x .= pcg(b, opA; kw...)
    return
end

# solve!(dfn), diffusion.jl:67-77 ; solve!(cdn), convectionDiffusion.jl:112-122
function solve!(dfn::Diffusion)
    fld = dfn.fld
    pcg!(fld.u, dfn.rhs, OpLHS(dfn.msh, dfn.ν, dfn.tstep.bdfB[1], fld.M); mult = dfn.msh.mult)
    fld.u .+= fld.ub
    return
end
function solve!(cdn::ConvectionDiffusion)
    fld = cdn.fld
    b0 = cdn.tstep.bdfB[1]
    pcg!(fld.u, cdn.rhs, OpLHS(cdn.mshV, cdn.ν, b0, fld.M); opM = DiagPrecond(cdn.mshV, b0), mult = cdn.mshV.mult)
    fld.u .+= fld.ub
    return
end

export OpLHS, DiagPrecond, FdmPrecond, set_precond!, StokesB200, release!, comm_unique_id, comm_init

end # module
