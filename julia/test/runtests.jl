#
# Parity of the B200 shim against the reference's own CPU methods, in the style of the reference's
# test/runtests.jl + test/examples.jl.  Run where Julia, SpectralElements.jl and a B200 are available:
#
#   LIBSEMB=/path/to/libsemb.so julia --project=<SpectralElements.jl checkout> julia/test/runtests.jl
#
# The shim REPLACES the reference's methods (same signatures), so every CPU result is computed BEFORE the shim is
# loaded and the same expressions are evaluated again afterwards.  Tolerances are the contract of the path:
# 1e-12 normwise per operator apply, bit-exact gatherScatter / mask, 1e-10 for converged PCG solutions.
#
# NOTE: written against the C ABI in an image without Julia; it has not been executed (see INTEGRATION.md).
#
using Test, LinearAlgebra
using SpectralElements

relerr(a, b) = norm(a .- b, Inf) / max(norm(b, Inf), floatmin(Float64))

# ---- inputs: the reference's own constructors (mesh.jl:66-133, examples/p2d.jl:41-43) ------------------------------
wavy(x, y) = (d = @. 0.1 * sin(pi * x) * sin(pi * y); (x .+ d, y .+ d))
cases = [
    ("box 8x8 order 8",      Mesh(9, 9, 8, 8, [false, false], wavy),                 ['D', 'D', 'D', 'D']),
    ("annulus periodic-y",   Mesh(8, 8, 5, 5, [false, true], SpectralElements.annulus), ['D', 'D', 'N', 'N']),
    ("order 12",             Mesh(13, 13, 3, 3, [false, false], wavy),               ['D', 'N', 'D', 'N']),
]

function evaluate(msh, bc)
    M  = generateMask(bc, msh)
    u  = @. sin(pi * msh.x) * cos(2pi * msh.y) + 0.3 * msh.x * msh.y     # SURVEY 8d closed-form input
    nu = @. 1.0 + 0.5 * msh.x^2
    Mf = convert(Array{Float64}, M)
    opA(v) = mask(gatherScatter(hlmz(v, 1.0, 1.0, msh), msh), Mf)        # diffusion.jl:36-45
    b  = gatherScatter(mask(mass(ones(size(u)), msh), Mf), msh)          # diffusion.jl:55,62-63
    return Dict(
        "lapl"          => lapl(u, msh),
        "lapl_nu"       => lapl(u, nu, msh),
        "hlmz"          => hlmz(u, nu, 2.5, msh),
        "mass"          => mass(u, msh),
        "gatherScatter" => gatherScatter(u, msh),
        "mask"          => mask(u, Mf),
        "ABu"           => ABu(msh.Ds, msh.Dr, u),
        "opLHS"         => opA(u),
        "pcg"           => pcg(b, opA; mult = msh.mult, tol = 1e-12),
    )
end

cpu = [evaluate(msh, bc) for (_, msh, bc) in cases]

# ---- the same expressions through libsemb ------------------------------------------------------------------------------
include(joinpath(@__DIR__, "..", "SpectralElementsB200.jl"))
using .SpectralElementsB200

function evaluate_gpu(msh, bc)
    M  = generateMask(bc, msh)
    u  = @. sin(pi * msh.x) * cos(2pi * msh.y) + 0.3 * msh.x * msh.y
    nu = @. 1.0 + 0.5 * msh.x^2
    Mf = convert(Array{Float64}, M)
    op = OpLHS(msh, 1.0, 1.0, Mf)                                        # the fused unit, device-resident in pcg
    b  = gatherScatter(mask(mass(ones(size(u)), msh), Mf), msh)
    return Dict(
        "lapl"          => lapl(u, msh),
        "lapl_nu"       => lapl(u, nu, msh),
        "hlmz"          => hlmz(u, nu, 2.5, msh),
        "mass"          => mass(u, msh),
        "gatherScatter" => gatherScatter(u, msh),
        "mask"          => mask(u, Mf),
        "ABu"           => ABu(msh.Ds, msh.Dr, u),
        "opLHS"         => op(u),
        "pcg"           => pcg(b, op; mult = msh.mult, tol = 1e-12),
    )
end

@testset "SpectralElementsB200 vs SpectralElements (CPU)" begin
    for (i, (name, msh, bc)) in enumerate(cases)
        gpu = evaluate_gpu(msh, bc)
        @testset "$name" begin
            for key in ("lapl", "lapl_nu", "hlmz", "mass", "ABu", "opLHS")
                @test relerr(gpu[key], cpu[i][key]) < 1e-12
            end
            @test gpu["gatherScatter"] == cpu[i]["gatherScatter"]          # two-term sums: bit-exact
            @test gpu["mask"] == cpu[i]["mask"]
            @test relerr(gpu["pcg"], cpu[i]["pcg"]) < 1e-10
        end
    end
end

# ---- FDM preconditioner (the reference has it only as commented-out sketches, lapl.jl:105-119: no CPU twin to compare
# with -- what can be held is that it is a preconditioner: same solution, several times fewer iterations) ---------------
@testset "FdmPrecond as opM of pcg" begin
    for (name, msh, bc) in cases
        M  = convert(Array{Float64}, generateMask(bc, msh))
        op = OpLHS(msh, 1.0, 1.0, M)
        b  = gatherScatter(mask(mass(ones(size(msh.x)), msh), M), msh)
        P  = FdmPrecond(msh, bc, 1.0, 1.0)
        x0 = pcg(b, op; mult = msh.mult, tol = 1e-12)
        x1 = pcg(b, op; opM = P, mult = msh.mult, tol = 1e-12)
        @test relerr(x1, x0) < 1e-9
        r  = mask(gatherScatter(b .* msh.mult, msh), M)
        h  = P(r)
        @test sum(r .* h .* msh.mult) > 0                                  # positive definite in pcg's inner product
    end
end
#
