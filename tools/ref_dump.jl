#!/usr/bin/env julia
# Regenerate TRUE-reference goldens with the Julia reference (SpectralElements.jl) on a machine that
# has Julia: writes raw little-endian Float64 column-major files matching tests/golden/*.npz keys, so
# tests/tools/compare_ref_dump.py can check oracle/sem_oracle.py and the CUDA path against them.
#   julia --project=/path/to/SpectralElements.jl tools/ref_dump.jl outdir
using SpectralElements, LinearAlgebra
wavy(x, y) = (d = @. 0.1 * sin(pi * x) * sin(pi * y); (x .+ d, y .+ d))
function splitmix(n; seed = UInt64(0x5EED))
    out = zeros(n)
    for i in 1:n
        z = seed + UInt64(i) * 0x9E3779B97F4A7C15
        z = (z ⊻ (z >> 30)) * 0xBF58476D1CE4E5B9
        z = (z ⊻ (z >> 27)) * 0x94D049BB133111EB
        z = z ⊻ (z >> 31)
        out[i] = 2.0 * (Float64(z >> 11) / 9007199254740992.0) - 1.0
    end
    out
end
cases = Dict(
    "p2d_annulus_5x5_nr8" => (8, 5, 5, [false, true], SpectralElements.annulus, ['D', 'D', 'N', 'N'], 1.0, 0.0),
    "cfg1_wavy_8x8_nr9" => (9, 8, 8, [false, false], wavy, ['D', 'D', 'D', 'D'], 1.0, 0.0),
    "helmholtz_wavy_6x4_nr9" => (9, 6, 4, [false, false], wavy, ['D', 'D', 'D', 'D'], 0.7, 1.3),  # needs the Ey fix of mesh.jl:80
)
outdir = length(ARGS) > 0 ? ARGS[1] : "ref_dump"
mkpath(outdir)
for (name, (nr, Ex, Ey, per, deform, bc, nu, k)) in cases
    msh = Mesh(nr, nr, Ex, Ey, per, deform)
    M = Array{Float64}(generateMask(bc, msh))
    u = reshape(splitmix(length(msh.x)), size(msh.x))
    f = ones(size(msh.x))
    opl(v) = mask(gatherScatter(hlmz(v, nu, k, msh), msh), M)
    b = gatherScatter(mask(mass(f, msh), M), msh)
    dump(key, a) = write(joinpath(outdir, "$(name).$(key).f64"), Array{Float64}(a))
    dump("G11", msh.G11); dump("G12", msh.G12); dump("G22", msh.G22); dump("B", msh.B); dump("mult", msh.mult)
    dump("x", msh.x); dump("y", msh.y); dump("M", M); dump("u", u)
    dump("lapl", lapl(u, msh)); dump("hlmz", hlmz(u, nu, k, msh)); dump("gs", gatherScatter(u, msh))
    dump("oplhs", opl(u)); dump("rhs", b)
    dump("pcg_x", pcg(b, opl; mult = msh.mult, tol = 1e-8)); dump("pcg_x_tol12", pcg(b, opl; mult = msh.mult, tol = 1e-12))
end
