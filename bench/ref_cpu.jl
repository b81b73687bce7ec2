#!/usr/bin/env julia
# True-reference CPU timing (for anyone with Julia): the same fused opLHS apply and PCG iterations
# bench.py times, through SpectralElements.jl's own functions.
#   julia --project=/path/to/SpectralElements.jl -t auto bench/ref_cpu.jl [nr] [E]
using SpectralElements, LinearAlgebra
nr = length(ARGS) > 0 ? parse(Int, ARGS[1]) : 9
E = length(ARGS) > 1 ? parse(Int, ARGS[2]) : 256
wavy(x, y) = (d = @. 0.1 * sin(pi * x) * sin(pi * y); (x .+ d, y .+ d))
msh = Mesh(nr, nr, E, E, [false, false], wavy)          # builds the dense QQt the reference uses (mesh.jl:81-82)
M = Array{Float64}(generateMask(['D', 'D', 'D', 'D'], msh))
u = rand(size(msh.x)...)
opl(v) = mask(gatherScatter(hlmz(v, 1.0, 0.0, msh), msh), M)
opl(u)
t = @elapsed for _ in 1:3 opl(u) end
println("opLHS: ", length(u) / (t / 3) / 1e9, " GDOF/s on ", Threads.nthreads(), " Julia threads, BLAS threads ", BLAS.get_num_threads())
b = gatherScatter(mask(mass(ones(size(u)), msh), M), msh)
t = @elapsed pcg(b, opl; mult = msh.mult, maxiter = 10)
println("pcg: ", 10 / t, " iterations/s")
