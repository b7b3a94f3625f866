# ref_timing.jl -- produces the TRUE reference figure for bench.py's workload wherever Julia exists:
#     julia --threads=auto bench/ref_timing.jl [KMS2party] [gates]
# Run from a checkout of SNUCP/MKTFHE (include path below).  Not executed in this repository (no Julia in the image);
# bench.py --impl reference times the C port of the same algorithm instead and says so.
include(joinpath(get(ENV, "MKTFHE_REF", "."), "src", "MKTFHE.jl"))
using .MKTFHE, Printf
pname = length(ARGS) ≥ 1 ? ARGS[1] : "KMS2party"
gates = length(ARGS) ≥ 2 ? parse(Int, ARGS[2]) : 16
params = getfield(MKTFHE, Symbol(pname))
a = CRS(params)
keys = [party_keygen(a, params) for _ = 1 : params.k]
scheme = setup(a, last.(keys), params)
lwekeys = first.(keys)
c1 = [lwe_ith_encrypt(rand(Bool), 1, lwekeys[1], params) for _ = 1 : gates]
c2 = [lwe_ith_encrypt(rand(Bool), params.k, lwekeys[params.k], params) for _ = 1 : gates]
NAND(c1[1], c2[1], scheme)                                  # compile
t = @elapsed for g = 1 : gates; NAND(c1[g], c2[g], scheme); end
@printf("%s: %d MK-NAND gates in %.2f s = %.2f gates/s on %d threads\n", pname, gates, t, gates / t, Threads.nthreads())
