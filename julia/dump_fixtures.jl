# dump_fixtures.jl -- turn "parity unpinned" into reference-pinned parity on any machine that has Julia.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI (the build image has no Julia).  Run it next to a checkout of SNUCP/MKTFHE:
#
#     cd <MKTFHE checkout>
#     julia --threads=auto <this repo>/julia/dump_fixtures.jl <this repo>/tests/golden_ref [CGGIparam KMS2party ...]
#
# For each of the five parameter sets the reference's own tests run (test/CGGI.jl:5, LMSS.jl:5, CCS.jl:5, KMS.jl:5,
# KMSblock.jl:5) it follows the reference's acceptance flow (test/KMS.jl:5-37): CRS -> party_keygen -> setup -> encrypt -> gates,
# and writes, in the blob container of mktfhe_b200/blob.py (magic "MKTFHEB2", JSON header, 4096-byte aligned arrays):
#
#     <set>.keys.blob      the REAL reference keys in the flat upload layouts of include/mktfhe_b200.h (brk, ksk, rlk, pubb per
#                          party, crs_fft) plus the LWE secret keys (p{i}.lwekey) so that decryptions can be checked
#     <set>.fixture.blob   in1, in2              [B][1 + n*k] uint32   input pairs: fresh single-party encryptions and
#                                                                      bootstrapped (full-support) ciphertexts
#                          out_NAND .. out_NOR   [B][1 + n*k] uint32   the reference's NAND(c1, c2, scheme) ... NOR(...)
#                          nand_linear           [B][1 + n*k] uint32   gate.jl:2-4 before bootstrapping!
#                          nand_tilde            [B][1 + n*k] uint32   bootstrapping.jl:8-9 (b~ first)
#                          nand_acc              [B][k+1][N]  torus    accumulator after blindrotate! (bootstrapping.jl:25)
#                          bits1, bits2          [B] uint8             plaintext bits
#
# tests/test_reference_fixtures.py consumes the directory: the CPU oracle (oracle/) and the GPU STRICT path must reproduce
# every array bit for bit from the same keys.  The reference is non-reproducible by construction (unseeded ChaCha20 streams),
# which is why keys and inputs are dumped rather than seeds.
include(joinpath(pwd(), "src", "MKTFHE.jl"))
include(joinpath(@__DIR__, "MKTFHEB200.jl"))
using .MKTFHE, SHA, Printf
import .MKTFHEB200: cparams, flatten_rgsw, flatten_unienc, flatten_polys, flatten_ksk

const MAGIC = b"MKTFHEB2"
const ALIGN = 4096
const SCHEME_NAMES = Dict(0 => "CGGI", 1 => "LMSS", 2 => "CCS", 3 => "KMS", 4 => "KMS_block")

dtype_str(::Type{Float64}) = "<f8"
dtype_str(::Type{UInt32}) = "<u4"
dtype_str(::Type{UInt64}) = "<u8"
dtype_str(::Type{UInt8}) = "|u1"

# (name, flat vector in C order, numpy shape)
const Arr = Tuple{String, Vector, Vector{Int}}

jstr(s::AbstractString) = "\"" * s * "\""
jnum(x::Integer) = string(x)
jnum(x::AbstractFloat) = isinteger(x) ? @sprintf("%.1f", x) : repr(Float64(x))     # 131072.0, 85.4084: parse back exactly

"params dict exactly as Python's dataclasses.asdict(mktfhe_b200.params.Params) spells it"
function params_json(name::String, p)
    c = cparams(p)
    fields = ["name" => jstr(name), "scheme" => jnum(c.scheme), "n" => jnum(c.n), "N" => jnum(c.N), "k" => jnum(c.k),
              "alpha" => jnum(c.alpha), "beta" => jnum(c.beta), "f" => jnum(c.f), "logD" => jnum(c.logD), "d" => jnum(c.d),
              "ell" => jnum(c.ell), "l_gsw" => jnum(c.l_gsw), "logB_gsw" => jnum(c.logB_gsw), "l_lev" => jnum(c.l_lev),
              "logB_lev" => jnum(c.logB_lev), "l_uni" => jnum(c.l_uni), "logB_uni" => jnum(c.logB_uni)]
    "{" * join([jstr(k) * ": " * v for (k, v) in fields], ", ") * "}"
end

function write_blob(path::String, kind::String, pjson::String, arrays::Vector{Arr})
    dir, offs = String[], Int[]
    off = 0
    for (name, data, shape) in arrays
        @assert prod(shape) == length(data) "$name: shape $(shape) does not match $(length(data)) elements"
        nbytes = sizeof(data)
        digest = bytes2hex(sha256(collect(reinterpret(UInt8, data))))
        push!(dir, "{\"name\": $(jstr(name)), \"dtype\": $(jstr(dtype_str(eltype(data)))), \"shape\": [$(join(shape, ", "))], " *
                   "\"offset\": $off, \"nbytes\": $nbytes, \"sha256\": $(jstr(digest))}")
        push!(offs, off)
        off += cld(nbytes, ALIGN) * ALIGN
    end
    header = "{\"kind\": $(jstr(kind)), \"params\": $pjson, \"seed\": null, \"source\": \"SNUCP/MKTFHE reference run (julia/dump_fixtures.jl)\", " *
             "\"arrays\": [" * join(dir, ", ") * "]}"
    hbytes = Vector{UInt8}(header)
    data0 = cld(16 + length(hbytes), ALIGN) * ALIGN
    open(path, "w") do io
        write(io, MAGIC); write(io, UInt32(1)); write(io, UInt32(length(hbytes))); write(io, hbytes)
        pos = 16 + length(hbytes)
        for (i, (name, data, shape)) in enumerate(arrays)
            target = data0 + offs[i]
            write(io, zeros(UInt8, target - pos)); pos = target
            write(io, data); pos += sizeof(data)
        end
        write(io, zeros(UInt8, data0 + off - pos))
    end
    path
end

lwe_words(c::MKTFHE.LWE) = vcat(c.b, c.a)
pack(cs) = reduce(vcat, lwe_words.(cs))

"bootstrapping.jl:4-27 with the intermediates kept: returns (tilde [b~; a~], accumulator words after blindrotate!, refreshed ciphertext)"
function bootstrap_traced(ctxt::MKTFHE.LWE{T}, scheme::MKTFHE.TFHEscheme{R, S}) where {T, R, S}
    N, logN = scheme.N, trailing_zeros(scheme.N)
    tildea = MKTFHE.divbits.(ctxt.a, MKTFHE.bits(T) - logN - 1)
    tildeb = MKTFHE.divbits(ctxt.b, MKTFHE.bits(T) - logN - 1)
    tilde = vcat(UInt32(tildeb), UInt32.(tildea))
    oneovereight = R(1) << (MKTFHE.bits(R) - 3)
    b = MKTFHE.zeronativepoly(N, R)
    if tildeb ≤ N
        for i = 1 : N
            b.coeffs[i] = i ≤ tildeb ? oneovereight : -oneovereight
        end
    else
        tb = tildeb - R(N)
        for i = 1 : N
            b.coeffs[i] = i ≤ tb ? -oneovereight : oneovereight
        end
    end
    acc = MKTFHE.RLWE(b, [MKTFHE.zeronativepoly(N, R) for _ = 1 : scheme.k])
    MKTFHE.blindrotate!(tildea, acc, scheme)
    accwords = vcat(copy(acc.b.coeffs), [copy(a.coeffs) for a in acc.a]...)
    res = MKTFHE.LWE(ctxt.b, copy(ctxt.a))
    MKTFHE.keyswitch!(res, acc, scheme)
    tilde, accwords, res
end

"gate.jl:2-4"
nand_linear(c1::MKTFHE.LWE{T}, c2::MKTFHE.LWE{T}) where T =
    MKTFHE.LWE(T(1) << (MKTFHE.bits(T) - 3) - c1.b - c2.b, (@. -c1.a - c2.a))

function dump_set(outdir::String, name::String, params)
    @printf("%s: key generation ...\n", name)
    mk = params isa MKTFHE.MKTFHEparams
    cp = cparams(params)
    n, k, N, f = Int(cp.n), Int(cp.k), Int(cp.N), Int(cp.f)
    arrays = Arr[]
    local scheme, lwekeys
    if mk
        a = CRS(params)
        keys = [party_keygen(a, params) for _ = 1 : params.k]
        lwekeys = first.(keys)
        btks = last.(keys)
        scheme = setup(a, btks, params)
        push!(arrays, ("crs_fft", flatten_polys(scheme.a), [Int(cp.l_uni), N ÷ 2, 2]))
    else
        lwekey, _ringkey, scheme = setup(params)
        lwekeys = [lwekey]
        btks = [scheme.btk]
    end
    brk_polys = cp.scheme == 2 ? 3 * Int(cp.l_uni) : 4 * Int(cp.l_gsw)
    Dk = size(btks[1].ksk, 1)
    for (i, btk) in enumerate(btks)
        brk = cp.scheme == 2 ? reduce(vcat, flatten_unienc.(btk.brk)) : flatten_rgsw(btk.brk)
        push!(arrays, ("p$(i-1).brk", brk, [n, brk_polys, N ÷ 2, 2]))
        push!(arrays, ("p$(i-1).ksk", flatten_ksk(btk.ksk, n, f), [N, Dk, f, n + 1]))
        hasproperty(btk, :rlk) && push!(arrays, ("p$(i-1).rlk", flatten_unienc(btk.rlk), [Int(cp.l_uni), 3, N ÷ 2, 2]))
        hasproperty(btk, :b) && push!(arrays, ("p$(i-1).pubb", flatten_polys(btk.b), [Int(cp.l_uni), N ÷ 2, 2]))
        push!(arrays, ("p$(i-1).lwekey", UInt32.(lwekeys[i].key), [n]))
    end
    pj = params_json(name, params)
    write_blob(joinpath(outdir, "$name.keys.blob"), "keys", pj, arrays)

    # ---- inputs: fresh encryptions (party i and party i+1 for the multi-key sets) and bootstrapped, full-support ciphertexts
    B0 = 4
    bits1, bits2 = rand(Bool, B0), rand(Bool, B0)
    enc(m, g) = mk ? lwe_ith_encrypt(m, 1 + (g % k), lwekeys[1 + (g % k)], params) : lwe_encrypt(m, lwekeys[1], params)
    in1 = [enc(bits1[g], g) for g = 1 : B0]
    in2 = [enc(bits2[g], g + 1) for g = 1 : B0]
    # two more pairs whose operands are gate outputs (all k blocks populated)
    push!(in1, NAND(in1[1], in2[1], scheme)); push!(bits1, !(bits1[1] & bits2[1]))
    push!(in2, OR(in1[2], in2[2], scheme));   push!(bits2, bits1[2] | bits2[2])
    push!(in1, XOR(in1[3], in2[3], scheme));  push!(bits1, bits1[3] ⊻ bits2[3])
    push!(in2, NOR(in1[4], in2[4], scheme));  push!(bits2, !(bits1[4] | bits2[4]))
    B = length(in1)
    words = 1 + n * k
    fx = Arr[("in1", pack(in1), [B, words]), ("in2", pack(in2), [B, words]),
             ("bits1", UInt8.(bits1), [B]), ("bits2", UInt8.(bits2), [B])]
    truth = Dict("NAND" => (x, y) -> !(x & y), "AND" => (x, y) -> x & y, "OR" => (x, y) -> x | y, "XOR" => (x, y) -> x ⊻ y,
                 "XNOR" => (x, y) -> !(x ⊻ y), "NOR" => (x, y) -> !(x | y))
    for (gname, g) in (("NAND", NAND), ("AND", AND), ("OR", OR), ("XOR", XOR), ("XNOR", XNOR), ("NOR", NOR))
        outs = [g(in1[i], in2[i], scheme) for i = 1 : B]
        for i = 1 : B                                    # the reference's own acceptance criterion (test/KMS.jl:37)
            dec = mk ? lwe_decrypt(outs[i], lwekeys, params) : lwe_decrypt(outs[i], lwekeys[1])
            @assert dec == truth[gname](bits1[i], bits2[i]) "$name $gname pair $i decrypts wrongly in the reference itself"
        end
        push!(fx, ("out_$gname", pack(outs), [B, words]))
    end
    lins = [nand_linear(in1[i], in2[i]) for i = 1 : B]
    traced = [bootstrap_traced(l, scheme) for l in lins]
    push!(fx, ("nand_linear", pack(lins), [B, words]))
    push!(fx, ("nand_tilde", reduce(vcat, [t[1] for t in traced]), [B, words]))
    push!(fx, ("nand_acc", reduce(vcat, [t[2] for t in traced]), [B, k + 1, N]))
    @assert pack([t[3] for t in traced]) == fx[findfirst(a -> a[1] == "out_NAND", fx)][2] "traced bootstrap differs from NAND()"
    write_blob(joinpath(outdir, "$name.fixture.blob"), "fixture", pj, fx)
    @printf("%s: %d gate pairs written\n", name, B)
end

function main()
    outdir = length(ARGS) ≥ 1 ? ARGS[1] : "golden_ref"
    mkpath(outdir)
    all = Dict("CGGIparam" => CGGIparam, "Blockparam" => Blockparam, "CCS2party" => CCS2party, "KMS2party" => KMS2party,
               "KMS2partyblock" => KMS2partyblock)
    names = length(ARGS) ≥ 2 ? ARGS[2:end] : ["CGGIparam", "Blockparam", "CCS2party", "KMS2party", "KMS2partyblock"]
    for name in names
        dump_set(outdir, name, all[name])
    end
end

main()
