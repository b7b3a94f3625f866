# MKTFHEB200.jl -- drop-in binding of SNUCP/MKTFHE to libmktfhe_b200.so (include/mktfhe_b200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  The Python package mktfhe_b200/ mirrors
# this file call for call (same flattening order, same entry points) and is what the tests exercise.
#
# Usage, next to the reference checkout (nothing in the reference is edited):
#
#     include("src/MKTFHE.jl"); include("MKTFHEB200.jl")
#     using .MKTFHE, .MKTFHEB200
#     a    = CRS(KMS2party)
#     keys = [party_keygen(a, KMS2party) for _ = 1 : KMS2party.k]          # unchanged, host
#     sch  = setup(a, last.(keys), KMS2party)                              # unchanged, host
#     gpu  = MKTFHEB200.upload(sch, KMS2party)                             # one-time key upload
#     res  = NAND(c1, c2, gpu)            # same name and argument order as src/tfhe/gate.jl:1
#     outs = NAND(cs1, cs2, gpu)          # Vector{LWE{UInt32}} -> one batched launch
#     lwe_decrypt(res, first.(keys), KMS2party)                            # unchanged, host
module MKTFHEB200

using ..MKTFHE
import ..MKTFHE: NAND, AND, OR, XOR, XNOR, NOR, bootstrapping!, LWE, TransNativePoly

const LIB = get(ENV, "MKTFHE_B200_LIB", joinpath(@__DIR__, "..", "mktfhe_b200", "lib", "libmktfhe_b200.so"))

# mirrors `mktfhe_params` (include/mktfhe_params.h)
struct CParams
    scheme::Int32; n::Int32; d::Int32; ell::Int32; f::Int32; logD::Int32; N::Int32; k::Int32
    l_gsw::Int32; logB_gsw::Int32; l_lev::Int32; logB_lev::Int32; l_uni::Int32; logB_uni::Int32
    alpha::Float64; beta::Float64
end
const CGGI_, LMSS_, CCS_, KMS_, KMSB_ = Int32(0), Int32(1), Int32(2), Int32(3), Int32(4)

cparams(p::MKTFHE.TFHEparams_bin)   = CParams(CGGI_, p.n, 0, 0, p.f, p.logD, p.N, p.k, p.l_gsw, p.logB_gsw, 0, 0, 0, 0, p.α, p.β)
cparams(p::MKTFHE.TFHEparams_block) = CParams(LMSS_, p.d * p.ℓ, p.d, p.ℓ, p.f, p.logD, p.N, p.k, p.l_gsw, p.logB_gsw, 0, 0, 0, 0, p.α, p.β)
cparams(p::MKTFHE.CCSparams)        = CParams(CCS_, p.n, 0, 0, p.f, p.logD, p.N, p.k, 0, 0, 0, 0, p.l_uni, p.logB_uni, p.α, p.β)
cparams(p::MKTFHE.KMSparams)        = CParams(KMS_, p.n, 0, 0, p.f, p.logD, p.N, p.k, p.l_gsw, p.logB_gsw, p.l_lev, p.logB_lev, p.l_uni, p.logB_uni, p.α, p.β)
cparams(p::MKTFHE.KMSparams_block)  = CParams(KMSB_, p.d * p.ℓ, p.d, p.ℓ, p.f, p.logD, p.N, p.k, p.l_gsw, p.logB_gsw, p.l_lev, p.logB_lev, p.l_uni, p.logB_uni, p.α, p.β)

mutable struct GPUScheme
    h::Ptr{Cvoid}
    words::Int          # 1 + n*k
    function GPUScheme(h, words)
        s = new(h, words)
        finalizer(x -> ccall((:mktfhe_ctx_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.h), s)
        s
    end
end

function check(rc::Cint, h)
    rc == 0 && return
    msg = unsafe_string(ccall((:mktfhe_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
    error("libmktfhe_b200: $rc: $msg")
end

# ---- flattening: SURVEY App. E traversal order = the order the hot path reads the structs -------------------
# A TransNativePoly's `coeffs::Vector{ComplexF64}` is already the H interleaved (re, im) pairs the library wants.
poly!(buf, p::TransNativePoly) = append!(buf, reinterpret(Float64, p.coeffs))

function flatten_rgsw(brk)          # Vector{TransRGSW}: basketb.stack[j].{b, a[1]} then basketa[1].stack[j].{b, a[1]}
    buf = Float64[]
    for g in brk, lev in (g.basketb, g.basketa[1]), row in lev.stack
        poly!(buf, row.b); poly!(buf, row.a[1])
    end
    buf
end
function flatten_unienc(u)          # TransUniEnc: d[j], f.stack[j].b, f.stack[j].a[1]
    buf = Float64[]
    for j in eachindex(u.d)
        poly!(buf, u.d[j]); poly!(buf, u.f.stack[j].b); poly!(buf, u.f.stack[j].a[1])
    end
    buf
end
flatten_polys(v) = (buf = Float64[]; foreach(p -> poly!(buf, p), v); buf)
function flatten_ksk(ksk, n, f)     # Array{LEV}(Dk, N) or (Dk, N, 1): [c][digit][level][b, a...]; undef entries (block) -> zeros
    Dk, N = size(ksk, 1), size(ksk, 2)
    # no slicing: `ksk[:, :, 1]` would touch the #undef entries of a block key (keygen.jl:43-52) and throw UndefRefError
    entry(dg, c) = ndims(ksk) == 3 ? (isassigned(ksk, dg, c, 1) ? ksk[dg, c, 1] : nothing) :
                                     (isassigned(ksk, dg, c) ? ksk[dg, c] : nothing)
    buf = zeros(UInt32, N * Dk * f * (n + 1))
    pos = 1
    for c = 1 : N, dg = 1 : Dk
        lev = entry(dg, c)
        if lev !== nothing
            for lwe in lev.stack
                buf[pos] = lwe.b; buf[pos+1 : pos+n] = lwe.a; pos += n + 1
            end
        else
            pos += f * (n + 1)
        end
    end
    buf
end

"""One-time upload: hooks after `setup` (src/tfhe/scheme.jl:151,190,244,292,343).
`devices = 0:7` puts ONE front context over several GPUs (mktfhe_ctx_create_multi): the keys go host -> first GPU once, are
replicated GPU -> GPU at finalize, and every NAND(cs1, cs2, gpu) batch is sharded in contiguous slices inside the library."""
function upload(scheme, params; device::Integer = 0, devices = nothing)
    cp = Ref(cparams(params))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    if devices === nothing
        check(ccall((:mktfhe_ctx_create, LIB), Cint, (Ref{CParams}, Cint, Ref{Ptr{Cvoid}}), cp, device, h), C_NULL)
    else
        devs = Cint.(collect(devices))
        GC.@preserve devs check(ccall((:mktfhe_ctx_create_multi, LIB), Cint, (Ref{CParams}, Cint, Ptr{Cint}, Ref{Ptr{Cvoid}}),
                                      cp, length(devs), devs, h), C_NULL)
    end
    n = cp[].n
    btks = scheme.btk isa AbstractVector ? scheme.btk : [scheme.btk]
    for (i, btk) in enumerate(btks)
        brk = cp[].scheme == CCS_ ? reduce(vcat, flatten_unienc.(btk.brk)) : flatten_rgsw(btk.brk)
        rlk = hasproperty(btk, :rlk) ? flatten_unienc(btk.rlk) : Float64[]
        pub = hasproperty(btk, :b) ? flatten_polys(btk.b) : Float64[]
        ksk = flatten_ksk(btk.ksk, n, Int(cp[].f))
        GC.@preserve brk rlk pub ksk check(ccall((:mktfhe_upload_party_key, LIB), Cint,
            (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt32}),
            h[], i - 1, brk, isempty(rlk) ? C_NULL : pointer(rlk), isempty(pub) ? C_NULL : pointer(pub), ksk), h[])
    end
    if hasproperty(scheme, :a)
        crs = flatten_polys(scheme.a)
        GC.@preserve crs check(ccall((:mktfhe_upload_common, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), h[], crs), h[])
    end
    check(ccall((:mktfhe_finalize_keys, LIB), Cint, (Ptr{Cvoid},), h[]), h[])
    GPUScheme(h[], 1 + n * (scheme isa MKTFHE.MKscheme ? cp[].k : 1))
end

pack(cs::Vector{LWE{UInt32}}) = reduce(vcat, [vcat(c.b, c.a) for c in cs])
unpack(buf::Vector{UInt32}, words) = [LWE(buf[(g-1)*words+1], buf[(g-1)*words+2 : g*words]) for g = 1 : length(buf) ÷ words]

function gate(op::Integer, c1::Vector{LWE{UInt32}}, c2::Vector{LWE{UInt32}}, s::GPUScheme)
    in1, in2 = pack(c1), pack(c2)
    out = similar(in1)
    GC.@preserve in1 in2 out check(ccall((:mktfhe_gate_batch, LIB), Cint,
        (Ptr{Cvoid}, Cint, Ptr{UInt32}, Ptr{UInt32}, Ptr{UInt32}, Csize_t), s.h, op, in1, in2, out, length(c1)), s.h)
    unpack(out, s.words)
end

for (i, g) in enumerate((:NAND, :AND, :OR, :XOR, :XNOR, :NOR))        # src/tfhe/gate.jl:1-52
    @eval $g(c1::Vector{LWE{UInt32}}, c2::Vector{LWE{UInt32}}, s::GPUScheme) = gate($(i - 1), c1, c2, s)
    @eval $g(c1::LWE{UInt32}, c2::LWE{UInt32}, s::GPUScheme) = gate($(i - 1), [c1], [c2], s)[1]
end

"""bootstrapping!(ctxt, scheme) -- src/tfhe/bootstrapping.jl:4-27; mutates ctxt like the reference."""
function bootstrapping!(c::LWE{UInt32}, s::GPUScheme)
    buf = vcat(c.b, c.a)
    GC.@preserve buf check(ccall((:mktfhe_bootstrap_batch, LIB), Cint,
        (Ptr{Cvoid}, Ptr{UInt32}, Ptr{UInt32}, Csize_t), s.h, buf, buf, 1), s.h)
    c.b = buf[1]; c.a .= @view buf[2:end]
    c
end

# ---- circuits: one call per level of independent gates over a device-resident wire table (include/mktfhe_b200.h) ----
# ops: 0..5 = NAND..NOR, -1 = bootstrapping! of src1, 6 = NOT! (no bootstrap).  Wire indices are 0-based rows of the table.
wires_resize(s::GPUScheme, n::Integer) =
    check(ccall((:mktfhe_wires_resize, LIB), Cint, (Ptr{Cvoid}, Csize_t), s.h, n), s.h)

function wires_write(s::GPUScheme, first::Integer, cs::Vector{LWE{UInt32}})
    buf = pack(cs)
    GC.@preserve buf check(ccall((:mktfhe_wires_write, LIB), Cint,
        (Ptr{Cvoid}, Csize_t, Csize_t, Ptr{UInt32}), s.h, first, length(cs), buf), s.h)
end

function wires_read(s::GPUScheme, first::Integer, count::Integer)
    buf = Vector{UInt32}(undef, count * s.words)
    GC.@preserve buf check(ccall((:mktfhe_wires_read, LIB), Cint,
        (Ptr{Cvoid}, Csize_t, Csize_t, Ptr{UInt32}), s.h, first, count, buf), s.h)
    unpack(buf, s.words)
end

function gate_level(s::GPUScheme, ops::Vector{Int32}, src1::Vector{Int32}, src2::Vector{Int32}, dst::Vector{Int32})
    GC.@preserve ops src1 src2 dst check(ccall((:mktfhe_gate_level, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Csize_t), s.h, ops, src1, src2, dst, length(ops)), s.h)
end

end # module
