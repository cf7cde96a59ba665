#!/usr/bin/env python3
"""Generates fr_asm.cuh: fully unrolled PTX carry-chain blocks for BLS12-381 Fr Montgomery arithmetic on sm_100a.

Why generated: the fast path is built from `mad.lo.cc/madc.hi.cc` pairs that ptxas fuses into ONE
`IMAD.WIDE.U32(.X)` each (measured on B200: 32 lane-ops/clk/SM, vs 64 for IMAD.LO and 32 for IMAD.HI, see
tools/microbench/pipes.cu), which only happens when every (lo,hi) accumulator pair stays on an aligned register
pair.  That is arranged by keeping two accumulators — EVEN (products landing on even 32-bit limb positions) and ODD
(odd positions, stored shifted down by one limb) — and by emitting each carry chain as one asm statement so the
carry flag never crosses statement boundaries.

Run:  python gen_fr_asm.py > fr_asm.cuh
"""
P32 = [0x00000001, 0xFFFFFFFF, 0xFFFE5BFE, 0x53BDA402, 0x09A1D805, 0x3339D808, 0x299D7D48, 0x73EDA753]

out = []
emit = out.append


def asm_block(lines, outs, ins, indent="    ", fresh=()):
    """outs: list of C lvalues used as "+r"; ins: list of C rvalues used as "r". `lines` use {o0}.. / {i0}.. names."""
    names = {}
    for k, _ in enumerate(outs):
        names[f"o{k}"] = f"%{k}"
    for k, _ in enumerate(ins):
        names[f"i{k}"] = f"%{len(outs) + k}"
    assert len(outs) + len(ins) <= 30, "asm operand limit"
    body = "\n".join(f'{indent}    "{ln.format(**names)};\\n\\t"' for ln in lines)
    o = ", ".join((f'"=r"({x})' if x in fresh else f'"+r"({x})') for x in outs)
    i = ", ".join(f'"r"({x})' for x in ins)
    emit(f"{indent}asm(\n{body}\n{indent}    : {o}\n{indent}    : {i});")


WRITTEN = None  # when a dict {"ev": set, "od": set}: track first touches so fresh limbs are produced, not accumulated


def chain(acc, base, n_prod, a_ops, b_op, prop_to, first_is_mul=False, first_pair=None, fresh=()):
    """acc[base .. base+2*n_prod) += a_ops[k] * b_op (k-th product on limbs base+2k, base+2k+1), carry chained, then
    carry propagated through acc[base+2*n_prod .. prop_to] (inclusive)."""
    outs = [f"{acc}[{base + k}]" for k in range(2 * n_prod)]
    tail = [f"{acc}[{k}]" for k in range(base + 2 * n_prod, prop_to + 1)]
    if WRITTEN is not None:
        touched = list(range(base, base + 2 * n_prod)) + list(range(base + 2 * n_prod, prop_to + 1))
        fresh = tuple(k for k in touched if k not in WRITTEN[acc])
        WRITTEN[acc].update(touched)
    fresh_names = {f"{acc}[{k}]" for k in fresh}
    ins = list(a_ops) + [b_op]
    nb = len(a_ops)
    if first_pair:
        ins += list(first_pair)
    lines = []
    for k in range(n_prod):
        lo, hi = f"{{o{2 * k}}}", f"{{o{2 * k + 1}}}"
        a, b = f"{{i{k}}}", f"{{i{nb}}}"
        last = (k == n_prod - 1) and not tail
        if first_is_mul:
            fresh_names.update(outs)
        if k == 0 and first_pair:
            # the product a_ops[0]*b_op is supplied as a ready (lo, hi) pair: two carry-chained adds, no multiply
            lines.append(f"add.cc.u32 {lo}, {lo}, {{i{nb + 1}}}")
            lines.append(f"addc.cc.u32 {hi}, {hi}, {{i{nb + 2}}}")
        elif first_is_mul:
            lines.append(f"mul.lo.u32 {lo}, {a}, {b}")
            lines.append(f"mul.hi.u32 {hi}, {a}, {b}")
        else:
            lo_add = "0" if outs[2 * k] in fresh_names else lo
            hi_add = "0" if outs[2 * k + 1] in fresh_names else hi
            lines.append(("mad.lo.cc.u32" if k == 0 else "madc.lo.cc.u32") + f" {lo}, {a}, {b}, {lo_add}")
            lines.append(("madc.hi.u32" if last else "madc.hi.cc.u32") + f" {hi}, {a}, {b}, {hi_add}")
    for k, t in enumerate(tail):
        o = f"{{o{2 * n_prod + k}}}"
        src = "0" if t in fresh_names else o
        lines.append(("addc.u32" if k == len(tail) - 1 else "addc.cc.u32") + f" {o}, {src}, 0")
    asm_block(lines, outs + tail, ins, fresh=fresh_names)


def add_chain(dst, src, cin=None, cout=None, indent="    "):
    """dst[k] += src[k] (src entry None = 0) with a carry chain; optional carry-in / carry-out through registers so long
    chains can be split to respect the 30-operand asm limit."""
    outs = list(dst) + ([cout] if cout else [])
    ins = [x for x in src if x is not None] + ([cin] if cin else [])
    names = {}
    for k, _ in enumerate(outs):
        names[f"o{k}"] = f"%{k}"
    m = {}
    for k, x in enumerate(ins):
        m[x] = f"%{len(outs) + k}"
    assert len(outs) + len(ins) <= 30
    body = []
    if cin:  # re-materialise the carry flag from a 0/1 register
        body.append("{ .reg .u32 t_; add.cc.u32 t_, " + m[cin] + ", 0xffffffff; }")
    for k, d in enumerate(dst):
        o = names[f"o{k}"]
        sx = m[src[k]] if src[k] is not None else "0"
        first = (k == 0 and not cin)
        last = (k == len(dst) - 1 and not cout)
        op = "add" if first else "addc"
        body.append(f"{op}{'' if last else '.cc'}.u32 {o}, {o}, {sx}")
    if cout:
        body.append(f"addc.u32 {names[f'o{len(dst)}']}, 0, 0")
    text = "\n".join(f'{indent}    "{ln}{"" if ln.endswith("}") else ";"}\\n\\t"' for ln in body)
    o = ", ".join(f'"+r"({x})' for x in dst) + (f', "=r"({cout})' if cout else "")
    i = ", ".join(f'"r"({x})' for x in ins)
    emit(f"{indent}asm(\n{text}\n{indent}    : {o}\n{indent}    : {i});")


emit("// GENERATED by gen_fr_asm.py — do not edit.  BLS12-381 Fr Montgomery kernels as PTX carry chains (see generator).")
emit("#pragma once")
emit("namespace fr {")
emit("")
emit("// Deliberately NOT initialised at compile time: with literal limbs ptxas switches the reduction chains to the")
emit("// immediate forms IMAD.X + IMAD.HI.U32.X (3 issue-slots per product) instead of IMAD.WIDE.U32.X (2).  The host")
emit("// fills it once per device via fr_init_constants().")
emit("__constant__ uint32_t FR_MODULUS[9];  // [8] = 0: an \"opaque zero\" — see redc_eo")
emit("static const uint32_t FR_MODULUS_HOST[9] = {" + ", ".join(f"0x{v:08x}u" for v in P32) + ", 0u};")
emit("static inline cudaError_t fr_init_constants() {")
emit("    return cudaMemcpyToSymbol(FR_MODULUS, FR_MODULUS_HOST, sizeof(FR_MODULUS_HOST));")
emit("}")
emit("")
emit("// ev[k] = limb k of the EVEN accumulator, od[k] = limb k+1 of the ODD accumulator; a*b = EV + OD*2^32.")
emit("// Every chain's (lo,hi) destination pair starts on an even index, so ptxas emits IMAD.WIDE.U32(.X).")
emit("__device__ __forceinline__ void mul_wide_eo(uint32_t (&ev)[16], uint32_t (&od)[16], const uint32_t (&a)[8], const uint32_t (&b)[8]) {")
emit("    od[15] = 0;  // every other limb is produced (not accumulated) on first touch")
WRITTEN = {"ev": set(), "od": set()}
A_EVEN = ["a[0]", "a[2]", "a[4]", "a[6]"]
A_ODD = ["a[1]", "a[3]", "a[5]", "a[7]"]
# row 0: plain products
chain("ev", 0, 4, A_EVEN, "b[0]", -1, first_is_mul=True)
chain("od", 0, 4, A_ODD, "b[0]", -1, first_is_mul=True)
for i in range(1, 8):
    if i % 2 == 1:
        # even-j products land on odd limb i+j -> od index i+j-1; top limb i+7 may carry into od[i+7]
        chain("od", i - 1, 4, A_EVEN, f"b[{i}]", i + 7)
        # odd-j products land on even limb i+j; top limb i+8 cannot carry (partial sum < 2^(32(i+9)))
        chain("ev", i + 1, 4, A_ODD, f"b[{i}]", -1)
    else:
        chain("ev", i, 4, A_EVEN, f"b[{i}]", i + 8)
        chain("od", i, 4, A_ODD, f"b[{i}]", -1)
assert WRITTEN["ev"] == set(range(16)) and WRITTEN["od"] == set(range(15)), WRITTEN
WRITTEN = None
emit("}")
emit("")

# ---- accumulate variant: EV/OD hold running sums of several products; every chain propagates to the top limb.
# ---- 4 x 4 limbs (128 x 128 bits): the building block of the one-level Karatsuba product of gemm_sum.cuh
emit("// 128 x 128-bit product, same even/odd scheme: a*b = EV + OD*2^32 (ev[0..7], od[0..6]; od[7] = 0).")
emit("__device__ __forceinline__ void mul_wide_eo4(uint32_t (&ev)[8], uint32_t (&od)[8], const uint32_t (&a)[4], const uint32_t (&b)[4]) {")
emit("    od[7] = 0;")
WRITTEN = {"ev": set(), "od": set()}
A4_EVEN = ["a[0]", "a[2]"]
A4_ODD = ["a[1]", "a[3]"]
chain("ev", 0, 2, A4_EVEN, "b[0]", -1, first_is_mul=True)
chain("od", 0, 2, A4_ODD, "b[0]", -1, first_is_mul=True)
for i in range(1, 4):
    if i % 2 == 1:
        chain("od", i - 1, 2, A4_EVEN, f"b[{i}]", i + 3)
        chain("ev", i + 1, 2, A4_ODD, f"b[{i}]", -1)
    else:
        chain("ev", i, 2, A4_EVEN, f"b[{i}]", i + 4)
        chain("od", i, 2, A4_ODD, f"b[{i}]", -1)
assert WRITTEN["ev"] == set(range(8)) and WRITTEN["od"] == set(range(7)), WRITTEN
WRITTEN = None
emit("}")
emit("")
emit("// ev/od (17 limbs each: 16 + one overflow limb) += a*b.  For lazily reduced inner products.")
emit("__device__ __forceinline__ void mac_wide_eo(uint32_t (&ev)[17], uint32_t (&od)[17], const uint32_t (&a)[8], const uint32_t (&b)[8]) {")
for i in range(8):
    if i % 2 == 1:
        chain("od", i - 1, 4, A_EVEN, f"b[{i}]", 16)
        chain("ev", i + 1, 4, A_ODD, f"b[{i}]", 16)
    else:
        chain("ev", i, 4, A_EVEN, f"b[{i}]", 16)
        chain("od", i, 4, A_ODD, f"b[{i}]", 16)
emit("}")
emit("")

# ---- Montgomery reduction, 32-bit digits.  -p^-1 mod 2^32 = 0xffffffff so m_i = -S_i; p[0] = 1 so the digit-0 product is
# folded into the carry c_{i+1} = (S_i + m_i) >> 32 and never issued.
for top, name, ndig in ((15, "redc_eo", 8), (16, "redc_eo17", 8), (9, "redc2_eo10", 2)):
    n = top + 1
    emit(f"// In place: (EV + OD*2^32) += M*p with M chosen so the low {32 * ndig} bits vanish; returns the carry into limb {ndig}.")
    emit(f"// Afterwards limbs {ndig}..{top} of EV and {ndig - 1}..{top - 1} of OD hold the (unmerged) quotient.")
    emit(f"__device__ __forceinline__ uint32_t {name}(uint32_t (&ev)[{n}], uint32_t (&od)[{n}]) {{")
    emit("    uint32_t c = 0, m, s, k1, k2, nz, h1;")
    emit("    uint32_t pc[8];  // modulus limbs from constant memory: register/uniform operands keep the IMAD.WIDE fusion")
    emit("#pragma unroll")
    emit("    for (int k = 0; k < 8; k++) pc[k] = FR_MODULUS[k];")
    emit("    // m = -S must not LOOK like a negation: ptxas folds `0 - s` into the multiply as a negated operand, which")
    emit("    // un-fuses every IMAD.WIDE of the chain into IMAD + IMAD.HI.  Subtracting from a zero it cannot see avoids that.")
    emit("    const uint32_t zero = FR_MODULUS[8];")
    for i in range(ndig):
        prev = f"od[{i - 1}]" if i > 0 else "0u"
        emit(f"    // digit {i}: S = ev[{i}] + {prev} + c")
        emit(f"    s = ev[{i}] + {prev}; k1 = (s < ev[{i}]) ? 1u : 0u; s += c; k2 = (s < c) ? 1u : 0u;")
        emit("    nz = (s != 0u) ? 1u : 0u; m = zero - s; c = k1 + k2 + nz;")
        emit("    h1 = m - nz;  // m * p[1] = m * (2^32 - 1) = (h1 << 32) | s : no multiply needed")
        podd = [f"pc[{j}]" for j in (1, 3, 5, 7)]
        peven = [f"pc[{j}]" for j in (2, 4, 6)]
        if i % 2 == 0:
            chain("od", i, 4, podd, "m", top, first_pair=("s", "h1"))       # limbs i+1,i+3,i+5,i+7 -> od idx i..i+7
            chain("ev", i + 2, 3, peven, "m", top)   # limbs i+2,i+4,i+6
        else:
            chain("ev", i + 1, 4, podd, "m", top, first_pair=("s", "h1"))   # limbs i+1.. (even)
            chain("od", i + 1, 3, peven, "m", top)   # limbs i+2,i+4,i+6 (odd) -> od idx i+1..
    emit("    return c;")
    emit("}")
    emit("")



# ---- multiplication by a per-round constant (the fold): V = sum_k d[k] * C[k], all rows land on limb 0.
emit("// ev/od (10 limbs) = sum_k d[k] * C[k] where C[k] (8 limbs each, shared memory, warp-uniform) are the per-round")
emit("// constants r * 2^(32k+64) mod p: then V == r * d * 2^64 (mod p) with V < 2^290, and a TWO-digit Montgomery")
emit("// reduction (redc2_eo10) brings it to r*d mod p — 64 + 12 IMAD.WIDE instead of 63 + 48 for a general product.")
emit("__device__ __forceinline__ void mulc_rows_eo(uint32_t (&ev)[10], uint32_t (&od)[10], const uint32_t* C, const uint32_t (&d)[8]) {")
emit("    uint32_t c[8];")
for k in range(8):
    emit(f"    {{ const uint4 lo = *reinterpret_cast<const uint4*>(C + {8 * k}), hi = *reinterpret_cast<const uint4*>(C + {8 * k + 4});")
    emit("      c[0] = lo.x; c[1] = lo.y; c[2] = lo.z; c[3] = lo.w; c[4] = hi.x; c[5] = hi.y; c[6] = hi.z; c[7] = hi.w; }")
    if k == 0:
        chain("ev", 0, 4, ["c[0]", "c[2]", "c[4]", "c[6]"], "d[0]", -1, first_is_mul=True)
        chain("od", 0, 4, ["c[1]", "c[3]", "c[5]", "c[7]"], "d[0]", -1, first_is_mul=True)
        emit("    ev[8] = 0; ev[9] = 0; od[8] = 0; od[9] = 0;")
    else:
        chain("ev", 0, 4, ["c[0]", "c[2]", "c[4]", "c[6]"], f"d[{k}]", 9)
        chain("od", 0, 4, ["c[1]", "c[3]", "c[5]", "c[7]"], f"d[{k}]", 9)
emit("}")
emit("")
emit("// w[0..16] += EV + OD*2^32 where (ev, od) = a*b: the merge of the two accumulators and the accumulation in two")
emit("// split carry chains each (asm statements are limited to 30 operands).")
emit("__device__ __forceinline__ void wide_mac_limbs(uint32_t (&w)[17], const uint32_t (&a)[8], const uint32_t (&b)[8]) {")
emit("    uint32_t ev[16], od[16], c0, c1;")
emit("    mul_wide_eo(ev, od, a, b);")
add_chain([f"w[{k}]" for k in range(0, 8)], [f"ev[{k}]" for k in range(0, 8)], cout="c0")
add_chain([f"w[{k}]" for k in range(8, 17)], [f"ev[{k}]" for k in range(8, 16)] + [None], cin="c0")
add_chain([f"w[{k}]" for k in range(1, 9)], [f"od[{k}]" for k in range(0, 8)], cout="c1")
add_chain([f"w[{k}]" for k in range(9, 17)], [f"od[{k}]" for k in range(8, 16)], cin="c1")
emit("}")
emit("")
emit("}  // namespace fr")
print("\n".join(out))
