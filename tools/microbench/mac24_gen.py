# Prototype for round 2 (not built by default): column-wise (product-scanning) multiply-accumulate + Montgomery reduction with a
# 25-word accumulator instead of the 59-register (e, o, c) form.  `python mac24_gen.py` writes /tmp/mac24/fp24_gen.cuh;
# mac24_regs.cu includes it: nvcc -Xptxas -v reports 72 registers, 0 spills under __launch_bounds__(384, 2); SASS per product +
# reduction: 277 IMAD.WIDE.U32 + 238 IADD3[.X] (the (e, o, c) MAC: 300 IMAD.WIDE + ~35 others, 128 registers with the interpreter).
# generates acc24 primitives (column-wise product scanning) as PTX asm
P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
PL = [(P >> (32 * i)) & 0xFFFFFFFF for i in range(12)]
N0 = (-pow(P, -1, 1 << 32)) % (1 << 32)
out=[]; w=out.append
w("// GENERATED"); w("#pragma once"); w("#include <cstdint>"); w("namespace fpc24 {")
w("struct Acc { uint32_t a[25]; };")
w("__device__ __forceinline__ void acc_zero(Acc& A) {\n#pragma unroll\n  for (int i = 0; i < 25; ++i) A.a[i] = 0; }")
# acc_mac: one asm statement for the whole product? too many operands (25 + 24 = 49 > 30 limit?). PTX asm operand limit is 30. So per column statements with (t64, t2) carried in C variables.
# column k: inputs: up to 12 x and 12 y; we pass the needed ones.
w("__device__ __forceinline__ void acc_mac(Acc& A, const uint32_t* __restrict__ x, const uint32_t* __restrict__ y) {")
w("  uint64_t t = 0; uint32_t t2 = 0;")
for k in range(23):
    pairs=[(i,k-i) for i in range(12) if 0<=k-i<12]
    # t += A[k] (into low word with carry), then products, then A[k]=lo(t); shift
    ops=[]
    lines=[".reg .u32 l,h;", "mov.b64 {l,h}, %0;", "add.cc.u32 l, l, %2;", "addc.cc.u32 h, h, 0;", "addc.u32 %1, %1, 0;"]
    n=3
    xs=[]; 
    for (i,j) in pairs:
        lines.append(f"mad.lo.cc.u32 l, %{n}, %{n+1}, l;")
        lines.append(f"madc.hi.cc.u32 h, %{n}, %{n+1}, h;")
        lines.append(f"addc.u32 %1, %1, 0;")
        xs.append(f'"r"(x[{i}]), "r"(y[{j}])')
        n+=2
    # output A[k] = l ; shift: t = (h, t2), t2 = 0
    lines.append("mov.u32 %2, l;")
    lines.append("mov.b64 %0, {h, %1};")
    lines.append("mov.u32 %1, 0;")
    body='"{\\n\\t"\n    "' + '\\n\\t"\n    "'.join(lines) + '\\n\\t"\n    "}"'
    w(f"  asm({body}\n    : \"+l\"(t), \"+r\"(t2), \"+r\"(A.a[{k}])\n    : {', '.join(xs)});")
# remaining: A[23] += lo(t), A[24] += hi(t) + carry
w('  asm("{\\n\\t.reg .u32 l,h;\\n\\tmov.b64 {l,h}, %2;\\n\\tadd.cc.u32 %0, %0, l;\\n\\taddc.u32 %1, %1, h;\\n\\t}" : "+r"(A.a[23]), "+r"(A.a[24]) : "l"(t));')
w("}")
# redc: column-wise
w("__device__ __forceinline__ void acc_redc(Acc& A, uint32_t* __restrict__ r) {")
w("  uint64_t t = 0; uint32_t t2 = 0; uint32_t m[12];")
for k in range(12):
    lines=[".reg .u32 l,h;", "mov.b64 {l,h}, %0;", "add.cc.u32 l, l, %3;", "addc.cc.u32 h, h, 0;", "addc.u32 %1, %1, 0;"]
    n=4; ins=[]
    for i in range(k):
        lines.append(f"mad.lo.cc.u32 l, %{n}, 0x{PL[k-i]:08x}, l;")
        lines.append(f"madc.hi.cc.u32 h, %{n}, 0x{PL[k-i]:08x}, h;")
        lines.append("addc.u32 %1, %1, 0;")
        ins.append(f'"r"(m[{i}])'); n+=1
    lines.append(f"mul.lo.u32 %2, l, 0x{N0:08x};")
    lines.append(f"mad.lo.cc.u32 l, %2, 0x{PL[0]:08x}, l;")
    lines.append(f"madc.hi.cc.u32 h, %2, 0x{PL[0]:08x}, h;")
    lines.append("addc.u32 %1, %1, 0;")
    lines.append("mov.b64 %0, {h, %1};")
    lines.append("mov.u32 %1, 0;")
    body='"{\\n\\t"\n    "' + '\\n\\t"\n    "'.join(lines) + '\\n\\t"\n    "}"'
    ins_s = (", " + ", ".join(ins)) if ins else ""
    w(f"  asm({body}\n    : \"+l\"(t), \"+r\"(t2), \"=r\"(m[{k}])\n    : \"r\"(A.a[{k}]){ins_s});")
for k in range(12,24):
    lines=[".reg .u32 l,h;", "mov.b64 {l,h}, %0;", "add.cc.u32 l, l, %3;", "addc.cc.u32 h, h, 0;", "addc.u32 %1, %1, 0;"]
    n=4; ins=[]
    for i in range(k-11,12):
        lines.append(f"mad.lo.cc.u32 l, %{n}, 0x{PL[k-i]:08x}, l;")
        lines.append(f"madc.hi.cc.u32 h, %{n}, 0x{PL[k-i]:08x}, h;")
        lines.append("addc.u32 %1, %1, 0;")
        ins.append(f'"r"(m[{i}])'); n+=1
    lines.append("mov.u32 %2, l;")
    lines.append("mov.b64 %0, {h, %1};")
    lines.append("mov.u32 %1, 0;")
    body='"{\\n\\t"\n    "' + '\\n\\t"\n    "'.join(lines) + '\\n\\t"\n    "}"'
    ins_s = (", " + ", ".join(ins)) if ins else ""
    w(f"  asm({body}\n    : \"+l\"(t), \"+r\"(t2), \"=r\"(r[{k-12}])\n    : \"r\"(A.a[{k}]){ins_s});")
w("  // the 25th word and the running accumulator hold the overflow above 2^384: must be zero by the caller's bound (k*p < 2^384)")
w("}")
w("}")
import os
os.makedirs("/tmp/mac24", exist_ok=True)
open("/tmp/mac24/fp24_gen.cuh",'w').write("\n".join(out)+"\n")
