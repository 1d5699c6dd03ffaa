#!/usr/bin/env python3
"""Rewrite OpenCL C vector literals so g++ can parse a reference .ocl file.

TEST INFRASTRUCTURE ONLY.  Reads a reference kernel file where it lies (e.g.
/root/reference/CLSuperPathTracer/pathtracer.ocl), writes C++ to stdout (the
Makefile redirects it into oracle/_ref/gen/, which is git-ignored: reference
sources are never committed to this repo).

The transformations are  `(vecN)(a, b, ...)`  ->  `vecN{a, b, ...}`  for the
vector types the hot-path kernels use, and the swizzle `.s012` -> `.s012()` (a member
function of the shim's vec4; the bidirectional kernel uses it).  Brace initialisation keeps the
left-to-right evaluation order that clang-based OpenCL compilers give the two
RNG calls inside the float4 literal at pathtracer.ocl:233; a function-call
rewrite would let g++ evaluate them right-to-left.
"""
import re
import sys

VEC_TYPES = ("float2", "float4", "float8", "uint2", "uint4", "int2", "int4", "uchar4")
PAT = re.compile(r"\(\s*(" + "|".join(VEC_TYPES) + r")\s*\)\s*\(")


def rewrite(src: str) -> str:
    out = []
    pos = 0
    closers = {}  # index of ')' that must become '}'
    i = 0
    text = src
    # first pass: find every literal and the index of its matching ')'
    opens = {}
    for m in PAT.finditer(text):
        start_paren = m.end() - 1
        depth = 0
        j = start_paren
        while j < len(text):
            c = text[j]
            if c == "(":
                depth += 1
            elif c == ")":
                depth -= 1
                if depth == 0:
                    break
            j += 1
        if depth != 0:
            raise SystemExit("unbalanced vector literal at offset %d" % m.start())
        opens[m.start()] = (m.end(), m.group(1))
        closers[j] = True
    while i < len(text):
        if i in opens:
            end, ty = opens[i]
            out.append(ty + "{")
            i = end
            continue
        if i in closers:
            out.append("}")
            i += 1
            continue
        out.append(text[i])
        i += 1
    return re.sub(r"\.s012\b", ".s012()", "".join(out))


if __name__ == "__main__":
    sys.stdout.write(rewrite(open(sys.argv[1]).read()))
