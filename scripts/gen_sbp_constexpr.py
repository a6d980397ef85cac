#!/usr/bin/env python3
"""sw4lite_b200/csrc/sbp4_tables.h (sparse exact rationals) -> sbp4_constexpr.h (the same tables as
constexpr functions for compile-time folding in k_closure_fast).  The generated header is committed."""
import os, re
HERE = os.path.dirname(os.path.abspath(__file__))
CS = os.path.join(HERE, "..", "sw4lite_b200", "csrc")
s = open(os.path.join(CS, "sbp4_tables.h")).read()


def tab(name):
    body = re.search(name + r'\[\d+\]\s*=\s*\{(.*?)\};', s, re.S).group(1)
    return [(int(a), b, c) for a, b, c in re.findall(r'\{\s*(\d+)\s*,\s*(-?[0-9.]+)\s*,\s*([0-9.]+)\s*\}', body)]


out = ["// GENERATED from sbp4_tables.h (scripts/gen_sbp_constexpr.py) -- do not edit.",
       "// The same tables as compile-time functions: with constant arguments (fully unrolled loops) the",
       "// compiler folds the values into the instruction stream and drops the zero entries altogether.",
       "// Used by k_closure_fast only when the runtime tables equal these built-in ones (api.cu checks).",
       "#ifndef SW4B200_SBP4_CONSTEXPR_H\n#define SW4B200_SBP4_CONSTEXPR_H", "namespace sw4b200 {"]
for fn, name in (("acof_c", "SW4B200_ACOF_NZ"), ("bope_c", "SW4B200_BOPE_NZ"), ("ghcof_c", "SW4B200_GHCOF_NZ")):
    out.append("__host__ __device__ constexpr double %s( int idx )\n{\n   switch( idx )\n   {" % fn)
    out += ["   case %d: return %s / %s;" % (i, n, d) for i, n, d in tab(name)]
    out.append("   default: return 0.0;\n   }\n}")
out.append("} // namespace sw4b200\n#endif")
open(os.path.join(CS, "sbp4_constexpr.h"), "w").write("\n".join(out) + "\n")
