"""Refreshes the copy of lean_sqrt / transit_loop inside scripts/micro/transit_bench2.cu from implicitpic1d_b200/csrc/particles.cu."""
import os, re
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
src = open(os.path.join(root, "implicitpic1d_b200", "csrc", "particles.cu")).read()
a = src.index("__device__ __forceinline__ double lean_sqrt(double s) {"); b = src.index("// The whole sweep of one species by one warp")
p = os.path.join(root, "scripts", "micro", "transit_bench2.cu")
t = open(p).read()
i = t.index("// ---- from particles.cu ----\n") + len("// ---- from particles.cu ----\n"); j = t.index("// ---- end of the copy ----")
open(p, "w").write(t[:i] + src[a:b] + t[j:])
