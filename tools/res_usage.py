"""Per-kernel resource usage of the built library (`cuobjdump -res-usage`): registers, stack (spills), static shared
memory -- the static counterpart of the ncu captures, written to profiles/ by hand:
    python tools/res_usage.py > profiles/r02_kernel_resource_usage.txt"""
import os, re, subprocess, sys

so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "afesp_b200", "lib", "libafesp_gpu.so")
txt = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True, check=True).stdout
dem = {}
rows = []
for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", txt):
    rows.append((m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(5))))
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
print("# %s (sm_100a): %d kernels; STACK > 0 would mean spills or local arrays" % (os.path.basename(so), len(rows)))
print("%-4s %-6s %-7s %-6s  kernel" % ("REG", "STACK", "SHARED", "LOCAL"))
for (mangled, reg, stack, sh, loc), nm in sorted(zip(rows, names), key=lambda x: (-x[0][1], x[1])):
    nm = re.sub(r"\(anonymous namespace\)::", "", nm)
    nm = re.sub(r"\(.*\)$", "", nm)
    print("%-4d %-6d %-7d %-6d  %s" % (reg, stack, sh, loc, nm[:150]))
