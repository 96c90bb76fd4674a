"""Opcode summary of every kernel in libdff_b200.so (cuobjdump -sass): the Blackwell-native evidence the profiling recipe asks for
(UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, HMMA = mma.sync, LDL / STL = spills).
usage: python tools/sass_summary.py [lib.so] > profiles/r02/sass_opcodes.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "two-for-one-diffusion_b200", "dff_b200", "libdff_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, ops = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        ops[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        ops[kern][m.group(1).split(".")[0]] += 1
keys = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "HMMA", "SYNCS", "BAR", "LDGSTS", "LDS", "STS", "LDG", "STG", "SHFL", "FFMA", "MUFU", "LDL", "STL"]
print(f"# cuobjdump -sass {os.path.basename(lib)}: instruction counts per kernel (static)")
print(f"{'kernel':70s} {'total':>7s} " + " ".join(f"{k:>7s}" for k in keys))
for k, c in ops.items():
    print(f"{k[:70]:70s} {sum(c.values()):7d} " + " ".join(f"{c.get(x, 0):7d}" for x in keys))
