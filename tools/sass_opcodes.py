#!/usr/bin/env python
"""Per-kernel SASS opcode counts of libgsb.so (cuobjdump -sass): the data-movement and special-function instructions the
north_star names.  usage: python tools/sass_opcodes.py [path/to/lib.so] > profiles/sass_opcodes.md"""
import collections, re, subprocess, sys, os
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gsorb_slam_b200", "libgsb.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
OPS = ["LDGSTS", "UBLKCP", "UTMALDG", "SYNCS", "LDG", "STG", "LDS", "STS", "RED", "REDG", "ATOMG", "ATOMS", "MUFU.EX2", "MUFU.RCP", "SHFL", "VOTE",
       "BAR", "VIMNMX", "BMSK", "PRMT", "REDUX", "LDGMC", "MULTIMEM"]
kern, counts, total = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1); counts[kern] = collections.Counter(); total[kern] = 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1); total[kern] += 1
        for o in OPS:
            if op == o or op.startswith(o + ".") or (o == "MULTIMEM" and "MULTIMEM" in op):
                counts[kern][o] += 1
print("# SASS opcode counts per kernel (`cuobjdump -sass gsorb_slam_b200/libgsb.so`, sm_100a)\n")
print("LDGSTS = cp.async (16-byte global->shared gathers of the splat records); UBLKCP = cp.async.bulk (TMA unit, 1-D); REDG = vector /")
print("scalar `red.global.add`; the `multimem.*` instructions of the exchange kernel appear as LDGMC / REDG-class opcodes with `.MC`.\n")
cols = ["LDGSTS", "UBLKCP", "UTMALDG", "LDG", "STG", "LDS", "STS", "RED", "REDG", "ATOMG", "MUFU.EX2", "MUFU.RCP", "SHFL", "VOTE", "BAR", "VIMNMX", "BMSK", "PRMT", "REDUX"]
print("| kernel | instr | " + " | ".join(cols) + " |")
print("|---|---|" + "---|" * len(cols))
for k, c in counts.items():
    name = demangle(k)
    name = re.sub(r"\(.*", "", name).replace("void gsb::", "").replace("gsb::", "")
    print(f"| `{name[:70]}` | {total[k]} | " + " | ".join(str(c[o]) if c[o] else "" for o in cols) + " |")
mm = [l for l in out.splitlines() if "MULTIMEM" in l.upper() or ".MC" in l]
print(f"\nInstructions carrying a multicast (NVSwitch) address in the exchange kernels: {len(mm)}")
for l in mm[:6]:
    print("    " + re.sub(r"/\*[0-9a-f]{16}\*/", "", l).strip())
