#!/usr/bin/env python
"""Static instruction mix of kernels in libhfkmc.so (cuobjdump -sass), the table of profiles/r0N_sass_excerpts.txt.
usage: sass_mix.py [regex of demangled kernel names]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "hyperfluorescence_b200", "libhfkmc.so")
want = re.compile(sys.argv[1] if len(sys.argv) > 1 else ".")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = {}
for fn in re.findall(r"Function : (\S+)", sass):
    names[fn] = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
cur, ops = None, collections.defaultdict(list)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = names[m.group(1)]
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        ops[cur].append(m.group(1))
SPECIAL = ("UTMALDG", "UBLKCP", "UTCMMA", "UTCHMMA", "HMMA", "LDTM", "STTM", "UTMASTG")
for k in sorted(ops):
    if not want.search(k):
        continue
    v = ops[k]
    base = collections.Counter(o.split(".")[0] for o in v)
    print(f"== {k}: {len(v)} instructions")
    print("   " + "  ".join(f"{o} {n}" for o, n in base.most_common(18)))
    g = lambda *p: sum(n for o, n in base.items() if o in p)
    print(f"   spills: STL {base['STL']}  LDL {base['LDL']} | MUFU {base['MUFU']}  DADD {base['DADD']}  DMUL {base['DMUL']}  DFMA {base['DFMA']} | "
          f"LDG {g('LDG', 'LD')}  LDS {base['LDS']}  STS {base['STS']}  SHFL {base['SHFL']} | IMAD.MOV.U32 {sum(1 for o in v if o.startswith('IMAD.MOV'))}")
    sp = [o for o in base if o.startswith(SPECIAL)]
    print("   TMA / tensor-core / TMEM instructions: " + (", ".join(sp) if sp else "none"))
    print()
