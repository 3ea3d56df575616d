"""Per-kernel SASS mnemonic counts of the built library (no GPU needed): evidence that the hot kernels are tcgen05 / TMEM /
TMA code.  UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,
SYNCS.* = mbarrier, UTCATOMSWS = TMEM alloc / dealloc, HMMA = mma.sync (the short-sequence attention kernel only).
Usage: python tools/sass_mnemonics.py [path/to/libaedit.so] > profiles/<name>.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "audioeditingcode_b200", "libaedit.so")
KEEP = re.compile(r"^(UTCHMMA|UTMALDG|UTMASTG|LDTM|STTM|UTCBAR|UTCATOMSWS|SYNCS|HMMA|MUFU\.EX2|LDG\.E\.128|STG\.E\.128|"
                  r"LDS\.128|STS\.128|LDSM|ATOM|RED|BAR\.SYNC|PREEXIT|ACQBULK|ELECT)")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
per = collections.defaultdict(collections.Counter)
inst = collections.Counter()
kern = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "").replace("aedit::", "")
        name = re.sub(r"\(.*", "", name)
        kern = re.sub(r"<.*", "", name)
        inst[kern] += 1
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)", line)
    if m and kern and KEEP.match(m.group(1)):
        per[kern][m.group(1)] += 1
print(f"# {os.path.basename(lib)}: SASS mnemonics per kernel family (summed over template instantiations)")
for k in sorted(per, key=lambda k: -sum(per[k].values())):
    print(f"== {k}  ({inst[k]} instantiation{'s' if inst[k] != 1 else ''})")
    for op, n in per[k].most_common():
        print(f"{n:8d} {op}")
