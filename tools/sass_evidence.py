"""profiles/r02_sass_tcgen05.txt: Blackwell-native SASS mnemonics per kernel of the built library (cuobjdump -sass)."""
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "pytracking_b200/libb200trk.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
print("# r02: cuobjdump -sass %s (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a) -- Blackwell-native mnemonics per kernel" % so)
print("# UTCHMMA = tcgen05.mma (kind::tf32), LDTM / STTM = tcgen05.ld / tcgen05.st, UTMALDG = cp.async.bulk.tensor (TMA), UTCBAR = tcgen05.commit, MAPA / UCGABAR = cluster DSMEM / barrier.cluster")
print()
keys = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "UCGABAR", "MAPA", "LDS.128", "SYNCS"]
for m in re.finditer(r"Function : (\S+)\n(.*?)(?=\n\s*Function : |\Z)", txt, re.S):
    name, body = m.group(1), m.group(2)
    cnt = {k: len(re.findall(r"\b" + re.escape(k), body)) for k in keys}
    if cnt["UTCHMMA"] == 0 and cnt["UTMALDG"] == 0 and cnt["LDTM"] == 0:
        continue
    print(demangle(name))
    print("   " + "  ".join("%s x%d" % (k, v) for k, v in cnt.items() if v))
    for k in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "MAPA"):
        l = next((ln for ln in body.splitlines() if re.search(r"\b" + re.escape(k), ln)), None)
        if l:
            print("      e.g. " + l.strip()[:150])
    print()
