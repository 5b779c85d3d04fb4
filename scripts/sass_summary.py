"""Per-kernel SASS mnemonic summary of the in-tree sm_100a objects (no GPU needed):
    python scripts/sass_summary.py > profiles/r2_sass_all_kernels.txt
Lists, for every kernel in every object, the instruction count and the counts of the mnemonics that prove the
Blackwell paths (UTCHMMA/UTCQMMA = tcgen05.mma f16/f8, UTMALDG = TMA load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, HMMA = mma.sync, *.SYS = system-scope peer-memory traffic, RED/ATOM = flag/counter updates)."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KEY = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UTMAPF", "UTMAREDG", "UTMACCTL", "LDTM", "STTM", "SYNCS", "HMMA", "LDSM",
       "LDGSTS", "LDGMC", "UCGABAR", "LDG", "STG", "LD.E", "ST.E", "RED", "ATOM", "MEMBAR", "FENCE", "ERRBAR", "CCTL", "SHFL", "MUFU", "BAR", "ELECT", "PRMT"]


def main():
    objs = sorted((ROOT / "polyrl_b200" / "csrc" / "build").glob("*.o"))
    for obj in objs:
        if obj.name == "bindings.o":
            continue
        out = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
        kern, counts, sysops = None, None, None
        results = []
        for line in out.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                if kern:
                    results.append((kern, counts, sysops))
                kern, counts, sysops = m.group(1), collections.Counter(), collections.Counter()
                continue
            m = re.match(r"\s*/\*[0-9a-f]{4,8}\*/\s+(?:@!?U?P[T\d]+\s+)?([A-Z0-9_.]+)", line)
            if m and kern:
                op = m.group(1)
                counts["_total"] += 1
                for k in KEY:
                    if op.startswith(k):
                        counts[k] += 1
                        break
                if ".SYS" in op or "MULTIMEM" in op or op.startswith("LDGMC") or ".MC" in op or "2CTA" in op:
                    sysops[op] += 1
        if kern:
            results.append((kern, counts, sysops))
        print(f"== {obj.name}")
        for kern, counts, sysops in results:
            dem = subprocess.run(["cu++filt", kern], capture_output=True, text=True).stdout.strip() or kern
            dem = re.sub(r"\s+", " ", dem)[:170]
            keys = " ".join(f"{k}={counts[k]}" for k in KEY if counts[k])
            print(f"  {dem}\n      insts={counts['_total']}  {keys}")
            if sysops:
                print("      system-scope / multicast / 2-CTA forms: " + " ".join(f"{k}x{v}" for k, v in sorted(sysops.items())))
    return 0


if __name__ == "__main__":
    sys.exit(main())
