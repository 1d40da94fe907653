"""Per-kernel SASS evidence for profiles/: counts of the Blackwell mnemonics and of local-memory spills.
usage: python tools_sass_summary.py [object_intrinsics_b200/lib/liboi_b200.so] > profiles/rNN_sass_summary.txt
(runs where cuobjdump is; no GPU needed)."""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "object_intrinsics_b200/lib/liboi_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
WANT = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "MUFU.SIN", "MUFU.COS",
        "FFMA2", "FMUL2", "FADD2", "HMMA", "STL", "LDL", "CCTL", "RED", "ATOM", "BAR.SYNC", "UCGABAR", "LDG", "STG", "LDS", "STS"]
kern, counts, total = None, collections.OrderedDict(), {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern], total[kern] = collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if kern and m:
        op = m.group(1)
        total[kern] += 1
        for w in WANT:
            if op == w or op.startswith(w + "."):
                counts[kern][w] += 1
regs = {}
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", line)
    if m and cur:
        regs[cur] = m.groups()
print(f"# SASS summary of {so} (cuobjdump -sass / -res-usage; sm_100a)")
for k, c in counts.items():
    short = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip() or k
    short = re.sub(r"\(anonymous namespace\)::", "", short).split("(")[0]
    r = regs.get(k)
    print(f"\n{short}\n  instructions {total[k]}" + (f", REG {r[0]}, static SHARED {r[1]}, LOCAL {r[2]} B" if r else ""))
    if c:
        print("  " + ", ".join(f"{w} {c[w]}" for w in WANT if c[w]))
