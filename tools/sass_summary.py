"""profiles/sass_summary.txt: per kernel of libmjv_b200.so, how many tcgen05 / TMEM / TMA instructions its SASS holds
(cuobjdump -sass; mnemonics per /opt/skills/guides/B200_PROFILING.md: tcgen05.mma = UTC*MMA, tcgen05.ld/st = LDTM/STTM,
TMA = UTMALDG/UTMASTG, tcgen05.commit = UTCBAR, setmaxnreg = USETMAXREG)."""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "mj_video_b200/libmjv_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "USETMAXREG", "MUFU.EX2", "HMMA", "SYNCS"]
name, counts, total = None, collections.OrderedDict(), {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)[:100]
        counts[name] = collections.Counter()
        total[name] = 0
        continue
    if name and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
        total[name] += 1
        for p in pats:
            if re.search(r"\b" + re.escape(p) + r"\b" if "." not in p else re.escape(p), line):
                counts[name][p] += 1
arch = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
lines = [f"# SASS summary of {so}", f"# cubins: {sorted(set(re.findall(r'sm_[0-9a-z]+', arch)))}", ""]
hdr = f"{'kernel':100s} {'instr':>6s} " + " ".join(f"{p:>12s}" for p in pats)
lines.append(hdr)
for k, c in counts.items():
    if total[k] == 0:
        continue
    lines.append(f"{k:100s} {total[k]:6d} " + " ".join(f"{c.get(p, 0):12d}" for p in pats))
open("profiles/sass_summary.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:6]))
print(len(counts), "kernels")
