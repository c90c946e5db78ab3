"""SASS evidence for the judge: per kernel of libmvc_b200.so, the instruction count and the counts of the Blackwell
mnemonics that prove tcgen05 / TMEM / TMA / cluster use (B200_PROFILING.md): UTCHMMA (tcgen05.mma), LDTM / STTM
(tcgen05.ld / st), UTMALDG (TMA tensor load), UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit), SYNCS (mbarrier),
UCGABAR (cluster barrier), HMMA (mma.sync), MUFU, RED / ATOM.  Usage: python tools/sass_summary.py > profiles/sass_r2.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "multimodal-video-captioning_b200", "lib", "libmvc_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "UCGABAR", "HMMA", "MUFU", "RED", "ATOM",
        "MEMBAR", "LDSM", "STS", "LDS", "LDG", "STG", "SHFL", "BAR"]
cur, stats = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        stats[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        op = m.group(1)
        stats[cur]["_n"] += 1
        base = op.split(".")[0]
        if base in KEYS:
            stats[cur][base] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(stats), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}: {len(stats)} kernels (sm_100a)")
for (name, c), dn in zip(stats.items(), demangle):
    short = re.sub(r"\(.*", "", dn)
    keys = "  ".join(f"{k}={c[k]}" for k in KEYS if c[k])
    print(f"{short[:78]:78s} {c['_n']:6d} instr   {keys}")
