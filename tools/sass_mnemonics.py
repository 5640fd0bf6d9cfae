"""Dev tool: count the tensor-core / TMEM / TMA / packed-FP32 SASS instructions per kernel of a built object.
    python tools/sass_mnemonics.py > profiles/rNN_sass_mnemonics.txt
(the evidence B200_PROFILING.md asks for: UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor load,
UBLKCP = cp.async.bulk, SYNCS = mbarrier, FFMA2 / FMNMX3 = packed FP32 FMA / 3-input min)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = re.compile(r"^(UTCHMMA|UTCBAR|UTMALDG|UTMAPF|LDTM|STTM|UBLKCP|SYNCS|FFMA2|FMNMX3|UTCATOMSWS|TCGEN|UTCCP)")
for obj in ("gemm_tf32x3.o", "chamfer.o", "gcn_linear.o"):
    path = os.path.join(ROOT, "active-3d-vision-and-touch_b200", "build", obj)
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    fn, cnt = None, collections.defaultdict(collections.Counter)
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn and KEEP.match(m.group(1)):
            cnt[fn][m.group(1).split(".")[0]] += 1
    print(f"## {obj}")
    for f, c in cnt.items():
        print(f"  {f[:90]}: " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items())))
