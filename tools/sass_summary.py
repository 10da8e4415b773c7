"""Per-kernel SASS mnemonic counts of the built library (what proves a Blackwell-native kernel, B200_PROFILING.md):
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA load / store, HMMA = legacy mma.sync, MUFU.EX2.
Usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt   (CPU only: cuobjdump -sass on the .so)"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "clip_diffusion_b200", "csrc", "libclipguide_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
PAT = collections.OrderedDict([("UTCHMMA", r"\bUTC[A-Z]*MMA"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
                               ("UTCBAR", r"\bUTCBAR"), ("HMMA", r"\bHMMA"), ("LDGSTS", r"\bLDGSTS"), ("MUFU.EX2", r"\bMUFU\.EX2"), ("ATOM/RED(f32)", r"\b(ATOMG|REDG|RED)\.E\.ADD\.F32")])
kern, counts, total = None, {}, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        kern = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", kern)
        kern = kern.split("(")[0][-90:]
        counts[kern] = collections.Counter()
        total[kern] = 0
        continue
    if kern and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        total[kern] += 1
        for k, p in PAT.items():
            if re.search(p, line):
                counts[kern][k] += 1
print("# SASS mnemonic counts per kernel of libclipguide_b200.so (sm_100a); blank = 0")
print("%-92s %6s " % ("kernel", "instr") + " ".join("%8s" % k[:8] for k in PAT))
for k in sorted(counts, key=lambda k: (-counts[k]["UTCHMMA"], -counts[k]["HMMA"], k)):
    c = counts[k]
    if not any(c.values()) and "--all" not in sys.argv:
        continue
    print("%-92s %6d " % (k, total[k]) + " ".join("%8s" % (c[p] if c[p] else "") for p in PAT))
print("# kernels with no tensor-core / TMA / TMEM / MUFU.EX2 / float-atomic instruction are omitted (--all lists them); total kernels: %d" % len(counts))
print("# float atomics anywhere: %d" % sum(c["ATOM/RED(f32)"] for c in counts.values()))
