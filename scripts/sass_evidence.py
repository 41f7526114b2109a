"""Counts of the Blackwell-native SASS mnemonics per kernel of libviquae_b200.so (cuobjdump -sass).
Usage: python scripts/sass_evidence.py > profiles/sass_tc_topk.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "viquae_b200", "libviquae_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
PAT = re.compile(r"\b(UTCHMMA(?:\.2CTA)?|UTCBAR(?:\.2CTA)?(?:\.MULTICAST)?|UTMALDG\.2D(?:\.2CTA)?|LDTM(?:\.x\d+)?|UTCATOMSWS[.\w]*|SYNCS\.PHASECHK[.\w]*|SYNCS\.ARRIVE[.\w]*|ELECT|HMMA[.\w]*|FFMA2|FFMA|FMNMX3?|UCGABAR_\w+|LDG\.E[.\w]*|STG\.E[.\w]*|R2UR(?:\.BROADCAST)?)\b")
fn, counts, arch = None, collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[fn] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch = m.group(1)
    if fn and "/*" in line:
        body = line.split("*/", 1)[-1]
        for mm in PAT.findall(body):
            counts[fn][mm] += 1
print(f"# cuobjdump -sass viquae_b200/libviquae_b200.so ({arch}); mnemonic counts per kernel")
print("# tcgen05.mma -> UTCHMMA, tcgen05.commit -> UTCBAR, tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG, mbarrier -> SYNCS.*, fma.rn.f32x2 -> FFMA2")
for fn, c in counts.items():
    if not c:
        continue
    keys = sorted(c, key=lambda k: (not k.startswith(("UTC", "UTMA", "LDTM")), k))
    print(f"{fn}\n    " + "  ".join(f"{k}={c[k]}" for k in keys))
