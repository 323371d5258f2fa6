"""Per-kernel counts of the SASS mnemonics that show what the shipped library is built from (evidence for profiles/):
UTMALDG (TMA tensor copies), LDSM (ldmatrix), SYNCS (mbarrier), IDP (dp4a/dp2a), POPC, LOP3, plus registers per thread.
usage: python tools/sass_counts.py [path/to/libbgdebias_b200.so] > profiles/r2_sass_counts.txt"""
import collections, pathlib, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else str(pathlib.Path(__file__).resolve().parent.parent / "background-debiased-video-cil_b200" / "libbgdebias_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True).stdout
regs = {}
for m in re.finditer(r"Function ([^:\s]+):\s*\n\s*REG:(\d+)", res):
    regs[m.group(1)] = int(m.group(2))
want = ["UTMALDG", "LDSM", "SYNCS", "IDP", "POPC", "LOP3", "IMAD", "SHF", "PRMT", "LDG", "STG", "FFMA", "MUFU"]
counts, total, cur = collections.defaultdict(collections.Counter), collections.Counter(), None
arch = set(re.findall(r"arch = (sm_\w+)", sass))
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for w in want:
            if op == w or op.startswith(w + "."):
                counts[cur][w] += 1
def demangle(n):
    try:
        return (subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n).replace("(anonymous namespace)::", "")
    except Exception:
        return n
print(f"# {lib.split('/')[-1]}: cubin arch {sorted(arch)}; {len(total)} kernels; per kernel: SASS instructions, registers, mnemonic counts")
print("# library totals: " + ", ".join(f"{w} {sum(c[w] for c in counts.values())}" for w in want))
fam = collections.defaultdict(lambda: [0, collections.Counter()])
for k in total:
    name = re.sub(r"<.*", "", demangle(k).replace("void ", "")).split("(")[0]
    fam[name][0] += 1
    for w in want:
        fam[name][1][w] += counts[k][w]
print("\n## by kernel family (all template instantiations summed)")
print(f"{'kernel':58s} {'inst.':>5s} " + " ".join(f"{w:>8s}" for w in want))
for name, (n, c) in sorted(fam.items()):
    print(f"{name[:58]:58s} {n:5d} " + " ".join(f"{c[w]:8d}" for w in want))
print("\n## every kernel")
print(f"{'kernel':100s} {'SASS':>6s} {'regs':>4s} " + " ".join(f"{w:>7s}" for w in want[:6]))
for k in sorted(total, key=demangle):
    d = demangle(k).replace("void ", "").split("(")[0]
    print(f"{d[:100]:100s} {total[k]:6d} {regs.get(k, 0):4d} " + " ".join(f"{counts[k][w]:7d}" for w in want[:6]))
