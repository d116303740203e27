"""Per-source-line stall-reason breakdown of k_analyse<16> from an ncu report (companion of ncu_lines.py).
Usage: ncu_stalls.py rep.ncu-rep   (needs the matching lib/libfxb200.so built with -lineinfo)"""
import collections, csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
lib = os.environ.get("FXLIB", os.path.join(ROOT, "feature-extractor_b200/lib/libfxb200.so"))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, "fx_analyse.sm_100a.cubin")], capture_output=True, text=True).stdout
cur, infunc, a2l = None, False, {}
for l in dis.split("\n"):
    if l.startswith(".text."):
        infunc = "k_analyseILi16E" in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+\S", l)
    if infunc and m:
        a2l[int(m.group(1), 16)] = cur
src_csv = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src_csv)))
h = rows[1]
ai = h.index("Address")
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
idx = {c: h.index(c) for c in stalls}
data = rows[2:]
base = int(data[0][ai], 16)
tot = collections.Counter()
byline = collections.defaultdict(collections.Counter)
for r in data:
    k = a2l.get(int(r[ai], 16) - base)
    for c in stalls:
        v = int(r[idx[c]] or 0)
        tot[c] += v
        byline[c][k] += v
T = sum(tot.values())
print("total samples", T)
for c, v in tot.most_common():
    print(f"{c:24s} {v / T:.3f}")
files = {"fx_analyse.cu": open(os.path.join(ROOT, "feature-extractor_b200/csrc/fx_analyse.cu")).read().split("\n"),
         "fx_fft.cuh": open(os.path.join(ROOT, "feature-extractor_b200/csrc/fx_fft.cuh")).read().split("\n")}
for c in os.environ.get("WHICH", "stall_barrier,stall_no_inst,stall_short_sb,stall_wait,stall_mio").split(","):
    print("==", c)
    for k, v in byline[c].most_common(int(os.environ.get("TOP", "12"))):
        f, ln = k if k else ("?", 0)
        t = files[f][ln - 1].strip()[:90] if f in files and 0 < ln <= len(files[f]) else ""
        print(f"  {v / T:.4f} {f[:13]:13s} {ln:4d} | {t}")
