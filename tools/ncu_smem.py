"""Per-source-line shared-memory wavefronts (actual vs ideal) of k_analyse<16> from an ncu report."""
import collections, csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 4096 * 468
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("FXLIB", os.path.join(ROOT, "feature-extractor_b200/lib/libfxb200.so"))], cwd=tmp, capture_output=True)
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
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
h = rows[1]
ai, wi, ii = h.index("Address"), h.index("L1 Wavefronts Shared"), h.index("L1 Wavefronts Shared Ideal")
data = rows[2:]
base = int(data[0][ai], 16)
w, wid = collections.Counter(), collections.Counter()
for r in data:
    k = a2l.get(int(r[ai], 16) - base)
    w[k] += int(r[wi] or 0); wid[k] += int(r[ii] or 0)
tw, ti = sum(w.values()), sum(wid.values())
print(f"wavefronts/frame {tw / frames:.0f} ideal {ti / frames:.0f}")
files = {"fx_analyse.cu": open(os.path.join(ROOT, "feature-extractor_b200/csrc/fx_analyse.cu")).read().split("\n"),
         "fx_fft.cuh": open(os.path.join(ROOT, "feature-extractor_b200/csrc/fx_fft.cuh")).read().split("\n")}
for k, v in w.most_common(int(os.environ.get("TOP", "30"))):
    f, ln = k if k else ("?", 0)
    t = files[f][ln - 1].strip()[:80] if f in files and 0 < ln <= len(files[f]) else ""
    print(f"{f[:13]:13s} {ln:4d} {v / frames:7.0f} wf/frame (ideal {wid[k] / frames:6.0f}) | {t}")
