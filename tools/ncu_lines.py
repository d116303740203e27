"""Per-source-line executed-instruction and stall-sample breakdown of k_analyse<16> from an ncu report.
Usage: ncu_lines.py rep.ncu-rep [n_frames]   (needs the matching lib/libfxb200.so built with -lineinfo)"""
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
src_csv = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src_csv)))
h = rows[1]
ai, si, ie = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed")
data = rows[2:]
base = int(data[0][ai], 16)
ex, sm = collections.Counter(), collections.Counter()
for r in data:
    k = a2l.get(int(r[ai], 16) - base)
    ex[k] += int(r[ie] or 0)
    sm[k] += int(r[si] or 0)
te, ts = sum(ex.values()), sum(sm.values())
files = {"fx_analyse.cu": open(os.path.join(ROOT, "feature-extractor_b200/csrc/fx_analyse.cu")).read().split("\n"),
         "fx_fft.cuh": open(os.path.join(ROOT, "feature-extractor_b200/csrc/fx_fft.cuh")).read().split("\n")}
print(f"instr/thread/frame total {te / frames / 8:.0f}; static {len(data)}")
byfile = collections.Counter()
for k, v in ex.items():
    byfile[k[0] if k else None] += v
print({k: round(v / frames / 8) for k, v in byfile.items()})
for k, v in ex.most_common(int(os.environ.get("TOP", "70"))):
    f, ln = k
    t = files[f][ln - 1].strip()[:100] if f in files and ln - 1 < len(files[f]) else ""
    print(f"{f[:13]:13s} {ln:4d} {v / frames / 8:7.1f} i/thr/frame  smp {sm[k] / ts:.3f} | {t}")
