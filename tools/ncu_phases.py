"""Per-PHASE breakdown of k_analyse<R1> from an ncu report captured with --import-source on: executed warp-instructions
per thread and frame, share of issue slots, warp-stall samples by reason, shared-memory wavefronts (actual / ideal ->
bank-conflict replays) for every phase of the frame loop, plus the source lines that own the conflict wavefronts.

Every SASS instruction is attributed to the OUTERMOST source line of its inline chain (nvdisasm -gi), i.e. the line of
k_analyse that the work was written on; the lines map to phases through the `// ====` section markers of fx_analyse.cu.
fft_core is one out-of-line function called twice per frame: it is reported as its own phase, split by stage.

Usage: ncu_phases.py rep.ncu-rep [n_frames] [R1] [MG]    (needs the matching lib/libfxb200.so built with -lineinfo;
                                                        FXLIB selects another build)"""
import collections, csv, io, json, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 4096 * 468
R1 = int(sys.argv[3]) if len(sys.argv) > 3 else 16
MG = int(sys.argv[4]) if len(sys.argv) > 4 else 0          # the instantiation with (1) / without (0) the decision margins
THREADS = 16 * R1
lib = os.environ.get("FXLIB", os.path.join(ROOT, "feature-extractor_b200/lib/libfxb200.so"))
src_path = os.path.join(ROOT, "feature-extractor_b200/csrc/fx_analyse.cu")
fft_path = os.path.join(ROOT, "feature-extractor_b200/csrc/fx_fft.cuh")
src = open(src_path).read().split("\n")
fft = open(fft_path).read().split("\n")


def find(lines, needle, start=0):
    for i in range(start, len(lines)):
        if needle in lines[i]:
            return i + 1
    raise SystemExit(f"marker {needle!r} not found")


# phase boundaries from the section markers of the frame loop (1-based line numbers, half-open ranges)
loop = find(src, "for (int f = f_begin; f < f_end; ++f)")
marks = [
    ("prologue + per-frame rematerialised invariants", find(src, "k_analyse (const AnalyseParams p)")),
    ("frame head (mbarrier wait)", loop),
    ("filter + window + RMS", find(src, "one-pole filter + window -> work array")),
    ("FFT-alpha gather", find(src, "FFT-alpha: z = x w + i onepole")),
    ("split + spectral pass 1", find(src, "previous non-silent spectrum of this thread's bins")),
    ("combine + flatness range events", find(src, "every thread needs the magnitude sum")),
    ("FFT-beta gather", find(src, "FFT-beta: z = x + i 2^k P")),
    ("flatness prefetch + hop prefetch", find(src, "The flatness product's continuation (record stage, below) starts")),
    ("pitch (cnd scan, lag search)", find(src, "pitch: cumulative normalised difference + lag search")),
    ("harmonic (peaks, inharmonicity)", find(src, "The harmonic and sub-octave bins of f0")),
    ("record stage", find(src, "---- what is left of the frame's record")),
    ("chunk epilogue", find(src, "---- chunk epilogue")),
    ("(end)", find(src, "K1b: the scalar tail of both analyser bodies")),
]
fft_marks = [
    ("fft stage 1 (radix-R1 + twiddle + store)", find(fft, "void fft_stage1_store")),
    ("fft stage 2 (radix-16 + twiddle)", find(fft, "void fft_stage2")),
    ("fft stage 3 (radix-16)", find(fft, "void fft_stage3")),
    ("(end)", find(fft, "position of spectrum bin k")),
]


def phase_of(fname, line, in_fft_core):
    if in_fft_core:
        # butterflies and packed arithmetic are inlined into the stage functions: outermost line inside fft_core is in fx_analyse.cu
        return None
    if fname != "fx_analyse.cu":
        return "other"
    for (name, a), (_, b) in zip(marks, marks[1:]):
        if a <= line < b:
            return name
    return "helpers / other"


tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, "fx_analyse.sm_100a.cubin")], capture_output=True, text=True).stdout

# address -> (outermost file, line, innermost file, line, in_fft_core, opcode)
kern = f"k_analyseILi{R1}ELb{MG}E"
a2 = {}
infunc = in_core = False
pending, cur = [], None                   # annotations since the last instruction: innermost first, outermost last
for l in dis.split("\n"):
    if l.startswith(".text."):
        infunc = kern in l
        in_core = False
        continue
    if not infunc:
        continue
    if l.lstrip().startswith(".type") and "fft_core" in l:
        in_core = True
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        pending.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(\S.*?);", l)
    if m:
        if pending:
            # FFT stage: the line info nests stages 2 / 3 under stage 1's last store, so any chain entry inside the body of
            # fft_stage2 / fft_stage3 decides; a chain that only touches fft_stage1_store (or the butterflies it inlines) is stage 1
            stage = None
            for fn, ln in pending:
                if fn == "fx_fft.cuh":
                    for (name, a_), (_, b_) in zip(fft_marks[1:], fft_marks[2:]):
                        if a_ <= ln < b_:
                            stage = name
                    if stage:
                        break
            if stage is None and any(fn == "fx_fft.cuh" for fn, _ in pending):
                stage = fft_marks[0][0]
            cur = (pending[-1], pending[0], stage)
            pending = []
        if cur:
            toks = m.group(2).split()
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            a2[int(m.group(1), 16)] = (cur[0][0], cur[0][1], cur[1][0], cur[1][1], in_core, op, cur[2])


rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
h = rows[1]
col = {c: i for i, c in enumerate(h)}
stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
data = rows[2:]
base = int(data[0][col["Address"]], 16)

P = collections.defaultdict(lambda: collections.Counter())
conflict_lines = collections.Counter()
conflict_ideal = collections.Counter()
for r in data:
    addr = int(r[col["Address"]], 16) - base
    info = a2.get(addr)
    if info is None:
        ph = "unattributed"
    else:
        of, ol, inf, inl, core, op, stage = info
        if core:
            ph = stage or "fft_core call overhead / barriers"
        else:
            ph = phase_of(of, ol, False)
    c = P[ph]
    ie = int(r[col["Instructions Executed"]] or 0)
    c["inst"] += ie
    c["samples"] += int(r[col["# Samples"]] or 0)
    c["wf"] += int(r[col["L1 Wavefronts Shared"]] or 0)
    c["wf_ideal"] += int(r[col["L1 Wavefronts Shared Ideal"]] or 0)
    for s in stall_cols:
        c[s] += int(r[col[s]] or 0)
    ex = int(r[col["L1 Wavefronts Shared"]] or 0) - int(r[col["L1 Wavefronts Shared Ideal"]] or 0)
    if ex > 0 and info is not None:
        key = (info[0], info[1], info[2], info[3], info[5])
        conflict_lines[key] += ex
        conflict_ideal[key] += int(r[col["L1 Wavefronts Shared Ideal"]] or 0)

tot_inst = sum(c["inst"] for c in P.values())
tot_samp = sum(c["samples"] for c in P.values())
tot_wf = sum(c["wf"] for c in P.values())
tot_ideal = sum(c["wf_ideal"] for c in P.values())
order = [m[0] for m in marks[:-1]] + ["fft stage 1 (radix-R1 + twiddle + store)", "fft stage 2 (radix-16 + twiddle)", "fft stage 3 (radix-16)",
                                      "fft_core call overhead / barriers", "fft_core (both transforms)", "helpers / other", "other", "unattributed"]
warps_per_frame = THREADS / 32.0
print(f"# k_analyse<{R1}> per-phase breakdown ({os.path.basename(rep)}; {frames} frames)")
print(f"# total: {tot_inst / frames / warps_per_frame:.0f} warp-instructions per thread per frame, {tot_samp} stall samples, "
      f"{tot_wf / frames:.0f} shared wavefronts/frame ({(tot_wf - tot_ideal) / frames:.0f} = {100.0 * (tot_wf - tot_ideal) / max(tot_wf, 1):.1f} % conflict replays)")
print(f"{'phase':44s} {'inst/thr/frame':>14s} {'inst %':>7s} {'samples %':>9s} {'wf/frame':>9s} {'replays':>8s}  top stalls (share of the phase's samples)")
out = []
for name in order:
    if name not in P:
        continue
    c = P[name]
    st = sorted(((c[s], s) for s in stall_cols), reverse=True)[:3]
    tops = ", ".join(f"{s[6:]} {v / max(c['samples'], 1):.2f}" for v, s in st if v)
    ipf = c["inst"] / frames / warps_per_frame
    print(f"{name:44s} {ipf:14.1f} {100.0 * c['inst'] / tot_inst:7.1f} {100.0 * c['samples'] / max(tot_samp, 1):9.1f} {c['wf'] / frames:9.0f} {(c['wf'] - c['wf_ideal']) / frames:8.0f}  {tops}")
    out.append({"phase": name, "inst_per_thread_frame": ipf, "inst_share": c["inst"] / tot_inst, "sample_share": c["samples"] / max(tot_samp, 1),
                "wavefronts_per_frame": c["wf"] / frames, "conflict_replays_per_frame": (c["wf"] - c["wf_ideal"]) / frames,
                "top_stalls": {s[6:]: v / max(c["samples"], 1) for v, s in st if v}})
print()
print("# shared-memory bank-conflict replays by source line (outermost line <- innermost line, SASS op), wavefronts per frame")
for (of, ol, inf, inl, op), v in conflict_lines.most_common(int(os.environ.get("TOP", "16"))):
    text = (src[ol - 1] if of == "fx_analyse.cu" else fft[ol - 1] if of == "fx_fft.cuh" else "").strip()[:100]
    print(f"{v / frames:8.1f} (+ideal {conflict_ideal[(of, ol, inf, inl, op)] / frames:7.1f})  {of}:{ol} <- {inf}:{inl} {op:10s} | {text}")
if os.environ.get("JSON"):
    json.dump({"kernel": f"k_analyse<{R1}>", "frames": frames, "phases": out}, open(os.environ["JSON"], "w"), indent=1)
