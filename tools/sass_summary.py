"""profiles/<name>: opcode histogram of one render_wave instantiation in libdrt.so + the lines that show the
sm_100-specific instructions the kernel relies on (packed FFMA2 / FMNMX3 of the slab filter) and the absence of
tensor-core / TMA opcodes (the path is not a contraction).  usage: python tools/sass_summary.py [feat mask] > profiles/..."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
feat = int(sys.argv[1]) if len(sys.argv) > 1 else 24
sym = f"_ZN3drt11render_waveIdLi{feat}ELb0EEEvNS_6ParamsIT_EE"
so = os.path.join(ROOT, "distraytracer_b200", "libdrt.so")
out = subprocess.run(["cuobjdump", "-sass", "-fun", sym, so], capture_output=True, text=True).stdout
ins = [l for l in out.splitlines() if re.match(r"^\s+/\*[0-9a-f]{4,6}\*/", l)]
ops = collections.Counter()
for l in ins:
    m = re.match(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
    if m:
        ops[m.group(1)] += 1
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
usage = [l for l in res.splitlines() if sym in l or ("REG:" in l)]
print(f"# SASS of render_wave<double, {feat}, false> in distraytracer_b200/libdrt.so (sm_100a), {len(ins)} instructions")
k = [i for i, l in enumerate(res.splitlines()) if sym in l]
if k:
    print("# " + res.splitlines()[k[0] + 1].strip())
print("# opcode histogram")
for op, n in ops.most_common():
    print(f"{n:6d}  {op}")
fam = lambda p: sum(n for o, n in ops.items() if o.startswith(p))
print("# families: FP64 (D*) %d, FP32 (F*) %d of which packed FFMA2 %d / FMNMX3 %d, integer+logic (I*,L*,S*HF*) %d, MUFU %d, "
      "LDG/STG %d, LDS/STS %d, LDL/STL (local) %d, ATOMS/ATOMG/RED %d, SHFL/VOTE %d, BAR %d"
      % (fam("D"), fam("F"), ops.get("FFMA2", 0), ops.get("FMNMX3", 0), fam("I") + fam("L") + fam("SHF"), fam("MUFU"),
         fam("LDG") + fam("STG"), fam("LDS") + fam("STS"), fam("LDL") + fam("STL"), fam("ATOM") + fam("RED"), fam("SHFL") + fam("VOTE"), fam("BAR")))
tc = [o for o in ops if any(t in o for t in ("TCGEN", "UTMA", "UTC", "HMMA", "IMMA", "DMMA", "QGMMA", "UBLKCP", "SYNCS"))]
print("# tensor-core / TMA / mbarrier opcodes:", ", ".join(tc) if tc else "none (the path is not a contraction; records are popped by plain LDG.128 / STG.128)")
print("# first FFMA2 / FMNMX3 lines (the packed two-geoms-per-instruction slab filter, drt_kernels.cuh slabMask)")
shown = 0
for l in ins:
    if ("FFMA2" in l or "FMNMX3" in l) and shown < 12:
        print(l.rstrip()); shown += 1
