"""SASS mnemonic counts per kernel of the in-tree sm_100a objects (quasiparticle-physics-simulation_b200/lib/obj/*.o):
python profiles/sass_summary.py > profiles/r2_sass_mnemonics.txt
UTMALDG / UTMASTG = TMA tensor loads / stores, SYNCS = mbarrier, LDGSTS = cp.async, DMMA = FP64 tensor-core MMA, REDUX =
warp reductions, UCGABAR = cluster barrier, STAS = st.async into another CTA's shared memory (distributed shared memory),
CCTL = cache control (L1 prefetch / invalidate).  Only the kernels the bench, smoke and test paths launch are listed."""
import collections, glob, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = re.compile(r"^(UTMALDG|UTMASTG|UTMAPF|SYNCS|LDGSTS|DMMA|REDUX|UCGABAR|STAS|ATOMS|BAR|MUFU|SHFL|LDS|STS|DFMA|DMUL|DADD|CCTL|MEMBAR|LDL|STL)")
WANT = re.compile(r"k_collide_struct<\(int\)(32|16), \(int\)512|k_collide_gemm|k_sweep_[xy]_pipe|k_sweep_x_tma|k_thomas|k_dct|k_kry_matvec|k_kry_line|"
                  r"k_pr_resident|k_sweep_[xy]_vard|k_generation_program|k_chunk_products|k_factor_vard")
print(__doc__.strip().replace("\n", "\n# ").join(["# ", ""]))
for obj in sorted(glob.glob(os.path.join(ROOT, "quasiparticle-physics-simulation_b200", "lib", "obj", "*.o"))):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True, text=True).stdout.split("\n")
    blocks = re.split(r"\n\s*Function : \S+\n", txt)[1:]
    seen = collections.OrderedDict()
    for name, body in zip(names, blocks):
        name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        if not WANT.search(name):
            continue
        cnt = collections.Counter()
        for m in re.finditer(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body, re.M):
            op = m.group(1)
            if KEEP.match(op):
                key = op if op.startswith(("UTMA", "SYNCS", "LDGSTS", "DMMA", "UCGABAR", "STAS", "CCTL", "MEMBAR")) else op.split(".")[0]
                cnt[key] += 1
        family = re.sub(r"<.*", "", name)
        seen.setdefault(family, []).append((name, cnt))
    PREFER = ("k_pr_resident<(int)32, (int)16, (int)256>", "k_collide_struct<(int)32, (int)512, (bool)1, (bool)1, (bool)1>",
              "k_sweep_x_pipe<(int)16, (int)16, (int)2, (int)256, (bool)0, (bool)1>",
              "k_sweep_y_pipe<(int)16, (int)16, (int)2, (int)256, (bool)0, (bool)1>", "k_sweep_x_vard<(int)16, (int)16>")
    for family, items in seen.items():
        pick = [it for it in items if any(p in it[0] for p in PREFER)]
        name, cnt = (pick or items)[0]
        extra = f"    [{len(items)} template instances, this one shown]" if len(items) > 1 else ""
        print(f"\n{os.path.basename(obj)[:-2]}: {name}{extra}")
        print("    " + "  ".join(f"{k}={v}" for k, v in sorted(cnt.items())))
