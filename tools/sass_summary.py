"""Mnemonic counts of the shipped sm_100a kernels: `python tools/sass_summary.py > profiles/r2_sass_summary.md` (CPU only: cuobjdump)."""
import collections, glob, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
objs = sorted(glob.glob(os.path.join(ROOT, "comfyui-keep_b200", "build", "*.o")))
KEY = ["UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "UTCATOMSWS", "SYNCS", "LDG.E.ENL2.256", "STG.E.ENL2.256", "LDL", "STL"]
per_fn, totals = {}, collections.Counter()
for o in objs:
    txt = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
    fn = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"keep::\(anonymous namespace\)::", "", fn).split("(")[0]
            per_fn[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and fn:
            op = m.group(1)
            per_fn[fn][op] += 1
            for k in KEY:
                if op.startswith(k):
                    totals[k] += 1
print("# SASS of the shipped sm_100a kernels (final build of round 2; `cuobjdump -sass comfyui-keep_b200/build/*.o`, nvcc 12.9,")
print("`-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo`; regenerate with `python tools/sass_summary.py`)\n")
print("Mnemonics that prove the Blackwell path (B200_PROFILING.md): `UTCHMMA` = tcgen05.mma (kind::f16), `LDTM` = tcgen05.ld (TMEM ->")
print("registers), `UTCBAR` = tcgen05.commit -> mbarrier, `UBLKCP` = cp.async.bulk (1-D TMA: weight panels, packed K / V stages),")
print("`SYNCS` = mbarrier try_wait / arrive, `UTCATOMSWS` = TMEM allocation, `LDG/STG.E.ENL2.256` = 256-bit global accesses; `LDL` / `STL` =")
print("local memory (spills).\n")
print("## Whole library\n\n| mnemonic (prefix) | instructions |\n|---|---|")
for k in KEY:
    print("| `%s` | %d |" % (k, totals[k]))
print("\n## Per kernel (tensor-core kernels and their helpers)\n")
print("| kernel | instructions | " + " | ".join("`%s`" % k for k in KEY) + " |")
print("|---|---|" + "---|" * len(KEY))
for fn in sorted(per_fn):
    c = per_fn[fn]
    if not any(s in fn for s in ("conv_tc_kernel", "attn_tc_kernel", "attn_pack_kv", "splitk_reduce", "tc_pack", "tc_repack", "conv_cout4", "conv_cin3_px4")):
        continue
    row = [sum(v for op, v in c.items() if op.startswith(k)) for k in KEY]
    print("| `%s` | %d | " % (fn, sum(c.values())) + " | ".join(str(v) for v in row) + " |")
