"""SASS evidence for the hot kernels (run here, no GPU needed): python profiles/sass_excerpts.py
Writes profiles/sass_<kernel>.txt: per kernel the instruction histogram of the mnemonics that prove the hardware path
(UTCIMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, SYNCS = mbarrier, UBLKCP = cp.async.bulk, VIMNMX3 ...)
and the first lines that contain each of them."""
import collections
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "rgbdslam_v2_b200" / "librgbdslam_b200.so"
KERNELS = {
    "tc_hamming_expand_kernel": ["UTCIMMA", "LDTM", "UTCBAR", "SYNCS", "VIMNMX3", "LOP3", "IMAD.SHL", "STS.128", "FENCE", "LDG.E.128"],
    "tc_match256_kernel": ["UTCHMMA", "UTCIMMA", "LDTM", "UTCBAR", "UBLKCP", "SYNCS"],
    "ransac_hyp_kernel": ["MUFU", "DFMA", "SHFL", "VOTE", "FFMA"],
    "k_fast_nms": ["VIMNMX3", "VIMNMX", "LDS", "STS"],
    "pg_pcg_resident_kernel": ["DFMA", "LD.E.64.STRONG.GPU", "ST.E.64.STRONG.GPU", "MEMBAR", "FENCE", "LDS", "SHFL", "BAR.SYNC"],
    "ba_cam_apply_kernel": ["DFMA", "SHFL", "LDG"],
}
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", sass)
for name, wanted in KERNELS.items():
    out = []
    for b in blocks[1:]:
        head = b.split("\n", 1)[0]
        if name not in head:
            continue
        lines = [l for l in b.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]
        ops = collections.Counter()
        for l in lines:
            m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
            if m:
                ops[m.group(1)] += 1
        out.append(f"== {head.strip()}\n   {len(lines)} SASS instructions; opcode histogram (static counts), top 24:")
        out.append("   " + ", ".join(f"{k} {v}" for k, v in ops.most_common(24)))
        for w in wanted:
            hits = [l.strip() for l in lines if w in l]
            out.append(f"   -- {w}: {len(hits)} occurrence(s)")
            for h in hits[:3]:
                out.append("      " + re.sub(r"\s+", " ", h)[:150])
    (ROOT / "profiles" / f"sass_{name}.txt").write_text("\n".join(out) + "\n")
    print(name, "->", len(out), "lines")
