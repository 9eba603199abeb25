"""Static evidence that the kernels are Blackwell-native, produced on the build machine (no GPU needed): per object file
the count of tcgen05 MMA (UTC*MMA), TMEM load (LDTM), TMA (UTMALDG / UTMASTG / UBLKCP), packed fp32 (FFMA2), cp.async
(LDGSTS) and legacy mma.sync (HMMA) SASS instructions, and per kernel the registers / spills ptxas reports.
    python tools/sass_evidence.py > profiles/r1_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from segmif_b200 import build  # noqa: E402

PAT = [("UTC*MMA (tcgen05.mma)", r"\bUTC[A-Z]*MMA\b"), ("LDTM (tcgen05.ld)", r"\bLDTM\b"), ("UTMALDG (TMA load)", r"\bUTMALDG\b"),
       ("UTMASTG (TMA store)", r"\bUTMASTG\b"), ("FFMA2 (packed fp32)", r"\bFFMA2\b"), ("LDGSTS (cp.async)", r"\bLDGSTS\b"),
       ("HMMA (mma.sync)", r"\bHMMA\b")]


def main():
    build.build()
    objs = sorted(f for f in os.listdir(build.OBJ) if f.endswith(".o"))
    print("SASS instruction counts per object (cuobjdump -sass, sm_100a)\n")
    print(f"{'object':22s}" + "".join(f"{n.split(' ')[0]:>10s}" for n, _ in PAT))
    for o in objs:
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(build.OBJ, o)], capture_output=True, text=True).stdout
        print(f"{o:22s}" + "".join(f"{len(re.findall(p, sass)):10d}" for _, p in PAT))
    print("\nlegend: " + "; ".join(n for n, _ in PAT))
    print("\nptxas resource usage per kernel (nvcc -Xptxas -v): registers, spill bytes, static shared memory\n")
    for src in sorted(f for f in os.listdir(build.CSRC) if f.endswith(".cu")):
        r = subprocess.run([build._nvcc()] + build.NVCC_FLAGS + ["-Xptxas", "-v", "-c", os.path.join(build.CSRC, src), "-o", os.devnull],
                           capture_output=True, text=True)
        cur, rows = None, collections.OrderedDict()
        for line in r.stderr.splitlines():
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                cur = re.sub(r"\(.*", "", name).replace("void ", "")
                rows[cur] = ["?", "0", "0"]
            m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and cur:
                rows[cur][1] = m.group(1)
            m = re.search(r"Used (\d+) registers", line)
            if m and cur:
                rows[cur][0] = m.group(1)
                sm = re.search(r"(\d+) bytes smem", line)
                rows[cur][2] = sm.group(1) if sm else "0"
        print(f"{src}")
        for k, (regs, spill, smem) in rows.items():
            print(f"   {k[:86]:86s} regs {regs:>3s}  spill {spill:>4s}  smem {smem:>6s}")


if __name__ == "__main__":
    main()
