"""Static and executed instruction accounting of the two fused 3-DOF kernels (main_3dof.py configuration, fp64):

    python tools/sass_summary.py [ncu source-page csv of the tensor-memory kernel] [warp-ticks] [same for the PID-phase kernel] [warp-ticks]

Static: registers / spill bytes from `nvcc -Xptxas -v`, SASS opcode counts from `cuobjdump -sass`.  Executed (optional):
opcode mix per warp-tick from `ncu -i rep --page source --csv --print-source sass` (Instructions Executed / warp-ticks)."""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "msc-thesis_b200", "csrc", "h3_loop_f64.cu")
KERNELS = (("closed_loop_tm_kernel<double, double, 6, 10, HYP=0, TAU1L=1, LANES=32> (ticks >= switch_step, tensor memory)",
            "closed_loop_tm_kernelIddLi6ELi10ELb0ELb1"),
           ("closed_loop_kernel<double, double, 6, 10, COL=0, SYM=1, HYP=0, TAU1L=1> (ticks < switch_step, PID collective)",
            "closed_loop_kernelIddLi6ELi10ELb0ELb1ELb0ELb1"))
OPS = ("DFMA", "DMUL", "DADD", "DSETP", "MUFU", "FSEL", "IMAD", "LDS", "STS", "LDL", "STL", "LDG", "STG", "LDTM", "STTM",
       "LDC", "LDCU", "BRA", "BSSY", "BSYNC", "WARPSYNC", "VOTE", "BAR")


def main():
    obj = "/tmp/h3_sass_summary.o"
    r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-I",
                        os.path.join(ROOT, "include"), "-DH3_FAST_BUILD", "-Xptxas", "-v", "-c", SRC, "-o", obj],
                       capture_output=True, text=True, check=True)
    ptxas = r.stderr
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    funcs = {f.split("\n", 1)[0].strip(): f for f in re.split(r"\n\s*Function : ", sass)[1:]}
    print("nvcc 12.9 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo; static counts are per kernel (every branch counted once)\n")
    for i, (title, key) in enumerate(KERNELS):
        name = [n for n in funcs if key in n][0]
        m = re.search(re.escape(name) + r".*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                      r"ptxas info\s+: Used (\d+) registers", ptxas, re.S)
        ins = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", funcs[name], re.M)
        c = collections.Counter(ins)
        print(title)
        print("  ptxas: %s registers, %s B stack frame, spill stores %s B, spill loads %s B" % (m.group(4), m.group(1), m.group(2), m.group(3)))
        print("  SASS instructions: %d (%.0f KB)" % (len(ins), len(ins) * 16 / 1024))
        print("  static: " + ", ".join("%s %d" % (o, c[o]) for o in OPS if c[o]))
        arg = 1 + 2 * i
        if len(sys.argv) > arg + 1:
            rows = list(csv.reader(open(sys.argv[arg])))
            hi = [j for j, rr in enumerate(rows) if rr and rr[0] == "Address"][0]
            h = rows[hi]
            isrc, iex, ith = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
            wt = float(sys.argv[arg + 1])
            ex, th = collections.Counter(), 0.0
            for rr in rows[hi + 1:]:
                mm = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", rr[isrc])
                if mm and rr[iex]:
                    ex[mm.group(1)] += float(rr[iex]) / wt
                    th += float(rr[ith]) / wt
            tot = sum(ex.values())
            print("  executed per warp-tick (ncu, %g warp-ticks): %.0f instructions, %.1f active lanes per instruction" % (wt, tot, th / tot))
            print("    " + ", ".join("%s %.0f" % (o, ex[o]) for o in OPS if ex[o] >= 0.5))
            fp64 = ex["DFMA"] + ex["DMUL"] + ex["DADD"] + ex["DSETP"]
            print("    FP64-pipe instructions %.0f (%.0f %%); per env-step: %.1f warp-instructions" % (fp64, 100 * fp64 / tot, tot / 32))
        print()


if __name__ == "__main__":
    main()
