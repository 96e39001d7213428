"""ptxas resource usage (registers, stack, spills, static shared memory) of every kernel in libsixdgs.so, as a markdown
table: `python tools/ptxas_report.py > profiles/ptxas_r2.md`.  Runs on the CPU-only build box (nvcc cross-compiles)."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "6dgs_b200", "csrc")
sys.path.insert(0, CSRC)
import build as b  # noqa: E402

PAT = re.compile(r"Compiling entry function '([^']+)' for 'sm_100a'\n(?:.*\n)*?.*?(\d+) bytes stack frame, (\d+) bytes spill "
                 r"stores, (\d+) bytes spill loads\n.*Used (\d+) registers(.*)")


def main():
    rows = []
    with tempfile.TemporaryDirectory() as tmp:
        procs = []
        for src, extra in b.SOURCES.items():
            cmd = [b._nvcc(), *b.ARCH, *b.COMMON, *extra, "-Xptxas=-v", "-c", os.path.join(CSRC, src), "-o",
                   os.path.join(tmp, src + ".o")]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        for src, p in procs:
            out, _ = p.communicate()
            for m in PAT.finditer(out):
                name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                name = re.sub(r"\(anonymous namespace\)::", "", name).replace("sixdgs::", "")
                name = re.sub(r"\(.*", "", name).replace("void ", "")
                smem = re.search(r"(\d+) bytes smem", m.group(6))
                rows.append((src, name, m.group(5), m.group(2), m.group(3), m.group(4), smem.group(1) if smem else "0"))
    print("# ptxas resource usage per kernel (sm_100a, the flags of csrc/build.py)\n")
    print("`python tools/ptxas_report.py`.  Dynamic shared memory (the TMA rings of the tensor-core kernels, up to 230.7 KB) is")
    print("requested at launch and does not show here.  The measured kernels (`score_tc_mq_kernel<pass, format>`) have no")
    print("stack frame and no spills; 168 registers x 384 threads (pass 2) and 105 x 384 (pass 1) fit the 64 K register file")
    print("of an SM with the one resident CTA their shared memory allows.\n")
    print("| file | kernel | registers | stack B | spill st B | spill ld B | static smem B |")
    print("|---|---|---|---|---|---|---|")
    for r in rows:
        print("| " + " | ".join((r[0], f"`{r[1]}`") + r[2:]) + " |")


if __name__ == "__main__":
    main()
