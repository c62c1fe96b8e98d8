"""Instruction mix of the hot kernels, read from the SASS of the built library (no GPU needed):
    python tools/sass_summary.py > profiles/sass_hot_kernels_<round>.txt
For every kernel of interest: SASS instruction count, the opcode histogram of the memory / reduction / special-function /
double-precision / synchronisation instructions, registers and shared memory from the ptxas logs next to the sources.
The mnemonics to look for on sm_100a are in /opt/skills/guides/B200_PROFILING.md (UBLKCP = cp.async.bulk, SYNCS = mbarrier,
LDG.E.ENL2.256 = 256-bit gather, REDG.E.ADD.F32x4 = red.global.add.v4.f32)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dem-engine_b200", "libdemcore.so")
WANT = [  # demangled prefix, what it is
    ("void demb::k_force_ss<0, false, 4, true>", "sphere-sphere force, Hertz-Mindlin, no record, 4 CTAs/SM, MUFU arithmetic (the bench kernel)"),
    ("void demb::k_force_ss_due<0, false, 4, true>", "the same for long-lived lists (candidates skipped until due)"),
    ("void demb::k_integrate<0>", "integration + wrench reset + max|v|"),
    ("demb::k_sweep_tma", "contact-pair sweep, count pass (TMA-staged)"),
    ("demb::k_sweep_fill", "contact-pair sweep, fill pass"),
    ("demb::k_history", "history carry-over"),
    ("demb::k_sphere_prep", "sphere -> cell keys, sphere-analytical candidates"),
    ("demb::k_scan_lookback", "single-pass scan"),
    ("demb::k_mg_exchange", "multi-GPU halo exchange over peer memory"),
    ("demb::k_reduce_spheres", "sphere-level inspector reduction"),
]
KEEP = re.compile(r"^(LDG|STG|LDS|STS|LDC|LDL|STL|REDG|RED|ATOMG|ATOMS|ATOM|UBLKCP|UTMALDG|UTMASTG|SYNCS|MUFU|DFMA|DADD|DMUL|F2F|I2F|F2I|"
                  r"SHFL|BAR|MEMBAR|FENCE|ERRBAR|CCTL|VOTE|MATCH|REDUX|WARPSYNC|ELECT|UCGABAR|LDGSTS|LDSM)")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    blocks = re.split(r"\n\s+Function : ", sass)[1:]
    names = subprocess.run(["c++filt"], input="\n".join(b.split("\n", 1)[0].strip() for b in blocks), stdout=subprocess.PIPE,
                           text=True).stdout.split("\n")
    regs = {}
    csrc = os.path.join(ROOT, "dem-engine_b200", "csrc")
    for log in os.listdir(csrc):
        if not log.endswith(".ptxas.log"):
            continue
        cur = None
        for line in open(os.path.join(csrc, log)):
            m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
            if m:
                cur = m.group(1)
                regs[cur] = {}
                continue
            if cur is None:
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m:
                regs[cur].update(stack=int(m.group(1)), spill=int(m.group(2)) + int(m.group(3)))
            m = re.search(r"Used (\d+) registers", line)
            if m:
                s = re.search(r"(\d+) bytes smem", line)
                regs[cur].update(regs=int(m.group(1)), smem=int(s.group(1)) if s else 0)
    print("SASS summary of %s (nvcc -gencode arch=compute_100a,code=sm_100a; cuobjdump -sass)\n" % os.path.relpath(LIB, ROOT))
    for prefix, what in WANT:
        for blk, name in zip(blocks, names):
            if not name.startswith(prefix):
                continue
            mangled = blk.split("\n", 1)[0].strip()
            ops = collections.Counter()
            total = 0
            for line in blk.split("\n"):
                m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
                if m:
                    total += 1
                    ops[m.group(1)] += 1
            r = regs.get(mangled, {})
            print("== %s\n   %s" % (name.split("(")[0], what))
            extra = ""
            if r:
                extra = "; %d registers, %d B static shared memory, %d B stack, %d B spills" % (
                    r.get("regs", -1), r.get("smem", 0), r.get("stack", 0), r.get("spill", 0))
            print("   %d SASS instructions%s" % (total, extra))
            shown = sorted(((k, v) for k, v in ops.items() if KEEP.match(k)), key=lambda kv: (-kv[1], kv[0]))
            print("   " + ", ".join("%s x%d" % kv for kv in shown))
            fp32 = sum(v for k, v in ops.items() if re.match(r"^(FFMA|FADD|FMUL|FSEL|FSETP|FMNMX|FCHK)", k))
            intg = sum(v for k, v in ops.items() if re.match(r"^(IADD3|IADD|IMAD|LOP3|SHF|LEA|ISETP|SEL|PRMT|IABS|POPC|FLO|BREV|UIADD3|ULOP3|UIMAD|USHF|ULEA)", k))
            print("   fp32 arithmetic %d, integer / address %d, other %d\n" % (fp32, intg, total - fp32 - intg - sum(v for _, v in shown)))
            break
        else:
            print("== %s: not found in the library\n" % prefix)


if __name__ == "__main__":
    sys.exit(main())
