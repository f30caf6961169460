#!/usr/bin/env python
"""Blackwell evidence of the shipped library, regenerated from the build (no GPU needed):

    python tools/sass_evidence.py            # writes profiles/r2_sass_opcodes.txt

Per kernel of sgnerf_b200/libsgnerf_b200.so (cuobjdump -sass, sm_100a): counts of the tensor-core / TMEM / TMA / cluster mnemonics
(UTCHMMA[.2CTA] = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, SYNCS = mbarrier, UCGABAR = cluster
barrier ...), and the tcgen05.* / cp.async.bulk / setmaxnreg lines of the PTX nvcc emits for the two tensor-core translation units."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "sgnerf_b200", "libsgnerf_b200.so")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "UCGABAR", "HMMA", "FFMA", "RED", "ATOMG"]


def sass_counts():
    out = subprocess.run([os.path.join(CUDA, "bin", "cuobjdump"), "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    arch = set(re.findall(r"arch = (sm_\w+)", out))
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + "."):
                    kernels[cur][mn] += 1
                    if mn == "UTCHMMA" and ".2CTA" in op:
                        kernels[cur]["UTCHMMA.2CTA"] += 1
    return arch, kernels


def demangle(names):
    try:
        out = subprocess.run([os.path.join(CUDA, "bin", "cu++filt")] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def ptx_lines(src):
    cmd = [os.path.join(CUDA, "bin", "nvcc"), "-gencode", "arch=compute_100a,code=compute_100a", "-O3", "-std=c++17", "--expt-relaxed-constexpr", "-ptx",
           os.path.join(ROOT, "sgnerf_b200", "csrc", src), "-o", "/dev/stdout"]
    out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    c = collections.Counter()
    for line in out.splitlines():
        m = re.search(r"\b(tcgen05\.[a-z0-9_.:]+|cp\.async\.bulk[a-z0-9_.:]*|setmaxnreg\.[a-z.]+|mbarrier\.[a-z_.:0-9]+|barrier\.cluster\.[a-z.]+)", line)
        if m:
            c[m.group(1)] += 1
    return c


def main():
    arch, kernels = sass_counts()
    names = demangle(list(kernels))
    lines = [f"# SASS of {os.path.relpath(SO, ROOT)} ({', '.join(sorted(arch))}); regenerate with: python tools/sass_evidence.py", ""]
    tot = collections.Counter()
    for k, c in kernels.items():
        hit = {m: c[m] for m in ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "UCGABAR", "RED", "ATOMG") if c[m]}
        tot.update({m: v for m, v in hit.items()})
        if any(m in hit for m in ("UTCHMMA", "LDTM", "UBLKCP", "UTCBAR", "UCGABAR")):
            short = re.sub(r"\(.*", "", names[k])
            lines.append(f"{short}: {c['_total']} instructions; " + ", ".join(f"{m} {v}" for m, v in hit.items()))
    lines += ["", "library totals: " + ", ".join(f"{m} {v}" for m, v in sorted(tot.items())), "",
              "mnemonics: UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM/STTM = tcgen05.ld/st (TMEM), UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA bulk copy),",
              "SYNCS = mbarrier ops, UCGABAR = barrier.cluster, RED/ATOMG = global reductions / atomics (scatter-add of the point-table gradients)", ""]
    for src in ("agg_tc.cu", "agg_fp32.cu"):
        c = ptx_lines(src)
        lines.append(f"# PTX of sgnerf_b200/csrc/{src} (nvcc -ptx, compute_100a): tensor-core / TMA / cluster instructions and their counts")
        lines += [f"  {v:5d}  {k}" for k, v in sorted(c.items())]
        lines.append("")
    out = os.path.join(ROOT, "profiles", "r2_sass_opcodes.txt")
    open(out, "w").write("\n".join(lines))
    print("\n".join(lines[:40]))
    print("->", out)


if __name__ == "__main__":
    main()
