#!/usr/bin/env python
"""SASS evidence for profiles/: per-kernel instruction mix of libfq_b200.so (cuobjdump -sass, sm_100a) and the
hot loops of the streaming kernels.  Runs without a GPU.

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "quantization", "mxnet_b200", "libfq_b200.so")
KEYS = ["LDG.E.128", "LDG.E.NA.128", "LDG.E", "STG.E.EF.128", "STG.E.128", "STG.E", "ATOMS", "ATOMG", "RED.E",
        "FMNMX3", "FMNMX", "VIMNMX", "F2I", "F2F", "DADD", "DMUL", "DFMA", "MUFU", "BAR.SYNC", "SHFL", "ACQBULK",
        "LDL", "STL", "CALL", "UTMALDG", "UTCMMA", "UTCIMMA", "UTCHMMA", "TCGEN05", "HMMA", "IMMA", "LDSM"]


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    arch = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m and name:
            funcs[name].append(m.group(2).strip())
    pretty = demangle(list(funcs))
    print("libfq_b200.so -- SASS summary (cuobjdump -sass); ELF images: %s" % ", ".join(sorted(set(re.findall(r"sm_\d+a?", arch)))))
    print("columns: total instructions | selected mnemonics (count)")
    print()
    for f, ins in funcs.items():
        counts = collections.Counter()
        for i in ins:
            op = i.split()[1] if i.startswith("@") else i.split()[0]
            for k in KEYS:
                if op.startswith(k):
                    counts[k] += 1
                    break
        short = re.sub(r"\(.*", "", pretty.get(f, f))
        mix = "  ".join("%s=%d" % (k, counts[k]) for k in KEYS if counts[k])
        print("%-70s %6d | %s" % (short[:70], len(ins), mix))
    print()
    print("Outside the two qconv_igemm kernels no tensor-core (UTC*MMA / HMMA / IMMA) or TMA (UTMALDG) instruction appears: the")
    print("north-star path has no dense contraction and streams with 128-bit LDG/STG; ACQBULK is the griddepcontrol.wait of")
    print("the dependent launches.  QConv2D's integer convolution: UTCIMMA (tcgen05.mma kind::i8; .2CTA = cta_group::2),")
    print("UTMALDG.2D (weights) and UTMALDG.4D.IM2COL (activations) by TMA, UTCBAR(.2CTA.MULTICAST) = tcgen05.commit, LDTM =")
    print("tcgen05.ld from tensor memory:")
    for f, ins in funcs.items():
        if "qconv_igemm" in f:
            c = collections.Counter(re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", i).group(1) for i in ins
                                    if re.match(r"(?:@!?U?P\d+\s+)?(UTMALDG|UTCIMMA|UTCBAR|LDTM|LDGSTS|SYNCS|UCGABAR)", i))
            print("  %s: %s" % (re.sub(r"\(.*", "", pretty.get(f, f)), "  ".join("%s=%d" % kv for kv in sorted(c.items()))))
    # hot loops: from the first 128-bit load to the last store / shared atomic of the unrolled body
    for want, title in (("hist_multi_kernel<true>", "hist_multi_kernel<CHECK=true>: the unrolled tile body (first 120 instructions after the first 128-bit load)"),
                        ("forward_scalar_kernel<true, fq::NoCode, false>", "forward_scalar_kernel<clip, no codes>: tile body"),
                        ("channel_stats_kernel", "channel_stats_kernel: streaming loop")):
        for f, ins in funcs.items():
            if want in pretty.get(f, ""):
                start = next((k for k, i in enumerate(ins) if "LDG.E" in i and ".128" in i), None)
                if start is None:
                    continue
                print()
                print("---- %s ----" % title)
                for i in ins[start:start + 120]:
                    print("    " + i)
                break


if __name__ == "__main__":
    sys.exit(main())
