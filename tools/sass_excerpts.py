#!/usr/bin/env python3
"""Writes profiles/sass/<kernel>.md for the hot kernels of libb200ais.so: the opcode histogram of
the SASS (cuobjdump -sass) and every line that shows how data moves (bulk / tensor copies,
mbarrier traffic, async copies, vector loads and stores, packed FP32), so that "sm_100a-native"
is evidenced by machine code rather than implied by the -arch flag.

    python tools/sass_excerpts.py            # needs only cuobjdump (no GPU)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "gr-ais_b200", "csrc", "build")
OUT = os.path.join(ROOT, "profiles", "sass")

# (object file, regex on the demangled kernel name, output name)
KERNELS = [
    ("corr_fft.o", r"k_corr_fft<8, 120, true>", "k_corr_fft_8_120_tma"),
    ("corr_fft.o", r"k_corr_fft<8, 120, false>", "k_corr_fft_8_120_ldg"),
    ("corr_fft.o", r"k_corr_fft<12, 1120, false>", "k_corr_fft_12_1120"),
    ("agc.o", r"k_mix_agc512_tma<false>", "k_mix_agc512_tma"),
    ("agc.o", r"k_mix_agc512<false>", "k_mix_agc512_ldg"),
    ("freqsync.o", r"k_sqfft_freqest_1024", "k_sqfft_freqest_1024"),
    ("msk.o", r"k_msk<false, 2, 7, true>", "k_msk_kind2"),
    ("msk.o", r"k_msk<false, 0, 1, true>", "k_msk_kind0"),
    ("channelizer.o", r"k_xlat_fir<64, 16", "k_xlat_fir_64_16"),
]
HILITE = re.compile(r"\b(UBLKCP|UTMALDG|UTMASTG|UTMAPF|UTMACMDFLUSH|SYNCS|LDGSTS|LDGDEPBAR|DEPBAR|"
                    r"LDG\.E\.(64|128)|STG\.E\.(64|128)|LDS\.(64|128)|STS\.(64|128)|FFMA2|FADD2|FMUL2|"
                    r"TEX|TLD|ATOMS|REDUX|BAR\.RED|ELECT|FENCE|MEMBAR|CCTL)\b")


def functions(obj):
    txt = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
    out, name, body = {}, None, []
    for ln in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            if name:
                out[name] = body
            name, body = m.group(1), []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            body.append(ln)
    if name:
        out[name] = body
    return out


def demangle(names):
    r = subprocess.run(["c++filt"], input="\n".join(names), stdout=subprocess.PIPE, text=True).stdout
    return dict(zip(names, r.splitlines()))


def main():
    os.makedirs(OUT, exist_ok=True)
    index = []
    cache = {}
    for obj, pat, outname in KERNELS:
        path = os.path.join(OBJ, obj)
        if path not in cache:
            f = functions(path)
            cache[path] = (f, demangle(list(f)))
        funcs, dem = cache[path]
        hit = [m for m, d in dem.items() if re.search(pat, d)]
        if not hit:
            print("not found:", pat, file=sys.stderr)
            continue
        body = funcs[hit[0]]
        ops = collections.Counter()
        lines = []
        for ln in body:
            m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?)\s*;", ln)
            if not m:
                continue
            ins = re.sub(r"^@!?U?P\d+\s+", "", m.group(2))
            ops[ins.split()[0].split(".")[0]] += 1
            if HILITE.search(ins):
                lines.append("%s  %s" % (m.group(1), m.group(2)))
        total = sum(ops.values())
        with open(os.path.join(OUT, outname + ".md"), "w") as fh:
            fh.write("# %s\n\n`%s` in `gr-ais_b200/csrc/build/%s` (cuobjdump -sass, sm_100a), %d instructions.\n\n"
                     % (outname, dem[hit[0]][:160], obj, total))
            fh.write("## Opcode histogram\n\n| opcode | count |\n|---|---|\n")
            for op, n in ops.most_common(28):
                fh.write("| %s | %d |\n" % (op, n))
            fh.write("\n## Data movement, synchronisation and packed-FP32 instructions (%d lines; the first 60 of each kind)\n\n```\n" % len(lines))
            seen = collections.Counter()
            for ln in lines:
                key = HILITE.search(ln).group(1)
                seen[key] += 1
                if seen[key] <= 60 if key in ("FFMA2", "FADD2", "FMUL2", "LDS.64", "LDS.128", "STS.64", "STS.128") and False else seen[key] <= 12:
                    fh.write(ln + "\n")
            fh.write("```\n\n| kind | lines |\n|---|---|\n")
            for k, n in seen.most_common():
                fh.write("| %s | %d |\n" % (k, n))
        index.append((outname, total, dict(ops)))
        print(outname, total, {k: ops[k] for k in ("FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "LDG", "STG", "LDS", "STS", "TEX") if ops[k]})
    return 0


if __name__ == "__main__":
    sys.exit(main())
