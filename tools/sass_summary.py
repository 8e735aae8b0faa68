"""SASS opcode summary of the shipped library: per kernel, how many tcgen05 / TMEM / TMA-engine / mbarrier / tensor-core
instructions ptxas emitted (the mnemonics /opt/skills/guides/B200_PROFILING.md lists as proof of the Blackwell path).

    python tools/sass_summary.py audio_diffusion_b200/libb200ad.so > profiles/sass_opcodes_<tag>.txt
"""
import collections
import re
import subprocess
import sys

WATCH = ["ACQBULK", "PREEXIT", "UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTCATOM", "UBLKCP", "UTMALDG", "UTMASTG", "UBLKRED",
         "SYNCS", "HMMA", "LDSM", "STSM", "LDGSTS", "FFMA2", "FADD2", "FMUL2", "MUFU.TANH", "MUFU.EX2", "DFMA", "RED", "ATOM", "BAR"]


def main(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    kern, counts, total = None, {}, {}
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(.*", "", kern)
            counts[kern] = collections.Counter()
            total[kern] = 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and kern:
            op = m.group(1)
            total[kern] += 1
            for w in WATCH:
                if op == w or op.startswith(w + ".") or (w.count(".") and op.startswith(w)):
                    counts[kern][w] += 1
    print(f"# {lib}: SASS opcode counts per kernel (cuobjdump -sass, sm_100a)")
    for k in sorted(counts, key=lambda k: -total[k]):
        c = counts[k]
        if not c:
            continue
        print(f"{k}  [{total[k]} instr]")
        print("    " + "  ".join(f"{w}={c[w]}" for w in WATCH if c[w]))
    agg = collections.Counter()
    for c in counts.values():
        agg.update(c)
    print("# whole library: " + "  ".join(f"{w}={agg[w]}" for w in WATCH if agg[w]))


if __name__ == "__main__":
    main(sys.argv[1])
